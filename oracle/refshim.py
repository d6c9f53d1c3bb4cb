"""Import the UNMODIFIED reference (pure Python) from /root/reference/py.

TEST INFRASTRUCTURE.  /root/reference only exists in the build container.  What CAN
travel to the GPU box is the one offline install of the unmodified reference the
bench contract allows (``pip install --no-deps --target baseline/_ref``, git-ignored,
see DESIGN.md section 8): when /root/reference is absent the shim imports that copy.
Used by oracle/make_golden.py (fixture generation, build container only), by the tests
that plug the B200 bounds/samplers into the real ``dynesty.NestedSampler`` /
``DynamicNestedSampler`` (they skip when neither copy is present) and by bench.py's
CPU arm (``cpu_baseline.kind = "reference"``).  Never on the product path.

dynesty/utils.py:21 does ``from . import __version__`` which needs installed
package metadata (py/dynesty/__init__.py:9-15); we provide it with a throwaway
``dynesty-3.0.0.dist-info`` directory in a temp dir on sys.path -- nothing is
written to /root/reference and no reference source is copied.
"""
import os
import sys
import tempfile

REF_PY = '/root/reference/py'
REF_INSTALLED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')


def source_tree_available():
    return os.path.isdir(os.path.join(REF_PY, 'dynesty'))


def installed_available():
    return os.path.isfile(os.path.join(REF_INSTALLED, 'dynesty', '__init__.py'))


def available():
    return source_tree_available() or installed_available()


def import_reference():
    """Returns the reference ``dynesty`` module (raises ImportError if absent)."""
    if 'dynesty' in sys.modules:
        return sys.modules['dynesty']
    if not available():
        raise ImportError("reference not present at %s or %s" % (REF_PY, REF_INSTALLED))
    if not source_tree_available():
        sys.path.insert(0, REF_INSTALLED)          # carries its own dist-info
        import dynesty  # noqa
        return dynesty
    d = tempfile.mkdtemp(prefix='b2n_refshim_')
    info = os.path.join(d, 'dynesty-3.0.0.dist-info')
    os.makedirs(info)
    with open(os.path.join(info, 'METADATA'), 'w') as f:
        f.write("Metadata-Version: 2.1\nName: dynesty\nVersion: 3.0.0\n")
    sys.path.insert(0, REF_PY)
    sys.path.insert(0, d)
    import dynesty  # noqa
    return dynesty
