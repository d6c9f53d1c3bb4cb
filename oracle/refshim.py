"""Import the UNMODIFIED reference (pure Python) from /root/reference/py.

TEST INFRASTRUCTURE.  Only usable in the build container: /root/reference does
not exist on the GPU box, so nothing that runs there may call this.  Used by
oracle/make_golden.py (fixture generation) and by the CPU-only tests that
exercise the drop-in seams against the real dynesty classes (they skip when
the reference is absent).

dynesty/utils.py:21 does ``from . import __version__`` which needs installed
package metadata (py/dynesty/__init__.py:9-15); we provide it with a throwaway
``dynesty-3.0.0.dist-info`` directory in a temp dir on sys.path -- nothing is
written to /root/reference and no reference source is copied.
"""
import os
import sys
import tempfile

REF_PY = '/root/reference/py'


def available():
    return os.path.isdir(os.path.join(REF_PY, 'dynesty'))


def import_reference():
    """Returns the reference ``dynesty`` module (raises ImportError if absent)."""
    if 'dynesty' in sys.modules:
        return sys.modules['dynesty']
    if not available():
        raise ImportError("reference not present at " + REF_PY)
    d = tempfile.mkdtemp(prefix='b2n_refshim_')
    info = os.path.join(d, 'dynesty-3.0.0.dist-info')
    os.makedirs(info)
    with open(os.path.join(info, 'METADATA'), 'w') as f:
        f.write("Metadata-Version: 2.1\nName: dynesty\nVersion: 3.0.0\n")
    sys.path.insert(0, REF_PY)
    sys.path.insert(0, d)
    import dynesty  # noqa
    return dynesty
