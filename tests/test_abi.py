"""CPU tier: the drop-in boundary itself.  libb200nest.so (built for sm_100a by
dynesty_b200/build.py) must load without a GPU, export every function include/b200nest.h
declares, and the ctypes table the host side binds must be exactly that set.  No compute entry
point is called here; the product path must FAIL LOUDLY without a CUDA device (no CPU
fallback)."""
import ctypes as C
import os
import re

import pytest

from dynesty_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'b200nest.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)          # comments mention functions too
    return set(re.findall(r'\b(b2n_[a-z0-9_]+)\s*\(', src))


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_lib.LIBPATH):
        build.build()
    return C.CDLL(_lib.LIBPATH)


def test_header_declares_the_path():
    d = _declared()
    for must in ('b2n_init', 'b2n_bounding_ellipsoid', 'b2n_multi_decompose', 'b2n_membership',
                 'b2n_scale_to_logvol', 'b2n_bootstrap_expand', 'b2n_bound_set', 'b2n_rwalk_batch',
                 'b2n_rslice_batch', 'b2n_slice_batch', 'b2n_unif_batch', 'b2n_peer_export', 'b2n_peer_import'):
        assert must in d
    assert len(d) >= 30


def test_library_exports_every_declared_symbol(lib):
    missing = [name for name in sorted(_declared()) if not hasattr(lib, name)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    assert set(_lib.SYMBOLS) == _declared()


def test_no_torch_types_in_signatures():
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    assert 'torch' not in src.lower() and 'at::' not in src and 'extern "C"' in src
    assert '#include <torch' not in open(HEADER).read()


def test_pure_host_entry_points(lib):
    l = _lib.load()
    assert b'sm_100a' in l.b2n_version()
    assert l.b2n_strerror(_lib.OK) == b'ok'
    for code in range(1, 13):
        assert l.b2n_strerror(code) != b'unknown status'
    al = lambda b: (b + 255) // 256 * 256
    assert l.b2n_peer_window_bytes(2000, 50) == 256 + 2 * (2 * al(2000 * 50 * 8) + al(2000 * 8) + 4 * al(2000 * 4))


def test_product_path_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip('a GPU is present')
    except ImportError:
        pass
    with pytest.raises(_lib.B200Unavailable, match='no CPU fallback'):
        _lib.Context(0)
    from dynesty_b200 import ops
    import numpy as np
    with pytest.raises(_lib.B200Unavailable):
        ops.bounding_ellipsoid(np.random.default_rng(0).random((20, 3)))
