"""GPU: ellipsoid membership masks are bit-exact vs the reference-generated
fixtures on shared (ctrs, ams) (reference tests/test_ellipsoid.py:106-133)."""
import numpy as np
import pytest

from dynesty_b200 import ops
from oracle import bounding as OB

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['c8', 'c2', 'blob', 'ring'])
def test_membership_golden(golden, name):
    g = golden['multi']
    p = 'me_%s_' % name
    x, ctrs, ams = g[p + 'query'], g[p + 'ctrs'], g[p + 'ams']
    mask, q, d2 = ops.membership(x, ctrs, ams, strict=True, want_d2=True)
    assert np.array_equal(mask, g[p + 'mask'])                 # bit-exact indexing
    assert np.array_equal(q, g[p + 'mask'].sum(axis=1))
    assert np.array_equal(q > 0, g[p + 'contains'])
    ref = OB.MultiEll.__new__(OB.MultiEll)
    ref.ctrs, ref.ams = ctrs, ams
    np.testing.assert_allclose(d2, ref.mahal2(x), rtol=1e-11)


def test_membership_brute_force_10000():
    """Clone of the reference's test_overlap: 10 000 points vs brute force."""
    rng = np.random.default_rng(56432)
    n, K = 2, 10
    ctrs = rng.random((K, n))
    ams = np.empty((K, n, n))
    for k in range(K):
        a = rng.standard_normal((n, n))
        ams[k] = np.linalg.inv(0.01 * (a @ a.T + 0.1 * np.eye(n)))
    x = rng.random((10000, n))
    mask, q = ops.membership(x, ctrs, ams)
    d = x[:, None, :] - ctrs[None]
    d2 = np.einsum('mki,kij,mkj->mk', d, ams, d)
    assert np.abs(d2 - 1).min() > 1e-9          # no point sits on a boundary
    assert np.array_equal(mask, d2 < 1)
    assert np.array_equal(q, (d2 < 1).sum(1))
    # non-strict variant (Ellipsoid.contains, bounding.py:302-305)
    mask2, _ = ops.membership(x, ctrs, ams, strict=False)
    assert np.array_equal(mask2, d2 <= 1)


@pytest.mark.parametrize('n,K,M', [(1, 1, 5), (50, 3, 257), (200, 2, 100), (3, 40, 1000)])
def test_membership_shapes(n, K, M):
    rng = np.random.default_rng(n * 100 + K)
    ctrs = rng.random((K, n))
    ams = np.empty((K, n, n))
    for k in range(K):
        a = rng.standard_normal((n, n + 2))
        ams[k] = np.linalg.inv((a @ a.T) * 0.3 / n)
    x = rng.random((M, n))
    mask, q, d2 = ops.membership(x, ctrs, ams, want_d2=True)
    d = x[:, None, :] - ctrs[None]
    want = np.einsum('mki,kij,mkj->mk', d, ams, d)
    np.testing.assert_allclose(d2, want, rtol=1e-10)
    safe = np.abs(want - 1) > 1e-9
    assert np.array_equal(mask[safe], (want < 1)[safe])
