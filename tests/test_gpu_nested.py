"""GPU: end-to-end nested-sampling runs through the B200 path (bounds + chains on the GPU,
dispatch logic in dynesty_b200.nested) against analytic evidences -- the shape of the
reference's integration tests (tests/test_gau.py:199-228, tests/test_egg.py:29-46).
Tolerance: |logZ - truth| < sig * logzerr + 0.1 with sig = 5 (the reference uses 4-8)."""
import numpy as np
import pytest

from dynesty_b200 import likelihoods as DL, nested

pytestmark = pytest.mark.gpu


def _check(res, truth, sig=5):
    assert abs(res.logz[-1] - truth) < sig * res.logzerr[-1] + 0.1, (res.logz[-1], res.logzerr[-1], truth)


@pytest.mark.parametrize('bound,sample', [('single', 'unif'), ('multi', 'unif'), ('multi', 'rwalk'),
                                          ('single', 'rwalk'), ('multi', 'rslice'), ('single', 'slice')])
def test_c1_gauss3d(bound, sample):
    """BASELINE C1: 3-D correlated Gaussian (tests/test_gau.py:67-102), truth -8.9872."""
    m = DL.gauss_test3d()
    s = nested.NestedSampler(m, nlive=500, bound=bound, sample=sample, seed=56432)
    res = s.run_nested()
    _check(res, m.logz_truth)
    mean, cov = res.posterior_moments()
    assert np.all(np.abs(mean - np.linspace(-1, 1, 3)) < 0.2)
    assert np.all(np.abs(cov - (0.95 + 0.05 * np.eye(3))) < 0.25)
    assert s.nbound > 2


def test_eggbox_2d():
    """tests/test_egg.py:29-46: logZ = 235.856 within 5 sigma (multi-modal: exercises the
    2-means decomposition inside a run)."""
    m = DL.eggbox(2)
    s = nested.NestedSampler(m, nlive=1000, bound='multi', sample='unif', seed=1)
    res = s.run_nested(dlogz=0.01)
    _check(res, 235.856)
    assert max(h[1] for h in res.bound_history) >= 6          # many modes -> many ellipsoids


def test_shells_2d_rslice():
    """demos/Examples -- Gaussian Shells.ipynb: analytic logZ = -1.75 in 2-D."""
    m = DL.shells(2)
    s = nested.NestedSampler(m, nlive=1000, bound='multi', sample='rslice', seed=2)
    res = s.run_nested()
    _check(res, -1.75, sig=5)


def test_c2_reduced_gauss20_rwalk():
    """C2 family at 20-D (full 50-D run is in bench.py): logZ = -20 ln 10."""
    m = DL.gauss_corr(20, 0.4, 5.0)
    s = nested.NestedSampler(m, nlive=1000, bound='multi', sample='rwalk', seed=3)
    res = s.run_nested()
    _check(res, m.logz_truth)


def test_deterministic_given_seed():
    """tests/test_misc.py:328-352: same seed => identical results."""
    m = DL.gauss_test3d()
    a = nested.NestedSampler(m, nlive=200, bound='multi', sample='rwalk', seed=7).run_nested(dlogz=0.5)
    b = nested.NestedSampler(m, nlive=200, bound='multi', sample='rwalk', seed=7).run_nested(dlogz=0.5)
    assert np.array_equal(a.logl, b.logl) and a.ncall == b.ncall
