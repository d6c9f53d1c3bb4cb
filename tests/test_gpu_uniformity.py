"""GPU: the reference's sampler-uniformity harness (tests/test_sampling.py:61-157, SURVEY 8c fixture iv).

The reference runs ONE chain of 100 000 sampler calls inside a hard-edged 2-D region (`diamond_logl` /
`checker_logl`, loglstar = 0, axes = I) and checks that the visited points are uniform over the region
through the marginal histograms (`pdf_test`: 100 bins, 6 sigma / 1 % margin).  Here the same regions are the
registry likelihood B2N_LIKE_REGION2D; 4096 chains run in parallel, each sampler call of every chain is one
kernel launch started from the previous end points, and the pooled end points of the later launches
(>= 1e5 points) go through the reference's `pdf_test`, restated below with the analytic bin integrals.
Covers rwalk (warp-per-chain kernel at n = 2, lock-step DMMA kernel at n = 16 with 14 free dimensions),
rslice (stepping-out and doubling) and slice."""
import numpy as np
import pytest

from dynesty_b200 import ops, likelihoods as DL

pytestmark = pytest.mark.gpu

NCHAIN = 4096


def pdf_test(cdf, curx, nbins=100, thresh=6):
    """tests/test_sampling.py:26-41 with the bin integrals from the analytic cdf instead of quad()."""
    hh, loc = np.histogram(curx, range=[0, 1], bins=nbins)
    width = loc[1] - loc[0]
    norm = width * len(curx)
    pdf = hh / norm
    model_pdf = np.diff(cdf(loc)) / width
    frac = 0.01 * model_pdf.max()
    epdf = (model_pdf * norm)**.5 / norm
    epdf1 = hh**.5 / norm
    margin = np.maximum(thresh * np.maximum(epdf, epdf1), frac)
    assert (np.abs(model_pdf - pdf) / margin).max() < 1


def diamond_cdf(x):
    """Integral of (1 - 2 sqrt(|x-.5| - (x-.5)^2)) / (1 - pi/4) (tests/test_sampling.py:46-48): with
    t = |x - .5| the integrand is 1 - 2 sqrt(t - t^2), the circle (t - .5)^2 + y^2 = .25."""
    x = np.asarray(x, dtype=float)
    t = np.abs(x - 0.5)
    s = t - 0.5                                           # in [-.5, 0]
    # int_0^t 2 sqrt(.25 - (t'-.5)^2) dt' = [ s sqrt(.25 - s^2) + .25 asin(2 s) ] from -.5 to s
    prim = s * np.sqrt(np.maximum(0.25 - s * s, 0.0)) + 0.25 * np.arcsin(np.clip(2 * s, -1, 1)) + 0.25 * np.pi / 2
    half = (t - prim) / (1 - np.pi / 4)                   # mass between the centre and x
    return 0.5 + np.sign(x - 0.5) * half


def uniform_cdf(x):
    return np.asarray(x, dtype=float)


def run_chains(model, sampler, nrounds, burn, scale, ndim=2, walks=10, slices=10, doubling=False, seed=11):
    mid = model.model_id()
    ops.bound_set(np.eye(ndim)[None])
    u = np.full((NCHAIN, ndim), 0.5)
    out = []
    for r in range(nrounds):
        c0 = r * NCHAIN
        if sampler == 'rwalk':
            o = ops.rwalk_batch(mid, u, 0.0, scale, walks, seed, chain0=c0)
        elif sampler == 'rslice':
            o = ops.rslice_batch(mid, u, 0.0, scale, slices, seed, chain0=c0, doubling=doubling)
        else:
            o = ops.slice_batch(mid, u, 0.0, scale, slices, seed, chain0=c0, doubling=doubling)
        u = o['u']
        assert np.all(o['logl'] > 0.0)                    # never leaves the region
        if r >= burn:
            out.append(u.copy())
    return np.concatenate(out)


def check_diamond(X):
    for i in range(2):
        pdf_test(diamond_cdf, X[:, i])
    for i in range(2, X.shape[1]):                        # free dimensions stay uniform
        pdf_test(uniform_cdf, X[:, i])


def test_diamond_cdf_is_the_reference_density():
    xs = np.linspace(0.003, 0.997, 400)
    pdf = (1 - 2 * np.sqrt(np.abs(xs - 0.5) - (xs - 0.5)**2)) / (1 - np.pi / 4)
    num = (diamond_cdf(xs + 1e-6) - diamond_cdf(xs - 1e-6)) / 2e-6
    assert np.allclose(num, pdf, atol=1e-5)
    assert abs(diamond_cdf(0.0)) < 1e-12 and abs(diamond_cdf(1.0) - 1) < 1e-12


def test_diamond_rwalk():
    """test_sampling.py:105-113 (scale .3, walks 10)."""
    check_diamond(run_chains(DL.region2d('diamond'), 'rwalk', 40, 12, 0.3, walks=10))


def test_diamond_rwalk_lockstep_kernel():
    """Same region on the first two of 16 dimensions: the lock-step DMMA kernel (16 <= n <= 64)."""
    check_diamond(run_chains(DL.region2d('diamond', 16), 'rwalk', 60, 24, 0.6, ndim=16, walks=30))


def test_diamond_rslice():
    """test_sampling.py:116-124 (scale .1, slices 10)."""
    check_diamond(run_chains(DL.region2d('diamond'), 'rslice', 36, 8, 0.1, slices=10))


def test_diamond_rslice_double():
    """test_sampling.py:127-136 (scale .001, doubling; the reference thins its single chain by 10 -- here every
    launch is 5 slices and only every second launch is kept)."""
    X = run_chains(DL.region2d('diamond'), 'rslice', 70, 20, 0.001, slices=5, doubling=True)
    check_diamond(X.reshape(-1, NCHAIN, 2)[::2].reshape(-1, 2))


def test_diamond_slice():
    """test_sampling.py:139-147 ('checkerboard_rslice' in name only: diamond, principal-axes slice, scale .3)."""
    check_diamond(run_chains(DL.region2d('diamond'), 'slice', 36, 8, 0.3, slices=1))


def test_checkerboard_rslice_double():
    """test_sampling.py:150-159: 16 x 16 x 2 cells, doubling slices at scale .001 must reach all of them."""
    X = run_chains(DL.region2d('checkerboard'), 'rslice', 90, 40, 0.001, slices=5, doubling=True)
    X = X.reshape(-1, NCHAIN, 2)[::2].reshape(-1, 2)
    for i in range(2):
        pdf_test(uniform_cdf, X[:, i])
