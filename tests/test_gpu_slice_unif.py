"""GPU: rslice / slice / unif chains vs fixtures produced by the UNMODIFIED reference
driven with the scripted Philox stream (oracle/make_golden.py), plus properties.

Tolerance: end points rtol 1e-8 (a slice chain compounds ~100 likelihood evaluations and
device libm differs from glibc in the last bits); the integer bookkeeping (ncall, n_expand,
n_contract) must agree exactly -- it only changes if a likelihood value lands within
round-off of loglstar."""
import numpy as np
import pytest

from dynesty_b200 import ops
from helpers import MODELS, device_model, close, SEED
from oracle import samplers as OS, philox, bounding as OB

pytestmark = pytest.mark.gpu

SLICE_CASES = [(k, n + '_' + t) for t in ('std', 'dbl')
               for k, n in (('rslice', 'egg'), ('rslice', 'shell'), ('rslice', 'g4'),
                            ('slice', 'g4'), ('slice', 'egg'))]


@pytest.mark.parametrize('kind,name', SLICE_CASES)
def test_slice_golden(golden, kind, name):
    g = golden['chains']
    p = '%s_%s_' % (kind, name)
    m = MODELS[name.split('_')[0]]
    dm = device_model(m)
    ops.bound_set(g[p + 'axes'])
    fn = ops.rslice_batch if kind == 'rslice' else ops.slice_batch
    o = fn(dm.model_id(), g[p + 'u0'], float(g[p + 'loglstar']), float(g[p + 'scale']),
           int(g[p + 'slices']), SEED, chain0=int(g[p + 'chain0']), doubling=bool(g[p + 'doubling']))
    assert np.array_equal(o['ncall'], g[p + 'ncall'])
    assert np.array_equal(o['n_expand'], g[p + 'n_expand'])
    assert np.array_equal(o['n_contract'], g[p + 'n_contract'])
    close(o['u'], g[p + 'u'], rtol=1e-8)
    close(o['v'], g[p + 'v'], rtol=1e-8)
    np.testing.assert_allclose(o['logl'], g[p + 'logl'], rtol=1e-8, atol=1e-8)
    assert np.all(o['flags'] == 0)


def test_rslice_c3_properties():
    """BASELINE C3 shape (25-D eggbox, rslice, slices=28) on a 1000-chain batch."""
    m = MODELS['egg'].__class__(25, MODELS['egg'].prior_kind, MODELS['egg'].like_kind,
                                **MODELS['egg'].p)
    dm = device_model(m)
    rng = np.random.default_rng(9)
    pts = rng.random((4000, 25))
    logl = m.loglike(m.prior_transform(pts))
    loglstar = float(np.quantile(logl, 0.5))
    e = OB.bounding_ellipsoid(pts)
    u0 = pts[logl > loglstar][:1000]
    ops.bound_set(e.axes)
    o = ops.rslice_batch(dm.model_id(), u0, loglstar, 1.0, 28, SEED)
    assert np.all(o['logl'] > loglstar)
    assert np.all((o['u'] > 0) & (o['u'] < 1))
    np.testing.assert_allclose(o['logl'], m.loglike(m.prior_transform(o['u'])), rtol=1e-9)
    assert np.all(o['n_contract'] >= 28)
    assert np.all(o['ncall'] == 2 * 28 + o['n_expand'] + o['n_contract'])
    for i in (0, 17, 999):
        r = OS.rslice_chain(u0[i], loglstar, e.axes, 1.0, m, philox.ChainStream(SEED, i), 28)
        assert r['ncall'] == o['ncall'][i]
        close(o['u'][i], r['u'], rtol=1e-8)


def test_slice_requires_full_ncdim():
    dm = device_model(MODELS['g6'])
    ops.bound_set(np.eye(4))
    with pytest.raises(ValueError):
        ops.rslice_batch(dm.model_id(), np.full((2, 6), 0.5), -1e9, 1.0, 3, 1)


@pytest.mark.parametrize('name', ['g3', 'g3nc', 'shell2'])
def test_unif_golden(golden, name):
    g = golden['chains']
    p = 'unif_%s_' % name
    m = MODELS[name]
    dm = device_model(m)
    ops.bound_set(g[p + 'axes'], g[p + 'ctrs'], g[p + 'ams'], g[p + 'logvols'])
    nchain = len(g[p + 'logl'])
    o = ops.unif_batch(dm.model_id(), nchain, int(g[p + 'ndim']), float(g[p + 'loglstar']), SEED,
                       chain0=int(g[p + 'chain0']), ncdim=g[p + 'ctrs'].shape[1])
    assert np.array_equal(o['ncall'], g[p + 'ncall'])
    close(o['u'], g[p + 'u'], rtol=1e-9)
    close(o['v'], g[p + 'v'], rtol=1e-9)
    np.testing.assert_allclose(o['logl'], g[p + 'logl'], rtol=1e-9, atol=1e-9)
    assert np.all(o['nprop'] >= o['ncall'])


def test_unif_uniformity_two_ellipsoids():
    """tests/test_ellipsoid.py:14-59 idea: samples from two overlapping unit-ish discs are
    uniform over the union (the 1/q rule), checked through area fractions."""
    m = MODELS['shell2']
    dm = device_model(m)
    ctrs = np.array([[0.4, 0.5], [0.6, 0.5]])
    r = 0.15
    ams = np.array([np.eye(2) / r**2] * 2)
    axes = np.array([np.eye(2) * r] * 2)
    lv = np.log(np.pi * r * r) * np.ones(2)
    ops.bound_set(axes, ctrs, ams, lv)
    o = ops.unif_batch(dm.model_id(), 20000, 2, -1e300, 4242)
    x = o['u']
    in0 = ((x - ctrs[0])**2).sum(1) < r * r
    in1 = ((x - ctrs[1])**2).sum(1) < r * r
    assert np.all(in0 | in1)
    # lens area for two radius-r discs at distance d
    d = 0.2
    lens = 2 * r * r * np.arccos(d / (2 * r)) - 0.5 * d * np.sqrt(4 * r * r - d * d)
    union = 2 * np.pi * r * r - lens
    frac = (in0 & in1).mean()
    assert abs(frac - lens / union) < 4 * np.sqrt(frac * (1 - frac) / len(x))
    assert np.all(o['ncall'] == 1)


def _two_sphere_vol(d, r1, r2, ndim):
    """Volume of the union of two n-balls (numerical 1-D integral over the axis joining them)."""
    from scipy.special import gammaln
    def ball(n, r):
        return np.exp(n / 2. * np.log(np.pi) - gammaln(n / 2. + 1) + n * np.log(r))
    xs = np.linspace(-r1, max(r1, d + r2), 400001)
    a = np.clip(r1**2 - xs**2, 0, None)
    b = np.clip(r2**2 - (xs - d)**2, 0, None)
    rad = np.sqrt(np.maximum(a, b))
    return np.trapezoid(ball(ndim - 1, 1.0) * rad**(ndim - 1), xs)


def test_mc_logvol_two_spheres():
    """tests/test_ellipsoid.py:174-194: MC volume of a two-ball union within 1e-2 (n=10, 1e4 draws
    in the reference; 2e5 here since the draws are one launch)."""
    from dynesty_b200.bounding import B200MultiEllipsoid
    import math
    from scipy.special import gammaln
    ndim, r1, r2 = 10, 1.0, 0.5
    pref = ndim / 2. * math.log(math.pi) - gammaln(ndim / 2. + 1)
    for D in (0.0, 0.6, 1.2, 2.0):
        b = B200MultiEllipsoid(ndim)
        c2 = np.zeros(ndim)
        c2[0] = D
        b.nells = 2
        b.ctrs = np.array([np.zeros(ndim), c2])
        b.covs = np.array([np.eye(ndim) * r1**2, np.eye(ndim) * r2**2])
        b.ams = np.array([np.eye(ndim) / r1**2, np.eye(ndim) / r2**2])
        b.axes_all = np.array([np.eye(ndim) * r1, np.eye(ndim) * r2])
        b.axlens_all = np.array([np.full(ndim, r1), np.full(ndim, r2)])
        b.logvol_ells = np.array([pref + ndim * math.log(r1), pref + ndim * math.log(r2)])
        b._refresh_logvol()
        lv, overlap = b.monte_carlo_logvol(200000, rstate=np.random.default_rng(int(D * 10)))
        assert abs(lv - math.log(_two_sphere_vol(D, r1, r2, ndim))) < 1e-2
        assert 0 <= overlap <= 1


def test_cube_overlap_half():
    """tests/test_ellipsoid.py:92-103: a ball centred on a cube face overlaps it by one half."""
    from dynesty_b200.bounding import B200Ellipsoid
    ndim = 10
    b = B200Ellipsoid(ndim)
    cen = np.full(ndim, 0.5)
    cen[0] = 0
    m = b._m
    m.ctrs, m.covs, m.ams = cen[None], (np.eye(ndim) * 0.25)[None], (np.eye(ndim) * 4.)[None]
    m.axes_all, m.axlens_all = (np.eye(ndim) * 0.5)[None], np.full((1, ndim), 0.5)
    m._refresh_logvol()
    frac = b.unitcube_overlap(100000, rstate=np.random.default_rng(3))
    assert abs(frac - 0.5) < 5 * np.sqrt(0.25 / 100000)
