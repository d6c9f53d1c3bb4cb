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
