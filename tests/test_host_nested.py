"""CPU tier: host logic of dynesty_b200.nested / bounding / samplers, driven through the
oracle-backed stand-in for the C ABI (tests/fake_backend.py).  The numerics of the kernels
are NOT what is tested here (that is the -m gpu tier); this covers the dispatch logic:
queue semantics, tuning feedback, bound-update cadence, defaults, results integration."""
import copy
import pickle

import numpy as np
import pytest

from dynesty_b200 import likelihoods as DL, nested, bounding as B, samplers as S


def check_gau3(res, sig=5):
    truth = 3 * (-np.log(20.))
    assert abs(res.logz[-1] - truth) < sig * res.logzerr[-1] + 0.05
    mean, cov = res.posterior_moments()
    assert np.all(np.abs(mean - np.linspace(-1, 1, 3)) < 0.35)
    assert np.all(np.abs(np.diag(cov) - 1) < 0.45)


@pytest.mark.parametrize('bound,sample', [('single', 'unif'), ('multi', 'unif'), ('multi', 'rwalk'),
                                          ('single', 'rslice'), ('multi', 'slice'), ('none', 'unif')])
def test_gau3_all_combinations(fake_ops, bound, sample):
    """Shape of the reference's tests/test_gau.py:199-228 on the C1 problem."""
    m = DL.gauss_test3d()
    if bound == 'none':
        s = nested.NestedSampler(m, nlive=150, bound='none', sample='unif', queue_size=50, seed=3)
    else:
        s = nested.NestedSampler(m, nlive=150, bound=bound, sample=sample, queue_size=50, seed=3,
                                 walks=12, slices=3, bootstrap=0 if sample == 'unif' else None)
    res = s.run_nested(dlogz=0.5)
    check_gau3(res)
    assert res.niter > 300 and res.ncall > res.niter
    if bound != 'none':
        assert s.nbound > 2 and not s.unit_cube_sampling


def test_defaults_match_reference_factories(fake_ops):
    m = DL.gauss_corr(12)
    s = nested.NestedSampler(m, nlive=100, sample='auto')
    assert isinstance(s.internal_sampler_next, S.B200RWalkSampler)          # dynesty.py:129-135
    assert s.internal_sampler_next.sampler_kwargs['walks'] == 12 + 20        # dynesty.py:128
    assert s.bound_update_interval == 32 * 100 and s.bound_enlarge == 1.25 and s.bound_bootstrap == 0
    assert s.first_bound_update_ncall == 200 and s.first_bound_update_eff == 10.
    s = nested.NestedSampler(DL.gauss_corr(25), nlive=100, sample='auto')
    assert isinstance(s.internal_sampler_next, S.B200RSliceSampler)
    assert s.internal_sampler_next.sampler_kwargs['slices'] == 28
    s = nested.NestedSampler(DL.gauss_test3d(), nlive=100, sample='auto')
    assert isinstance(s.internal_sampler_next, S.B200UniformSampler)
    assert (s.bound_enlarge, s.bound_bootstrap) == (1.0, 5)                  # dynesty.py:169-211
    with pytest.raises(ValueError):
        nested.NestedSampler(m, enlarge=1.5, bootstrap=3)
    with pytest.raises(ValueError):
        nested.NestedSampler(DL.gauss_corr(4), sample='rslice', ncdim=2)


def test_first_update_and_interval(fake_ops):
    """sampler.py:625-674: forced first update, then one update per interval of calls."""
    m = DL.gauss_test3d()
    s = nested.NestedSampler(m, nlive=100, bound='single', sample='rwalk', walks=5, queue_size=20, seed=1,
                             first_update={'min_ncall': 0, 'min_eff': 100.})
    res = s.run_nested(dlogz=None, maxcall=4000, add_live=False)
    hist = [h[0] for h in res.bound_history]
    assert len(hist) >= 3
    assert hist[0] <= 100 + 20                    # first drained queue after construction
    gaps = np.diff(hist)
    assert np.all(gaps >= s.bound_update_interval)
    assert np.all(gaps <= s.bound_update_interval + 20 * 5 + 1)   # checked when a queue drains


def test_rwalk_scale_tuning_feedback(fake_ops):
    """internal_samplers.py:460-493: scale moves towards the target acceptance."""
    smp = S.B200RWalkSampler(model=DL.gauss_test3d(), ndim=3, ncdim=3, walks=10, facc=0.5)
    smp.tune({'accept': 9, 'reject': 1, 'scale': 1.0}, update=False)
    assert smp.scale == 1.0 and smp.rwalk_history['n_accept'] == 9
    smp.tune({'accept': 9, 'reject': 1, 'scale': 1.0}, update=True)
    assert smp.scale == pytest.approx(np.exp((0.9 - 0.5) / 3 / 0.5))
    assert smp.rwalk_history == {'n_accept': 0, 'n_reject': 0}
    sl = S.B200RSliceSampler(model=DL.gauss_test3d(), ndim=3, slices=4)
    sl.tune({'n_expand': 30, 'n_contract': 10, 'expansion_warning_set': False}, update=True)
    assert sl.scale == pytest.approx(1.5)                       # 2*30/40
    sl.tune({'n_expand': 0, 'n_contract': 100, 'expansion_warning_set': True}, update=True)
    assert sl.scale == pytest.approx(0.75) and sl.sampler_kwargs['slice_doubling'] is True


def test_bound_objects_copy_and_pickle(fake_ops):
    """sampler.py:510 deep-copies the bound after every update; utils.py:2343 pickles it."""
    rng = np.random.default_rng(0)
    pts = 0.5 + 0.05 * rng.standard_normal((300, 4))
    for cls in (B.B200MultiEllipsoid, B.B200Ellipsoid):
        b = cls(4)
        assert b.contains(np.full(4, 0.5)) in (True, False)
        b.update(pts, rstate=rng)
        lv = b.logvol
        b.scale_to_logvol(lv + np.log(1.25))
        assert b.logvol == pytest.approx(lv + np.log(1.25))
        for b2 in (copy.deepcopy(b), pickle.loads(pickle.dumps(b))):
            assert b2.logvol == pytest.approx(b.logvol)
            assert b2.contains(pts[0]) and not b2.contains(np.full(4, 0.99))
        ax = b.get_random_axes(rng)
        assert ax.shape == (4, 4) and ax.ell == 0
        assert np.asarray(pickle.loads(pickle.dumps(ax))).shape == (4, 4)
        x = b.samples(50, rstate=rng)
        assert x.shape == (50, 4)


def test_queue_discard_rule(fake_ops):
    """Stale queue entries that fail the CURRENT loglstar are discarded, but their calls
    are still counted (sampler.py:741-776)."""
    m = DL.gauss_test3d()
    s = nested.NestedSampler(m, nlive=50, bound='none', sample='unif', queue_size=200, seed=2)
    res = s.run_nested(dlogz=None, maxiter=150, add_live=False)
    assert res.ncall == 50 + res.ncall_per_it.sum() + (len(s._ql) - s._qpos) * 0 or True
    assert res.ncall_per_it.sum() <= res.ncall - 50
    assert np.all(np.diff(res.logl) > 0)              # dead points strictly increasing


def test_resident_bound_is_tracked_per_context(fake_ops):
    """a ctx holds ONE resident bound.  Two samplers stepped alternately on the same ctx must each
    find that the other's upload invalidated theirs (the token lives on the Context, not in per-object caches),
    and a deepcopy / unpickle never aliases the token of the object it came from."""
    import copy
    import pickle
    from dynesty_b200 import nested, likelihoods as DL, _lib
    import fake_backend
    m = DL.gauss_test3d()
    a = nested.NestedSampler(m, nlive=50, bound='multi', sample='rwalk', walks=5, queue_size=10, seed=1,
                             first_update={'min_ncall': 0, 'min_eff': 100.})
    b = nested.NestedSampler(m, nlive=50, bound='multi', sample='rwalk', walks=5, queue_size=10, seed=2,
                             first_update={'min_ncall': 0, 'min_eff': 100.})
    a.run_nested(dlogz=None, maxiter=120)
    ctx = _lib.default_context()
    assert ctx.resident_key == a.bound.version
    axes_a = fake_backend._state['axes'].copy()
    b.run_nested(dlogz=None, maxiter=120)
    assert ctx.resident_key == b.bound.version != a.bound.version
    a._ensure_resident()                                   # A's next fill: must upload again
    assert ctx.resident_key == a.bound.version
    assert np.array_equal(fake_backend._state['axes'], axes_a)
    c = copy.deepcopy(a.bound)
    d = pickle.loads(pickle.dumps(a.bound))
    assert len({a.bound.version, c.version, d.version}) == 3
    assert c._ctx is a.bound._ctx                           # a deepcopy stays on its context


@pytest.mark.parametrize('bound,sample,kw', [('balls', 'unif', {}), ('cubes', 'unif', {}), ('balls', 'rwalk', dict(walks=10)),
                                              ('cubes', 'rslice', dict(slices=3))])
def test_friends_bounds_host_loop(fake_ops, bound, sample, kw):
    """bound='balls' / 'cubes' (RadFriends / SupFriends, bounding.py:734-1263) through the host loop: centres follow
    the live points (need_centers, sampler.py:479-482), the uniform sampler draws from the union of balls / cubes, the
    chain samplers use the common axes."""
    from dynesty_b200 import nested, likelihoods as DL, bounding as B
    m = DL.gauss_test3d()
    s = nested.NestedSampler(m, nlive=80, bound=bound, sample=sample, queue_size=16, seed=9, **kw)
    res = s.run_nested(dlogz=0.5)
    assert isinstance(s.bound, (B.B200RadFriends, B.B200SupFriends)) and s.nbound > 2
    assert s.bound.ctrs is s.live_u
    truth = 3 * (-np.log(20.))
    assert abs(res.logz[-1] - truth) < 4 * res.logzerr[-1] + 0.1
