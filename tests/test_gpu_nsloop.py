"""GPU tier: the device-resident nested-sampling rounds (csrc/b2n_ns.cu, C ABI b2n_ns_*) against
their oracle (oracle/nsloop.py) on the same seeded inputs, their stop flags, and whole runs
(``run_nested(loop='device')``) against analytic evidences.

Tolerances: dead-point order / counts / call totals exact; dead logl exact (they are copies of the
inputs' values); ln X per dead point 1e-14; running logZ 1e-10; live set after the rounds rtol 1e-8
(chains replay the oracle's to ~1e-9, tests/test_gpu_rwalk.py); tuned scale 1e-10."""
import math

import numpy as np
import pytest

from oracle import nsloop, likelihoods as OL, bounding as OB
from dynesty_b200 import ops, likelihoods as DL, nested
from helpers import close

pytestmark = pytest.mark.gpu


def _bound(points_list, strict=True, enlarge=1.25):
    ells = []
    for p in points_list:
        e = OB.bounding_ellipsoid(p)
        e.scale_to_logvol(e.logvol + math.log(enlarge))
        ells.append(e)
    return dict(ctrs=np.array([e.ctr for e in ells]), ams=np.array([e.am for e in ells]),
                axes=np.array([e.axes for e in ells]), logvols=np.array([e.logvol for e in ells]), strict=strict)


def _models(kind, n):
    if kind == 'gauss':
        return DL.gauss_corr(n, 0.4, 5.0), OL.gauss_corr(n, 0.4, 5.0)
    return DL.shells(n), OL.shells(n)


def _live(om, n, N, rng, two=False):
    if two:        # two clusters around the two shell centres
        half = N // 2
        c1, c2 = 0.5 + np.zeros(n), 0.5 + np.zeros(n)
        c1[0], c2[0] = 0.5 - 3.5 / 12, 0.5 + 3.5 / 12
        u = np.concatenate([c1 + 0.03 * rng.standard_normal((half, n)), c2 + 0.03 * rng.standard_normal((N - half, n))])
        groups = [u[:half], u[half:]]
    else:
        u = 0.5 + 0.05 * rng.standard_normal((N, n))
        groups = [u]
    v = om.prior_transform(u)
    l = np.array([float(om.loglike(x)) for x in v])
    return u, v, l, groups


CASES = [
    # kind, n, N, K, sampler, steps, two-ellipsoid bound, rounds
    ('gauss', 6, 64, 16, 'rwalk', 10, False, 3),       # warp-per-chain rwalk kernel
    ('gauss', 20, 96, 24, 'rwalk', 12, False, 3),      # lock-step DMMA rwalk kernel
    ('gauss', 5, 64, 8, 'rslice', 4, False, 3),
    ('gauss', 4, 48, 12, 'slice', 1, False, 2),
    ('shells', 4, 80, 20, 'rwalk', 8, True, 3),        # 2 ellipsoids: volume-weighted picks + grouped worklist
    ('shells', 4, 80, 10, 'rslice', 3, True, 2),
    ('gauss', 4, 64, 16, 'unif', 1, False, 3),          # uniform sampler: chains draw from the bound themselves
    ('shells', 3, 80, 16, 'unif', 1, True, 2),          # ... from a two-ellipsoid bound (1/q acceptance)
    ('gauss', 6, 64, 16, 'rwalk', 10, False, 3, dict(ncdim=4)),                  # bound on the first 4 dims only
    ('gauss', 6, 64, 16, 'rwalk', 10, False, 3, dict(dimflags=[1, 2, 0, 0, 1, 0])),   # periodic / reflective dims
]
CASES = [c if len(c) == 9 else c + ({},) for c in CASES]



def _abort_at(k_stop):
    """on_checkpoint callback that kills the run after its k_stop-th checkpoint (the reference's
    tests/test_resume.py kills the process instead)."""
    def cb(k):
        if k >= k_stop:
            raise KeyboardInterrupt('test: run aborted after checkpoint %d' % k)
    return cb


@pytest.mark.parametrize('kind,n,N,K,sampler,steps,two,rounds,extra', CASES)
def test_rounds_match_oracle(kind, n, N, K, sampler, steps, two, rounds, extra):
    dm, om = _models(kind, n)
    rng = np.random.default_rng(100 + n + K)
    u, v, l, groups = _live(om, n, N, rng, two)
    nc = extra.get('ncdim', n)
    flags = extra.get('dimflags')
    b = _bound([g[:, :nc] for g in groups])
    seed, chain0, scale0 = 56432, 1000, 0.7
    o = nsloop.BatchNS(om, u, v, l, K, sampler, steps, seed, chain0=chain0, scale=scale0, logvol=-2.5, logz=-40.0,
                       loglstar=float(l.min()) - 0.5, ncall=500, bound=b, dlogz=1e-6, dimflags=flags)
    ops.ns_create(dm.model_id(), N, n, K, ('rwalk', 'rslice', 'slice', 'unif').index(sampler), steps, seed, chain0=chain0,
                  ncdim=nc, dlogz=1e-6, dead_capacity=rounds * K + 5,
                  dimflags=None if flags is None else np.array(flags, dtype=np.uint8))
    try:
        ops.ns_set_state(u, v, l, -2.5, -40.0, float(l.min()) - 0.5, 500, scale0)
        for r in range(rounds):
            # the same bound on both sides, rebuilt from the oracle's live set before every round
            # (chains leave a static bound; a start outside it raises need_bound = 2 on both sides)
            lu = o.live_u[:, :nc]
            b = _bound([lu[lu[:, 0] < 0.5], lu[lu[:, 0] >= 0.5]] if two else [lu])
            o.bound = b
            ops.bound_set(b['axes'], b['ctrs'], b['ams'], b['logvols'])
            assert o.step(), (o.done, o.need_bound)
            st = ops.ns_run(1, 0)
        assert (st['done'], st['need_bound'], st['error']) == (0, 0, 0)
        assert st['rounds'] == rounds and st['it'] == rounds * K
        assert st['ncall'] == o.ncall
        du, dv, dl, dlv, dnc = ops.ns_get_dead(0, st['it'], n)
        ou, ov, ol, olv, onc = o.dead_arrays()
        assert np.array_equal(dl[:K], ol[:K])                       # first round: copies of the inputs, same order
        assert np.allclose(dl, ol, rtol=1e-9, atol=0) and np.allclose(du, ou, rtol=1e-8, atol=1e-12)
        assert np.allclose(dlv, olv, rtol=0, atol=1e-13)
        assert np.array_equal(dnc, onc)
        # live sets as SETS (sorted by logl): a chain that never moved returns a clone of its start row whose
        # logl the device recomputes (last-bit different from the host value of the original), so the order of
        # such a pair -- hence which freed slot receives which chain -- may differ between the two sides
        lu, lv_, ll = ops.ns_get_live(N, n)
        pd, po = np.argsort(ll, kind='stable'), np.argsort(o.live_logl, kind='stable')
        assert np.allclose(ll[pd], o.live_logl[po], rtol=1e-8, atol=1e-10)
        assert np.allclose(lu[pd], o.live_u[po], rtol=1e-8, atol=1e-12)
        assert np.allclose(lv_[pd], o.live_v[po], rtol=1e-8, atol=1e-11)
        assert st['logz'] == pytest.approx(o.logz, rel=1e-10)
        assert st['logvol'] == pytest.approx(o.logvol, abs=1e-13)
        assert st['scale'] == pytest.approx(o.scale, rel=1e-10)
        assert st['loglstar'] == pytest.approx(o.loglstar, rel=1e-9)
    finally:
        ops.ns_destroy()


def test_stop_flags_and_dead_capacity():
    dm, om = _models('gauss', 6)
    rng = np.random.default_rng(3)
    N, K, n = 64, 16, 6
    u, v, l, groups = _live(om, n, N, rng)
    b = _bound(groups, enlarge=3.0**n)       # axes x3: children (<= 1.5 axes from their parents) stay inside
    ops.bound_set(b['axes'], b['ctrs'], b['ams'], b['logvols'])
    ops.ns_create(dm.model_id(), N, n, K, 0, 5, 1, update_interval=100, dlogz=1e-9, dead_capacity=2 * K)
    try:
        ops.ns_set_state(u, v, l, 0.0, -1e300, -1e300, 0, 0.5)
        st = ops.ns_run(8, 0)                                    # 16 x 5 = 80 calls per round, interval 100
        assert st['rounds'] == 2 and st['need_bound'] == 1 and st['ncall'] == 160
        st = ops.ns_run(4, 0)                                    # flag still up: nothing runs
        assert st['rounds'] == 2
        lu, _, _ = ops.ns_get_live(N, n)
        b2 = _bound([lu])
        ops.bound_set(b2['axes'], b2['ctrs'], b2['ams'], b2['logvols'])
        ops.ns_bound_updated()
        st = ops.ns_run(1, 0)                                    # dead buffer (2K rows) is full
        assert st['need_bound'] == 3 and st['rounds'] == 2
        ops.ns_reserve_dead(10 * K)
        st = ops.ns_run(1, 0)
        assert st['rounds'] == 3 and st['it'] == 3 * K
        _, _, dl, _, _ = ops.ns_get_dead(0, 3 * K, n)
        assert np.all(np.diff(dl) >= 0)                          # rows survived the reallocation, still ascending
        # a bound that does not contain the live points -> forced update request
        ops.bound_set(b2['axes'], b2['ctrs'] + 5.0, b2['ams'], b2['logvols'])
        ops.ns_bound_updated()
        st = ops.ns_run(2, 0)
        assert st['need_bound'] == 2 and st['rounds'] == 3
    finally:
        ops.ns_destroy()
    # termination: dlogz huge -> done at the first propose
    ops.bound_set(b['axes'], b['ctrs'], b['ams'], b['logvols'])
    ops.ns_create(dm.model_id(), N, n, K, 0, 5, 1, dlogz=1e9)
    try:
        ops.ns_set_state(u, v, l, 0.0, -10.0, -1e300, 0, 0.5)
        st = ops.ns_run(3, 0)
        assert st['done'] == 1 and st['rounds'] == 0 and st['it'] == 0
    finally:
        ops.ns_destroy()


@pytest.mark.parametrize('ndim,nlive,sample,batch', [(20, 1000, 'rwalk', 100), (20, 600, 'rslice', 120),
                                                     (6, 400, 'slice', 40), (3, 500, 'unif', 50)])
def test_device_loop_logz(ndim, nlive, sample, batch):
    """Whole runs with the rounds on the device: logZ against the analytic evidence of the C2 family."""
    m = DL.gauss_corr(ndim, 0.4, 5.0)
    s = nested.NestedSampler(m, nlive=nlive, bound='multi', sample=sample, seed=5, queue_size=max(32, nlive // 10))
    res = s.run_nested(loop='device', batch=batch)
    assert abs(res.logz[-1] - m.logz_truth) < 3.5 * res.logzerr[-1] + 0.1, (res.logz[-1], res.logzerr[-1], m.logz_truth)
    assert s.device_rounds > 10 and s.nbound > 3
    assert np.all(np.diff(res.logl) >= 0) and np.all(np.diff(res.logvol) < 0)
    mean, cov = res.posterior_moments()
    assert np.all(np.abs(mean) < 0.5)
    assert np.all(np.abs(np.diag(cov) - 1.0) < 0.5)


def test_device_loop_shells_two_ellipsoids():
    """C5 likelihood (Gaussian shells, two modes): multi-ellipsoid bound with K > 1 inside the device loop."""
    m = DL.shells(2)
    s = nested.NestedSampler(m, nlive=600, bound='multi', sample='rslice', seed=9, queue_size=60)
    res = s.run_nested(loop='device', batch=60)
    assert abs(res.logz[-1] - m.logz_truth) < 3.5 * res.logzerr[-1] + 0.1
    assert max(h[1] for h in res.bound_history) >= 2


def test_rounds_c2_size_properties():
    """BASELINE C2 size (50-D, nlive 2000, batch 50, walks 70) -- too big for the per-chain Python oracle, so
    size-independent properties of the rounds: the batch lowest points die in ascending order, every live
    point beats the last threshold, (u, v, logl) stay consistent with the model, ln X telescopes, the call
    count is rounds x batch x walks, the sorted order the device maintains by merging equals a full sort."""
    m = DL.gauss_corr(50, 0.4, 5.0)
    om = OL.gauss_corr(50, 0.4, 5.0)
    rng = np.random.default_rng(8)
    N, n, K, walks, R = 2000, 50, 50, 70, 12
    Cm = np.full((n, n), 0.4)
    np.fill_diagonal(Cm, 1.0)
    u = 0.5 + 0.03 * rng.standard_normal((N, n)) @ np.linalg.cholesky(Cm).T
    v, l = m.evaluate(u)
    b = _bound([u], enlarge=3.0**n)
    ops.bound_set(b['axes'], b['ctrs'], b['ams'], b['logvols'])
    ops.ns_create(m.model_id(), N, n, K, 0, walks, 3, dlogz=1e-9, dead_capacity=R * K)
    try:
        ops.ns_set_state(u, v, l, 0.0, -1e300, -1e300, 0, 0.05)
        st = ops.ns_run(R, 0)
        assert (st['done'], st['need_bound'], st['rounds'], st['it']) == (0, 0, R, R * K)
        assert st['ncall'] == R * K * walks
        du, dv, dl, dlv, dnc = ops.ns_get_dead(0, R * K, n)
        assert np.all(np.diff(dl) >= 0) and np.all(dnc == walks)
        assert np.array_equal(dl[:K], np.sort(l)[:K])
        lu, lv_, ll = ops.ns_get_live(N, n)
        assert ll.min() > dl[-1] and st['loglstar'] == dl[-1] and st['lmax'] == ll.max()
        close(lv_, om.prior_transform(lu), rtol=1e-13)          # (atol scaled by max |v|: v passes through 0)
        close(ll, om.loglike(lv_), rtol=1e-10)
        # ln X: after r full rounds ln X = r ln((N-K+1)/(N+1)); inside a round ln((N-j)/(N+1)) on top
        j = np.arange(R * K)
        expect = (j // K) * math.log((N - K + 1) / (N + 1.0)) + np.log((N - j % K) / (N + 1.0))
        np.testing.assert_allclose(dlv, expect, rtol=0, atol=1e-12)
        assert st['logvol'] == pytest.approx(R * math.log((N - K + 1) / (N + 1.0)), abs=1e-12)
        # running logZ == post-hoc trapezoid integral of the dead points
        _, logz, _, _ = nested._integrate(dl, dlv)
        assert st['logz'] == pytest.approx(logz[-1], rel=1e-11)
    finally:
        ops.ns_destroy()


@pytest.mark.parametrize('sample,kw', [('rwalk', dict(walks=30)), ('rslice', dict(slices=8))])
def test_checkpoint_resume_is_bit_identical(tmp_path, sample, kw):
    """tests/test_resume.py of the reference: kill after a checkpoint, restore, resume -> the same results as the
    uninterrupted run, bit for bit (device state restored through b2n_ns_set_state + b2n_ns_set_counters)."""
    m = DL.gauss_corr(10, 0.4, 5.0)
    mk = lambda: nested.NestedSampler(m, nlive=400, bound='multi', sample=sample, queue_size=40, seed=11, **kw)
    ref = mk().run_nested(loop='device', batch=20)
    f = str(tmp_path / 'ckpt.pkl')
    s = mk()
    with pytest.raises(KeyboardInterrupt):
        s.run_nested(loop='device', batch=20, checkpoint_file=f, checkpoint_every=0., on_checkpoint=_abort_at(4))
    del s
    r = nested.NestedSampler.restore(f)
    assert r._dev_snap['rounds'] > 0 and len(r._dev_snap['dead'][2]) > 0
    res = r.run_nested(resume=True)
    assert res.niter == ref.niter and res.ncall == ref.ncall
    assert np.array_equal(res.logl, ref.logl) and np.array_equal(res.logvol, ref.logvol)
    assert np.array_equal(res.samples_u, ref.samples_u) and np.array_equal(res.samples, ref.samples)
    assert res.logz[-1] == ref.logz[-1] and res.logzerr[-1] == ref.logzerr[-1]
    assert abs(res.logz[-1] - m.logz_truth) < 4 * res.logzerr[-1] + 0.1


# ---- round 2: the phase before the first bound, the device-side bound update, ties at the threshold -----------
def test_unitcube_batch_matches_oracle():
    """b2n_unitcube_batch == UnitCubeSampler.sample (internal_samplers.py:420-441) on the B2N stream: same draws,
    same call counts (the oracle chain is pinned to the reference in tests/golden/chains.npz: uc_*)."""
    from oracle import samplers as OS, philox
    dm, om = _models('gauss', 6)
    thr = -30.0
    o = ops.unitcube_batch(dm.model_id(), 48, 6, thr, 77, chain0=9)
    for i in (0, 5, 47):
        c = OS.unitcube_chain(thr, om, philox.ChainStream(77, 9 + i), 6)
        assert c['ncall'] == o['ncall'][i]
        close(o['u'][i], c['u'], rtol=1e-15)
        close(o['v'][i], c['v'], rtol=1e-12)
        assert o['logl'][i] == pytest.approx(c['logl'], rel=1e-11)
    assert np.all(o['logl'] > thr) and o['ncall'].max() > 1


def test_unitcube_golden(golden):
    """The reference's own UnitCubeSampler.sample replayed on the Philox stream (oracle/make_golden.py)."""
    g = golden['chains']
    dm = DL.gauss_test3d()
    thr = float(g['uc_loglstar'])
    o = ops.unitcube_batch(dm.model_id(), len(g['uc_u']), 3, thr, int(g['uc_seed']), chain0=int(g['uc_chain0']))
    assert np.array_equal(o['ncall'], g['uc_ncall'])
    close(o['u'], g['uc_u'], rtol=1e-15)
    close(o['v'], g['uc_v'], rtol=1e-12)
    np.testing.assert_allclose(o['logl'], g['uc_logl'], rtol=1e-11)


def test_unitcube_phase_rounds_match_oracle():
    """Rounds before the first bound (unit_cube_phase): prior draws at the round's threshold, dead records and
    evidence as in the bounded rounds, need_bound = 4 exactly when the reference's first-update test fires
    (ncall >= min_ncall and eff < min_eff, sampler.py:640-647)."""
    dm, om = _models('gauss', 4)
    rng = np.random.default_rng(8)
    N, K, n = 60, 6, 4
    u = rng.random((N, n))
    v = om.prior_transform(u)
    l = np.array([float(om.loglike(x)) for x in v])
    kw = dict(unit_cube_phase=True, first_min_ncall=2 * N, first_min_eff=25.0, it0=1)
    o = nsloop.BatchNS(om, u, v, l, K, 'rwalk', 9, 5, chain0=0, ncall=N, dlogz=1e-6, **kw)
    ops.ns_create(dm.model_id(), N, n, K, 0, 9, 5, chain0=0, dlogz=1e-6, **kw)
    try:
        ops.ns_set_state(u, v, l, 0.0, -1e300, -1e300, N, 1.0)
        nr = 0
        while o.step():
            nr += 1
        assert o.need_bound == 4 and nr >= 3
        st = ops.ns_run(nr + 5, 0)                       # the extra rounds must be no-ops
        assert (st['rounds'], st['need_bound'], st['done'], st['it'], st['ncall']) == (nr, 4, 0, nr * K, o.ncall)
        assert st['logz'] == pytest.approx(o.logz, rel=1e-11) and st['logvol'] == pytest.approx(o.logvol, abs=1e-13)
        du, dv, dl, dlv, dnc = ops.ns_get_dead(0, st['it'], n)
        ou, ov, ol, olv, onc = o.dead_arrays()
        assert np.array_equal(du, ou) and np.array_equal(dl[:K], ol[:K]) and np.array_equal(dnc, onc)
        np.testing.assert_allclose(dl, ol, rtol=1e-11)
        lu, lv_, ll = ops.ns_get_live(N, n)
        assert np.array_equal(np.sort(lu, axis=0), np.sort(o.live_u, axis=0))
        # first bound on the device, then bounded rounds continue and still agree with the oracle
        nells, lv, warn = ops.ns_update_bound(True, 1.25)
        ops.ns_bound_updated()
        e = OB.bounding_ellipsoid(o.live_u)
        e.scale_to_logvol(e.logvol + math.log(1.25))
        assert nells == 1 and lv == pytest.approx(e.logvol, abs=1e-8)
        # the oracle continues with the DEVICE-built bound (same ellipsoid, but the Jacobi solver's eigenvector signs /
        # order -- hence the directions `axes @ z` of the chains -- are its own)
        db = ops.ns_get_bound(nells, n)
        close(db['ctrs'][0], e.ctr, rtol=1e-9)
        close(db['covs'][0], e.cov, rtol=1e-9)
        o.bound_updated(dict(ctrs=db['ctrs'], ams=db['ams'], axes=db['axes'], logvols=db['logvols'], strict=True))
        assert o.step()
        st = ops.ns_run(1, 0)
        assert (st['rounds'], st['ncall'], st['need_bound']) == (nr + 1, o.ncall, o.need_bound)
        assert st['logz'] == pytest.approx(o.logz, rel=1e-10) and st['scale'] == pytest.approx(o.scale, rel=1e-10)
    finally:
        ops.ns_destroy()


def test_start_rows_are_strictly_above_threshold():
    """Ties at the threshold (sampler.py:471 requires live_logl > loglstar): live points that share the
    K-th lowest logl are never start rows; when NO point is above the threshold the run ends with the reference's
    plateau error (sampler.py:473-475)."""
    dm, om = _models('gauss', 4)
    rng = np.random.default_rng(2)
    N, K, n = 40, 8, 4
    u = 0.5 + 0.04 * rng.standard_normal((N, n))
    v = om.prior_transform(u)
    far = np.argsort([float(om.loglike(x)) for x in v])[:14]
    u[far] = u[far[0]]                               # 14 clones: the 8 lowest and 6 survivors share one logl
    v = om.prior_transform(u)
    l = np.array([float(om.loglike(x)) for x in v])
    order = np.argsort(l, kind='stable')
    assert l[order[K - 1]] == l[order[K + 3]]        # the tie straddles the removal boundary
    b = _bound([0.5 + 0.2 * rng.standard_normal((200, n))])
    o = nsloop.BatchNS(om, u, v, l, K, 'rwalk', 6, 3, scale=0.3, logvol=-1.0, logz=-30.0, loglstar=float(l.min()) - 1, ncall=10,
                       bound=b, dlogz=1e-9)
    ops.bound_set(b['axes'], b['ctrs'], b['ams'], b['logvols'])
    ops.ns_create(dm.model_id(), N, n, K, 0, 6, 3, dlogz=1e-9)
    try:
        ops.ns_set_state(u, v, l, -1.0, -30.0, float(l.min()) - 1, 10, 0.3)
        assert o.step()
        assert np.all(l[o.last['starts']] > o.last['thr'])
        st = ops.ns_run(1, 0)
        assert st['ncall'] == o.ncall and st['logz'] == pytest.approx(o.logz, rel=1e-10)
        assert st['scale'] == pytest.approx(o.scale, rel=1e-10)          # same chains => same accept counts
    finally:
        ops.ns_destroy()
    # plateau: every live point at the same logl except the K lowest -> nothing above the threshold
    l2 = np.full(N, 1.0)
    l2[:3] = 0.0
    ops.ns_create(dm.model_id(), N, n, K, 0, 6, 3, dlogz=1e-9)
    try:
        ops.ns_set_state(u, v, l2, -1.0, -30.0, 0.0, 10, 0.3)
        with pytest.raises(RuntimeError, match='plateau'):
            ops.ns_run(1, 0)
    finally:
        ops.ns_destroy()


@pytest.mark.parametrize('bound,sample,kw', [('multi', 'rwalk', dict(walks=15)), ('single', 'rslice', dict(slices=4)),
                                              ('multi', 'rwalk', dict(walks=15, ncdim=4))])
def test_device_bound_update_equals_host_path(bound, sample, kw):
    """b2n_ns_update_bound (fit + enlarge + make resident without leaving the device) runs the same kernels on the
    same live points as the host route (b2n_ns_get_live -> b2n_multi_decompose -> b2n_scale_to_logvol ->
    b2n_bound_set): whole runs are bit-identical."""
    from dynesty_b200 import _lib
    out = []
    for dev in (True, False):
        ctx = _lib.Context(0)       # a fresh context per run: a ctx remembers whether the Cholesky candidate path failed
        s = nested.NestedSampler(DL.gauss_corr(6, 0.4, 5.0), nlive=200, bound=bound, sample=sample, seed=12, ctx=ctx,
                                 **kw)                          # recently (bound_fast_skip) -- run history, not input
        s.device_bound = dev
        r = s.run_nested(loop='device', batch=10, dlogz=0.5)
        out.append((r, s))
    (a, sa), (b, sb) = out
    assert a.niter == b.niter and a.ncall == b.ncall and sa.nbound == sb.nbound > 3
    assert np.array_equal(a.logl, b.logl) and np.array_equal(a.samples, b.samples)
    assert a.logz[-1] == b.logz[-1]
    assert [h[1] for h in sa.bound_history] == [h[1] for h in sb.bound_history]
    np.testing.assert_allclose([h[2] for h in sa.bound_history], [h[2] for h in sb.bound_history], rtol=0, atol=1e-12)
    close(sa.bound.ctrs, sb.bound.ctrs, rtol=1e-15)            # the host object is synchronised at the end
    close(sa.bound.ams, sb.bound.ams, rtol=1e-15)
