"""One measured line per SURVEY 8 row (the contract benchmark of the headline row a14 is bench.py):
the B200 entry point on the BASELINE shapes, its algorithmic bytes against the measured HBM
peak, and the UNMODIFIED reference's function for the same row timed beside it on ONE host core of the same
box (baseline/_ref; no reference => the CPU columns are null).

usage (GPU box): python scripts/row_bench.py > gpurun_out/rows.jsonl
Every record: {row, what, shape, gpu_ms, units, unit, gpu_units_per_s, algorithmic_bytes, achieved_GBps,
hbm_frac, ref_cpu_ms (same units of work, scaled from a bounded sample when `ref_sample` says so), speedup}.
"""
import json
import math
import os
import sys
import time

os.environ.setdefault('OMP_NUM_THREADS', '1')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
import numpy as np

from dynesty_b200 import _lib, ops, likelihoods as DL, bounding as B
from bench_configs import top_points, ball_state

SEED = 56432
REF = None
try:
    sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
    import dynesty                                    # noqa: F401  (the offline install of the unmodified reference)
    from dynesty import bounding as RB, internal_samplers as RIS, utils as RU
    REF = dynesty
except Exception:                                     # pragma: no cover
    RB = RIS = RU = None


def hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs'
    except Exception:
        return 6650.0, 'fallback 6650 GB/s (B200_PROFILING.md)'


PEAK, PEAK_SRC = hbm_peak()


def gpu_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts))


def cpu_ms(fn, budget_s=6.0, min_reps=1):
    """median wall of fn() over as many repetitions as fit in the budget"""
    ts = []
    t_end = time.perf_counter() + budget_s
    while len(ts) < min_reps or (time.perf_counter() < t_end and len(ts) < 50):
        t0 = time.perf_counter()
        fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts)), len(ts)


def emit(row, what, shape, g_ms, units, unit, abytes, ref_ms=None, ref_sample=None, kernel_ms=None, note=None):
    t = (kernel_ms if kernel_ms else g_ms) * 1e-3
    rec = dict(row=row, what=what, shape=shape, gpu_ms=round(g_ms, 4), kernel_ms=None if kernel_ms is None else round(kernel_ms, 4),
               units=units, unit=unit, gpu_units_per_s=units / (g_ms * 1e-3),
               algorithmic_bytes=abytes, achieved_GBps=None if abytes is None else abytes / t / 1e9,
               hbm_frac=None if abytes is None else abytes / t / 1e9 / PEAK, hbm_peak_GBps=PEAK,
               ref_cpu_ms=None if ref_ms is None else round(ref_ms, 3), ref_sample=ref_sample,
               speedup_vs_one_core=None if ref_ms is None else ref_ms / g_ms, note=note)
    print(json.dumps(rec), flush=True)


# ---- numpy likelihoods for the reference's chains (what a dynesty user writes; same densities as the device registry)
def np_models():
    def eggbox_l(x, tmax=5.0 * math.pi):
        t = 2.0 * tmax * x - tmax
        return (2.0 + np.prod(np.cos(t / 2.0)))**5.0

    def shells_l(x, r=2.0, w=0.1, c=3.5):
        c1 = np.zeros(len(x)); c1[0] = -c
        c2 = np.zeros(len(x)); c2[0] = c
        lc = lambda ctr: -0.5 * ((np.sqrt(np.sum((x - ctr)**2)) - r) / w)**2 - math.log(math.sqrt(2 * math.pi * w * w))
        return np.logaddexp(lc(c1), lc(c2))

    C3 = np.full((3, 3), 0.95); np.fill_diagonal(C3, 1.0)
    P3, m3 = np.linalg.inv(C3), np.linspace(-1, 1, 3)
    l3 = -0.5 * (math.log(2 * math.pi) * 3 + np.linalg.slogdet(C3)[1])
    C50 = np.full((50, 50), 0.4); np.fill_diagonal(C50, 1.0)
    P50 = np.linalg.inv(C50)
    l50 = -0.5 * (math.log(2 * math.pi) * 50 + np.linalg.slogdet(C50)[1])
    return dict(
        eggbox=(lambda u: u, eggbox_l),
        shells=(lambda u: 12.0 * u - 6.0, shells_l),
        gauss3=(lambda u: 20.0 * u - 10.0, lambda x: -0.5 * np.dot(x - m3, np.dot(P3, x - m3)) + l3),
        gauss50=(lambda u: 10.0 * u - 5.0, lambda x: -0.5 * np.dot(x, np.dot(P50, x)) + l50))


def ref_chains(sampler_cls, u0s, loglstar, axes, scale, pt, ll, ndim, kwargs):
    """the reference's static per-chain method (what dynesty's pool maps over a queue); returns calls"""
    L = RU.LogLikelihood(ll, ndim)
    calls = 0
    for i, u0 in enumerate(u0s):
        a = RIS.SamplerArgument(u=u0, loglstar=loglstar, axes=axes, scale=scale, prior_transform=pt, loglikelihood=L,
                                rseed=SEED + i, kwargs=kwargs)
        calls += sampler_cls.sample(a).ncalls
    return calls


def main():
    ctx = _lib.default_context()
    rng = np.random.default_rng(SEED)
    M = np_models()
    which = set(sys.argv[1:])
    on = lambda k: not which or k in which

    # ------------------------------------------------------------ a1-a3: bounding_ellipsoid
    if on('ell'):
        for tag, N, n in (('C2', 2000, 50), ('C4', 8000, 200)):
            pts = 0.5 + 0.03 * rng.standard_normal((N, n))
            g = gpu_ms(lambda: ops.bounding_ellipsoid(pts), reps=8)
            r = None
            if REF:
                r, k = cpu_ms(lambda: RB.bounding_ellipsoid(pts), budget_s=5.0)
            emit('a1-a3', 'bounding_ellipsoid (mean, cov ddof=1, improve_covar_mat ladder, fmax rescale, axes/am/logvol)',
                 '%s %dx%d' % (tag, N, n), g, 1, 'fit', 16 * N * n, r, None,
                 note='host arrays in and out (H2D of the points inside); latency bound: a chain of single-CTA kernels')

    # ------------------------------------------------------------ a4, a5, a8: MultiEllipsoid.update (+ enlarge)
    clouds = {}
    if on('multi') or on('member') or on('samples') or on('boot'):
        Cm = np.full((50, 50), 0.4); np.fill_diagonal(Cm, 1.0)
        clouds['C2 unimodal 2000x50'] = ball_state(50, 2000, rng, lambda v: (v @ np.linalg.cholesky(Cm).T + 5) / 10)
        cc = 0.2 + 0.6 * rng.random((8, 25))
        clouds['C3-shaped 8 clusters 4000x25'] = np.concatenate([c + 0.01 * rng.standard_normal((500, 25)) for c in cc])
    if on('multi'):
        for tag, pts in clouds.items():
            N, n = pts.shape
            b = B.B200MultiEllipsoid(n, ctx=ctx)

            def gfn():
                b.update(pts, rstate=np.random.default_rng(SEED))
                b.scale_to_logvol(b.logvol + math.log(1.25))
            g = gpu_ms(gfn, reps=10)
            gs = gpu_ms(lambda: b.scale_to_logvol(b.logvol + 1e-3), reps=10)
            r = rs = None
            nref = None
            if REF:
                rb = RB.MultiEllipsoid(n)

                def rfn():
                    rb.update(pts, rstate=np.random.default_rng(SEED))
                    rb.scale_to_logvol(rb.logvol + math.log(1.25))
                r, _ = cpu_ms(rfn, budget_s=6.0)
                rs, _ = cpu_ms(lambda: rb.scale_to_logvol(rb.logvol + 1e-3), budget_s=1.0)
                nref = rb.nells
            depth = max(1, int(math.log2(max(N / (2.0 * n), 1))) + 1)
            emit('a4-a5', 'MultiEllipsoid.update + enlarge 1.25 (candidate tree of 2-means splits, accept tests, containment check)',
                 tag, g, 1, 'update', 96 * N * n * depth, r, None,
                 note='nells %d (reference %s); bytes = SURVEY 8d per level x %d levels' % (b.nells, nref, depth))
            emit('a8', 'MultiEllipsoid.scale_to_logvol (scalar target)', tag + ' K=%d' % b.nells, gs, 1, 'call',
                 24 * n * n * b.nells, rs, None)

    # ------------------------------------------------------------ a6: membership, a7: samples
    if on('member') or on('samples'):
        pts = clouds['C3-shaped 8 clusters 4000x25']
        N, n = pts.shape
        b = B.B200MultiEllipsoid(n, ctx=ctx)
        b.update(pts, rstate=np.random.default_rng(SEED))
        rb = None
        if REF:
            rb = RB.MultiEllipsoid(n)
            rb.update(pts, rstate=np.random.default_rng(SEED))
        K = b.nells
        if on('member'):
            Mq = 200000
            x = pts[rng.integers(N, size=Mq)] + 0.002 * rng.standard_normal((Mq, n))
            g = gpu_ms(lambda: ops.membership(x, b.ctrs, b.ams), reps=8)
            r = None
            samp = None
            if REF:
                xs = x[:4000]
                r, _ = cpu_ms(lambda: [rb.within(xi) for xi in xs], budget_s=4.0)
                r *= Mq / len(xs)
                samp = '4000 points, scaled'
            emit('a6', 'MultiEllipsoid.within / overlap / contains for a batch of points (strict < 1, M x K mask + q)',
                 'M=%d K=%d n=%d' % (Mq, K, n), g, Mq, 'points', 8 * Mq * n + 8 * K * n * n + Mq * K, r, samp,
                 note='host arrays: H2D of the points and D2H of the mask inside the time')
        if on('samples'):
            S = 200000
            g = gpu_ms(lambda: b.samples(S, rstate=np.random.default_rng(1)), reps=8)
            r = None
            samp = None
            if REF:
                r, _ = cpu_ms(lambda: rb.samples(2000, rstate=np.random.default_rng(1)), budget_s=4.0)
                r *= S / 2000
                samp = '2000 draws, scaled'
            emit('a7', 'MultiEllipsoid.samples (volume-weighted pick, ball draw, 1/q acceptance)', 'S=%d K=%d n=%d' % (S, K, n),
                 g, S, 'draws', S * (8 * n * n + K * 8 * n * n), r, samp,
                 note='bytes = SURVEY 8d no-reuse model (axes + K precision matrices per draw); the matrices live in L2 / shared memory')

    # ------------------------------------------------------------ a9: bootstrap expansion
    if on('boot'):
        for tag, pts, multi in (('C1-shaped 500x3 single, 5 replicas', 0.5 + 0.05 * rng.standard_normal((500, 3)), 0),
                                ('C2 2000x50 multi, 5 replicas', clouds['C2 unimodal 2000x50'], 1)):
            N, n = pts.shape
            g = gpu_ms(lambda: ops.bootstrap_expand(pts, multi, 5, SEED, 1000), reps=5, warm=1)
            r = None
            if REF:
                def rfn():
                    for s in range(5):
                        RB._ellipsoid_bootstrap_expand((bool(multi), pts, np.random.SeedSequence(s)))
                try:
                    r, _ = cpu_ms(rfn, budget_s=6.0)
                except Exception as e:          # signature drift must not lose the GPU number
                    r = None
                    tag += ' [reference call failed: %r]' % (e,)
            emit('a9', '_ellipsoid_bootstrap_expand x 5 (resample, refit, max out-of-bag distance)', tag, g, 5, 'replicas',
                 5 * 16 * N * n, r, None)

    # ------------------------------------------------------------ a15-a17 + f2: chains other than rwalk
    def chain_row(row, what, tag, model, u_live, loglstar, bound_kind, sampler, steps, Q, npkey, ref_cls, ref_kwargs, scale=1.0,
                  nref=24, bytes_per_call=None):
        n = model.ndim
        bnd = (B.B200MultiEllipsoid if bound_kind == 'multi' else B.B200Ellipsoid)(n, ctx=ctx)
        bnd.update(u_live, rstate=np.random.default_rng(SEED), bootstrap=5 if sampler == 'unif' else 0)
        if sampler != 'unif':
            bnd.scale_to_logvol(bnd.logvol + math.log(1.25))
        bnd.make_resident()
        mid = model.model_id(ctx)
        r2 = np.random.default_rng(1)
        ctx.set_timing(True)
        kms, walls, calls = [], [], []
        for it in range(10):
            starts = r2.integers(len(u_live), size=Q)
            ell = bnd.random_ells(r2, Q)
            t0 = time.perf_counter()
            if sampler in ('rslice', 'slice'):
                fn = ops.rslice_batch if sampler == 'rslice' else ops.slice_batch
                o = fn(mid, u_live[starts], loglstar, scale, steps, SEED, chain0=it * Q, ell=ell, ctx=ctx)
                ne, nc = max(int(o['n_expand'].sum()), 1), int(o['n_contract'].sum())
                if it < 5:
                    scale *= min(max(ne * 2. / (ne + nc), 0.5), 2.)
            elif sampler == 'unif':
                o = ops.unif_batch(mid, Q, n, loglstar, SEED, chain0=it * Q, ctx=ctx)
            else:
                o = ops.unitcube_batch(mid, Q, n, loglstar, SEED, chain0=it * Q, ctx=ctx)
            wall = time.perf_counter() - t0
            if it >= 5:
                kms.append(ctx.last_kernel_ms()); walls.append(1e3 * wall); calls.append(int(o['ncall'].sum()))
        ctx.set_timing(False)
        ncall = float(np.mean(calls))
        r_ms = None
        samp = None
        if REF:
            pt, ll = M[npkey]
            kw = dict(ref_kwargs)
            axes = None
            if sampler == 'unif':
                rbnd = RB.Ellipsoid(n)
                rbnd.update(u_live, rstate=np.random.default_rng(SEED), bootstrap=5)
                kw.update(bound=rbnd, ndim=n, n_cluster=n, nonbounded=None)
            elif sampler != 'unitcube':
                axes = np.asarray(bnd.get_random_axes(np.random.default_rng(2)))
            u0s = u_live[r2.integers(len(u_live), size=nref)]
            t0 = time.perf_counter()
            rc = ref_chains(ref_cls, u0s, loglstar, axes, scale, pt, ll, n, kw)
            dt = 1e3 * (time.perf_counter() - t0)
            r_ms = dt * ncall / max(rc, 1)            # the same number of likelihood calls on one core
            samp = '%d chains (%d calls, %.1f s), scaled to the fill\'s calls' % (nref, rc, dt / 1e3)
        emit(row, what, tag, float(np.mean(walls)), ncall, 'likelihood calls',
             None if bytes_per_call is None else bytes_per_call * ncall, r_ms, samp, kernel_ms=float(np.mean(kms)),
             note='gpu_ms = plug-in call with host arrays; achieved bytes against kernel_ms')

    if on('chains'):
        m = DL.eggbox(25)
        u, ls = top_points(m, 4000, 10, rng, ctx)
        chain_row('a15', 'RSliceSampler.sample x queue (28 slices per chain, stepping out + shrink)', 'C3 25-D eggbox, 4000 chains',
                  m, u, ls, 'multi', 'rslice', 28, 4000, 'eggbox', RIS.RSliceSampler if REF else None,
                  dict(slices=28, nonperiodic=None, slice_doubling=False), bytes_per_call=16 * 25)
        chain_row('a16', 'SliceSampler.sample x queue (3 x 25 axis-aligned slices per chain)', '25-D eggbox, 4000 chains',
                  m, u, ls, 'multi', 'slice', 3, 4000, 'eggbox', RIS.SliceSampler if REF else None,
                  dict(slices=3, nonperiodic=None, slice_doubling=False), nref=8, bytes_per_call=16 * 25)
        m5 = DL.shells(10)
        u5, ls5 = top_points(m5, 500, 400, rng, ctx)
        chain_row('a15', 'RSliceSampler.sample x queue (13 slices per chain)', 'C5 10-D shells, 500 chains (one batch)',
                  m5, u5, ls5, 'multi', 'rslice', 13, 500, 'shells', RIS.RSliceSampler if REF else None,
                  dict(slices=13, nonperiodic=None, slice_doubling=False), bytes_per_call=16 * 10 + 16 * 10)
        m1 = DL.gauss_test3d()
        u1 = ball_state(3, 500, rng, lambda v: (v @ np.linalg.cholesky(0.95 + 0.05 * np.eye(3)).T + np.linspace(-1, 1, 3) + 10) / 20)
        _, l1 = m1.evaluate(u1, ctx=ctx)
        chain_row('a17', 'UniformBoundSampler.sample x queue (draw in the bound until logl > L*)', 'C1 3-D Gaussian single, 500 chains',
                  m1, u1, float(l1.min()) - 1e-9, 'single', 'unif', 1, 500, 'gauss3', RIS.UniformBoundSampler if REF else None, {},
                  nref=200, bytes_per_call=8 * 9 + 8 * 9)
        m2 = DL.gauss_corr(50, 0.4, 5.0)
        u2 = rng.random((2000, 50))
        _, l2 = m2.evaluate(u2, ctx=ctx)
        chain_row('f2', 'UnitCubeSampler.sample x queue (prior draws until logl > L*, the phase before the first bound)',
                  'C2 50-D, 2000 chains at the median of the prior sample', m2, u2, float(np.median(l2)), 'single', 'unitcube', 1, 2000,
                  'gauss50', RIS.UnitCubeSampler if REF else None, dict(ndim=50), nref=200, bytes_per_call=8 * 50 * 50 + 24 * 50)

    # ------------------------------------------------------------ f3: RadFriends / SupFriends update
    if on('friends'):
        N, n = 1000, 10
        pts = np.concatenate([0.3 + 0.03 * rng.standard_normal((N // 2, n)), 0.7 + 0.03 * rng.standard_normal((N // 2, n))])
        for kind, cls in (('balls', 'RadFriends'), ('cubes', 'SupFriends')):
            prev = ops.friends_update(pts, kind, use_clustering=False)
            g = gpu_ms(lambda: ops.friends_update(pts, kind, am_prev=prev['am'], use_clustering=True), reps=8)
            r = None
            if REF:
                rbf = getattr(RB, cls)(n)
                rbf.update(pts, rstate=np.random.default_rng(1), use_clustering=False)
                r, _ = cpu_ms(lambda: rbf.update(pts, rstate=np.random.default_rng(1), use_clustering=True), budget_s=5.0)
            emit('f3', '%s.update (clusters under the current metric, re-centred covariance, leave-one-out radius)' % cls,
                 '%dx%d, 2 clusters' % (N, n), g, 1, 'update', 8 * N * N * 2 + 16 * N * n, r, None,
                 note='bytes: the N x N pair tests (adjacency + nearest neighbour) at 8 B per pair per pass')


if __name__ == '__main__':
    main()
