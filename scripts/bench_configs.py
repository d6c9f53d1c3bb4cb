"""Supplementary throughput table for the BASELINE.json configs other than the headline C2
(bench.py is the contract benchmark).  One queue fill per step, host-pointer plug-in calls
(e2e) + the kernel time recorded by the library.  usage (GPU box): python scripts/bench_configs.py
"""
import json
import math
import sys
import time

sys.path.insert(0, '.')
import numpy as np

from dynesty_b200 import _lib, ops, likelihoods as DL, bounding as B

SEED = 56432


def top_points(model, nlive, factor, rng, ctx):
    u = rng.random((nlive * factor, model.ndim))
    _, l = model.evaluate(u, ctx=ctx)
    keep = np.argsort(l)[-nlive:]
    return np.ascontiguousarray(u[keep]), float(l[keep].min()) - 1e-9


def ball_state(n, nlive, rng, to_u):
    z = rng.standard_normal((nlive, n))
    z *= (rng.random(nlive)**(1. / n) / np.linalg.norm(z, axis=1))[:, None]
    return to_u(math.sqrt(n) * z)


def run(tag, model, u_live, loglstar, bound_kind, sampler, steps_per_chain, Q, ctx, reps=10, scale=1.0):
    n = model.ndim
    bound = (B.B200MultiEllipsoid if bound_kind == 'multi' else B.B200Ellipsoid)(n, ctx=ctx)
    bound.update(u_live, rstate=np.random.default_rng(SEED), bootstrap=5 if sampler == 'unif' else 0)
    t0 = time.perf_counter()
    t_scale = 0.0
    for _ in range(3):
        bound.update(u_live, rstate=np.random.default_rng(SEED), bootstrap=5 if sampler == 'unif' else 0)
        if sampler != 'unif':
            t1 = time.perf_counter()
            bound.scale_to_logvol(bound.logvol + math.log(1.25))
            t_scale += time.perf_counter() - t1
    bound_ms = 1e3 * (time.perf_counter() - t0) / 3
    scale_ms = 1e3 * t_scale / 3
    bound.make_resident()
    rng = np.random.default_rng(1)
    mid = model.model_id(ctx)
    ctx.set_timing(True)
    kms, walls, calls = [], [], []
    chain0 = 0
    for it in range(reps + 6):
        starts = rng.integers(len(u_live), size=Q)
        ell = bound.random_ells(rng, Q)
        t0 = time.perf_counter()
        if sampler == 'rwalk':
            o = ops.rwalk_batch(mid, u_live[starts], loglstar, scale, steps_per_chain, SEED, chain0=chain0, ell=ell, ctx=ctx)
            acc, rej = int(o['n_accept'].sum()), int(o['n_reject'].sum())
            if it < 6:
                scale *= math.exp((acc / (acc + rej) - 0.5) / n / 0.5 * 8)      # faster tuning for the warm-up
        elif sampler in ('rslice', 'slice'):
            fn = ops.rslice_batch if sampler == 'rslice' else ops.slice_batch
            o = fn(mid, u_live[starts], loglstar, scale, steps_per_chain, SEED, chain0=chain0, ell=ell, ctx=ctx)
            ne, nc = max(int(o['n_expand'].sum()), 1), int(o['n_contract'].sum())
            if it < 6:
                scale *= min(max(ne * 2. / (ne + nc), 0.5), 2.)
        else:
            o = ops.unif_batch(mid, Q, n, loglstar, SEED, chain0=chain0, ctx=ctx)
        wall = time.perf_counter() - t0
        chain0 += Q
        if it >= 6:
            kms.append(ctx.last_kernel_ms())
            walls.append(wall)
            calls.append(int(o['ncall'].sum()))
    ctx.set_timing(False)
    ncall = float(np.mean(calls))
    out = dict(config=tag, ndim=n, nlive=len(u_live), bound=bound_kind, nells=int(getattr(bound, 'nells', 1)),
               sampler=sampler, steps_per_chain=steps_per_chain, queue=Q, calls_per_fill=ncall,
               kernel_ms=round(float(np.mean(kms)), 4), kernel_calls_per_s=ncall / (np.mean(kms) * 1e-3),
               e2e_ms=round(1e3 * float(np.mean(walls)), 4), e2e_calls_per_s=ncall / float(np.mean(walls)),
               bound_update_ms=round(bound_ms, 3), of_which_scale_to_logvol_ms=round(scale_ms, 3),
               scale=round(scale, 4))
    print(json.dumps(out), flush=True)


def main():
    ctx = _lib.Context(0)
    rng = np.random.default_rng(SEED)
    which = sys.argv[1:] or ['c1', 'c2', 'c3', 'c4', 'c4shard', 'c5']
    if 'c1' in which:
        m = DL.gauss_test3d()
        u = ball_state(3, 500, rng, lambda v: (v @ np.linalg.cholesky(0.95 + 0.05 * np.eye(3)).T + np.linspace(-1, 1, 3) + 10) / 20)
        _, l = m.evaluate(u, ctx=ctx)
        run('C1 3-D Gaussian single/unif nlive=500', m, u, float(l.min()) - 1e-9, 'single', 'unif', 1, 500, ctx)
    if 'c2' in which:
        m = DL.gauss_corr(50, 0.4, 5.0)
        Cm = np.full((50, 50), 0.4)
        np.fill_diagonal(Cm, 1.0)
        u = ball_state(50, 2000, rng, lambda v: (v @ np.linalg.cholesky(Cm).T + 5) / 10)
        _, l = m.evaluate(u, ctx=ctx)
        run('C2 50-D corr Gaussian multi/rwalk nlive=2000', m, u, float(l.min()) - 1e-9, 'multi', 'rwalk', 70, 2000, ctx, scale=0.2)
        run('C2 (rslice instead of rwalk)', m, u, float(l.min()) - 1e-9, 'multi', 'rslice', 53, 2000, ctx, scale=0.3)
    if 'c3' in which:
        m = DL.eggbox(25)
        u, ls = top_points(m, 4000, 10, rng, ctx)
        run('C3 25-D eggbox multi/rslice nlive=4000', m, u, ls, 'multi', 'rslice', 28, 4000, ctx)
    if 'c4' in which or 'c4shard' in which:
        from scipy.special import ndtr
        m = DL.iid_normal_ppf(200)
        u = ball_state(200, 8000, rng, lambda v: ndtr(v))
        _, l = m.evaluate(u, ctx=ctx)
        if 'c4' in which:
            run('C4 200-D iid normal single/rwalk nlive=8000 (whole queue on 1 GPU)', m, u, float(l.min()) - 1e-9, 'single', 'rwalk', 220, 8000, ctx, reps=5, scale=0.1)
        if 'c4shard' in which:
            run('C4 per-GPU shard of the 8-GPU config (1000 chains)', m, u, float(l.min()) - 1e-9, 'single', 'rwalk', 220, 1000, ctx, reps=5, scale=0.1)
    if 'c5' in which:
        m = DL.shells(10)
        u, ls = top_points(m, 500, 400, rng, ctx)
        run('C5 10-D shells multi/rslice nlive=500 (one batch)', m, u, ls, 'multi', 'rslice', 13, 500, ctx)


if __name__ == '__main__':
    main()
