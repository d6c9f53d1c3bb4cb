"""GPU: the BASELINE configs end to end with the rounds on the device (run_nested(loop='device')):
logZ against the analytic truth where there is one, wall time, calls/s.  One JSON line per run.
usage: python scripts/configs_device_loop.py [c1 c2r c3 c4 c5 ...]"""
import json
import sys
import time

sys.path.insert(0, '.')
import numpy as np
from dynesty_b200 import likelihoods as DL, nested


def run(tag, model, loop='device', batch=None, seed=56432, **kw):
    rk = {k: kw.pop(k) for k in ('dlogz', 'maxiter', 'maxcall') if k in kw}
    t0 = time.time()
    s = nested.NestedSampler(model, seed=seed, **kw)
    r = s.run_nested(loop=loop, batch=batch, **rk)
    wall = time.time() - t0
    K = [h[1] for h in r.bound_history]
    out = dict(config=tag, loop=loop, batch=getattr(s, 'batch', None), logz=round(float(r.logz[-1]), 3),
               logzerr=round(float(r.logzerr[-1]), 3), truth=model.logz_truth, niter=int(r.niter), ncall=int(r.ncall),
               nbound=int(r.nbound), nells_max=max(K) if K else 0, wall_s=round(wall, 2),
               calls_per_s=round(r.ncall / wall), rounds=getattr(s, 'device_rounds', None))
    if hasattr(s, 'device_timing'):
        out.update(rounds_s=round(s.device_timing['rounds_s'], 3), bound_s=round(s.device_timing['bound_s'], 3))
    print(json.dumps(out), flush=True)


which = sys.argv[1:] or ['c1', 'c2', 'c2r', 'c3', 'c4', 'c5']
if 'c1' in which:     # 3-D Gaussian, single/unif (bootstrap 5, the reference's default for unif)
    run('C1 3-D gauss single/unif nlive=500', DL.gauss_test3d(), batch=50, nlive=500, bound='single', sample='unif',
        queue_size=64)
    run('C1 3-D gauss single/unif nlive=500 (host loop)', DL.gauss_test3d(), loop='host', nlive=500, bound='single',
        sample='unif', queue_size=64)
if 'c2' in which:
    run('C2 50-D gauss multi/rwalk nlive=2000', DL.gauss_corr(50, 0.4, 5.0), batch=50, nlive=2000, bound='multi',
        sample='rwalk', queue_size=200)
if 'c2r' in which:
    run('C2 50-D gauss multi/rslice nlive=2000', DL.gauss_corr(50, 0.4, 5.0), batch=200, nlive=2000, bound='multi',
        sample='rslice', queue_size=200)
if 'c3' in which:     # eggbox 25-D: logl nearly flat, default dlogz stops early -> fixed iteration budget
    run('C3 25-D eggbox multi/rslice nlive=4000 (maxiter 40000)', DL.eggbox(25), batch=400, nlive=4000, bound='multi',
        sample='rslice', queue_size=400, dlogz=1e-9, maxiter=40000)
    run('C3-pinned 2-D eggbox multi/rslice nlive=1000', DL.eggbox(2), batch=50, nlive=1000, bound='multi',
        sample='rslice', queue_size=100)
if 'c4' in which:
    run('C4 200-D iid normal single/rwalk nlive=8000 (one GPU)', DL.iid_normal_ppf(200), batch=200, nlive=8000,
        bound='single', sample='rwalk', queue_size=800)
if 'c5' in which:     # shells: static run of the C5 likelihood (the dynamic driver is dynesty's, tests/test_gpu_dropin.py)
    run('C5-static 10-D shells multi/rslice nlive=500', DL.shells(10), batch=25, nlive=500, bound='multi',
        sample='rslice', queue_size=50)
    run('C5-static 10-D shells multi/rslice nlive=2000', DL.shells(10), batch=100, nlive=2000, bound='multi',
        sample='rslice', queue_size=200)
