"""One run alone per BASELINE config on one GPU: wall time split into rounds and bound updates (device-resident runs).
usage (GPU box): python scripts/config_runs_timing.py > gpurun_out/config_runs.jsonl"""
import json
import sys
import time

sys.path.insert(0, '.')
from dynesty_b200 import likelihoods as DL, nested


def one(tag, model, **kw):
    run_kw = kw.pop('run', {})
    s = nested.NestedSampler(model, seed=11, **kw)
    t0 = time.perf_counter()
    r = s.run_nested(loop='device', **run_kw)
    wall = time.perf_counter() - t0
    st = getattr(s, 'device_timing', None) or {}
    print(json.dumps(dict(config=tag, wall_s=round(wall, 3), logz=float(r.logz[-1]), niter=int(r.niter), ncall=int(s.ncall),
                          nbound=int(s.nbound), truth=getattr(model, 'logz_truth', None),
                          rounds_s=st.get('rounds_s'), bound_s=st.get('bound_s'))), flush=True)


one('C1 3-D Gaussian single/unif nlive 500', DL.gauss_test3d(), nlive=500, bound='single', sample='unif', run=dict(batch=12))
one('C2 50-D Gaussian multi/rwalk nlive 2000', DL.gauss_corr(50, 0.4, 5.0), nlive=2000, bound='multi', sample='rwalk', walks=70, run=dict(batch=50))
one('C3 25-D eggbox multi/rslice nlive 4000, maxiter 40000', DL.eggbox(25), nlive=4000, bound='multi', sample='rslice', slices=28,
    run=dict(batch=100, dlogz=None, maxiter=40000))
one('C5 10-D shells multi/rslice nlive 500 (static)', DL.shells(10), nlive=500, bound='multi', sample='rslice', run=dict(batch=12))
