"""GPU: C2 full-run logZ under variations of the dispatch parameters (diagnostic)."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from dynesty_b200 import likelihoods as DL, nested

def run(tag, ndim=50, nlive=2000, **kw):
    m = DL.gauss_corr(ndim, 0.4, 5.0)
    t0 = time.time()
    s = nested.NestedSampler(m, nlive=nlive, **kw)
    r = s.run_nested()
    K = [h[1] for h in r.bound_history]
    sc = [x[1] for x in r.scale_history]
    print(json.dumps(dict(tag=tag, logz=round(float(r.logz[-1]), 3), err=round(float(r.logzerr[-1]), 3),
                          truth=round(m.logz_truth, 3), niter=r.niter, ncall=r.ncall, nbound=r.nbound,
                          nells_max=max(K) if K else 0, nells_last=K[-1] if K else 0,
                          scale_last=round(sc[-1], 4) if sc else None, wall=round(time.time() - t0, 2))), flush=True)

which = sys.argv[1:] or ['base', 'single', 'q200', 'walks200', 'seed2', 'rslice', 'n20']
if 'base' in which: run('base multi rwalk Q=nlive', bound='multi', sample='rwalk')
if 'single' in which: run('single rwalk', bound='single', sample='rwalk')
if 'q200' in which: run('multi rwalk Q=200', bound='multi', sample='rwalk', queue_size=200)
if 'q16' in which: run('multi rwalk Q=16', bound='multi', sample='rwalk', queue_size=16)
if 'walks200' in which: run('multi rwalk walks=200', bound='multi', sample='rwalk', walks=200)
if 'seed2' in which: run('base seed 2', bound='multi', sample='rwalk', seed=2)
if 'rslice' in which: run('multi rslice', bound='multi', sample='rslice')
if 'n20' in which: run('n=20 multi rwalk', ndim=20, bound='multi', sample='rwalk')
if 'nlive500' in which: run('nlive=500 multi rwalk', nlive=500, bound='multi', sample='rwalk')
