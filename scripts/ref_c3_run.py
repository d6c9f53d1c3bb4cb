"""Build-container only: the UNMODIFIED reference on C3 (25-D eggbox, multi / rslice, nlive 4000) run to a FIXED
iteration count (no analytic truth, compare logZ at the same maxiter).  Appends one JSON line
per seed: logz of the dead points only (add_live=False) and the logZ trajectory every 5000 iterations.
usage: python scripts/ref_c3_run.py MAXITER SEED [NLIVE [NDIM]]"""
import json, sys, time, math
import numpy as np
sys.path.insert(0, '.')
from oracle import refshim
dynesty = refshim.import_reference()
maxiter = int(sys.argv[1]); seed = int(sys.argv[2])
nlive = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmax = 5.0 * math.pi
def loglike(x):
    t = 2.0 * tmax * x - tmax
    return (2.0 + np.prod(np.cos(t / 2.0)))**5.0
ptform = lambda u: u
t0 = time.time()
s = dynesty.NestedSampler(loglike, ptform, n, nlive=nlive, bound='multi', sample='rslice', rstate=np.random.default_rng(seed))
s.run_nested(print_progress=False, dlogz=None, maxiter=maxiter, add_live=False)
r = s.results
lz = np.asarray(r['logz'])
print(json.dumps(dict(seed=seed, ndim=n, nlive=nlive, sample='rslice', slices=3 + n, maxiter=maxiter, logz_dead=float(lz[-1]),
                      niter=int(r['niter']), ncall=int(np.sum(r['ncall'])), logz_every_5000=[float(x) for x in lz[4999::5000]],
                      wall=time.time() - t0)))
