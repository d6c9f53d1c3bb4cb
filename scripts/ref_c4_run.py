"""Build-container only: BASELINE C4 (200-D iid normal, normal-ppf prior, single/rwalk, nlive 8000) with the
UNMODIFIED reference on one CPU core, for the logZ comparison in DESIGN.md.  usage: python scripts/ref_c4_run.py [SEED]"""
import sys, time, math
import numpy as np
from scipy.special import ndtri
sys.path.insert(0, '.')
from oracle import refshim
dynesty = refshim.import_reference()
n = 200; seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lnorm = -0.5 * n * math.log(2 * math.pi)
loglike = lambda x: -0.5 * float(np.dot(x, x)) + lnorm
ptform = lambda u: ndtri(u)
t0 = time.time()
s = dynesty.NestedSampler(loglike, ptform, n, nlive=8000, bound='single', sample='rwalk', rstate=np.random.default_rng(seed))
s.run_nested(print_progress=False)
r = s.results
print(dict(seed=seed, ndim=n, nlive=8000, sample='rwalk', bound='single', logz=float(r['logz'][-1]), logzerr=float(r['logzerr'][-1]),
           truth=lnorm - 0.5 * n * math.log(2), niter=int(r['niter']), ncall=int(np.sum(r['ncall'])), wall=time.time() - t0))
