"""Build-container only: full C2-family run with the UNMODIFIED reference (CPU), for the
logZ comparison quoted in DESIGN.md.  usage: python scripts/ref_c2_run.py NDIM NLIVE [SAMPLE [SEED]]"""
import sys, time, math
import numpy as np
sys.path.insert(0, '.')
from oracle import refshim
dynesty = refshim.import_reference()
n = int(sys.argv[1]); nlive = int(sys.argv[2]); sample = sys.argv[3] if len(sys.argv) > 3 else 'rwalk'
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 56432
C = np.full((n, n), 0.4); np.fill_diagonal(C, 1.0)
Cinv = np.linalg.inv(C)
lnorm = -0.5 * (math.log(2 * math.pi) * n + np.linalg.slogdet(C)[1])
loglike = lambda x: -0.5 * x @ Cinv @ x + lnorm
ptform = lambda u: 10. * u - 5.
t0 = time.time()
s = dynesty.NestedSampler(loglike, ptform, n, nlive=nlive, bound='multi', sample=sample,
                          rstate=np.random.default_rng(seed))
s.run_nested(print_progress=False)
r = s.results
print(dict(seed=seed, ndim=n, nlive=nlive, sample=sample, logz=float(r['logz'][-1]), logzerr=float(r['logzerr'][-1]),
           truth=-n * math.log(10.), niter=int(r['niter']), ncall=int(np.sum(r['ncall'])), wall=time.time() - t0))
