"""improve_covar on ill-conditioned matrices: which (n, kind) breaks cov @ am = I ?"""
import sys
sys.path.insert(0, '.')
import numpy as np
from dynesty_b200 import ops
from oracle import bounding as OB
for n in (40, 100, 117, 118, 130, 200):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    S = A @ A.T / n
    lam, V = np.linalg.eigh(S)
    for kind in ('good', 'illcond', 'indefinite', 'singular'):
        l2 = lam.copy()
        if kind == 'illcond':
            l2[0] = l2[-1] * 1e-15
        elif kind == 'indefinite':
            l2[:2] = -l2[:2]
        elif kind == 'singular':
            l2[:max(1, n // 4)] = 0.0
        M = (V * l2) @ V.T
        M = 0.5 * (M + M.T)
        good, cov, am, axes, warn = ops.improve_covar(M)
        og, oc, oa, ox, _ = OB.improve_covar_mat(M)
        w = np.linalg.eigvalsh(cov)
        print(n, kind, 'good', good, og, 'cov err %.2e' % (np.abs(cov - oc).max() / np.abs(oc).max()),
              'am err %.2e' % (np.abs(am - oa).max() / np.abs(oa).max()), '|cov am - I| %.2e' % np.abs(cov @ am - np.eye(n)).max(),
              'oracle %.2e' % np.abs(oc @ oa - np.eye(n)).max(), 'axes %.2e' % (np.abs(axes @ axes.T - cov).max() / np.abs(cov).max()),
              'cond %.2e' % (w.max() / w.min()), flush=True)
