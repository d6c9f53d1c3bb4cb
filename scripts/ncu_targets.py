"""One launch of every kernel family that still lacked an `ncu --set full` summary, at the
BASELINE sizes: bound update at C2 (k-means on clusters, Cholesky candidates, Jacobi ladder, moments), the slice
kernel at C3, the uniform kernel at C1, one device round at C2 (fused step kernel + one-chain-per-CTA rwalk).
Run under ncu with -k regex:<kernel> -c 1.  usage (GPU box): python scripts/ncu_targets.py"""
import math
import sys

sys.path.insert(0, '.')
import numpy as np

from dynesty_b200 import ops, likelihoods as DL, bounding as B, nested

sys.path.insert(0, '.')
from bench import make_state

rng = np.random.default_rng(1)
# ---- C2 bound update
u, loglstar = make_state(50, 2000)
b = B.B200MultiEllipsoid(50)
for _ in range(2):
    b.update(u, rstate=rng)
    b.scale_to_logvol(b.logvol + math.log(1.25))
# ---- C3: 25-D eggbox, rslice 28 slices, Q = 4000
m3 = DL.eggbox(25)
u3 = rng.random((4000, 25))
_, l3 = m3.evaluate(u3)
b3 = B.B200MultiEllipsoid(25)
b3.update(u3, rstate=rng)
b3.scale_to_logvol(b3.logvol + math.log(1.25))
b3.make_resident()
thr = float(np.quantile(l3, 0.2))
st = u3[l3 > thr][:4000]
ops.rslice_batch(m3.model_id(), np.resize(st, (4000, 25)), thr, 1.0, 28, 1, ell=b3.random_ells(rng, 4000))
# ---- C1: 3-D Gaussian, single / unif
m1 = DL.gauss_test3d()
u1 = 0.5 + 0.02 * rng.standard_normal((500, 3))
b1 = B.B200Ellipsoid(3)
b1.update(u1, rstate=rng)
b1.make_resident()
_, l1 = m1.evaluate(u1)
ops.unif_batch(m1.model_id(), 500, 3, float(np.quantile(l1, 0.5)), 1)
# ---- a short device-resident C2 run (fused step kernel, unit-cube kernel, one-chain-per-CTA rwalk)
s = nested.NestedSampler(DL.gauss_corr(50, 0.4, 5.0), nlive=2000, bound='multi', sample='rwalk', walks=70, seed=3)
s.run_nested(loop='device', dlogz=None, maxiter=12000)
print('done')
