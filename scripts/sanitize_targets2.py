"""Small instances of the code paths added in the second half of round 2, for compute-sanitizer (memcheck / racecheck /
synccheck): the restructured bound update -- speculative root fit on the side stream (adopted and not adopted), the two
halves of the candidate fit on two streams, symmetric sweeps in registers (both instantiations: n <= 64 and n > 64),
thread-per-row k-means on staged rows, the warp-per-row k-means with a partly staged CTA (rows beyond the stage),
deferred stats, half-warp Jacobi (even and odd n) -- the streamed DMMA rwalk kernel with helper warps (1, 3, 16 chains
per CTA), the friends metric (shared Jacobi), and a device-resident run whose updates use all of it.
usage (GPU box): compute-sanitizer --tool racecheck python scripts/sanitize_targets2.py"""
import math
import sys

sys.path.insert(0, '.')
import numpy as np

from dynesty_b200 import ops, likelihoods as DL, bounding as B, nested

rng = np.random.default_rng(7)
print('bound updates'); sys.stdout.flush()
uni = 0.5 + 0.03 * rng.standard_normal((600, 8))
two = np.concatenate([0.3 + 0.02 * rng.standard_normal((300, 8)), 0.7 + 0.02 * rng.standard_normal((300, 8))])
wide = 0.5 + 0.03 * rng.standard_normal((640, 70))                      # sweeps <8, 4>, eig n = 70
big = np.concatenate([0.3 + 0.02 * rng.standard_normal((9000, 8)), 0.7 + 0.02 * rng.standard_normal((7000, 8))])
for pts in (uni, two, wide):
    o = ops.multi_decompose(pts)
    print(pts.shape, 'nells', o['nells']); sys.stdout.flush()
o = ops.multi_decompose(big)                                             # 2000 rows per k-means CTA: partly staged
print(big.shape, 'nells', o['nells']); sys.stdout.flush()
for n in (3, 6, 7):
    ops.bounding_ellipsoid(0.5 + 0.05 * rng.standard_normal((120, n)))
ops.improve_covar(np.diag([1., -1., 2., 0.5, 3.]))
print('friends'); sys.stdout.flush()
prev = ops.friends_update(two, 'balls', use_clustering=False)
ops.friends_update(two, 'balls', am_prev=prev['am'], use_clustering=True)
print('rwalk n > 64'); sys.stdout.flush()
m = DL.iid_normal_ppf(72)
u = 0.5 + 0.05 * rng.standard_normal((400, 72))
b = B.B200Ellipsoid(72)
b.update(u, rstate=rng)
b.scale_to_logvol(b.logvol + math.log(1.25))
b.make_resident()
_, l = m.evaluate(u)
thr = float(np.quantile(l, 0.3))
st = u[l > thr]
for Q in (1, 3, 24, 160 * 16):          # 1 and 3 chains on a CTA (helpers), 16 per CTA on every SM
    ops.rwalk_batch(m.model_id(), np.resize(st, (Q, 72)), thr, 0.3, 3, 5)
print('device run'); sys.stdout.flush()
s = nested.NestedSampler(DL.gauss_corr(6, 0.4, 5.0), nlive=240, bound='multi', sample='rwalk', walks=6, seed=3)
s.run_nested(loop='device', batch=8, dlogz=None, maxiter=1500)
print('nbound', s.nbound)
print('done')
