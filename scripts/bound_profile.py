"""A few bound updates of one BASELINE shape (for `ncu --metrics gpu__time_duration.sum`).
usage: python scripts/bound_profile.py c4|c2|c3"""
import sys, time, math
sys.path.insert(0, '.')
import numpy as np
from dynesty_b200 import bounding as B
which = sys.argv[1] if len(sys.argv) > 1 else 'c4'
rng = np.random.default_rng(56432)
if which == 'c4':
    n, N, kind = 200, 8000, 'single'
    pts = 0.5 + 0.05 * rng.standard_normal((N, n))
elif which == 'c2':
    n, N, kind = 50, 2000, 'multi'
    pts = 0.5 + 0.03 * rng.standard_normal((N, n))
else:
    n, N, kind = 25, 4000, 'multi'
    ctr = 0.15 + 0.7 * rng.random((8, n))
    pts = ctr[rng.integers(8, size=N)] + 0.01 * rng.standard_normal((N, n))
b = (B.B200MultiEllipsoid if kind == 'multi' else B.B200Ellipsoid)(n)
b.update(pts)
t0 = time.perf_counter()
for _ in range(3):
    b.update(pts)
    b.scale_to_logvol(b.logvol + math.log(1.25))
print(which, 'update ms', 1e3 * (time.perf_counter() - t0) / 3, 'nells', getattr(b, 'nells', 1))
