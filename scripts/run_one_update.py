"""GPU: a few C2-sized bound updates (for ncu captures of the bounding kernels)."""
import sys
sys.path.insert(0, '.')
import numpy as np
import bench
from dynesty_b200 import bounding as B
u_live, _ = bench.make_state(50, 2000)
b = B.B200MultiEllipsoid(50)
for _ in range(3):
    b.update(u_live, rstate=np.random.default_rng(1))
print(b.nells, b.logvol)
