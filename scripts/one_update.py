"""Two C2 bound updates (b2n_multi_decompose on the bench's live set) for ncu launch lists / captures."""
import sys
sys.path.insert(0, '.')
import numpy as np
from dynesty_b200 import ops
import bench
u, _ = bench.make_state(50, 2000)
for _ in range(3):
    o = ops.multi_decompose(u)
print('nells', o['nells'])
