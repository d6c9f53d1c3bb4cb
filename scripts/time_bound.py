"""Where does a single-ellipsoid bound update spend its wall time? (GPU box)"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
from dynesty_b200 import _lib, ops
n, N = int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 8000
rng = np.random.default_rng(1)
pts = 0.5 + 0.05 * rng.standard_normal((N, n))
ctx = _lib.default_context()
def t(fn, k=5):
    fn(); t0 = time.perf_counter()
    for _ in range(k): fn()
    return 1e3 * (time.perf_counter() - t0) / k
print('host pageable  bounding_ellipsoid ms', t(lambda: ops.bounding_ellipsoid(pts)))
pp = torch.from_numpy(pts).pin_memory().numpy()
print('host pinned    bounding_ellipsoid ms', t(lambda: ops.bounding_ellipsoid(pp)))
d = torch.from_numpy(pts).cuda()
outs = dict(ctr=torch.empty(n, dtype=torch.float64, device='cuda'), cov=torch.empty(n, n, dtype=torch.float64, device='cuda'),
            am=torch.empty(n, n, dtype=torch.float64, device='cuda'), axes=torch.empty(n, n, dtype=torch.float64, device='cuda'),
            axlens=torch.empty(n, dtype=torch.float64, device='cuda'), lv=torch.empty(1, dtype=torch.float64, device='cuda'))
import ctypes as C
def dev():
    ctx.set_pointer_mode(_lib.PTR_DEVICE)
    w = C.c_uint32(0)
    ctx.check(ctx.lib.b2n_bounding_ellipsoid(ctx.h, d.data_ptr(), N, n, outs['ctr'].data_ptr(), outs['cov'].data_ptr(),
              outs['am'].data_ptr(), outs['axes'].data_ptr(), outs['axlens'].data_ptr(), outs['lv'].data_ptr(), C.addressof(w)))
    ctx.set_pointer_mode(_lib.PTR_HOST)
print('device-resident bounding_ellipsoid ms', t(dev))
print('multi_decompose (host) ms', t(lambda: ops.multi_decompose(pts), 3))
