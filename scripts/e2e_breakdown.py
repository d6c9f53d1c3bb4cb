"""Where does the host-buffer (e2e) C2 rwalk fill spend its wall time?  (GPU box)
Splits one plug-in call into: host-side proposal bookkeeping, the C-ABI call with pinned host
buffers, the same call with device pointers (+ sync), and the kernel itself."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
import bench
from dynesty_b200 import _lib, ops

n, Q, walks = 50, 2000, 70
ctx = _lib.Context(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ctx.set_stream(st.cuda_stream)
from dynesty_b200 import likelihoods as DL, bounding as B
model = DL.gauss_corr(n, 0.4, 5.0)
u_live, loglstar = bench.make_state(n, Q)
bound = B.B200MultiEllipsoid(n, ctx=ctx)
bound.update(u_live, rstate=np.random.default_rng(1))
bound.scale_to_logvol(bound.logvol + np.log(1.25))
bound.make_resident()
mid = model.model_id(ctx)
rng = np.random.default_rng(1)
h_u0 = torch.empty((Q, n), dtype=torch.float64).pin_memory()
h = dict(u=torch.empty((Q, n), dtype=torch.float64).pin_memory(), v=torch.empty((Q, n), dtype=torch.float64).pin_memory(),
         logl=torch.empty(Q, dtype=torch.float64).pin_memory(), n_accept=torch.empty(Q, dtype=torch.int32).pin_memory(),
         n_reject=torch.empty(Q, dtype=torch.int32).pin_memory(), ncall=torch.empty(Q, dtype=torch.int32).pin_memory())
h_np = {k: v.numpy() for k, v in h.items()}
d_u0 = torch.empty((Q, n), dtype=torch.float64, device='cuda')
d_out = {k: torch.empty_like(v, device='cuda') for k, v in h.items()}
scale = 0.25


def t(fn, k=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / k


def propose():
    starts = rng.integers(len(u_live), size=Q)
    ell = bound.random_ells(rng, Q)
    np.take(u_live, starts, axis=0, out=h_u0.numpy(), mode='clip')
    return ell


ell = propose()
chain = [0]


def host_call():
    ctx.set_pointer_mode(_lib.PTR_HOST)
    chain[0] += Q
    ops.rwalk_batch(mid, h_u0.numpy(), loglstar, scale, walks, 1, chain0=chain[0], ell=ell, ctx=ctx, out=h_np)


def dev_call():
    ctx.set_pointer_mode(_lib.PTR_DEVICE)
    chain[0] += Q
    ops.rwalk_batch(mid, d_u0, loglstar, scale, walks, 1, chain0=chain[0], ell=ell, ctx=ctx, out=d_out)
    ctx.synchronize()


def copies_only():
    d_u0.copy_(h_u0, non_blocking=True)
    for k in h:
        h[k].copy_(d_out[k], non_blocking=True)
    st.synchronize()


d_u0.copy_(h_u0)
print('propose + np.take           us', round(t(propose), 1))
print('C-ABI call, pinned host     us', round(t(host_call), 1))
print('C-ABI call, device ptr+sync us', round(t(dev_call), 1))
ctx.set_timing(True)
dev_call()
print('kernel                      us', round(1e3 * ctx.last_kernel_ms(), 1))
ctx.set_timing(False)
print('7 torch copies + sync       us', round(t(copies_only), 1))
