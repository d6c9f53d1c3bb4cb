"""Chain-kernel capacity under concurrency: S contexts (streams), each launching rounds of K = 50 one-chain-per-CTA
rwalk chains back to back (device pointers, no step kernels, no bound updates).  Aggregate proposals/s vs S and packing."""
import json, sys, threading, time
sys.path.insert(0, '.')
import numpy as np, torch
from dynesty_b200 import _lib, ops, likelihoods as DL, bounding as B
from bench import make_state
n, K, walks = 50, 50, 70
u, loglstar = make_state(n, 2000)
m = DL.gauss_corr(n, 0.4, 5.0)
b = B.B200MultiEllipsoid(n)
b.update(u, rstate=np.random.default_rng(1))
b.scale_to_logvol(b.logvol + np.log(1.25))
def run(S, pack, reps=300):
    ctxs = [_lib.Context(0) for _ in range(S)]
    bufs = []
    for c in ctxs:
        if pack > 1:
            c.set_chain_pack(pack)
        b.make_resident(c)
        d_u0 = torch.from_numpy(u[:K].copy()).cuda()
        out = dict(u=torch.empty(K, n, dtype=torch.float64, device='cuda'), v=torch.empty(K, n, dtype=torch.float64, device='cuda'),
                   logl=torch.empty(K, dtype=torch.float64, device='cuda'), n_accept=torch.empty(K, dtype=torch.int32, device='cuda'),
                   n_reject=torch.empty(K, dtype=torch.int32, device='cuda'), ncall=torch.empty(K, dtype=torch.int32, device='cuda'))
        bufs.append((d_u0, out, m.model_id(c)))
        c.set_pointer_mode(_lib.PTR_DEVICE)
    torch.cuda.synchronize()
    def work(i):
        c, (d_u0, out, mid) = ctxs[i], bufs[i]
        for r in range(reps):
            ops.rwalk_batch(mid, d_u0, loglstar, 0.2, walks, 7, chain0=r * K, ctx=c, out=out)
        c.synchronize()
    for i in range(S):
        work_i = threading.Thread(target=work, args=(i,))
    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(S)]
    [t.start() for t in ts]; [t.join() for t in ts]
    wall = time.perf_counter() - t0
    for c in ctxs:
        c.set_pointer_mode(_lib.PTR_HOST); c.close()
    return dict(streams=S, pack=pack, rounds_per_s=round(S * reps / wall), proposals_per_s=round(S * reps * K * walks / wall),
                us_per_round_per_stream=round(1e6 * wall / reps, 1))
run(1, 1, 50)
for S, pack in [(1, 1), (4, 1), (8, 1), (16, 1), (32, 1), (8, 4), (16, 4), (32, 4), (32, 8)]:
    print(json.dumps(run(S, pack)), flush=True)
