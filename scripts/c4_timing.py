"""C4 (200-D, single / rwalk 220 walks, nlive 8000) timings on one GPU: the chain kernel at the queue sizes of a
device round (200 chains = 2 per CTA) and of a whole-queue fill (8000 chains = 16 per CTA), and one run alone.
usage (GPU box): python scripts/c4_timing.py > gpurun_out/c4_timing.jsonl"""
import json
import sys
import time

sys.path.insert(0, '.')
sys.path.insert(0, 'scripts')
import numpy as np
from scipy.special import ndtr

from dynesty_b200 import _lib, ops, likelihoods as DL, bounding as B, nested
from bench_configs import ball_state

ctx = _lib.default_context()
rng = np.random.default_rng(56432)
m = DL.iid_normal_ppf(200)
u = ball_state(200, 8000, rng, lambda v: ndtr(v))
_, l = m.evaluate(u, ctx=ctx)
loglstar = float(l.min()) - 1e-9
b = B.B200Ellipsoid(200, ctx=ctx)
b.update(u, rstate=np.random.default_rng(1))
b.scale_to_logvol(b.logvol + np.log(1.25))
b.make_resident()
mid = m.model_id(ctx)
ctx.set_timing(True)
for Q in (200, 1000, 8000):
    kms = []
    for it in range(6):
        st = u[rng.integers(8000, size=Q)]
        o = ops.rwalk_batch(mid, st, loglstar, 0.1, 220, 56432, chain0=it * Q, ctx=ctx)
        if it >= 2:
            kms.append(ctx.last_kernel_ms())
    k = float(np.mean(kms))
    print(json.dumps(dict(what='rwalk_mmas_kernel 200-D', chains=Q, walks=220, kernel_ms=round(k, 4), us_per_step=round(1e3 * k / 220, 3),
                          calls_per_s=Q * 220 / (k * 1e-3), accept=float(o['n_accept'].mean() / 220))), flush=True)
ctx.set_timing(False)
t0 = time.perf_counter()
s = nested.NestedSampler(m, nlive=8000, bound='single', sample='rwalk', walks=220, seed=11, ctx=ctx)
r = s.run_nested(loop='device', batch=200)
wall = time.perf_counter() - t0
print(json.dumps(dict(what='one C4 run alone, batch 200', wall_s=round(wall, 3), logz=float(r.logz[-1]), niter=int(r.niter), ncall=int(s.ncall),
                      nbound=int(s.nbound), truth=m.logz_truth)), flush=True)
