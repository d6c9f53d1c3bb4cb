"""C4 diagnostics: where does the logZ of the 200-D run go wrong?  solo runs, device vs host bound update, device vs
host unit-cube phase."""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
from dynesty_b200 import likelihoods as DL, nested, _lib

m = DL.iid_normal_ppf(200)
ctx = _lib.Context(0)
for tag, kw, dev_bound in [('dev_init+dev_bound', dict(), True), ('dev_init+host_bound', dict(), False),
                           ('host_init+dev_bound', dict(device_init=False), True),
                           ('host_init+host_bound', dict(device_init=False), False)]:
    for batch in (200,):
        t0 = time.perf_counter()
        s = nested.NestedSampler(m, nlive=8000, bound='single', sample='rwalk', walks=220, seed=11, ctx=ctx, queue_size=200)
        s.device_bound = dev_bound
        try:
            r = s.run_nested(loop='device', batch=batch, **kw)
            print(json.dumps(dict(tag=tag, batch=batch, logz=float(r.logz[-1]), niter=int(r.niter), ncall=int(r.ncall), nbound=s.nbound,
                                  wall=round(time.perf_counter() - t0, 2),
                                  bound_hist=[(int(a), int(b), round(c, 2)) for a, b, c in s.bound_history[:6]],
                                  scale_hist=[round(x[1], 4) for x in s.scale_history[::max(1, len(s.scale_history) // 12)]])), flush=True)
        except Exception as e:
            print(json.dumps(dict(tag=tag, batch=batch, error=repr(e)[:300], nbound=s.nbound,
                                  bound_hist=[(int(a), int(b), round(c, 2)) for a, b, c in s.bound_history[:6]])), flush=True)
