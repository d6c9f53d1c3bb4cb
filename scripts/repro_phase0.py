import json, sys
sys.path.insert(0, '.')
from dynesty_b200 import likelihoods as DL, replicas, ops, nested, _lib
m = DL.gauss_corr(50, 0.4, 5.0)
kw = dict(nlive=2000, bound='multi', sample='rwalk', sampler_kwargs=dict(walks=70), batch=50)
bad = []
for pack in (4, 1):
    for s0 in range(100, 196, 16):
        try:
            outs, wall = replicas.run_replicas(m, range(s0, s0 + 16), max_in_flight=16, chain_pack=pack, **kw)
        except Exception as e:
            print('FAIL block', pack, s0, repr(e)[:200], flush=True)
            for s in range(s0, s0 + 16):
                try:
                    replicas.run_replicas(m, [s], max_in_flight=1, chain_pack=pack, **kw)
                except Exception as e2:
                    print('   solo FAIL seed', s, repr(e2)[:200], flush=True)
                    bad.append((pack, s))
print('bad', bad)
if bad:
    pack, s = bad[0]
    ctx = _lib.Context(0)
    ctx.set_chain_pack(pack)
    ns = nested.NestedSampler(m, nlive=2000, bound='multi', sample='rwalk', walks=70, seed=s, ctx=ctx)
    import numpy as np
    print('live logl finite', np.isfinite(ns.live_logl).all(), ns.live_logl.min(), ns.live_logl.max(), np.isnan(ns.live_u).any())
