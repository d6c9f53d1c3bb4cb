"""Ensemble throughput under different host / driver settings (env): quick scan of replicas in flight."""
import json, os, sys, time
sys.path.insert(0, '.')
from dynesty_b200 import likelihoods as DL, replicas
m = DL.gauss_corr(50, 0.4, 5.0)
kw = dict(nlive=2000, bound='multi', sample='rwalk', sampler_kwargs=dict(walks=70), batch=50)
tag = dict(maxconn=os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS'), blocking=os.environ.get('B2N_BLOCKING_SYNC'),
           graph=os.environ.get('B2N_NS_GRAPH'))
replicas.run_replicas(m, range(4), max_in_flight=4, **kw)
for inflight, pack, nrep in [(8, 1, 32), (16, 1, 48), (32, 1, 64), (32, 4, 64)]:
    t0 = time.perf_counter()
    outs, wall = replicas.run_replicas(m, range(100, 100 + nrep), max_in_flight=inflight, chain_pack=pack, **kw)
    wall = time.perf_counter() - t0
    print(json.dumps(dict(tag, in_flight=inflight, pack=pack, replicas=nrep, wall=round(wall, 3),
                          calls_per_s=round(sum(o['ncall'] for o in outs) / wall),
                          rounds_s=round(sum(o['rounds_s'] for o in outs) / nrep, 3), bound_s=round(sum(o['bound_s'] for o in outs) / nrep, 3))), flush=True)
