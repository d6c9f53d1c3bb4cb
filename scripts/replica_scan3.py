"""Ensemble throughput: replicas in flight x chains per CTA x step-kernel threads."""
import json, os, sys, time
sys.path.insert(0, '.')
from dynesty_b200 import likelihoods as DL, replicas
m = DL.gauss_corr(50, 0.4, 5.0)
kw = dict(nlive=2000, bound='multi', sample='rwalk', sampler_kwargs=dict(walks=70), batch=50)
replicas.run_replicas(m, range(4), max_in_flight=4, **kw)
for T in (1024, 256):
    os.environ['B2N_NS_THREADS'] = str(T)
    for inflight, pack in [(16, 1), (16, 4), (32, 4), (48, 4), (48, 8)]:
        nrep = 2 * inflight
        t0 = time.perf_counter()
        outs, wall = replicas.run_replicas(m, range(100, 100 + nrep), max_in_flight=inflight, chain_pack=pack, **kw)
        wall = time.perf_counter() - t0
        print(json.dumps(dict(ns_threads=T, in_flight=inflight, pack=pack, replicas=nrep, wall=round(wall, 3),
                              calls_per_s=round(sum(o['ncall'] for o in outs) / wall),
                              rounds_s=round(sum(o['rounds_s'] for o in outs) / nrep, 3), bound_s=round(sum(o['bound_s'] for o in outs) / nrep, 3))), flush=True)
