"""Throughput of an ensemble of full C2 runs vs replicas in flight and chains per CTA (one GPU)."""
import json, sys, time
sys.path.insert(0, '.')
from dynesty_b200 import likelihoods as DL, replicas
m = DL.gauss_corr(50, 0.4, 5.0)
out = open(sys.argv[1], 'a') if len(sys.argv) > 1 else sys.stdout
kw = dict(nlive=2000, bound='multi', sample='rwalk', sampler_kwargs=dict(walks=70), batch=50)
replicas.run_replicas(m, range(4), max_in_flight=4, **kw)          # warm-up
for inflight, pack, nrep in [(1, 1, 4), (8, 1, 32), (16, 1, 48), (16, 2, 48), (32, 2, 64), (32, 4, 64), (48, 4, 96), (64, 8, 128)]:
    t0 = time.perf_counter()
    outs, wall = replicas.run_replicas(m, range(100, 100 + nrep), max_in_flight=inflight, chain_pack=pack, **kw)
    wall = time.perf_counter() - t0
    s = replicas.summarize(outs, wall)
    rec = dict(in_flight=inflight, chain_pack=pack, replicas=nrep, wall_s=round(wall, 3), calls_per_s=round(s['calls_per_s']),
               logz_mean=round(s['logz_mean'], 3), logz_std=round(s['logz_std'], 3), run_wall_mean=round(s['run_wall_s_mean'], 3),
               rounds_s_mean=round(sum(o['rounds_s'] for o in outs) / nrep, 3), bound_s_mean=round(sum(o['bound_s'] for o in outs) / nrep, 3))
    out.write(json.dumps(rec) + '\n')
    out.flush()
