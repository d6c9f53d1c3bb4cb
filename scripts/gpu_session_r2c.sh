#!/bin/bash
# round 2, GPU session C: full GPU test tier after the fixes, C4 sweep with the new tuning, replica scan with graphs, bench
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2c_pytest.log
timeout 200 python scripts/icm_debug.py > $O/r2c_icm.log 2>&1
timeout 600 python scripts/c4_batch_sweep.py $O/r2c_c4_sweep.jsonl > $O/r2c_c4_sweep.log 2>&1
timeout 600 python scripts/replica_scan.py $O/r2c_replica_scan.jsonl > $O/r2c_replica_scan.log 2>&1
B2N_NS_GRAPH=0 timeout 300 python - > $O/r2c_nograph.log 2>&1 <<'PY'
import sys, time, json
sys.path.insert(0, '.')
from dynesty_b200 import likelihoods as DL, replicas
m = DL.gauss_corr(50, 0.4, 5.0)
kw = dict(nlive=2000, bound='multi', sample='rwalk', sampler_kwargs=dict(walks=70), batch=50)
replicas.run_replicas(m, range(4), max_in_flight=4, **kw)
for inflight, nrep in [(1, 4), (16, 48)]:
    outs, wall = replicas.run_replicas(m, range(100, 100 + nrep), max_in_flight=inflight, **kw)
    print(json.dumps(dict(graph=0, in_flight=inflight, replicas=nrep, wall=round(wall, 3), calls_per_s=round(sum(o['ncall'] for o in outs) / wall))))
PY
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2c_bench.json 2> $O/r2c_bench.err
echo "bench rc=$?" >> $O/r2c_bench.err
