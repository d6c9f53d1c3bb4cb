#!/bin/bash
# round 2, session zg: start points by index (b2n_set_start_rows): parity test, e2e with and without the caller's gather
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_rwalk.py tests/test_gpu_peer.py -q --timeout 300 -p no:cacheprovider > $O/r2zg_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2zg_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --ensemble 0 --cpu-baseline 0 --gather-starts 1 > $O/r2zg_bench_gather.json 2> $O/r2zg_bench_gather.err
timeout 300 python bench.py --steps 20 --warmup 5 --ensemble 0 --cpu-baseline 0 > $O/r2zg_bench_index.json 2> $O/r2zg_bench_index.err
tail -n 4 $O/r2zg_pytest.log
python - <<'PY'
import json
for f in ('gather', 'index'):
    try:
        d = json.loads([l for l in open('gpurun_out/r2zg_bench_%s.json' % f) if l.startswith('{')][0])
        print(f, 'value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), d['e2e'].get('start_points'))
    except Exception as e:
        print(f, 'failed', e)
PY
tail -n 3 $O/r2zg_bench_index.err
