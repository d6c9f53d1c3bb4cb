#!/bin/bash
# round 2, session w: bound-update restructuring (speculative root fit, staged k-means rows, split candidate fit):
# parity tests, A/B timings, per-row measurements, a short bench line
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bounding.py -q --timeout 300 -p no:cacheprovider -x > $O/r2w_pytest_bounding.log 2>&1
echo "pytest rc=$?" >> $O/r2w_pytest_bounding.log
timeout 300 python scripts/bound_ab.py > $O/r2w_bound_ab.jsonl 2> $O/r2w_bound_ab.err
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2w_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2w_pytest_gpu.log
timeout 600 python scripts/row_bench.py > $O/r2w_rows.jsonl 2> $O/r2w_rows.err
timeout 300 python bench.py --steps 10 --warmup 3 --ensemble 8 --cpu-baseline 0 > $O/r2w_bench_short.json 2> $O/r2w_bench_short.err
tail -3 $O/r2w_pytest_bounding.log $O/r2w_pytest_gpu.log
cat $O/r2w_bound_ab.jsonl | cut -c1-400
tail -5 $O/r2w_bound_ab.err $O/r2w_rows.err
