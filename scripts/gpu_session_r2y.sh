#!/bin/bash
# round 2, session y: sweep-based candidate fit, finer moment jobs; parity tests, A/B, launch list of an update, ncu captures
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bounding.py -q --timeout 300 -p no:cacheprovider > $O/r2y_pytest_bounding.log 2>&1
echo "pytest rc=$?" >> $O/r2y_pytest_bounding.log
timeout 300 python scripts/bound_ab.py > $O/r2y_bound_ab.jsonl 2> $O/r2y_bound_ab.err
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2y_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2y_pytest_gpu.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2y_update_launches.csv \
    python scripts/one_update.py > $O/r2y_ncu_launch.log 2>&1
for spec in "kmeans2_kernel 4" "chol_node_kernel<1> 5" "chol_node_kernel<2> 5" "cov_partial_kernel 5"; do
  set -- $spec
  tag=$(echo $1 | tr -d "<>")
  timeout 240 ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c 1 -f -o $O/r2y_$tag \
      python scripts/one_update.py > $O/r2y_ncu_$tag.log 2>&1
done
timeout 300 python bench.py --steps 10 --warmup 3 --ensemble 8 --cpu-baseline 0 > $O/r2y_bench_short.json 2> $O/r2y_bench_short.err
tail -n 3 $O/r2y_pytest_bounding.log $O/r2y_pytest_gpu.log
cut -c1-330 $O/r2y_bound_ab.jsonl
