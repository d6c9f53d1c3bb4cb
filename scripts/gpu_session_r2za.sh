#!/bin/bash
# round 2, session za: half-warp Jacobi pairs, k-means staging with prefetched row indices, parallel max in scale_finish; parity tests, A/B, launch list of an update, ncu captures
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bounding.py -q --timeout 300 -p no:cacheprovider > $O/r2za_pytest_bounding.log 2>&1
echo "pytest rc=$?" >> $O/r2za_pytest_bounding.log
timeout 300 python scripts/bound_ab.py > $O/r2za_bound_ab.jsonl 2> $O/r2za_bound_ab.err
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2za_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2za_pytest_gpu.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2za_update_launches.csv \
    python scripts/one_update.py > $O/r2za_ncu_launch.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:kmeans2_kernel -s 4 -c 1 -f -o $O/r2za_kmeans2_kernel \
    python scripts/one_update.py > $O/r2za_ncu_kmeans2_kernel.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:eig_ladder -s 1 -c 1 -f -o $O/r2za_eig_ladder_kernel \
    python scripts/one_update.py > $O/r2za_ncu_eig_ladder_kernel.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --ensemble 8 --cpu-baseline 0 > $O/r2za_bench_short.json 2> $O/r2za_bench_short.err
tail -n 3 $O/r2za_pytest_bounding.log $O/r2za_pytest_gpu.log
cut -c1-330 $O/r2za_bound_ab.jsonl
