#!/bin/bash
# round 2: warp-specialised rwalk kernel (8 step + 4 draw warps) against the 8-warp and 16-warp kernels
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_rwalk.py tests/test_gpu_nsloop.py tests/test_gpu_uniformity.py -x -q -m gpu > $O/r2q_pytest_w12.log 2>&1
echo "pytest w12 rc=$?" >> $O/r2q_pytest_w12.log
for W in 8 12; do
  B2N_RWALK_WARPS=$W timeout 300 python bench.py --steps 30 --warmup 5 --ensemble 0 --cpu-baseline 0 > $O/r2q_bench_w$W.json 2> $O/r2q_bench_w$W.err
  echo "bench w$W rc=$?" >> $O/r2q_bench_w$W.err
done
B2N_RWALK_WARPS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rwalk_mmaws -s 6 -c 1 -o $O/r2q_mmaws -f python bench.py --steps 2 --warmup 1 --ensemble 0 --cpu-baseline 0 > $O/r2q_ncu.log 2>&1
ncu -i $O/r2q_mmaws.ncu-rep --page raw --csv > $O/r2q_mmaws_raw.csv 2>/dev/null
ncu -i $O/r2q_mmaws.ncu-rep --page source --csv > $O/r2q_mmaws_src.csv 2>/dev/null
