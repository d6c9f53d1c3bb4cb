#!/bin/bash
# round 2: rwalk_mma16_kernel (16 warps per 8 chains) against the 8-warp kernel: parity tests, kernel time, ncu
set -u
O=gpurun_out
mkdir -p $O
export B2N_RWALK_WARPS=16
timeout 900 python -m pytest tests/test_gpu_rwalk.py tests/test_gpu_nsloop.py tests/test_gpu_uniformity.py -x -q -m gpu > $O/r2n_pytest_w16.log 2>&1
echo "pytest w16 rc=$?" >> $O/r2n_pytest_w16.log
for W in 8 16; do
  B2N_RWALK_WARPS=$W timeout 600 python bench.py --steps 30 --warmup 5 --ensemble 0 --cpu-baseline 0 > $O/r2n_bench_w$W.json 2> $O/r2n_bench_w$W.err
done
B2N_RWALK_WARPS=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rwalk_mma16 -s 6 -c 1 -o $O/r2n_mma16 -f python bench.py --steps 2 --warmup 1 --ensemble 0 --cpu-baseline 0 > $O/r2n_ncu.log 2>&1
ncu -i $O/r2n_mma16.ncu-rep --page raw --csv > $O/r2n_mma16_raw.csv 2>/dev/null
