#!/bin/bash
# round 2: ring of 4 buffers x 4 steps with one pipeline over all slots (B2N_RWALK_RING=4) against 2 x 8
set -u
O=gpurun_out
mkdir -p $O
B2N_RWALK_RING=4 timeout 600 python -m pytest tests/test_gpu_rwalk.py tests/test_gpu_nsloop.py -x -q -m gpu > $O/r2t_pytest_ring4.log 2>&1
echo "pytest rc=$?" >> $O/r2t_pytest_ring4.log
for R in 8 4 8 4; do
  B2N_RWALK_RING=$R timeout 300 python bench.py --steps 30 --warmup 5 --ensemble 0 --cpu-baseline 0 >> $O/r2t_bench_ring$R.json 2>> $O/r2t_bench_ring$R.err
done
B2N_RWALK_RING=4 timeout 600 python bench.py --steps 20 --warmup 5 --ensemble 256 --cpu-baseline 0 > $O/r2t_bench_ring4_ens.json 2> $O/r2t_bench_ring4_ens.err
B2N_RWALK_RING=4 timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 1 python -m pytest tests/test_gpu_rwalk.py -x -q -m gpu -k "variants_agree and 32-8 or mma_kernel_golden and g50" > $O/r2t_sanitizer_racecheck_ring4.log 2>&1
echo "racecheck rc=$?" >> $O/r2t_sanitizer_racecheck_ring4.log
