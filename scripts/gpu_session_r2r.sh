#!/bin/bash
# round 2: rwalk_mmaws_kernel with one pipeline over all ring slots; sanitizers on the lock-step variants
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_rwalk.py tests/test_gpu_nsloop.py tests/test_gpu_uniformity.py -x -q -m gpu > $O/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2r_pytest.log
for W in 8 12; do
  B2N_RWALK_WARPS=$W timeout 300 python bench.py --steps 30 --warmup 5 --ensemble 0 --cpu-baseline 0 > $O/r2r_bench_w$W.json 2> $O/r2r_bench_w$W.err
  echo "bench w$W rc=$?" >> $O/r2r_bench_w$W.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rwalk_mmaws -s 6 -c 1 -o $O/r2r_mmaws -f python bench.py --steps 2 --warmup 1 --ensemble 0 --cpu-baseline 0 > $O/r2r_ncu.log 2>&1
ncu -i $O/r2r_mmaws.ncu-rep --page raw --csv > $O/r2r_mmaws_raw.csv 2>/dev/null
ncu -i $O/r2r_mmaws.ncu-rep --page source --csv > $O/r2r_mmaws_src.csv 2>/dev/null
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_rwalk.py -x -q -m gpu -k "variants_agree or mma_kernel_golden" > $O/r2r_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/r2r_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 1 python -m pytest tests/test_gpu_rwalk.py -x -q -m gpu -k "variants_agree and 32-8 or mma_kernel_golden and g50" > $O/r2r_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/r2r_sanitizer_racecheck.log
