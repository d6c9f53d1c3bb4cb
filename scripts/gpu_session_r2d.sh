#!/bin/bash
# round 2, GPU session D: host/driver settings for the replica ensemble, sanitizer sweep, C4 test
set -u
O=gpurun_out
mkdir -p $O
nproc > $O/r2d_nproc.txt
: > $O/r2d_scan2.jsonl
timeout 200 python scripts/replica_scan2.py >> $O/r2d_scan2.jsonl 2>> $O/r2d_scan2.err
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 200 python scripts/replica_scan2.py >> $O/r2d_scan2.jsonl 2>> $O/r2d_scan2.err
B2N_BLOCKING_SYNC=1 timeout 200 python scripts/replica_scan2.py >> $O/r2d_scan2.jsonl 2>> $O/r2d_scan2.err
CUDA_DEVICE_MAX_CONNECTIONS=32 B2N_BLOCKING_SYNC=1 timeout 200 python scripts/replica_scan2.py >> $O/r2d_scan2.jsonl 2>> $O/r2d_scan2.err
timeout 300 python -m pytest tests/test_gpu_fullrun.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2d_pytest.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_targets.py > $O/r2d_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/r2d_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_targets.py > $O/r2d_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/r2d_sanitizer_racecheck.log
