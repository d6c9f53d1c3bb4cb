#!/bin/bash
# round 2, GPU session F: does the 1024-thread step kernel starve under replicas?  scan step-kernel threads x packing
set -u
O=gpurun_out
mkdir -p $O
: > $O/r2f_scan.jsonl
for T in 1024 512 256; do
  B2N_NS_THREADS=$T CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 200 python scripts/replica_scan2.py 2>> $O/r2f_scan.err | sed "s/^{/{\"ns_threads\": $T, /" >> $O/r2f_scan.jsonl
done
timeout 600 python -m pytest tests/test_gpu_nsloop.py tests/test_gpu_replicas.py tests/test_gpu_fullrun.py tests/test_gpu_rwalk.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2f_pytest.log
