#!/bin/bash
# round 2, GPU session B: diagnostics (C4, improve_covar), new tests (friends, fixed ones), replica throughput scan
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python scripts/icm_debug.py > $O/r2b_icm.log 2>&1
timeout 600 python scripts/c4_debug.py > $O/r2b_c4_debug.log 2>&1
timeout 900 python -m pytest tests/test_gpu_friends.py tests/test_gpu_uniformity.py tests/test_gpu_nsloop.py tests/test_gpu_bounding.py tests/test_gpu_replicas.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2b_pytest.log
timeout 600 python scripts/replica_scan.py $O/r2b_replica_scan.jsonl > $O/r2b_replica_scan.log 2>&1
