#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 120 python scripts/launch_rate.py > $O/r2g_launch_rate.jsonl 2> $O/r2g_launch_rate.err
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python scripts/replica_scan2.py > $O/r2g_scan.jsonl 2> $O/r2g_scan.err
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > $O/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2g_pytest.log
