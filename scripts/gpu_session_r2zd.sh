#!/bin/bash
# round 2, session zd: rwalk_mmas_kernel, 25 fragment loads in flight in the one-tile contraction: parity tests at n > 64, C4 timings
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_rwalk.py -q --timeout 600 -p no:cacheprovider > $O/r2zd_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2zd_pytest.log
timeout 600 python scripts/c4_timing.py > $O/r2zd_c4_timing.jsonl 2> $O/r2zd_c4_timing.err
tail -n 5 $O/r2zd_pytest.log
cat $O/r2zd_c4_timing.jsonl
tail -n 3 $O/r2zd_c4_timing.err
