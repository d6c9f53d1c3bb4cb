#!/bin/bash
# round 2, session zj (last GPU seconds of the round): the GPU suite on the final tree; then, if time is left, one run
# alone per config with its rounds / bound-update split
set -u
O=gpurun_out
mkdir -p $O
timeout 80 python -m pytest tests -m gpu -q -x --timeout 60 -p no:cacheprovider > $O/r2zj_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2zj_pytest_gpu.log
tail -n 3 $O/r2zj_pytest_gpu.log
timeout 60 python scripts/config_runs_timing.py > $O/r2zj_config_runs.jsonl 2> $O/r2zj_config_runs.err
cat $O/r2zj_config_runs.jsonl
