#!/bin/bash
# round 2, session zj (last GPU seconds of the round): one run alone per config, rounds / bound-update split
set -u
O=gpurun_out
mkdir -p $O
timeout 80 python scripts/config_runs_timing.py > $O/r2zj_config_runs.jsonl 2> $O/r2zj_config_runs.err
cat $O/r2zj_config_runs.jsonl
tail -n 3 $O/r2zj_config_runs.err
