#!/bin/bash
# round 2, session ze: sanitizer sweep over the code paths added in the second half of the round; C4 timing with the
# slab order changed
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python scripts/sanitize_targets2.py > $O/r2ze_targets_plain.log 2>&1
echo "plain rc=$?" >> $O/r2ze_targets_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_targets2.py > $O/r2ze_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/r2ze_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_targets2.py > $O/r2ze_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/r2ze_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python scripts/sanitize_targets2.py > $O/r2ze_sanitizer_synccheck.log 2>&1
echo "synccheck rc=$?" >> $O/r2ze_sanitizer_synccheck.log
timeout 300 python scripts/c4_timing.py > $O/r2ze_c4_timing.jsonl 2> $O/r2ze_c4_timing.err
tail -n 4 $O/r2ze_targets_plain.log $O/r2ze_sanitizer_memcheck.log $O/r2ze_sanitizer_racecheck.log $O/r2ze_sanitizer_synccheck.log
cat $O/r2ze_c4_timing.jsonl
