#!/bin/bash
# round 2, session zh: final one-GPU validation of the round: GPU suite, per-row measurements, both bench arms (ensemble
# of 512 runs), ncu launch list of the bench command, smoke
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2zh_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2zh_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2zh_pytest_gpu.log
timeout 600 python scripts/row_bench.py > $O/r2zh_rows.jsonl 2> $O/r2zh_rows.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2zh_bench_ref.json 2> $O/r2zh_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2zh_bench.json 2> $O/r2zh_bench.err
echo "bench rc=$?" >> $O/r2zh_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2zh_launches.csv python bench.py --steps 2 --warmup 1 --ensemble 2 --in-flight 1 --chain-pack 1 --cpu-baseline 0 > $O/r2zh_ncu_launches.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2zh_smoke.log 2>&1
echo "smoke rc=$?" >> $O/r2zh_smoke.log
tail -n 3 $O/r2zh_pytest_gpu.log $O/r2zh_smoke.log
tail -n 2 $O/r2zh_bench.err
