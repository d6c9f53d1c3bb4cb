#!/bin/bash
# round 2, session zb: Jacobi on shared-memory addresses; then the complete one-GPU validation: GPU suite, A/B of the
# update, per-row measurements, both bench arms (ensemble of 512 runs), ncu launch list of the bench command, smoke
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2zb_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2zb_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2zb_pytest_gpu.log
timeout 300 python scripts/bound_ab.py > $O/r2zb_bound_ab.jsonl 2> $O/r2zb_bound_ab.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2zb_update_launches.csv \
    python scripts/one_update.py > $O/r2zb_ncu_launch.log 2>&1
timeout 600 python scripts/row_bench.py > $O/r2zb_rows.jsonl 2> $O/r2zb_rows.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2zb_bench_ref.json 2> $O/r2zb_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2zb_bench.json 2> $O/r2zb_bench.err
echo "bench rc=$?" >> $O/r2zb_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2zb_launches.csv python bench.py --steps 2 --warmup 1 --ensemble 2 --in-flight 1 --chain-pack 1 --cpu-baseline 0 > $O/r2zb_ncu_launches.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2zb_smoke.log 2>&1
echo "smoke rc=$?" >> $O/r2zb_smoke.log
tail -n 3 $O/r2zb_pytest_gpu.log $O/r2zb_smoke.log
grep -h '"all o' $O/r2zb_bound_ab.jsonl | cut -c1-260
