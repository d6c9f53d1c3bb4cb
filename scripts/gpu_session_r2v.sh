#!/bin/bash
# round 2 final validation on one GPU: full GPU suite, both bench arms (ensemble of 512 runs), ncu launch list of the
# bench command, smoke
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2v_pytest.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2v_bench_ref.json 2> $O/r2v_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2v_bench.json 2> $O/r2v_bench.err
echo "bench rc=$?" >> $O/r2v_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2v_launches.csv python bench.py --steps 2 --warmup 1 --ensemble 2 --in-flight 1 --chain-pack 1 --cpu-baseline 0 > $O/r2v_ncu_launches.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2v_smoke.log 2>&1
echo "smoke rc=$?" >> $O/r2v_smoke.log
