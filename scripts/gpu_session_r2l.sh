#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2l_pytest.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2l_bench_ref.json 2> $O/r2l_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2l_bench.json 2> $O/r2l_bench.err
echo "bench rc=$?" >> $O/r2l_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2l_smoke.log 2>&1
