#!/bin/bash
# round 2 validation with rwalk_mmaws_kernel as the default: full GPU suite, both bench arms, the ensemble with the
# 8-warp kernel for comparison, smoke
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2s_pytest.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2s_bench_ref.json 2> $O/r2s_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2s_bench.json 2> $O/r2s_bench.err
echo "bench rc=$?" >> $O/r2s_bench.err
B2N_RWALK_WARPS=8 timeout 900 python bench.py --steps 20 --warmup 5 --cpu-baseline 0 > $O/r2s_bench_warps8.json 2> $O/r2s_bench_warps8.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s_smoke.log 2>&1
echo "smoke rc=$?" >> $O/r2s_smoke.log
