#!/bin/bash
# round 2, GPU session A: GPU test tier, contract bench, C4 batch sweep, launch list + ncu --set full of the kernels
# that had no summary yet.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2a_bench.json 2> $O/r2a_bench.err
echo "bench rc=$?" >> $O/r2a_bench.err
timeout 400 python scripts/c4_batch_sweep.py $O/r2a_c4_sweep.jsonl > $O/r2a_c4_sweep.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2a_launches.csv \
    python scripts/ncu_targets.py > $O/r2a_ncu_launch.log 2>&1
for k in slice_kernel unif_kernel kmeans2_kernel eig_ladder_kernel chol_node_kernel ns_step_kernel unitcube_kernel; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $O/r2a_$k \
      python scripts/ncu_targets.py > $O/r2a_ncu_$k.log 2>&1
done
ls -la $O > $O/r2a_ls.txt
