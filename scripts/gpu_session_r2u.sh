#!/bin/bash
# round 2, 8-GPU session with the final kernels: the contract bench at N = 8 and N = 4 (ensemble of 512 full runs)
set -u
O=gpurun_out
mkdir -p $O
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 > $O/r2u_bench_${N}gpu.json 2> $O/r2u_bench_${N}gpu.err
  echo "bench N=$N rc=$?" >> $O/r2u_bench_${N}gpu.err
done
