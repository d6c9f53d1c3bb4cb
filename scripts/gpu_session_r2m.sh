#!/bin/bash
# round 2, 8-GPU session: the contract bench at N = 8 and N = 4
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r2m_topo.txt 2>&1
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > $O/r2m_bench_${N}gpu.json 2> $O/r2m_bench_${N}gpu.err
  echo "bench N=$N rc=$?" >> $O/r2m_bench_${N}gpu.err
done
