#!/bin/bash
# round 2, session zi (2 GPUs): the contract bench at N = 2 with the start points by index in the fused-exchange path
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-baseline 0 --ensemble 64 > $O/r2zi_bench_2gpu.json 2> $O/r2zi_bench_2gpu.err
echo "bench rc=$?" >> $O/r2zi_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --impl reference > $O/r2zi_bench_ref_2gpu.json 2> $O/r2zi_bench_ref_2gpu.err
echo "ref rc=$?" >> $O/r2zi_bench_ref_2gpu.err
cut -c1-600 $O/r2zi_bench_2gpu.json
tail -n 3 $O/r2zi_bench_2gpu.err $O/r2zi_bench_ref_2gpu.err
cut -c1-300 $O/r2zi_bench_ref_2gpu.json
