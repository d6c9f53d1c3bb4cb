#!/bin/bash
# round 2, GPU session E (2 GPUs): the multi-GPU tests and the contract bench at N = 2, both arms
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r2e_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_peer.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2e_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2e_bench_2gpu.json 2> $O/r2e_bench_2gpu.err
echo "bench rc=$?" >> $O/r2e_bench_2gpu.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-baseline 0 > $O/r2e_bench_1gpu.json 2> $O/r2e_bench_1gpu.err
