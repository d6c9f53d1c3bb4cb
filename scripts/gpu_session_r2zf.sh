#!/bin/bash
# round 2, session zf (2 GPUs): the multi-GPU tests with the restructured bound update / helper-warp chain kernel, and
# the contract bench at N = 2 (fused exchange, gathered-row check, the fixed ensemble = strong scaling point)
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r2zf_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_peer.py tests/test_gpu_replicas.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/r2zf_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2zf_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-baseline 0 > $O/r2zf_bench_2gpu.json 2> $O/r2zf_bench_2gpu.err
echo "bench rc=$?" >> $O/r2zf_bench_2gpu.err
tail -n 4 $O/r2zf_pytest_2gpu.log
cut -c1-400 $O/r2zf_bench_2gpu.json
tail -n 3 $O/r2zf_bench_2gpu.err
