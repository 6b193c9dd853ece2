#!/bin/bash
# usage: gpu_r02_multi.sh N  -- sharding-invariance check and bench at N GPUs of one box
N=$1
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r02_dist_check_n$N.txt 2>&1; echo "dist_check rc=$?" >> gpurun_out/r02_dist_check_n$N.txt
grep -v "^W\|^\[W\|Warning" gpurun_out/r02_dist_check_n$N.txt | tail -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n$N.err
