#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/n8
nvidia-smi -L | head -8 > gpurun_out/n8/gpus.txt
CSPM_BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/n8/r02_bench_n8_hd_ss_v2.json 2> gpurun_out/n8/n8.err
tail -3 gpurun_out/n8/n8.err | cut -c1-300; cut -c1-700 gpurun_out/n8/r02_bench_n8_hd_ss_v2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/n8/r02_bench_n8_reference.json 2>> gpurun_out/n8/n8.err
cut -c1-300 gpurun_out/n8/r02_bench_n8_reference.json
