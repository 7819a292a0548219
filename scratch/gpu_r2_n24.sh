#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/n24
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/n24/n$N.err | grep "^{" > gpurun_out/n24/r02_bench_n${N}_hd_ss.json
python -c "
import json; d=json.loads(open('gpurun_out/n24/r02_bench_n${N}_hd_ss.json').readline()); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],1))"
