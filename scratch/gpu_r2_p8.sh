#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/p8
for P in 6 8; do timeout 200 python bench.py --config C1 --pairs-in-flight $P --no-cpu-baseline --no-tap-count --steps 2 --warmup 1 > gpurun_out/p8/cfg_C1_gpu_batch$P.json 2>/dev/null; python -c "import json; d=json.loads(open('gpurun_out/p8/cfg_C1_gpu_batch$P.json').readline()); print($P, round(d['value'],1), {k: round(v,1) for k,v in d['stage_ms'].items() if v>0.1})"; done
