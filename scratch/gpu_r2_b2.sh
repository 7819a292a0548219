#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/b2
timeout 200 python bench.py --steps 3 --warmup 3 > gpurun_out/b2/bench.json 2> gpurun_out/b2/err.log; python -c "
import json; d=json.loads(open('gpurun_out/b2/bench.json').readline()); print(round(d['value'],1), d['parity_checked'], d['roofline']['binding'])" || tail -5 gpurun_out/b2/err.log
