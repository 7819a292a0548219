#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/x_pytest.log 2>&1; tail -3 gpurun_out/x_pytest.log
for cfg in C1 C2 C3; do timeout 300 python bench.py --config $cfg --no-cpu-baseline --no-tap-count --no-worst-case --steps 3 --warmup 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$cfg', round(d['value'],1), round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['stage_ms'].items() if v>0.1})"; done
