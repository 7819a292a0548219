#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/z
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/z/pytest.log 2>&1; tail -2 gpurun_out/z/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
timeout 600 python bench.py > gpurun_out/z/bench.json 2> gpurun_out/z/bench.err; python -c "
import json; d=json.loads(open('gpurun_out/z/bench.json').readline()); print(round(d['value'],1), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity_checked'], 'launches', d['gpu_launches'], 'onchip', bool(d['roofline']['onchip']), d['clocks'])"
