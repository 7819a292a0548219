#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/zz; O=gpurun_out/zz
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 600 python bench.py --config C4cs > $O/cfg_C4cs_gpu.json 2> $O/err.log
for cfg in C2 C3; do timeout 300 python bench.py --config $cfg --no-cpu-baseline > $O/cfg_${cfg}_gpu_nocpu.json 2>> $O/err.log; done
for f in $O/*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').readline()); print(round(d['value'],1), round(d.get('ms_per_step',0),1),'ms', {k: round(v,1) for k,v in d['stage_ms'].items() if v>0.1}, 'parity', d.get('parity_checked'), 'wc', round((d.get('worst_case') or {}).get('value',0),1), 'fe', round(d['roofline'].get('frac_executed',0),2))")"; done
