#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/k_pytest.log 2>&1; tail -3 gpurun_out/k_pytest.log
for lib in main coopcs0; do
  for a in "450x375x60 cs" "1242x375x128 cs" "1920x1080x256 cs" "1920x1080x256 ss"; do
    set -- $a
    echo "$lib $a: $(CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_$lib.so timeout 300 python scratch/time_step.py $1 $2 2 2>&1 | grep rep1 | sed 's/.*rep1//')" >> gpurun_out/k_variants.log
  done
done
cat gpurun_out/k_variants.log
for cfg in C1 C3; do for P in 1 2 4; do
  timeout 300 python bench.py --config $cfg --pairs-in-flight $P --steps 2 --warmup 1 --no-cpu-baseline --no-tap-count --no-worst-case 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$cfg P=$P', round(d['value'],1), 'Mpx*d/s', round(d['ms_per_step'],1), 'ms/step', {k: round(v,1) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['value'],1))"
done; done
