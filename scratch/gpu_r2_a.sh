#!/bin/bash
# round 2, GPU call A: parity suite + wave-vs-ring stage timings
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
for mode in wave ring; do
  CSPM_SPATIAL=$mode CSPM_BENCH_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_bench_$mode.json 2> gpurun_out/a_bench_$mode.err
  for cfg in C1 C3; do
    CSPM_SPATIAL=$mode CSPM_BENCH_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config $cfg > gpurun_out/a_bench_${mode}_$cfg.json 2> gpurun_out/a_bench_${mode}_$cfg.err
  done
done
grep -h "stage_ms" gpurun_out/a_bench_*.err | cut -c1-400
