#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; tail -4 gpurun_out/f_pytest.log
CSPM_BENCH_DEBUG=1 timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; tail -3 gpurun_out/f_bench.err | cut -c1-300; cat gpurun_out/f_bench.json | cut -c1-3000
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; cat gpurun_out/f_bench_ref.json | cut -c1-800
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
