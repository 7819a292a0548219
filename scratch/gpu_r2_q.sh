#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/q_pytest.log 2>&1; tail -3 gpurun_out/q_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scratch/gpu_r2_m.sh 2>&1 | tail -14
