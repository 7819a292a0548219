#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-100
timeout 300 python -m pytest tests/test_gpu_configs.py tests/test_gpu_vs_reference_golden.py -x -q -m gpu -k "C1 or C3 or C4 or golden" 2>&1 | tail -2
