#!/bin/bash
cd "$GRAFT_REPO_ROOT"
CSPM_SPATIAL=ring CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_spprof.so timeout 300 python profiles/sp_phase_cycles.py 1920x1080x256 2>&1 | tail -2
