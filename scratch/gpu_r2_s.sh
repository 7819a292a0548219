#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for lib in main col608u1 col608u2 col608u3 col608u5 col512u4; do
  for a in "1920x1080x256 ss" "1920x1080x256 cs" "960x540x128 ss"; do
    set -- $a
    echo "$lib ring $a: $(CSPM_SPATIAL=ring CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_$lib.so timeout 300 python scratch/time_step.py $1 $2 2 2>&1 | grep rep1 | sed 's/.*rep1//')" >> gpurun_out/s_variants.log
  done
done
cat gpurun_out/s_variants.log
CSPM_SPATIAL=ring CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_col608u2.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stages or full_patchmatch or shortcuts" 2>&1 | tail -2
