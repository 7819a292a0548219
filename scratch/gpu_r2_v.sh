#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for lib in main guard ilp1t512; do
  for a in "1920x1080x256 ss" "1920x1080x256 cs"; do
    set -- $a
    echo "$lib ring $a: $(CSPM_SPATIAL=ring CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_$lib.so timeout 120 python scratch/time_step.py $1 $2 2 2>&1 | grep rep1 | sed 's/.*rep1//')" >> gpurun_out/v_variants.log
  done
done
cat gpurun_out/v_variants.log
