#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for lib in main ldg; do
  for a in "1920x1080x256 ss" "1920x1080x256 cs" "450x375x60 ss"; do
    set -- $a
    echo "$lib $a: $(CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_$lib.so timeout 120 python scratch/time_step.py $1 $2 2 2>&1 | grep rep1 | sed 's/.*rep1//')" >> gpurun_out/y_variants.log
  done
done
cat gpurun_out/y_variants.log
