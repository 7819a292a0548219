#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for lib in bycs16 bycs18 bycs24; do
  for a in "1242x375x128 cs" "1920x1080x256 cs"; do
    set -- $a
    echo "$lib $a: $(CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_$lib.so timeout 120 python scratch/time_step.py $1 $2 2 2>&1 | grep rep1 | sed 's/.*rep1//')" >> gpurun_out/by_variants.log
  done
done
cat gpurun_out/by_variants.log
