#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for lib in main wbf2 wbf4 wbf5t640 wbf3t640; do
  for a in "450x375x60 ss" "450x375x60 cs" "1242x375x128 cs"; do
    set -- $a
    echo "$lib $a: $(CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_$lib.so timeout 120 python scratch/time_step.py $1 $2 2 2>&1 | grep rep1 | sed 's/.*rep1//')" >> gpurun_out/w_variants.log
  done
done
cat gpurun_out/w_variants.log
CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_wbf2.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu -k "stages or full_patchmatch or shortcuts or C1 or C2 or batch" 2>&1 | tail -2
