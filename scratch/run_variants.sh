#!/bin/bash
# GPU box: time every scratch/variants/libcspm_*.so on the HD single-scale step; parity-test the ones listed in $PARITY
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/variants_$(date +%H%M%S).log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $out
for so in scratch/variants/libcspm_*.so; do
  CSPM_LIB_PATH=$PWD/$so timeout 300 python scratch/time_step.py ${SIZE:-1920x1080x256} ${MODE:-ss} ${REPS:-3} >> $out 2>&1 || echo "FAILED $so" >> $out
done
for v in $PARITY; do
  echo "== parity $v" >> $out
  CSPM_LIB_PATH=$PWD/scratch/variants/libcspm_$v.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 >> $out
done
cat $out
