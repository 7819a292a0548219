#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for mode in wave ring; do
  for a in "450x375x60 ss" "450x375x60 cs" "1242x375x128 cs" "1920x1080x256 ss" "1920x1080x256 cs" "960x540x128 ss" "3840x2160x256 ss"; do
    echo "$mode $a: $(CSPM_SPATIAL=$mode timeout 300 python scratch/time_step.py $a 2 2>&1 | grep rep1)" >> gpurun_out/e_modes.log
  done
done
cat gpurun_out/e_modes.log
