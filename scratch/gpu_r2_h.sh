#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/h_pytest.log 2>&1; tail -3 gpurun_out/h_pytest.log
for lib in main lutc0 acqpoll; do
  for a in "wave 450x375x60 ss" "wave 1242x375x128 cs" "ring 1920x1080x256 ss" "wave 1920x1080x256 ss" "wave 960x540x128 ss" "ring 960x540x128 ss" "ring 1920x1080x256 cs" "wave 1920x1080x256 cs"; do
    set -- $a
    echo "$lib $a: $(CSPM_SPATIAL=$1 CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_$lib.so timeout 300 python scratch/time_step.py $2 $3 2 2>&1 | grep rep1 | sed 's/.*rep1//')" >> gpurun_out/h_variants.log
  done
done
cat gpurun_out/h_variants.log
for sz in 450x375x60 1920x1080x256; do
  CSPM_SPATIAL=wave CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_waveprof.so timeout 300 python profiles/wave_phase_cycles.py $sz 2>&1 | grep sweep | cut -c1-330
done
