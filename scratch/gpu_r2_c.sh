#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu -k "config or stages or sweep_impl or redblack or full_patchmatch or error" > gpurun_out/d_pytest.log 2>&1; tail -4 gpurun_out/d_pytest.log
for sz in 450x375x60 1920x1080x256; do
  CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_waveprof.so timeout 300 python profiles/wave_phase_cycles.py $sz >> gpurun_out/d_phase.log 2>&1
done
CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_waveprof.so timeout 300 python profiles/wave_phase_cycles.py 1242x375x128 cs >> gpurun_out/d_phase.log 2>&1
cat gpurun_out/d_phase.log
for so in scratch/variants_r2/libcspm_t*.so; do
  for a in "450x375x60 ss" "1920x1080x256 ss" "1242x375x128 cs"; do
    CSPM_LIB_PATH=$PWD/$so timeout 300 python scratch/time_step.py $a 2 2>&1 | grep rep1 >> gpurun_out/d_variants.log
  done
done
CSPM_SPATIAL=ring timeout 300 python scratch/time_step.py 1920x1080x256 ss 2 2>&1 | grep rep1 >> gpurun_out/d_variants.log
cat gpurun_out/d_variants.log
