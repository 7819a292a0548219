#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu -k "C1 or C2 or C3 or stages or sweep_impl or full_patchmatch or other_window" > gpurun_out/i_pytest.log 2>&1; tail -3 gpurun_out/i_pytest.log
for lib in main u4 t640u5 t640u3; do
  for a in "450x375x60 ss" "450x375x60 cs" "1242x375x128 cs" "960x540x128 ss" "1920x1080x256 ss"; do
    set -- $a
    echo "$lib wave $a: $(CSPM_SPATIAL=wave CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_$lib.so timeout 300 python scratch/time_step.py $1 $2 2 2>&1 | grep rep1 | sed 's/.*rep1//')" >> gpurun_out/i_variants.log
  done
done
cat gpurun_out/i_variants.log
for sz in 450x375x60 1920x1080x256; do
  CSPM_SPATIAL=wave CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_waveprof.so timeout 300 python profiles/wave_phase_cycles.py $sz 2>&1 | grep sweep | cut -c1-330
done
