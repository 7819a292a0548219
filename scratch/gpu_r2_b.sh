#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for sz in 450x375x60 1920x1080x256; do
  CSPM_LIB_PATH=$PWD/scratch/variants_r2/libcspm_waveprof.so timeout 300 python profiles/wave_phase_cycles.py $sz >> gpurun_out/b_phase.log 2>&1
done
cat gpurun_out/b_phase.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spatial_wave -c 1 -o gpurun_out/r02_wave_c1 -f python profiles/run_one_step.py 450x375x60 > gpurun_out/b_ncu_c1.log 2>&1
tail -3 gpurun_out/b_ncu_c1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spatial_wave -c 1 -o gpurun_out/r02_wave_hd -f python profiles/run_one_step.py 1920x1080x256 > gpurun_out/b_ncu_hd.log 2>&1
tail -3 gpurun_out/b_ncu_hd.log
