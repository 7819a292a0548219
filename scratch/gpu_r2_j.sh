#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spatial_wave -s 1 -c 1 -f -o gpurun_out/r02_wave_c1_v4 python profiles/run_one_step.py 450x375x60 > gpurun_out/j_ncu.log 2>&1
tail -2 gpurun_out/j_ncu.log; ls -la gpurun_out/
