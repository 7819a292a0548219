#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spatial_persist -s 1 -c 1 -f -o gpurun_out/r02_ring_hd_src python profiles/run_one_step.py 1920x1080x256 > gpurun_out/t_ncu.log 2>&1
tail -2 gpurun_out/t_ncu.log; ls -la gpurun_out/*.ncu-rep
