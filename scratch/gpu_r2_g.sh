#!/bin/bash
# round 2 profiles: launch lists + ncu --set full of the dominant kernels at the benched kernel version
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N="ncu --clock-control none"
timeout 600 $N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_hd_ss.csv python profiles/run_one_step.py 1920x1080x256 > gpurun_out/g1.log 2>&1
timeout 600 $N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_c1.csv python profiles/run_one_step.py 450x375x60 > gpurun_out/g2.log 2>&1
timeout 900 $N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_hd_cs3.csv python profiles/run_one_step.py 1920x1080x256 cs > gpurun_out/g3.log 2>&1
timeout 600 $N --set full --import-source on -k regex:k_refine -s 1 -c 1 -f -o gpurun_out/r02_refine_hd_ss python profiles/run_one_step.py 1920x1080x256 > gpurun_out/g4.log 2>&1
timeout 900 $N --set full --import-source on -k regex:k_refine -s 1 -c 1 -f -o gpurun_out/r02_refine_hd_cs3 python profiles/run_one_step.py 1920x1080x256 cs > gpurun_out/g5.log 2>&1
timeout 900 $N --set full --import-source on -k regex:k_spatial_persist -s 1 -c 1 -f -o gpurun_out/r02_spatial_ring_hd_cs3 python profiles/run_one_step.py 1920x1080x256 cs > gpurun_out/g6.log 2>&1
timeout 600 $N --set full --import-source on -k regex:k_spatial_persist -s 1 -c 1 -f -o gpurun_out/r02_spatial_ring_hd_ss python profiles/run_one_step.py 1920x1080x256 > gpurun_out/g7.log 2>&1
timeout 600 $N --set full --import-source on -k regex:k_spatial_wave -s 1 -c 1 -f -o gpurun_out/r02_spatial_wave_c1 python profiles/run_one_step.py 450x375x60 > gpurun_out/g8.log 2>&1
timeout 600 $N --set full --import-source on -k regex:k_spatial_wave -s 1 -c 1 -f -o gpurun_out/r02_spatial_wave_c3 python profiles/run_one_step.py 1242x375x128 cs > gpurun_out/g9.log 2>&1
ls -la gpurun_out/r02_*; tail -2 gpurun_out/g4.log
