#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/p2; O=gpurun_out/p2
N="ncu --clock-control none"
ll() { timeout 900 $N --metrics gpu__time_duration.sum --csv --log-file /tmp/ll.csv python profiles/run_one_step.py $2 $3 $4 > /tmp/ll.log 2>&1; python profiles/summarize.py launches /tmp/ll.csv > $O/r02_launches_$1.txt; tail -1 /tmp/ll.log >> $O/r02_launches_$1.txt; }
ll hd_ss 1920x1080x256 ss
ll hd_cs3 1920x1080x256 cs
ll c1 450x375x60 ss
ll c3 1242x375x128 cs
kk() { timeout 900 $N --set full -k regex:$2 -s $3 -c 1 -f -o /tmp/$1 python profiles/run_one_step.py $4 $5 $6 > /tmp/kk.log 2>&1; python profiles/summarize.py kernel /tmp/$1.ncu-rep > $O/r02_ncu_$1.txt 2>&1; }
kk k_spatial_persist_hd_ss k_spatial_persist 1 1920x1080x256 ss
kk k_spatial_persist_hd_cs3 k_spatial_persist 1 1920x1080x256 cs
kk k_spatial_wave_c1 k_spatial_wave 1 450x375x60 ss
kk k_spatial_wave_c3 k_spatial_wave 1 1242x375x128 cs
head -4 $O/r02_launches_hd_ss.txt
