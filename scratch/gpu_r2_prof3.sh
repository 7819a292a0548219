#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/p3; O=gpurun_out/p3
N="ncu --clock-control none"
timeout 600 $N --metrics gpu__time_duration.sum --csv --log-file /tmp/ll.csv python profiles/run_one_step.py 1920x1080x256 cs > /tmp/ll.log 2>&1; python profiles/summarize.py launches /tmp/ll.csv > $O/r02_launches_hd_cs3.txt; tail -1 /tmp/ll.log >> $O/r02_launches_hd_cs3.txt
timeout 600 $N --set full -k regex:k_refine -s 1 -c 1 -f -o /tmp/kr python profiles/run_one_step.py 1920x1080x256 cs > /tmp/kk.log 2>&1; python profiles/summarize.py kernel /tmp/kr.ncu-rep > $O/r02_ncu_k_refine_hd_cs3.txt 2>&1
head -4 $O/r02_launches_hd_cs3.txt; grep -E "registers_per|thread_inst_executed_per|issue_active" $O/r02_ncu_k_refine_hd_cs3.txt
