#!/bin/bash
# round-2 measurement lines at the final kernels
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/m; O=gpurun_out/m
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $O/gpu.txt; nproc >> $O/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu.txt
timeout 900 python bench.py > $O/r02_bench_n1_hd_ss.json 2> $O/bench_hd.err
for cfg in C1 C2 C3; do timeout 1200 python bench.py --config $cfg --cpu-full > $O/cfg_${cfg}_gpu.json 2> $O/cfg_$cfg.err; done
timeout 900 python bench.py --config C4cs > $O/cfg_C4cs_gpu.json 2> $O/cfg_C4cs.err
for sch in 1 2; do timeout 600 python bench.py --schedule $sch --no-cpu-baseline --no-worst-case > $O/r02_bench_n1_hd_ss_redblack$sch.json 2>> $O/bench_hd.err; done
for cfg in C1 C2 C3; do timeout 600 python bench.py --config $cfg --pairs-in-flight 4 --no-cpu-baseline --no-tap-count > $O/cfg_${cfg}_gpu_batch4.json 2>> $O/bench_hd.err; done
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/r02_bench_reference_arm.json 2>> $O/bench_hd.err
N="ncu --clock-control none"
timeout 900 $N --metrics gpu__time_duration.sum --csv --log-file /tmp/ll.csv python profiles/run_one_step.py 1920x1080x256 cs > /tmp/ll.log 2>&1; python profiles/summarize.py launches /tmp/ll.csv > $O/r02_launches_hd_cs3.txt
timeout 900 $N --metrics gpu__time_duration.sum --csv --log-file /tmp/ll.csv python profiles/run_one_step.py 1242x375x128 cs > /tmp/ll.log 2>&1; python profiles/summarize.py launches /tmp/ll.csv > $O/r02_launches_c3.txt
timeout 900 $N --set full -k regex:k_refine -s 1 -c 1 -f -o /tmp/kr python profiles/run_one_step.py 1920x1080x256 cs > /tmp/kk.log 2>&1; python profiles/summarize.py kernel /tmp/kr.ncu-rep > $O/r02_ncu_k_refine_hd_cs3.txt 2>&1
for f in $O/*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').readline()); print(round(d['value'],1), d['unit'], round(d.get('ms_per_step',0),1),'ms', 'parity', d.get('parity_checked'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))")"; done
