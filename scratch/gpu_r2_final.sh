#!/bin/bash
# refresh of the small-image lines and wavefront summaries at the final kernels
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/f; O=gpurun_out/f
for cfg in C1 C2 C3; do timeout 1200 python bench.py --config $cfg --cpu-full > $O/cfg_${cfg}_gpu.json 2> $O/cfg_$cfg.err; done
for cfg in C1 C2 C3; do timeout 600 python bench.py --config $cfg --pairs-in-flight 4 --no-cpu-baseline --no-tap-count > $O/cfg_${cfg}_gpu_batch4.json 2>> $O/err.log; done
N="ncu --clock-control none"
ll() { timeout 900 $N --metrics gpu__time_duration.sum --csv --log-file /tmp/ll.csv python profiles/run_one_step.py $2 $3 $4 > /tmp/ll.log 2>&1; python profiles/summarize.py launches /tmp/ll.csv > $O/r02_launches_$1.txt; tail -1 /tmp/ll.log >> $O/r02_launches_$1.txt; }
ll c1 450x375x60 ss
ll c3 1242x375x128 cs
kk() { timeout 900 $N --set full -k regex:$2 -s $3 -c 1 -f -o /tmp/$1 python profiles/run_one_step.py $4 $5 $6 > /tmp/kk.log 2>&1; python profiles/summarize.py kernel /tmp/$1.ncu-rep > $O/r02_ncu_$1.txt 2>&1; }
kk k_spatial_wave_c1 k_spatial_wave 1 450x375x60 ss
kk k_spatial_wave_c3 k_spatial_wave 1 1242x375x128 cs
for f in $O/*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').readline()); print(round(d['value'],1), round(d.get('ms_per_step',0),1),'ms', 'parity', d.get('parity_checked'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))")"; done
