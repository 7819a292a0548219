#!/bin/bash
# One GPU-box pass: tests, smoke, headline bench, config benches, launch list, ncu captures of the two dominant kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
for c in C1 C2 C3 C4cs; do python bench.py --config $c --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/cfg_${c}_gpu.json 2>> gpurun_out/bench_final.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python profiles/run_one_step.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_cs3.csv python profiles/run_one_step.py 1920x1080x256 cs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_refine -s 1 -c 1 -o gpurun_out/final_refine -f python profiles/run_one_step.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spatial_persist -s 1 -c 1 -o gpurun_out/final_spatial -f python profiles/run_one_step.py > /dev/null 2>&1
head -c 400 gpurun_out/bench_final.json; echo; for c in C1 C2 C3 C4cs; do python -c "import json;d=json.load(open('gpurun_out/cfg_${c}_gpu.json'));print('$c',round(d['value'],1),round(d['ms_per_step'],1),d['roofline']['frac'])"; done
