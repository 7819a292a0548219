#!/bin/bash
# Light GPU-box pass: tests, headline bench, two config lines, launch list (no ncu --set full captures).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --config C4cs --no-cpu-baseline > gpurun_out/cfg_C4cs_gpu.json 2>> gpurun_out/bench_final.err
python bench.py --config C1 --no-cpu-baseline > gpurun_out/cfg_C1_gpu.json 2>> gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python profiles/run_one_step.py > /dev/null 2>&1
head -c 300 gpurun_out/bench_final.json; echo
for c in C1 C4cs; do python -c "import json;d=json.load(open('gpurun_out/cfg_${c}_gpu.json'));print('$c',round(d['value'],1),round(d['ms_per_step'],1),d['roofline']['frac'])"; done
ncu --set full --clock-control none --import-source on -k regex:k_refine -s 1 -c 1 -o gpurun_out/final_refine -f python profiles/run_one_step.py > /dev/null 2>&1
ls -la gpurun_out/final_refine.ncu-rep
