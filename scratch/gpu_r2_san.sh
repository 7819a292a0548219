#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/san; O=gpurun_out/san
CS=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; shift; env "$@" timeout 900 $CS --tool $TOOL --print-limit 5 --error-exitcode 7 python profiles/sanitize_small.py > $O/${TOOL}_$name.log 2>&1; echo "$TOOL $name rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small OK' $O/${TOOL}_$name.log | tr '\n' ' ')"; }
for TOOL in memcheck racecheck; do
  run wave CSPM_SPATIAL=wave CSPM_SP_CTAS=2
  run wave_exact CSPM_SPATIAL=wave CSPM_SP_CTAS=3 CSPM_WAVE_EXACT=1
  run ring CSPM_SPATIAL=ring CSPM_SP_CTAS=2
done
