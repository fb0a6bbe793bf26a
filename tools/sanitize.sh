#!/usr/bin/env bash
# compute-sanitizer memcheck + racecheck over small batches of every kernel family
# (SURVEY.md section 5).  Run on the GPU box:  gpurun -- 'bash tools/sanitize.sh'
# Logs go to gpurun_out/sanitize_<tool>_<family>.log; copy the summaries to profiles/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
    for fam in lp_small lp_generic lp_lane lp_cta diff hull sets; do
        log=gpurun_out/sanitize_${tool}_${fam}.log
        timeout 900 compute-sanitizer --tool $tool --print-limit 20 --log-file $log \
            python tools/sanitize_driver.py $fam > gpurun_out/sanitize_${tool}_${fam}.out 2>&1
        echo "$tool $fam rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
    done
done
