#!/usr/bin/env bash
# One GPU-box pass that produces the evidence files of a build: GPU tests, bench lines (both arms),
# ncu launch list of the bench command, ncu --set full of the lane kernels.
#   gpurun --timeout 1500 -- 'bash tools/evidence.sh r02f'
# Outputs under gpurun_out/<tag>_*; summarise with tools/ncu_summary.py / tools/update_traffic.py.
set -u
tag=${1:-rXX}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
o=gpurun_out/${tag}
timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 > ${o}_pytest.log 2>&1; echo "pytest rc=$?" >> ${o}_pytest.log
tail -4 ${o}_pytest.log
timeout 600 python bench.py > ${o}_bench.json 2> ${o}_bench.err; tail -c 400 ${o}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${o}_bench_ref.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${o}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${o}_b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 2 -c 2 -o ${o}_lane \
    python tools/profile_reduce.py 10000 32 8 2 > ${o}_ncu.log 2>&1; tail -2 ${o}_ncu.log
python - <<PY
import json
d = json.load(open("${o}_bench.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["stage_ms"], d["config"].get("strong"))
PY
