#!/bin/bash
# GPU-box visit r2i: next-tile input staging moved under barriers 2 / 3 -- parity suite, bench, phase timeline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2i_tests.txt; cat gpurun_out/r2i_tests.txt
timeout 500 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -2 gpurun_out/r2i_bench.err
timeout 200 python tools/tc_timeline.py > gpurun_out/r2i_timeline.txt 2>&1; cat gpurun_out/r2i_timeline.txt | tail -28
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2i_bench.json") if l.startswith("{")][-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "dev-order", d["value_device_order"]["ms_per_step"], d["e2e_device_order"]["ms_per_step"], "us/step", d["roofline"]["us_per_optimiser_step"], "frac", d["roofline"]["frac"])
PY
