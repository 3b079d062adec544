#!/bin/bash
# What the driver does at round end, on one GPU: build check, smoke, parity suite, bench (both arms).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1; tail -2 gpurun_out/final_smoke.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/final_tests.txt; cat gpurun_out/final_tests.txt
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -2 gpurun_out/final_bench_reference.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/final_bench.json") if l.startswith("{")][-1])
print("value", d["value"], d["ms_per_step"], "steps", d["steps"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "dev-order", d["value_device_order"]["ms_per_step"], d["e2e_device_order"]["ms_per_step"], d["default_over_device_order"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], d["clocks"])
r=json.loads([l for l in open("gpurun_out/final_bench_reference.json") if l.startswith("{")][-1])
print("reference arm", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["sample"][:120])
PY
