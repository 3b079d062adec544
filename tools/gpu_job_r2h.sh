#!/bin/bash
# GPU-box visit r2h: parity suite + bench + host timeline with the parked thread crew.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2h_tests.txt
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
timeout 120 python tools/default_order_hosttrace.py > gpurun_out/r2h_default_order_hosttrace.txt 2>&1
timeout 200 python tools/default_order_probe.py > gpurun_out/r2h_default_order_probe.txt 2>&1
tail -6 gpurun_out/r2h_tests.txt
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2h_bench.json") if l.startswith("{")][-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "dev-order", d["value_device_order"]["ms_per_step"], d["e2e_device_order"]["ms_per_step"], d["default_over_device_order"], "us/step", d["roofline"]["us_per_optimiser_step"])
PY
tail -3 gpurun_out/r2h_bench.err
tail -16 gpurun_out/r2h_default_order_hosttrace.txt
cat gpurun_out/r2h_default_order_probe.txt
