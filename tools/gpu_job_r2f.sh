#!/bin/bash
# GPU-box visit r2f: parity suite + bench with the split-K net_gemm + probes of the default-order path.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2f_tests.txt
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 200 python tools/default_order_probe.py > gpurun_out/r2f_default_order_probe.txt 2>&1
timeout 120 python tools/host_pass_trace.py > gpurun_out/r2f_host_pass_trace.txt 2>&1
tail -8 gpurun_out/r2f_tests.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench.json"))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "dev-order", d["value_device_order"]["ms_per_step"], d["e2e_device_order"]["ms_per_step"], d["default_over_device_order"], "us/step", d["roofline"]["us_per_optimiser_step"])
print("offpolicy", json.dumps(d.get("offpolicy"))[:1200])
PY
tail -3 gpurun_out/r2f_bench.err
cat gpurun_out/r2f_default_order_probe.txt
tail -14 gpurun_out/r2f_host_pass_trace.txt
