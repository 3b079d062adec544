#!/bin/bash
# GPU-box visit r2e: parity suite + bench (default-order path fed by the copy stream) + host-side trace of one update.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2e_tests.txt
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
TS_B200_PERM_TRACE=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_bench_trace.json 2> gpurun_out/r2e_bench_trace.err
tail -25 gpurun_out/r2e_tests.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2e_bench.json"))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "dev-order", d["value_device_order"], d["e2e_device_order"]["ms_per_step"], d["default_over_device_order"], "us/step", d["roofline"]["us_per_optimiser_step"])
print("offpolicy", json.dumps(d.get("offpolicy"))[:400])
PY
tail -3 gpurun_out/r2e_bench.err
grep "pass  [09]" gpurun_out/r2e_bench_trace.err | tail -8
timeout 120 python tools/host_pass_trace.py > gpurun_out/r2e_host_pass_trace.txt 2>&1; tail -14 gpurun_out/r2e_host_pass_trace.txt
