#!/bin/bash
# One GPU-box visit: parity suite, bench, phase timeline, ncu launch list of bench.py's own timed step, ncu --set full of the epoch kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2d_tests.txt
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
TS_B200_PERM_TRACE=1 timeout 100 python tools/perm_timing.py > gpurun_out/r2d_perm_timing.txt 2>&1
timeout 200 python tools/tc_timeline.py > gpurun_out/r2d_timeline.txt 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_bench_step.csv python bench.py --profile-one-step --no-extras > gpurun_out/r2d_ncu_launches.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ppo_tc_kernel -c 1 -o gpurun_out/r2d_epoch python bench.py --profile-one-step --no-extras > gpurun_out/r2d_ncu_full.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:net_gemm_kernel -c 3 -o gpurun_out/r2d_gemm python -c "
import sys; sys.path.insert(0,'.')
import torch, bench
torch.cuda.profiler.start()
print(bench.offpolicy_extras(torch.device('cuda:0')).keys())
torch.cuda.profiler.stop()" > gpurun_out/r2d_ncu_gemm.log 2>&1
tail -40 gpurun_out/r2d_tests.txt
cat gpurun_out/r2d_timeline.txt
grep -v "pass" gpurun_out/r2d_perm_timing.txt | tail -8
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2d_bench.json"))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "dev-order", d["value_device_order"], d["default_over_device_order"], "us/step", d["roofline"]["us_per_optimiser_step"])
print("offpolicy", json.dumps(d.get("offpolicy"))[:600])
PY
tail -3 gpurun_out/r2d_bench.err
