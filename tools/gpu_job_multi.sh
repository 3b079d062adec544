#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): 2-rank parity tests (N >= 2), weak-scaling bench (default), strong-scaling configs[4].
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_dist_tests_n${N}.txt; else echo "dist tests run in the 2-GPU visit" > gpurun_out/r2_dist_tests_n${N}.txt; fi
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_weak_n${N}.json 2> gpurun_out/r2_bench_weak_n${N}.err
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --scaling strong --config c5 --no-extras > gpurun_out/r2_bench_strong_c5_n${N}.json 2> gpurun_out/r2_bench_strong_c5_n${N}.err
[ "$N" = "2" ] && timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-extras > gpurun_out/r2_bench_strong_c2_n${N}.json 2> gpurun_out/r2_bench_strong_c2_n${N}.err
TS_B200_PERM_TRACE=1 timeout 300 $TR bench.py --gpus $N --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_weak_trace_n${N}.json 2> gpurun_out/r2_bench_weak_trace_n${N}.err
grep "pass  [0-3]" gpurun_out/r2_bench_weak_trace_n${N}.err | tail -32 | sort | uniq -c | tail -12
cat gpurun_out/r2_dist_tests_n${N}.txt
python - <<PY
import json
for f in ("weak", "strong_c5", "strong_c2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2_bench_{f}_n$N.json") if l.startswith("{")][-1])
        print(f, d["n_gpus"], "value", round(d["value"] / 1e6, 2), "M/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"] / 1e6, 2), "dev-order", round(d["value_device_order"]["value"] / 1e6, 2), d.get("multi_gpu_check"), (d.get("config4") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2_bench_weak_n${N}.err

