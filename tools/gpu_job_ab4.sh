#!/bin/bash
# A/B of the host permutation crew under several ranks per box (weak scaling, default order), N GPUs.
cd "$(dirname "$0")/.."
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; shift; env "$@" timeout 300 $TR bench.py --gpus $N --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ab_${tag}_n${N}.json 2> gpurun_out/ab_${tag}_n${N}.err; }
run base TS_B200_PERM_NICE=0
run nice TS_B200_PERM_NICE=10
run w2 TS_B200_PERM_WORKERS=2
run w2nice TS_B200_PERM_WORKERS=2 TS_B200_PERM_NICE=10
run nopin TS_B200_PERM_PIN=0
python - <<PY
import json
for t in ("base","nice","w2","w2nice","nopin"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/ab_{t}_n$N.json") if l.startswith("{")][-1])
        print(t, "value", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "dev", round(d["value_device_order"]["ms_per_step"],2), round(d["e2e_device_order"]["ms_per_step"],2))
    except Exception as e:
        print(t, "failed", e)
PY
