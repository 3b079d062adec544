"""TEST INFRASTRUCTURE (build box only: needs /root/reference) -- calibrate bench.py's CPU arm.

bench.py's ``--impl reference`` / ``cpu_baseline`` time the numpy PORT of the reference's PPO update
(oracle/oracle_np.py), because the reference tree does not travel to the GPU box.  This script times the
UNMODIFIED imported reference (``tianshou.algorithm.PPO.update`` through oracle/ref_shim.py) and the port on
the SAME sample, same threads, same minibatch size / repeat / hyper-parameters, and writes

    profiles/cpu_port_calibration.json   {"reference_tps", "port_tps", "port_over_reference", ...}

so that a reader can convert the bench line's ``vs_reference`` (measured against the port) into a ratio
against the real ``tianshou/algorithm/modelfree/ppo.py``.

    python tools/cpu_port_calibration.py [--envs 256] [--steps 3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OBS, ACT, T = 17, 6, 128
BATCH_SIZE, REPEAT = 16384, 10


def time_reference(E: int, steps: int, warmup: int) -> dict:
    from oracle import gen_golden as gg                     # imports the reference through the stub shim
    from tianshou.data import Batch, VectorReplayBuffer
    from tianshou.utils.torch_utils import policy_within_training_step

    from tianshou_b200.synthetic import MUJOCO_PPO_KWARGS, synth_rollout
    algo, actor, critic = gg.build_ref_ppo(OBS, ACT, 0, **dict(MUJOCO_PPO_KWARGS))
    buf = VectorReplayBuffer(E * T, E)
    for s in synth_rollout(np.random.default_rng(0), E, T, OBS, ACT):
        buf.add(Batch(**s), buffer_ids=np.arange(E))
    times = []
    np.random.seed(0)
    torch.manual_seed(0)
    for it in range(warmup + steps):
        with policy_within_training_step(algo.policy):
            t0 = time.perf_counter()
            algo.update(buffer=buf, batch_size=min(BATCH_SIZE, E * T), repeat=REPEAT)
            dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return {"tps": E * T / (sum(times) / len(times)), "ms_per_update": 1e3 * sum(times) / len(times), "calls": len(times)}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    import bench
    threads = torch.get_num_threads()
    ref = time_reference(args.envs, args.steps, args.warmup)
    port = bench.cpu_reference_run(args.envs, T, args.steps, args.warmup)
    out = {
        "what": "imported reference PPO.update (tianshou 2.0.1, /root/reference via oracle/ref_shim.py) vs the numpy port "
                "(oracle/oracle_np.py) that bench.py times on the GPU box; same sample, same host",
        "sample": f"{args.envs} envs x {T} steps = {args.envs * T} transitions, minibatch {min(BATCH_SIZE, args.envs * T)}, "
                  f"repeat {REPEAT}, {args.steps} timed update() calls each after {args.warmup} warm-up",
        "host_cores": os.cpu_count(), "torch_threads": threads, "port_blas_threads": port["cores"],
        "reference_tps": ref["tps"], "reference_ms_per_update": ref["ms_per_update"],
        "port_tps": port["value"], "port_ms_per_update": port["ms_per_step"],
        "port_over_reference": port["value"] / ref["tps"],
        "note": "bench.py reports vs the port; multiply a GPU/port ratio by port_over_reference to read it against the "
                "real reference on this host class",
    }
    path = os.path.join(ROOT, "profiles", "cpu_port_calibration.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
