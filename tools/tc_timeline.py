"""Print the phase timeline of the fused tensor-core PPO step kernel (CTA 0), averaged over steps."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tianshou_b200 import _cabi
from tianshou_b200.data import VectorReplayBuffer
from tianshou_b200.synthetic import build_mujoco_ppo, fill_vector_buffer
from tianshou_b200.utils import policy_within_training_step
dev = torch.device("cuda:0")
E, T = 4096, 128
buf = VectorReplayBuffer(E * T, E, device=dev)
fill_vector_buffer(buf, np.random.default_rng(0), E, T, 17, 6)
algo, _, _ = build_mujoco_ppo(17, 6, dev, minibatch_shuffle="device")
lib = _cabi.use_diagnostics_library()
names = {0: "start", 1: "tile inputs staged", 2: "critic weights staged", 3: "critic fwd (3 MMA stages + 2 epi)", 4: "critic loss epi",
         16: "  c: dW3 MMA", 17: "  c: dz2 epi", 18: "  c: dW2/db2/dH1 MMA", 19: "  c: dz1 epi", 20: "  c: dW1/db1 MMA",
         5: "critic bwd done (REDs)", 6: "actor weights staged", 7: "actor fwd", 8: "actor loss epi", 9: "actor bwd done", 10: "tile loop end",
         11: "grid barrier 1", 23: "  next tile's inputs stored", 24: "  fold: partial rows loaded + combined", 12: "fold done (ss reduced)",
         13: "grid barrier 2", 25: "  clip coefficient known", 26: "  adam + image scatter", 14: "fence.proxy.async", 15: "grid barrier 3 (params visible)"}
with policy_within_training_step(algo.policy):
    algo.update(buffer=buf, batch_size=16384, repeat=1)
    torch.cuda.synchronize()
    CTA = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    lib.ts_tc_timeline(1 + CTA, None)
    acc = {}
    n = 0
    for it in range(3):
        algo.update(buffer=buf, batch_size=16384, repeat=1)
        torch.cuda.synchronize()
        out = (C.c_uint64 * 32)()
        lib.ts_tc_timeline(1 + CTA, out)
        t = np.array(out[:], dtype=np.int64)
        order = [0, 1, 2, 3, 4, 16, 17, 18, 19, 20, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]
        # slots 16..20 are overwritten by the actor pass; report the actor's inner split separately
        prev = t[0]
        seq = [(22, t[22])] + [(k, t[k]) for k in (1, 2, 3, 4, 5, 6, 7, 8, 16, 17, 18, 19, 20, 9, 10, 11, 23, 24, 12, 13, 25, 26, 14, 15)]
        for (k0, t0), (k1, t1) in zip(seq[:-1], seq[1:]):
            acc[(k0, k1)] = acc.get((k0, k1), 0) + (t1 - t0)
        acc["total"] = acc.get("total", 0) + (t[15] - t[22])
        n += 1
    lib.ts_tc_timeline(0, None)
for k, v in acc.items():
    if k == "total":
        print(f"TOTAL {v / n / 1e3:8.2f} us")
    else:
        print(f"{names.get(k[1], k[1]):40s} {v / n / 1e3:8.2f} us   (slot {k[0]} -> {k[1]})")
