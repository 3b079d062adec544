"""One warm-up PPO update, then ONE update inside a cudaProfilerStart/Stop range (for ncu
--profile-from-start off).  Workload = bench.py's (4096 envs x 128 steps, bs 16384, repeat 10)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tianshou_b200.data import VectorReplayBuffer
from tianshou_b200.synthetic import build_mujoco_ppo, fill_vector_buffer
from tianshou_b200.utils import policy_within_training_step

E = int(os.environ.get("TS_PROF_ENVS", "4096")); T = 128
repeat = int(os.environ.get("TS_PROF_REPEAT", "10"))
dev = torch.device("cuda:0")
buf = VectorReplayBuffer(E * T, E, device=dev)
fill_vector_buffer(buf, np.random.default_rng(0), E, T, 17, 6)
algo, _, _ = build_mujoco_ppo(17, 6, dev, minibatch_shuffle="device")
with policy_within_training_step(algo.policy):
    algo.update(buffer=buf, batch_size=16384, repeat=2)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    algo.update(buffer=buf, batch_size=16384, repeat=repeat)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
