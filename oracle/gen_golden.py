"""TEST INFRASTRUCTURE -- generate golden input/output vectors by running the UNMODIFIED reference
(thu-ml/tianshou 2.0.1, imported from /root/reference through oracle/ref_shim.py).

Run in the build container only (the reference tree does not travel to the GPU box):

    python -m oracle.gen_golden            # writes tests/golden/*.npz

The fixtures pin (a) the numpy/C oracle and (b) the CUDA path to the reference's actual outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle.ref_shim import import_reference  # noqa: E402

ts = import_reference()
from tianshou.algorithm import PPO, Algorithm  # noqa: E402
from tianshou.algorithm.modelfree.reinforce import ProbabilisticActorPolicy  # noqa: E402
from tianshou.algorithm.optim import AdamOptimizerFactory  # noqa: E402
from tianshou.data import (  # noqa: E402
    Batch,
    PrioritizedVectorReplayBuffer,
    ReplayBuffer,
    SegmentTree,
    VectorReplayBuffer,
)
from tianshou.utils.net.common import Net  # noqa: E402
from tianshou.utils.net.continuous import ContinuousActorProbabilistic, ContinuousCritic  # noqa: E402
from tianshou.utils.torch_utils import policy_within_training_step  # noqa: E402


def synth_rollout(rng, E, steps, obs_dim, act_dim, p_term, trunc_len):
    """Synthetic HalfCheetah-shaped rollout: per step arrays of shape [E, ...]."""
    t_in_ep = np.zeros(E, dtype=np.int64)
    obs = rng.standard_normal((E, obs_dim)).astype(np.float32)
    out = []
    for _ in range(steps):
        act = rng.standard_normal((E, act_dim)).astype(np.float32)
        rew = rng.standard_normal(E)
        obs_next = rng.standard_normal((E, obs_dim)).astype(np.float32)
        term = rng.random(E) < p_term
        t_in_ep += 1
        trunc = (t_in_ep >= trunc_len) & ~term
        out.append(dict(obs=obs, act=act, rew=rew, terminated=term, truncated=trunc, obs_next=obs_next))
        done = term | trunc
        t_in_ep[done] = 0
        fresh = rng.standard_normal((E, obs_dim)).astype(np.float32)
        obs = np.where(done[:, None], fresh, obs_next)
    return out


def fill(buf, steps, BatchCls=Batch):
    for s in steps:
        buf.add(BatchCls(obs=s["obs"], act=s["act"], rew=s["rew"], terminated=s["terminated"],
                         truncated=s["truncated"], obs_next=s["obs_next"]),
                buffer_ids=np.arange(len(s["rew"])))


def meta_of(buf):
    return dict(offset=np.asarray(buf._extend_offset, dtype=np.int64), done=np.asarray(buf.done, dtype=bool),
                last_index=np.asarray(buf.last_index, dtype=np.int64),
                lengths=np.asarray(buf._lengths, dtype=np.int64))


# ---------------------------------------------------------------------------------------------
def gen_returns():
    rng = np.random.default_rng(1234)
    out = {}
    case = 0
    for E, cap, steps, p_term, trunc_len in [(4, 16, 16, 0.1, 7), (8, 12, 18, 0.05, 5), (3, 40, 25, 0.2, 1000),
                                             (16, 64, 64, 0.02, 30), (1, 50, 50, 0.1, 9)]:
        buf = VectorReplayBuffer(E * cap, E)
        fill(buf, synth_rollout(rng, E, steps, 3, 2, p_term, trunc_len))
        batch, indices = buf.sample(0)
        n = len(indices)
        for vdtype, gamma, lam in [(np.float32, 0.99, 0.95), (np.float64, 0.9, 1.0), (np.float32, 1.0, 0.5)]:
            v_s = rng.standard_normal(n).astype(vdtype)
            v_s_ = rng.standard_normal(n).astype(vdtype)
            ret, adv = Algorithm.compute_episodic_return(batch, buf, indices, torch.from_numpy(v_s_.copy()),
                                                         torch.from_numpy(v_s.copy()), gamma=gamma, gae_lambda=lam)
            pre = f"gae{case}_"
            out.update({pre + k: v for k, v in meta_of(buf).items()})
            out.update({
                pre + "indices": indices, pre + "rew": batch.rew, pre + "terminated": batch.terminated,
                pre + "truncated": batch.truncated, pre + "buf_terminated": np.asarray(buf.terminated),
                pre + "unfinished": buf.unfinished_index(), pre + "v_s": v_s, pre + "v_s_": v_s_,
                pre + "gamma": gamma, pre + "lam": lam, pre + "returns": ret, pre + "adv": adv,
            })
            case += 1
        # Monte-Carlo form: v_s_ = None (gae_lambda must be 1)
        ret, adv = Algorithm.compute_episodic_return(batch, buf, indices, gamma=0.97, gae_lambda=1.0)
        pre = f"mc{E}_"
        out.update({pre + k: v for k, v in meta_of(buf).items()})
        out.update({pre + "indices": indices, pre + "rew": batch.rew, pre + "terminated": batch.terminated,
                    pre + "truncated": batch.truncated, pre + "unfinished": buf.unfinished_index(),
                    pre + "returns": ret, pre + "adv": adv})
    out["n_gae_cases"] = case

    # ---- n-step ---------------------------------------------------------------------------
    case = 0
    for E, cap, steps, p_term, trunc_len, A in [(4, 32, 40, 0.1, 9, 1), (8, 16, 16, 0.05, 6, 51), (1, 64, 80, 0.1, 11, 3)]:
        buf = VectorReplayBuffer(E * cap, E)
        fill(buf, synth_rollout(rng, E, steps, 3, 2, p_term, trunc_len))
        table = rng.standard_normal((buf.maxsize, A)).astype(np.float32)

        def tq_fn(buffer, idx, table=table):
            return torch.from_numpy(table[idx].copy())

        for n_step, gamma in [(1, 0.99), (3, 0.99), (10, 0.9)]:
            batch, indices = buf.sample(64)
            b = Algorithm.compute_nstep_return(batch, buf, indices, tq_fn, gamma=gamma, n_step=n_step)
            pre = f"nstep{case}_"
            out.update({pre + k: v for k, v in meta_of(buf).items()})
            out.update({pre + "indices": indices, pre + "rew": np.asarray(buf.rew), pre + "table": table,
                        pre + "buf_terminated": np.asarray(buf.terminated), pre + "n_step": n_step,
                        pre + "gamma": gamma, pre + "returns": b.returns.numpy()})
            case += 1
    out["n_nstep_cases"] = case
    np.savez_compressed(os.path.join(OUT, "returns_ref.npz"), **out)
    print("returns_ref.npz", len(out), "arrays")


# ---------------------------------------------------------------------------------------------
def gen_index():
    rng = np.random.default_rng(77)
    out = {}
    case = 0
    for E, cap, steps, p_term, trunc_len in [(4, 5, 3, 0.3, 4), (4, 5, 12, 0.3, 4), (7, 9, 20, 0.1, 5),
                                             (32, 16, 10, 0.05, 7), (1, 10, 12, 0.25, 100), (5, 8, 0, 0.1, 5)]:
        buf = VectorReplayBuffer(E * cap, E)
        roll = synth_rollout(rng, E, max(steps, 1), 2, 1, p_term, trunc_len)
        if steps == 0:  # only some sub-buffers written, others empty
            s = roll[0]
            ids = np.array([0, 2])
            buf.add(Batch(obs=s["obs"][ids], act=s["act"][ids], rew=s["rew"][ids], terminated=s["terminated"][ids],
                          truncated=s["truncated"][ids], obs_next=s["obs_next"][ids]), buffer_ids=ids)
        else:
            fill(buf, roll)
            # ragged: a few extra steps for a subset of envs
            s = synth_rollout(rng, E, 1, 2, 1, p_term, trunc_len)[0]
            ids = np.arange(0, E, 2)
            buf.add(Batch(obs=s["obs"][ids], act=s["act"][ids], rew=s["rew"][ids], terminated=s["terminated"][ids],
                          truncated=s["truncated"][ids], obs_next=s["obs_next"][ids]), buffer_ids=ids)
        pre = f"idx{case}_"
        out.update({pre + k: v for k, v in meta_of(buf).items()})
        q = np.concatenate([np.arange(-3, buf.maxsize + 3), rng.integers(0, buf.maxsize, 50)]).astype(np.int64)
        out[pre + "query"] = q
        out[pre + "prev"] = buf.prev(q)
        out[pre + "next"] = buf.next(q)
        out[pre + "unfinished"] = buf.unfinished_index()
        out[pre + "all"] = buf.sample_indices(0)
        out[pre + "sample37"] = buf.sample_indices(37)   # RNG stream of the reference (RandomState(42) each)
        out[pre + "sample5"] = buf.sample_indices(5)
        case += 1
    out["n_cases"] = case
    # single (non-vector) ReplayBuffer
    buf = ReplayBuffer(10)
    for i in range(12):
        buf.add(Batch(obs=0, act=0, rew=i + 1, terminated=i % 4 == 3, truncated=False))
    out.update({"single_" + k: v for k, v in meta_of(buf).items()} if hasattr(buf, "_extend_offset") else {})
    out["single_done"] = np.asarray(buf.done, dtype=bool)
    out["single_last_index"] = np.asarray(buf.last_index, dtype=np.int64)
    out["single_size"] = len(buf)
    q = np.arange(10)
    out["single_prev"], out["single_next"] = buf.prev(q), buf.next(q)
    out["single_all"], out["single_unfinished"] = buf.sample_indices(0), buf.unfinished_index()
    np.savez_compressed(os.path.join(OUT, "index_ref.npz"), **out)
    print("index_ref.npz", len(out), "arrays")


# ---------------------------------------------------------------------------------------------
def gen_segtree():
    rng = np.random.default_rng(5)
    out = {}
    case = 0
    for size in [1, 2, 6, 100, 1000, 4097]:
        tree = SegmentTree(size)
        ops = []
        for rnd in range(4):
            n = int(rng.integers(1, min(size, 64) + 1))
            idx = rng.integers(0, size, n)            # duplicates allowed: last write wins
            val = rng.random(n) * (10.0 ** rng.integers(-3, 3))
            if rnd == 3:
                val = val.astype(np.float32)
            tree[idx] = val
            ops.append((idx, val))
            pre = f"seg{case}_r{rnd}_"
            out[pre + "idx"], out[pre + "val"] = idx, val
            out[pre + "tree"] = tree._value.copy()
            total = tree.reduce()
            u = rng.random(33)
            out[pre + "u"] = u
            out[pre + "prefix_idx"] = tree.get_prefix_sum_idx(u * total) if total > 0 else np.zeros(33, np.int64)
            lo = int(rng.integers(0, size))
            hi = int(rng.integers(lo, size)) + 1
            out[pre + "range"] = np.array([lo, hi])
            out[pre + "range_sum"] = tree.reduce(lo, hi)
        out[f"seg{case}_size"] = size
        case += 1
    out["n_cases"] = case
    # prefix-sum corner cases of test/base/test_buffer.py:617-624
    tree = SegmentTree(6)
    tree[np.arange(6)] = np.array([0.0, 1.0, 0.5, 0.0, 0.0, 0.5])
    out["corner_tree"] = tree._value.copy()
    qs = np.array([0.0, 0.5, 1.0, 1.5, 1.99])
    out["corner_q"] = qs
    out["corner_idx"] = tree.get_prefix_sum_idx(qs.copy())  # the reference mutates its argument
    # PER end to end: index stream + weights of the reference
    np.random.seed(3)
    buf = PrioritizedVectorReplayBuffer(64, 4, alpha=0.6, beta=0.4)
    fill(buf, synth_rollout(np.random.default_rng(9), 4, 20, 3, 2, 0.1, 6))
    b, idx = buf.sample(16)
    out["per_idx0"], out["per_w0"] = idx, b.weight
    td = np.random.default_rng(10).standard_normal(16).astype(np.float32)
    buf.update_weight(idx, td)
    out["per_td"] = td
    b, idx = buf.sample(16)
    out["per_idx1"], out["per_w1"] = idx, b.weight
    out["per_tree"] = buf.weight._value.copy()
    out["per_minmax"] = np.array([buf._max_prio, buf._min_prio])
    np.savez_compressed(os.path.join(OUT, "segtree_ref.npz"), **out)
    print("segtree_ref.npz", len(out), "arrays")


# ---------------------------------------------------------------------------------------------
class _Box:
    def __init__(self, dim):
        self.shape = (dim,)
        self.low = -np.ones(dim, np.float32)
        self.high = np.ones(dim, np.float32)


def build_ref_ppo(obs_dim, act_dim, seed, **ppo_kw):
    from gymnasium.spaces import Box
    torch.manual_seed(seed)
    net_a = Net(state_shape=(obs_dim,), hidden_sizes=(64, 64), activation=torch.nn.Tanh)
    actor = ContinuousActorProbabilistic(preprocess_net=net_a, action_shape=(act_dim,), unbounded=True)
    net_c = Net(state_shape=(obs_dim,), hidden_sizes=(64, 64), activation=torch.nn.Tanh)
    critic = ContinuousCritic(preprocess_net=net_c)
    torch.nn.init.constant_(actor.sigma_param, -0.5)
    for m in list(actor.modules()) + list(critic.modules()):
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.orthogonal_(m.weight, gain=np.sqrt(2))
            torch.nn.init.zeros_(m.bias)
    for m in actor.mu.modules():
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.zeros_(m.bias)
            m.weight.data.copy_(0.01 * m.weight.data)

    def dist(loc_scale):
        loc, scale = loc_scale
        return torch.distributions.Independent(torch.distributions.Normal(loc, scale), 1)

    policy = ProbabilisticActorPolicy(actor=actor, dist_fn=dist, action_scaling=True, action_bound_method="clip",
                                      action_space=Box(-1.0, 1.0, (act_dim,)))
    lr = ppo_kw.pop("lr", 3e-4)
    algo = PPO(policy=policy, critic=critic, optim=AdamOptimizerFactory(lr=lr), **ppo_kw)
    return algo, actor, critic


def flat_named_params(actor, critic):
    """Parameters in the flat-buffer order of include/ts_b200.h (ts_actor_critic_desc)."""
    a1, a2 = [m for m in actor.preprocess.model.model if isinstance(m, torch.nn.Linear)]
    a3 = actor.mu.model[0]
    c1, c2 = [m for m in critic.preprocess.model.model if isinstance(m, torch.nn.Linear)]
    c3 = critic.last.model[0]
    return {
        "a_w1": a1.weight, "a_b1": a1.bias, "a_w2": a2.weight, "a_b2": a2.bias, "a_w3": a3.weight, "a_b3": a3.bias,
        "a_logstd": actor.sigma_param,
        "c_w1": c1.weight, "c_b1": c1.bias, "c_w2": c2.weight, "c_b2": c2.bias, "c_w3": c3.weight, "c_b3": c3.bias,
    }


def gen_ppo():
    import tianshou.algorithm.modelfree.ppo as ref_ppo
    variants = {
        "A": dict(E=16, cap=32, steps=32, bs=128, repeat=3, seed=0,
                  kw=dict(gamma=0.99, gae_lambda=0.95, max_grad_norm=0.5, vf_coef=0.25, ent_coef=0.0,
                          return_scaling=True, eps_clip=0.2, value_clip=True, dual_clip=None,
                          advantage_normalization=False, recompute_advantage=True)),
        "B": dict(E=16, cap=32, steps=32, bs=100, repeat=2, seed=1,
                  kw=dict(gamma=0.98, gae_lambda=0.9, max_grad_norm=None, vf_coef=0.5, ent_coef=0.01,
                          return_scaling=False, eps_clip=0.1, value_clip=False, dual_clip=2.0,
                          advantage_normalization=True, recompute_advantage=False, lr=1e-3)),
        "C": dict(E=8, cap=24, steps=36, bs=64, repeat=2, seed=2,          # wrapped sub-buffers
                  kw=dict(gamma=0.99, gae_lambda=0.95, max_grad_norm=0.5, vf_coef=0.25, ent_coef=0.0,
                          return_scaling=True, eps_clip=0.2, value_clip=True, dual_clip=None,
                          advantage_normalization=False, recompute_advantage=True)),
        "D": dict(E=6, cap=20, steps=13, bs=None, repeat=2, seed=3,         # partially filled, one minibatch
                  kw=dict(gamma=0.99, gae_lambda=0.95, max_grad_norm=0.5, vf_coef=0.25, ent_coef=0.001,
                          return_scaling=True, eps_clip=0.2, value_clip=True, dual_clip=None,
                          advantage_normalization=True, recompute_advantage=True)),
    }
    obs_dim, act_dim = 17, 6
    for name, cfg in variants.items():
        rng = np.random.default_rng(100 + cfg["seed"])
        algo, actor, critic = build_ref_ppo(obs_dim, act_dim, cfg["seed"], **dict(cfg["kw"]))
        steps = synth_rollout(rng, cfg["E"], cfg["steps"], obs_dim, act_dim, 0.03, 20)
        buf = VectorReplayBuffer(cfg["E"] * cfg["cap"], cfg["E"])
        fill(buf, steps)
        out = {"p0_" + k: v.detach().numpy().copy() for k, v in flat_named_params(actor, critic).items()}
        # second update on fresh data exercises Adam state / ret_rms carry-over
        steps2 = synth_rollout(rng, cfg["E"], cfg["steps"], obs_dim, act_dim, 0.03, 20)
        captured = {"pre": [], "seq": []}
        orig_pre = algo._preprocess_batch

        def pre_hook(batch, buffer, indices, orig_pre=orig_pre, captured=captured):
            b = orig_pre(batch, buffer, indices)
            captured["pre"].append({k: b[k].detach().numpy().copy() for k in ("v_s", "returns", "adv", "logp_old")}
                                   | {"indices": np.asarray(indices).copy()})
            return b

        algo._preprocess_batch = pre_hook
        orig_from = ref_ppo.SequenceSummaryStats.from_sequence

        def rec(seq, orig_from=orig_from, captured=captured):
            captured["seq"].append(np.asarray(seq, dtype=np.float64))
            return orig_from(seq)

        ref_ppo.SequenceSummaryStats.from_sequence = rec
        for u, st in enumerate([steps, steps2]):
            if u == 1:
                buf.reset(keep_statistics=True)
                fill(buf, st)
            np.random.seed(1000 + u)
            torch.manual_seed(2000 + u)
            N = len(buf)
            with policy_within_training_step(algo.policy):
                stats = algo.update(buffer=buf, batch_size=cfg["bs"], repeat=cfg["repeat"])
            # replay the RNG stream the update consumed: `repeat` permutations of N
            np.random.seed(1000 + u)
            perms = np.stack([np.random.permutation(N) for _ in range(cfg["repeat"])])
            pre = captured["pre"][u]
            seqs = captured["seq"][4 * u: 4 * u + 4]
            o = f"u{u}_"
            out.update({o + "perms": perms, o + "indices": pre["indices"], o + "v_s": pre["v_s"],
                        o + "returns": pre["returns"], o + "adv": pre["adv"], o + "logp_old": pre["logp_old"],
                        o + "losses": np.stack(seqs, axis=1),   # columns: loss, clip, vf, ent
                        o + "gradient_steps": stats.gradient_steps,
                        o + "rms": np.array([float(algo.ret_rms.mean), float(algo.ret_rms.var), float(algo.ret_rms.count)])})
            out.update({o + "p_" + k: v.detach().numpy().copy() for k, v in flat_named_params(actor, critic).items()})
            for key in ("obs", "act", "rew", "terminated", "truncated", "obs_next", "done"):
                out[o + "buf_" + key] = np.asarray(buf._meta[key]).copy()
            out.update({o + "meta_" + k: v for k, v in meta_of(buf).items()})
            out[o + "unfinished"] = buf.unfinished_index()
        ref_ppo.SequenceSummaryStats.from_sequence = orig_from
        out["cfg_E"], out["cfg_cap"], out["cfg_steps"] = cfg["E"], cfg["cap"], cfg["steps"]
        out["cfg_bs"] = -1 if cfg["bs"] is None else cfg["bs"]
        out["cfg_repeat"] = cfg["repeat"]
        for k, v in cfg["kw"].items():
            out["kw_" + k] = np.nan if v is None else v
        np.savez_compressed(os.path.join(OUT, f"ppo_ref_{name}.npz"), **out)
        print(f"ppo_ref_{name}.npz", len(out), "arrays; gradient_steps", int(out["u0_gradient_steps"]))


# ---------------------------------------------------------------------------------------------
# BASELINE configs[0]: the reference's own discrete PPO test network (test/discrete/test_ppo_discrete.py:90-125):
# ONE Net(obs -> 64 -> 64, ReLU) shared by DiscreteActor(softmax_output=True) and DiscreteCritic,
# orthogonal init, Categorical(probs), CartPole shapes (obs 4, 2 actions).
def build_ref_ppo_discrete(obs_dim, n_act, seed, shared=True, **ppo_kw):
    from gymnasium.spaces import Discrete
    from tianshou.algorithm.modelfree.reinforce import DiscreteActorPolicy
    from tianshou.utils.net.common import ActorCritic
    from tianshou.utils.net.discrete import DiscreteActor, DiscreteCritic
    torch.manual_seed(seed)
    net = Net(state_shape=(obs_dim,), hidden_sizes=(64, 64))
    net_c = net if shared else Net(state_shape=(obs_dim,), hidden_sizes=(64, 64))
    actor = DiscreteActor(preprocess_net=net, action_shape=(n_act,))
    critic = DiscreteCritic(preprocess_net=net_c)
    for m in ActorCritic(actor, critic).modules():
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.orthogonal_(m.weight)
            torch.nn.init.zeros_(m.bias)
    policy = DiscreteActorPolicy(actor=actor, dist_fn=torch.distributions.Categorical, action_space=Discrete(n_act),
                                 deterministic_eval=True)
    lr = ppo_kw.pop("lr", 3e-4)
    algo = PPO(policy=policy, critic=critic, optim=AdamOptimizerFactory(lr=lr), **ppo_kw)
    return algo, actor, critic


def discrete_named_params(actor, critic):
    t1, t2 = [m for m in actor.preprocess.model.model if isinstance(m, torch.nn.Linear)]
    c1, c2 = [m for m in critic.preprocess.model.model if isinstance(m, torch.nn.Linear)]
    a3, c3 = actor.last.model[0], critic.last.model[0]
    d = {"a_w1": t1.weight, "a_b1": t1.bias, "a_w2": t2.weight, "a_b2": t2.bias, "a_w3": a3.weight, "a_b3": a3.bias}
    if c1 is not t1:
        d.update({"c_w1": c1.weight, "c_b1": c1.bias, "c_w2": c2.weight, "c_b2": c2.bias})
    d.update({"c_w3": c3.weight, "c_b3": c3.bias})
    return d


def synth_rollout_discrete(rng, E, steps, obs_dim, n_act, p_term, trunc_len):
    """CartPole-shaped: obs f64 (gymnasium's CartPole returns float32; the buffer keeps what it gets), integer
    actions, reward 1.0 per step."""
    t_in_ep = np.zeros(E, dtype=np.int64)
    obs = rng.standard_normal((E, obs_dim)).astype(np.float32)
    out = []
    for _ in range(steps):
        act = rng.integers(0, n_act, E)
        rew = np.ones(E) + 0.1 * rng.standard_normal(E)
        obs_next = (obs + 0.1 * rng.standard_normal((E, obs_dim))).astype(np.float32)
        term = rng.random(E) < p_term
        t_in_ep += 1
        trunc = (t_in_ep >= trunc_len) & ~term
        out.append(dict(obs=obs, act=act, rew=rew, terminated=term, truncated=trunc, obs_next=obs_next))
        done = term | trunc
        t_in_ep[done] = 0
        fresh = rng.standard_normal((E, obs_dim)).astype(np.float32)
        obs = np.where(done[:, None], fresh, obs_next)
    return out


def gen_ppo_discrete():
    import tianshou.algorithm.modelfree.ppo as ref_ppo
    variants = {
        # the reference test's hyper-parameters (test_ppo_discrete.py:28-63) on 10 envs (BASELINE configs[0])
        "C1": dict(E=10, cap=40, steps=40, bs=64, repeat=3, seed=0, shared=True,
                   kw=dict(gamma=0.99, gae_lambda=0.95, max_grad_norm=0.5, vf_coef=0.5, ent_coef=0.0,
                           return_scaling=False, eps_clip=0.2, value_clip=False, dual_clip=None,
                           advantage_normalization=False, recompute_advantage=False)),
        # every optional term on (entropy bonus, advantage normalisation, value clip, dual clip, return scaling)
        "C1b": dict(E=8, cap=24, steps=30, bs=50, repeat=2, seed=1, shared=True,
                    kw=dict(gamma=0.98, gae_lambda=0.9, max_grad_norm=0.5, vf_coef=0.25, ent_coef=0.01,
                            return_scaling=True, eps_clip=0.2, value_clip=True, dual_clip=3.0,
                            advantage_normalization=True, recompute_advantage=True, lr=1e-3)),
        # separate trunks
        "C1c": dict(E=6, cap=20, steps=20, bs=None, repeat=2, seed=2, shared=False,
                    kw=dict(gamma=0.99, gae_lambda=0.95, max_grad_norm=None, vf_coef=0.5, ent_coef=0.005,
                            return_scaling=False, eps_clip=0.2, value_clip=False, dual_clip=None,
                            advantage_normalization=False, recompute_advantage=False)),
    }
    obs_dim, n_act = 4, 2
    for name, cfg in variants.items():
        rng = np.random.default_rng(300 + cfg["seed"])
        algo, actor, critic = build_ref_ppo_discrete(obs_dim, n_act, cfg["seed"], shared=cfg["shared"], **dict(cfg["kw"]))
        steps = synth_rollout_discrete(rng, cfg["E"], cfg["steps"], obs_dim, n_act, 0.04, 25)
        buf = VectorReplayBuffer(cfg["E"] * cfg["cap"], cfg["E"])
        fill(buf, steps)
        out = {"p0_" + k: v.detach().numpy().copy() for k, v in discrete_named_params(actor, critic).items()}
        steps2 = synth_rollout_discrete(rng, cfg["E"], cfg["steps"], obs_dim, n_act, 0.04, 25)
        captured = {"pre": [], "seq": []}
        orig_pre = algo._preprocess_batch

        def pre_hook(batch, buffer, indices, orig_pre=orig_pre, captured=captured):
            b = orig_pre(batch, buffer, indices)
            captured["pre"].append({k: b[k].detach().numpy().copy() for k in ("v_s", "returns", "adv", "logp_old")}
                                   | {"indices": np.asarray(indices).copy()})
            return b

        algo._preprocess_batch = pre_hook
        orig_from = ref_ppo.SequenceSummaryStats.from_sequence

        def rec(seq, orig_from=orig_from, captured=captured):
            captured["seq"].append(np.asarray(seq, dtype=np.float64))
            return orig_from(seq)

        ref_ppo.SequenceSummaryStats.from_sequence = rec
        for u, st in enumerate([steps, steps2]):
            if u == 1:
                buf.reset(keep_statistics=True)
                fill(buf, st)
            np.random.seed(1000 + u)
            torch.manual_seed(2000 + u)
            N = len(buf)
            with policy_within_training_step(algo.policy):
                stats = algo.update(buffer=buf, batch_size=cfg["bs"], repeat=cfg["repeat"])
            np.random.seed(1000 + u)
            perms = np.stack([np.random.permutation(N) for _ in range(cfg["repeat"])])
            pre = captured["pre"][u]
            seqs = captured["seq"][4 * u: 4 * u + 4]
            o = f"u{u}_"
            out.update({o + "perms": perms, o + "indices": pre["indices"], o + "v_s": pre["v_s"],
                        o + "returns": pre["returns"], o + "adv": pre["adv"], o + "logp_old": pre["logp_old"],
                        o + "losses": np.stack(seqs, axis=1), o + "gradient_steps": stats.gradient_steps,
                        o + "rms": np.array([float(algo.ret_rms.mean), float(algo.ret_rms.var), float(algo.ret_rms.count)])})
            out.update({o + "p_" + k: v.detach().numpy().copy() for k, v in discrete_named_params(actor, critic).items()})
            for key in ("obs", "act", "rew", "terminated", "truncated", "obs_next", "done"):
                out[o + "buf_" + key] = np.asarray(buf._meta[key]).copy()
            out.update({o + "meta_" + k: v for k, v in meta_of(buf).items()})
            out[o + "unfinished"] = buf.unfinished_index()
        ref_ppo.SequenceSummaryStats.from_sequence = orig_from
        out["cfg_E"], out["cfg_cap"], out["cfg_steps"] = cfg["E"], cfg["cap"], cfg["steps"]
        out["cfg_bs"] = -1 if cfg["bs"] is None else cfg["bs"]
        out["cfg_repeat"], out["cfg_shared"] = cfg["repeat"], int(cfg["shared"])
        for k, v in cfg["kw"].items():
            out["kw_" + k] = np.nan if v is None else v
        np.savez_compressed(os.path.join(OUT, f"ppo_ref_{name}.npz"), **out)
        print(f"ppo_ref_{name}.npz", len(out), "arrays; gradient_steps", int(out["u0_gradient_steps"]))


def gen_a2c():
    """Reference A2C (a2c.py:156-299) on the MuJoCo-shaped nets: same capture as gen_ppo (no logp_old)."""
    import tianshou.algorithm.modelfree.a2c as ref_a2c
    from tianshou.algorithm import A2C
    obs_dim, act_dim = 17, 6
    cfg = dict(E=16, cap=32, steps=32, bs=128, repeat=2, seed=5,
               kw=dict(gamma=0.99, gae_lambda=0.95, max_grad_norm=0.5, vf_coef=0.5, ent_coef=0.01, return_scaling=True))
    rng = np.random.default_rng(500)
    ppo_algo, actor, critic = build_ref_ppo(obs_dim, act_dim, cfg["seed"])       # nets + policy, PPO wrapper discarded
    algo = A2C(policy=ppo_algo.policy, critic=critic, optim=AdamOptimizerFactory(lr=3e-4), **cfg["kw"])
    steps = synth_rollout(rng, cfg["E"], cfg["steps"], obs_dim, act_dim, 0.03, 20)
    steps2 = synth_rollout(rng, cfg["E"], cfg["steps"], obs_dim, act_dim, 0.03, 20)
    buf = VectorReplayBuffer(cfg["E"] * cfg["cap"], cfg["E"])
    fill(buf, steps)
    out = {"p0_" + k: v.detach().numpy().copy() for k, v in flat_named_params(actor, critic).items()}
    captured = {"pre": [], "seq": []}
    orig_pre = algo._preprocess_batch

    def pre_hook(batch, buffer, indices):
        b = orig_pre(batch, buffer, indices)
        captured["pre"].append({k: b[k].detach().numpy().copy() for k in ("v_s", "returns", "adv")}
                               | {"indices": np.asarray(indices).copy()})
        return b

    algo._preprocess_batch = pre_hook
    orig_from = ref_a2c.SequenceSummaryStats.from_sequence

    def rec(seq):
        captured["seq"].append(np.asarray(seq, dtype=np.float64))
        return orig_from(seq)

    ref_a2c.SequenceSummaryStats.from_sequence = rec
    for u, st in enumerate([steps, steps2]):
        if u == 1:
            buf.reset(keep_statistics=True)
            fill(buf, st)
        np.random.seed(1000 + u)
        torch.manual_seed(2000 + u)
        N = len(buf)
        with policy_within_training_step(algo.policy):
            stats = algo.update(buffer=buf, batch_size=cfg["bs"], repeat=cfg["repeat"])
        np.random.seed(1000 + u)
        perms = np.stack([np.random.permutation(N) for _ in range(cfg["repeat"])])
        pre = captured["pre"][u]
        seqs = captured["seq"][4 * u: 4 * u + 4]
        o = f"u{u}_"
        out.update({o + "perms": perms, o + "indices": pre["indices"], o + "v_s": pre["v_s"], o + "returns": pre["returns"],
                    o + "adv": pre["adv"], o + "losses": np.stack(seqs, axis=1), o + "gradient_steps": stats.gradient_steps,
                    o + "rms": np.array([float(algo.ret_rms.mean), float(algo.ret_rms.var), float(algo.ret_rms.count)])})
        out.update({o + "p_" + k: v.detach().numpy().copy() for k, v in flat_named_params(actor, critic).items()})
        for key in ("obs", "act", "rew", "terminated", "truncated", "obs_next", "done"):
            out[o + "buf_" + key] = np.asarray(buf._meta[key]).copy()
        out.update({o + "meta_" + k: v for k, v in meta_of(buf).items()})
        out[o + "unfinished"] = buf.unfinished_index()
    ref_a2c.SequenceSummaryStats.from_sequence = orig_from
    out["cfg_E"], out["cfg_cap"], out["cfg_steps"], out["cfg_bs"], out["cfg_repeat"] = cfg["E"], cfg["cap"], cfg["steps"], cfg["bs"], cfg["repeat"]
    for k, v in cfg["kw"].items():
        out["kw_" + k] = np.nan if v is None else v
    np.savez_compressed(os.path.join(OUT, "a2c_ref.npz"), **out)
    print("a2c_ref.npz", len(out), "arrays; gradient_steps", int(out["u0_gradient_steps"]))


def gen_fromdata():
    """sample_indices(0) / prev / next / unfinished_index on buffers whose insertion index was NOT produced by add():
    ``ReplayBuffer.from_data`` (buffer_base.py:382-418: size = N, _insertion_idx stays 0, last_index stays 0), then a few
    ``add`` calls on top (wrap-around from slot 0)."""
    from tianshou.data import ReplayBuffer as RefRB
    rng = np.random.default_rng(99)
    out = {}
    for case, (n, extra) in enumerate([(7, 0), (7, 3), (12, 12), (5, 7)]):
        obs = rng.standard_normal((n, 2)).astype(np.float32)
        act = rng.standard_normal((n, 1)).astype(np.float32)
        rew = rng.standard_normal(n)
        term = rng.random(n) < 0.2
        trunc = (rng.random(n) < 0.1) & ~term
        obs_next = rng.standard_normal((n, 2)).astype(np.float32)
        buf = RefRB.from_data(obs, act, rew, term, trunc, term | trunc, obs_next)
        p = f"fd{case}_"
        out.update({p + "obs": obs, p + "act": act, p + "rew": rew, p + "terminated": term, p + "truncated": trunc,
                    p + "obs_next": obs_next, p + "extra": extra})
        for j in range(extra):
            b = Batch(obs=rng.standard_normal(2).astype(np.float32), act=rng.standard_normal(1).astype(np.float32),
                      rew=float(rng.standard_normal()), terminated=bool(rng.random() < 0.2), truncated=False,
                      obs_next=rng.standard_normal(2).astype(np.float32))
            for k in ("obs", "act", "rew", "terminated", "truncated", "obs_next"):
                out[p + f"add{j}_" + k] = np.asarray(b[k])
            buf.add(b)
        q = np.arange(-2, n + 2)
        out.update({p + "all": buf.sample_indices(0), p + "query": q, p + "prev": buf.prev(q % n), p + "next": buf.next(q % n),
                    p + "unfinished": buf.unfinished_index(), p + "len": len(buf), p + "last_index": np.asarray(buf.last_index),
                    p + "done": np.asarray(buf.done, dtype=bool)})
    out["n_cases"] = 4
    np.savez_compressed(os.path.join(OUT, "fromdata_ref.npz"), **out)
    print("fromdata_ref.npz", len(out), "arrays")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["returns", "index", "segtree", "ppo", "ppo_discrete", "a2c", "fromdata"]
    for w in which:
        {"returns": gen_returns, "index": gen_index, "segtree": gen_segtree, "ppo": gen_ppo,
         "ppo_discrete": gen_ppo_discrete, "a2c": gen_a2c, "fromdata": gen_fromdata}[w]()
