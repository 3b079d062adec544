"""BASELINE configs[0] on the GPU: the reference's own discrete PPO test network (ONE ReLU Net shared by
DiscreteActor(softmax_output=True) and DiscreteCritic, Categorical(probs); test/discrete/test_ppo_discrete.py:90-125)
through the public ``PPO.update`` and through the C ABI, against outputs of the imported reference
(tests/golden/ppo_ref_C1*.npz) and against the numpy oracle (oracle/oracle_discrete.py)."""
import ctypes as C

import numpy as np
import pytest
import torch

from test_ppo_gpu import ppo_kwargs
from ts_testutil import load_golden, restore_vector_buffer

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = ["a_w1", "a_b1", "a_w2", "a_b2", "a_w3", "a_b3", "c_w1", "c_b1", "c_w2", "c_b2", "c_w3", "c_b3"]


class Discrete:
    def __init__(self, n):
        self.n = n


def build_discrete(g, device):
    from tianshou_b200.algorithm import PPO, AdamOptimizerFactory, DiscreteActorPolicy
    from tianshou_b200.utils.net.common import Net
    from tianshou_b200.utils.net.discrete import DiscreteActor, DiscreteCritic
    shared = bool(g["cfg_shared"])
    net = Net(state_shape=(4,), hidden_sizes=(64, 64))                 # default activation: ReLU, like the reference
    net_c = net if shared else Net(state_shape=(4,), hidden_sizes=(64, 64))
    actor = DiscreteActor(preprocess_net=net, action_shape=(2,)).to(device)
    critic = DiscreteCritic(preprocess_net=net_c).to(device)
    named = named_params(actor, critic)
    with torch.no_grad():
        for k, p in named.items():
            p.copy_(torch.as_tensor(g["p0_" + k]).reshape(p.shape))
    policy = DiscreteActorPolicy(actor=actor, dist_fn=torch.distributions.Categorical, action_space=Discrete(2),
                                 deterministic_eval=True)
    lr = float(g["kw_lr"]) if "kw_lr" in g.files else 3e-4
    algo = PPO(policy=policy, critic=critic, optim=AdamOptimizerFactory(lr=lr), **ppo_kwargs(g))
    return algo, actor, critic


def named_params(actor, critic):
    t1, t2 = [m for m in actor.preprocess.model.model if isinstance(m, torch.nn.Linear)]
    c1, c2 = [m for m in critic.preprocess.model.model if isinstance(m, torch.nn.Linear)]
    d = {"a_w1": t1.weight, "a_b1": t1.bias, "a_w2": t2.weight, "a_b2": t2.bias,
         "a_w3": actor.last.model[0].weight, "a_b3": actor.last.model[0].bias}
    if c1 is not t1:
        d.update({"c_w1": c1.weight, "c_b1": c1.bias, "c_w2": c2.weight, "c_b2": c2.bias})
    d.update({"c_w3": critic.last.model[0].weight, "c_b3": critic.last.model[0].bias})
    return d


@pytest.mark.parametrize("variant", ["C1", "C1b", "C1c"])
def test_discrete_ppo_update_matches_reference(variant):
    from tianshou_b200.utils import policy_within_training_step
    g = load_golden(f"ppo_ref_{variant}.npz")
    E, cap = int(g["cfg_E"]), int(g["cfg_cap"])
    bs = int(g["cfg_bs"])
    bs = None if bs < 0 else bs
    algo, actor, critic = build_discrete(g, DEV)
    assert algo._desc.flags == 3 and (algo._desc.c_w1 == algo._desc.a_w1) == bool(g["cfg_shared"])
    captured = {}
    orig = algo._preprocess_batch

    def hook(batch, buffer, indices):
        b = orig(batch, buffer, indices)
        captured["pre"] = {k: b[k].detach().cpu().numpy().copy() for k in ("v_s", "returns", "adv", "logp_old")}
        captured["indices"] = indices.cpu().numpy().copy()
        return b

    algo._preprocess_batch = hook
    for u in range(2):
        o = f"u{u}_"
        buf = restore_vector_buffer(g, o, E, cap, device=DEV)
        np.random.seed(1000 + u)
        with policy_within_training_step(algo.policy):
            stats = algo.update(buffer=buf, batch_size=bs, repeat=int(g["cfg_repeat"]))
        assert np.array_equal(captured["indices"], g[o + "indices"])
        pre = captured["pre"]
        np.testing.assert_allclose(pre["v_s"], g[o + "v_s"], rtol=2e-5, atol=5e-6)
        assert np.allclose(pre["returns"], g[o + "returns"], rtol=1e-4, atol=2e-5)
        assert np.allclose(pre["adv"], g[o + "adv"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(pre["logp_old"], g[o + "logp_old"], rtol=2e-5, atol=5e-6)
        assert stats.gradient_steps == int(g[o + "gradient_steps"])
        ref_losses = g[o + "losses"]
        for col, name in enumerate(["loss", "actor_loss", "vf_loss", "ent_loss"]):
            s = getattr(stats, name)
            np.testing.assert_allclose(s.mean, ref_losses[:, col].mean(), rtol=5e-4, atol=2e-5, err_msg=name)
        for k, pv in named_params(actor, critic).items():
            np.testing.assert_allclose(pv.detach().cpu().numpy(), g[o + "p_" + k], rtol=2e-3, atol=3e-5,
                                       err_msg=f"{variant} update {u} param {k}")
        if bool(g["kw_return_scaling"]):
            np.testing.assert_allclose([algo.ret_rms.mean, algo.ret_rms.var, algo.ret_rms.count], g[o + "rms"], rtol=1e-5)


def test_discrete_grad_kernel_vs_oracle():
    """ts_ppo_grad + ts_grad_reduce (C ABI) on one minibatch vs the numpy oracle's manual backward, entropy bonus and
    advantage normalisation on; shared trunk: both losses' gradients land in the same slots."""
    from oracle import oracle_discrete as od
    from tianshou_b200._cabi import call, ptr, stream_ptr
    g = load_golden("ppo_ref_C1b.npz")
    algo, actor, critic = build_discrete(g, DEV)
    rng = np.random.default_rng(0)
    n = 300
    obs = rng.standard_normal((n, 4)).astype(np.float32)
    act = rng.integers(0, 2, n)
    p = {k: g["p0_" + k].copy() for k in KEYS if "p0_" + k in g.files}
    probs, _, _ = od.actor_forward(p, obs)
    logp_old = od.categorical(probs, act)[0] + (0.3 * rng.standard_normal(n)).astype(np.float32)
    adv = rng.standard_normal(n).astype(np.float32)
    ret = rng.standard_normal(n).astype(np.float32)
    v_s = od.critic_forward(p, obs) + (0.1 * rng.standard_normal(n)).astype(np.float32)
    hp_np = dict(eps_clip=0.2, dual_clip=3.0, vf_coef=0.25, ent_coef=0.01, adv_eps=1e-8, value_clip=True,
                 advantage_normalization=True)
    grads, ls = od.minibatch_grad(p, dict(obs=obs, act=act, adv=adv, returns=ret, logp_old=logp_old, v_s=v_s), hp_np)
    t = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a)).to(DEV, dt)   # noqa: E731
    d_obs, d_act, d_adv, d_ret, d_lp, d_vs = t(obs), t(act.reshape(-1, 1)), t(adv), t(ret), t(logp_old), t(v_s)
    hp = algo._ppo_hparams()
    mom = torch.tensor([adv.mean(), adv.std(ddof=1)], dtype=torch.float32, device=DEV)
    f = algo._flat
    n_part = C.c_int32(0)
    call("ts_ppo_grad", ptr(f.flat), C.byref(algo._desc), C.byref(hp), ptr(d_obs), ptr(d_act), ptr(d_adv), ptr(d_ret),
         ptr(d_lp), ptr(d_vs), None, 0, n, n, ptr(mom), ptr(f.partials), C.byref(n_part), stream_ptr(torch.device(DEV)))
    call("ts_grad_reduce", ptr(f.partials), n_part.value, C.byref(algo._desc), ptr(f.grad), stream_ptr(torch.device(DEV)))
    gflat = f.grad.cpu().numpy()
    for k in grads:
        off = getattr(algo._desc, k)
        got = gflat[off: off + grads[k].size].reshape(grads[k].shape)
        np.testing.assert_allclose(got, grads[k], rtol=2e-4, atol=2e-6, err_msg=k)
    ex = gflat[algo._desc.n_params:]
    np.testing.assert_allclose([-ex[0] / n, ex[1] / n, ex[2] / n], ls[1:], rtol=2e-5, atol=1e-6)


def test_discrete_policy_forward_fused_inference():
    from tianshou_b200.data import Batch
    g = load_golden("ppo_ref_C1.npz")
    algo, actor, critic = build_discrete(g, DEV)
    pol = algo.policy
    obs = np.random.default_rng(1).standard_normal((77, 4)).astype(np.float32)
    with torch.no_grad():
        fused = pol(Batch(obs=obs, info=Batch())).logits
        pol.use_fused_inference = False
        ref = pol(Batch(obs=obs, info=Batch())).logits
    np.testing.assert_allclose(fused.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-6)
