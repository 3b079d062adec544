"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float32, manual backward) of the reference's PPO update for its
own discrete test network (BASELINE configs[0], test/discrete/test_ppo_discrete.py:90-125):

  Net(obs -> 64 -> 64, ReLU)  [ONE instance shared by actor and critic, or two]   (utils/net/common.py:223-369)
  DiscreteActor: Linear(64 -> A) -> softmax                                        (utils/net/discrete.py:29-92)
  DiscreteCritic: Linear(64 -> 1)                                                  (utils/net/discrete.py:94-123)
  dist = torch.distributions.Categorical(probs)                                    (reinforce.py:167-192)

Pinned against outputs of the imported reference: tests/golden/ppo_ref_C1*.npz (oracle/gen_golden.py
``ppo_discrete``), checked by tests/test_oracle.py.  Only tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline legs may import this module.

Parameter dict: a_w1, a_b1, a_w2, a_b2 (actor trunk), a_w3, a_b3 (logits head), c_w3, c_b3 (value head) and, for
separate trunks, c_w1, c_b1, c_w2, c_b2.  A shared trunk is the absence of the c_* trunk keys.
"""
from __future__ import annotations

import numpy as np

from . import oracle_np as onp

F = np.float32
EPS = np.finfo(np.float32).eps


def _ctrunk(p):
    return (p["c_w1"], p["c_b1"], p["c_w2"], p["c_b2"]) if "c_w1" in p else (p["a_w1"], p["a_b1"], p["a_w2"], p["a_b2"])


def trunk_forward(x, w1, b1, w2, b2):
    h1 = np.maximum(x @ w1.T + b1, F(0))
    h2 = np.maximum(h1 @ w2.T + b2, F(0))
    return h1, h2


def critic_forward(p, obs):
    _, h2 = trunk_forward(obs.astype(F), *_ctrunk(p))
    return (h2 @ p["c_w3"].T + p["c_b3"]).reshape(-1)


def actor_forward(p, obs):
    """-> (probs [n, A] = softmax(logits), h1, h2); torch.softmax subtracts the row max."""
    h1, h2 = trunk_forward(obs.astype(F), p["a_w1"], p["a_b1"], p["a_w2"], p["a_b2"])
    z = h2 @ p["a_w3"].T + p["a_b3"]
    e = np.exp(z - z.max(axis=1, keepdims=True)).astype(F)
    return (e / e.sum(axis=1, keepdims=True, dtype=F)).astype(F), h1, h2


def categorical(probs, act):
    """torch.distributions.Categorical(probs): renormalise, logits = log(clamp(p, eps, 1 - eps))
    (torch/distributions/categorical.py, utils.probs_to_logits).  -> (log_prob(act), entropy, pn, logits, s2)."""
    s2 = probs.sum(axis=1, keepdims=True, dtype=F)
    pn = (probs / s2).astype(F)
    lg = np.log(np.clip(pn, EPS, F(1.0) - EPS)).astype(F)
    idx = np.asarray(act).astype(np.int64).reshape(-1)
    logp = lg[np.arange(len(idx)), idx]
    ent = -(pn * lg).sum(axis=1, dtype=F)
    return logp.astype(F), ent.astype(F), pn, lg, s2.astype(F)


def add_returns_and_advantages(p, rollout, rms, gamma, lam, eps=1e-8):
    """a2c.py:115-153 with this network's critic."""
    v_s = critic_forward(p, rollout["obs"])
    v_s_ = critic_forward(p, rollout["obs_next"])
    vs_np, vn_np = v_s, v_s_
    if rms is not None:
        scale = np.sqrt(rms.var + eps)
        vs_np, vn_np = v_s * scale, v_s_ * scale
    ret, adv = onp.compute_episodic_return(
        rollout["rew"], rollout["terminated"].copy(), rollout["truncated"], rollout["unfinished"],
        ~rollout["terminated"], vn_np, vs_np, gamma, lam)
    if rms is not None:
        returns = ret / np.sqrt(rms.var + eps)
        rms.update(ret)
    else:
        returns = ret
    return v_s.astype(F), returns.astype(F), adv.astype(F)


def minibatch_grad(p, mb, hp):
    """Loss and gradients of one minibatch (ppo.py:179-211 + autograd through Categorical / softmax / ReLU)."""
    n = len(mb["adv"])
    B = F(n)
    obs = mb["obs"].astype(F)
    act = np.asarray(mb["act"]).astype(np.int64).reshape(-1)
    adv = mb["adv"].astype(F)
    if hp["advantage_normalization"]:
        mean, std = adv.mean(dtype=F), adv.std(ddof=1, dtype=F)
        adv = (adv - mean) / (std + F(hp["adv_eps"]))
    probs, ah1, ah2 = actor_forward(p, obs)
    logp, ent, pn, lg, s2 = categorical(probs, act)
    ratio = np.exp(logp - mb["logp_old"]).astype(F)
    lo, hi = F(1.0 - hp["eps_clip"]), F(1.0 + hp["eps_clip"])
    rc = np.clip(ratio, lo, hi)
    in_range = (ratio >= lo) & (ratio <= hi)
    surr1, surr2 = ratio * adv, rc * adv
    clip1 = np.minimum(surr1, surr2)
    g_ratio = np.where(surr1 < surr2, adv, np.where(surr1 > surr2, np.where(in_range, adv, F(0)),
                                                    np.where(in_range, adv, F(0.5) * adv))).astype(F)
    obj = clip1
    if hp["dual_clip"]:
        c2 = F(hp["dual_clip"]) * adv
        clip2 = np.maximum(clip1, c2)
        neg = adv < 0
        obj = np.where(neg, clip2, clip1)
        g_ratio = np.where(neg & (clip1 < c2), F(0), np.where(neg & (clip1 == c2), F(0.5) * g_ratio, g_ratio))
    clip_loss = -obj.mean(dtype=F)
    ent_loss = ent.mean(dtype=F)
    gl = (-g_ratio * ratio / B).astype(F)              # d loss / d log_prob
    ge = F(-hp["ent_coef"]) / B                          # d loss / d entropy_row
    # autograd chain: gather + entropy -> log o clamp -> renormalisation -> softmax
    onehot = np.zeros_like(pn)
    onehot[np.arange(n), act] = F(1)
    dlg = onehot * gl[:, None] - ge * pn
    passes = (pn >= EPS) & (pn <= F(1.0) - EPS)
    dpn = -ge * lg + np.where(passes, dlg / pn, F(0))
    dp = (dpn - (dpn * pn).sum(axis=1, keepdims=True, dtype=F)) / s2
    dz = (probs * (dp - (dp * probs).sum(axis=1, keepdims=True, dtype=F))).astype(F)
    # critic
    ch1, ch2 = trunk_forward(obs, *_ctrunk(p))
    value = (ch2 @ p["c_w3"].T + p["c_b3"]).reshape(-1)
    R, vs = mb["returns"].astype(F), mb["v_s"].astype(F)
    if hp["value_clip"]:
        eps = F(hp["eps_clip"])
        dlt = value - vs
        v_clip = vs + np.clip(dlt, -eps, eps)
        e1, e2 = R - value, R - v_clip
        vf1, vf2 = e1 * e1, e2 * e2
        vf = np.maximum(vf1, vf2)
        inr = ((dlt >= -eps) & (dlt <= eps)).astype(F)
        g1, g2 = F(-2.0) * e1, F(-2.0) * e2 * inr
        g = np.where(vf1 > vf2, g1, np.where(vf1 < vf2, g2, F(0.5) * (g1 + g2)))
    else:
        e1 = R - value
        vf = e1 * e1
        g = F(-2.0) * e1
    vf_loss = vf.mean(dtype=F)
    dv = (F(hp["vf_coef"]) * g / B).astype(F)[:, None]
    loss = clip_loss + F(hp["vf_coef"]) * vf_loss - F(hp["ent_coef"]) * ent_loss

    def trunk_backward(dout, w3, h1, h2, w2, x):
        gw3 = dout.T @ h2
        gb3 = dout.sum(axis=0)
        dz2 = (dout @ w3) * (h2 > 0)
        gw2 = dz2.T @ h1
        gb2 = dz2.sum(axis=0)
        dz1 = (dz2 @ w2) * (h1 > 0)
        gw1 = dz1.T @ x
        gb1 = dz1.sum(axis=0)
        return [a.astype(F) for a in (gw1, gb1, gw2, gb2, gw3, gb3)]

    cw1, cb1, cw2, cb2 = _ctrunk(p)
    ga = trunk_backward(dz, p["a_w3"], ah1, ah2, p["a_w2"], obs)
    gc = trunk_backward(dv, p["c_w3"], ch1, ch2, cw2, obs)
    grads = dict(zip(["a_w1", "a_b1", "a_w2", "a_b2", "a_w3", "a_b3"], ga, strict=True))
    if "c_w1" in p:
        grads.update(dict(zip(["c_w1", "c_b1", "c_w2", "c_b2"], gc[:4], strict=True)))
    else:       # shared trunk: both losses flow into the same parameters
        for k, gk in zip(["a_w1", "a_b1", "a_w2", "a_b2"], gc[:4], strict=True):
            grads[k] = grads[k] + gk
    grads["c_w3"], grads["c_b3"] = gc[4], gc[5]
    return grads, (float(loss), float(clip_loss), float(vf_loss), float(ent_loss))


def ppo_update(p, m, v, step, rollout, perms, batch_size, repeat, hp, rms, gamma, lam, recompute_adv):
    """``PPO._preprocess_batch`` + ``_update_with_batch`` (ppo.py:146-224) for this network family."""
    v_s, returns, adv = add_returns_and_advantages(p, rollout, rms, gamma, lam)
    probs, _, _ = actor_forward(p, rollout["obs"])
    logp_old = categorical(probs, rollout["act"])[0]
    first = dict(v_s=v_s.copy(), returns=returns.copy(), adv=adv.copy(), logp_old=logp_old.copy())
    n = len(adv)
    losses = []
    for r in range(repeat):
        if recompute_adv and r > 0:
            v_s, returns, adv = add_returns_and_advantages(p, rollout, rms, gamma, lam)
        for lo, hi in onp.minibatch_bounds(n, batch_size or n):
            idx = perms[r][lo:hi]
            mb = dict(obs=rollout["obs"][idx], act=rollout["act"][idx], adv=adv[idx], returns=returns[idx],
                      logp_old=logp_old[idx], v_s=v_s[idx])
            grads, ls = minibatch_grad(p, mb, hp)
            step, _ = onp.clip_adam_step(p, grads, m, v, step, hp)
            losses.append(ls)
    return dict(first=first, losses=np.array(losses), step=step, v_s=v_s, returns=returns, adv=adv)
