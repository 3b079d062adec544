"""BASELINE configs[0] plumbing: the UNMODIFIED reference ``Collector`` + ``OnPolicyTrainer`` (imported from
/root/reference through oracle/ref_shim.py) drive this package's ``Batch`` / ``VectorReplayBuffer`` /
``DiscreteActorPolicy`` / ``OnPolicyAlgorithm`` on a 10-env CartPole-shaped environment for two epochs.

Runs on the build box only (the reference tree does not travel to the GPU box; there is no GPU here), so the numeric body
of ``_update_with_batch`` is the numpy oracle (tests may use it) -- everything AROUND it is the product: the collector's
``isinstance`` gates (collector.py:360,375-384) through ``tianshou_b200.compat``, ``buffer.add`` of the Collector's own
Batch objects (:927-930), ``reset_buffer`` (trainer.py:1131), ``Algorithm.update`` with its training-step guard and lr
scheduler (algorithm_base.py:586-631), the stats dataclass the trainer serialises (trainer.py:733-737,1122).
"""
import dataclasses

import numpy as np
import pytest
import torch

from oracle.ref_shim import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs the reference tree (build box only)")


class FakeCartPole:
    """CartPole-shaped stand-in (gymnasium is not installed): obs 4, 2 actions, reward 1 per step."""

    metadata: dict = {}
    spec = None

    def __init__(self, seed: int = 0):
        import gymnasium as gym        # the shim's stand-in spaces
        self.action_space = gym.spaces.Discrete(2)
        self.observation_space = gym.spaces.Box(-1.0, 1.0, (4,))
        self.rng = np.random.default_rng(seed)
        self.t = 0
        self.unwrapped = self

    def reset(self, seed=None, options=None):
        self.t = 0
        return self.rng.standard_normal(4).astype(np.float32), {}

    def step(self, a):
        self.t += 1
        term = bool(self.rng.random() < 0.05)
        return self.rng.standard_normal(4).astype(np.float32), 1.0, term, self.t >= 30 and not term, {}

    def close(self):
        pass

    def seed(self, s=None):
        return [s]

    def render(self):
        return None


def test_reference_collector_and_trainer_drive_b200_types():
    from oracle import oracle_discrete as od
    from oracle import oracle_np as onp
    from oracle.ref_shim import import_reference
    import_reference()
    import gymnasium as gym
    import tianshou.data.collector as refcol
    from tianshou.env import DummyVectorEnv
    from tianshou.trainer import OnPolicyTrainer, OnPolicyTrainerParams

    import tianshou_b200.compat as compat
    from tianshou_b200.algorithm import A2CTrainingStats, AdamOptimizerFactory, DiscreteActorPolicy, OnPolicyAlgorithm
    from tianshou_b200.algorithm.optim import LRSchedulerFactoryLinear
    from tianshou_b200.data import Batch, ReplayBufferManager, SequenceSummaryStats, VectorReplayBuffer
    from tianshou_b200.utils.net.common import ActorCritic, Net
    from tianshou_b200.utils.net.discrete import DiscreteActor, DiscreteCritic

    compat.install_into_reference()
    compat.install_into_reference()                      # idempotent
    assert ReplayBufferManager in refcol.ReplayBufferManager

    torch.manual_seed(0)
    np.random.seed(0)
    E = 10
    net = Net(state_shape=(4,), hidden_sizes=(64, 64))
    actor = DiscreteActor(preprocess_net=net, action_shape=2, softmax_output=True)
    critic = DiscreteCritic(preprocess_net=net)
    policy = DiscreteActorPolicy(actor=actor, dist_fn=lambda p: torch.distributions.Categorical(probs=p),
                                 action_space=gym.spaces.Discrete(2))

    class OraclePPO(OnPolicyAlgorithm):
        """This package's Algorithm skeleton with the numeric update body from the numpy oracle (CPU stand-in for the kernels)."""

        def __init__(self):
            super().__init__(policy=policy)
            self.critic = critic
            self.optim = self._create_optimizer(ActorCritic(actor, critic), AdamOptimizerFactory(lr=3e-4).with_lr_scheduler_factory(
                LRSchedulerFactoryLinear(max_epochs=2, epoch_num_steps=400, collection_step_num_env_steps=200)))
            self.calls, self.rms = [], onp.RunningMeanStd()
            self.m = self.v = None
            self.step = 0

        def _params(self):
            t1, t2 = [m for m in actor.preprocess.model.model if isinstance(m, torch.nn.Linear)]
            a3, c3 = actor.last.model[0], critic.last.model[0]
            return {"a_w1": t1.weight, "a_b1": t1.bias, "a_w2": t2.weight, "a_b2": t2.bias, "a_w3": a3.weight, "a_b3": a3.bias,
                    "c_w3": c3.weight, "c_b3": c3.bias}

        def _sample(self, buffer, sample_size):
            assert sample_size == 0 and isinstance(buffer, VectorReplayBuffer)
            idx = np.concatenate([np.arange(e * buffer._cap[e], e * buffer._cap[e] + buffer._sizes[e]) for e in range(E)])
            return buffer._meta[idx], idx                # un-wrapped sub-buffers: chronological order per env

        def _update_with_batch(self, batch, batch_size, repeat):
            assert self.policy.is_within_training_step and isinstance(batch, Batch)
            self.calls.append((len(batch), batch_size, repeat))
            tp = self._params()
            p = {k: v.detach().numpy().copy() for k, v in tp.items()}
            if self.m is None:
                self.m = {k: np.zeros_like(v) for k, v in p.items()}
                self.v = {k: np.zeros_like(v) for k, v in p.items()}
            n = len(batch)
            done = np.asarray(batch.terminated) | np.asarray(batch.truncated)
            unf = np.zeros(n, bool)
            ends = np.cumsum([self._last_sizes[e] for e in range(E)]) - 1
            unf[ends] = ~done[ends]
            roll = dict(obs=np.asarray(batch.obs, np.float32), obs_next=np.asarray(batch.obs_next, np.float32), act=np.asarray(batch.act),
                        rew=np.asarray(batch.rew, np.float64), terminated=np.asarray(batch.terminated), truncated=np.asarray(batch.truncated),
                        unfinished=unf)
            perms = [np.random.permutation(n) for _ in range(repeat)]
            hp = dict(eps_clip=0.2, dual_clip=None, vf_coef=0.5, ent_coef=0.0, max_grad_norm=0.5, adv_eps=1e-8, value_clip=False,
                      advantage_normalization=True, lr=self.optim._optim.param_groups[0]["lr"], beta1=0.9, beta2=0.999, adam_eps=1e-8,
                      weight_decay=0.0)
            res = od.ppo_update(p, self.m, self.v, self.step, roll, perms, batch_size, repeat, hp, None, 0.99, 0.95, False)
            self.step = res["step"]
            with torch.no_grad():
                for k, t in tp.items():
                    t.copy_(torch.from_numpy(p[k]))
            ls = res["losses"]
            return A2CTrainingStats(loss=SequenceSummaryStats.from_sequence(ls[:, 0]), actor_loss=SequenceSummaryStats.from_sequence(ls[:, 1]),
                                    vf_loss=SequenceSummaryStats.from_sequence(ls[:, 2]), ent_loss=SequenceSummaryStats.from_sequence(ls[:, 3]),
                                    gradient_steps=len(ls))

        def update(self, buffer, batch_size, repeat):
            self._last_sizes = buffer._sizes.copy()
            return super().update(buffer, batch_size, repeat)

    algo = OraclePPO()
    buf = VectorReplayBuffer(E * 40, E)
    envs = DummyVectorEnv([lambda i=i: FakeCartPole(i) for i in range(E)])
    collector = refcol.Collector(algo, envs, buf)                          # Algorithm -> .policy through the widened gate
    assert collector.policy is policy and collector.buffer is buf
    w0 = actor.last.model[0].weight.detach().clone()
    with pytest.raises(RuntimeError):                                       # update outside a training step (algorithm_base.py:612-617)
        algo.update(buf, 64, 1)
    params = OnPolicyTrainerParams(training_collector=collector, test_collector=None, max_epochs=2, epoch_num_steps=400,
                                   collection_step_num_env_steps=200, batch_size=64, update_step_num_repetitions=2,
                                   test_in_training=False, verbose=False, show_progress=False)
    result = OnPolicyTrainer(algo, params).run()
    assert len(algo.calls) == 4 and all(c[1:] == (64, 2) for c in algo.calls)          # 2 epochs x (400 / 200) update steps
    assert all(c[0] == 200 for c in algo.calls), algo.calls                             # the buffer was reset between collects
    assert not torch.equal(w0, actor.last.model[0].weight)                              # parameters moved
    assert algo.optim._optim.param_groups[0]["lr"] < 3e-4                               # lr scheduler stepped once per update()
    assert isinstance(dataclasses.asdict(result), dict)
    assert len(buf) == 0 or len(buf) <= E * 40
