"""Parity of the CUDA kernels (through the C ABI) against the oracle and the reference's outputs.

Bar: bit-exact for integer / index work and for f64 paths that keep the reference's operation
order (n-step, sum tree); 1e-5 relative (np.allclose, the reference's own test style) for the GAE
scan, which re-associates the f64 recurrence.
"""
import numpy as np
import pytest
import torch

from oracle import oracle_np as onp
from ts_testutil import load_golden

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def ops():
    from tianshou_b200 import ops as _ops
    return _ops


def meta_from(g, p):
    return ops().DeviceBufferMeta.from_host(g[p + "offset"], g[p + "done"], g[p + "last_index"], g[p + "lengths"], DEV)


# ---------------------------------------------------------------------------------------- GAE
def test_gae_vs_reference_outputs():
    g = load_golden("returns_ref.npz")
    for c in range(int(g["n_gae_cases"])):
        p = f"gae{c}_"
        idx = g[p + "indices"]
        unf = np.isin(idx, g[p + "unfinished"])
        term_mask = g[p + "buf_terminated"][idx]
        end = g[p + "terminated"] | g[p + "truncated"]
        adv, ret = ops().gae(g[p + "v_s"], g[p + "v_s_"], g[p + "rew"], term_mask, end, unf, gamma=float(g[p + "gamma"]),
                             gae_lambda=float(g[p + "lam"]), out_dtype=torch.float64, terminated_ends=False, device=DEV)
        assert np.allclose(ret.cpu().numpy(), g[p + "returns"], rtol=1e-5, atol=1e-8), c
        assert np.allclose(adv.cpu().numpy(), g[p + "adv"], rtol=1e-5, atol=1e-8), c
        # tighter than the bar: the f64 scan differs from the sequential f64 loop by rounding only
        np.testing.assert_allclose(adv.cpu().numpy(), g[p + "adv"], rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("n", [1, 7, 8, 2047, 2048, 2049, 10_000, 4096 * 128, 8192 * 256])   # ... C2, C5 (BASELINE configs[1], [4])
@pytest.mark.parametrize("vdtype", [np.float32, np.float64])
def test_gae_sizes_vs_oracle(n, vdtype):
    rng = np.random.default_rng(n)
    v_s, v_n = rng.standard_normal(n).astype(vdtype), rng.standard_normal(n).astype(vdtype)
    rew = rng.standard_normal(n)
    term = rng.random(n) < 0.01
    trunc = (rng.random(n) < 0.01) & ~term
    extra = np.zeros(n, dtype=bool)
    extra[-1] = True
    out_dtype = torch.float32 if vdtype == np.float32 else torch.float64
    adv, ret = ops().gae(v_s, v_n, rew, term, trunc, extra, gamma=0.99, gae_lambda=0.95, out_dtype=out_dtype, device=DEV)
    end = term | trunc | extra
    ref_adv = onp.gae(v_s, v_n * ~term, rew, end, 0.99, 0.95)
    assert np.allclose(adv.cpu().numpy(), ref_adv, rtol=1e-5, atol=1e-6)
    assert np.allclose(ret.cpu().numpy(), ref_adv + v_s, rtol=1e-5, atol=1e-6)


def test_gae_no_end_flags_long_chain():
    """No segment cut anywhere: the look-back must chain through every tile."""
    n = 50_000
    rng = np.random.default_rng(0)
    v_s, v_n, rew = (rng.standard_normal(n) for _ in range(3))
    adv, _ = ops().gae(v_s, v_n, rew, None, None, None, gamma=0.999, gae_lambda=0.99, out_dtype=torch.float64, device=DEV)
    ref = onp.gae(v_s, v_n, rew, np.zeros(n, dtype=bool), 0.999, 0.99)
    np.testing.assert_allclose(adv.cpu().numpy(), ref, rtol=1e-10, atol=1e-10)


def test_gae_return_scaling_and_running_stats():
    n = 30_000
    rng = np.random.default_rng(5)
    v_s, v_n = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    rew = rng.standard_normal(n) * 3 + 1
    term = rng.random(n) < 0.02
    trunc = np.zeros(n, dtype=bool)
    trunc[-1] = True
    rms = onp.RunningMeanStd()
    state = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64, device=DEV)
    for _ in range(3):   # statistics carry over between calls
        scale = np.sqrt(rms.var + 1e-8)
        ref_adv = onp.gae(v_s * scale, (v_n * scale) * ~term, rew, term | trunc, 0.99, 0.95)
        ref_ret = ref_adv + v_s * scale
        adv, ret = ops().gae(v_s, v_n, rew, term, trunc, None, gamma=0.99, gae_lambda=0.95, rms_state=state,
                             rms_eps=1e-8, device=DEV)
        assert np.allclose(adv.cpu().numpy(), ref_adv, rtol=1e-5, atol=1e-6)
        assert np.allclose(ret.cpu().numpy(), ref_ret / scale, rtol=1e-5, atol=1e-6)
        rms.update(ref_ret)
        np.testing.assert_allclose(state.cpu().numpy(), [rms.mean, rms.var, rms.count], rtol=1e-9)


def test_gae_empty():
    adv, ret = ops().gae(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0), None, None, None, gamma=0.9,
                         gae_lambda=0.9, device=DEV)
    assert adv.numel() == 0 and ret.numel() == 0


# ------------------------------------------------------------------------------------- indices
def test_index_kernels_vs_reference_outputs():
    g = load_golden("index_ref.npz")
    for c in range(int(g["n_cases"])):
        p = f"idx{c}_"
        m = meta_from(g, p)
        q = g[p + "query"]
        assert np.array_equal(ops().prev_index(m, q).cpu().numpy(), g[p + "prev"]), c
        assert np.array_equal(ops().next_index(m, q).cpu().numpy(), g[p + "next"]), c
        assert np.array_equal(ops().unfinished_index(m).cpu().numpy(), g[p + "unfinished"]), c
        assert np.array_equal(ops().sample_all_indices(m).cpu().numpy(), g[p + "all"]), c
        assert np.array_equal(ops().buffer_end_flags(m).cpu().numpy().astype(bool),
                              onp.buffer_end_flags(g[p + "done"], g[p + "last_index"], g[p + "lengths"])), c
        st = ops().stack_next_indices(m, g[p + "all"], 4).cpu().numpy()
        ref = [g[p + "all"]]
        for _ in range(3):
            ref.append(onp.next_index(ref[-1], g[p + "offset"], g[p + "done"], g[p + "last_index"], g[p + "lengths"]))
        assert np.array_equal(st, np.stack(ref)), c


def test_index_kernels_many_subbuffers_vs_oracle():
    """4096 sub-buffers (the reference loops over all of them per call)."""
    rng = np.random.default_rng(3)
    E, cap = 4096, 128
    offset = np.arange(E + 1, dtype=np.int64) * cap
    lengths = rng.integers(0, cap + 1, E)
    lengths[:10] = 0
    lengths[10:20] = cap
    last = offset[:-1] + np.where(lengths > 0, rng.integers(0, cap, E) % np.maximum(lengths, 1), 0)
    done = rng.random(E * cap) < 0.05
    m = ops().DeviceBufferMeta.from_host(offset, done, last, lengths, DEV)
    q = rng.integers(-5, E * cap + 5, 20_000)
    assert np.array_equal(ops().next_index(m, q).cpu().numpy(), onp.next_index(q, offset, done, last, lengths))
    assert np.array_equal(ops().prev_index(m, q).cpu().numpy(), onp.prev_index(q, offset, done, last, lengths))
    assert np.array_equal(ops().unfinished_index(m).cpu().numpy(), onp.unfinished_index(offset, done, last, lengths))
    assert np.array_equal(ops().sample_all_indices(m).cpu().numpy(), onp.sample_all_indices(offset, last, lengths))


def test_gather_and_mark():
    rng = np.random.default_rng(1)
    for shape, dtype in [((1000, 17), np.float32), ((1000, 4), np.float32), ((1000,), np.float64), ((1000,), np.uint8),
                         ((1000, 3), np.uint8)]:
        src = (rng.random(shape) * 100).astype(dtype)
        idx = rng.integers(0, 1000, 777)
        out = ops().gather_rows(torch.from_numpy(src).to(DEV), torch.from_numpy(idx).to(DEV))
        assert np.array_equal(out.cpu().numpy(), src[idx])
    idx = torch.from_numpy(rng.integers(0, 500, 300)).to(DEV)
    members = torch.tensor([3, 77, 499, 12], device=DEV)
    mark = ops().mark_members(idx, members, 500)
    assert np.array_equal(mark.cpu().numpy().astype(bool), np.isin(idx.cpu().numpy(), members.cpu().numpy()))


# --------------------------------------------------------------------------------------- n-step
def test_nstep_vs_reference_outputs():
    g = load_golden("returns_ref.npz")
    for c in range(int(g["n_nstep_cases"])):
        p = f"nstep{c}_"
        m = meta_from(g, p)
        n_step, gamma = int(g[p + "n_step"]), float(g[p + "gamma"])
        stacked = ops().stack_next_indices(m, g[p + "indices"], n_step)
        last = stacked[-1].cpu().numpy()
        tq = torch.from_numpy(g[p + "table"][last].copy()).to(DEV)
        term = torch.from_numpy(g[p + "buf_terminated"].view(np.uint8)).to(DEV)
        ops().value_mask_rows(tq, term, stacked[-1].contiguous())
        end_flag = ops().buffer_end_flags(m)
        rew = torch.from_numpy(g[p + "rew"]).to(DEV)
        out32 = ops().nstep_return(rew, end_flag, tq, stacked, gamma, n_step, torch.float32)
        assert np.array_equal(out32.cpu().numpy(), g[p + "returns"]), c     # bit-exact after the f32 rounding
        out64 = ops().nstep_return(rew, end_flag, tq, stacked, gamma, n_step, torch.float64).cpu().numpy()
        ref = onp.nstep_return(g[p + "rew"], end_flag.cpu().numpy().astype(bool), tq.cpu().numpy(),
                               stacked.cpu().numpy(), gamma, n_step)
        assert np.array_equal(out64, ref), c                                # bit-exact f64


def test_nstep_large_buffer_vs_oracle():
    rng = np.random.default_rng(11)
    B, I, A, n_step = 1_000_000, 256, 6, 3
    E = 64
    offset = np.arange(E + 1, dtype=np.int64) * (B // E)
    lengths = np.full(E, B // E)
    last = offset[:-1] + rng.integers(0, B // E, E)
    done = rng.random(offset[-1]) < 0.01
    rew = rng.standard_normal(offset[-1])
    m = ops().DeviceBufferMeta.from_host(offset, done, last, lengths, DEV)
    idx = rng.integers(0, offset[-1], I)
    stacked = ops().stack_next_indices(m, idx, n_step)
    tq = rng.standard_normal((I, A)).astype(np.float32)
    end_flag = ops().buffer_end_flags(m)
    out = ops().nstep_return(torch.from_numpy(rew).to(DEV), end_flag, torch.from_numpy(tq).to(DEV), stacked, 0.99,
                             n_step, torch.float64).cpu().numpy()
    rows = [idx]
    for _ in range(n_step - 1):
        rows.append(onp.next_index(rows[-1], offset, done, last, lengths))
    ref = onp.nstep_return(rew, onp.buffer_end_flags(done, last, lengths), tq, np.stack(rows), 0.99, n_step)
    assert np.array_equal(stacked.cpu().numpy(), np.stack(rows))
    assert np.array_equal(out, ref)


# ------------------------------------------------------------------------------------- sum tree
def test_segtree_vs_reference_outputs():
    from tianshou_b200.data import SegmentTree
    g = load_golden("segtree_ref.npz")
    for c in range(int(g["n_cases"])):
        size = int(g[f"seg{c}_size"])
        tree = SegmentTree(size, device=DEV)
        for r in range(4):
            p = f"seg{c}_r{r}_"
            tree[g[p + "idx"]] = g[p + "val"]
            assert np.array_equal(tree.tree.cpu().numpy(), g[p + "tree"]), (c, r)    # bit-exact f64 tree
            total = tree.reduce()
            assert total == g[p + "tree"][1]
            if total > 0:
                assert np.array_equal(tree.get_prefix_sum_idx(g[p + "u"] * total), g[p + "prefix_idx"]), (c, r)
                u = torch.from_numpy(g[p + "u"]).to(DEV)
                assert np.array_equal(tree.sample_device(u).cpu().numpy(), g[p + "prefix_idx"]), (c, r)
            lo, hi = g[p + "range"]
            assert tree.reduce(int(lo), int(hi)) == float(g[p + "range_sum"])
    tree = SegmentTree(6, device=DEV)
    tree[np.arange(6)] = np.array([0.0, 1.0, 0.5, 0.0, 0.0, 0.5])
    assert np.array_equal(tree.get_prefix_sum_idx(g["corner_q"]), g["corner_idx"])   # test_buffer.py:617-624


def test_segtree_randomised_like_reference_test():
    """test/base/test_buffer.py:553-633 style: random writes vs naive sums."""
    from tianshou_b200.data import SegmentTree
    rng = np.random.default_rng(0)
    size = 100
    tree = SegmentTree(size, device=DEV)
    naive = np.zeros(size)
    for _ in range(50):
        idx = rng.integers(0, size, rng.integers(1, 30))
        val = rng.random(len(idx))
        naive[idx] = val
        tree[idx] = val
        for _ in range(5):
            lo = int(rng.integers(0, size)); hi = int(rng.integers(lo, size)) + 1
            assert np.allclose(tree.reduce(lo, hi), naive[lo:hi].sum())
        assert np.allclose(tree.reduce(), naive.sum())
        assert np.allclose(tree[np.arange(size)], naive)
        q = rng.random(20) * naive.sum() * 0.999
        got = tree.get_prefix_sum_idx(q)
        cs = np.cumsum(naive)
        assert np.all(cs[got] >= q - 1e-9) and np.all(np.where(got > 0, cs[np.maximum(got - 1, 0)], 0.0) <= q + 1e-9)


def test_prioritized_buffer_stream_matches_reference():
    from tianshou_b200.data import Batch, PrioritizedVectorReplayBuffer
    from ts_testutil import synth_rollout
    g = load_golden("segtree_ref.npz")
    np.random.seed(3)
    buf = PrioritizedVectorReplayBuffer(64, 4, alpha=0.6, beta=0.4, device=DEV)
    for s in synth_rollout(np.random.default_rng(9), 4, 20, 3, 2, 0.1, 6):
        buf.add(Batch(**s), buffer_ids=np.arange(4))
    b, idx = buf.sample(16)
    assert np.array_equal(idx, g["per_idx0"])
    np.testing.assert_allclose(b.weight, g["per_w0"], rtol=1e-12)
    buf.update_weight(idx, g["per_td"])
    b, idx = buf.sample(16)
    assert np.array_equal(idx, g["per_idx1"])
    np.testing.assert_allclose(b.weight, g["per_w1"], rtol=1e-12)
    assert np.array_equal(buf.weight.tree.cpu().numpy(), g["per_tree"])
    np.testing.assert_allclose([buf._max_prio, buf._min_prio], g["per_minmax"], rtol=0, atol=0)
