"""Layered networks on the device for the off-policy update bodies.

A ``FusedStack`` is the kernel-side view of a torch module stack made of ``nn.Linear`` / ``nn.Conv2d`` /
``nn.ReLU`` / ``nn.Flatten`` (the reference's ``MLP`` / ``Net`` / ``ContinuousCritic`` / ``DQNet``,
utils/net/common.py:76-369, utils/net/continuous.py:96-238, env/atari/atari_network.py:60-122): every layer's
forward, input gradient and weight gradient is ONE ``ts_net_gemm`` launch (tcgen05, fp32-faithful), convolutions
run as implicit GEMM over im2col rows.  Parameters live in a ``FlatGroup`` (one flat fp32 buffer per optimiser,
``nn.Parameter``s are views of it) so Adam and the Polyak update are single kernels and ``state_dict()`` keeps
working.  There is no autograd graph and no eager-PyTorch path: unsupported layers raise ``UnsupportedModelError``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any

import torch
from torch import nn

from .._cabi import call, load_library, ptr, stream_ptr
from .flat_params import UnsupportedModelError, adam_hyperparams

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2


class FlatGroup:
    """Flat fp32 storage (parameters, gradient, Adam moments) of one optimiser's parameters."""

    def __init__(self, params: list[nn.Parameter], device: torch.device) -> None:
        self.params = list(params)
        self.device = device
        self.n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(self.n, dtype=torch.float32, device=device)
        self.grad = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.norm_scratch = torch.zeros(256, dtype=torch.float64, device=device)
        self.step = 0
        self.step_dev: torch.Tensor | None = None     # device-resident step counter (CUDA-graph mode: see adam_step_device)
        self._offsets: dict[int, int] = {}
        off = 0
        for p in self.params:
            self._offsets[id(p)] = off
            off += p.numel()
        self._ptrs: list[int] = []
        self.adopt()

    def offset(self, p: nn.Parameter) -> int:
        return self._offsets[id(p)]

    def view(self, buf: torch.Tensor, p: nn.Parameter) -> torch.Tensor:
        o = self._offsets[id(p)]
        return buf[o:o + p.numel()]

    def adopt(self) -> None:
        with torch.no_grad():
            for p in self.params:
                v = self.view(self.flat, p).view(p.shape)
                if p.data.data_ptr() != v.data_ptr():
                    v.copy_(p.data.to(self.device, torch.float32))
                    p.data = v
        self._ptrs = [p.data.data_ptr() for p in self.params]

    def ensure_adopted(self) -> None:
        if [p.data.data_ptr() for p in self.params] != self._ptrs:
            self.adopt()

    def adam_step(self, optimizer: torch.optim.Optimizer, max_grad_norm: float | None) -> None:
        """``Algorithm.Optimizer.step`` after backward: clip_grad_norm_ (optional) + Adam (algorithm_base.py:496-500)."""
        hp = adam_hyperparams(optimizer)
        self.sync_step_from_device()
        self.step += 1
        if self.step_dev is not None:
            self.step_dev.fill_(self.step)
        call("ts_adam_step", ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.n, self.step,
             hp["lr"], hp["beta1"], hp["beta2"], hp["adam_eps"], hp["weight_decay"], float(max_grad_norm or 0.0),
             ptr(self.norm_scratch), stream_ptr(self.device))

    def adam_step_device(self, optimizer: torch.optim.Optimizer, max_grad_norm: float | None) -> None:
        """Same step with the step number read from / advanced in DEVICE memory: nothing in the launch depends on host state,
        so it can live inside a captured CUDA graph (the host mirror ``self.step`` is re-read by ``sync_step_from_device``)."""
        hp = adam_hyperparams(optimizer)
        if self.step_dev is None:
            self.step_dev = torch.tensor([self.step], dtype=torch.int64, device=self.device)
        call("ts_adam_step_dev", ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.n, ptr(self.step_dev),
             hp["lr"], hp["beta1"], hp["beta2"], hp["adam_eps"], hp["weight_decay"], float(max_grad_norm or 0.0),
             ptr(self.norm_scratch), stream_ptr(self.device))

    def sync_step_from_device(self) -> None:
        if self.step_dev is not None:
            self.step = int(self.step_dev.item())

    def export_state(self, optimizer: torch.optim.Optimizer) -> None:
        self.sync_step_from_device()
        if self.step == 0 and len(optimizer.state) == 0:
            return
        for p in self.params:
            optimizer.state[p] = {"step": torch.tensor(float(self.step), dtype=torch.float32),
                                  "exp_avg": self.view(self.exp_avg, p).view(p.shape),
                                  "exp_avg_sq": self.view(self.exp_avg_sq, p).view(p.shape)}

    def import_state(self, optimizer: torch.optim.Optimizer) -> None:
        steps = []
        with torch.no_grad():
            for p in self.params:
                st = optimizer.state.get(p)
                m, v = self.view(self.exp_avg, p), self.view(self.exp_avg_sq, p)
                if not st:
                    m.zero_(); v.zero_()
                    continue
                m.copy_(st["exp_avg"].to(self.device, torch.float32).reshape(-1))
                v.copy_(st["exp_avg_sq"].to(self.device, torch.float32).reshape(-1))
                steps.append(float(st["step"]))
        if steps and max(steps) != min(steps):
            raise UnsupportedModelError("per-parameter Adam step counts differ; cannot fuse")
        self.step = int(round(steps[0])) if steps else 0
        if self.step_dev is not None:
            self.step_dev.fill_(self.step)


def polyak_update(target: FlatGroup, source: FlatGroup, tau: float) -> None:
    """``polyak_parameter_update`` (utils/lagged_network.py:8-18) on the flat buffers."""
    assert target.n == source.n
    call("ts_polyak_update", ptr(target.flat), ptr(source.flat), target.n, float(tau), stream_ptr(target.device))


@dataclass
class _Layer:
    kind: str                      # "linear" | "conv" | "flatten"
    weight: nn.Parameter | None = None
    bias: nn.Parameter | None = None
    act: int = ACT_NONE
    in_dim: int = 0
    out_dim: int = 0
    # conv: input NHWC [B, H, W, C] -> output NHWC [B, Ho, Wo, Cout]
    C: int = 0
    H: int = 0
    W: int = 0
    k: int = 0
    s: int = 0
    Ho: int = 0
    Wo: int = 0


def _activation_code(m: nn.Module) -> int:
    if type(m) is nn.ReLU:
        return ACT_RELU
    if type(m) is nn.Tanh:
        return ACT_TANH
    raise UnsupportedModelError(f"fused layered networks support ReLU / Tanh activations only, got {m}")


def compile_sequential(mods: list[nn.Module], input_shape: tuple[int, ...]) -> list[_Layer]:
    """Linear / Conv2d / ReLU / Flatten chain -> layer list.  ``input_shape`` = (features,) or (C, H, W)."""
    layers: list[_Layer] = []
    shape = tuple(int(x) for x in input_shape)
    for m in mods:
        if isinstance(m, nn.Sequential):
            sub = compile_sequential(list(m), shape)
            layers += sub
            shape = _out_shape(sub, shape)
            continue
        if isinstance(m, nn.Linear):
            if len(shape) != 1 or shape[0] != m.in_features:
                raise UnsupportedModelError(f"Linear({m.in_features}) after shape {shape}")
            if m.bias is None:
                raise UnsupportedModelError("Linear layers need a bias")
            layers.append(_Layer("linear", m.weight, m.bias, ACT_NONE, m.in_features, m.out_features))
            shape = (m.out_features,)
        elif isinstance(m, nn.Conv2d):
            if len(shape) != 3:
                raise UnsupportedModelError(f"Conv2d after shape {shape}")
            k, s = m.kernel_size, m.stride
            if (k[0] != k[1] or s[0] != s[1] or m.padding not in ((0, 0), 0) or m.dilation != (1, 1) or m.groups != 1
                    or m.bias is None or m.in_channels != shape[0]):
                raise UnsupportedModelError(f"unsupported Conv2d configuration {m}")
            Cc, H, W = shape
            Ho, Wo = (H - k[0]) // s[0] + 1, (W - k[0]) // s[0] + 1
            layers.append(_Layer("conv", m.weight, m.bias, ACT_NONE, Cc * k[0] * k[0], m.out_channels, C=Cc, H=H, W=W, k=k[0],
                                 s=s[0], Ho=Ho, Wo=Wo))
            shape = (m.out_channels, Ho, Wo)
        elif isinstance(m, nn.Flatten):
            if len(shape) == 3:
                layers.append(_Layer("flatten", C=shape[0], H=shape[1], W=shape[2], in_dim=shape[0] * shape[1] * shape[2],
                                     out_dim=shape[0] * shape[1] * shape[2]))
                shape = (shape[0] * shape[1] * shape[2],)
        elif isinstance(m, (nn.ReLU, nn.Tanh, nn.Sigmoid, nn.GELU, nn.ELU, nn.LeakyReLU)):
            code = _activation_code(m)
            if not layers or layers[-1].kind == "flatten" or layers[-1].act != ACT_NONE:
                raise UnsupportedModelError("activation without a producing layer")
            layers[-1].act = code
        elif isinstance(m, nn.Identity):
            continue
        else:
            raise UnsupportedModelError(f"layer {m} is outside the fused layered-network family")
    return layers


def _out_shape(layers: list[_Layer], shape: tuple[int, ...]) -> tuple[int, ...]:
    for L in layers:
        if L.kind == "linear":
            shape = (L.out_dim,)
        elif L.kind == "conv":
            shape = (L.out_dim, L.Ho, L.Wo)
        else:
            shape = (L.out_dim,)
    return shape


class FusedStack:
    """Forward / backward of a layer list on ``[rows, features]`` fp32 matrices (NHWC between conv layers)."""

    def __init__(self, layers: list[_Layer], group: FlatGroup, name: str = "net") -> None:
        self.layers = layers
        self.group = group
        self.name = name
        self.device = group.device
        self._bufs: dict[tuple, torch.Tensor] = {}
        self._ws_sizes: dict[tuple[int, int, int], int] = {}
        self._lib = load_library()

    # ------------------------------------------------------------------ scratch
    def _buf(self, key: tuple, n: int) -> torch.Tensor:
        t = self._bufs.get(key)
        if t is None or t.numel() < n:
            t = self._bufs[key] = torch.empty(max(n, 1), dtype=torch.float32, device=self.device)
        return t

    def _gemm(self, a, lda, a_mn, b, ldb, b_mn, c, ldc, M, N, K, bias=None, act=ACT_NONE, mask=None, ld_mask=0,
              mask_kind=ACT_RELU, accumulate=False) -> None:
        ws_n = self._ws_sizes.get((M, N, K))
        if ws_n is None:
            ws_n = self._ws_sizes[(M, N, K)] = int(self._lib.ts_net_gemm_workspace_floats(M, N, K))
        ws = self._buf(("ws",), ws_n) if ws_n > 0 else None
        call("ts_net_gemm", a, lda, a_mn, b, ldb, b_mn, c, ldc, M, N, K, bias, act, mask, ld_mask, int(mask_kind), int(accumulate),
             ptr(ws) if ws is not None else None, ws_n, stream_ptr(self.device))

    def _w(self, L: _Layer, flat: torch.Tensor | None = None) -> int:
        g = self.group
        base = (flat if flat is not None else g.flat).data_ptr()
        return base + 4 * g.offset(L.weight)

    def _b(self, L: _Layer, flat: torch.Tensor | None = None) -> int:
        g = self.group
        base = (flat if flat is not None else g.flat).data_ptr()
        return base + 4 * g.offset(L.bias)

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor | None, rows: int, tag: str = "a", *, frames: tuple | None = None,
                params: torch.Tensor | None = None) -> list[torch.Tensor]:
        """Returns the activation list ``[x0, y1, ..., yL]`` (``y_i`` = post-activation output of layer i; for conv layers
        rows * Ho * Wo NHWC rows).  ``frames`` = (uint8 frames, stack_idx int64 [rows, C], denom) feeds the first conv layer
        straight from single-frame storage (frame-stack gather + im2col in one kernel).  ``params``: evaluate with another
        flat parameter buffer of the same layout (the lagged / target copy)."""
        self.group.ensure_adopted()
        st = stream_ptr(self.device)
        acts: list[torch.Tensor] = [x]
        cur = x
        for i, L in enumerate(self.layers):
            if L.kind == "linear":
                y = self._buf((tag, "y", i), rows * L.out_dim)[: rows * L.out_dim].view(rows, L.out_dim)
                self._gemm(ptr(cur), L.in_dim, 0, self._w(L, params), L.in_dim, 0, ptr(y), L.out_dim, rows, L.out_dim, L.in_dim,
                           bias=self._b(L, params), act=L.act)
            elif L.kind == "conv":
                R = rows * L.Ho * L.Wo
                col = self._buf((tag, "col", i), R * L.in_dim)[: R * L.in_dim].view(R, L.in_dim)
                if i == 0 and frames is not None:
                    fr, sidx, denom = frames
                    call("ts_im2col_u8", ptr(fr), ptr(sidx), rows, L.C, L.H, L.W, L.k, L.s, float(denom), ptr(col), st)
                else:
                    call("ts_im2col_f32", ptr(cur), rows, L.C, L.H, L.W, L.k, L.s, ptr(col), st)
                y = self._buf((tag, "y", i), R * L.out_dim)[: R * L.out_dim].view(R, L.out_dim)
                self._gemm(ptr(col), L.in_dim, 0, self._w(L, params), L.in_dim, 0, ptr(y), L.out_dim, R, L.out_dim, L.in_dim,
                           bias=self._b(L, params), act=L.act)
            else:  # flatten NHWC -> NCHW order
                y = self._buf((tag, "y", i), rows * L.out_dim)[: rows * L.out_dim].view(rows, L.out_dim)
                call("ts_nhwc_to_nchw_flat", ptr(cur), rows, L.H * L.W, L.C, ptr(y), st)
            acts.append(y)
            cur = y
        return acts

    # ------------------------------------------------------------------ backward
    def backward(self, acts: list[torch.Tensor], dy: torch.Tensor, rows: int, tag: str = "a", *, param_grads: bool = True,
                 input_grad: bool = False, input_cols: tuple[int, int] | None = None, dy_preact: bool = False,
                 input_act: tuple[int, torch.Tensor] | None = None, dx_out: torch.Tensor | None = None,
                 dx_accumulate: bool = False, grad_accumulate: bool = False) -> torch.Tensor | None:
        """Back-propagate ``dy`` = gradient w.r.t. the LAST layer's output (its activation must be none, or ``dy_preact``: the
        caller already folded the last activation's derivative in).  Weight / bias gradients are STORED into the group's
        gradient buffer (``zero_grad`` + ``backward`` of algorithm_base.py:497-498) or added with ``grad_accumulate`` (a trunk
        shared by two losses).  With ``input_grad`` returns d loss / d input (columns ``input_cols`` of the first Linear's
        input), multiplied by the derivative of the activation ``input_act = (kind, y)`` that produced this stack's input (a
        head on top of a trunk), written to / accumulated into ``dx_out`` when given."""
        g = self.group
        st = stream_ptr(self.device)
        n = len(self.layers)
        if self.layers[-1].act != ACT_NONE and not dy_preact:
            raise UnsupportedModelError("backward expects a linear output layer")
        dz = dy                                   # gradient w.r.t. the pre-activation of layer i (mask already applied)
        for i in range(n - 1, -1, -1):
            L = self.layers[i]
            x_in = acts[i]
            prev_act = self._producer_act(i) if i > 0 else (input_act[0] if input_act is not None else ACT_NONE)
            act_src = x_in if i > 0 else (input_act[1] if input_act is not None else None)
            need_dx = i > 0 or input_grad
            acc = int(grad_accumulate)
            if L.kind == "linear":
                M_rows = rows
                if param_grads:
                    gw = g.grad.data_ptr() + 4 * g.offset(L.weight)
                    gb = g.grad.data_ptr() + 4 * g.offset(L.bias)
                    self._gemm(ptr(dz), L.out_dim, 1, ptr(x_in), L.in_dim, 1, gw, L.in_dim, L.out_dim, L.in_dim, M_rows,
                               accumulate=grad_accumulate)
                    call("ts_net_colsum", ptr(dz), L.out_dim, M_rows, L.out_dim, gb, acc, st)
                if need_dx:
                    lo, hi = (0, L.in_dim) if (i > 0 or input_cols is None) else input_cols
                    width = hi - lo
                    if i == 0 and dx_out is not None:
                        dx = dx_out
                    else:
                        dx = self._buf((tag, "dx", i), M_rows * width)[: M_rows * width].view(M_rows, width)
                    mask = ptr(act_src) if prev_act != ACT_NONE else None
                    self._gemm(ptr(dz), L.out_dim, 0, self._w(L) + 4 * lo, L.in_dim, 1, ptr(dx), width, M_rows, width, L.out_dim,
                               mask=mask, ld_mask=L.in_dim, mask_kind=prev_act if prev_act != ACT_NONE else ACT_RELU,
                               accumulate=(i == 0 and dx_accumulate))
                    dz = dx
            elif L.kind == "conv":
                if prev_act == ACT_TANH:
                    raise UnsupportedModelError("tanh in front of a convolution is not supported")
                R = rows * L.Ho * L.Wo
                col = self._bufs[(tag, "col", i)][: R * L.in_dim].view(R, L.in_dim)
                if param_grads:
                    gw = g.grad.data_ptr() + 4 * g.offset(L.weight)
                    gb = g.grad.data_ptr() + 4 * g.offset(L.bias)
                    self._gemm(ptr(dz), L.out_dim, 1, ptr(col), L.in_dim, 1, gw, L.in_dim, L.out_dim, L.in_dim, R,
                               accumulate=grad_accumulate)
                    call("ts_net_colsum", ptr(dz), L.out_dim, R, L.out_dim, gb, acc, st)
                if i > 0:
                    dcol = self._buf((tag, "dcol", i), R * L.in_dim)[: R * L.in_dim].view(R, L.in_dim)
                    self._gemm(ptr(dz), L.out_dim, 0, self._w(L), L.in_dim, 1, ptr(dcol), L.in_dim, R, L.in_dim, L.out_dim)
                    dx = self._buf((tag, "dx", i), rows * L.H * L.W * L.C)[: rows * L.H * L.W * L.C]
                    call("ts_col2im_f32", ptr(dcol), rows, L.C, L.H, L.W, L.k, L.s, ptr(x_in) if prev_act == ACT_RELU else None,
                         ptr(dx), st)
                    dz = dx
                elif input_grad:
                    raise UnsupportedModelError("input gradients through the first convolution are not provided")
            else:  # flatten: NCHW-flat gradient back to NHWC rows (+ the producer's ReLU mask)
                if prev_act == ACT_TANH:
                    raise UnsupportedModelError("tanh in front of a flatten is not supported")
                dx = self._buf((tag, "dx", i), rows * L.out_dim)[: rows * L.out_dim]
                call("ts_nchw_flat_to_nhwc", ptr(dz), rows, L.H * L.W, L.C, ptr(x_in) if prev_act == ACT_RELU else None, ptr(dx), st)
                dz = dx
        return dz if input_grad else None

    def _producer_act(self, i: int) -> int:
        """Activation that produced the input of layer i (looking through a flatten)."""
        j = i - 1
        while j >= 0 and self.layers[j].kind == "flatten":
            j -= 1
        return self.layers[j].act if j >= 0 else ACT_NONE


def module_layers(mod: Any) -> list[nn.Module]:
    """Flat module list of the reference-shaped containers: MLP / Net (``.model`` chains) or a plain Sequential."""
    if isinstance(mod, nn.Sequential):
        return list(mod)
    inner = getattr(mod, "model", None)
    if inner is not None:
        return module_layers(inner)
    net = getattr(mod, "net", None)            # DQNet
    if isinstance(net, nn.Sequential):
        return list(net)
    raise UnsupportedModelError(f"cannot read a layer chain from {type(mod).__name__}")
