"""Asynchronous device mirror of a replay buffer's transition arrays.

The reference uploads nothing: its update reads a full fancy-indexed copy of the buffer
(``buffer.sample(0)``, buffer_base.py:605-649, ~95 MB at 4096 x 128) every ``update()``.  Without a
mirror this package does the same upload once per update (a2c.py ``_sample``).  With
``VectorReplayBuffer(..., device_mirror=True)`` every ``add()`` additionally ships the E rows it just
wrote to the device -- pinned staging, one async H2D per key on a side stream, one scatter kernel
(``ts_scatter_rows``) to the rows' slots -- so that by the time the rollout is complete the device copy is
complete too and ``update()`` starts without any bulk transfer.  Host numpy storage stays the source of
truth (the reference ``Collector`` keeps reading / writing it).

Validity: the mirror tracks the buffer's mutation counter.  Every mutation that does not go through
``add()`` / ``reset()`` (``set_batch``, ``update``, ``set_array_at_key``, ``dropnull``, unpickling) makes it
stale and the update falls back to the bulk upload until ``resync()``.  In-place edits of the numpy arrays
obtained through attribute access (``buf.rew[3] = 0``) are NOT detectable -- call ``resync()`` after them.
"""
from __future__ import annotations

from typing import TYPE_CHECKING, Any

import numpy as np
import torch

from ..._cabi import call, ptr
from ..batch import Batch

if TYPE_CHECKING:
    from .base import ReplayBuffer

MIRROR_DTYPES: dict[str, torch.dtype] = {
    "obs": torch.float32, "act": torch.float32, "rew": torch.float64, "terminated": torch.bool,
    "truncated": torch.bool, "done": torch.bool, "obs_next": torch.float32,
}
_NP = {torch.float32: np.float32, torch.float64: np.float64, torch.bool: np.bool_}


class DeviceMirror:
    def __init__(self, buffer: "ReplayBuffer", slots: int = 4) -> None:
        self.buffer = buffer
        self.device = buffer.device
        self.stream = torch.cuda.Stream(device=self.device)
        self.cols: dict[str, torch.Tensor] = {}
        self.supported = True
        self.expected_version = -1
        self._slots = slots
        self._ring: list[dict[str, Any]] = []
        self._next = 0
        self.pushed_rows = 0

    # ------------------------------------------------------------------ allocation
    def _keys(self) -> list[str]:
        meta = self.buffer._meta
        keys = [k for k in MIRROR_DTYPES if k in meta.get_keys()]
        for k in keys:
            if isinstance(meta[k], Batch) or not isinstance(meta[k], np.ndarray) or meta[k].dtype == object:
                self.supported = False      # dict observations etc.: not a dense array, no mirror
        return keys

    def _ensure(self, rows: int) -> bool:
        """(Re)allocate device columns / staging when the host layout changed.  Returns True when the device
        columns were (re)created and therefore need a full resync."""
        meta, fresh = self.buffer._meta, False
        for k in self._keys():
            if not self.supported:
                return False
            shape = (self.buffer.maxsize, *meta[k].shape[1:])
            t = self.cols.get(k)
            # uint8 observations (Atari frames, 1 M x 84 x 84) stay uint8: 4x less HBM, converted inside the im2col gather
            dtype = torch.uint8 if meta[k].dtype == np.uint8 else MIRROR_DTYPES[k]
            if t is None or tuple(t.shape) != shape or t.dtype != dtype:
                self.cols[k] = torch.zeros(shape, dtype=dtype, device=self.device)
                self._ring, fresh = [], True
        if not self._ring or self._ring[0]["rows"] < rows:
            self._ring = []
            for _ in range(self._slots):
                slot: dict[str, Any] = {"rows": rows, "event": None, "host": {}, "dev": {}}
                for k, t in self.cols.items():
                    slot["host"][k] = torch.empty((rows, *t.shape[1:]), dtype=t.dtype, pin_memory=True)
                    slot["dev"][k] = torch.empty((rows, *t.shape[1:]), dtype=t.dtype, device=self.device)
                slot["host"]["_idx"] = torch.empty(rows, dtype=torch.int64, pin_memory=True)
                slot["dev"]["_idx"] = torch.empty(rows, dtype=torch.int64, device=self.device)
                self._ring.append(slot)
        return fresh

    # ------------------------------------------------------------------ data path
    def push(self, idx: np.ndarray, batch: Batch) -> None:
        """Ship the rows ``batch`` that ``add()`` just wrote at absolute slots ``idx``."""
        if not self.supported:
            return
        n = len(idx)
        was_valid = self.valid_before_add()
        if self._ensure(max(n, self.buffer.buffer_num)) or not was_valid:
            if self.supported:
                self.resync()           # first add / layout change / stale mirror: one bulk upload
            return
        if not self.supported:
            return
        slot = self._ring[self._next]
        self._next = (self._next + 1) % len(self._ring)
        if slot["event"] is not None:
            slot["event"].synchronize()     # the staging memory of this slot is free again (normally long ago)
        slot["host"]["_idx"][:n].numpy()[...] = idx
        for k in self.cols:
            dst = slot["host"][k][:n].numpy()
            a = np.asarray(batch[k])
            a = a.reshape(dst.shape) if a.size == dst.size else a[:n].reshape(dst.shape)   # un-stacked single transition
            np.copyto(dst, a, casting="unsafe")
        with torch.cuda.stream(self.stream):
            slot["dev"]["_idx"][:n].copy_(slot["host"]["_idx"][:n], non_blocking=True)
            for k, dst in self.cols.items():
                src = slot["dev"][k]
                src[:n].copy_(slot["host"][k][:n], non_blocking=True)
                row_bytes = (dst[0].numel() if dst.dim() > 1 else 1) * dst.element_size()
                call("ts_scatter_rows", ptr(src), row_bytes, ptr(slot["dev"]["_idx"]), n, ptr(dst),
                     self.stream.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            slot["event"] = ev
        self.pushed_rows += n
        self.expected_version = self.buffer._version

    def valid_before_add(self) -> bool:
        # add() bumps the version exactly once (in _advance) before it calls push()
        return self.expected_version == self.buffer._version - 1 and bool(self.cols)

    def resync(self) -> None:
        """Bulk upload of every mirrored column from the host arrays (blocking on the side stream only)."""
        if not self.supported:
            return
        self._ensure(self.buffer.buffer_num)
        if not self.supported:
            return
        meta = self.buffer._meta
        with torch.cuda.stream(self.stream):
            for k, dst in self.cols.items():
                host = torch.from_numpy(np.ascontiguousarray(meta[k]))
                dst.copy_(host.reshape(dst.shape), non_blocking=False)
        self.expected_version = self.buffer._version

    def on_reset(self) -> None:
        """``reset()`` only rewinds the bookkeeping; the arrays (host and device) keep their contents."""
        if self.cols:
            self.expected_version = self.buffer._version

    def columns(self) -> dict[str, torch.Tensor] | None:
        """The device columns if they reflect the host buffer, ordered after every pending copy; else None."""
        if not (self.supported and self.cols and self.expected_version == self.buffer._version):
            return None
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self.cols
