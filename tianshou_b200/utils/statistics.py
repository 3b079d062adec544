"""Running mean / variance (API of tianshou/utils/statistics.py:68-114).

The host ``update`` is kept for API users; inside the fused PPO update the same Chan merge is
executed by the GAE kernel on a 3-double device state (csrc/gae.cu) and mirrored back here with
``load_device_state`` once per ``update()`` call.
"""
from __future__ import annotations

import numpy as np
import torch


class RunningMeanStd:
    def __init__(
        self,
        mean: float | np.ndarray = 0.0,
        std: float | np.ndarray = 1.0,
        clip_max: float | None = 10.0,
        epsilon: float = np.finfo(np.float32).eps.item(),
    ) -> None:
        # NB: the reference stores `std` into `var` (statistics.py:88); kept for parity.
        self.mean, self.var = mean, std
        self.clip_max = clip_max
        self.count = 0
        self.eps = epsilon

    def norm(self, data_array: float | np.ndarray) -> float | np.ndarray:
        data_array = (data_array - self.mean) / np.sqrt(self.var + self.eps)
        if self.clip_max:
            data_array = np.clip(data_array, -self.clip_max, self.clip_max)
        return data_array

    def update(self, data_array: np.ndarray) -> None:
        """Chan's parallel merge of (mean, population var, count)."""
        batch_mean, batch_var = np.mean(data_array, axis=0), np.var(data_array, axis=0)
        batch_count = len(data_array)
        delta = batch_mean - self.mean
        total_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / total_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + delta**2 * self.count * batch_count / total_count
        self.mean, self.var = new_mean, m_2 / total_count
        self.count = total_count

    # -- device mirror (scalar statistics only) ---------------------------------------------
    def device_state(self, device: torch.device) -> torch.Tensor:
        return torch.tensor([float(self.mean), float(self.var), float(self.count)],
                            dtype=torch.float64, device=device)

    def load_device_state(self, state: torch.Tensor) -> None:
        m, v, c = state.detach().cpu().tolist()
        self.mean, self.var, self.count = np.float64(m), np.float64(v), int(round(c))
