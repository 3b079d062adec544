"""Diagnostic: TS-mode tcgen05.mma (A operand in tensor memory, written by tcgen05.st) vs an f64 matmul."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tianshou_b200 import _cabi
from tianshou_b200._cabi import call, ptr, stream_ptr
_cabi.use_diagnostics_library()      # ts_umma_selftest lives in the diagnostics build (python -m tianshou_b200.csrc.build --diag)
dev = "cuda:0"
rng = np.random.default_rng(0)
for b_mn in (0, 1):
    for N, K in ((64, 64), (64, 32), (16, 64), (64, 16)):
        a = rng.standard_normal((128, K)).astype(np.float32); b = rng.standard_normal((N, K)).astype(np.float32)
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        d = torch.full((128, N), float("nan"), dtype=torch.float32, device=dev)
        ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        try:
            call("ts_umma_selftest", ptr(ta), ptr(tb), ptr(d), 128, N, K, 1, 2, b_mn, 0, stream_ptr())
            torch.cuda.synchronize()
            err = float(np.abs(d.cpu().numpy() - ref).max() / np.abs(ref).max())
        except Exception as e:  # noqa: BLE001
            err = f"ERR {e}"
        print(f"TS bf16x3 M=128 b_mn={b_mn} N={N} K={K}: rel err {err}", flush=True)
