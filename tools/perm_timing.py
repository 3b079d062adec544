"""Stage times of the host minibatch-order job (csrc/hostperm.cu) on this box: when each of the `repeat` rows of one update()
is ready, for every instruction-set path.  Host only (no GPU needed).  TS_B200_PERM_TRACE=1 adds the per-pass walk / apply times."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from tianshou_b200.data.batch import NumpyGlobalPermutationJob


def main() -> None:
    n, rep = 524288, 10
    rows = torch.empty((rep, n), dtype=torch.int32, pin_memory=torch.cuda.is_available())
    cpu = subprocess.run("lscpu | grep -E 'Model name|^CPU\\(s\\)|Thread|Socket'", shell=True, capture_output=True, text=True).stdout
    print(cpu.strip())
    for isa in ("scalar", "avx2", "avx512"):
        os.environ["TS_B200_PERM_ISA"] = isa
        best = None
        for trial in range(4):
            np.random.seed(trial)
            t0 = time.perf_counter()
            with NumpyGlobalPermutationJob(rows, rep) as job:
                ts = []
                for r in range(rep):
                    job.wait(r)
                    ts.append(1e3 * (time.perf_counter() - t0))
            if trial and (best is None or ts[-1] < best[-1]):
                best = ts
        np.random.seed(3)
        ok = np.array_equal(np.random.permutation(n), rows[0].numpy())
        print(f"isa<={isa:7s} rows ready at (ms): " + " ".join(f"{t:.2f}" for t in best) + f"   bit-exact row 0: {ok}")
    t0 = time.perf_counter()
    for _ in range(3):
        np.random.permutation(n).astype(np.int32)
    print(f"np.random.permutation + astype: {1e3 * (time.perf_counter() - t0) / 3:.2f} ms per row")


if __name__ == "__main__":
    main()
