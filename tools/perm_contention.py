"""How the host minibatch-order job behaves when 8 ranks of one box run theirs at the same time (weak scaling, N = 8): P
processes start a job for one C2 update (10 rows x 524 288) at the same instant; prints when the rows were ready (median
over processes, worst process) for several applier counts, with and without NUMA confinement.  Host only (pinned rows if CUDA
is there)."""
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def worker(rank, P, nw, pin, start_at, q):
    os.environ["LOCAL_WORLD_SIZE"] = str(P)
    os.environ["TS_B200_PERM_PIN"] = "1" if pin else "0"
    import numpy as np
    import torch

    from tianshou_b200.data.batch import NumpyGlobalPermutationJob
    n, rep = 524288, 10
    rows = torch.empty((rep, n), dtype=torch.int32, pin_memory=torch.cuda.is_available())
    res = []
    for trial in range(4):
        np.random.seed(100 * rank + trial)
        while time.time() < start_at + 0.5 * trial:
            pass
        t0 = time.perf_counter()
        with NumpyGlobalPermutationJob(rows, rep, n_workers=nw) as job:
            ts = []
            for r in range(rep):
                job.wait(r)
                ts.append(1e3 * (time.perf_counter() - t0))
        if trial:
            res.append(ts)
    q.put((rank, res))


def main():
    import numpy as np
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ctx = mp.get_context("spawn")
    for nw in ((4,) if len(sys.argv) > 2 else (4, 2, 1)):
        for pin in (True, False):
            q = ctx.Queue()
            start_at = time.time() + 25.0          # after every process has imported torch
            ps = [ctx.Process(target=worker, args=(r, P, nw, pin, start_at, q)) for r in range(P)]
            for p in ps:
                p.start()
            out = [q.get() for _ in ps]
            for p in ps:
                p.join()
            a = np.array([t for _, res in out for t in res])        # [P * trials, rep]
            med, worst = np.median(a, axis=0), a.max(axis=0)
            print(f"P={P} appliers={nw} numa_pin={int(pin)}  row ready, median over ranks (ms): " + " ".join(f"{x:.1f}" for x in med))
            print(f"{'':34s}worst rank (ms): " + " ".join(f"{x:.1f}" for x in worst), flush=True)


if __name__ == "__main__":
    main()
