"""The `hungarian` assignment at the size of the reference's 15k meshes: N points of an ellipsoid matched to displaced,
permuted copies (long augmenting paths).  On the GPU: `python tools/lsap_large.py gpu [N]` (writes the assignment to
gpurun_out/); on a host: `python tools/lsap_large.py scipy [N]` (scipy.optimize.linear_sum_assignment on one core, compared
with the GPU's assignment if its file is present)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def problem(n):
    rng = np.random.RandomState(11)
    p = rng.randn(n, 3)
    p = p / np.linalg.norm(p, axis=1)[:, None] * np.array([1.0, 0.8, 0.6])
    q = p[rng.permutation(n)] + 0.3 * rng.randn(n, 3)
    return q, p


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "gpu"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 15000
    q, p = problem(n)
    out = os.path.join(ROOT, "gpurun_out", "lsap_%d_col.npy" % n)
    if mode == "gpu":
        import torch

        from pyfocusr_b200 import _device

        cost = _device.cdist(q, p)
        _device.linear_sum_assignment(cost[:512, :512].contiguous())   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, col = _device.linear_sum_assignment(cost)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        os.makedirs(os.path.dirname(out), exist_ok=True)
        np.save(out, col)
        total = float(cost[torch.arange(n, device="cuda"), torch.from_numpy(col).cuda()].sum())
        print("focusr_lsap %d x %d: %.2f s, total cost %.9f" % (n, n, dt, total), flush=True)
    else:
        from scipy.optimize import linear_sum_assignment
        from scipy.spatial.distance import cdist

        c = cdist(q, p)
        t0 = time.perf_counter()
        _, col = linear_sum_assignment(c)
        dt = time.perf_counter() - t0
        msg = "scipy linear_sum_assignment %d x %d on one core: %.1f s, total cost %.9f" % (n, n, dt, float(c[np.arange(n), col].sum()))
        if os.path.exists(out):
            msg += "; equal to the GPU's assignment: %s" % bool(np.array_equal(col, np.load(out)))
        print(msg, flush=True)


if __name__ == "__main__":
    main()
