"""Per-mesh convergence of the batched eigensolve on the bench workload: how evenly the meshes of one batch
finish (a batch runs every filter pass over all of its meshes).  Usage: python tools/batch_convergence.py [pairs]"""
import collections
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import bench
    from pyfocusr_b200._device import DeviceGraph

    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    pts, _, _, n, _, base = bench.make_pairs(list(range(n_pairs)))
    pts = pts.reshape(2 * n_pairs, n, 3)
    g = DeviceGraph(list(pts), [base.tris] * (2 * n_pairs))
    for _ in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vals, vecs, info = g.eigs_smallest(k=7, n_k_needed=6)
        e1.record()
        torch.cuda.synchronize()
    print("solve ms", e0.elapsed_time(e1), "block", info["block_size"])
    print("outer iterations:", sorted(collections.Counter(info["outer_iterations"].tolist()).items()))
    print("filter degree   :", sorted(collections.Counter(info["filter_degree"].tolist()).items()))
    print("max residual    : max %.2e median %.2e" % (np.max(info["max_residual"]), np.median(info["max_residual"])))
    print("spectrum bound  : min %.4f max %.4f" % (np.min(info["spectrum_bound"]), np.max(info["spectrum_bound"])))


if __name__ == "__main__":
    main()
