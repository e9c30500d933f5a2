"""Smoothing stage alone on the bench workload: one launch per pass vs the persistent cluster kernel (variants).
Usage: python tools/smooth_bench.py [pairs]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import bench
    from pyfocusr_b200 import _lib
    from pyfocusr_b200._device import DeviceGraph

    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    pts, _, _, n, _, base = bench.make_pairs(list(range(n_pairs)))
    pts = pts.reshape(2 * n_pairs, n, 3)
    g = DeviceGraph(list(pts), [base.tris] * (2 * n_pairs))

    def timed(**kw):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = g.mean_filter(g.points, 300, 0, g.n_points, **kw)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, out

    t_ref, ref = timed(cluster=False)
    print("per-pass launches: %.2f ms for 300 passes over %d meshes" % (t_ref, 2 * n_pairs))
    for var, name in ((0, "1024 thr x 4"), (1, "512 x 8"), (2, "512 x 4"), (3, "1024 x 8")):
        _lib.call("focusr_set_tuning", 2, var)
        t, out = timed(cluster=True)
        print("cluster variant %d (%s): %.2f ms, bit-identical %s" % (var, name, t, bool(torch.equal(out, ref))))
    _lib.call("focusr_set_tuning", 2, 0)


if __name__ == "__main__":
    main()
