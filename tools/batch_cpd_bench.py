"""Batched pipeline WITH the CPD step: pairs registered one after the other vs concurrently on several streams.
Usage: python tools/batch_cpd_bench.py [pairs]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from pyfocusr_b200 import SpectralBatch, ellipsoid_pair, icosphere

    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    base = icosphere(39)
    pairs = [ellipsoid_pair(i, 39, base=base) for i in range(n_pairs)]
    t, s = [p[0] for p in pairs], [p[1] for p in pairs]
    ref = None
    for streams in (1, 8, 16, 32, 8):
        sb = SpectralBatch(registration="b200", cpd_streams=streams)
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = sb.run_meshes(t, s, record_events=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        fin = out["final_idx"].cpu().numpy()
        same = True if ref is None else bool(np.array_equal(fin, ref))
        ref = fin if ref is None else ref
        print("cpd_streams=%2d: %d pairs in %.3f s = %.1f pairs/s (cpd stage %.1f ms); same result as 1 stream: %s" %
              (streams, n_pairs, dt, n_pairs / dt, sb.timings.get("cpd", float("nan")), same))


if __name__ == "__main__":
    main()
