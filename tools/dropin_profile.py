"""Where the time of the drop-in `Focusr(target, source)` (every default) goes on the shipped 15k pair: wall-clock
per stage with a device synchronise at every boundary (so this is latency, not throughput)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import pyfocusr_b200 as pyfocusr
    from pyfocusr_b200 import _device, focusr as fmod
    from pyfocusr_b200.graph import Graph

    g = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    mt = pyfocusr.PolyData(g["target_mesh_15k_points"], g["target_mesh_15k_tris"])
    ms = pyfocusr.PolyData(g["source_mesh_15k_points"], g["source_mesh_15k_tris"])
    stages = {}

    def wrap(obj, name, label):
        orig = getattr(obj, name)

        def timed(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = orig(*a, **k)
            torch.cuda.synchronize()
            stages[label] = stages.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
            return out

        setattr(obj, name, timed)
        return orig

    wrap(fmod, "_icp_transform", "icp")
    wrap(_device, "curvatures", "curvature (2 meshes)")
    wrap(Graph, "get_graph_spectrum", "laplacian + eigensolve (2 graphs)")
    wrap(pyfocusr.Focusr, "register_target_to_source", "cpd affine + deformable + transforms")
    wrap(pyfocusr.Focusr, "get_initial_correspondences", "knn initial")
    wrap(pyfocusr.Focusr, "get_smoothed_correspondences", "smoothing 300 + 40, knn final")
    wrap(pyfocusr.Focusr, "get_weighted_final_node_locations", "k=3 weighted positions")
    wrap(fmod.eigsort, "sort_eigenmaps", "eigsort")
    for rep in range(3):
        stages.clear()
        np.random.seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f = pyfocusr.Focusr(mt, ms)
        t1 = time.perf_counter()
        f.align_maps()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    tot = (t2 - t0) * 1e3
    print("Focusr(target, source) + align_maps(): %.1f ms  (ctor %.1f, align_maps %.1f)" % (tot, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    for k, v in sorted(stages.items(), key=lambda kv: -kv[1]):
        print("  %-45s %7.1f ms  %4.1f%%" % (k, v, 100 * v / tot))
    print("  %-45s %7.1f ms" % ("(host glue, numpy <-> device copies, prints)", tot - sum(stages.values())))


if __name__ == "__main__":
    main()
