"""A/B of the kernel forms of the three HBM streams of the hot path on the bench workload (128 pairs = 256 meshes of
15 212 vertices): the fp32 filter steps of the eigensolver (sliced-ELL vs CSR, cache policies, occupancy), the smoothing
passes (sliced-ELL padded / packed vs CSR) and the Laplacian build.  Prints one line per variant with the live CUDA-event
time per launch (the library's own brackets), the GB/s by algorithmic bytes and whether the result is bit-identical to the
first variant.  Usage: python tools/kernel_ab.py [pairs]"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import bench
    from pyfocusr_b200 import _lib
    from pyfocusr_b200._device import DeviceGraph

    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    peak = 6534.8
    pts, tris, off, n, f, base = bench.make_pairs(list(range(n_pairs)))
    pts_d, tris_d = torch.from_numpy(pts).cuda(), torch.from_numpy(tris).cuda()
    lib = _lib.load()

    def sha(t):
        return hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:12]

    # --- Laplacian build
    for _ in range(3):
        g = DeviceGraph.from_device(pts_d, tris_d, off)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        g = DeviceGraph.from_device(pts_d, tris_d, off)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    nnz, meshes = 3.0 * f, 2.0 * n_pairs
    lap_bytes = meshes * (24.0 * n + 12.0 * f + 12.0 * nnz + 20.0 * n)
    print("laplacian build: %.3f ms wall per batch (%d meshes)  %.0f GB/s by compulsory bytes (%.1f%% of %.1f)" % (
        ms, 2 * n_pairs, lap_bytes / ms / 1e6, lap_bytes / ms / 1e6 / peak * 100, peak), flush=True)

    # --- filter steps
    variants = [("pol=3 pf=1 (default)", {}),
                ("pol=2 pf=1", dict(filter_policy=2)),
                ("pol=0 pf=1", dict(filter_policy=0)),
                ("pol=2 pf=0", dict(filter_prefetch=0)),
                ("pol=2 pf=1 6 CTAs/SM", dict(filter_min_blocks=6)),
                ("pol=3 pf=1 pdl=1", dict(filter_pdl=1)),
                ("pol=3 pf=1 pdl=0", dict(filter_pdl=0))]
    ref = None
    for name, opt in variants:
        for _ in range(2):
            vals, vecs, info = g.eigs_smallest(k=7, n_k_needed=6, options=opt)
        lib.focusr_profile_reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            vals, vecs, info = g.eigs_smallest(k=7, n_k_needed=6, options=opt)
        e1.record()
        torch.cuda.synchronize()
        line = "%-28s solve %.1f ms" % (name, e0.elapsed_time(e1) / 2)
        for kind, tag in ((1, "fp32"), (2, "corr")):
            pr = np.zeros(4)
            lib.focusr_profile_get_kind(kind, pr.ctypes.data)
            if pr[1] > 0:
                gbs = pr[2] / (pr[0] / 1e3) / 1e9
                line += " | %s %.4f ms/launch %.0f GB/s (%.1f%%) x%d" % (tag, pr[0] / pr[1], gbs, gbs / peak * 100, int(pr[1]))
        h = sha(vecs)
        ref = ref or h
        line += " | degree %d residual %.1e identical=%s" % (int(info["filter_degree"].max()), info["max_residual"].max(), h == ref)
        print(line, flush=True)

    # --- smoothing: 300 passes over the targets
    nt = int(off[n_pairs])
    smooth_bytes = n_pairs * 300 * (12.0 * nnz + 60.0 * n)
    for _rep in range(1):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = g.mean_filter(g.points, 300, 0, nt)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        gbs = smooth_bytes / best / 1e6
        print("smoothing: %.2f ms for 300 passes over %d meshes  %.0f GB/s (%.1f%%)  sha %s" % (
            best, n_pairs, gbs, gbs / peak * 100, sha(out[:nt])), flush=True)


if __name__ == "__main__":
    main()
