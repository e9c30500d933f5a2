"""Times the 300 smoothing passes over the 128 target meshes of the bench workload (the library chosen by FOCUSR_B200_LIB,
see tools/smooth_ab.sh) and prints the SHA of the result: the A/B of kernel forms of Graph.mean_filter_graph."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import bench
    from pyfocusr_b200._device import DeviceGraph

    n_pairs = 128
    pts, tris, off, n, f, base = bench.make_pairs(list(range(n_pairs)))
    g = DeviceGraph.from_device(torch.from_numpy(pts).cuda(), torch.from_numpy(tris).cuda(), off)
    nt = int(off[n_pairs])
    nnz = 3.0 * f
    smooth_bytes = n_pairs * 300 * (12.0 * nnz + 60.0 * n)
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = g.mean_filter(g.points, 300, 0, nt)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    sha = hashlib.sha256(out[:nt].cpu().numpy().tobytes()).hexdigest()[:12]
    print("smoothing: %.2f ms for 300 passes over %d meshes  %.0f GB/s (%.1f%% of 6534.8)  sha %s" % (
        best, n_pairs, smooth_bytes / best / 1e6, smooth_bytes / best / 1e6 / 6534.8 * 100, sha), flush=True)


if __name__ == "__main__":
    main()
