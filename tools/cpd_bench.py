"""CPD stage timing on one GPU: the reference's default sizes (5000 x 5000 subsets, D = 3, num_eig = 100,
affine <= 100 iterations then deformable <= 1000 iterations, tolerance 1e-8) on synthetic coordinates, then
the whole drop-in Focusr.align_maps() on the shipped 15k pair with the CPD step on the GPU.
Usage: python tools/cpd_bench.py [n_points]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def problem(seed, n, m, d):
    rng = np.random.RandomState(seed)
    x = rng.rand(n, d) - 0.5
    a = np.eye(d) + 0.08 * rng.randn(d, d)
    full = x @ a + 0.04 + 0.03 * np.sin(3.0 * x[:, ::-1])
    return np.ascontiguousarray(x), np.ascontiguousarray(full[rng.permutation(n)][:m])


def main():
    import torch

    import pyfocusr_b200 as pyfocusr
    from pyfocusr_b200 import _lib
    from pyfocusr_b200.cpd import affine_registration, deformable_registration

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    x, y = problem(0, n, n, 3)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        aff = affine_registration(X=xd, Y=yd, max_iterations=100, tolerance=1e-8)
        aff.register()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ty = aff.transform_point_cloud(yd)
        l0 = _lib.launch_count()
        reg = deformable_registration(X=xd, Y=ty, max_iterations=1000, tolerance=1e-8, alpha=0.5, beta=3.0, num_eig=100)
        reg.register()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        # eigen-decomposition alone: max_iterations = 0
        reg0 = deformable_registration(X=xd, Y=ty, max_iterations=0, tolerance=1e-8, alpha=0.5, beta=3.0, num_eig=100)
        reg0.register()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print("rep %d  n=%d  affine: %d it, %.1f ms (%.3f ms/it)   deformable: %d it, %.1f ms total, low-rank setup %.1f ms "
              "(%d eig it, residual %.1e), EM %.3f ms/it, launches %d" %
              (rep, n, aff.iteration, (t1 - t0) * 1e3, (t1 - t0) * 1e3 / max(aff.iteration, 1), reg.iteration,
               (t2 - t1) * 1e3, (t3 - t2) * 1e3, reg0.eig_info["iterations"], reg0.eig_info["residual"],
               ((t2 - t1) - (t3 - t2)) * 1e3 / max(reg.iteration, 1), _lib.launch_count() - l0))
    # drop-in on the shipped pair
    g = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    mt = pyfocusr.PolyData(g["target_mesh_15k_points"], g["target_mesh_15k_tris"])
    ms = pyfocusr.PolyData(g["source_mesh_15k_points"], g["source_mesh_15k_tris"])
    for rep in range(2):
        np.random.seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f = pyfocusr.Focusr(mt, ms, icp_register_first=False, list_features_to_calc=[])
        t1 = time.perf_counter()
        f.align_maps()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("drop-in 15k pair with GPU CPD: ctor %.1f ms, align_maps %.1f ms, unique correspondences %d" %
              ((t1 - t0) * 1e3, (t2 - t1) * 1e3, len(np.unique(f.corresponding_target_idx_for_each_source_pt))))
    # every default of the reference, ICP included (100 iterations, 1000 landmarks, closest points on 30k triangles)
    from pyfocusr_b200 import _device

    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _device.icp(mt.points, mt.tris, ms.points)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        np.random.seed(0)
        f = pyfocusr.Focusr(mt, ms)
        t2 = time.perf_counter()
        f.align_maps()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print("ICP alone %.1f ms; Focusr(target, source) with all defaults: ctor %.1f ms + align_maps %.1f ms = %.1f ms" %
              ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t1) * 1e3))


if __name__ == "__main__":
    main()
