"""Evidence for the dense / FP64-ALU side of the path, one JSON object:
  * the FP64 denominators measured on this box: hand DMMA / DFMA issue peaks (tools/micro/fp64_peak.cu) and cuBLAS DGEMM
    (torch.matmul on float64, 8192^3);
  * BASELINE.json configs[4] top end: k = 65 smallest eigenpairs of the 100 002-vertex perturbed sphere (block 96):
    seconds, eigenvalues against the scipy golden vector;
  * the pruned KNN at bench shape (128 segments of 15 212 x 15 212, d = 3, k = 1 and 3): queries/s, distance evaluations/s
    (live counter of the kernel) and their FP64 instruction rate as a fraction of the measured DFMA issue peak.
Usage: python tools/dense_evidence.py  (under ncu: -k regex:'k_gram|k_rotate|k_pk_search' for the per-kernel counters)"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from pyfocusr_b200 import _device, _lib
    from pyfocusr_b200._device import DeviceGraph
    from pyfocusr_b200.mesh import perturbed_ellipsoid

    out = {}
    knn_only = "--knn-only" in sys.argv
    exe = os.path.join(ROOT, "tools", "micro", "fp64_peak")
    if os.path.exists(exe) and "--no-micro" not in sys.argv:
        out["fp64_peaks"] = json.loads(subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1])
    if not knn_only:
        dgemm_and_sweep(out, torch, DeviceGraph, perturbed_ellipsoid)
    knn_evidence(out, torch, bench, _device, _lib)
    print(json.dumps(out))


def dgemm_and_sweep(out, torch, DeviceGraph, perturbed_ellipsoid):
    # cuBLAS DGEMM (library reference for the DMMA kernels)
    a = torch.randn((8192, 8192), dtype=torch.float64, device="cuda")
    b = torch.randn((8192, 8192), dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["cublas_dgemm_8192_tflops"] = 2 * 8192.0 ** 3 / best / 1e9
    del a, b
    # configs[4], k = 65
    m = perturbed_ellipsoid(100, seed=5, semi_axes=(1.0, 1.0, 1.0))
    g = DeviceGraph([m.points], [m.tris])
    gold = np.load(os.path.join(ROOT, "tests", "golden", "large_eigs.npz"))["nu100_seed5_k65"]
    res = {}
    for k in ((65,) if "--k65-only" in sys.argv else (17, 33, 65)):
        g.eigs_smallest(k=k, n_k_needed=k - 1)
        best = 1e9
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            vals, vecs, info = g.eigs_smallest(k=k, n_k_needed=k - 1)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        v = vals[0, : k - 1].cpu().numpy()
        res["k%d" % k] = {"seconds": best, "block": int(info["block_size"]), "outer_iterations": int(info["outer_iterations"][0]),
                          "filter_degree": int(info["filter_degree"][0]), "max_rel_err_vs_scipy": float(np.max(np.abs(v - gold[: k - 1]) / gold[: k - 1])),
                          "max_residual": float(info["max_residual"][0])}
    out["config5_100k_vertices"] = res
    del g


def knn_evidence(out, torch, bench, _device, _lib):
    # pruned KNN at bench shape
    lib = _lib.load()
    P = 128
    pts, tris, off, n, f, base = bench.make_pairs(list(range(P)))
    refs = torch.from_numpy(pts[: P * n]).cuda()
    qs = torch.from_numpy(pts[P * n:]).cuda()
    seg = torch.arange(P + 1, dtype=torch.int32, device="cuda") * n
    knn = {}
    for k in (1, 3):
        kw = dict(k=k, ref_off=seg, query_off=seg, max_queries=n, max_refs=n)
        _device.knn(refs, qs, **kw)
        lib.focusr_profile_reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _device.knn(refs, qs, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        pr = np.zeros(4)
        lib.focusr_profile_get_kind(4, pr.ctypes.data)
        evals = float(pr[0])
        rec = {"ms": ms, "queries_per_s": P * n / ms * 1e3, "distance_evaluations": evals,
               "evaluations_per_query": evals / (P * n), "of_brute_force": evals / (float(P) * n * n),
               "evaluations_per_s": evals / ms * 1e3,
               # 3 DSUB + 3 DMUL + 2 DADD per evaluation at d = 3 (no FMA: the sums must round as numpy's do)
               "fp64_instructions_per_s": 8.0 * evals / ms * 1e3}
        if "fp64_peaks" in out:
            rec["frac_of_dfma_issue_peak"] = rec["fp64_instructions_per_s"] / (out["fp64_peaks"]["dfma_ginstr_per_s"] * 1e9)
        knn["k%d_d3" % k] = rec
    out["knn_pruned_128x15212"] = knn


if __name__ == "__main__":
    main()
