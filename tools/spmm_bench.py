"""Micro-benchmark of the CSR x block kernel (MODE 2, y = L x) on a batch of 15 212-vertex meshes.  Algorithmic bytes per launch:
12 nnz + 4 N + 16 N + 16 b N.  Usage: python tools/spmm_bench.py [n_meshes] [b]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyfocusr_b200 import _lib
from pyfocusr_b200._device import DeviceGraph
from pyfocusr_b200.mesh import icosphere, perturbed_ellipsoid

M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
b = int(sys.argv[2]) if len(sys.argv) > 2 else 16
base = icosphere(39)
ms = [perturbed_ellipsoid(39, s, base=base) for s in range(8)]
g = DeviceGraph([ms[i % 8].points for i in range(M)], [ms[i % 8].tris for i in range(M)])
x = torch.randn((g.n_points, b), dtype=torch.float64, device="cuda")
lib = _lib.load()
nbytes = 12.0 * g.nnz + 20.0 * g.n_points + 16.0 * b * g.n_points
ref = None
for var in (0, 1):
    lib.focusr_set_tuning(0, var)
    y = g.laplacian_apply(x)
    if ref is None:
        ref = y.clone()
    ok = torch.equal(y, ref)
    for _ in range(3):
        g.laplacian_apply(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        g.laplacian_apply(x)
    e1.record(); torch.cuda.synchronize()
    ms_ = e0.elapsed_time(e1) / 20
    print("variant %d: %.4f ms/launch  %.0f GB/s (%.1f%% of 6534.8)  identical=%s" % (var, ms_, nbytes / ms_ / 1e6, nbytes / ms_ / 1e6 / 65.348, ok), flush=True)
