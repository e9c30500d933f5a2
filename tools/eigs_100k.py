"""BASELINE.json configs[4]: eigenpair sweep on the 100 002-vertex perturbed sphere (stresses the
Gram / Rayleigh-Ritz DMMA kernels).  Usage: python tools/eigs_100k.py [k ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyfocusr_b200._device import DeviceGraph
from pyfocusr_b200.mesh import perturbed_ellipsoid

ks = [int(a) for a in sys.argv[1:]] or [65]
m = perturbed_ellipsoid(100, seed=5, semi_axes=(1.0, 1.0, 1.0))
g = DeviceGraph([m.points], [m.tris])
gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "large_eigs.npz"))["nu100_seed5_k65"]
for k in ks:
    g.eigs_smallest(k=k, n_k_needed=k - 1)  # warm-up
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vals, vecs, info = g.eigs_smallest(k=k, n_k_needed=k - 1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    v = vals[0, : k - 1].cpu().numpy()
    print("k=%d block=%d outer=%d degree=%d time=%.3f s max rel err vs scipy %.2e" % (
        k, info["block_size"], info["outer_iterations"][0], info["filter_degree"][0], dt, np.max(np.abs(v - gold[: k - 1]) / gold[: k - 1])), flush=True)
