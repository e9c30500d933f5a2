"""Golden eigenvalues for the large synthetic configs (BASELINE.json configs[3], configs[4]) from the
reference's own call -- scipy eigs(L, k, sigma=1e-10, which="LM", ncv=4k) as in graph.py:372 -- on the
Laplacian assembled by oracle/port.py (pinned bitwise against the unmodified reference).  The 1M-vertex
solve needs ~5 GB and minutes of SuperLU time, so it is run once here and the values are committed.

Usage: python oracle/make_golden_large.py [--skip-1m]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402
from pyfocusr_b200.mesh import icosphere, perturbed_ellipsoid  # noqa: E402

out = {}
path = os.path.join(ROOT, "tests", "golden", "large_eigs.npz")
if os.path.exists(path):
    out.update(np.load(path))

# config 5: 100 002-vertex perturbed sphere (multiplets split), k = 65 smallest
t = time.time()
m = perturbed_ellipsoid(100, seed=5, semi_axes=(1.0, 1.0, 1.0))
lap = port.laplacian(port.adjacency(m.points, m.tris))
vals, _ = port.recursive_eig(lap, 65, 64, 1)
out["nu100_seed5_k65"] = np.sort(vals)
print("nu=100 k=65: %d values in %.1f s" % (vals.size, time.time() - t), flush=True)
np.savez_compressed(path, **out)

if "--skip-1m" not in sys.argv:
    t = time.time()
    m = icosphere(316)
    lap = port.laplacian(port.adjacency(m.points, m.tris))
    vals, _ = port.recursive_eig(lap, 11, 10, 1)
    out["nu316_k11"] = np.sort(vals)
    print("nu=316 (998 562 vertices) k=11: %d values in %.1f s" % (vals.size, time.time() - t), vals, flush=True)
    np.savez_compressed(path, **out)
