"""Small eigensolves that reach every form of the filter step (fp32 blocks with fp64 in / out steps, the fp32 correction
form and its last step, fp64) and the fp32 matrix copy at the workspace tail, in a batch whose symmetric run does not
start at row 0.  Written for `compute-sanitizer --tool memcheck python tools/sanitize_eigs_mixed.py`; the sanitizer is closed on
this GPU pool (gpurun_out/san_memcheck.log), so it serves as a plain smoke run of these paths (gpurun_out/san_plain.log)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pyfocusr_b200.mesh as fmesh
from pyfocusr_b200 import SpectralBatch, _lib
from pyfocusr_b200._device import DeviceGraph

rng = np.random.RandomState(1)
base = fmesh.perturbed_ellipsoid(8, 5)
t = base.tris.copy()
fl = rng.choice(len(t), 6, replace=False)
t[fl] = t[fl][:, [0, 2, 1]]                       # flipped triangles: one-way entries, non-symmetric run first
ms = [fmesh.PolyData(base.points, t), fmesh.perturbed_ellipsoid(7, 1), fmesh.perturbed_ellipsoid(9, 2), fmesh.icosphere(6)]
g = DeviceGraph([m.points for m in ms], [m.tris for m in ms])
for block in (0, 24, 32):
    vals, vecs, info = g.eigs_smallest(k=5, n_k_needed=4, block_size=block)
    for opt in (dict(filter_policy=3), dict(filter_policy=0, filter_prefetch=0), dict(filter_min_blocks=6)):
        g.eigs_smallest(k=5, n_k_needed=4, block_size=block, options=opt)
    print("block", block, "status", info["status"].tolist(), "fp32 steps", info["fp32_filter_degree"].tolist(),
          "degree", info["filter_degree"].tolist(), "residual %.1e" % info["max_residual"].max())
vals, vecs, info = g.eigs_smallest(k=5, n_k_needed=4, options=dict(mixed_precision=0))
print("fp64 only: status", info["status"].tolist(), "fp32 steps", info["fp32_filter_degree"].tolist())
sb = SpectralBatch(n_coords_spectral_ordering=200, graph_smoothing_iterations=5, projection_smooth_iterations=3)
jobs = [sb.pack_meshes([ms[1]], [ms[2]]), sb.pack_meshes([ms[2], ms[3]], [ms[1], ms[1]])]
outs = sb.run_concurrent(jobs)
torch.cuda.synchronize()
print("concurrent ok", [int(o["final_idx"].sum()) for o in outs])
