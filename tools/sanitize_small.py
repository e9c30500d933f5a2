"""Small invocations of every entry point added for the SURVEY 8f rows (CPD, curvature, ICP) plus one tiny batch of the
hot path, for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pyfocusr_b200 as pyfocusr
from pyfocusr_b200 import SpectralBatch, _device
from pyfocusr_b200.cpd import affine_registration, deformable_registration

rng = np.random.RandomState(0)
t, s = pyfocusr.perturbed_ellipsoid(6, 0), pyfocusr.perturbed_ellipsoid(6, 1)
out = SpectralBatch(n_coords_spectral_ordering=200, graph_smoothing_iterations=5, projection_smooth_iterations=3).run_meshes([t], [s])
print("batch ok", int(out["final_idx"].sum()))
x, y = rng.rand(333, 3), rng.rand(301, 3)
a = affine_registration(X=x, Y=y, max_iterations=9, tolerance=0.0)
a.register()
for num_eig, m in ((100, 301), (40, 120), (100, 270)):
    d = deformable_registration(X=x, Y=y[:m], max_iterations=9, tolerance=0.0, alpha=0.5, beta=3.0, num_eig=num_eig)
    d.register()
    d.transform_point_cloud(rng.rand(777, 3))
d6 = deformable_registration(X=rng.rand(150, 6), Y=rng.rand(140, 6), max_iterations=4, tolerance=0.0, alpha=0.5, beta=3.0, num_eig=100)
d6.register()
print("cpd ok", a.iteration, d.iteration, d6.iteration)
c = _device.curvatures(t.points, t.tris)
print("curvature ok", float(c["mean"].sum()))
mat, moved = _device.icp(t.points, t.tris, s.points + 1.0, max_iterations=5, max_landmarks=77)
print("icp ok", float(mat[0, 3]))
torch.cuda.synchronize()
