"""Robustness at size: the drop-in path with every default on a 100k-vertex (nu = 100) perturbed-ellipsoid pair."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pyfocusr_b200 as pyfocusr

nu = int(sys.argv[1]) if len(sys.argv) > 1 else 100
base = pyfocusr.icosphere(nu)
t, s = pyfocusr.perturbed_ellipsoid(nu, 0, base=base), pyfocusr.perturbed_ellipsoid(nu, 1, base=base)
for rep in range(2):
    np.random.seed(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f = pyfocusr.Focusr(t, s)
    t1 = time.perf_counter()
    f.align_maps()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
idx = f.corresponding_target_idx_for_each_source_pt
err = np.linalg.norm(f.weighted_avg_transformed_points - s.points, axis=1)
print("nu=%d: %d vertices per mesh; ctor %.1f ms, align_maps %.1f ms; unique correspondences %d; eigenvalues %s" %
      (nu, t.points.shape[0], (t1 - t0) * 1e3, (t2 - t1) * 1e3, len(np.unique(idx)), np.array2string(f.graph_target.eig_vals, precision=3)))
print("mean |matched target position - source position| = %.3f (mesh diameter ~86)" % float(np.mean(err)))
