"""CPU oracle for the ICP pre-alignment  --  TEST INFRASTRUCTURE ONLY.  ** PARITY UNPINNED **

The reference's default ``icp_register_first=True`` (focusr.py:106-131) calls ``vtk_functions.icp_transform``
(vtk_functions.py:12-29: ``vtkIterativeClosestPointTransform`` with a rigid-body or similarity landmark
transform, 100 iterations, ``StartByMatchingCentroidsOn``) and ``apply_transform`` (vtk_functions.py:32-37).
VTK is an UNPINNED dependency (requirements.txt:2), not installed here and not in the wheelhouse, so nothing
below could be run against it.  This file restates VTK 9's algorithm:

  * ``vtkIterativeClosestPointTransform::InternalUpdate``: landmarks = every ``step``-th source point,
    ``step = n_source // max_landmarks`` when ``n_source > max_landmarks``; start by translating the source
    centroid (all points) onto the target centroid (all points); every iteration finds, for each landmark, the
    closest point ON THE TARGET SURFACE (``vtkCellLocator::FindClosestPoint``), fits the landmark transform,
    left-multiplies it onto the accumulated matrix and moves the landmarks; ``CheckMeanDistance`` is off, so
    exactly ``max_iterations`` fits are made.
  * ``vtkLandmarkTransform`` (Horn 1987, unit quaternions): centroids, M = sum a b^T, the 4x4 matrix N, its
    dominant eigenvector as quaternion, optional scale sqrt(sum |b|^2 / sum |a|^2), translation from the
    centroids.  The collinear / two-point special case of VTK is not restated (cannot occur with >= 3
    non-collinear landmarks).
  * ``max_landmarks``: the reference calls ``SetMaximumNumberOfLandmarks(1000)`` AFTER ``icp.Update()``
    (vtk_functions.py:26-28); that call bumps the transform's modification time, so the
    ``vtkTransformPolyDataFilter`` of ``apply_transform`` re-runs the registration with 1000 landmarks before
    transforming the points.  The effective value is therefore 1000 (the first run with VTK's default 200 is
    discarded).

The closest point on a triangle is unique, so any exact method agrees with VTK's up to rounding; this file and
the CUDA kernel both use Ericson's region test ("Real-Time Collision Detection", 5.1.5); among triangles at
equal distance the lowest index wins (VTK's locator order is unspecified; the closest POINT is the same unless
two different surface points are exactly equidistant).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np


def closest_points_on_triangles(p, a, b, c):
    """Closest point of triangle (a, b, c) to p, broadcasting over leading dimensions; returns (point, dist2)."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = np.sum(ab * ap, -1), np.sum(ac * ap, -1)
    bp = p - b
    d3, d4 = np.sum(ab * bp, -1), np.sum(ac * bp, -1)
    cp = p - c
    d5, d6 = np.sum(ab * cp, -1), np.sum(ac * cp, -1)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    with np.errstate(divide="ignore", invalid="ignore"):
        v_ab = d1 / (d1 - d3)
        w_ac = d2 / (d2 - d6)
        w_bc = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        denom = 1.0 / (va + vb + vc)
    conds = [
        (d1 <= 0) & (d2 <= 0),
        (d3 >= 0) & (d4 <= d3),
        (vc <= 0) & (d1 >= 0) & (d3 <= 0),
        (d6 >= 0) & (d5 <= d6),
        (vb <= 0) & (d2 >= 0) & (d6 <= 0),
        (va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0),
    ]
    choices = [
        a + 0.0 * p,
        b + 0.0 * p,
        a + v_ab[..., None] * ab,
        c + 0.0 * p,
        a + w_ac[..., None] * ac,
        b + w_bc[..., None] * (c - b),
    ]
    inside = a + ab * (vb * denom)[..., None] + ac * (vc * denom)[..., None]
    out = inside
    for cond, ch in zip(conds[::-1], choices[::-1]):       # first matching region wins
        out = np.where(cond[..., None], ch, out)
    d = p - out
    return out, np.sum(d * d, -1)


def closest_points_on_mesh(points, target_points, target_tris, chunk=64):
    a, b, c = (target_points[target_tris[:, k]] for k in range(3))
    out = np.empty_like(points)
    for s in range(0, points.shape[0], chunk):
        p = points[s:s + chunk, None, :]
        cp, d2 = closest_points_on_triangles(p, a[None], b[None], c[None])
        best = np.argmin(d2, axis=1)                        # first minimum = lowest triangle index
        out[s:s + chunk] = cp[np.arange(cp.shape[0]), best]
    return out


def landmark_transform(src, dst, similarity=False):
    """vtkLandmarkTransform (rigid body / similarity) as a 4x4 matrix acting on column vectors."""
    n = src.shape[0]
    cs, ct = src.sum(0) / n, dst.sum(0) / n
    mat = np.eye(4)
    if n == 1:
        mat[:3, 3] = ct - cs
        return mat
    a, b = src - cs, dst - ct
    M = a.T @ b
    sa, sb = float(np.sum(a * a)), float(np.sum(b * b))
    if sa == 0.0 or sb == 0.0:
        mat[:3, 3] = ct - cs
        return mat
    N = np.empty((4, 4))
    N[0, 0] = M[0, 0] + M[1, 1] + M[2, 2]
    N[1, 1] = M[0, 0] - M[1, 1] - M[2, 2]
    N[2, 2] = -M[0, 0] + M[1, 1] - M[2, 2]
    N[3, 3] = -M[0, 0] - M[1, 1] + M[2, 2]
    N[0, 1] = N[1, 0] = M[1, 2] - M[2, 1]
    N[0, 2] = N[2, 0] = M[2, 0] - M[0, 2]
    N[0, 3] = N[3, 0] = M[0, 1] - M[1, 0]
    N[1, 2] = N[2, 1] = M[0, 1] + M[1, 0]
    N[1, 3] = N[3, 1] = M[2, 0] + M[0, 2]
    N[2, 3] = N[3, 2] = M[1, 2] + M[2, 1]
    ev, evec = np.linalg.eigh(N)
    w, x, y, z = evec[:, np.argmax(ev)]
    R = np.array([
        [w * w + x * x - y * y - z * z, 2.0 * (-w * z + x * y), 2.0 * (w * y + x * z)],
        [2.0 * (w * z + x * y), w * w - x * x + y * y - z * z, 2.0 * (-w * x + y * z)],
        [2.0 * (-w * y + x * z), 2.0 * (w * x + y * z), w * w - x * x - y * y + z * z],
    ])
    if similarity:
        R = R * np.sqrt(sb / sa)
    mat[:3, :3] = R
    mat[:3, 3] = ct - R @ cs
    return mat


def icp_transform(target_points, target_tris, source_points, max_iterations=100, max_landmarks=1000, similarity=False,
                  start_by_matching_centroids=True):
    """vtkIterativeClosestPointTransform as configured at vtk_functions.py:12-29.  Returns the 4x4 matrix."""
    tp = np.asarray(target_points, dtype=np.float64)
    sp = np.asarray(source_points, dtype=np.float64)
    tris = np.asarray(target_tris, dtype=np.int64)
    n = sp.shape[0]
    step = n // max_landmarks if n > max_landmarks else 1
    nb = n // step
    acc = np.eye(4)
    if start_by_matching_centroids:
        acc[:3, 3] = tp.sum(0) / tp.shape[0] - sp.sum(0) / n
    a = sp[np.arange(nb) * step] @ acc[:3, :3].T + acc[:3, 3]
    for it in range(max_iterations):
        closest = closest_points_on_mesh(a, tp, tris)
        L = landmark_transform(a, closest, similarity)
        acc = L @ acc
        if it + 1 >= max_iterations:
            break
        a = a @ L[:3, :3].T + L[:3, 3]
    return acc


def apply_transform(points, matrix):
    """vtk_functions.py:32-37 on a point array."""
    return np.asarray(points, dtype=np.float64) @ matrix[:3, :3].T + matrix[:3, 3]
