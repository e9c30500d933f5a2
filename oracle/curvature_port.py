"""CPU oracle for the curvature node features  --  TEST INFRASTRUCTURE ONLY.  ** PARITY UNPINNED **

The reference's default ``list_features_to_calc=["curvature"]`` (focusr.py:59, graph.py:11-15,86-87) calls
``vtk.vtkCurvatures`` with ``SetCurvatureTypeToMinimum`` / ``SetCurvatureTypeToMaximum``
(vtk_functions.py:40-74).  VTK is an UNPINNED dependency (requirements.txt:2) that is not installed in this
image and not in the wheelhouse, so nothing below could be run against it.  This file restates the algorithm
of ``vtkCurvatures`` (VTK 9.x ``Filters/General/vtkCurvatures.cxx``; discrete curvatures of a triangle mesh):

  * Gauss curvature  K_v = 3 (2 pi - sum of the interior angles at v) / (sum of the areas of the triangles at v),
    0 where that area is 0; angles as pi - atan2(|a x b|, a . b) of consecutive edge vectors
    (``vtkMath::AngleBetweenVectors``), areas by Heron's formula on squared edge lengths
    (``vtkTriangle::TriangleArea``); contributions accumulate face by face in cell order.
  * Mean curvature   for every edge (v_l, v_r) of face f, in face order then edge order, that has exactly ONE
    other face n on it and n > f:  Hf = |e| * atan2((n_f x n_n) . e^, n_f . n_n) (0 if both arguments are 0),
    scaled by 3 / (area(f) + area(n)) when that sum is non-zero, added to both end points;
    H_v = 0.5 * (sum of Hf at v) / (number of such edges at v), 0 where there is none.  n_f is the unit normal
    of (v_l, v_r, v_o) in f's winding, n_n the unit normal of the neighbour in ITS OWN stored vertex order.
  * Minimum / maximum  k = H -/+ sqrt(H^2 - K) where H^2 - K >= 0, else 0.

Plain-loop reference (`curvatures_loops`, small meshes) and a vectorised version (`curvatures`) that is
asserted against it in the tests.  The CUDA implementation (csrc/curvature.cu) is compared with THIS file only.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np


def _tri_area(p1, p2, p3):
    a = np.sum((p1 - p2) ** 2, axis=-1)
    b = np.sum((p2 - p3) ** 2, axis=-1)
    c = np.sum((p3 - p1) ** 2, axis=-1)
    return 0.25 * np.sqrt(np.abs(4.0 * a * c - (a - b + c) * (a - b + c)))


def _unit_normal(p1, p2, p3):
    """vtkTriangle::ComputeNormal: (p3 - p2) x (p1 - p2), normalised when non-zero."""
    n = np.cross(p3 - p2, p1 - p2)
    ln = np.sqrt(np.sum(n * n, axis=-1, keepdims=True))
    return np.where(ln != 0.0, n / np.where(ln != 0.0, ln, 1.0), n)


def _angle_between(a, b):
    cr = np.cross(a, b)
    return np.arctan2(np.sqrt(np.sum(cr * cr, axis=-1)), np.sum(a * b, axis=-1))


def curvatures_loops(points, tris):
    """Literal transcription with Python loops (use on small meshes only)."""
    pts = np.asarray(points, dtype=np.float64)
    tris = np.asarray(tris, dtype=np.int64)
    n, f_count = pts.shape[0], tris.shape[0]
    K = np.full(n, 2.0 * math.pi)
    dA = np.zeros(n)
    for f in range(f_count):
        i0, i1, i2 = tris[f]
        v0, v1, v2 = pts[i0], pts[i1], pts[i2]
        e0, e1, e2 = v1 - v0, v2 - v1, v0 - v2
        alpha0 = math.pi - float(_angle_between(e1, e2))
        alpha1 = math.pi - float(_angle_between(e2, e0))
        alpha2 = math.pi - float(_angle_between(e0, e1))
        A = float(_tri_area(v0, v1, v2))
        dA[i0] += A
        dA[i1] += A
        dA[i2] += A
        K[i0] -= alpha1
        K[i1] -= alpha2
        K[i2] -= alpha0
    gauss = np.where(dA > 0.0, 3.0 * K / np.where(dA > 0.0, dA, 1.0), 0.0)
    faces_of = [[] for _ in range(n)]
    for f in range(f_count):
        for v in tris[f]:
            faces_of[v].append(f)
    mean_acc = np.zeros(n)
    num = np.zeros(n, dtype=np.int64)
    for f in range(f_count):
        for c in range(3):
            vl, vr, vo = tris[f][c], tris[f][(c + 1) % 3], tris[f][(c + 2) % 3]
            nb = [g for g in faces_of[vl] if g != f and vr in tris[g]]
            if len(nb) == 1 and nb[0] > f:
                g = nb[0]
                ore, end, oth = pts[vl], pts[vr], pts[vo]
                n_f = _unit_normal(ore, end, oth)
                e = end - ore
                length = math.sqrt(float(e @ e))
                if length != 0.0:
                    e = e / length
                Af = float(_tri_area(ore, end, oth))
                w0, w1, w2 = pts[tris[g][0]], pts[tris[g][1]], pts[tris[g][2]]
                Af += float(_tri_area(w0, w1, w2))
                n_n = _unit_normal(w0, w1, w2)
                cs = float(n_f @ n_n)
                sn = float(np.cross(n_f, n_n) @ e)
                Hf = length * math.atan2(sn, cs) if (sn != 0.0 or cs != 0.0) else 0.0
                if Af != 0.0:
                    Hf = Hf / Af * 3.0
                mean_acc[vl] += Hf
                mean_acc[vr] += Hf
                num[vl] += 1
                num[vr] += 1
    mean = np.where(num > 0, 0.5 * mean_acc / np.where(num > 0, num, 1), 0.0)
    return _finish(gauss, mean)


def _finish(gauss, mean):
    tmp = mean * mean - gauss
    root = np.sqrt(np.where(tmp >= 0.0, tmp, 0.0))
    kmax = np.where(tmp >= 0.0, mean + root, 0.0)
    kmin = np.where(tmp >= 0.0, mean - root, 0.0)
    return dict(gauss=gauss, mean=mean, minimum=kmin, maximum=kmax)


def curvatures(points, tris):
    """Vectorised restatement; accumulation in cell order (``np.add.at`` is sequential in index order)."""
    pts = np.asarray(points, dtype=np.float64)
    tris = np.asarray(tris, dtype=np.int64)
    n, F = pts.shape[0], tris.shape[0]
    v0, v1, v2 = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    e0, e1, e2 = v1 - v0, v2 - v1, v0 - v2
    alpha0 = math.pi - _angle_between(e1, e2)     # at v2
    alpha1 = math.pi - _angle_between(e2, e0)     # at v0
    alpha2 = math.pi - _angle_between(e0, e1)     # at v1
    area = _tri_area(v0, v1, v2)
    K = np.full(n, 2.0 * math.pi)
    dA = np.zeros(n)
    # per face: K[v0] -= alpha1; K[v1] -= alpha2; K[v2] -= alpha0, in face order
    np.add.at(K, tris.reshape(-1), -np.stack([alpha1, alpha2, alpha0], axis=1).reshape(-1))
    np.add.at(dA, tris.reshape(-1), np.repeat(area, 3))
    gauss = np.where(dA > 0.0, 3.0 * K / np.where(dA > 0.0, dA, 1.0), 0.0)
    # edges: half-edge h = 3 f + c runs v_l = tris[f, c] -> v_r = tris[f, c + 1]
    vl = tris.reshape(-1)
    vr = tris[:, [1, 2, 0]].reshape(-1)
    vo = tris[:, [2, 0, 1]].reshape(-1)
    face = np.repeat(np.arange(F), 3)
    key = np.minimum(vl, vr) * np.int64(n) + np.maximum(vl, vr)
    order = np.argsort(key, kind="stable")
    sk = key[order]
    start = np.flatnonzero(np.concatenate([[True], sk[1:] != sk[:-1]]))
    count = np.diff(np.concatenate([start, [sk.size]]))
    # exactly one other face on the edge <=> the undirected edge has exactly two half-edges (from different faces)
    two = np.repeat(count == 2, count)
    partner = np.empty(sk.size, dtype=np.int64)
    pos = np.arange(sk.size) - np.repeat(start, count)
    idx2 = np.flatnonzero(two)
    partner[:] = -1
    partner[idx2] = np.where(pos[idx2] == 0, idx2 + 1, idx2 - 1)
    nb_face = np.full(3 * F, -1, dtype=np.int64)
    nb_face[order[idx2]] = face[order[partner[idx2]]]
    sel = np.flatnonzero((nb_face > face))          # also drops -1; a face meeting itself on an edge cannot have n > f
    ore, end, oth = pts[vl[sel]], pts[vr[sel]], pts[vo[sel]]
    n_f = _unit_normal(ore, end, oth)
    e = end - ore
    length = np.sqrt(np.sum(e * e, axis=1))
    e = np.where(length[:, None] != 0.0, e / np.where(length != 0.0, length, 1.0)[:, None], e)
    g = nb_face[sel]
    w0, w1, w2 = pts[tris[g, 0]], pts[tris[g, 1]], pts[tris[g, 2]]
    Af = _tri_area(ore, end, oth) + _tri_area(w0, w1, w2)
    n_n = _unit_normal(w0, w1, w2)
    cs = np.sum(n_f * n_n, axis=1)
    sn = np.sum(np.cross(n_f, n_n) * e, axis=1)
    Hf = np.where((sn != 0.0) | (cs != 0.0), length * np.arctan2(sn, cs), 0.0)
    Hf = np.where(Af != 0.0, Hf / np.where(Af != 0.0, Af, 1.0) * 3.0, Hf)
    mean_acc = np.zeros(n)
    num = np.zeros(n, dtype=np.int64)
    # VTK's order: half-edges ascending (face, then edge), v_l before v_r
    np.add.at(mean_acc, np.stack([vl[sel], vr[sel]], axis=1).reshape(-1), np.repeat(Hf, 2))
    np.add.at(num, np.stack([vl[sel], vr[sel]], axis=1).reshape(-1), 1)
    mean = np.where(num > 0, 0.5 * mean_acc / np.where(num > 0, num, 1), 0.0)
    return _finish(gauss, mean)


def curvature_features(points, tris, which="curvature"):
    """The node-feature lists of graph.py:11-15: "curvature" -> [min, max]; "min_curvature"; "max_curvature"."""
    c = curvatures(points, tris)
    if which == "curvature":
        return [c["minimum"], c["maximum"]]
    if which == "min_curvature":
        return [c["minimum"]]
    if which == "max_curvature":
        return [c["maximum"]]
    raise KeyError(which)
