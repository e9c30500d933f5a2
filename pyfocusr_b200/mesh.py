"""Mesh containers, IO and synthetic generators for the spectral-correspondence path.

The reference consumes ``vtkPolyData`` objects and touches only four accessors on the hot
path (``GetNumberOfPoints/GetPoint/GetNumberOfCells/GetCell``; reference
``pyfocusr/graph.py:58-62,155-164``) plus ``GetPointData()`` for optional scalars
(``graph.py:88-104``).  VTK is not installable in this image, so :class:`PolyData` is a
duck-typed stand-in exposing the same accessors (so the *unmodified* reference can consume it
in the oracle harness) together with flat ``points``/``tris`` arrays that the CUDA path
uploads directly.

Generators implement the synthetic inputs of BASELINE.json ``configs[2..4]`` (SURVEY.md
§8d): class-I geodesic icospheres (``10*nu**2 + 2`` vertices) and perturbed ellipsoids.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "PolyData",
    "read_vtk_mesh",
    "mesh_arrays",
    "icosphere",
    "perturbed_ellipsoid",
    "ellipsoid_pair",
]


class _Edge:
    __slots__ = ("_a", "_b")

    def __init__(self, a, b):
        self._a = a
        self._b = b

    def GetPointId(self, i):
        return self._a if i == 0 else self._b


class _Cell:
    __slots__ = ("_ids",)

    def __init__(self, ids):
        self._ids = ids

    def GetNumberOfEdges(self):
        return len(self._ids)

    def GetNumberOfPoints(self):
        return len(self._ids)

    def GetPointId(self, i):
        return int(self._ids[i])

    def GetEdge(self, j):
        # VTK polygon edge order: (0,1), (1,2), ..., (nv-1,0)
        n = len(self._ids)
        return _Edge(int(self._ids[j]), int(self._ids[(j + 1) % n]))


class _Array:
    def __init__(self, name, values):
        self._name = name
        self.values = np.ascontiguousarray(values)

    def GetName(self):
        return self._name

    # numpy_support.vtk_to_numpy stub in the oracle harness is np.asarray
    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)


class _PointData:
    def __init__(self, arrays, owner=None):
        self._arrays = arrays
        self._owner = owner

    def SetScalars(self, values):
        """vtkPointData::SetScalars for the PolyData stand-in: the active scalars become ``values`` (a numpy array, or
        whatever ``numpy_to_vtk`` returned when VTK is installed)."""
        if self._owner is None:
            raise RuntimeError("this point data is not attached to a PolyData")
        try:
            from vtk.util.numpy_support import vtk_to_numpy  # type: ignore

            values = vtk_to_numpy(values) if not isinstance(values, np.ndarray) else values
        except ImportError:
            pass
        v = np.asarray(values)
        if v.shape[0] != self._owner.points.shape[0]:
            raise ValueError("scalars must have one value per point")
        old = dict(self._owner.point_scalars)
        self._owner.point_scalars = {"scalars": v}
        self._owner.point_scalars.update({k: a for k, a in old.items() if k != "scalars"})

    def GetNumberOfArrays(self):
        return len(self._arrays)

    def GetArray(self, idx):
        return self._arrays[idx]

    def GetScalars(self):
        return self._arrays[0] if self._arrays else None


class PolyData:
    """Triangle surface mesh: ``points (N,3) float64`` and ``tris (F,3) int32``.

    Exposes the vtkPolyData accessors used by the reference (``graph.py:58-62,155-164``).
    """

    def __init__(self, points, tris, point_scalars=None):
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self.tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        self.point_scalars = dict(point_scalars or {})

    # --- vtkPolyData duck-typing -------------------------------------------------
    def GetNumberOfPoints(self):
        return self.points.shape[0]

    def GetPoint(self, i):
        p = self.points[i]
        return (float(p[0]), float(p[1]), float(p[2]))

    def GetNumberOfCells(self):
        return self.tris.shape[0]

    def GetCell(self, i):
        return _Cell(self.tris[i])

    def GetPointData(self):
        return _PointData([_Array(k, v) for k, v in self.point_scalars.items()], owner=self)

    def copy(self):
        return PolyData(self.points.copy(), self.tris.copy(), dict(self.point_scalars))


_VTK_TYPES = {
    "float": ">f4", "double": ">f8", "int": ">i4", "unsigned_int": ">u4", "long": ">i8", "unsigned_long": ">u8",
    "short": ">i2", "unsigned_short": ">u2", "char": ">i1", "unsigned_char": ">u1", "vtktypeint64": ">i8",
    "vtktypeint32": ">i4", "vtkidtype": ">i8", "bit": ">u1",
}


class _LegacyCursor:
    """Walks a legacy .vtk file held in memory: header lines as text, data blocks as ASCII tokens or
    big-endian binary (the legacy format's byte order)."""

    def __init__(self, buf, binary):
        self.buf, self.pos, self.binary = buf, 0, binary

    def line(self):
        """Next non-blank line as a list of words (None at the end of the file)."""
        n = len(self.buf)
        while self.pos < n:
            end = self.buf.find(b"\n", self.pos)
            end = n if end < 0 else end
            raw = self.buf[self.pos:end]
            self.pos = end + 1
            words = raw.decode("ascii", "replace").split()
            if words:
                return words
        return None

    def values(self, count, vtk_type):
        dt = _VTK_TYPES.get(vtk_type.lower())
        if dt is None:
            raise ValueError("unsupported VTK data type %r" % vtk_type)
        if self.binary:
            nbytes = count * np.dtype(dt).itemsize
            if self.pos + nbytes > len(self.buf):
                raise ValueError("truncated binary section")
            out = np.frombuffer(self.buf, dtype=dt, count=count, offset=self.pos)
            self.pos += nbytes
            return out
        parts = self.buf[self.pos:].split(None, count)
        if len(parts) < count:
            raise ValueError("truncated ASCII section")
        rest = parts[count] if len(parts) > count else b""
        self.pos = len(self.buf) - len(rest)
        if np.dtype(dt).kind != "f":
            return np.array(parts[:count], dtype=np.int64)
        out = np.array(parts[:count], dtype=np.float64)
        return out.astype(np.float32) if np.dtype(dt).itemsize == 4 else out   # a `float` array holds float32 values


def read_vtk_mesh(path_to_file):
    """Legacy ``DATASET POLYDATA`` reader, ASCII and BINARY (replaces ``vtk_functions.py:5-9``, which goes
    through ``vtkPolyDataReader``).

    Handles ``POINTS``, ``POLYGONS`` in the classic layout (``n size`` then ``3 i j k`` records) and in the
    version 5.x layout (``OFFSETS`` / ``CONNECTIVITY`` arrays), and ``POINT_DATA`` with ``SCALARS`` and
    ``FIELD`` arrays; other sections (``VERTICES``, ``LINES``, ``CELL_DATA``, ``NORMALS``, ...) are skipped.
    All polygons must be triangles.
    """
    with open(path_to_file, "rb") as f:
        buf = f.read()
    head = buf.split(b"\n", 3)
    if len(head) < 4 or not head[0].startswith(b"# vtk DataFile"):
        raise ValueError("not a legacy VTK file: %s" % path_to_file)
    mode = head[2].strip().upper()
    if mode not in (b"ASCII", b"BINARY"):
        raise ValueError("legacy VTK file must be ASCII or BINARY: %s" % path_to_file)
    cur = _LegacyCursor(buf, mode == b"BINARY")
    cur.pos = len(head[0]) + len(head[1]) + len(head[2]) + 3
    points = tris = None
    scalars = {}
    n_points, in_point_data = 0, False
    while True:
        words = cur.line()
        if words is None:
            break
        key = words[0].upper()
        if key == "DATASET":
            if words[1].upper() != "POLYDATA":
                raise ValueError("DATASET must be POLYDATA")
        elif key == "POINTS":
            n_points = int(words[1])
            points = np.ascontiguousarray(cur.values(3 * n_points, words[2]).reshape(-1, 3), dtype=np.float64)
        elif key in ("POLYGONS", "VERTICES", "LINES", "TRIANGLE_STRIPS"):
            a, b = int(words[1]), int(words[2])
            nxt_pos = cur.pos
            nxt = cur.line()
            if nxt is not None and nxt[0].upper() == "OFFSETS":          # version 5.x: a = n_cells + 1, b = connectivity size
                offs = cur.values(a, nxt[1])
                cw = cur.line()
                if cw is None or cw[0].upper() != "CONNECTIVITY":
                    raise ValueError("OFFSETS without CONNECTIVITY")
                conn = cur.values(b, cw[1])
                if key == "POLYGONS":
                    if not np.all(np.diff(offs) == 3):
                        raise ValueError("only triangle meshes are supported")
                    tris = np.ascontiguousarray(conn.reshape(-1, 3), dtype=np.int32)
            else:                                                          # classic: a = n_cells, b = total ints
                cur.pos = nxt_pos
                raw = cur.values(b, "int")
                if key == "POLYGONS":
                    if b != 4 * a or not np.all(raw[0::4] == 3):
                        raise ValueError("only triangle meshes are supported")
                    tris = np.ascontiguousarray(raw.reshape(-1, 4)[:, 1:], dtype=np.int32)
        elif key == "POINT_DATA":
            in_point_data, n_tuples = True, int(words[1])
        elif key == "CELL_DATA":
            in_point_data, n_tuples = False, int(words[1])
        elif key == "SCALARS":
            ncomp = int(words[3]) if len(words) > 3 else 1
            lt_pos = cur.pos
            lt = cur.line()
            if lt is None or lt[0].upper() != "LOOKUP_TABLE":
                cur.pos = lt_pos
            vals = cur.values(n_tuples * ncomp, words[2])
            if in_point_data and ncomp == 1:
                scalars[words[1]] = np.asarray(vals, dtype=np.float64)
        elif key in ("VECTORS", "NORMALS"):
            cur.values(3 * n_tuples, words[2])
        elif key == "TEXTURE_COORDINATES":
            cur.values(int(words[2]) * n_tuples, words[3])
        elif key == "COLOR_SCALARS":
            cur.values(int(words[2]) * n_tuples, "unsigned_char" if cur.binary else "float")
        elif key == "LOOKUP_TABLE":
            cur.values(4 * int(words[2]), "unsigned_char" if cur.binary else "float")
        elif key == "FIELD":
            for _ in range(int(words[2])):
                fw = cur.line()
                if fw is None:
                    break
                if fw[0].upper() == "METADATA":      # 5.x information block: skip to the blank-line-terminated end
                    fw = cur.line()
                    while fw is not None and fw[0].upper() in ("INFORMATION", "NAME", "DATA"):
                        fw = cur.line()
                    if fw is None:
                        break
                ncomp, ntup = int(fw[1]), int(fw[2])
                vals = cur.values(ncomp * ntup, fw[3])
                if in_point_data and ncomp == 1 and ntup == n_points:
                    scalars[fw[0]] = np.asarray(vals, dtype=np.float64)
        # anything else (METADATA, INFORMATION, blank keywords): ignored
    if points is None or tris is None:
        raise ValueError("file has no POINTS/POLYGONS section: %s" % path_to_file)
    if tris.size and (tris.min() < 0 or tris.max() >= n_points):
        raise ValueError("polygon index out of range in %s" % path_to_file)
    return PolyData(points, tris, scalars)


def write_vtk_mesh(mesh, path_to_file, binary=False):
    """Write a legacy (version 3.0, classic POLYGONS layout) ``DATASET POLYDATA`` file: points as double,
    triangles, and the point scalars.  ASCII output uses 17 significant digits, so both encodings round-trip
    bit-exactly through :func:`read_vtk_mesh`."""
    pts, tris = mesh_arrays(mesh)
    scalars = getattr(mesh, "point_scalars", {}) or {}
    n, f = pts.shape[0], tris.shape[0]
    with open(path_to_file, "wb") as out:
        out.write(b"# vtk DataFile Version 3.0\npyfocusr_b200 mesh\n" + (b"BINARY\n" if binary else b"ASCII\n"))
        out.write(b"DATASET POLYDATA\nPOINTS %d double\n" % n)
        if binary:
            out.write(np.ascontiguousarray(pts, dtype=">f8").tobytes() + b"\n")
        else:
            out.write("\n".join(" ".join("%.17g" % v for v in row) for row in pts).encode() + b"\n")
        out.write(b"POLYGONS %d %d\n" % (f, 4 * f))
        cells = np.concatenate([np.full((f, 1), 3, dtype=np.int64), tris.astype(np.int64)], axis=1)
        if binary:
            out.write(np.ascontiguousarray(cells, dtype=">i4").tobytes() + b"\n")
        else:
            out.write("\n".join(" ".join(str(v) for v in row) for row in cells).encode() + b"\n")
        if scalars:
            out.write(b"POINT_DATA %d\n" % n)
            for name, vals in scalars.items():
                out.write(("SCALARS %s double 1\nLOOKUP_TABLE default\n" % name.replace(" ", "_")).encode())
                if binary:
                    out.write(np.ascontiguousarray(vals, dtype=">f8").tobytes() + b"\n")
                else:
                    out.write("\n".join("%.17g" % v for v in np.asarray(vals, dtype=np.float64)).encode() + b"\n")


def mesh_arrays(vtk_mesh):
    """Return ``(points f64 (N,3), tris i32 (F,3))`` for a PolyData-like object.

    Fast paths: our own :class:`PolyData`; real ``vtkPolyData`` through numpy_support when
    VTK is importable.  Otherwise the generic accessor walk the reference itself performs
    (``graph.py:58-62,155-164``).
    """
    if isinstance(vtk_mesh, PolyData):
        return vtk_mesh.points, vtk_mesh.tris
    try:  # real VTK, zero python loops
        from vtk.util.numpy_support import vtk_to_numpy  # type: ignore

        pts = np.ascontiguousarray(vtk_to_numpy(vtk_mesh.GetPoints().GetData()), dtype=np.float64)
        polys = vtk_mesh.GetPolys()
        conn = vtk_to_numpy(polys.GetConnectivityArray())
        offs = vtk_to_numpy(polys.GetOffsetsArray())
        if not np.all(np.diff(offs) == 3):
            raise ValueError("only triangle meshes are supported")
        return pts, np.ascontiguousarray(conn.reshape(-1, 3), dtype=np.int32)
    except (ImportError, AttributeError):
        pass
    n = vtk_mesh.GetNumberOfPoints()
    pts = np.zeros((n, 3))
    for i in range(n):
        pts[i, :] = vtk_mesh.GetPoint(i)
    nc = vtk_mesh.GetNumberOfCells()
    tris = np.zeros((nc, 3), dtype=np.int32)
    for c in range(nc):
        cell = vtk_mesh.GetCell(c)
        if cell.GetNumberOfEdges() != 3:
            raise ValueError("only triangle meshes are supported")
        for e in range(3):
            tris[c, e] = cell.GetEdge(e).GetPointId(0)
    return pts, tris


# ---------------------------------------------------------------------------------------
# synthetic meshes
# ---------------------------------------------------------------------------------------
def _icosahedron():
    t = (1.0 + 5.0**0.5) / 2.0
    v = np.array(
        [
            [-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
            [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
            [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1],
        ],
        dtype=np.float64,
    )
    v /= np.linalg.norm(v, axis=1)[:, None]
    f = np.array(
        [
            [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
            [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
            [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
            [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1],
        ],
        dtype=np.int64,
    )
    return v, f


def icosphere(nu, radius=1.0):
    """Class-I geodesic icosphere of frequency ``nu``: ``10*nu**2+2`` vertices, ``20*nu**2`` faces.

    Vertices are emitted face by face in lattice order (first occurrence wins for points
    shared between icosahedron faces), which keeps graph neighbours close in index — the
    locality the CSR SpMM kernel's L1/shared-memory staging relies on.  Triangles are
    consistently oriented (outward), so the reference's directed edge fill
    (``graph.py:158-178``) produces a symmetric adjacency.
    """
    nu = int(nu)
    if nu < 1:
        raise ValueError("nu must be >= 1")
    v0, f0 = _icosahedron()
    # integer barycentric lattice on one face: (i, j) with i + j <= nu
    ii, jj = np.meshgrid(np.arange(nu + 1), np.arange(nu + 1), indexing="ij")
    keep = (ii + jj) <= nu
    li = ii[keep]
    lj = jj[keep]
    lk = nu - li - lj
    n_loc = li.size
    loc_id = -np.ones((nu + 1, nu + 1), dtype=np.int64)
    loc_id[li, lj] = np.arange(n_loc)
    # local triangles (two orientations of lattice cells)
    a_i, a_j = np.meshgrid(np.arange(nu), np.arange(nu), indexing="ij")
    up = (a_i + a_j) <= nu - 1
    ui, uj = a_i[up], a_j[up]
    tri_up = np.stack([loc_id[ui, uj], loc_id[ui + 1, uj], loc_id[ui, uj + 1]], axis=1)
    dn = (a_i + a_j) <= nu - 2
    di, dj = a_i[dn], a_j[dn]
    tri_dn = np.stack([loc_id[di + 1, dj], loc_id[di + 1, dj + 1], loc_id[di, dj + 1]], axis=1)
    tri_loc = np.concatenate([tri_up, tri_dn], axis=0)

    # exact integer keys for lattice points so shared edge/corner points merge exactly:
    # key = sorted tuple of (icosahedron vertex id, weight) with non-zero weight.
    all_pts = []
    all_keys = []
    for f in f0:
        A, B, C = v0[f[0]], v0[f[1]], v0[f[2]]
        # point = (lk*A + li*B + lj*C)/nu  (lattice index (li,lj) -> weights (lk, li, lj))
        p = (lk[:, None] * A[None] + li[:, None] * B[None] + lj[:, None] * C[None]) / float(nu)
        all_pts.append(p)
        w = np.stack([lk, li, lj], axis=1)  # weights for vertices f[0], f[1], f[2]
        ids = np.broadcast_to(f[None, :], w.shape)
        # encode up to three (id, weight) pairs, zero-weight pairs dropped, sorted by id
        code = np.where(w > 0, ids * (nu + 1) + w, -1)
        code = np.sort(code, axis=1)  # -1 entries first
        base = 12 * (nu + 1) + 1
        key = (code[:, 0] + 1) * base * base + (code[:, 1] + 1) * base + (code[:, 2] + 1)
        all_keys.append(key)
    pts = np.concatenate(all_pts, axis=0)
    keys = np.concatenate(all_keys, axis=0)
    _, first_idx, inverse = np.unique(keys, return_index=True, return_inverse=True)
    # rank unique points by first occurrence to keep face-major lattice order
    order = np.argsort(first_idx, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    vid = rank[inverse]  # global vertex id for every (face, local) point
    verts = pts[first_idx[order]]
    verts /= np.linalg.norm(verts, axis=1)[:, None]
    tris = np.concatenate([vid[tri_loc + k * n_loc] for k in range(20)], axis=0)
    # orientation: make all faces outward (centroid . normal > 0)
    p0, p1, p2 = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    nrm = np.cross(p1 - p0, p2 - p0)
    flip = np.einsum("ij,ij->i", nrm, p0 + p1 + p2) < 0
    tris[flip] = tris[flip][:, [0, 2, 1]]
    assert verts.shape[0] == 10 * nu * nu + 2 and tris.shape[0] == 20 * nu * nu
    return PolyData(verts * float(radius), tris.astype(np.int32))


def perturbed_ellipsoid(nu=39, seed=0, semi_axes=(43.0, 32.0, 33.0), bump_amp=0.05, jitter=0.02,
                        base=None):
    """Perturbed ellipsoid of SURVEY.md §8d config 3.

    Icosphere of frequency ``nu`` (39 -> 15 212 vertices) scaled to ``semi_axes``; radial
    perturbation = sum of 6 random smooth bumps (total amplitude <= ``bump_amp``); vertex
    jitter ``N(0, (jitter * mean_edge)^2)``.  ``base`` may carry a precomputed unit icosphere.
    """
    sph = base if base is not None else icosphere(nu)
    rng = np.random.RandomState(int(seed))
    u = sph.points / np.linalg.norm(sph.points, axis=1)[:, None]
    centers = rng.normal(size=(6, 3))
    centers /= np.linalg.norm(centers, axis=1)[:, None]
    amps = rng.uniform(-1.0, 1.0, size=6) * (bump_amp / 6.0)
    widths = rng.uniform(0.5, 1.2, size=6)
    radial = np.ones(u.shape[0])
    for c, a, w in zip(centers, amps, widths):
        ang2 = np.sum((u - c[None]) ** 2, axis=1)  # chord^2 on unit sphere
        radial += a * np.exp(-ang2 / (2.0 * w * w))
    pts = u * radial[:, None] * np.asarray(semi_axes, dtype=np.float64)[None]
    t = sph.tris
    mean_edge = np.mean(np.linalg.norm(pts[t[:, 0]] - pts[t[:, 1]], axis=1))
    pts = pts + rng.normal(scale=jitter * mean_edge, size=pts.shape)
    return PolyData(pts, t.copy())


def ellipsoid_pair(i, nu=39, base=None):
    """Pair ``i`` of config 3: target seed ``2i``, source seed ``2i+1``."""
    base = base if base is not None else icosphere(nu)
    return (perturbed_ellipsoid(nu, 2 * i, base=base), perturbed_ellipsoid(nu, 2 * i + 1, base=base))
