"""Drop-in for ``pyfocusr.graph`` (reference ``pyfocusr/graph.py``) on the B200 CUDA library.

Same constructor, attributes and methods as the reference ``Graph`` (graph.py:18-354) and
``recursive_eig`` (graph.py:357-389); the arithmetic runs in ``libfocusr_b200.so``:
adjacency / degree / Laplacian assembly (K1), the smallest-k eigensolve with the reference's
retry contract (K2), eigenvector normalisation (B2) and graph smoothing (K5).  Host attributes are
numpy / scipy objects exactly as in the reference (eigsort mutates ``eig_vecs`` in place).

Deliberate differences, all documented in DESIGN.md:
  * ``feature_weights=None`` is accepted (the reference raises AttributeError, graph.py:41-42);
  * curvature features (graph.py:11-15: vtkCurvatures min / max) are computed by the library's own
    kernel (csrc/curvature.cu) instead of VTK; the features-in-G option does not terminate in the
    reference itself and raises NotImplementedError (mesh-scalar and curvature features, features in
    the adjacency and features as coordinates ARE supported);
  * eigenpairs come back in ascending order with a fixed sign convention (ARPACK's order is
    ascending up to near-ties and its sign is random).
"""
from __future__ import annotations

import numpy as np
from scipy import sparse

from . import _lib
from ._device import DeviceGraph
from .mesh import mesh_arrays

__all__ = ["Graph", "recursive_eig"]


def _curvature_arrays(vtk_mesh, names):
    from . import _device

    _lib.require_cuda()
    pts, tris = mesh_arrays(vtk_mesh)
    out = _device.curvatures(pts, tris)
    return [out[n].cpu().numpy() for n in names]


def get_min_max_curvature_values(vtk_mesh):
    """vtk_functions.py:67-74 on the GPU: (minimum, maximum) principal curvature per vertex."""
    return tuple(_curvature_arrays(vtk_mesh, ("minimum", "maximum")))


def get_min_curvature(vtk_mesh):
    """vtk_functions.py:59-64."""
    return _curvature_arrays(vtk_mesh, ("minimum",))


def get_max_curvature(vtk_mesh):
    """vtk_functions.py:51-56."""
    return _curvature_arrays(vtk_mesh, ("maximum",))


# graph.py:11-15
features_dictionary = {
    "curvature": get_min_max_curvature_values,
    "min_curvature": get_min_curvature,
    "max_curvature": get_max_curvature,
}


class Graph(object):
    def __init__(
        self,
        vtk_mesh,
        n_spectral_features=3,
        norm_eig_vecs=True,
        n_rand_samples=10000,
        list_features_to_calc=[],
        list_features_to_get_from_mesh=[],
        feature_weights=None,
        include_features_in_adj_matrix=False,
        include_features_in_G_matrix=False,
        G_matrix_p_function="exp",
        norm_node_features_std=True,
        norm_node_features_cap_std=3,
        norm_node_features_0_1=True,
    ):
        self.vtk_mesh = vtk_mesh
        self.n_spectral_features = n_spectral_features
        self.norm_eig_vecs = norm_eig_vecs
        self.include_features_in_adj_matrix = include_features_in_adj_matrix
        self.include_features_in_G_matrix = include_features_in_G_matrix
        self.G_matrix_p_function = G_matrix_p_function
        self.norm_node_features_std = norm_node_features_std
        self.norm_node_features_cap_std = norm_node_features_cap_std
        self.norm_node_features_0_1 = norm_node_features_0_1

        # graph.py:58-67 (the per-point GetPoint loop is replaced by a vectorised extraction)
        pts, tris = mesh_arrays(vtk_mesh)
        self.points = np.array(pts, dtype=np.float64)
        self._tris = np.ascontiguousarray(tris, dtype=np.int32)
        self.n_points = self.points.shape[0]
        self.pts_scale_range = np.ptp(self.points, axis=0)
        self.max_pts_scale_range = np.max(self.pts_scale_range)
        self.mean_pts_scale_range = np.mean(self.pts_scale_range)
        self.normed_points = (self.points - np.min(self.points, axis=0)) / self.mean_pts_scale_range

        # graph.py:82 creates an empty n x n lil_matrix here; that is 2 n Python lists (30 000 container objects for a 15k
        # mesh), enough to trigger a full garbage-collection pass every other construction (a bimodal 43 / 90 ms for the
        # Focusr constructor on the shipped 15k pair).  An empty CSR matrix is the same placeholder without the objects;
        # get_weighted_adjacency_matrix replaces it either way.
        self.adjacency_matrix = sparse.csr_matrix((self.n_points, self.n_points))
        self.degree_matrix = None
        self.degree_matrix_inv = None
        self.laplacian_matrix = None
        self.G = None

        self.eig_vals = None
        self.eig_vecs = None
        self.eig_val_gap = None
        self.rand_idxs = self.get_list_rand_idxs(n_rand_samples)

        # graph.py:84-119: extra node features (outside the hot path; host-side only)
        self.node_features = []
        for feature in list_features_to_calc:
            self.node_features += list(features_dictionary[feature](self.vtk_mesh))
        for feature in list_features_to_get_from_mesh:
            pd = vtk_mesh.GetPointData()
            found = None
            for idx in range(pd.GetNumberOfArrays()):
                if pd.GetArray(idx).GetName() == feature:
                    found = np.array(np.asarray(pd.GetArray(idx)), dtype=np.float64)
                    break
            if found is None:
                raise Exception("NO SCALARS WITH SPECIFIED NAME: %s" % feature)
            self.node_features.append(found)
        self.norm_node_features(
            norm_using_std=self.norm_node_features_std,
            norm_range_0_to_1=self.norm_node_features_0_1,
            cap_std=self.norm_node_features_cap_std,
        )
        self.n_extra_features = len(self.node_features)
        self.feature_weights = np.eye(self.n_extra_features) if feature_weights is None else feature_weights
        self.mean_xyz_range_scaled_features = [f * self.mean_pts_scale_range for f in self.node_features]
        if self.n_extra_features > 0 and include_features_in_G_matrix:
            raise NotImplementedError(
                "features in the G matrix (graph.py:191-210) are not supported: the reference's own branch does not "
                "terminate with current numpy/scipy (np.ptp of a sparse matrix), so there is nothing to match"
            )
        self._dev = None

    # graph.py:121-142
    def norm_node_features(self, norm_using_std=True, norm_range_0_to_1=True, cap_std=3):
        for idx in range(len(self.node_features)):
            if norm_using_std is True:
                self.node_features[idx] = (
                    self.node_features[idx] - np.mean(self.node_features[idx])
                ) / np.std(self.node_features[idx])
                if cap_std is not False:
                    self.node_features[idx][self.node_features[idx] > cap_std] = cap_std
                    self.node_features[idx][self.node_features[idx] < -cap_std] = -cap_std
            if norm_range_0_to_1 is True:
                self.node_features[idx] = (
                    self.node_features[idx] - np.min(self.node_features[idx])
                ) / np.ptp(self.node_features[idx])

    # --- device graph ---------------------------------------------------------------------------
    def _device_graph(self):
        if self._dev is None:
            edge_pts = None
            if (self.n_extra_features > 0) & (self.include_features_in_adj_matrix is True):
                # graph.py:166-175: the features, scaled to the mean xyz range, extend the position
                edge_pts = [np.concatenate([self.points] + [f[:, None] for f in self.mean_xyz_range_scaled_features], axis=1)]
            self._dev = DeviceGraph([self.points], [self._tris], edge_pts)
        return self._dev

    # graph.py:148-178  (K1)
    def get_weighted_adjacency_matrix(self):
        rp, ci, w = self._device_graph().adjacency_host()
        self.adjacency_matrix = sparse.csr_matrix((w, ci, rp), shape=(self.n_points, self.n_points))

    # graph.py:216-219
    def get_degree_matrix(self):
        dev = self._device_graph()
        self.degree_matrix = sparse.diags(dev.degree.cpu().numpy())
        self.degree_matrix_inv = sparse.diags(dev.degree_inv.cpu().numpy())

    # graph.py:180-214 (default branch only)
    def get_G_matrix(self, p_function="exp"):
        if (self.n_extra_features > 0) & (self.include_features_in_G_matrix is True):
            raise NotImplementedError("features in the G matrix are outside the B200 hot path")
        self.G = self.degree_matrix_inv

    # graph.py:221-226
    def get_laplacian_matrix(self):
        if self.G is None:
            self.G = self.degree_matrix_inv
        rp, ci, v = self._device_graph().laplacian_host()
        self.laplacian_matrix = sparse.csr_matrix((v, ci, rp), shape=(self.n_points, self.n_points))

    # graph.py:228-257
    def get_graph_spectrum(self):
        self.get_weighted_adjacency_matrix()
        self.get_degree_matrix()
        self.get_G_matrix(p_function=self.G_matrix_p_function)
        self.get_laplacian_matrix()
        dev = self._device_graph()
        n = self.n_spectral_features
        vals, vecs, info = dev.eigs_smallest(k=n + 1, n_k_needed=n, k_buffer=1)
        m = int(info["n_found"][0])
        if self.norm_eig_vecs is True:
            dev.normalize_columns(vecs, [m])
        self.eig_vals = vals[0, :m].cpu().numpy()
        self.eig_vecs = np.ascontiguousarray(vecs[:, :m].cpu().numpy())
        self.eigs_info = info

    # graph.py:263-290
    def get_eig_val_gap(self):
        self.eig_val_gap = np.mean(np.diff(self.eig_vals))

    def get_rand_eig_vecs(self):
        return self.eig_vecs[self.rand_idxs, :]

    def get_rand_normalized_points(self):
        return (
            self.points[self.rand_idxs, :] - np.min(self.points[self.rand_idxs, :], axis=0)
        ) / np.ptp(self.points[self.rand_idxs, :], axis=0)

    def get_list_rand_idxs(self, n_rand_samples, replace=False, force_randomization=False):
        if n_rand_samples > self.n_points:
            list_points = np.arange(self.n_points)
            if force_randomization is True:
                np.random.shuffle(list_points)
            return list_points
        return np.random.choice(self.n_points, size=n_rand_samples, replace=replace)

    # graph.py:296-314 (viewers: out of scope)
    def view_mesh_existing_scalars(self):
        raise ImportError("itkwidgets viewers are not part of the B200 hot path")

    view_mesh_eig_vec = view_mesh_features = view_mesh_existing_scalars

    # graph.py:320-354  (K5)
    def mean_filter_graph(self, values, iterations=300):
        torch = _lib.require_cuda()
        dev = self._device_graph()
        v = np.asarray(values, dtype=np.float64)
        one_d = v.ndim == 1
        v2 = np.ascontiguousarray(v.reshape(self.n_points, -1))
        out = dev.mean_filter(torch.from_numpy(v2).to(dev.device), iterations).cpu().numpy()
        return out[:, 0] if one_d else out


def recursive_eig(matrix, k, n_k_needed, k_buffer=1, sigma=1e-10, which="LM"):
    """graph.py:357-389 for any sparse matrix with a real, non-negative low spectrum (the
    reference only ever passes a graph Laplacian).  The matrix is applied as is (general CSR,
    Euclidean Rayleigh-Ritz); eigenvalues <= 1e-10 are dropped and the request grows by
    ``k_buffer + n_k_needed`` until ``n_k_needed`` remain, as the reference does.  ``sigma`` /
    ``which`` are accepted for signature compatibility (only sigma~0, which='LM' is supported)."""
    if which != "LM" or abs(sigma) > 1e-6:
        raise NotImplementedError("only the reference's shift-invert-at-zero call (sigma~0, which='LM') is supported")
    torch = _lib.require_cuda()
    lap = sparse.csr_matrix(matrix, dtype=np.float64)
    lap.sort_indices()
    n = lap.shape[0]
    diag = lap.diagonal()
    off = lap - sparse.diags(diag)
    off.eliminate_zeros()
    off = off.tocsr()
    off.sort_indices()
    g = DeviceGraph.__new__(DeviceGraph)
    dev = torch.device("cuda", torch.cuda.current_device())
    g.device = dev
    g.n_meshes, g.n_points, g.max_mesh_points = 1, n, n
    g.mesh_off_host = np.array([0, n], dtype=np.int32)
    g.mesh_off = torch.from_numpy(g.mesh_off_host).to(dev)
    g.row_ptr = torch.from_numpy(off.indptr.astype(np.int32)).to(dev)
    g.cols = torch.from_numpy(off.indices.astype(np.int32)).to(dev)
    g.weights = torch.from_numpy(-off.data).to(dev)  # y = diag*x - sum_j w_ij x_j
    g.degree = torch.from_numpy(np.ascontiguousarray(diag)).to(dev)
    g.degree_inv = torch.ones(n, dtype=torch.float64, device=dev)
    empty_rows = int(np.sum((np.diff(lap.indptr) == 0)))
    # general matrix: force the Euclidean (non-symmetric) Rayleigh-Ritz path
    max_row = int(np.max(np.diff(off.indptr))) if n else 0
    g.mesh_info_host = np.array([[off.nnz, max(1, off.nnz), empty_rows, 0, max_row, 0, 0, 0]], dtype=np.int32)
    g.nnz = off.nnz
    # start block from a 1-D embedding of the row index (no geometry available)
    t = np.linspace(-1.0, 1.0, n)
    g.points = torch.from_numpy(np.stack([t, np.cos(np.pi * t), np.sin(np.pi * t)], axis=1).copy()).to(dev)
    beta = float(np.max(np.abs(lap).sum(axis=1)))
    vals, vecs, info = g.eigs_smallest(k=k, n_k_needed=n_k_needed, k_buffer=k_buffer, min_eig_val=1e-10,
                                       spectrum_upper_bound=beta)
    m = int(info["n_found"][0])
    return vals[0, :m].cpu().numpy(), np.ascontiguousarray(vecs[:, :m].cpu().numpy())
