"""Drop-in for ``pyfocusr.Focusr`` (reference ``pyfocusr/focusr.py:22-807``) on the B200 library.

Constructor signature, attribute names and method names follow the reference.  On the hot path
 the work runs in libfocusr_b200.so: both graphs' Laplacians and spectra,
eigsort, KNN correspondences (focusr.py:351-366), the 300 + 40 smoothing passes and the second
KNN (focusr.py:368-396), the k=3 weighted positions (focusr.py:401-426) and the nearest-neighbour
gather (focusr.py:428-431), and the ICP pre-alignment of focusr.py:106-131 (csrc/icp.cu).  The extra
keyword ``registration`` (after the reference's last argument) selects the CPD step of
focusr.py:297-334: "b200" (default) runs affine + deformable Coherent Point Drift on the GPU
(pyfocusr_b200.cpd, same call surface as cycpd); "cycpd" uses the reference's own dependency when it
is installed; "identity" skips the step, as in the throughput configuration of BASELINE.md section 3.
"""
from __future__ import annotations

import numpy as np

from . import _device, _lib
from .eigsort import eigsort
from .graph import Graph
from .mesh import PolyData

__all__ = ["Focusr"]


def print_header(message, banner_length=72):
    print("=" * banner_length)
    print("")
    print(message)
    print("")
    print("=" * banner_length)


class IcpTransform(object):
    """What ``vtk_functions.icp_transform`` returns, reduced to what the drop-in needs: the accumulated 4x4
    matrix (``.matrix``; ``GetMatrix().GetElement(i, j)`` for code written against vtkTransform)."""

    def __init__(self, matrix):
        self.matrix = np.asarray(matrix, dtype=np.float64).reshape(4, 4)

    def GetMatrix(self):
        return self

    def GetElement(self, i, j):
        return float(self.matrix[i, j])


def _icp_transform(target, source, transform_mode):
    """vtk_functions.py:12-37 on the GPU (csrc/icp.cu): register ``source`` onto ``target``; returns the transform
    and a copy of ``source`` moved by it.  100 iterations; 1000 landmarks -- the reference sets the landmark count
    after its first Update(), and the re-run triggered by ``apply_transform`` is the one that counts."""
    from .mesh import mesh_arrays

    if transform_mode not in ("rigid", "similarity"):
        raise TypeError("Error invalid transform mode")                       # vtk_functions.py:20-21 raises a str
    _lib.require_cuda()
    tp, tt = mesh_arrays(target)
    sp, st = mesh_arrays(source)
    mat, moved = _device.icp(tp, tt, sp, max_iterations=100, max_landmarks=1000, similarity=transform_mode == "similarity")
    scalars = dict(getattr(source, "point_scalars", {}) or {})
    return IcpTransform(mat), PolyData(moved.cpu().numpy(), np.array(st, copy=True), scalars)


def _mesh_with_points(mesh, points):
    """Copy of ``mesh`` with its vertices moved to ``points`` (focusr.py:605-625)."""
    if isinstance(mesh, PolyData):
        out = mesh.copy()
        out.points = np.array(points, dtype=np.float64)
        return out
    import vtk  # type: ignore  # pragma: no cover

    out = vtk.vtkPolyData()
    out.DeepCopy(mesh)
    pts = out.GetPoints()
    for i in range(points.shape[0]):
        pts.SetPoint(i, points[i])
    return out


class Focusr(object):
    def __init__(
        self,
        vtk_mesh_target,
        vtk_mesh_source,
        icp_register_first=True,
        icp_registration_mode="rigid",
        icp_reg_target_to_source=False,
        n_spectral_features=3,
        n_extra_spectral=3,
        target_eigenmap_as_reference=True,
        norm_physical_and_spectral=True,
        n_coords_spectral_ordering=5000,
        n_coords_spectral_registration=5000,
        rigid_before_non_rigid_reg=True,
        rigid_reg_max_iterations=100,
        rigid_tolerance=1e-8,
        non_rigid_max_iterations=1000,
        non_rigid_tolerance=1e-8,
        non_rigid_alpha=0.5,
        non_rigid_beta=3.0,
        non_rigid_n_eigens=100,
        include_points_as_features=False,
        get_weighted_spectral_coords=True,
        graph_smoothing_iterations=300,
        feature_smoothing_iterations=40,
        smooth_correspondences=True,
        return_average_final_points=True,
        return_nearest_final_points=True,
        return_transformed_mesh=True,
        projection_smooth_iterations=40,
        feature_weights=None,
        initial_correspondence_type="kd",
        final_correspondence_type="kd",
        list_features_to_calc=["curvature"],
        list_features_to_get_from_mesh=[],
        use_features_as_coords=False,
        use_features_in_graph=False,
        include_features_in_adj_matrix=False,
        G_matrix_p_function="exp",
        norm_node_features_std=True,
        norm_node_features_cap_std=3,
        norm_node_features_0_1=True,
        verbose=False,
        registration="b200",
    ):
        self.verbose = verbose
        self.registration = registration
        self.n_spectral_features = n_spectral_features
        self.n_extra_spectral = n_extra_spectral
        self.n_total_spectral_features = self.n_spectral_features + self.n_extra_spectral
        self.target_eigenmap_as_reference = target_eigenmap_as_reference
        self.norm_physical_and_spectral = norm_physical_and_spectral
        self.include_points_as_features = include_points_as_features
        self.get_weighted_spectral_coords = get_weighted_spectral_coords
        self.feature_smoothing_iterations = feature_smoothing_iterations
        self.n_coords_spectral_registration = n_coords_spectral_registration
        self.rigid_before_non_rigid_reg = rigid_before_non_rigid_reg
        self.rigid_reg_max_iterations = rigid_reg_max_iterations
        self.rigid_tolerance = rigid_tolerance
        self.non_rigid_max_iterations = non_rigid_max_iterations
        self.non_rigid_tolerance = non_rigid_tolerance
        self.non_rigid_alpha = non_rigid_alpha
        self.non_rigid_beta = non_rigid_beta
        self.non_rigid_n_eigens = non_rigid_n_eigens
        self.initial_correspondence_type = initial_correspondence_type
        self.smooth_correspondences = smooth_correspondences
        self.return_average_final_points = return_average_final_points
        self.return_nearest_final_points = return_nearest_final_points
        self.graph_smoothing_iterations = graph_smoothing_iterations
        self.projection_smooth_iterations = projection_smooth_iterations
        self.final_correspondence_type = final_correspondence_type
        self.return_transformed_mesh = return_transformed_mesh
        for ctype in (initial_correspondence_type, final_correspondence_type):
            if ctype not in ("kd", "hungarian"):
                raise ValueError("correspondence type must be 'kd' or 'hungarian'")

        # focusr.py:110-131 (vtkIterativeClosestPointTransform restated on the GPU: csrc/icp.cu)
        if icp_register_first is True:
            if icp_reg_target_to_source is True:
                icp, vtk_mesh_target = _icp_transform(vtk_mesh_source, vtk_mesh_target, icp_registration_mode)
            else:
                icp, vtk_mesh_source = _icp_transform(vtk_mesh_target, vtk_mesh_source, icp_registration_mode)
            self._icp_transform = icp

        graph_kwargs = dict(
            n_spectral_features=self.n_total_spectral_features,
            n_rand_samples=n_coords_spectral_ordering,
            list_features_to_calc=list_features_to_calc,
            list_features_to_get_from_mesh=list_features_to_get_from_mesh,
            feature_weights=feature_weights,
            include_features_in_G_matrix=use_features_in_graph,
            include_features_in_adj_matrix=include_features_in_adj_matrix,
            G_matrix_p_function=G_matrix_p_function,
            norm_node_features_std=norm_node_features_std,
            norm_node_features_cap_std=norm_node_features_cap_std,
            norm_node_features_0_1=norm_node_features_0_1,
        )
        # focusr.py:134-169
        self.graph_target = Graph(vtk_mesh_target, **graph_kwargs)
        self.graph_target.get_graph_spectrum()
        self.graph_source = Graph(vtk_mesh_source, **graph_kwargs)
        self.graph_source.get_graph_spectrum()

        # focusr.py:174-208
        self.Q = None
        self.spec_weights = None
        self.spectral_weights = None
        self.source_spectral_coords = None
        self.target_spectral_coords = None
        self.source_extra_features = None
        self.target_extra_features = None
        self.use_features_as_coords = use_features_as_coords
        self.source_spectral_coords_after_rigid = None
        self.source_spectral_coords_b4_reg = None
        self.rigid_params = None
        self.non_rigid_params = None
        self.smoothed_target_coords = None
        self.source_projected_on_target = None
        self.weighted_avg_transformed_mesh = None
        self.nearest_neighbour_transformed_mesh = None
        self.corresponding_target_idx_for_each_source_pt = None
        self.nearest_neighbor_transformed_points = None
        self.weighted_avg_transformed_points = None
        self.average_mesh = None

    # ------------------------------------------------------------------ focusr.py:271-295
    def append_pts_to_spectral_coords(self):
        if self.norm_physical_and_spectral is True:
            self.source_spectral_coords = np.concatenate(
                (self.source_spectral_coords, self.graph_source.normed_points), axis=1)
            self.target_spectral_coords = np.concatenate(
                (self.target_spectral_coords, self.graph_target.normed_points), axis=1)
        else:
            self.source_spectral_coords = np.concatenate(
                (self.source_spectral_coords * self.graph_source.mean_pts_scale_range, self.graph_source.points), axis=1)
            self.target_spectral_coords = np.concatenate(
                (self.target_spectral_coords * self.graph_target.mean_pts_scale_range, self.graph_target.points), axis=1)

    # ------------------------------------------------------------------ focusr.py:218-269
    def append_features_to_spectral_coords(self):
        if self.graph_source.n_extra_features != self.graph_target.n_extra_features:
            raise Exception(
                "Number of extra features between"
                " target ({}) and source ({}) dont match!".format(
                    self.graph_target.n_extra_features, self.graph_source.n_extra_features))
        out = []
        for graph, coords in ((self.graph_source, self.source_spectral_coords),
                              (self.graph_target, self.target_spectral_coords)):
            feats = np.zeros((graph.n_points, graph.n_extra_features))
            for k in range(graph.n_extra_features):
                f = graph.mean_filter_graph(graph.node_features[k], iterations=self.feature_smoothing_iterations)  # K5
                f = f - np.min(f)
                f = f / np.max(f)
                feats[:, k] = np.ptp(coords) * f
            out.append(feats)
        self.source_extra_features, self.target_extra_features = out
        self.source_spectral_coords = np.concatenate((self.source_spectral_coords, self.source_extra_features), axis=1)
        self.target_spectral_coords = np.concatenate((self.target_spectral_coords, self.target_extra_features), axis=1)

    # ------------------------------------------------------------------ focusr.py:297-334 (CPD: cycpd)
    def register_target_to_source(self, reg_type="deformable"):
        if self.registration == "identity":
            return
        if self.registration == "cycpd":
            try:
                import cycpd  # type: ignore
            except ImportError as e:
                raise ImportError(
                    "registration='cycpd' asks for the reference's own CPD dependency (focusr.py:297-334) and cycpd "
                    "is not installed; use registration='b200' (GPU CPD) or 'identity'"
                ) from e
        elif self.registration == "b200":
            from . import cpd as cycpd
        else:
            raise ValueError("registration must be 'b200', 'cycpd' or 'identity'")
        x = self.source_spectral_coords[self.graph_source.get_list_rand_idxs(self.n_coords_spectral_registration), :]
        y = self.target_spectral_coords[self.graph_target.get_list_rand_idxs(self.n_coords_spectral_registration), :]
        if reg_type == "deformable":
            reg = cycpd.deformable_registration(
                X=x, Y=y, num_eig=self.non_rigid_n_eigens, max_iterations=self.non_rigid_max_iterations,
                tolerance=self.non_rigid_tolerance, alpha=self.non_rigid_alpha, beta=self.non_rigid_beta,
                verbose=self.verbose)
            _, self.non_rigid_params = reg.register()
        elif reg_type == "affine":
            reg = cycpd.affine_registration(X=x, Y=y, max_iterations=self.rigid_reg_max_iterations,
                                            tolerance=self.rigid_tolerance)
            _, self.rigid_params = reg.register()
        self.target_spectral_coords = reg.transform_point_cloud(self.target_spectral_coords)

    # ------------------------------------------------------------------ focusr.py:340-366
    def get_hungarian_correspondence(self, target_pts, spectral_pts):
        """focusr.py:340-349: the N x N distance matrix and its assignment, both on the GPU -- ``focusr_cdist`` and
        ``focusr_lsap``, scipy's own shortest-augmenting-path algorithm with its scan order and tie rules in one
        thread-block cluster, so ``target_idx`` is what ``scipy.optimize.linear_sum_assignment`` returns."""
        dist = _device.cdist(spectral_pts, target_pts)
        _, target_idx = _device.linear_sum_assignment(dist)
        self.corresponding_target_idx_for_each_source_pt = target_idx

    def get_kd_correspondence(self, target_pts, spectral_pts):
        torch = _lib.require_cuda()
        refs = torch.from_numpy(np.ascontiguousarray(target_pts, dtype=np.float64)).cuda()
        qs = torch.from_numpy(np.ascontiguousarray(spectral_pts, dtype=np.float64)).cuda()
        idx, _ = _device.knn(refs, qs, k=1, want_dist=False)
        self.corresponding_target_idx_for_each_source_pt = idx[:, 0].cpu().numpy()

    def get_initial_correspondences(self):
        if self.initial_correspondence_type == "kd":
            self.get_kd_correspondence(self.target_spectral_coords, self.source_spectral_coords)
        elif self.initial_correspondence_type == "hungarian":
            self.get_hungarian_correspondence(self.target_spectral_coords, self.source_spectral_coords)

    # ------------------------------------------------------------------ focusr.py:368-396
    def get_smoothed_correspondences(self):
        self.smoothed_target_coords = self.graph_target.mean_filter_graph(
            self.graph_target.points, iterations=self.graph_smoothing_iterations)
        if self.smoothed_target_coords.shape[0] != self.graph_source.n_points and self.initial_correspondence_type == "hungarian":
            raise Exception(   # focusr.py:377-385
                "If number vertices between source & target don't match, initial_correspondence_type must\n"
                "be 'kd' and not 'hungarian'. Current type is: {}".format(self.initial_correspondence_type))
        self.source_projected_on_target = self.graph_source.mean_filter_graph(
            self.smoothed_target_coords[self.corresponding_target_idx_for_each_source_pt, :],
            iterations=self.projection_smooth_iterations)
        if self.final_correspondence_type == "kd":
            self.get_kd_correspondence(self.smoothed_target_coords, self.source_projected_on_target)
        elif self.final_correspondence_type == "hungarian":
            self.get_hungarian_correspondence(self.smoothed_target_coords, self.source_projected_on_target)

    # ------------------------------------------------------------------ focusr.py:401-431
    def get_weighted_final_node_locations(self, n_closest_pts=3):
        if n_closest_pts != 3:
            raise NotImplementedError("the reference only ever uses n_closest_pts=3")
        torch = _lib.require_cuda()
        refs = torch.from_numpy(np.ascontiguousarray(self.smoothed_target_coords)).cuda()
        qs = torch.from_numpy(np.ascontiguousarray(self.source_projected_on_target)).cuda()
        tp = torch.from_numpy(np.ascontiguousarray(self.graph_target.points)).cuda()
        idx3, dist3 = _device.knn(refs, qs, k=3)
        self.weighted_avg_transformed_points = _device.weighted_positions(idx3, dist3, tp).cpu().numpy()

    def get_nearest_neighbour_final_node_locations(self):
        self.nearest_neighbor_transformed_points = self.graph_target.points[
            self.corresponding_target_idx_for_each_source_pt, :]

    # ------------------------------------------------------------------ focusr.py:433-453
    def get_average_shape(self, align_type="weighted"):
        if align_type == "nearest":
            new = self.graph_target.points[self.corresponding_target_idx_for_each_source_pt]
        else:
            new = self.weighted_avg_transformed_points
        self.average_mesh = _mesh_with_points(self.graph_source.vtk_mesh, (self.graph_source.points + new) / 2)

    # ------------------------------------------------------------------ focusr.py:459-508
    def calc_c_weighting_spectral(self):
        self.spectral_weights = self.Q[: self.n_spectral_features] * np.max(
            (self.graph_source.eig_vals[: self.n_spectral_features],
             self.graph_target.eig_vals[: self.n_spectral_features]), axis=0)
        sigma = np.mean(self.spectral_weights)
        self.spectral_weights = np.exp(-(self.spectral_weights**2) / (2 * sigma**2))

    def calc_weighted_spectral_coords(self):
        self.calc_c_weighting_spectral()
        self.source_spectral_coords = (
            self.graph_source.eig_vecs[:, : self.n_spectral_features] * self.spectral_weights[None, :])
        self.target_spectral_coords = (
            self.graph_target.eig_vecs[:, : self.n_spectral_features] * self.spectral_weights[None, :])

    def calc_spectral_coords(self):
        if self.get_weighted_spectral_coords is True:
            self.calc_weighted_spectral_coords()
        else:
            self.source_spectral_coords = self.graph_source.eig_vecs[:, : self.n_spectral_features]
            self.target_spectral_coords = self.graph_target.eig_vecs[:, : self.n_spectral_features]

    # ------------------------------------------------------------------ focusr.py:514-568
    def align_maps(self):
        eig_map_sorter = eigsort(graph_target=self.graph_target, graph_source=self.graph_source,
                                 n_features=self.n_total_spectral_features,
                                 target_as_reference=self.target_eigenmap_as_reference)
        self.Q = eig_map_sorter.sort_eigenmaps()
        self.calc_spectral_coords()
        if (self.graph_source.n_extra_features > 0) & (self.use_features_as_coords is True):
            self.append_features_to_spectral_coords()
        if self.include_points_as_features is True:
            self.append_pts_to_spectral_coords()
        self.source_spectral_coords_b4_reg = np.copy(self.source_spectral_coords)
        if self.rigid_before_non_rigid_reg is True:
            self.register_target_to_source(reg_type="affine")
            self.source_spectral_coords_after_rigid = np.copy(self.source_spectral_coords)
        self.register_target_to_source("deformable")
        self.get_initial_correspondences()
        if self.smooth_correspondences is True:
            self.get_smoothed_correspondences()
        if self.return_average_final_points is True:
            self.get_weighted_final_node_locations()
        if self.return_nearest_final_points is True:
            self.get_nearest_neighbour_final_node_locations()
        if self.return_transformed_mesh is True:
            if self.return_average_final_points is True:
                self.get_source_mesh_transformed_weighted_avg()
            if self.return_nearest_final_points is True:
                self.get_source_mesh_transformed_nearest_neighbour()

    # ------------------------------------------------------------------ focusr.py:605-625
    def get_source_mesh_transformed_weighted_avg(self):
        self.weighted_avg_transformed_mesh = _mesh_with_points(
            self.graph_source.vtk_mesh, self.weighted_avg_transformed_points)

    def get_source_mesh_transformed_nearest_neighbour(self):
        self.nearest_neighbour_transformed_mesh = _mesh_with_points(
            self.graph_source.vtk_mesh, self.nearest_neighbor_transformed_points)

    # ------------------------------------------------------------------ focusr.py:576-599 (scalars for visualisation)
    @staticmethod
    def _set_scalars(mesh, values):
        if isinstance(mesh, PolyData):
            mesh.GetPointData().SetScalars(np.asarray(values))
        else:  # a real vtkPolyData  # pragma: no cover
            from vtk.util.numpy_support import numpy_to_vtk  # type: ignore

            mesh.GetPointData().SetScalars(numpy_to_vtk(np.asarray(values)))

    def set_transformed_source_scalars_to_corresp_target_idx(self):
        for mesh in (self.weighted_avg_transformed_mesh, self.nearest_neighbour_transformed_mesh):
            if mesh is not None:
                self._set_scalars(mesh, self.corresponding_target_idx_for_each_source_pt)

    def set_source_scalars_to_corresp_target_idx(self):
        self._set_scalars(self.graph_source.vtk_mesh, self.corresponding_target_idx_for_each_source_pt)

    def set_target_scalars_to_corresp_target_idx(self):
        self._set_scalars(self.graph_target.vtk_mesh, np.arange(self.graph_target.n_points))

    def set_all_mesh_scalars_to_corresp_target_idx(self):
        self.set_target_scalars_to_corresp_target_idx()
        self.set_source_scalars_to_corresp_target_idx()
        self.set_transformed_source_scalars_to_corresp_target_idx()

    # viewers (focusr.py:646-795) need itkwidgets: out of scope
    def view_aligned_spectral_coords(self, *a, **k):
        raise ImportError("itkwidgets viewers are not part of the B200 hot path")

    view_meshes = view_meshes_colored_by_spectral_correspondences = view_aligned_smoothed_spectral_coords = (
        view_aligned_spectral_coords)
