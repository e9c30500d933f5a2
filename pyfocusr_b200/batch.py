"""Batched spectral-correspondence stage: many (target, source) mesh pairs per launch.

This is the throughput path of BASELINE.json config 3 ("batch of 1024 synthetic 15k-vertex mesh
pairs sharded across 1/2/4/8 B200").  It is the same sequence ``Focusr.__init__`` +
``Focusr.align_maps`` run for one pair (reference focusr.py:134-169, 514-562, CPD = identity as
in BASELINE.md section 3), but every kernel works on the block-diagonal graph of all 2P meshes, so
one SpMM launch streams gigabytes instead of 1.3 MB and nothing visits the host between the upload
of the vertices and the download of the correspondences except the eigensolver's convergence test
(a few hundred Ritz values and residuals per outer iteration).

Mesh order inside the batch graph: targets 0..P-1, then sources 0..P-1.
"""
from __future__ import annotations

import os
import threading

import numpy as np

from . import _device, _lib
from ._device import DeviceGraph

__all__ = ["SpectralBatch"]


class SpectralBatch:
    def __init__(self, n_spectral_features=3, n_extra_spectral=3, n_coords_spectral_ordering=5000,
                 graph_smoothing_iterations=300, projection_smooth_iterations=40,
                 target_eigenmap_as_reference=True, get_weighted_spectral_coords=True, tol=1e-10,
                 block_size=0, seed=0, registration="identity", n_coords_spectral_registration=5000,
                 rigid_before_non_rigid_reg=True, rigid_reg_max_iterations=100, rigid_tolerance=1e-8,
                 non_rigid_max_iterations=1000, non_rigid_tolerance=1e-8, non_rigid_alpha=0.5, non_rigid_beta=3.0,
                 non_rigid_n_eigens=100, cpd_streams=8):
        self.ns = int(n_spectral_features)
        self.n = int(n_spectral_features + n_extra_spectral)
        self.n_samples = int(n_coords_spectral_ordering)
        self.graph_smoothing_iterations = int(graph_smoothing_iterations)
        self.projection_smooth_iterations = int(projection_smooth_iterations)
        self.target_as_reference = bool(target_eigenmap_as_reference)
        self.weighted = bool(get_weighted_spectral_coords)
        self.tol = float(tol)
        self.block_size = int(block_size)
        self.seed = int(seed)
        # focusr.py:297-334, 537-543: "identity" (BASELINE.json's throughput configuration) or "b200" = affine +
        # deformable CPD on the GPU, one pair after the other (pyfocusr_b200.cpd; Focusr's keyword names)
        if registration not in ("identity", "b200"):
            raise ValueError("registration must be 'identity' or 'b200'")
        self.registration = registration
        self.n_coords_spectral_registration = int(n_coords_spectral_registration)
        self.rigid_before_non_rigid_reg = bool(rigid_before_non_rigid_reg)
        self.cpd_streams = max(1, int(cpd_streams))
        self.cpd_kwargs = dict(rigid_reg_max_iterations=rigid_reg_max_iterations, rigid_tolerance=rigid_tolerance,
                               non_rigid_max_iterations=non_rigid_max_iterations, non_rigid_tolerance=non_rigid_tolerance,
                               non_rigid_alpha=non_rigid_alpha, non_rigid_beta=non_rigid_beta,
                               non_rigid_n_eigens=non_rigid_n_eigens)
        self._tls = threading.local()
        self._pool = None
        self.eigs_options = None  # _lib.EigsOptions or dict: per-call options of the eigensolver (A/B measurements)
        # The smoothing of the target vertices (graph.py:349-354, 300 passes) depends only on the target graphs, not on
        # the spectral stages: with overlap_smoothing it is enqueued on a second CUDA stream right after the Laplacian
        # build and joined where its result is first read; its CTAs fill the gaps the latency-bound filter steps leave
        # (B200, 128 pairs: 1059 -> 1127 pairs/s).  Same kernels, same results; `timings` then attributes the overlapped
        # passes to the eigensolve stage.
        self.overlap_smoothing = os.environ.get("FOCUSR_OVERLAP_SMOOTHING", "1") == "1"
        self.timings = {}

    # ------------------------------------------------------------------------------------------
    def sample_indices(self, sizes, rng=None):
        """Graph.get_list_rand_idxs (graph.py:274-290) for every mesh; all meshes must yield the
        same count (n_samples <= every size, or n_samples > every size)."""
        rng = rng or np.random.RandomState(self.seed)
        out = []
        for sz in sizes:
            if self.n_samples > sz:
                out.append(np.arange(sz, dtype=np.int64))
            else:
                out.append(rng.choice(sz, size=self.n_samples, replace=False).astype(np.int64))
        lens = {len(o) for o in out}
        if len(lens) != 1:
            raise ValueError("meshes of a batch must yield the same number of ordering samples")
        return np.stack(out)

    # ------------------------------------------------------------------------------------------
    def run(self, points, tris, mesh_off_host, n_pairs, idx_t=None, idx_s=None, record_events=False,
            keep_presort=False):
        """points: torch tensor [n_points_total, 3] float64 (host pinned or device), meshes in the
        order targets then sources; tris: int32 [n_tris_total, 3] with GLOBAL vertex ids;
        mesh_off_host: int32 [2P+1].  Returns a dict of device tensors + host metadata."""
        torch = _lib.require_cuda()
        P = int(n_pairs)
        ev = []

        def mark(name):
            if record_events:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev.append((name, e))

        mark("start")
        g = DeviceGraph.from_device(points, tris, mesh_off_host)
        mark("laplacian")
        n, ns = self.n, self.ns
        nt_total = int(g.mesh_off_host[P])
        smoothed_t, side = None, None
        if self.overlap_smoothing:
            main = torch.cuda.current_stream()
            if getattr(self._tls, "side_stream", None) is None:  # one stream per host thread for the life of the object:
                self._tls.side_stream = torch.cuda.Stream(device=g.device)  # the caching allocator keeps a pool per stream
            side = self._tls.side_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                smoothed_t = g.mean_filter(g.points, self.graph_smoothing_iterations, 0, nt_total)
            smoothed_t.record_stream(main)
        try:
            vals, vecs, info = g.eigs_smallest(k=n + 1, n_k_needed=n, k_buffer=1, tol=self.tol,
                                               block_size=self.block_size, options=self.eigs_options)
        except _lib.FocusrB200Error as exc:
            info = getattr(exc, "eigs_info", None)
            if info is None:
                raise
            bad = np.nonzero(info["status"] != 0)[0]
            raise RuntimeError("eigensolve failed for %d of %d meshes: %s (mesh m is the %s of pair m %% %d); the other "
                               "meshes of the batch were solved -- drop the named pairs and rerun" % (
                                   bad.size, 2 * P, ", ".join("mesh %d: %s" % (m, _device.EIG_STATUS.get(int(info["status"][m]), "?"))
                                                              for m in bad[:16]), "target if m < %d else source" % P, P)) from exc
        n_found = info["n_found"]
        if int(n_found.min()) < n:
            short = np.nonzero(n_found < n)[0]
            raise RuntimeError("meshes %s returned fewer than %d eigenpairs (pairs %s)" % (
                short[:16].tolist(), n, sorted(set(int(m) % P for m in short[:16]))))
        g.normalize_columns(vecs, n_found)
        mark("eigensolve")

        off = g.mesh_off_host
        sizes = np.diff(off)
        if idx_t is None:
            rng = np.random.RandomState(self.seed)
            idx_t = self.sample_indices(sizes[:P], rng)
            idx_s = self.sample_indices(sizes[P:], rng)
        t_mesh = np.arange(P, dtype=np.int32)
        s_mesh = np.arange(P, 2 * P, dtype=np.int32)
        ch, chf, cs, csf, _ = _device.eigsort_costs(g, vecs, t_mesh, s_mesh, idx_t, idx_s, n)
        # n x n decisions per pair (eigsort.py:66-122, 142-160; focusr.py:459-490) on the device: c_lambda, min(c, c_f),
        # the assignment (scipy's algorithm and tie rules), flips, column moves and spectral weights -- no host visit
        Q, dst, src, sign, weights, decide_status = _device.eigsort_decide(
            g, vals, n_found, t_mesh, s_mesh, (ch, chf, cs, csf), n, ns, self.target_as_reference, self.weighted)
        costs = (ch, chf, cs, csf)
        presort = vecs.clone() if keep_presort else None
        g.flip_permute(vecs, dst, src, sign)
        coords = g.spectral_coords(vecs, weights, ns)
        mark("eigsort")

        dev = g.device
        ref_off = g.mesh_off[: P + 1].contiguous()
        qry_off = (g.mesh_off[P:] - nt_total).contiguous()
        max_q, max_r = int(sizes[P:].max()), int(sizes[:P].max())
        cpd_idx, coords_b4_reg = None, None
        if self.registration == "b200":  # register every pair's target coordinates onto its source's
            coords_b4_reg = coords.clone()
            cpd_idx = self._register_pairs(coords, off, P, sizes)
            mark("cpd")
        idx0, _ = _device.knn(coords[:nt_total], coords[nt_total:], k=1, ref_off=ref_off, query_off=qry_off,
                              max_queries=max_q, max_refs=max_r, want_dist=False)
        mark("knn_initial")
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        else:
            smoothed_t = g.mean_filter(g.points, self.graph_smoothing_iterations, 0, nt_total)
        base_q = torch.repeat_interleave(g.mesh_off[:P], torch.from_numpy(sizes[P:].astype(np.int64)).to(dev)).to(torch.int32)
        staged = torch.empty_like(g.points)
        _lib.call("focusr_gather_rows", _lib.ptr(smoothed_t), _lib.ptr(idx0), _lib.ptr(base_q), g.n_points - nt_total, 3,
                  _lib.ptr(staged[nt_total:]), _lib.stream_ptr())
        src_proj = g.mean_filter(staged, self.projection_smooth_iterations, nt_total, g.n_points)
        mark("smoothing")
        # one k=3 search serves both focusr.py:391-392 (k=1: its first column, same distances and tie
        # rule) and focusr.py:409-413 (k=3)
        idx3, dist3 = _device.knn(smoothed_t[:nt_total], src_proj[nt_total:], k=3, ref_off=ref_off, query_off=qry_off,
                                  max_queries=max_q, max_refs=max_r)
        idx1 = idx3[:, :1]
        weighted = _device.weighted_positions(idx3, dist3, g.points, base_q)
        nearest = _device.gather_rows(g.points, idx1[:, 0].contiguous(), base_q)
        mark("knn_final")
        if record_events:
            torch.cuda.synchronize()
            self.timings = {ev[i][0]: ev[i - 1][1].elapsed_time(ev[i][1]) for i in range(1, len(ev))}
        return dict(graph=g, eig_vals=vals, eig_vecs=vecs, eigs_info=info, Q=Q, spectral_weights=weights[:P],
                    eigsort_status=decide_status,
                    coords=coords, initial_idx=idx0[:, 0], smoothed_target_coords=smoothed_t[:nt_total],
                    source_projected_on_target=src_proj[nt_total:], final_idx=idx1[:, 0], knn3_idx=idx3,
                    knn3_dist=dist3, weighted_avg_transformed_points=weighted,
                    nearest_neighbor_transformed_points=nearest, idx_t=idx_t, idx_s=idx_s, costs=costs,
                    eig_vecs_presort=presort, cpd_idx=cpd_idx, coords_b4_reg=coords_b4_reg)

    # ------------------------------------------------------------------------------------------
    def run_concurrent(self, sub_batches, **kw):
        """Several independent sub-batches at once, one host thread and one CUDA stream each: ``sub_batches`` is a list
        of ``(points, tris, mesh_off_host, n_pairs)`` or of dicts of ``run`` keyword arguments.  A batch is a chain of
        HBM-bound stages (Laplacian, filter steps, smoothing) followed by FP64-ALU / latency-bound ones (eigsort costs,
        KNN) with a few host visits in between; two sub-batches in flight let one's KNN and host visits hide under the
        other's filter steps (B200, 128 pairs as 2 x 64: 1186 -> see DESIGN.md).  Results are those of ``run`` on each
        sub-batch (deterministic kernels, no shared state); returns the list of result dicts in order."""
        from concurrent.futures import ThreadPoolExecutor

        torch = _lib.require_cuda()
        jobs = [dict(zip(("points", "tris", "mesh_off_host", "n_pairs"), b)) if not isinstance(b, dict) else dict(b)
                for b in sub_batches]
        if len(jobs) == 1:
            return [self.run(**jobs[0], **kw)]
        if self._pool is None or self._pool._max_workers < len(jobs):
            self._pool = ThreadPoolExecutor(max_workers=len(jobs))  # kept: workers own pinned pools and streams
        main = torch.cuda.current_stream()
        device = torch.cuda.current_device()

        def work(job):
            with torch.cuda.device(device):
                if getattr(self._tls, "stream", None) is None:
                    self._tls.stream = torch.cuda.Stream(device=device)
                st = self._tls.stream
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    out = self.run(**job, **kw)
                return out, st

        def hand_over(obj):  # results were allocated on the worker's stream and are used on the caller's from here on
            for v in (obj.values() if isinstance(obj, dict) else vars(obj).values()):
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    v.record_stream(main)
                elif isinstance(v, (tuple, list)):
                    for w in v:
                        if isinstance(w, torch.Tensor) and w.is_cuda:
                            w.record_stream(main)

        results = []
        for out, st in self._pool.map(work, jobs):
            main.wait_stream(st)
            hand_over(out)
            if out.get("graph") is not None:
                hand_over(out["graph"])
            results.append(out)
        return results

    # ------------------------------------------------------------------------------------------
    def run_pipelined(self, jobs, depth=2, consume=None, **kw):
        """A stream of independent batches with at most ``depth`` of them in flight, one host thread and one CUDA stream
        each: ``jobs`` is an iterable of dicts of ``run`` keyword arguments (successive steps of a production queue).
        Unlike ``run_concurrent`` the launches keep their full size -- every kernel works on a whole batch -- while the
        FP64-ALU / latency-bound tail of one batch (eigsort, KNN, host visits) hides under the HBM-bound filter steps of
        the next.  ``consume(out, k)`` runs in the worker thread right after batch k has been enqueued, with its stream
        current (read results back there, keep what is needed); what it returns is collected in order.  Default: the
        whole result dict is kept (its tensors are handed over to the caller's stream)."""
        from concurrent.futures import ThreadPoolExecutor

        torch = _lib.require_cuda()
        jobs = [dict(j) for j in jobs]
        depth = max(1, min(int(depth), len(jobs)))
        main = torch.cuda.current_stream()
        device = torch.cuda.current_device()
        if depth == 1:
            return [consume(self.run(**j, **kw), k) if consume else self.run(**j, **kw) for k, j in enumerate(jobs)]
        if self._pool is None or self._pool._max_workers < depth:
            self._pool = ThreadPoolExecutor(max_workers=depth)  # kept: workers own pinned pools and streams
        streams = set()
        lock = threading.Lock()

        def work(item):
            k, job = item
            with torch.cuda.device(device):
                if getattr(self._tls, "stream", None) is None:
                    self._tls.stream = torch.cuda.Stream(device=device)
                st = self._tls.stream
                with lock:
                    streams.add(st)
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    out = self.run(**job, **kw)
                    if consume is not None:
                        return consume(out, k)
                    for obj in (out, out.get("graph")):
                        if obj is None:
                            continue
                        for v in (obj.values() if isinstance(obj, dict) else vars(obj).values()):
                            if isinstance(v, torch.Tensor) and v.is_cuda:
                                v.record_stream(main)
                    return out

        # ThreadPoolExecutor hands items to idle workers in order: at most `depth` batches are in flight
        results = list(self._pool.map(work, list(enumerate(jobs))))
        for st in streams:
            main.wait_stream(st)
        return results

    # ------------------------------------------------------------------------------------------
    def _register_pairs(self, coords, off, P, sizes):
        """focusr.py:537-543 per pair: affine on fresh random subsets, transform all target coordinates, then
        deformable on fresh subsets, transform again.  Returns the index draws [(s_aff, t_aff, s_def, t_def)].

        The pairs are independent, and one registration leaves most of the GPU idle during its latency-bound
        steps (the r x r solve, the host Rayleigh-Ritz of the low-rank set-up), so ``cpd_streams`` host threads
        run registrations concurrently, one CUDA stream each (ctypes releases the GIL inside the library).  All
        random draws are made up front in pair order: the result does not depend on the thread schedule."""
        from concurrent.futures import ThreadPoolExecutor

        from .cpd import affine_registration, deformable_registration

        torch = _lib.require_cuda()
        rng = np.random.RandomState(self.seed + 7919)
        kw = self.cpd_kwargs
        stages = (["affine"] if self.rigid_before_non_rigid_reg else []) + ["deformable"]

        def draw(n):  # Graph.get_list_rand_idxs (graph.py:274-290)
            if self.n_coords_spectral_registration > n:
                return np.arange(n, dtype=np.int64)
            return rng.choice(n, size=self.n_coords_spectral_registration, replace=False).astype(np.int64)

        draws = []
        for p in range(P):
            nt, ns = int(off[p + 1] - off[p]), int(off[P + p + 1] - off[P + p])
            rec = []
            for _ in stages:
                rec += [draw(ns), draw(nt)]   # source first, as the reference's dict literal evaluates
            draws.append(tuple(rec))

        main = torch.cuda.current_stream()
        device = coords.device

        def work(pairs):
            st = torch.cuda.Stream(device=device)
            st.wait_stream(main)
            with torch.cuda.device(device), torch.cuda.stream(st):
                for p in pairs:
                    t0, t1, s0, s1 = int(off[p]), int(off[p + 1]), int(off[P + p]), int(off[P + p + 1])
                    tgt, src = coords[t0:t1], coords[s0:s1]
                    for k, stage in enumerate(stages):
                        i_s, i_t = draws[p][2 * k], draws[p][2 * k + 1]
                        x = src[torch.from_numpy(i_s).to(device)]
                        y = tgt[torch.from_numpy(i_t).to(device)]
                        if stage == "affine":
                            reg = affine_registration(X=x, Y=y, max_iterations=kw["rigid_reg_max_iterations"],
                                                      tolerance=kw["rigid_tolerance"])
                        else:
                            reg = deformable_registration(X=x, Y=y, num_eig=kw["non_rigid_n_eigens"],
                                                          max_iterations=kw["non_rigid_max_iterations"],
                                                          tolerance=kw["non_rigid_tolerance"], alpha=kw["non_rigid_alpha"],
                                                          beta=kw["non_rigid_beta"])
                        reg.register()
                        tgt.copy_(reg.transform_point_cloud(tgt.contiguous()))
            st.synchronize()
            return st

        n_workers = min(self.cpd_streams, P)
        chunks = [list(range(w, P, n_workers)) for w in range(n_workers)]
        if n_workers == 1:
            work(chunks[0])
        else:
            with ThreadPoolExecutor(max_workers=n_workers) as pool:
                for st in pool.map(work, chunks):
                    main.wait_stream(st)
        return draws

    def fetch(self, out, keys=("final_idx", "weighted_avg_transformed_points"), slot=0):
        """Device -> host read-back of the per-vertex results into reusable pinned buffers (one
        asynchronous copy each, then a single synchronisation).  Returns numpy views that stay valid
        until the next ``fetch`` with the same ``slot`` (one slot per sub-batch of ``run_concurrent``; ``slot=None`` =
        one slot per calling thread, for the workers of ``run_pipelined``)."""
        torch = _lib.require_cuda()
        if slot is None:
            slot = ("thread", threading.get_ident())
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        res = {}
        for k in keys:
            t = out[k]
            buf = self._pinned.get((slot, k))
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                self._pinned[(slot, k)] = buf
            buf.copy_(t, non_blocking=True)
            res[k] = buf
        torch.cuda.current_stream().synchronize()
        return {k: v.numpy() for k, v in res.items()}

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def pack_meshes(targets, sources):
        """Lists of PolyData-like meshes (``.points``, ``.tris``) -> ``(points, tris, mesh_off_host, n_pairs)`` of ``run``
        (host tensors; targets first, then sources; triangle vertex ids made global)."""
        import torch

        if len(targets) != len(sources) or not targets:
            raise ValueError("need as many source meshes as target meshes (at least one pair)")
        meshes = list(targets) + list(sources)
        sizes = [m.points.shape[0] for m in meshes]
        off = np.zeros(len(meshes) + 1, dtype=np.int32)
        off[1:] = np.cumsum(sizes)
        pts = torch.from_numpy(np.ascontiguousarray(np.concatenate([m.points for m in meshes])))
        tris = torch.from_numpy(np.ascontiguousarray(np.concatenate(
            [m.tris.astype(np.int64) + int(o) for m, o in zip(meshes, off[:-1])]).astype(np.int32)))
        return pts, tris, off, len(targets)

    def run_meshes(self, targets, sources, **kw):
        """Convenience: lists of PolyData-like meshes (``.points``, ``.tris``)."""
        pts, tris, off, n_pairs = self.pack_meshes(targets, sources)
        return self.run(pts, tris, off, n_pairs, **kw)
