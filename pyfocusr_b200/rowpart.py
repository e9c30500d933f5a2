"""Row-partitioned eigensolve of ONE large mesh across the GPUs of a node
(BASELINE.json configs[3]: "synthetic 1M-vertex icosphere ... row-partitioned ... NCCL halo over
NVLink").

The vertices are first ordered along the 3-D Morton (Z-order) curve of their positions, so that a
contiguous block of rows is a compact patch of the surface whatever order the file listed them in;
rank r owns one such block.  A rank never holds the whole graph: it keeps the triangles that touch
its rows, numbers their vertices [own rows | ghosts sorted by global id] and assembles only that
sub-mesh on its GPU with the same K1 kernels (``focusr_laplacian_build``) -- the rows of its own
vertices come out complete, with the columns already in the local numbering the solver wants.
The host-side partition logic below is pure numpy (tested on CPU for world sizes 1..8); the solve is
``focusr_eigs_smallest_dist``: the Chebyshev-filtered subspace iteration of the single-GPU path with
the halo fused into the kernels over NVLink peer memory (default) or exchanged by grouped
ncclSend/ncclRecv, and small all-reduces for the Gram blocks, residuals and norms.
``torch.distributed`` is used only to agree on the NCCL unique id / IPC handles and to exchange the
ghost lists once.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import dist as fdist

__all__ = ["row_bounds", "morton_order", "partition_mesh", "local_partition", "send_lists", "push_lists",
           "RowPartitionedSolver"]


def row_bounds(n_rows, world):
    """Contiguous, nearly equal row blocks: bounds[r] .. bounds[r+1]."""
    base, rem = divmod(int(n_rows), int(world))
    sizes = np.full(world, base, dtype=np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def _spread_bits_21(v):
    """Bits of a 21-bit integer moved to every third position (the classic magic-number interleave)."""
    v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def morton_order(points):
    """Permutation ``order`` (new index -> old index) that sorts the vertices along the Morton curve of their
    positions, quantised to 21 bits per axis over the bounding box (ties keep the input order)."""
    p = np.asarray(points, dtype=np.float64)
    lo, hi = p.min(axis=0), p.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    q = np.minimum(((p - lo) / span * 2097151.0).astype(np.uint64), np.uint64(2097151))
    key = _spread_bits_21(q[:, 0]) | (_spread_bits_21(q[:, 1]) << np.uint64(1)) | (_spread_bits_21(q[:, 2]) << np.uint64(2))
    return np.argsort(key, kind="stable").astype(np.int64)


def partition_mesh(points, tris, world, rank, reorder="morton"):
    """What ``rank`` of ``world`` keeps of the mesh.  Returns a dict:
    ``order`` (new -> old vertex id), ``bounds`` (row blocks in the new order), ``n_local``, ``row_begin``,
    ``ghosts`` (sorted new global ids of the foreign vertices its triangles touch, hence grouped by owner),
    ``points_local`` [n_local + n_ghost][3] and ``tris_local`` (the triangles touching its rows, in the local numbering
    [own rows | ghosts]), ``recv_counts`` [world]."""
    points = np.asarray(points, dtype=np.float64)
    tris = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
    n = points.shape[0]
    if reorder == "morton":
        order = morton_order(points)
    elif reorder in (None, "none"):
        order = np.arange(n, dtype=np.int64)
    else:
        raise ValueError("reorder must be 'morton' or None")
    inv = np.empty(n, dtype=np.int64)
    inv[order] = np.arange(n, dtype=np.int64)
    t_new = inv[tris]
    bounds = row_bounds(n, world)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    n_loc = r1 - r0
    mine = ((t_new >= r0) & (t_new < r1)).any(axis=1)
    t_sel = t_new[mine]
    verts = np.unique(t_sel)
    ghosts = verts[(verts < r0) | (verts >= r1)]
    is_own = (t_sel >= r0) & (t_sel < r1)
    t_loc = np.where(is_own, t_sel - r0, n_loc + np.searchsorted(ghosts, t_sel)).astype(np.int32)
    owner = np.searchsorted(bounds[1:], ghosts, side="right")
    recv_counts = np.bincount(owner, minlength=world).astype(np.int32)
    pts_loc = np.ascontiguousarray(np.concatenate([points[order[r0:r1]], points[order[ghosts]]]))
    return dict(order=order, bounds=bounds, n_local=n_loc, row_begin=r0, ghosts=ghosts, points_local=pts_loc,
                tris_local=np.ascontiguousarray(t_loc), recv_counts=recv_counts)


def local_partition(indptr, indices, bounds, rank):
    """Local CSR of rank's rows cut out of a GLOBAL CSR, columns remapped to [local | ghosts] numbering (the form the
    solver takes; ``partition_mesh`` + the device build produce the same thing without the global matrix).

    Returns dict(row_ptr, cols_local, entry_slice, ghosts (sorted global ids, hence grouped by owner),
    recv_counts [world])."""
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    world = len(bounds) - 1
    e0, e1 = int(indptr[r0]), int(indptr[r1])
    cols = np.asarray(indices[e0:e1], dtype=np.int64)
    is_local = (cols >= r0) & (cols < r1)
    ghosts = np.unique(cols[~is_local])
    owner = np.searchsorted(bounds[1:], ghosts, side="right")
    recv_counts = np.bincount(owner, minlength=world).astype(np.int32)
    n_loc = r1 - r0
    cols_local = np.where(is_local, cols - r0, n_loc + np.searchsorted(ghosts, cols)).astype(np.int32)
    row_ptr = (np.asarray(indptr[r0 : r1 + 1], dtype=np.int64) - e0).astype(np.int32)
    return dict(row_ptr=row_ptr, cols_local=cols_local, entry_slice=(e0, e1), ghosts=ghosts, recv_counts=recv_counts,
                n_local=n_loc, row_begin=r0)


def send_lists(all_ghosts, bounds, rank):
    """What ``rank`` must ship: for every peer p (in rank order) the local ids of the rows of
    ``rank`` that appear in p's ghost list.  Returns (send_idx int32, send_counts int32 [world])."""
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    idx, counts = [], []
    for p, gh in enumerate(all_ghosts):
        gh = np.asarray(gh, dtype=np.int64)
        mine = gh[(gh >= r0) & (gh < r1)] - r0 if p != rank else np.zeros(0, dtype=np.int64)
        idx.append(mine)
        counts.append(mine.size)
    return np.concatenate(idx).astype(np.int32), np.asarray(counts, dtype=np.int32)


def push_lists(send_idx, send_counts, recv_counts_all, rank):
    """The halo as a PUSH: which of ``rank``'s rows go where.  ``send_idx`` / ``send_counts`` are ``send_lists`` of this
    rank, ``recv_counts_all[q]`` the ``recv_counts`` of rank q.  Rank q's ghost list is sorted by global id, so this rank's
    rows sit in it contiguously, in the order of its own send list, after the ghosts q gets from lower ranks.  Returns
    (rows int64, dst int64) sorted by row: local row ``rows[i]`` is copied to ghost slot ``dst[i] & 0xffffff`` of peer
    ``dst[i] >> 24``."""
    rows, dsts, so = [], [], 0
    for q in range(len(send_counts)):
        cnt = int(send_counts[q])
        if cnt:
            first_slot = int(np.sum(np.asarray(recv_counts_all[q])[:rank]))
            rows.append(np.asarray(send_idx[so:so + cnt], dtype=np.int64))
            dsts.append((q << 24) | (first_slot + np.arange(cnt, dtype=np.int64)))
        so += cnt
    if not rows:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    rows, dsts = np.concatenate(rows), np.concatenate(dsts)
    o = np.argsort(rows, kind="stable")
    return rows[o], dsts[o]


class RowPartitionedSolver:
    """All ranks construct it with the same mesh, then call :meth:`eigs_smallest` collectively.

    ``order`` maps the solver's row numbering back to the caller's vertex ids: row ``row_begin + i`` of this rank is
    vertex ``order[row_begin + i]``; :meth:`gather_vectors` returns the full eigenvectors in the caller's order."""

    def __init__(self, points, tris, reorder="morton"):
        torch = _lib.require_cuda()
        from ._device import DeviceGraph

        self.rank, self.local_rank, self.world = fdist.world()
        lib = _lib.load()
        # own NCCL communicator of the library: rank 0 draws the id, torch.distributed hands it round
        uid = (C.c_char * 128)()
        if self.rank == 0:
            _lib.call("focusr_dist_unique_id", C.cast(uid, C.c_void_p))
        ids = fdist.gather_objects(bytes(uid.raw))
        uid_all = (C.c_char * 128).from_buffer_copy(ids[0])
        if self.world > 1:
            _lib.call("focusr_dist_init", C.cast(uid_all, C.c_void_p), self.rank, self.world)
        part = partition_mesh(points, tris, self.world, self.rank, reorder)
        self.order, self.bounds = part["order"], part["bounds"]
        self.n_global = int(self.bounds[-1])
        n_loc, ghosts = part["n_local"], part["ghosts"]
        n_tot = n_loc + int(ghosts.size)
        # K1 on the rank's sub-mesh: "mesh" 0 = its own rows (complete), "mesh" 1 = the ghost vertices (rows incomplete,
        # dropped); the per-mesh facts of mesh 0 are this rank's share of the global ones
        off = np.array([0, n_loc, n_tot] if ghosts.size else [0, n_loc], dtype=np.int32)
        sub = DeviceGraph.from_device(torch.from_numpy(part["points_local"]), torch.from_numpy(part["tris_local"]), off)
        own = sub.mesh_info_host[0]
        one_way = int(round(fdist.all_reduce_sum(int(own[1]))))
        if int(round(fdist.all_reduce_sum(int(own[3])))) != 0:
            raise ValueError("the mesh has zero-length edges (non-finite weights, reference graph.py:177-178)")
        if one_way != 0:
            raise NotImplementedError("the row-partitioned solve needs a symmetric adjacency (closed, oriented mesh); "
                                      "%d one-way entries" % one_way)
        self.n_zero_rows = int(round(fdist.all_reduce_sum(int(own[2]))))
        self.max_row_entries = int(own[4])
        self.nnz_local = int(own[0])
        dev = sub.device
        self.device = dev
        self.n_local, self.n_ghost, self.row_begin = n_loc, int(ghosts.size), part["row_begin"]
        self.row_ptr = sub.row_ptr[: n_loc + 1].clone()
        self.cols_local = sub.cols[: max(self.nnz_local, 1)].clone()
        self.weights = sub.weights[: max(self.nnz_local, 1)].clone()
        self.degree = sub.degree[:n_loc].clone()
        self.degree_inv = sub.degree_inv[:n_loc].clone()
        self.points = sub.points[:n_loc].clone()
        del sub
        # NCCL send/receive mode: what this rank ships to whom
        all_ghosts = fdist.gather_objects(ghosts)
        send_idx, send_counts = send_lists(all_ghosts, self.bounds, self.rank)
        self.send_idx = torch.from_numpy(send_idx if send_idx.size else np.zeros(1, np.int32)).to(dev)
        self.n_send = int(send_idx.size)
        self.send_counts = np.ascontiguousarray(send_counts, dtype=np.int32)
        self.recv_counts = np.ascontiguousarray(part["recv_counts"], dtype=np.int32)
        # P2P mode: where every ghost row lives (owner rank, row inside the owner's block)
        owner = np.searchsorted(self.bounds[1:], ghosts, side="right").astype(np.int32)
        g_row = (ghosts - self.bounds[owner]).astype(np.int32)
        self.ghost_peer = torch.from_numpy(owner if owner.size else np.zeros(1, np.int32)).to(dev)
        self.ghost_row = torch.from_numpy(g_row if g_row.size else np.zeros(1, np.int32)).to(dev)
        # persistent-kernel halo: the fp32 views carry ghost rows behind the owned rows (same layout on every rank), and
        # every rank pushes the rows its peers gather into their ghost slots.  Rank q's ghost list is sorted by global id,
        # so this rank's rows sit in it contiguously, in the order of its own send list, after the ghosts q gets from
        # lower ranks.
        self.ghost_base = int(np.max(np.diff(self.bounds)))
        self.rows_cap = self.ghost_base + max(int(gh.size) for gh in all_ghosts)
        rows, dsts = push_lists(send_idx, send_counts, fdist.gather_objects(self.recv_counts), self.rank)
        self.n_push = int(rows.size)
        self.push_row = torch.from_numpy((rows if rows.size else np.zeros(1)).astype(np.int32)).to(dev)
        self.push_dst = torch.from_numpy((dsts if dsts.size else np.zeros(1)).astype(np.int32)).to(dev)
        self._shared_for_block = None
        self._lib = lib

    def _ensure_shared(self, block):
        """IPC-shared region holding this rank's three vector blocks + barrier flags (P2P mode)."""
        if self._shared_for_block == block or self.world == 1:
            return
        lib = self._lib
        nbytes = int(lib.focusr_dist_shared_bytes(self.rows_cap, block, self.world))
        handle = (C.c_char * 64)()
        _lib.call("focusr_dist_shared_alloc", nbytes, C.cast(handle, C.c_void_p))
        handles = b"".join(fdist.gather_objects(bytes(handle.raw)))
        buf = (C.c_char * (64 * self.world)).from_buffer_copy(handles)
        _lib.call("focusr_dist_shared_open", C.cast(buf, C.c_void_p), self.rank, self.world)
        fdist.barrier()  # every rank has zeroed its flags and mapped its peers before anyone signals
        self._shared_for_block = block

    def eigs_smallest(self, k, n_k_needed, k_buffer=1, min_eig_val=1e-10, tol=1e-10, max_outer=60, block_size=0,
                      p2p=True, options=None):
        """``p2p=True`` (default when world > 1): halo fused into the kernels over NVLink peer memory, filter passes in
        the fp32 forms with all steps of a pass in one persistent kernel per GPU; ``p2p=False``: fp64 steps with the
        packed boundary rows exchanged by grouped ncclSend/ncclRecv.  Returns ``(vals, vecs, info)``: the eigenvalues
        (identical on every rank) and this rank's rows of the eigenvectors."""
        torch = _lib.require_cuda()
        lib = self._lib
        b = int(block_size) if block_size else int(lib.focusr_eigs_block_size(k, n_k_needed, k_buffer, 0, self.n_zero_rows))
        use_p2p = bool(p2p) and self.world > 1
        if use_p2p:
            self._ensure_shared(b)
        if isinstance(options, dict):
            options = _lib.EigsOptions(**options)
        opt_ptr = C.byref(options) if options is not None else None
        ldv = b
        vals = torch.zeros(ldv, dtype=torch.float64, device=self.device)
        vecs = torch.zeros((self.n_local, ldv), dtype=torch.float64, device=self.device)
        ws_bytes = int(lib.focusr_eigs_dist_workspace_bytes(self.n_local, self.n_ghost, self.n_send, self.max_row_entries,
                                                            b, self.world))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        res_i, res_d = np.zeros(8, dtype=np.int32), np.zeros(2)
        _lib.call("focusr_eigs_smallest_dist", _lib.ptr(self.row_ptr), _lib.ptr(self.cols_local), _lib.ptr(self.weights),
                  _lib.ptr(self.degree), _lib.ptr(self.degree_inv), _lib.ptr(self.points), self.n_local, self.n_ghost,
                  self.row_begin, self.nnz_local, _lib.ptr(self.send_idx), self.n_send, _lib.ptr(self.send_counts),
                  _lib.ptr(self.recv_counts), _lib.ptr(self.ghost_peer), _lib.ptr(self.ghost_row), self.ghost_base,
                  _lib.ptr(self.push_row), _lib.ptr(self.push_dst), self.n_push, int(use_p2p),
                  self.rows_cap, self.n_zero_rows, self.max_row_entries, int(k), int(n_k_needed), int(k_buffer),
                  float(min_eig_val), float(tol), int(max_outer), b, 0.0, _lib.ptr(vals), _lib.ptr(vecs), ldv,
                  _lib.ptr(res_i), _lib.ptr(res_d), _lib.ptr(ws), ws_bytes, opt_ptr, _lib.stream_ptr())
        m = int(res_i[1])
        info = dict(status=int(res_i[0]), n_found=m, k_final=int(res_i[2]), outer_iterations=int(res_i[3]),
                    filter_degree=int(res_i[4]), block_size=b, fp32_filter_degree=int(res_i[6]), world=int(res_i[7]),
                    max_residual=float(res_d[0]), spectrum_bound=float(res_d[1]),
                    n_local=self.n_local, n_ghost=self.n_ghost, n_send=self.n_send, p2p=use_p2p)
        return vals[:m], vecs[:, :m], info

    def gather_vectors(self, vecs):
        """Every rank's rows -> the full eigenvectors [n_global][m] in the CALLER's vertex order, on every rank (a
        check / export helper, not part of the solve)."""
        torch = _lib.require_cuda()
        if self.world > 1:
            import torch.distributed as dist

            sizes = [int(b1 - b0) for b0, b1 in zip(self.bounds[:-1], self.bounds[1:])]
            parts = [torch.zeros((s, vecs.shape[1]), dtype=vecs.dtype, device=vecs.device) for s in sizes]
            dist.all_gather(parts, vecs.contiguous())
            new = torch.cat(parts)
        else:
            new = vecs
        out = torch.empty_like(new)
        out[torch.from_numpy(self.order).to(new.device)] = new
        return out

    def close(self):
        if self.world > 1:
            fdist.barrier()
            _lib.call("focusr_dist_shared_free")
            _lib.call("focusr_dist_finalize")
