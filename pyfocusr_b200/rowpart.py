"""Row-partitioned eigensolve of ONE large mesh across the GPUs of a node
(BASELINE.json configs[3]: "synthetic 1M-vertex icosphere ... row-partitioned ... NCCL halo over
NVLink").

Rank r owns a contiguous block of rows of the adjacency.  The host-side partition logic below is
pure numpy (tested on CPU, world sizes 1..8, by emulating the exchange); the solve itself is
``focusr_eigs_smallest_dist``: the same Chebyshev-filtered subspace iteration as the single-GPU
path, with one grouped ncclSend/ncclRecv of boundary rows before every SpMM and small all-reduces
for the Gram blocks, residuals and norms.  ``torch.distributed`` is used only to agree on the NCCL
unique id and to exchange the ghost lists once.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import dist as fdist

__all__ = ["row_bounds", "local_partition", "send_lists", "RowPartitionedSolver"]


def row_bounds(n_rows, world):
    """Contiguous, nearly equal row blocks: bounds[r] .. bounds[r+1]."""
    base, rem = divmod(int(n_rows), int(world))
    sizes = np.full(world, base, dtype=np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def local_partition(indptr, indices, bounds, rank):
    """Local CSR of rank's rows with columns remapped to [local | ghosts] numbering.

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


class RowPartitionedSolver:
    """All ranks construct it with the same mesh, then call :meth:`eigs_smallest` collectively."""

    def __init__(self, points, tris):
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
        # every rank assembles the whole adjacency on its own GPU (K1 is cheap) and keeps its rows
        full = DeviceGraph([points], [tris])
        info = full.mesh_info_host[0]
        if info[1] != 0:
            raise NotImplementedError("the row-partitioned solve needs a symmetric adjacency (closed, oriented mesh)")
        self.n_global = full.n_points
        self.n_zero_rows = int(info[2])
        indptr, indices, weights = full.adjacency_host()
        self.bounds = row_bounds(self.n_global, self.world)
        part = local_partition(indptr, indices, self.bounds, self.rank)
        all_ghosts = fdist.gather_objects(part["ghosts"])
        send_idx, send_counts = send_lists(all_ghosts, self.bounds, self.rank)
        r0, n_loc = part["row_begin"], part["n_local"]
        e0, e1 = part["entry_slice"]
        dev = full.device
        self.device = dev
        self.n_local, self.n_ghost, self.row_begin = n_loc, int(part["ghosts"].size), r0
        self.nnz_local = e1 - e0
        self.row_ptr = torch.from_numpy(part["row_ptr"]).to(dev)
        self.cols_local = torch.from_numpy(part["cols_local"]).to(dev)
        self.weights = full.weights[e0:e1].clone()
        self.degree = full.degree[r0 : r0 + n_loc].clone()
        self.degree_inv = full.degree_inv[r0 : r0 + n_loc].clone()
        self.points = full.points[r0 : r0 + n_loc].clone()
        self.send_idx = torch.from_numpy(send_idx if send_idx.size else np.zeros(1, np.int32)).to(dev)
        self.n_send = int(send_idx.size)
        self.send_counts = np.ascontiguousarray(send_counts, dtype=np.int32)
        self.recv_counts = np.ascontiguousarray(part["recv_counts"], dtype=np.int32)
        # P2P mode: where every ghost row lives (owner rank, row inside the owner's block)
        ghosts = part["ghosts"]
        owner = np.searchsorted(self.bounds[1:], ghosts, side="right").astype(np.int32)
        g_row = (ghosts - self.bounds[owner]).astype(np.int32)
        self.ghost_peer = torch.from_numpy(owner if owner.size else np.zeros(1, np.int32)).to(dev)
        self.ghost_row = torch.from_numpy(g_row if g_row.size else np.zeros(1, np.int32)).to(dev)
        self.rows_cap = int(np.max(np.diff(self.bounds)))
        self._shared_for_block = None
        del full
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
                      p2p=True):
        """``p2p=True`` (default when world > 1): halo fused into the SpMM over NVLink peer memory;
        ``p2p=False``: packed boundary rows exchanged with grouped ncclSend/ncclRecv."""
        torch = _lib.require_cuda()
        lib = self._lib
        b = int(block_size) if block_size else int(lib.focusr_eigs_block_size(k, n_k_needed, k_buffer, 0, self.n_zero_rows))
        use_p2p = bool(p2p) and self.world > 1
        if use_p2p:
            self._ensure_shared(b)
        ldv = b
        vals = torch.zeros(ldv, dtype=torch.float64, device=self.device)
        vecs = torch.zeros((self.n_local, ldv), dtype=torch.float64, device=self.device)
        ws_bytes = int(lib.focusr_eigs_dist_workspace_bytes(self.n_local, self.n_ghost, self.n_send, b, self.world))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        res_i, res_d = np.zeros(8, dtype=np.int32), np.zeros(2)
        _lib.call("focusr_eigs_smallest_dist", _lib.ptr(self.row_ptr), _lib.ptr(self.cols_local), _lib.ptr(self.weights),
                  _lib.ptr(self.degree), _lib.ptr(self.degree_inv), _lib.ptr(self.points), self.n_local, self.n_ghost,
                  self.row_begin, self.nnz_local, _lib.ptr(self.send_idx), self.n_send, _lib.ptr(self.send_counts),
                  _lib.ptr(self.recv_counts), _lib.ptr(self.ghost_peer), _lib.ptr(self.ghost_row), int(use_p2p),
                  self.rows_cap, self.n_zero_rows, int(k), int(n_k_needed), int(k_buffer),
                  float(min_eig_val), float(tol), int(max_outer), b, 0.0, _lib.ptr(vals), _lib.ptr(vecs), ldv,
                  _lib.ptr(res_i), _lib.ptr(res_d), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        m = int(res_i[1])
        info = dict(status=int(res_i[0]), n_found=m, k_final=int(res_i[2]), outer_iterations=int(res_i[3]),
                    filter_degree=int(res_i[4]), block_size=b, world=int(res_i[7]), max_residual=float(res_d[0]), spectrum_bound=float(res_d[1]),
                    n_local=self.n_local, n_ghost=self.n_ghost, n_send=self.n_send, p2p=use_p2p)
        return vals[:m], vecs[:, :m], info

    def close(self):
        if self.world > 1:
            fdist.barrier()
            _lib.call("focusr_dist_shared_free")
            _lib.call("focusr_dist_finalize")
