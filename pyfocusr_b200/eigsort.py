"""Drop-in for ``pyfocusr.eigsort`` (reference ``pyfocusr/eigsort.py:9-249``).

The N-sized work -- sampling, the 3n sorted log-columns and 2 n^2 Wasserstein distances of
``calc_c_hist`` (eigsort.py:162-189), the nearest-neighbour search and 2 n^2 RMS differences of
``calc_c_spatial`` (eigsort.py:191-233), and the flip + column reorder of ``eigen_sort``
(eigsort.py:108-122) -- runs in libfocusr_b200.so.  What stays on the host is the n x n
arithmetic the reference also does in Python: ``c_lambda`` (eigsort.py:142-160), ``min(c, c_f)``,
scipy's ``linear_sum_assignment`` and the flip list (eigsort.py:66-105), with n <= ~70.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment

from . import _lib
from ._device import DeviceGraph, eigsort_costs

__all__ = ["eigsort", "decide_matches", "moves_from_matches"]


def c_lambda_matrix(eig_vals_t, eig_vals_s, n):
    """eigsort.py:142-160: the gap averages over ALL eigenvalues each graph returned."""
    gap = (np.mean(np.diff(eig_vals_t)) + np.mean(np.diff(eig_vals_s))) / 2
    lt = np.asarray(eig_vals_t)[:n, None]
    ls = np.asarray(eig_vals_s)[None, :n]
    # same elementwise operations as the reference's double loop (bitwise equal on the golden vectors)
    return np.exp((lt - ls) ** 2 / (2 * gap**2))


def decide_matches(c_lambda, c_hist, c_hist_f, c_spatial, c_spatial_f, target_as_reference=True):
    """eigsort.py:66-105 verbatim: returns (Q per pair, target_matches, source_matches, flipped_pairs)."""
    c = c_spatial * c_lambda * c_hist
    c_f = c_spatial_f * c_lambda * c_hist_f
    q = np.min((c, c_f), axis=0)
    s = c > c_f
    (target_flipped, source_flipped) = np.where(s == True)  # noqa: E712
    if target_as_reference is True:
        target_matches, source_matches = linear_sum_assignment(q)
    else:
        source_matches, target_matches = linear_sum_assignment(q.T)
    q = q[target_matches, source_matches]
    flipped_pairs = [
        p2 for p1 in zip(target_flipped, source_flipped) for p2 in zip(target_matches, source_matches) if p2 == p1
    ]
    return q, target_matches, source_matches, flipped_pairs


def moves_from_matches(target_matches, source_matches, flipped_pairs, target_as_reference=True):
    """Column moves (dst, src, sign) of eigsort.py:108-122 for the graph that gets permuted:
    the source when the target is the reference (new[:, t] = +-old[:, s]), else the target."""
    flipped = set((int(a), int(b)) for a, b in flipped_pairs)
    dst, src, sign = [], [], []
    for t, s in zip(target_matches, source_matches):
        t, s = int(t), int(s)
        if target_as_reference:
            dst.append(t)
            src.append(s)
        else:
            dst.append(s)
            src.append(t)
        sign.append(-1 if (t, s) in flipped else 1)
    return np.array(dst, np.int32), np.array(src, np.int32), np.array(sign, np.int32)


_PERMS = {}


def _perms(n):
    if n not in _PERMS:
        import itertools

        _PERMS[n] = np.array(list(itertools.permutations(range(n))), dtype=np.int64)
    return _PERMS[n]


def decide_batch(c_lambda, c_hist, c_hist_f, c_spatial, c_spatial_f, target_as_reference=True):
    """``decide_matches`` + ``moves_from_matches`` for P pairs at once (inputs [P][n][n]).

    For n <= 8 the assignment is found by scoring all n! permutations in one vectorised pass -- the
    same optimum scipy's ``linear_sum_assignment`` returns whenever it is unique (ties have measure
    zero for these costs; tests/test_host_logic.py checks the two agree) -- otherwise scipy is called
    per pair.  Returns (Q [P][n], dst, src, sign int32 [P][n])."""
    c = c_spatial * c_lambda * c_hist
    c_f = c_spatial_f * c_lambda * c_hist_f
    q = np.minimum(c, c_f)  # == np.min((c, c_f), axis=0)
    s = c > c_f
    P, n, _ = q.shape
    rows = np.arange(n)
    if n <= 8:
        perms = _perms(n)
        qq = q if target_as_reference else np.swapaxes(q, 1, 2)
        # sum in row order exactly like the per-pair total scipy minimises
        tot = np.zeros((P, perms.shape[0]))
        for i in range(n):
            tot = tot + qq[:, i, perms[:, i]]
        assign = perms[np.argmin(tot, axis=1)]          # [P][n]: column matched to row i
    else:
        assign = np.empty((P, n), dtype=np.int64)
        for p in range(P):
            assign[p] = linear_sum_assignment(q[p] if target_as_reference else q[p].T)[1]
    if target_as_reference:
        tm, sm = np.broadcast_to(rows, (P, n)), assign   # target i <-> source assign[i]
    else:
        sm, tm = np.broadcast_to(rows, (P, n)), assign   # source i <-> target assign[i]
    pi = np.arange(P)[:, None]
    q_pairs = q[pi, tm, sm]
    sign = np.where(s[pi, tm, sm], -1, 1).astype(np.int32)
    dst, src = (tm, sm) if target_as_reference else (sm, tm)
    return q_pairs, np.ascontiguousarray(dst, dtype=np.int32), np.ascontiguousarray(src, dtype=np.int32), sign


class eigsort(object):
    def __init__(self, graph_target, graph_source, n_features, target_as_reference=True):
        self.graph_target = graph_target
        self.graph_source = graph_source
        self.n_features = n_features
        self.target_as_reference = target_as_reference
        # eigsort.py:34-41 (kept as host attributes for API parity; the kernels re-gather on device)
        self.rand_target_points = self.graph_target.get_rand_normalized_points()
        self.rand_source_points = self.graph_source.get_rand_normalized_points()
        self.rand_target_eig_vecs = self.graph_target.get_rand_eig_vecs()
        self.rand_source_eig_vecs = self.graph_source.get_rand_eig_vecs()
        self.c_lambda = np.zeros((self.n_features, self.n_features))
        self.c_hist = np.zeros_like(self.c_lambda)
        self.c_hist_f = np.zeros_like(self.c_lambda)
        self.c_spatial = np.zeros_like(self.c_lambda)
        self.c_spatial_f = np.zeros_like(self.c_lambda)
        self.Q = None
        self._costs_done = False

    # --- device side ----------------------------------------------------------------------------
    def _pair_graph(self):
        gt, gs = self.graph_target, self.graph_source
        torch = _lib.require_cuda()
        g = DeviceGraph([gt.points, gs.points], [gt._tris, gs._tris])
        ld = max(gt.eig_vecs.shape[1], gs.eig_vecs.shape[1])
        host = np.zeros((gt.n_points + gs.n_points, ld))
        host[: gt.n_points, : gt.eig_vecs.shape[1]] = gt.eig_vecs
        host[gt.n_points :, : gs.eig_vecs.shape[1]] = gs.eig_vecs
        return g, torch.from_numpy(host).to(g.device)

    def _device_costs(self):
        if self._costs_done:
            return
        g, vecs = self._pair_graph()
        ch, chf, cs, csf, _ = eigsort_costs(g, vecs, [0], [1], np.asarray(self.graph_target.rand_idxs)[None, :],
                                            np.asarray(self.graph_source.rand_idxs)[None, :], self.n_features)
        self.c_hist, self.c_hist_f = ch[0].cpu().numpy(), chf[0].cpu().numpy()
        self.c_spatial, self.c_spatial_f = cs[0].cpu().numpy(), csf[0].cpu().numpy()
        self._dev_pair = (g, vecs)
        self._costs_done = True

    # eigsort.py:142-160
    def calc_c_lambda(self):
        for graph in [self.graph_source, self.graph_target]:
            if graph.eig_val_gap is None:
                graph.get_eig_val_gap()
        eigen_gap = (self.graph_target.eig_val_gap + self.graph_source.eig_val_gap) / 2
        for i in range(self.n_features):
            for j in range(self.n_features):
                self.c_lambda[i, j] = np.exp(
                    (self.graph_target.eig_vals[i] - self.graph_source.eig_vals[j]) ** 2 / (2 * eigen_gap**2)
                )

    # eigsort.py:162-189
    def calc_c_hist(self):
        self._device_costs()

    # eigsort.py:191-233
    def calc_c_spatial(self):
        self._device_costs()

    # eigsort.py:54-140
    def eigen_sort(self):
        self.Q, target_matches, source_matches, flipped_pairs = decide_matches(
            self.c_lambda, self.c_hist, self.c_hist_f, self.c_spatial, self.c_spatial_f, self.target_as_reference
        )
        dst, src, sign = moves_from_matches(target_matches, source_matches, flipped_pairs, self.target_as_reference)
        g, vecs = self._dev_pair if self._costs_done else self._pair_graph()
        n = len(dst)
        ident = (np.arange(n, dtype=np.int32), np.arange(n, dtype=np.int32), np.ones(n, np.int32))
        moves = (dst, src, sign)
        per_mesh = [ident, moves] if self.target_as_reference else [moves, ident]
        g.flip_permute(vecs, np.stack([m[0] for m in per_mesh]), np.stack([m[1] for m in per_mesh]),
                       np.stack([m[2] for m in per_mesh]))
        out = vecs.cpu().numpy()
        nt = self.graph_target.n_points
        if self.target_as_reference:
            self.graph_source.eig_vecs[:, :] = out[nt:, : self.graph_source.eig_vecs.shape[1]]
        else:
            self.graph_target.eig_vecs[:, :] = out[:nt, : self.graph_target.eig_vecs.shape[1]]
        self.target_matches, self.source_matches, self.flipped_pairs = target_matches, source_matches, flipped_pairs

    # eigsort.py:235-249
    def sort_eigenmaps(self):
        self.calc_c_lambda()
        self.calc_c_hist()
        self.calc_c_spatial()
        self.eigen_sort()
        return self.Q
