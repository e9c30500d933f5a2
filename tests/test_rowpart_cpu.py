"""Host logic of the row-partitioned solve (pyfocusr_b200/rowpart.py): partition, ghost lists, send lists.
The exchange is emulated with numpy for world sizes 1..8 and the partitioned SpMM is compared with scipy."""
import numpy as np
import pytest

from oracle import port
from pyfocusr_b200 import rowpart
from pyfocusr_b200.mesh import icosphere, perturbed_ellipsoid


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partitioned_spmm_equals_global(world):
    m = perturbed_ellipsoid(9, 2)  # 812 vertices
    a = port.adjacency(m.points, m.tris)
    n = a.shape[0]
    bounds = rowpart.row_bounds(n, world)
    assert bounds[0] == 0 and bounds[-1] == n and np.all(np.diff(bounds) >= n // world)
    parts = [rowpart.local_partition(a.indptr, a.indices, bounds, r) for r in range(world)]
    sends = [rowpart.send_lists([p["ghosts"] for p in parts], bounds, r) for r in range(world)]
    x = np.random.RandomState(0).standard_normal((n, 4))
    y_ref = a @ x
    for r in range(world):
        p = parts[r]
        r0, r1 = bounds[r], bounds[r + 1]
        # emulate the halo exchange: peer q ships x[local send rows] and r stores them behind its local rows
        ext = np.zeros((p["n_local"] + p["ghosts"].size, 4))
        ext[: p["n_local"]] = x[r0:r1]
        ro = 0
        for q in range(world):
            idx_q, cnt_q = sends[q]
            so = int(np.sum(cnt_q[:r]))
            rows = idx_q[so : so + cnt_q[r]] + bounds[q]
            assert cnt_q[r] == p["recv_counts"][q]
            ext[p["n_local"] + ro : p["n_local"] + ro + cnt_q[r]] = x[rows]
            assert np.array_equal(rows, p["ghosts"][ro : ro + cnt_q[r]])
            ro += cnt_q[r]
        assert ro == p["ghosts"].size and sends[r][1][r] == 0
        e0, e1 = p["entry_slice"]
        y = np.zeros((p["n_local"], 4))
        for i in range(p["n_local"]):
            s, e = p["row_ptr"][i], p["row_ptr"][i + 1]
            y[i] = a.data[e0 + s : e0 + e] @ ext[p["cols_local"][s:e]]
        assert np.allclose(y, y_ref[r0:r1], rtol=1e-14, atol=1e-14)


def test_halo_is_thin_for_lattice_ordered_icosphere():
    m = icosphere(40)  # 16 002 vertices, face-major lattice order
    a = port.adjacency(m.points, m.tris)
    bounds = rowpart.row_bounds(a.shape[0], 8)
    for r in range(8):
        p = rowpart.local_partition(a.indptr, a.indices, bounds, r)
        assert p["ghosts"].size < 0.25 * p["n_local"]


@pytest.mark.parametrize("world", [1, 2, 5, 8])
def test_partition_mesh_rows_equal_the_permuted_global_matrix(world):
    """partition_mesh: a rank assembles only the triangles touching its rows, in the numbering [own | ghosts]; the rows
    of its own vertices must equal the rows of the globally assembled, Morton-permuted adjacency (same weights, bit for
    bit: an edge weight depends only on the two endpoints)."""
    m = perturbed_ellipsoid(9, 3)
    rng = np.random.RandomState(1)
    shuffle = rng.permutation(m.points.shape[0])          # the file order carries no locality
    pts = m.points[shuffle]
    inv = np.empty_like(shuffle)
    inv[shuffle] = np.arange(shuffle.size)
    tris = inv[m.tris]
    a = port.adjacency(pts, tris).tocsr()
    seen = np.zeros(pts.shape[0], dtype=int)
    for r in range(world):
        p = rowpart.partition_mesh(pts, tris, world, r)
        order, bounds, n_loc, ghosts = p["order"], p["bounds"], p["n_local"], p["ghosts"]
        r0 = p["row_begin"]
        assert sorted(order.tolist()) == list(range(pts.shape[0]))
        seen[order[r0:r0 + n_loc]] += 1
        glob = np.concatenate([order[r0:r0 + n_loc], order[ghosts]])       # old id of every local vertex
        assert np.array_equal(p["points_local"], pts[glob])
        a_loc = port.adjacency(p["points_local"], p["tris_local"]).tocsr()[:n_loc]
        a_ref = a[order[r0:r0 + n_loc]][:, glob]                            # own rows, local column numbering
        assert a_ref.nnz == a[order[r0:r0 + n_loc]].nnz                     # no neighbour outside own + ghosts
        assert (a_loc != a_ref).nnz == 0
        owner = np.searchsorted(bounds[1:], ghosts, side="right")
        assert np.all(owner != r) and np.array_equal(np.bincount(owner, minlength=world), p["recv_counts"])
        assert np.all(np.diff(ghosts) > 0)
    assert np.all(seen == 1)


def test_morton_order_keeps_halos_thin_on_a_shuffled_icosphere():
    """with a locality ordering the halo stays a few percent of the rows even when the input
    order is random (without it, every row of a shuffled mesh is a ghost of somebody)."""
    m = icosphere(100)                                   # 100 002 vertices
    rng = np.random.RandomState(0)
    shuffle = rng.permutation(m.points.shape[0])
    pts = m.points[shuffle]
    inv = np.empty_like(shuffle)
    inv[shuffle] = np.arange(shuffle.size)
    tris = inv[m.tris]
    for world in (2, 8):
        worst = 0.0
        for r in range(world):
            p = rowpart.partition_mesh(pts, tris, world, r)
            worst = max(worst, p["ghosts"].size / p["n_local"])
        print("world", world, "worst halo fraction", worst)
        assert worst < 0.05, (world, worst)   # 1.3% / 4.7% at 100k vertices (patches of 12.5k rows at world 8)
    p = rowpart.partition_mesh(pts, tris, 8, 3, reorder=None)
    assert p["ghosts"].size > 2 * p["n_local"]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_push_lists_fill_every_ghost_slot_with_the_right_row(world):
    """The halo of the persistent filter kernel is a push: emulate it with numpy and check that every rank's ghost rows
    end up holding exactly the rows its ghost list names."""
    m = perturbed_ellipsoid(9, 4)
    n = m.points.shape[0]
    parts = [rowpart.partition_mesh(m.points, m.tris, world, r) for r in range(world)]
    bounds = parts[0]["bounds"]
    all_ghosts = [p["ghosts"] for p in parts]
    recv_all = [p["recv_counts"] for p in parts]
    x = np.random.RandomState(0).standard_normal(n)            # one value per (new) global row
    ghost_rows = [np.full(p["ghosts"].size, np.nan) for p in parts]
    for r in range(world):
        send_idx, send_counts = rowpart.send_lists(all_ghosts, bounds, r)
        rows, dst = rowpart.push_lists(send_idx, send_counts, recv_all, r)
        assert np.all(np.diff(rows) >= 0) and rows.size == send_idx.size
        for row, d in zip(rows, dst):
            peer, slot = int(d) >> 24, int(d) & 0xFFFFFF
            assert peer != r and np.isnan(ghost_rows[peer][slot])
            ghost_rows[peer][slot] = x[bounds[r] + row]
    for r in range(world):
        assert np.array_equal(ghost_rows[r], x[parts[r]["ghosts"]])
