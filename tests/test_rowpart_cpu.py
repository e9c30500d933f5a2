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
