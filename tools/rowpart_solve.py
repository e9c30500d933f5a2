"""Row-partitioned eigensolve of one icosphere across the visible GPUs (BASELINE.json configs[3]).

  python tools/rowpart_solve.py [nu] [k]                       (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29517 tools/rowpart_solve.py [nu] [k]      (N GPUs, NCCL over NVLink)

nu=316 is the 998 562-vertex mesh; eigenvalues are checked against tests/golden/large_eigs.npz (scipy) when
available, else against a single-GPU solve on rank 0.  Prints one JSON line on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyfocusr_b200 import dist as fdist
from pyfocusr_b200.mesh import icosphere
from pyfocusr_b200.rowpart import RowPartitionedSolver

nu = int(sys.argv[1]) if len(sys.argv) > 1 else 316
k = int(sys.argv[2]) if len(sys.argv) > 2 else 11
p2p = (sys.argv[3] != "nccl") if len(sys.argv) > 3 else True   # halo: fused P2P loads (default) or ncclSend/Recv
rank, local, world = fdist.world()
torch.cuda.set_device(local)
fdist.init("nccl")
m = icosphere(nu)
solver = RowPartitionedSolver(m.points, m.tris)
solver.eigs_smallest(k=k, n_k_needed=k - 1, p2p=p2p)          # warm-up (NCCL channels, kernels)
fdist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
vals, vecs, info = solver.eigs_smallest(k=k, n_k_needed=k - 1, p2p=p2p)
e1.record()
fdist.barrier()
ms = fdist.all_reduce_max(e0.elapsed_time(e1))
v = vals.cpu().numpy()
# residual of this rank's rows needs the neighbours' rows: gather the full vectors on every rank (check only)
full = [None] * world
if world > 1:
    import torch.distributed as dist
    sizes = [int(b1 - b0) for b0, b1 in zip(solver.bounds[:-1], solver.bounds[1:])]
    parts = [torch.zeros((s, vecs.shape[1]), dtype=torch.float64, device="cuda") for s in sizes]
    dist.all_gather(parts, vecs.contiguous())
    allv = torch.cat(parts)
else:
    allv = vecs
ok_ref, rel = None, None
gold_path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "large_eigs.npz")
if nu == 316 and k == 11 and os.path.exists(gold_path):
    gold = np.load(gold_path)["nu316_k11"]
    rel = float(np.max(np.abs(v - gold) / gold))
    ok_ref = rel <= 1e-6
if rank == 0:
    from pyfocusr_b200._device import DeviceGraph
    g = DeviceGraph([m.points], [m.tris])
    nn = v.size
    b = (nn + 7) // 8 * 8
    x = torch.zeros((g.n_points, b), dtype=torch.float64, device="cuda")
    x[:, :nn] = allv
    r = g.laplacian_apply(x)[:, :nn] - x[:, :nn] * vals[None, :]
    res = float(torch.linalg.vector_norm(r, dim=0).max())
    norms = torch.linalg.vector_norm(allv, dim=0).cpu().numpy()
    print(json.dumps({"config": "configs[3]: icosphere nu=%d (%d vertices), k=%d smallest, row-partitioned over %d GPU(s)" % (nu, g.n_points, k, world),
                      "n_gpus": world, "halo": ("p2p-fused" if info["p2p"] else ("nccl send/recv" if world > 1 else "none")), "seconds": ms / 1e3, "status": info["status"], "n_found": info["n_found"],
                      "outer_iterations": info["outer_iterations"], "filter_degree": info["filter_degree"], "block": info["block_size"],
                      "n_local": info["n_local"], "n_ghost": info["n_ghost"], "max_residual_global": res,
                      "max_rel_err_vs_scipy": rel, "within_1e-6": ok_ref, "unit_norm_err": float(np.max(np.abs(norms - 1.0))),
                      "eig_vals": v.tolist()}))
    assert info["status"] == 0 and res <= 1e-9 and (ok_ref is None or ok_ref)
solver.close()
fdist.finalize()
