"""Wall-clock of the drop-in single-pair API on the shipped 15k pair (BASELINE.json configs[0]) and the 5k pair
(configs[1], n_spectral_features=10), ICP off / CPD identity, next to the CPU oracle port on the same inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfocusr_b200 as pyfocusr
from pyfocusr_b200.mesh import PolyData
from oracle import port

if "--fp64" in sys.argv:   # A/B: every filter pass in fp64 (the default lets them iterate in fp32)
    from pyfocusr_b200._device import DeviceGraph
    _orig_eigs = DeviceGraph.eigs_smallest
    def _eigs_fp64(self, *a, **k):
        k.setdefault("options", dict(mixed_precision=0))
        return _orig_eigs(self, *a, **k)
    DeviceGraph.eigs_smallest = _eigs_fp64
    print("A/B: options.mixed_precision = 0")

z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "meshes.npz"))
def mesh(n): return PolyData(z[n + "_points"], z[n + "_tris"])
for tag, (tn, sn), kw in (("configs[0] 15k pair, defaults", ("target_mesh_15k", "source_mesh_15k"), {}),
                          ("configs[1] 5k pair, n_spectral_features=10", ("target_mesh", "source_mesh"), dict(n_spectral_features=10))):
    mt, ms = mesh(tn), mesh(sn)
    reps = []
    for rep in range(6):
        np.random.seed(0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        f = pyfocusr.Focusr(mt, ms, icp_register_first=False, list_features_to_calc=[], registration="identity", **kw)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        f.align_maps()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        reps.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    print("   all repetitions, constructor + align_maps [ms]: " + ", ".join("%.1f + %.1f" % r for r in reps), flush=True)
    # where the constructor's time goes (one more run, device synchronised at every boundary)
    from pyfocusr_b200.graph import Graph
    stages, orig = [], Graph.get_graph_spectrum
    def timed_spectrum(self, *a, **k):
        torch.cuda.synchronize(); s0 = time.perf_counter()
        out = orig(self, *a, **k)
        torch.cuda.synchronize(); stages.append((time.perf_counter() - s0) * 1e3)
        return out
    Graph.get_graph_spectrum = timed_spectrum
    np.random.seed(0)
    torch.cuda.synchronize(); c0 = time.perf_counter()
    f = pyfocusr.Focusr(mt, ms, icp_register_first=False, list_features_to_calc=[], registration="identity", **kw)
    torch.cuda.synchronize(); c1 = time.perf_counter()
    Graph.get_graph_spectrum = orig
    print("   constructor %.1f ms, of which Laplacian + eigensolve per graph: %s ms; eigensolver reports: %s" % (
        (c1 - c0) * 1e3, ", ".join("%.1f" % v for v in stages),
        [getattr(g, "eigs_info", None) and {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in g.eigs_info.items()
                                             if k in ("outer_iterations", "filter_degree", "block_size", "k_final")}
         for g in (f.graph_target, f.graph_source)]), flush=True)
    ns = kw.get("n_spectral_features", 3)
    np.random.seed(0)
    t3 = time.perf_counter()
    if "--no-oracle" not in sys.argv:
        port.spectral_stage(mt.points, mt.tris, ms.points, ms.tris, ns, 3, 5000)
    t4 = time.perf_counter()
    print("%s: B200 drop-in Focusr() %.3f s + align_maps() %.3f s = %.3f s;  CPU oracle port %.2f s  (eigenpairs kept: %d / %d)" % (
        tag, t1 - t0, t2 - t1, t2 - t0, t4 - t3, f.graph_target.eig_vals.size, f.graph_source.eig_vals.size), flush=True)
