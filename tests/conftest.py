import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "reference_outputs.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def shipped_meshes():
    """The four meshes shipped with the reference (data/*.vtk), stored as arrays by oracle/make_golden.py."""
    from pyfocusr_b200.mesh import PolyData

    z = np.load(os.path.join(GOLDEN, "meshes.npz"))
    names = ["target_mesh", "source_mesh", "target_mesh_15k", "source_mesh_15k"]
    return {n: PolyData(z[n + "_points"], z[n + "_tris"], {"thickness_change_(mm)": z[n + "_scalar"]}) for n in names}


@pytest.fixture(scope="session")
def hostsim():
    """The host test double of the solver backend (tests/hostsim), built on demand."""
    import ctypes as C
    import subprocess

    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libhostsim.so")
    src_m = max(os.path.getmtime(os.path.join(d, "hostsim.cpp")),
                *(os.path.getmtime(os.path.join(ROOT, "pyfocusr_b200", "csrc", f))
                  for f in ("chfsi_driver.hpp", "dense_small.h", "nonsym_host.hpp", "nonsym_small.h", "rowops.h", "eigsort_decide.h")))
    if not os.path.exists(so) or os.path.getmtime(so) < src_m:
        subprocess.check_call(["sh", os.path.join(d, "build.sh")])
    lib = C.CDLL(so)
    dp = np.ctypeslib.ndpointer(np.float64, flags="C")
    ip = np.ctypeslib.ndpointer(np.int32, flags="C")
    lib.hostsim_eigs.argtypes = [ip, ip, dp, dp, dp, dp, C.c_int, ip, C.c_int, C.c_int, ip, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int,
                                 dp, dp, ip, dp]
    lib.hostsim_set_lowp.argtypes = [C.c_double]
    lib.hostsim_set_lowp.restype = None
    lib.hostsim_rr_sym.argtypes = [dp, dp, dp, dp, C.c_int]
    lib.hostsim_rr_sym_rev.argtypes = [dp, dp, dp, dp, C.c_int]
    lib.hostsim_eig_general.argtypes = [dp, C.c_int, dp, dp]
    lib.hostsim_rr_nonsym.argtypes = [C.c_int, dp, dp, C.c_int, C.c_double, dp, dp, ip]
    lib.hostsim_set_nonsym_small.argtypes = [C.c_int]
    lib.hostsim_edge_weight.argtypes = [dp, dp, C.c_int]
    lib.hostsim_edge_weight.restype = C.c_double
    lib.hostsim_lsap.argtypes = [dp, C.c_int, ip]
    lib.hostsim_eigsort_decide.argtypes = [dp, C.c_int, dp, C.c_int, dp, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                           dp, ip, ip, ip, dp]
    return lib
