"""Coherent Point Drift on the GPU behind the ``cycpd`` call surface the reference uses
(focusr.py:297-334): ``affine_registration(X=, Y=, max_iterations=, tolerance=)`` and
``deformable_registration(X=, Y=, num_eig=, max_iterations=, tolerance=, alpha=, beta=)``, each with
``register() -> (TY, params)`` and ``transform_point_cloud(Y)``.

The arithmetic is ``focusr_cpd_*`` in ``csrc/cpd.cu`` (E-step, M-step and the low-rank kernel
eigen-decomposition all on the device).  cycpd itself is not available offline, so parity is against the
restated algorithm in ``oracle/cpd_port.py`` only -- see that file's header.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["affine_registration", "deformable_registration"]


def _dev(a, torch):
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


class _Registration:
    def __init__(self, X, Y, max_iterations=100, tolerance=0.001, w=0.0, verbose=False, **_ignored):
        torch = _lib.require_cuda()
        self._torch = torch
        self.X, self.Y = _dev(X, torch), _dev(Y, torch)
        if self.X.ndim != 2 or self.Y.ndim != 2 or self.X.shape[1] != self.Y.shape[1]:
            raise ValueError("X and Y must be (n, D) arrays with the same D")
        (self.N, self.D), self.M = self.X.shape, self.Y.shape[0]
        self.max_iterations = 100 if max_iterations is None else int(max_iterations)
        self.tolerance = 0.001 if tolerance is None else float(tolerance)
        self.w = 0.0 if w is None else float(w)
        self.verbose = verbose
        self.iteration, self.sigma2, self.diff = 0, None, np.inf
        self.TY = None

    def _workspace(self, num_eig):
        lib = _lib.load()
        nbytes = int(lib.focusr_cpd_workspace_bytes(self.N, self.M, self.D, num_eig))
        return self._torch.empty(nbytes, dtype=self._torch.uint8, device="cuda"), nbytes

    def _as_input_type(self, t, like):
        return t if isinstance(like, self._torch.Tensor) else t.cpu().numpy()


class affine_registration(_Registration):
    """``cycpd.affine_registration`` (reference call: focusr.py:319-331).  ``register()`` returns
    ``(TY, (B, t))`` with ``TY = Y @ B + t``."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.B, self.t, self.q = np.eye(self.D), np.zeros(self.D), np.inf

    def register(self, callback=None):
        torch = self._torch
        ws, nbytes = self._workspace(0)
        self._B = torch.empty((self.D, self.D), dtype=torch.float64, device="cuda")
        self._t = torch.empty(self.D, dtype=torch.float64, device="cuda")
        ty = torch.empty_like(self.Y)
        res = np.zeros(4)
        _lib.call("focusr_cpd_affine", _lib.ptr(self.X), self.N, _lib.ptr(self.Y), self.M, self.D, self.max_iterations,
                  self.tolerance, self.w, _lib.ptr(self._B), _lib.ptr(self._t), _lib.ptr(ty), _lib.ptr(res), _lib.ptr(ws),
                  nbytes, _lib.stream_ptr())
        self.iteration, self.sigma2, self.q, self.diff = int(res[0]), float(res[1]), float(res[2]), float(res[3])
        self.B, self.t = self._B.cpu().numpy(), self._t.cpu().numpy()
        self._TY = ty
        self.TY = ty.cpu().numpy()
        return self.TY, self.get_registration_parameters()

    def get_registration_parameters(self):
        return self.B, self.t

    def transform_point_cloud(self, Y=None):
        torch = self._torch
        if Y is None:
            return None
        pts = _dev(Y, torch)
        out = torch.empty_like(pts)
        _lib.call("focusr_cpd_affine_apply", _lib.ptr(pts), pts.shape[0], self.D, _lib.ptr(self._B), _lib.ptr(self._t),
                  _lib.ptr(out), _lib.stream_ptr())
        return self._as_input_type(out, Y)


class deformable_registration(_Registration):
    """``cycpd.deformable_registration`` (reference call: focusr.py:299-316).  ``register()`` returns
    ``(TY, (G, W))``; ``G`` (M x M) is materialised lazily, only if the caller indexes the tuple."""

    def __init__(self, *args, alpha=2.0, beta=2.0, num_eig=100, **kwargs):
        super().__init__(*args, **kwargs)
        self.alpha = 2.0 if alpha is None else float(alpha)
        self.beta = 2.0 if beta is None else float(beta)
        self.num_eig = int(num_eig)
        self.W = np.zeros((self.M, self.D))
        self.eig_info = None

    def register(self, callback=None):
        torch = self._torch
        ws, nbytes = self._workspace(self.num_eig)
        self._W = torch.empty((self.M, self.D), dtype=torch.float64, device="cuda")
        ty = torch.empty_like(self.Y)
        res = np.zeros(6)
        _lib.call("focusr_cpd_deformable", _lib.ptr(self.X), self.N, _lib.ptr(self.Y), self.M, self.D, self.max_iterations,
                  self.tolerance, self.w, self.alpha, self.beta, self.num_eig, _lib.ptr(self._W), _lib.ptr(ty),
                  _lib.ptr(res), _lib.ptr(ws), nbytes, _lib.stream_ptr())
        self.iteration, self.sigma2, self.diff = int(res[0]), float(res[1]), float(res[2])
        self.eig_info = dict(iterations=int(res[3]), residual=float(res[4]), smallest_kept=float(res[5]))
        self.W = self._W.cpu().numpy()
        self._TY = ty
        self.TY = ty.cpu().numpy()
        return self.TY, self.get_registration_parameters()

    @property
    def G(self):
        """exp(-|y_i - y_j|^2 / (2 beta^2)) on the host (only built when asked for)."""
        g = self._torch.empty((self.M, self.M), dtype=self._torch.float64, device="cuda")
        _lib.call("focusr_cpd_kernel_matrix", _lib.ptr(self.Y), self.M, self.D, self.beta, _lib.ptr(g), _lib.stream_ptr())
        return g.cpu().numpy()

    def get_registration_parameters(self):
        return _LazyParams(self)

    def transform_point_cloud(self, Y=None):
        torch = self._torch
        if Y is None:
            return None
        pts = _dev(Y, torch)
        out = torch.empty_like(pts)
        _lib.call("focusr_cpd_deformable_apply", _lib.ptr(pts), pts.shape[0], _lib.ptr(self.Y), self.M, self.D,
                  _lib.ptr(self._W), self.beta, _lib.ptr(out), _lib.stream_ptr())
        return self._as_input_type(out, Y)


class _LazyParams:
    """Behaves like the tuple ``(G, W)`` but only forms the M x M kernel matrix when item 0 is read."""

    def __init__(self, reg):
        self._reg = reg

    def __len__(self):
        return 2

    def __getitem__(self, i):
        if i in (0, -2):
            return self._reg.G
        if i in (1, -1):
            return self._reg.W
        raise IndexError(i)

    def __iter__(self):
        yield self._reg.G
        yield self._reg.W
