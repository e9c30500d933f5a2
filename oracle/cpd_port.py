"""CPU oracle for the Coherent Point Drift stage  --  TEST INFRASTRUCTURE ONLY.  ** PARITY UNPINNED **

The reference registers the spectral coordinates with ``cycpd.affine_registration`` followed by
``cycpd.deformable_registration`` (focusr.py:297-334: random subsets of at most 5000 points as X
(source) and Y (target), then ``reg.transform_point_cloud`` on ALL target coordinates).  ``cycpd``
(github.com/gattia/cycpd, a Cython fork of pycpd by the reference's author) is an UNPINNED dependency
(requirements.txt:8) that is not installed in this image, not in the wheelhouse and not vendored under
``/root/reference``; neither is pycpd.  Nothing here could therefore be checked against the dependency
itself or against reference outputs: this file restates the PUBLISHED algorithm

    A. Myronenko, X. Song, "Point Set Registration: Coherent Point Drift", IEEE TPAMI 32(12), 2010
    (affine: Fig. 3; non-rigid: Fig. 4; low-rank acceleration: section 7 "fast implementation")

in the formulation of pycpd / cycpd that the reference's call sites imply -- the keyword names the
reference passes (``X, Y, max_iterations, tolerance, alpha, beta, num_eig``), ``register()`` returning
``(TY, params)`` and ``transform_point_cloud(Y)`` -- with these stated conventions:

  * sigma2 is initialised to the mean squared distance over all pairs / D; the outlier weight w is 0;
  * the E-step uses direct squared differences; a zero denominator is clipped to machine epsilon;
  * affine: objective q = (xPx - 2 tr(AB) + tr(B YPY B)) / (2 sigma2) + D Np / 2 log sigma2, the loop ends
    when |q - q_prev| <= tolerance or at max_iterations; sigma2 <= 0 is replaced by tolerance / 10;
  * deformable: G = exp(-|y_i - y_j|^2 / (2 beta^2)); because the reference passes ``num_eig`` the low-rank
    path is taken: the ``num_eig`` eigenpairs of G of largest magnitude (dense symmetric eigensolver),
    W by the Woodbury identity, TY = Y + Q S Q^T W; the loop ends when |sigma2 - sigma2_prev| <= tolerance;
  * ``transform_point_cloud(Y_new)`` = Y_new + G(Y_new, Y) W  (deformable),  Y_new B + t  (affine).

The CUDA implementation (csrc/cpd.cu) is compared with THIS file only; DESIGN.md says the same.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np


def initialize_sigma2(X, Y):
    """Mean squared pair distance / D (CPD paper eq. after (6); pycpd ``initialize_sigma2``)."""
    (N, D), M = X.shape, Y.shape[0]
    diff = X[None, :, :] - Y[:, None, :]
    return float(np.sum(diff * diff) / (D * M * N))


def gaussian_kernel(X, beta, Y=None):
    Y = X if Y is None else Y
    diff = X[:, None, :] - Y[None, :, :]
    return np.exp(-np.sum(diff * diff, axis=2) / (2.0 * beta ** 2))


def low_rank_eigen(G, num_eig):
    """The ``num_eig`` eigenpairs of largest |eigenvalue| (CPD paper section 7)."""
    S, Q = np.linalg.eigh(G)
    idx = np.argsort(np.abs(S))[::-1][:num_eig]
    return Q[:, idx], S[idx]


class _EM:
    def __init__(self, X, Y, sigma2=None, max_iterations=100, tolerance=1e-3, w=0.0):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.Y = np.ascontiguousarray(Y, dtype=np.float64)
        self.TY = self.Y.copy()
        (self.N, self.D), self.M = self.X.shape, self.Y.shape[0]
        self.sigma2 = initialize_sigma2(self.X, self.Y) if sigma2 is None else float(sigma2)
        self.max_iterations, self.tolerance, self.w = int(max_iterations), float(tolerance), float(w)
        self.iteration, self.diff, self.q = 0, np.inf, np.inf

    def expectation(self):
        d2 = np.sum((self.X[None, :, :] - self.TY[:, None, :]) ** 2, axis=2)       # (M, N)
        P = np.exp(-d2 / (2.0 * self.sigma2))
        c = (2.0 * np.pi * self.sigma2) ** (self.D / 2.0) * self.w / (1.0 - self.w) * self.M / self.N
        den = np.clip(np.sum(P, axis=0, keepdims=True), np.finfo(np.float64).eps, None) + c
        P = P / den
        self.Pt1, self.P1 = np.sum(P, axis=0), np.sum(P, axis=1)
        self.Np = float(np.sum(self.P1))
        self.PX = P @ self.X
        self.P = P

    def iterate(self):
        self.expectation()
        self.update_transform()
        self.transform_point_cloud()
        self.update_variance()
        self.iteration += 1

    def register(self):
        self.transform_point_cloud()
        while self.iteration < self.max_iterations and self.diff > self.tolerance:
            self.iterate()
        return self.TY, self.get_registration_parameters()


class AffineRegistration(_EM):
    """CPD paper Fig. 3 (affine).  Reference call site: focusr.py:319-331."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.B, self.t = np.eye(self.D), np.zeros(self.D)

    def update_transform(self):
        muX = np.sum(self.PX, axis=0) / self.Np
        muY = (self.P1 @ self.Y) / self.Np
        self.X_hat, Y_hat = self.X - muX, self.Y - muY
        self.A = (self.P @ self.X_hat).T @ Y_hat
        self.YPY = (Y_hat.T * self.P1) @ Y_hat
        self.B = np.linalg.solve(self.YPY.T, self.A.T)
        self.t = muX - self.B.T @ muY

    def transform_point_cloud(self, Y=None):
        if Y is None:
            self.TY = self.Y @ self.B + self.t
            return None
        return np.asarray(Y, dtype=np.float64) @ self.B + self.t

    def update_variance(self):
        qprev = self.q
        trAB = np.trace(self.A @ self.B)
        xPx = self.Pt1 @ np.sum(self.X_hat * self.X_hat, axis=1)
        trBYPYB = np.trace(self.B @ self.YPY @ self.B)
        self.q = (xPx - 2.0 * trAB + trBYPYB) / (2.0 * self.sigma2) + self.D * self.Np / 2.0 * np.log(self.sigma2)
        self.diff = abs(self.q - qprev)
        self.sigma2 = (xPx - trAB) / (self.Np * self.D)
        if self.sigma2 <= 0:
            self.sigma2 = self.tolerance / 10.0

    def get_registration_parameters(self):
        return self.B, self.t


class DeformableRegistration(_EM):
    """CPD paper Fig. 4 with the low-rank kernel of section 7.  Reference call site: focusr.py:299-316."""

    def __init__(self, *args, alpha=2.0, beta=2.0, num_eig=100, **kwargs):
        super().__init__(*args, **kwargs)
        self.alpha, self.beta, self.num_eig = float(alpha), float(beta), int(min(num_eig, self.Y.shape[0]))
        self.W = np.zeros((self.M, self.D))
        self.G = gaussian_kernel(self.Y, self.beta)
        self.Q, self.S = low_rank_eigen(self.G, self.num_eig)
        self.E = 0.0

    def update_transform(self):
        dPQ = self.P1[:, None] * self.Q
        F = self.PX - self.P1[:, None] * self.Y
        lam = self.alpha * self.sigma2
        inner = np.linalg.solve(np.diag(lam / self.S) + self.Q.T @ dPQ, self.Q.T @ F)
        self.W = (F - dPQ @ inner) / lam
        QtW = self.Q.T @ self.W
        self.E += self.alpha / 2.0 * np.trace(QtW.T @ (self.S[:, None] * QtW))

    def transform_point_cloud(self, Y=None):
        if Y is None:
            self.TY = self.Y + self.Q @ (self.S[:, None] * (self.Q.T @ self.W))
            return None
        Y = np.asarray(Y, dtype=np.float64)
        out = np.empty_like(Y)
        for s in range(0, Y.shape[0], 2048):                                        # bounded memory
            out[s:s + 2048] = Y[s:s + 2048] + gaussian_kernel(Y[s:s + 2048], self.beta, self.Y) @ self.W
        return out

    def update_variance(self):
        qprev = self.sigma2
        xPx = self.Pt1 @ np.sum(self.X * self.X, axis=1)
        yPy = self.P1 @ np.sum(self.TY * self.TY, axis=1)
        trPXY = np.sum(self.TY * self.PX)
        self.sigma2 = (xPx - 2.0 * trPXY + yPy) / (self.Np * self.D)
        if self.sigma2 <= 0:
            self.sigma2 = self.tolerance / 10.0
        self.diff = abs(self.sigma2 - qprev)

    def get_registration_parameters(self):
        return self.G, self.W


def register_target_to_source(target_coords, source_coords, idx_t, idx_s, rigid_first=True, rigid_max_iterations=100,
                              rigid_tolerance=1e-8, max_iterations=1000, tolerance=1e-8, alpha=0.5, beta=3.0,
                              num_eig=100, idx_t2=None, idx_s2=None):
    """focusr.py:537-543 + 297-334: affine then deformable CPD of the target coordinates onto the source's,
    fitted on the given subsets and applied to all target points.  The reference draws fresh random subsets for
    each of the two registrations (focusr.py:301-306, 321-326): ``idx_t2`` / ``idx_s2`` are the deformable
    stage's (default: the same as the affine stage's).  Returns the transformed target coords and a dict of
    parameters/iteration counts."""
    info = {}
    tc = np.asarray(target_coords, dtype=np.float64)
    idx_t2 = idx_t if idx_t2 is None else idx_t2
    idx_s2 = idx_s if idx_s2 is None else idx_s2
    if rigid_first:
        reg = AffineRegistration(source_coords[idx_s], tc[idx_t], max_iterations=rigid_max_iterations,
                                 tolerance=rigid_tolerance)
        reg.register()
        tc = reg.transform_point_cloud(tc)
        info.update(B=reg.B, t=reg.t, affine_iterations=reg.iteration, affine_sigma2=reg.sigma2)
    reg = DeformableRegistration(source_coords[idx_s2], tc[idx_t2], max_iterations=max_iterations, tolerance=tolerance,
                                 alpha=alpha, beta=beta, num_eig=num_eig)
    reg.register()
    tc = reg.transform_point_cloud(tc)
    info.update(W=reg.W, iterations=reg.iteration, sigma2=reg.sigma2)
    return tc, info
