// Host-side dense eigen-solver for the small (b x b, b <= 128) NON-symmetric Rayleigh quotient.
//
// Open / non-manifold meshes give a structurally non-symmetric adjacency (reference
// graph.py:158-178 writes one direction per cell edge), so L = D^-1 (D - A) has a few complex
// eigenvalue pairs.  The block iteration then projects onto a subspace
// in the Euclidean inner product and needs the eigen-decomposition of a small general real
// matrix -- what ARPACK's dneupd/dlahqr do inside scipy `eigs` (reference graph.py:372).  It is
// O(b^3) work on a matrix of at most a few hundred entries per outer iteration, so it stays on
// the host next to the driver loop; the N-sized work stays on the GPU.
//
// Algorithm: Householder reduction to Hessenberg form, explicit single-shift (Wilkinson) QR in
// complex arithmetic to a complex Schur form A = Q T Q^H, eigenvectors by back substitution.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace fb {

typedef std::complex<double> cplx;

// a: n x n row-major real matrix.  evals[n]; evecs n x n row-major, column j = eigenvector j (unit
// 2-norm).  Returns 0 on success, >0 = number of eigenvalues that failed to converge.
inline int eig_general(const double* a, int n, cplx* evals, cplx* evecs) {
  std::vector<cplx> H(n * n), Q(n * n, cplx(0.0));
  for (int i = 0; i < n * n; ++i) H[i] = a[i];
  for (int i = 0; i < n; ++i) Q[i * n + i] = 1.0;
  auto h = [&](int i, int j) -> cplx& { return H[i * n + j]; };
  auto q = [&](int i, int j) -> cplx& { return Q[i * n + j]; };

  // --- Hessenberg reduction (Householder), Q accumulates the reflectors
  std::vector<cplx> v(n);
  for (int k = 0; k + 2 < n; ++k) {
    double nrm = 0.0;
    for (int i = k + 1; i < n; ++i) nrm += std::norm(h(i, k));
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) continue;
    const cplx x0 = h(k + 1, k);
    const cplx phase = (std::abs(x0) > 0.0) ? x0 / std::abs(x0) : cplx(1.0);
    const cplx alpha = -phase * nrm;
    double vn = 0.0;
    for (int i = k + 1; i < n; ++i) {
      v[i] = h(i, k) - (i == k + 1 ? alpha : cplx(0.0));
      vn += std::norm(v[i]);
    }
    vn = std::sqrt(vn);
    if (vn == 0.0) continue;
    for (int i = k + 1; i < n; ++i) v[i] /= vn;
    // H <- (I - 2 v v^H) H
    for (int j = 0; j < n; ++j) {
      cplx s = 0.0;
      for (int i = k + 1; i < n; ++i) s += std::conj(v[i]) * h(i, j);
      s *= 2.0;
      for (int i = k + 1; i < n; ++i) h(i, j) -= v[i] * s;
    }
    // H <- H (I - 2 v v^H),  Q <- Q (I - 2 v v^H)
    for (int i = 0; i < n; ++i) {
      cplx s = 0.0, t = 0.0;
      for (int j = k + 1; j < n; ++j) {
        s += h(i, j) * v[j];
        t += q(i, j) * v[j];
      }
      s *= 2.0;
      t *= 2.0;
      for (int j = k + 1; j < n; ++j) {
        h(i, j) -= s * std::conj(v[j]);
        q(i, j) -= t * std::conj(v[j]);
      }
    }
    for (int i = k + 2; i < n; ++i) h(i, k) = 0.0;
  }

  // --- shifted QR to (complex) Schur form
  const double eps = 2.220446049250313e-16;
  double hnorm = 0.0;
  for (int i = 0; i < n * n; ++i) hnorm = std::max(hnorm, std::abs(H[i]));
  if (hnorm == 0.0) hnorm = 1.0;
  std::vector<cplx> cs(n), sn(n);
  int failed = 0;
  int ihi = n - 1;
  int iter = 0;
  while (ihi > 0) {
    int l = ihi;
    while (l > 0) {
      double s = std::abs(h(l - 1, l - 1)) + std::abs(h(l, l));
      if (s == 0.0) s = hnorm;
      if (std::abs(h(l, l - 1)) <= eps * s) {
        h(l, l - 1) = 0.0;
        break;
      }
      --l;
    }
    if (l == ihi) {
      --ihi;
      iter = 0;
      continue;
    }
    if (++iter > 60 * 1) {  // give up on this eigenvalue, deflate by force
      ++failed;
      h(ihi, ihi - 1) = 0.0;
      --ihi;
      iter = 0;
      continue;
    }
    cplx shift;
    if (iter == 10 || iter == 20 || iter == 30) {
      shift = h(ihi, ihi) + cplx(std::abs(h(ihi, ihi - 1)) + (ihi > 1 ? std::abs(h(ihi - 1, ihi - 2)) : 0.0));
    } else {
      const cplx aa = h(ihi - 1, ihi - 1), bb = h(ihi - 1, ihi), cc = h(ihi, ihi - 1), dd = h(ihi, ihi);
      const cplx half = 0.5 * (aa + dd);
      const cplx disc = std::sqrt(0.25 * (aa - dd) * (aa - dd) + bb * cc);
      const cplx m1 = half + disc, m2 = half - disc;
      shift = (std::abs(m1 - dd) < std::abs(m2 - dd)) ? m1 : m2;
    }
    for (int i = l; i <= ihi; ++i) h(i, i) -= shift;
    for (int k = l; k < ihi; ++k) {
      const cplx x = h(k, k), y = h(k + 1, k);
      const double r = std::sqrt(std::norm(x) + std::norm(y));
      cplx c = 1.0, s = 0.0;
      if (r > 0.0) {
        c = x / r;
        s = y / r;
      }
      cs[k] = c;
      sn[k] = s;
      for (int j = k; j < n; ++j) {
        const cplx t0 = h(k, j), t1 = h(k + 1, j);
        h(k, j) = std::conj(c) * t0 + std::conj(s) * t1;
        h(k + 1, j) = -s * t0 + c * t1;
      }
      h(k + 1, k) = 0.0;
    }
    for (int k = l; k < ihi; ++k) {
      const cplx c = cs[k], s = sn[k];
      const int rmax = std::min(k + 2, ihi);
      for (int i = 0; i <= rmax; ++i) {
        const cplx t0 = h(i, k), t1 = h(i, k + 1);
        h(i, k) = t0 * c + t1 * s;
        h(i, k + 1) = -t0 * std::conj(s) + t1 * std::conj(c);
      }
      for (int i = 0; i < n; ++i) {
        const cplx t0 = q(i, k), t1 = q(i, k + 1);
        q(i, k) = t0 * c + t1 * s;
        q(i, k + 1) = -t0 * std::conj(s) + t1 * std::conj(c);
      }
    }
    for (int i = l; i <= ihi; ++i) h(i, i) += shift;
  }

  // --- eigenvectors of the triangular factor, back-transformed
  std::vector<cplx> y(n);
  const double small = eps * hnorm;
  for (int k = 0; k < n; ++k) {
    const cplx lam = h(k, k);
    evals[k] = lam;
    for (int i = 0; i < n; ++i) y[i] = 0.0;
    y[k] = 1.0;
    for (int i = k - 1; i >= 0; --i) {
      cplx s = 0.0;
      for (int j = i + 1; j <= k; ++j) s += h(i, j) * y[j];
      cplx d = h(i, i) - lam;
      if (std::abs(d) < small) d = small;
      y[i] = -s / d;
    }
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) {
      cplx s = 0.0;
      for (int j = 0; j <= k; ++j) s += q(i, j) * y[j];
      evecs[i * n + k] = s;
      nrm += std::norm(s);
    }
    nrm = std::sqrt(nrm);
    if (nrm > 0.0)
      for (int i = 0; i < n; ++i) evecs[i * n + k] /= nrm;
  }
  return failed;
}

}  // namespace fb
