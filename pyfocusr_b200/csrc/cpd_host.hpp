// Host-side small dense algebra of the CPD stage: the b x b (b ~ 128) symmetric eigenproblem of the
// subspace iteration that finds the leading eigenpairs of the Gaussian kernel matrix (csrc/cpd.cu).
// Householder tridiagonalisation followed by implicit-shift QL (the classic EISPACK tred2/tql2 pair),
// O(b^3) once per subspace iteration -- a few milliseconds -- where the shared-memory Jacobi of
// dense_small.h would not fit (3 b^2 doubles > 227 KB for b > 96).  Pure C++; tests/hostsim exports it.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "dense_small.h"

namespace fb {

// Symmetric tridiagonalisation.  On entry v (n x n, row-major) holds the symmetric matrix; on return
// v holds the orthogonal transformation, d the diagonal and e the sub-diagonal (e[0] = 0).
inline void tridiagonalize_sym(double* v, double* d, double* e, int n) {
  auto V = [&](int i, int j) -> double& { return v[(size_t)i * n + j]; };
  for (int j = 0; j < n; ++j) d[j] = V(n - 1, j);
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) {
        d[j] = V(i - 1, j);
        V(i, j) = 0.0;
        V(j, i) = 0.0;
      }
    } else {
      for (int k = 0; k < i; ++k) {
        d[k] /= scale;
        h += d[k] * d[k];
      }
      double f = d[i - 1];
      double g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j];
        V(j, i) = f;
        g = e[j] + V(j, j) * f;
        for (int k = j + 1; k <= i - 1; ++k) {
          g += V(k, j) * d[k];
          e[k] += V(k, j) * f;
        }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) {
        e[j] /= h;
        f += e[j] * d[j];
      }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j];
        g = e[j];
        for (int k = j; k <= i - 1; ++k) V(k, j) -= (f * e[k] + g * d[k]);
        d[j] = V(i - 1, j);
        V(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; ++i) {
    V(n - 1, i) = V(i, i);
    V(i, i) = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; ++k) d[k] = V(k, i + 1) / h;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        for (int k = 0; k <= i; ++k) g += V(k, i + 1) * V(k, j);
        for (int k = 0; k <= i; ++k) V(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) V(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) {
    d[j] = V(n - 1, j);
    V(n - 1, j) = 0.0;
  }
  V(n - 1, n - 1) = 1.0;
  e[0] = 0.0;
}

// Implicit QL on the tridiagonal (d, e) accumulating into v, which is held TRANSPOSED (vt[j][k] = V[k][j]):
// every plane rotation then updates two contiguous rows instead of two strided columns (3x faster at n = 128).
// Returns 0, or -1 after 60 iterations on one eigenvalue.  On return d holds the eigenvalues (unsorted) and the
// rows of vt the eigenvectors.
inline int ql_implicit(double* vt, double* d, double* e, int n) {
  auto V = [&](int i, int j) -> double& { return vt[(size_t)j * n + i]; };
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) {
      if (std::fabs(e[m]) <= eps * tst1) break;
      ++m;
    }
    if (m > l) {
      int iter = 0;
      do {
        if (++iter > 60) return -1;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c;
        const double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; ++k) {
            h = V(k, i + 1);
            V(k, i + 1) = s * V(k, i) + c * h;
            V(k, i) = c * V(k, i) - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1);
    }
    d[l] += f;
    e[l] = 0.0;
  }
  return 0;
}

// Eigen-decomposition of the symmetric a (n x n row-major, overwritten by the eigenvectors as columns);
// evals unsorted.  Returns 0 or -1.
inline int eig_sym_host(double* a, double* evals, int n) {
  std::vector<double> e(n);
  tridiagonalize_sym(a, evals, e.data(), n);
  auto transpose = [&]() {
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) std::swap(a[(size_t)i * n + j], a[(size_t)j * n + i]);
  };
  transpose();
  const int rc = ql_implicit(a, evals, e.data(), n);
  transpose();
  return rc;
}

// Rayleigh-Ritz for the leading eigenpairs of a symmetric operator: g = X^T X, h = X^T (G X) (b x b
// row-major, both overwritten).  Produces w (b x b) and theta ordered by DESCENDING |theta| with
// (X W)^T (X W) = I and (X W)^T G (X W) = diag(theta).  Returns the number of clamped Cholesky pivots,
// or -1 if the QL iteration failed.
inline int rr_leading_host(double* g, double* h, int b, double* w, double* theta) {
  SeqPar par;
  std::vector<double> diag0(b), ev(b);
  const int bad = cholesky_upper(g, diag0.data(), b, par);
  congruence_upper(h, g, b, par);
  if (eig_sym_host(h, ev.data(), b) != 0) return -1;
  std::vector<int> order(b);
  for (int i = 0; i < b; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return std::fabs(ev[x]) > std::fabs(ev[y]); });
  // w[:, c] = R^-1 y[:, order[c]]
  for (int c = 0; c < b; ++c) {
    const int src = order[c];
    theta[c] = ev[src];
    for (int i = b - 1; i >= 0; --i) {
      double v = h[(size_t)i * b + src];
      for (int k = i + 1; k < b; ++k) v -= g[(size_t)i * b + k] * w[(size_t)k * b + c];
      w[(size_t)i * b + c] = v / g[(size_t)i * b + i];
    }
  }
  return bad;
}

}  // namespace fb
