// Small dense kernels of the Rayleigh-Ritz step (b x b, b <= 96), written once for host and
// device.  On the GPU one warp owns one mesh's matrices in shared memory and `Par` spreads the
// inner loops over the 32 lanes; on the host (tests/hostsim, never shipped) the same code runs
// with a sequential `Par`, so the control flow exercised by the CPU tests is the control flow
// the B200 executes.
//
// Replaces, for the symmetric (closed, consistently oriented mesh) case, the dense part of
// scipy's ARPACK `dneupd` extraction used at reference graph.py:372.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif

namespace fb {

struct SeqPar {
  template <class F>
  FB_HD void for_n(int n, F f) const {
    for (int i = 0; i < n; ++i) f(i);
  }
  FB_HD void sync() const {}
  FB_HD double sum(double v) const { return v; }  // sequential for_n already accumulated everything
  FB_HD int lane() const { return 0; }
};

#if defined(__CUDACC__)
// a whole CTA owns the matrix (large blocks, one mesh): red is blockDim.x doubles of shared scratch
struct BlockPar {
  double* red;
  template <class F>
  __device__ __forceinline__ void for_n(int n, F f) const {
    for (int i = threadIdx.x; i < n; i += blockDim.x) f(i);
  }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ double sum(double v) const {
    __syncthreads();
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
  }
  __device__ __forceinline__ int lane() const { return threadIdx.x; }
};

struct WarpPar {
  int lane_;
  template <class F>
  __device__ __forceinline__ void for_n(int n, F f) const {
    for (int i = lane_; i < n; i += 32) f(i);
  }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ double sum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  __device__ __forceinline__ int lane() const { return lane_; }
};
#endif

// In-place upper Cholesky of the SPD matrix g (row-major, ld = b):  g = R^T R, R stored in the
// upper triangle (strict lower triangle left untouched).  Returns the number of pivots that had
// to be clamped (0 = clean factorisation).
template <class Par>
FB_HD int cholesky_upper(double* g, double* diag0, int b, const Par& par) {
  // diag0 (b doubles of scratch) keeps the original diagonal: a pivot that has lost ~14 digits
  // against it means column k is numerically dependent on the previous ones.
  par.for_n(b, [&](int i) { diag0[i] = g[i * b + i]; });
  int bad = 0;
  for (int k = 0; k < b; ++k) {
    par.sync();
    double piv = g[k * b + k];
    const double ref = diag0[k];
    if (!(piv > 1e-14 * ref) || !(ref > 0.0)) {  // also catches NaN
      piv = (ref > 0.0 ? ref : 1.0) * 1e-14;
      ++bad;
    }
    const double r = sqrt(piv);
    const double rinv = 1.0 / r;
    par.sync();
    par.for_n(b - k, [&](int t) {
      const int j = k + t;
      if (j == k)
        g[k * b + k] = r;
      else
        g[k * b + j] *= rinv;
    });
    par.sync();
    par.for_n(b - k - 1, [&](int t) {
      const int j = k + 1 + t;
      const double gkj = g[k * b + j];
      for (int i = k + 1; i <= j; ++i) g[i * b + j] -= g[k * b + i] * gkj;
    });
  }
  par.sync();
  return bad;
}

// h <- R^-T h R^-1 (R upper triangular in the upper triangle of r), then symmetrised.
template <class Par>
FB_HD void congruence_upper(double* h, const double* r, int b, const Par& par) {
  // T = R^-T H : row i of T from rows < i (forward substitution), parallel over columns
  for (int i = 0; i < b; ++i) {
    par.sync();
    const double dinv = 1.0 / r[i * b + i];
    par.for_n(b, [&](int j) {
      double v = h[i * b + j];
      for (int k = 0; k < i; ++k) v -= r[k * b + i] * h[k * b + j];
      h[i * b + j] = v * dinv;
    });
  }
  // H' = T R^-1 : column j from columns < j, parallel over rows
  for (int j = 0; j < b; ++j) {
    par.sync();
    const double dinv = 1.0 / r[j * b + j];
    par.for_n(b, [&](int i) {
      double v = h[i * b + j];
      for (int k = 0; k < j; ++k) v -= h[i * b + k] * r[k * b + j];
      h[i * b + j] = v * dinv;
    });
  }
  par.sync();
  par.for_n(b, [&](int i) {
    for (int j = i + 1; j < b; ++j) {
      const double v = 0.5 * (h[i * b + j] + h[j * b + i]);
      h[i * b + j] = v;
    }
  });
  par.sync();
  par.for_n(b, [&](int i) {
    for (int j = 0; j < i; ++j) h[i * b + j] = h[j * b + i];
  });
  par.sync();
}

// Jacobi eigen-decomposition of the symmetric matrix a (row-major, full storage) with the
// round-robin ("chess tournament") parallel ordering: each of the m-1 rounds of a sweep applies
// m/2 rotations on disjoint index pairs, so all threads of the warp / CTA owning the matrix work at
// once (column phase A <- A J, Y <- Y J, then row phase A <- J^T A).  On return diag(a) holds the
// eigenvalues and the columns of y the eigenvectors (a_in = y D y^T).
// rot: 2*(b/2+1) doubles (cos, sin per pair), pq: 2*(b/2+1) ints.  Returns the number of sweeps.
template <class Par>
FB_HD int jacobi_sym(double* a, double* y, int b, double* rot, int* pq, const Par& par) {
  par.for_n(b, [&](int i) {
    for (int j = 0; j < b; ++j) y[i * b + j] = (i == j) ? 1.0 : 0.0;
  });
  par.sync();
  double loc = 0.0;
  par.for_n(b, [&](int i) {
    for (int j = 0; j < b; ++j) loc += a[i * b + j] * a[i * b + j];
  });
  const double normf = sqrt(par.sum(loc));
  const double thr = 1e-17 * normf;
  const int m = (b + 1) & ~1;  // players (one bye if b is odd)
  const int half = m / 2;
  double* cs = rot;
  double* sn = rot + half;
  int* pp = pq;
  int* qq = pq + half;
  int sweep = 0;
  for (; sweep < 40; ++sweep) {
    double nrot = 0.0;  // rotations applied in this sweep (every thread's own count, summed below)
    for (int r = 0; r < m - 1; ++r) {
      par.for_n(half, [&](int k) {
        const int x = (r + k) % (m - 1);
        const int z = (k == 0) ? (m - 1) : (r - k + (m - 1)) % (m - 1);
        const int p = x < z ? x : z, q = x < z ? z : x;
        double c = 1.0, s = 0.0;
        if (q < b) {
          const double apq = a[p * b + q];
          if (fabs(apq) > thr) {
            const double tau = (a[q * b + q] - a[p * b + p]) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        }
        cs[k] = c;
        sn[k] = s;
        pp[k] = p;
        qq[k] = q;
        if (s != 0.0) nrot += 1.0;
      });
      par.sync();
      par.for_n(half * b, [&](int e) {  // column phase
        const int k = e / b, i = e - k * b;
        const int p = pp[k], q = qq[k];
        const double c = cs[k], s = sn[k];
        if (q < b && s != 0.0) {
          const double aip = a[i * b + p], aiq = a[i * b + q];
          a[i * b + p] = c * aip - s * aiq;
          a[i * b + q] = s * aip + c * aiq;
          const double yip = y[i * b + p], yiq = y[i * b + q];
          y[i * b + p] = c * yip - s * yiq;
          y[i * b + q] = s * yip + c * yiq;
        }
      });
      par.sync();
      par.for_n(half * b, [&](int e) {  // row phase
        const int k = e / b, j = e - k * b;
        const int p = pp[k], q = qq[k];
        const double c = cs[k], s = sn[k];
        if (q < b && s != 0.0) {
          const double apj = a[p * b + j], aqj = a[q * b + j];
          a[p * b + j] = c * apj - s * aqj;
          a[q * b + j] = s * apj + c * aqj;
        }
      });
      par.sync();
      par.for_n(half, [&](int k) {  // the annihilated pair, exactly
        const int p = pp[k], q = qq[k];
        if (q < b && sn[k] != 0.0) {
          a[p * b + q] = 0.0;
          a[q * b + p] = 0.0;
        }
      });
      par.sync();
    }
    if (!(par.sum(nrot) > 0.0)) break;  // a full sweep without a rotation: every |a_pq| <= thr
  }
  par.sync();
  return sweep;
}

// Ascending rank of diag(a): rank[i] = position of eigenvalue i in sorted order (ties by index).
template <class Par>
FB_HD void rank_ascending(const double* a, int* rank, int b, const Par& par) {
  par.for_n(b, [&](int i) {
    const double ti = a[i * b + i];
    int r = 0;
    for (int j = 0; j < b; ++j) {
      const double tj = a[j * b + j];
      r += (tj < ti) || (tj == ti && j < i);
    }
    rank[i] = r;
  });
  par.sync();
}

// w[:, rank[j]] = R^-1 y[:, j]  (back substitution), theta[rank[j]] = a[j][j].
template <class Par>
FB_HD void ritz_basis(const double* r, const double* y, const double* a, const int* rank, double* w,
                      double* theta, int b, const Par& par) {
  par.for_n(b, [&](int j) {
    const int jo = rank[j];
    for (int i = b - 1; i >= 0; --i) {
      double v = y[i * b + j];
      for (int k = i + 1; k < b; ++k) v -= r[i * b + k] * w[k * b + jo];
      w[i * b + jo] = v / r[i * b + i];
    }
    theta[jo] = a[j * b + j];
  });
  par.sync();
}

// Whole symmetric Rayleigh-Ritz step:  given G = X^T D X (SPD) and H = X^T (D-A) X (symmetric),
// find W (b x b) and theta ascending with  (X W)^T D (X W) = I,  (X W)^T (D-A) (X W) = diag(theta).
// g, h are overwritten (g <- R, h <- rotated); y, w are b*b scratch/outputs; rank is b ints;
// rot (b+2 doubles) and pq (b+2 ints) are scratch of the Jacobi rounds.
// Returns (#clamped pivots << 8) | #sweeps.
template <class Par>
FB_HD int rayleigh_ritz_sym(double* g, double* h, double* y, double* w, double* theta, int* rank,
                            double* rot, int* pq, int b, const Par& par) {
  const int bad = cholesky_upper(g, theta, b, par);
  congruence_upper(h, g, b, par);
  const int sweeps = jacobi_sym(h, y, b, rot, pq, par);
  rank_ascending(h, rank, b, par);
  ritz_basis(g, y, h, rank, w, theta, b, par);
  return (bad << 8) | sweeps;
}

}  // namespace fb
