// K8  Coherent Point Drift on the B200 -- the registration step in the middle of the path
// (reference focusr.py:297-334: cycpd.affine_registration then cycpd.deformable_registration on random
// subsets of <= 5000 spectral coordinates, then transform_point_cloud on all target coordinates;
// SURVEY.md section 8f-1).  Algorithm: Myronenko & Song, TPAMI 2010 (affine Fig. 3, non-rigid Fig. 4 with the
// low-rank kernel of section 7); conventions as stated in oracle/cpd_port.py.
//
// Design: the whole EM loop runs on the device.  All scalars of a registration (sigma2, objective, Np, the
// `active` flag, ...) live in a CpdState in HBM; every per-iteration kernel starts with `if (!active) return`,
// and the last kernel of an iteration re-evaluates `active = iteration < max && diff > tolerance`.  The host
// enqueues iterations in batches and reads the state back once per batch, so the convergence test costs no
// per-iteration synchronisation and the iteration count is exactly the one a per-iteration test would give.
//
//   E-step      two passes over the M x N Gaussian affinities, never materialised:
//               k_cpd_colsum (thread = one x_n, TY chunk in shared memory) -> column sums -> 1/den, Pt1;
//               k_cpd_rowsum (thread = one y_m, X chunk in shared memory)  -> P1, P X.
//               fp64 exp dominated: 2 M N (3 D + exp) flops per iteration, FP64-ALU-bound.
//   affine M    centred moments (D x D) by chunk + fixed-order sums, LU with partial pivoting in one thread.
//   deformable  G = Q S Q^T: leading eigenpairs by subspace iteration (DMMA GEMM G X, DMMA tall-skinny Gram
//               blocks, b x b Rayleigh-Ritz on the host: cpd_host.hpp);  M-step by the Woodbury identity:
//               T1 = Q^T dP Q (DMMA), T2 = Q^T F, (lambda S^-1 + T1) Z = T2 by LU with partial pivoting in one
//               CTA (shared memory), W = (F - dP Q Z) / lambda, TY = Y + Q S Q^T W.
// Every reduction has a fixed order: results are bit-reproducible run to run.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "cpd_host.hpp"
#include "rowops.h"

namespace fb {

constexpr int CPD_MAXD = 16;
constexpr int CPD_T = 128;      // threads per CTA of the E-step kernels
constexpr int CPD_CHUNK = 128;  // points of the other set staged in shared memory per CTA
constexpr int CPD_ACH = 128;    // control points per stage of the deformable transform kernel
constexpr int CPD_MCH = 128;    // rows per CTA of the moment kernels
constexpr int CPD_TN_ROWS = 128;  // rows per warp of the tall-skinny Gram kernel
constexpr int CPD_MAX_RP = 152;   // largest padded rank the one-CTA LU holds in shared memory

struct CpdState {
  double sigma2, q, diff, np, tolerance, w, lam;
  double mu_x[CPD_MAXD], mu_y[CPD_MAXD];
  int iteration, max_iterations, active, singular;
};

__device__ __forceinline__ void dmma884c(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// fixed-tree block sum (blockDim.x a power of two <= 1024); every thread gets the total
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int t = threadIdx.x;
  __syncthreads();
  red[t] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (t < s) red[t] += red[t + s];
    __syncthreads();
  }
  return red[0];
}

// ------------------------------------------------------------------------------------------- E-step
// MODE 0: partial[ch][n] = sum_{m in chunk} exp(-|x_n - ty_m|^2 / (2 sigma2));  MODE 1: sum of |x_n - ty_m|^2
template <int D, int MODE>
__global__ void __launch_bounds__(CPD_T)
k_cpd_colsum(const double* __restrict__ x, int N, const double* __restrict__ ty, int M, const CpdState* __restrict__ st,
             double* __restrict__ partial) {
  if (MODE == 0 && !st->active) return;
  __shared__ double s[CPD_CHUNK * D];
  const int ch = blockIdx.y, m0 = ch * CPD_CHUNK, cnt = min(M - m0, CPD_CHUNK);
  for (int i = threadIdx.x; i < cnt * D; i += CPD_T) s[i] = ty[(size_t)m0 * D + i];
  __syncthreads();
  const int n = blockIdx.x * CPD_T + threadIdx.x;
  if (n >= N) return;
  double xv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) xv[d] = x[(size_t)n * D + d];
  const double ninv = MODE == 0 ? -0.5 / st->sigma2 : 0.0;
  double acc = 0.0;
  for (int m = 0; m < cnt; ++m) {
    double d2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double t = xv[d] - s[m * D + d];
      d2 += t * t;
    }
    acc += MODE == 0 ? exp(ninv * d2) : d2;
  }
  partial[(size_t)ch * N + n] = acc;
}

// sigma2 = sum of all squared pair distances / (D M N)   (single CTA, fixed order)
__global__ void __launch_bounds__(1024) k_cpd_sigma_init(const double* __restrict__ partial, long long count, int D, int M, int N,
                                                        CpdState* st) {
  __shared__ double red[1024];
  double v = 0.0;
  for (long long i = threadIdx.x; i < count; i += 1024) v += partial[i];
  const double tot = block_sum(v, red);
  if (threadIdx.x == 0) st->sigma2 = tot / ((double)D * (double)M * (double)N);
}

__global__ void k_cpd_den(const double* __restrict__ partial, int chunks, int N, int M, int D, const CpdState* __restrict__ st,
                          double* __restrict__ inv_den, double* __restrict__ pt1) {
  if (!st->active) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double sum = 0.0;
  for (int c = 0; c < chunks; ++c) sum += partial[(size_t)c * N + n];
  const double w = st->w;
  const double c = w > 0.0 ? pow(2.0 * 3.141592653589793 * st->sigma2, 0.5 * D) * w / (1.0 - w) * (double)M / (double)N : 0.0;
  const double den = fmax(sum, DBL_EPSILON) + c;
  inv_den[n] = 1.0 / den;
  pt1[n] = sum / den;
}

// partial[ch][m][0..D-1] = sum_n P[m][n] x_n,  partial[ch][m][D] = sum_n P[m][n]   (n in chunk)
template <int D>
__global__ void __launch_bounds__(CPD_T)
k_cpd_rowsum(const double* __restrict__ x, int N, const double* __restrict__ ty, int M, const double* __restrict__ inv_den,
             const CpdState* __restrict__ st, double* __restrict__ partial) {
  if (!st->active) return;
  __shared__ double s[CPD_CHUNK * (D + 1)];
  const int ch = blockIdx.y, n0 = ch * CPD_CHUNK, cnt = min(N - n0, CPD_CHUNK);
  for (int i = threadIdx.x; i < cnt * D; i += CPD_T) s[(i / D) * (D + 1) + (i % D)] = x[(size_t)n0 * D + i];
  for (int i = threadIdx.x; i < cnt; i += CPD_T) s[i * (D + 1) + D] = inv_den[n0 + i];
  __syncthreads();
  const int m = blockIdx.x * CPD_T + threadIdx.x;
  if (m >= M) return;
  double yv[D], px[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    yv[d] = ty[(size_t)m * D + d];
    px[d] = 0.0;
  }
  const double ninv = -0.5 / st->sigma2;
  double p1 = 0.0;
  for (int n = 0; n < cnt; ++n) {
    const double* sx = s + n * (D + 1);
    double d2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double t = sx[d] - yv[d];
      d2 += t * t;
    }
    const double k = exp(ninv * d2) * sx[D];
    p1 += k;
#pragma unroll
    for (int d = 0; d < D; ++d) px[d] += k * sx[d];
  }
  double* o = partial + ((size_t)ch * M + m) * (D + 1);
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = px[d];
  o[D] = p1;
}

__global__ void k_cpd_rowreduce(const double* __restrict__ partial, int chunks, int M, int D, const CpdState* __restrict__ st,
                                double* __restrict__ p1, double* __restrict__ px) {
  if (!st->active) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * (D + 1)) return;
  double sum = 0.0;
  for (int c = 0; c < chunks; ++c) sum += partial[(size_t)c * M * (D + 1) + i];
  const int m = i / (D + 1), e = i % (D + 1);
  if (e == D)
    p1[m] = sum;
  else
    px[(size_t)m * D + e] = sum;
}

// ------------------------------------------------------------------------------------------- affine M-step
// Np, muX = sum PX / Np, muY = sum P1 y / Np   (single CTA)
template <int D>
__global__ void __launch_bounds__(1024)
k_cpd_means(const double* __restrict__ p1, const double* __restrict__ px, const double* __restrict__ y, int M, CpdState* st) {
  if (!st->active) return;
  __shared__ double red[1024];
  double np = 0.0, sx[D], sy[D];
#pragma unroll
  for (int d = 0; d < D; ++d) sx[d] = sy[d] = 0.0;
  for (int m = threadIdx.x; m < M; m += 1024) {
    const double p = p1[m];
    np += p;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      sx[d] += px[(size_t)m * D + d];
      sy[d] += p * y[(size_t)m * D + d];
    }
  }
  np = block_sum(np, red);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double a = block_sum(sx[d], red), b = block_sum(sy[d], red);
    if (threadIdx.x == 0) {
      st->mu_x[d] = a / np;
      st->mu_y[d] = b / np;
    }
  }
  if (threadIdx.x == 0) st->np = np;
}

// CTAs [0, chm): partial[b][i*D+j] = sum_r (PX_r - P1_r muX)_i (y_r - muY)_j ; partial[b][D*D + i*D+j] = sum_r P1_r yh_i yh_j
// CTAs [chm, chm+chn): partial[b][2*D*D] = sum_n Pt1_n |x_n - muX|^2
template <int D>
__global__ void __launch_bounds__(256)
k_cpd_affine_moments(const double* __restrict__ p1, const double* __restrict__ px, const double* __restrict__ y, int M,
                     const double* __restrict__ pt1, const double* __restrict__ x, int N, const CpdState* __restrict__ st,
                     double* __restrict__ partial, int chm) {
  if (!st->active) return;
  constexpr int STRIDE = 2 * D * D + 1;
  __shared__ double sa[CPD_MCH * D], sy[CPD_MCH * D], sp[CPD_MCH], red[256];
  const int b = blockIdx.x, t = threadIdx.x;
  double* out = partial + (size_t)b * STRIDE;
  if (b < chm) {
    const int r0 = b * CPD_MCH, cnt = min(M - r0, CPD_MCH);
    for (int i = t; i < cnt * D; i += 256) {
      const int r = i / D, d = i % D;
      const double p = p1[r0 + r];
      sa[i] = px[(size_t)(r0 + r) * D + d] - p * st->mu_x[d];
      sy[i] = y[(size_t)(r0 + r) * D + d] - st->mu_y[d];
      if (d == 0) sp[r] = p;
    }
    __syncthreads();
    if (t < D * D) {
      const int i = t / D, j = t % D;
      double a = 0.0, ypy = 0.0;
      for (int r = 0; r < cnt; ++r) {
        a += sa[r * D + i] * sy[r * D + j];
        ypy += sp[r] * sy[r * D + i] * sy[r * D + j];
      }
      out[t] = a;
      out[D * D + t] = ypy;
    }
    if (t == 0) out[2 * D * D] = 0.0;
  } else {
    const int r0 = (b - chm) * CPD_MCH, cnt = min(N - r0, CPD_MCH);
    double v = 0.0;
    if (t < cnt) {
      double d2 = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double u = x[(size_t)(r0 + t) * D + d] - st->mu_x[d];
        d2 += u * u;
      }
      v = pt1[r0 + t] * d2;
    }
    const double tot = block_sum(v, red);
    for (int i = t; i < 2 * D * D; i += 256) out[i] = 0.0;
    if (t == 0) out[2 * D * D] = tot;
  }
}

// sums the chunk moments in order, solves YPY^T B = A^T (LU, partial pivoting), t = muX - B^T muY, then the
// objective and the new sigma2 (one CTA; the D x D algebra runs in thread 0)
template <int D>
__global__ void __launch_bounds__(512)
k_cpd_affine_solve(const double* __restrict__ partial, int n_part, CpdState* st, double* __restrict__ bm, double* __restrict__ tv) {
  if (!st->active) return;
  constexpr int STRIDE = 2 * D * D + 1;
  __shared__ double mom[STRIDE];
  for (int i = threadIdx.x; i < STRIDE; i += blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < n_part; ++c) s += partial[(size_t)c * STRIDE + i];
    mom[i] = s;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double* A = mom;          // [i][j]
  const double* YPY = mom + D * D;
  const double xPx = mom[2 * D * D];
  // solve YPY^T B = A^T : lu = YPY^T, rhs = A^T
  double lu[D][D], rhs[D][D];
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      lu[i][j] = YPY[j * D + i];
      rhs[i][j] = A[j * D + i];
    }
  int singular = 0;
  for (int k = 0; k < D; ++k) {
    int piv = k;
    double best = fabs(lu[k][k]);
    for (int i = k + 1; i < D; ++i)
      if (fabs(lu[i][k]) > best) {
        best = fabs(lu[i][k]);
        piv = i;
      }
    if (!(best > 0.0)) {
      singular = 1;
      break;
    }
    if (piv != k)
      for (int j = 0; j < D; ++j) {
        double tmp = lu[k][j];
        lu[k][j] = lu[piv][j];
        lu[piv][j] = tmp;
        tmp = rhs[k][j];
        rhs[k][j] = rhs[piv][j];
        rhs[piv][j] = tmp;
      }
    const double inv = 1.0 / lu[k][k];
    for (int i = k + 1; i < D; ++i) {
      const double f = lu[i][k] * inv;
      for (int j = k + 1; j < D; ++j) lu[i][j] -= f * lu[k][j];
      for (int j = 0; j < D; ++j) rhs[i][j] -= f * rhs[k][j];
    }
  }
  if (singular) {
    st->singular = 1;
    st->diff = 0.0;
    return;
  }
  double B[D][D];
  for (int j = 0; j < D; ++j)
    for (int i = D - 1; i >= 0; --i) {
      double v = rhs[i][j];
      for (int k = i + 1; k < D; ++k) v -= lu[i][k] * B[k][j];
      B[i][j] = v / lu[i][i];
    }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) bm[i * D + j] = B[i][j];
  for (int j = 0; j < D; ++j) {
    double v = 0.0;
    for (int i = 0; i < D; ++i) v += B[i][j] * st->mu_y[i];
    tv[j] = st->mu_x[j] - v;
  }
  double trAB = 0.0, trBYB = 0.0;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) trAB += A[i * D + j] * B[j][i];
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      double v = 0.0;
      for (int k = 0; k < D; ++k) v += YPY[j * D + k] * B[k][i];  // (YPY B)[j][i]
      trBYB += B[i][j] * v;
    }
  const double s2 = st->sigma2, np = st->np;
  const double q = (xPx - 2.0 * trAB + trBYB) / (2.0 * s2) + D * np / 2.0 * log(s2);
  st->diff = fabs(q - st->q);
  st->q = q;
  double ns = (xPx - trAB) / (np * D);
  if (ns <= 0.0) ns = st->tolerance / 10.0;
  st->sigma2 = ns;
}

// out = pts B + t.  `st` nullable: the in-loop call is conditional, the final call is not.
template <int D>
__global__ void k_cpd_affine_apply(const double* __restrict__ pts, int n, const double* __restrict__ bm, const double* __restrict__ tv,
                                   double* __restrict__ out, const CpdState* __restrict__ st) {
  if (st && !st->active) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[D];
#pragma unroll
  for (int d = 0; d < D; ++d) v[d] = pts[(size_t)i * D + d];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) s += v[d] * bm[d * D + j];
    out[(size_t)i * D + j] = s + tv[j];
  }
}

__global__ void k_cpd_finish(CpdState* st) {
  if (!st->active) return;
  st->iteration += 1;
  st->active = (st->iteration < st->max_iterations && st->diff > st->tolerance && !st->singular) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------- dense helpers
// partial[chunk][p][q] = sum_{r in chunk} A[r][p] scale[r] B[r][q]   (DMMA m8n8k4: lane l holds A[l>>2][l&3],
// B[l&3][l>>2], D[l>>2][2(l&3)+{0,1}]).  CTA = 4 warps on one chunk of 128 rows: each warp takes 32 rows as 8
// fully unrolled k-steps (all loads in flight at once -- the kernel is latency-bound, not flop-bound), then the
// four accumulator tiles are added in warp order through shared memory.  grid (chunks, ra/8, ceil(rb/8/QT)).
template <int QT>
__global__ void __launch_bounds__(128)
k_tn_gram(const double* __restrict__ A, int lda, const double* __restrict__ Bm, int ldb, const double* __restrict__ scale,
          int rows, int ra, int rb, double* __restrict__ partial, const CpdState* __restrict__ st) {
  if (st && !st->active) return;
  __shared__ double red[4][QT][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, kq = lane & 3, cc = lane >> 2;
  const int chunk = blockIdx.x, p0 = blockIdx.y * 8, t0 = blockIdx.z * QT, nbt = rb >> 3;
  const int r0 = chunk * CPD_TN_ROWS + warp * 32, r1 = min(rows, chunk * CPD_TN_ROWS + CPD_TN_ROWS);
  double acc[QT][2];
#pragma unroll
  for (int t = 0; t < QT; ++t) acc[t][0] = acc[t][1] = 0.0;
  double av[8], bv[8][QT];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int row = r0 + 4 * s + kq;
    const bool ok = row < r1;
    av[s] = ok ? A[(size_t)row * lda + p0 + cc] : 0.0;
    if (ok && scale) av[s] *= scale[row];
#pragma unroll
    for (int t = 0; t < QT; ++t) bv[s][t] = (ok && t0 + t < nbt) ? Bm[(size_t)row * ldb + 8 * (t0 + t) + cc] : 0.0;
  }
#pragma unroll
  for (int s = 0; s < 8; ++s)
#pragma unroll
    for (int t = 0; t < QT; ++t) dmma884c(acc[t][0], acc[t][1], av[s], bv[s][t]);
#pragma unroll
  for (int t = 0; t < QT; ++t) {
    red[warp][t][2 * lane] = acc[t][0];
    red[warp][t][2 * lane + 1] = acc[t][1];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int t = 0; t < QT; ++t)
      if (t0 + t < nbt) {
        const double v0 = ((red[0][t][2 * lane] + red[1][t][2 * lane]) + red[2][t][2 * lane]) + red[3][t][2 * lane];
        const double v1 = ((red[0][t][2 * lane + 1] + red[1][t][2 * lane + 1]) + red[2][t][2 * lane + 1]) + red[3][t][2 * lane + 1];
        double* o = partial + ((size_t)chunk * ra + p0 + cc) * rb + 8 * (t0 + t) + 2 * kq;
        o[0] = v0;
        o[1] = v1;
      }
  }
}

__global__ void k_sum_partials(const double* __restrict__ partial, int chunks, int n, double* __restrict__ out,
                               const CpdState* __restrict__ st) {
  if (st && !st->active) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * n + i];
  out[i] = s;
}

// C[M][ncols] = A[M][K] B[K][ncols]   (ncols a multiple of 8).  CTA = 32 rows x 64 columns, 4 warps, 16-deep
// shared-memory stages; DMMA fragments as above.  Leading dimensions padded by 4 doubles: the 16 lanes of a
// half-warp then read 16 distinct 8-byte banks.
constexpr int GE_ROWS = 32, GE_KT = 16, GE_COLS = 64;
__global__ void __launch_bounds__(128)
k_gemm_nn(const double* __restrict__ A, int lda, const double* __restrict__ Bm, int ldb, double* __restrict__ Cm, int ldc, int M,
          int K, int ncols) {
  __shared__ double As[GE_ROWS][GE_KT + 4];
  __shared__ double Bs[GE_KT][GE_COLS + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, kq = lane & 3, cc = lane >> 2;
  const int row0 = blockIdx.x * GE_ROWS, col0 = blockIdx.y * GE_COLS;
  double acc[GE_COLS / 8][2];
#pragma unroll
  for (int t = 0; t < GE_COLS / 8; ++t) acc[t][0] = acc[t][1] = 0.0;
  for (int k0 = 0; k0 < K; k0 += GE_KT) {
    for (int i = threadIdx.x; i < GE_ROWS * GE_KT; i += 128) {
      const int r = i / GE_KT, c = i % GE_KT, gr = row0 + r, gc = k0 + c;
      As[r][c] = (gr < M && gc < K) ? A[(size_t)gr * lda + gc] : 0.0;
    }
    for (int i = threadIdx.x; i < GE_KT * GE_COLS; i += 128) {
      const int r = i / GE_COLS, c = i % GE_COLS, gr = k0 + r, gc = col0 + c;
      Bs[r][c] = (gr < K && gc < ncols) ? Bm[(size_t)gr * ldb + gc] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < GE_KT; k4 += 4) {
      const double a = As[warp * 8 + cc][k4 + kq];
#pragma unroll
      for (int t = 0; t < GE_COLS / 8; ++t) dmma884c(acc[t][0], acc[t][1], a, Bs[k4 + kq][8 * t + cc]);
    }
    __syncthreads();
  }
  const int row = row0 + warp * 8 + cc;
  if (row < M) {
#pragma unroll
    for (int t = 0; t < GE_COLS / 8; ++t) {
      const int col = col0 + 8 * t + 2 * kq;
      if (col < ncols) {
        Cm[(size_t)row * ldc + col] = acc[t][0];
        Cm[(size_t)row * ldc + col + 1] = acc[t][1];
      }
    }
  }
}

// G[i][j] = exp(-|y_i - y_j|^2 / (2 beta^2))
template <int D>
__global__ void k_gauss_gram(const double* __restrict__ y, int M, double ninv, double* __restrict__ G) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= M) return;
  double d2 = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double t = y[(size_t)i * D + d] - y[(size_t)j * D + d];
    d2 += t * t;
  }
  G[(size_t)i * M + j] = exp(ninv * d2);
}

__global__ void k_cpd_eig_start(double* __restrict__ x, int M, int rb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * rb) return;
  x[i] = hash_unit((uint32_t)(i / rb), (uint32_t)(i % rb), 0x5EEDu);
}

// residual partials and the next block: res2[chunk][j] = sum_r (Zr - theta_j Xr)^2; Xn = Zr / theta_j for the
// columns that carry signal (|theta_j| >= thr), the Ritz vector itself otherwise
__global__ void k_cpd_eig_next(const double* __restrict__ xr, const double* __restrict__ zr, const double* __restrict__ theta, int M,
                               int rb, double thr, double* __restrict__ xn, double* __restrict__ res_partial) {
  const int j = threadIdx.x, chunk = blockIdx.x;
  if (j >= rb) return;
  const int r0 = chunk * CPD_TN_ROWS, r1 = min(M, r0 + CPD_TN_ROWS);
  const double th = theta[j];
  const bool power = fabs(th) >= thr;
  const double inv = power ? 1.0 / th : 0.0;
  double acc = 0.0;
  for (int r = r0; r < r1; ++r) {
    const double a = xr[(size_t)r * rb + j], z = zr[(size_t)r * rb + j];
    const double e = z - th * a;
    acc += e * e;
    xn[(size_t)r * rb + j] = power ? z * inv : a;
  }
  res_partial[(size_t)chunk * rb + j] = acc;
}

// Q[m][p] = Xr[m][p] for p < r, 0 for the padding columns
__global__ void k_cpd_store_q(const double* __restrict__ xr, int M, int rb, int r, int rp, double* __restrict__ q) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * rp) return;
  const int m = (int)(i / rp), p = (int)(i % rp);
  q[i] = p < r ? xr[(size_t)m * rb + p] : 0.0;
}

// ------------------------------------------------------------------------------------------- deformable M-step
// F = PX - P1 y  (padded to DP columns)
__global__ void k_cpd_def_prep(const double* __restrict__ p1, const double* __restrict__ px, const double* __restrict__ y, int M, int D,
                               int DP, double* __restrict__ f, const CpdState* __restrict__ st) {
  if (!st->active) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * DP) return;
  const int m = i / DP, d = i % DP;
  f[i] = d < D ? px[(size_t)m * D + d] - p1[m] * y[(size_t)m * D + d] : 0.0;
}

// (lambda S^-1 + T1) Z = T2 by blocked LU with partial pivoting; one CTA, the augmented matrix [A | T2] in
// shared memory.  Panels of 8 columns are factorised by warp 0 with warp-synchronous steps (pivot search by
// shuffles, ties to the lower row), the row swaps and the triangular solve of the panel's row block are
// column-local (one thread per column), and the trailing update is the only block-wide step: 3 barriers per
// panel instead of ~14 per column.  Back substitution: one warp per right-hand side.
constexpr int LU_PW = 8;
__global__ void __launch_bounds__(512)
k_cpd_def_solve(const double* __restrict__ t1, const double* __restrict__ t2, const double* __restrict__ S, int rp, int dp, int d_real,
                double alpha, CpdState* st, double* __restrict__ z) {
  if (!st->active) return;
  extern __shared__ double sm[];
  const int ld = rp + dp + 1, t = threadIdx.x, nt = blockDim.x, lane = t & 31, warp = t >> 5;
  double* a = sm;
  __shared__ int piv[LU_PW];
  __shared__ int bad;
  const double lam = alpha * st->sigma2;
  for (int i = t; i < rp * (rp + dp); i += nt) {
    const int r = i / (rp + dp), c = i % (rp + dp);
    double v;
    if (c < rp)
      v = t1[(size_t)r * rp + c] + (r == c ? lam / S[r] : 0.0);
    else
      v = t2[(size_t)r * dp + (c - rp)];
    a[r * ld + c] = v;
  }
  if (t == 0) {
    st->lam = lam;
    bad = 0;
  }
  __syncthreads();
  const int ncol = rp + dp;
  for (int k0 = 0; k0 < rp; k0 += LU_PW) {
    const int k1 = min(k0 + LU_PW, rp);
    if (warp == 0) {
      // the panel lives in registers: lane l owns rows l, l+32, l+64, l+96 (rp <= 160 -> 5 slots)
      constexpr int SL = 5;
      double pr[SL][LU_PW];
#pragma unroll
      for (int sl = 0; sl < SL; ++sl) {
        const int i = lane + 32 * sl;
#pragma unroll
        for (int c = 0; c < LU_PW; ++c) pr[sl][c] = (i < rp && k0 + c < k1) ? a[i * ld + k0 + c] : 0.0;
      }
#pragma unroll
      for (int kk = 0; kk < LU_PW; ++kk) {
        const int k = k0 + kk;
        if (k < k1) {
          double best = -1.0;
          int bi = k;
#pragma unroll
          for (int sl = 0; sl < SL; ++sl) {
            const int i = lane + 32 * sl;
            const double v = fabs(pr[sl][kk]);
            if (i >= k && i < rp && v > best) {
              best = v;
              bi = i;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) {
              best = ov;
              bi = oi;
            }
          }
          if (lane == 0) {
            piv[kk] = bi;
            if (!(best > 0.0)) bad = 1;
          }
          // pivot row and row k, broadcast from their owners; then the swap
          double rowp[LU_PW], rowk[LU_PW];
          const int slp = bi >> 5, slk = k >> 5;
#pragma unroll
          for (int c = 0; c < LU_PW; ++c) {
            double mp = 0.0, mk = 0.0;
#pragma unroll
            for (int sl = 0; sl < SL; ++sl) {
              if (sl == slp) mp = pr[sl][c];
              if (sl == slk) mk = pr[sl][c];
            }
            rowp[c] = __shfl_sync(0xffffffffu, mp, bi & 31);
            rowk[c] = __shfl_sync(0xffffffffu, mk, k & 31);
          }
          if (bi != k) {
#pragma unroll
            for (int sl = 0; sl < SL; ++sl)
#pragma unroll
              for (int c = 0; c < LU_PW; ++c) {
                if (sl == slk && lane == (k & 31)) pr[sl][c] = rowp[c];
                if (sl == slp && lane == (bi & 31)) pr[sl][c] = rowk[c];
              }
          }
          const double inv = 1.0 / rowp[kk];
#pragma unroll
          for (int sl = 0; sl < SL; ++sl) {
            const int i = lane + 32 * sl;
            if (i > k && i < rp) {
              const double f = pr[sl][kk] * inv;
              pr[sl][kk] = f;
#pragma unroll
              for (int c = kk + 1; c < LU_PW; ++c) pr[sl][c] -= f * rowp[c];
            }
          }
        }
      }
#pragma unroll
      for (int sl = 0; sl < SL; ++sl) {
        const int i = lane + 32 * sl;
        if (i >= k0 && i < rp)
#pragma unroll
          for (int c = 0; c < LU_PW; ++c)
            if (k0 + c < k1) a[i * ld + k0 + c] = pr[sl][c];
      }
    }
    __syncthreads();
    if (bad) {
      if (t == 0) {
        st->singular = 1;
        st->diff = 0.0;
      }
      return;
    }
    // columns right of the panel (right-hand sides included): apply the swaps, then U12 = L11^-1 A12
    for (int c = k1 + t; c < ncol; c += nt) {
      for (int k = k0; k < k1; ++k) {
        const int p = piv[k - k0];
        if (p != k) {
          const double tmp = a[k * ld + c];
          a[k * ld + c] = a[p * ld + c];
          a[p * ld + c] = tmp;
        }
      }
      for (int k = k0; k < k1; ++k) {
        const double u = a[k * ld + c];
        for (int i = k + 1; i < k1; ++i) a[i * ld + c] -= a[i * ld + k] * u;
      }
    }
    __syncthreads();
    // trailing update A22 -= L21 U12
    const int w = ncol - k1, h = rp - k1;
    for (int e = t; e < w * h; e += nt) {
      const int i = k1 + e / w, c = k1 + e % w;
      double v = a[i * ld + c];
#pragma unroll
      for (int k = 0; k < LU_PW; ++k)
        if (k0 + k < k1) v -= a[i * ld + k0 + k] * a[(k0 + k) * ld + c];
      a[i * ld + c] = v;
    }
    __syncthreads();
  }
  // back substitution, one warp per real right-hand side (the padding columns are zero)
  if (warp < d_real) {
    const int c = rp + warp;
    for (int k = rp - 1; k >= 0; --k) {
      double part = 0.0;
      for (int j = k + 1 + lane; j < rp; j += 32) part += a[k * ld + j] * a[j * ld + c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) a[k * ld + c] = (a[k * ld + c] - part) / a[k * ld + k];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = t; i < rp * dp; i += nt) {
    const int k = i / dp, d = i % dp;
    z[i] = d < d_real ? a[k * ld + rp + d] : 0.0;
  }
}

// The same system for the common ranks (a thread's share of the matrix fits 28 registers): Gaussian
// elimination WITHOUT pivoting with the matrix held in registers.  lambda S^-1 + Q^T dP Q is symmetric positive
// definite on the eigenpairs with S > 0, and the rows of (rounding-noise) eigenvalues S < 0 have diagonals
// ~ -lambda / |S| that dominate their row by >= 1e8, so no pivot can be small against its column.  Thread t owns
// column c = t % CW and the rows rg, rg + RG, ... of it; step k publishes row k and column k through
// double-buffered shared memory (one barrier per column instead of a serial panel factorisation), then every
// thread updates its own elements.  Back substitution: one warp per right-hand side, the running right-hand
// side in registers, the solved component broadcast by shuffle, reciprocal diagonal precomputed.
template <int MAXJ>
__global__ void __launch_bounds__(512)
k_cpd_def_solve_reg(const double* __restrict__ t1, const double* __restrict__ t2, const double* __restrict__ S, int rp, int dp,
                    int d_real, double alpha, CpdState* st, double* __restrict__ z) {
  if (!st->active) return;
  extern __shared__ double sm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int CW = rp + d_real, RG = 512 / CW, ld = CW + 1;
  const int c = t % CW, rg = t / CW;
  const bool on = rg < RG;
  double* a = sm;                        // [rp][ld]   (written after the elimination)
  double* rowbuf = a + (size_t)rp * ld;  // [2][CW]
  double* colbuf = rowbuf + 2 * CW;      // [2][rp]
  double* dinv = colbuf + 2 * rp;        // [rp]
  __shared__ int bad;
  const double lam = alpha * st->sigma2;
  double v[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    const int r = rg + RG * j;
    double x = 0.0;
    if (on && r < rp) {
      if (c < rp) {
        x = t1[(size_t)r * rp + c];
        if (r == c) {
          double sv = S[r];
          if (fabs(sv) < 1e-300) sv = sv < 0.0 ? -1e-300 : 1e-300;
          x += lam / sv;
        }
      } else {
        x = t2[(size_t)r * dp + (c - rp)];
      }
    }
    v[j] = x;
  }
  if (t == 0) {
    st->lam = lam;
    bad = 0;
  }
  for (int k = 0; k < rp; ++k) {
    double* rb = rowbuf + (k & 1) * CW;
    double* cb = colbuf + (k & 1) * rp;
    const int jk = k / RG, rgk = k - jk * RG;
    if (on && rg == rgk) {
      double x = 0.0;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j)
        if (j == jk) x = v[j];
      rb[c] = x;
    }
    if (on && c == k) {
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int r = rg + RG * j;
        if (r < rp) cb[r] = v[j];
      }
    }
    __syncthreads();
    if (on && c > k) {
      const double piv = rb[k];
      if (!(fabs(piv) > 0.0)) bad = 1;
      const double s = rb[c] / piv;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int r = rg + RG * j;
        if (r > k && r < rp) v[j] -= cb[r] * s;
      }
    }
  }
  if (on) {
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int r = rg + RG * j;
      if (r < rp) a[r * ld + c] = v[j];
    }
  }
  __syncthreads();
  if (bad) {
    if (t == 0) {
      st->singular = 1;
      st->diff = 0.0;
    }
    return;
  }
  for (int r = t; r < rp; r += 512) dinv[r] = 1.0 / a[r * ld + r];
  __syncthreads();
  if (warp < d_real) {
    constexpr int SL = 5;  // rp <= 160
    const int cr = rp + warp;
    double b[SL];
#pragma unroll
    for (int sl = 0; sl < SL; ++sl) {
      const int i = lane + 32 * sl;
      b[sl] = i < rp ? a[i * ld + cr] : 0.0;
    }
    for (int k = rp - 1; k >= 0; --k) {
      const int slk = k >> 5;
      double mine = 0.0;
#pragma unroll
      for (int sl = 0; sl < SL; ++sl)
        if (sl == slk) mine = b[sl];
      const double zk = __shfl_sync(0xffffffffu, mine, k & 31) * dinv[k];
#pragma unroll
      for (int sl = 0; sl < SL; ++sl) {
        const int i = lane + 32 * sl;
        if (i < k)
          b[sl] -= a[i * ld + k] * zk;
        else if (i == k)
          b[sl] = zk;
      }
    }
#pragma unroll
    for (int sl = 0; sl < SL; ++sl) {
      const int i = lane + 32 * sl;
      if (i < rp) z[(size_t)i * dp + warp] = b[sl];
    }
  }
  for (int i = t; i < rp * dp; i += 512)
    if (i % dp >= d_real) z[i] = 0.0;
}

// W = (F - P1 (Q Z)) / lambda  (padded [M][DP]); one warp per control point, lanes over the rank
template <int D>
__global__ void __launch_bounds__(256)
k_cpd_def_w(const double* __restrict__ f, const double* __restrict__ p1, const double* __restrict__ q, int rp,
            const double* __restrict__ z, int M, int DP, const CpdState* __restrict__ st, double* __restrict__ w) {
  if (!st->active) return;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  double acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.0;
  for (int p = lane; p < rp; p += 32) {
    const double qv = q[(size_t)m * rp + p];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] += qv * z[(size_t)p * DP + d];
  }
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
  if (lane < DP) {
    double v = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d)
      if (lane == d) v = (f[(size_t)m * DP + d] - p1[m] * acc[d]) / st->lam;
    w[(size_t)m * DP + lane] = v;
  }
}

// TY = Y + Q (S .* QtW); one warp per control point
template <int D>
__global__ void __launch_bounds__(256)
k_cpd_def_ty(const double* __restrict__ y, const double* __restrict__ q, int rp, const double* __restrict__ S,
             const double* __restrict__ qtw, int M, int DP, const CpdState* __restrict__ st, double* __restrict__ ty) {
  if (!st->active) return;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  double acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.0;
  for (int p = lane; p < rp; p += 32) {
    const double qv = q[(size_t)m * rp + p] * S[p];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] += qv * qtw[(size_t)p * DP + d];
  }
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
#pragma unroll
  for (int d = 0; d < D; ++d)
    if (lane == d) ty[(size_t)m * D + d] = y[(size_t)m * D + d] + acc[d];
}

// sigma2 update (one CTA): Np, xPx, yPy, trPXY by fixed-tree sums
template <int D>
__global__ void __launch_bounds__(1024)
k_cpd_def_variance(const double* __restrict__ x, int N, const double* __restrict__ pt1, const double* __restrict__ ty, int M,
                   const double* __restrict__ p1, const double* __restrict__ px, CpdState* st) {
  if (!st->active) return;
  __shared__ double red[1024];
  double xpx = 0.0, ypy = 0.0, tr = 0.0, np = 0.0;
  for (int n = threadIdx.x; n < N; n += 1024) {
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double v = x[(size_t)n * D + d];
      s += v * v;
    }
    xpx += pt1[n] * s;
  }
  for (int m = threadIdx.x; m < M; m += 1024) {
    double s = 0.0, u = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double v = ty[(size_t)m * D + d];
      s += v * v;
      u += v * px[(size_t)m * D + d];
    }
    const double p = p1[m];
    ypy += p * s;
    tr += u;
    np += p;
  }
  xpx = block_sum(xpx, red);
  ypy = block_sum(ypy, red);
  tr = block_sum(tr, red);
  np = block_sum(np, red);
  if (threadIdx.x == 0) {
    const double prev = st->sigma2;
    double ns = (xpx - 2.0 * tr + ypy) / (np * D);
    if (ns <= 0.0) ns = st->tolerance / 10.0;
    st->np = np;
    st->sigma2 = ns;
    st->diff = fabs(ns - prev);
  }
}

// out = pts + G(pts, y) W   (thread = one point; y and W staged in shared memory by chunks)
template <int D>
__global__ void __launch_bounds__(CPD_T)
k_cpd_def_apply(const double* __restrict__ pts, int n, const double* __restrict__ y, int M, const double* __restrict__ w, double ninv,
                double* __restrict__ out) {
  __shared__ double s[CPD_ACH * 2 * D];
  const int i = blockIdx.x * CPD_T + threadIdx.x;
  double pv[D], acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    pv[d] = i < n ? pts[(size_t)i * D + d] : 0.0;
    acc[d] = 0.0;
  }
  for (int m0 = 0; m0 < M; m0 += CPD_ACH) {
    const int cnt = min(M - m0, CPD_ACH);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * D; e += CPD_T) {
      s[(e / D) * 2 * D + (e % D)] = y[(size_t)m0 * D + e];
      s[(e / D) * 2 * D + D + (e % D)] = w[(size_t)m0 * D + e];
    }
    __syncthreads();
    for (int m = 0; m < cnt; ++m) {
      const double* sy = s + m * 2 * D;
      double d2 = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double t = pv[d] - sy[d];
        d2 += t * t;
      }
      const double k = exp(ninv * d2);
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] += k * sy[D + d];
    }
  }
  if (i < n)
#pragma unroll
    for (int d = 0; d < D; ++d) out[(size_t)i * D + d] = pv[d] + acc[d];
}

__global__ void k_cpd_unpad(const double* __restrict__ in, int M, int D, int DP, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * D) return;
  out[i] = in[(size_t)(i / D) * DP + (i % D)];
}

// ------------------------------------------------------------------------------------------- host side
#define FB_CPD_DISPATCH(D, MACRO) \
  switch (D) {                    \
    case 1: MACRO(1); break;      \
    case 2: MACRO(2); break;      \
    case 3: MACRO(3); break;      \
    case 4: MACRO(4); break;      \
    case 5: MACRO(5); break;      \
    case 6: MACRO(6); break;      \
    case 7: MACRO(7); break;      \
    case 8: MACRO(8); break;      \
    case 9: MACRO(9); break;      \
    case 10: MACRO(10); break;    \
    case 11: MACRO(11); break;    \
    case 12: MACRO(12); break;    \
    case 13: MACRO(13); break;    \
    case 14: MACRO(14); break;    \
    case 15: MACRO(15); break;    \
    case 16: MACRO(16); break;    \
    default: break;               \
  }

struct CpdWs {
  CpdState* st;
  double *ty, *part_col, *part_row, *inv_den, *pt1, *p1, *px, *mom, *bm, *tv;
  // deformable
  double *G, *X, *Z, *Xr, *Zr, *gram, *gh, *wmat, *theta, *res_part, *res;
  double *Q, *S, *F, *W, *T1, *T2, *Zs, *QtW, *tn_part;
  int rp, rb, dp;
};

static int cpd_rank_block(int num_eig) {  // columns of the subspace iteration
  const int guard = num_eig / 4 > 8 ? num_eig / 4 : 8;
  return (num_eig + guard + 7) / 8 * 8;
}

static size_t cpd_layout(int N, int M, int D, int num_eig, CpdWs* w, void* ws, size_t ws_bytes) {
  Carver cv(ws, ws_bytes);
  const int chm = div_up(M, CPD_CHUNK), chn = div_up(N, CPD_CHUNK);
  CpdWs l;
  memset(&l, 0, sizeof(l));
  l.st = cv.take<CpdState>(1);
  l.ty = cv.take<double>((size_t)M * D);
  l.part_col = cv.take<double>((size_t)chm * N);
  l.part_row = cv.take<double>((size_t)chn * M * (D + 1));
  l.inv_den = cv.take<double>(N);
  l.pt1 = cv.take<double>(N);
  l.p1 = cv.take<double>(M);
  l.px = cv.take<double>((size_t)M * D);
  l.mom = cv.take<double>((size_t)(div_up(M, CPD_MCH) + div_up(N, CPD_MCH)) * (2 * D * D + 1));
  l.bm = cv.take<double>(D * D);
  l.tv = cv.take<double>(D);
  if (num_eig > 0) {
    const int r = num_eig < M ? num_eig : M;
    l.rp = (r + 7) / 8 * 8;
    l.rb = cpd_rank_block(r);
    l.dp = (D + 7) / 8 * 8;
    const int chunks = div_up(M, CPD_TN_ROWS);
    const int big = l.rb > l.rp ? l.rb : l.rp;
    l.G = cv.take<double>((size_t)M * M);
    l.X = cv.take<double>((size_t)M * l.rb);
    l.Z = cv.take<double>((size_t)M * l.rb);
    l.Xr = cv.take<double>((size_t)M * l.rb);
    l.Zr = cv.take<double>((size_t)M * l.rb);
    l.gh = cv.take<double>((size_t)2 * big * big);
    l.wmat = cv.take<double>((size_t)big * big);
    l.theta = cv.take<double>(big);
    l.res_part = cv.take<double>((size_t)chunks * big);
    l.res = cv.take<double>(big);
    l.Q = cv.take<double>((size_t)M * l.rp);
    l.S = cv.take<double>(l.rp);
    l.F = cv.take<double>((size_t)M * l.dp);
    l.W = cv.take<double>((size_t)M * l.dp);
    l.T1 = cv.take<double>((size_t)l.rp * l.rp);
    l.T2 = cv.take<double>((size_t)l.rp * l.dp);
    l.Zs = cv.take<double>((size_t)l.rp * l.dp);
    l.QtW = cv.take<double>((size_t)l.rp * l.dp);
    l.tn_part = cv.take<double>((size_t)chunks * big * big);
  }
  if (w) *w = l;
  return cv.used + 256;
}

static int tn_gram(const double* A, int lda, const double* Bm, int ldb, const double* scale, int rows, int ra, int rb,
                   double* partial, double* out, const CpdState* st, cudaStream_t stream) {
  const int chunks = div_up(rows, CPD_TN_ROWS), nbt = rb / 8;
  if (nbt > 2) {  // wide right operand: 8 column tiles per warp, so each A fragment feeds 8 MMAs
    dim3 grid(chunks, ra / 8, div_up(nbt, 8));
    k_tn_gram<8><<<grid, 128, 0, stream>>>(A, lda, Bm, ldb, scale, rows, ra, rb, partial, st);
  } else {
    dim3 grid(chunks, ra / 8, 1);
    k_tn_gram<2><<<grid, 128, 0, stream>>>(A, lda, Bm, ldb, scale, rows, ra, rb, partial, st);
  }
  k_sum_partials<<<div_up(ra * rb, 256), 256, 0, stream>>>(partial, chunks, ra * rb, out, st);
  FB_COUNT_LAUNCH(2);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

static int gemm_nn(const double* A, int lda, const double* Bm, int ldb, double* Cm, int ldc, int M, int K, int ncols,
                   cudaStream_t stream) {
  dim3 grid(div_up(M, GE_ROWS), div_up(ncols, GE_COLS));
  k_gemm_nn<<<grid, 128, 0, stream>>>(A, lda, Bm, ldb, Cm, ldc, M, K, ncols);
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

static int cpd_init_state(CpdWs& w, const double* x, int N, const double* y, int M, int D, int max_iterations, double tolerance,
                          double outlier_w, cudaStream_t stream) {
  CpdState h;
  memset(&h, 0, sizeof(h));
  h.q = INFINITY;
  h.diff = INFINITY;
  h.tolerance = tolerance;
  h.w = outlier_w;
  h.max_iterations = max_iterations;
  h.active = max_iterations > 0 ? 1 : 0;
  FB_CUDA(cudaMemcpyAsync(w.st, &h, sizeof(h), cudaMemcpyHostToDevice, stream));
  FB_CUDA(cudaStreamSynchronize(stream));  // h is on this stack frame
  FB_CUDA(cudaMemcpyAsync(w.ty, y, sizeof(double) * (size_t)M * D, cudaMemcpyDeviceToDevice, stream));
  const int chm = div_up(M, CPD_CHUNK);
  dim3 grid(div_up(N, CPD_T), chm);
#define FB_L(DD) k_cpd_colsum<DD, 1><<<grid, CPD_T, 0, stream>>>(x, N, y, M, w.st, w.part_col)
  FB_CPD_DISPATCH(D, FB_L)
#undef FB_L
  k_cpd_sigma_init<<<1, 1024, 0, stream>>>(w.part_col, (long long)chm * N, D, M, N, w.st);
  FB_COUNT_LAUNCH(2);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

static int cpd_estep(CpdWs& w, const double* x, int N, int M, int D, cudaStream_t stream) {
  const int chm = div_up(M, CPD_CHUNK), chn = div_up(N, CPD_CHUNK);
  dim3 g1(div_up(N, CPD_T), chm), g2(div_up(M, CPD_T), chn);
#define FB_L(DD)                                                                              \
  k_cpd_colsum<DD, 0><<<g1, CPD_T, 0, stream>>>(x, N, w.ty, M, w.st, w.part_col);             \
  k_cpd_den<<<div_up(N, 256), 256, 0, stream>>>(w.part_col, chm, N, M, D, w.st, w.inv_den, w.pt1); \
  k_cpd_rowsum<DD><<<g2, CPD_T, 0, stream>>>(x, N, w.ty, M, w.inv_den, w.st, w.part_row)
  FB_CPD_DISPATCH(D, FB_L)
#undef FB_L
  k_cpd_rowreduce<<<div_up(M * (D + 1), 256), 256, 0, stream>>>(w.part_row, chn, M, D, w.st, w.p1, w.px);
  FB_COUNT_LAUNCH(4);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

static int cpd_read_state(CpdWs& w, CpdState* h, cudaStream_t stream) {
  FB_CUDA(cudaMemcpyAsync(h, w.st, sizeof(CpdState), cudaMemcpyDeviceToHost, stream));
  FB_CUDA(cudaStreamSynchronize(stream));
  return FB_OK;
}

// leading eigenpairs of G into w.Q / w.S; info3 = {iterations, max residual / |theta_0|, smallest kept |theta|}
static int cpd_low_rank(CpdWs& w, const double* y, int M, int D, double beta, int num_eig, double* info3, cudaStream_t stream) {
  const int r = num_eig < M ? num_eig : M, rp = w.rp, rb = w.rb;
  const double ninv = -0.5 / (beta * beta);
  dim3 gg(div_up(M, 256), M);
#define FB_L(DD) k_gauss_gram<DD><<<gg, 256, 0, stream>>>(y, M, ninv, w.G)
  FB_CPD_DISPATCH(D, FB_L)
#undef FB_L
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  std::vector<double> S_host(rp, 1.0);
  info3[0] = info3[1] = info3[2] = 0.0;
  if (M <= 256 || rb > M) {
    // small problem: the whole matrix goes through the host eigensolver
    std::vector<double> g((size_t)M * M), ev(M);
    FB_CUDA(cudaMemcpyAsync(g.data(), w.G, sizeof(double) * g.size(), cudaMemcpyDeviceToHost, stream));
    FB_CUDA(cudaStreamSynchronize(stream));
    if (eig_sym_host(g.data(), ev.data(), M) != 0) {
      set_error("cpd: the dense eigensolver did not converge");
      return FB_ERR_UNSUPPORTED;
    }
    std::vector<int> order(M);
    for (int i = 0; i < M; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return std::fabs(ev[a]) > std::fabs(ev[b]); });
    std::vector<double> q((size_t)M * rp, 0.0);
    for (int c = 0; c < r; ++c) {
      S_host[c] = ev[order[c]];
      for (int m = 0; m < M; ++m) q[(size_t)m * rp + c] = g[(size_t)m * M + order[c]];
    }
    FB_CUDA(cudaMemcpyAsync(w.Q, q.data(), sizeof(double) * q.size(), cudaMemcpyHostToDevice, stream));
    FB_CUDA(cudaMemcpyAsync(w.S, S_host.data(), sizeof(double) * rp, cudaMemcpyHostToDevice, stream));
    FB_CUDA(cudaStreamSynchronize(stream));
    info3[2] = std::fabs(S_host[r - 1]);
    return FB_OK;
  }
  k_cpd_eig_start<<<div_up((long long)M * rb, 256), 256, 0, stream>>>(w.X, M, rb);
  FB_COUNT_LAUNCH(1);
  std::vector<double> gh((size_t)2 * rb * rb), wm((size_t)rb * rb), th(rb), res(rb);
  const int chunks = div_up(M, CPD_TN_ROWS);
  const int max_it = 40;
  for (int it = 0; it < max_it; ++it) {
    int rc = gemm_nn(w.G, M, w.X, rb, w.Z, rb, M, M, rb, stream);
    if (rc) return rc;
    if ((rc = tn_gram(w.X, rb, w.X, rb, nullptr, M, rb, rb, w.tn_part, w.gh, nullptr, stream))) return rc;
    if ((rc = tn_gram(w.X, rb, w.Z, rb, nullptr, M, rb, rb, w.tn_part, w.gh + (size_t)rb * rb, nullptr, stream))) return rc;
    FB_CUDA(cudaMemcpyAsync(gh.data(), w.gh, sizeof(double) * gh.size(), cudaMemcpyDeviceToHost, stream));
    FB_CUDA(cudaStreamSynchronize(stream));
    double* hm = gh.data() + (size_t)rb * rb;
    for (int i = 0; i < rb; ++i)  // X^T G X is symmetric up to rounding
      for (int j = i + 1; j < rb; ++j)
        hm[(size_t)i * rb + j] = hm[(size_t)j * rb + i] = 0.5 * (hm[(size_t)i * rb + j] + hm[(size_t)j * rb + i]);
    if (rr_leading_host(gh.data(), hm, rb, wm.data(), th.data()) < 0) {
      set_error("cpd: Rayleigh-Ritz of the kernel matrix failed");
      return FB_ERR_UNSUPPORTED;
    }
    FB_CUDA(cudaMemcpyAsync(w.wmat, wm.data(), sizeof(double) * wm.size(), cudaMemcpyHostToDevice, stream));
    FB_CUDA(cudaMemcpyAsync(w.theta, th.data(), sizeof(double) * rb, cudaMemcpyHostToDevice, stream));
    if ((rc = gemm_nn(w.X, rb, w.wmat, rb, w.Xr, rb, M, rb, rb, stream))) return rc;  // Ritz vectors
    if ((rc = gemm_nn(w.Z, rb, w.wmat, rb, w.Zr, rb, M, rb, rb, stream))) return rc;  // G times them
    // residuals, and the next block (one power step on the columns that carry signal) over the old X
    const double thr = 1e-12 * std::fabs(th[0]);
    k_cpd_eig_next<<<chunks, (rb + 31) / 32 * 32, 0, stream>>>(w.Xr, w.Zr, w.theta, M, rb, thr, w.X, w.res_part);
    k_sum_partials<<<div_up(rb, 256), 256, 0, stream>>>(w.res_part, chunks, rb, w.res, nullptr);
    FB_COUNT_LAUNCH(2);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaMemcpyAsync(res.data(), w.res, sizeof(double) * rb, cudaMemcpyDeviceToHost, stream));
    FB_CUDA(cudaStreamSynchronize(stream));
    double worst = 0.0;
    for (int j = 0; j < r; ++j) worst = std::max(worst, std::sqrt(res[j]));
    info3[0] = it + 1;
    info3[1] = worst / std::fabs(th[0]);
    if ((it >= 1 && worst <= 1e-11 * std::fabs(th[0])) || it == max_it - 1) {
      k_cpd_store_q<<<div_up((long long)M * rp, 256), 256, 0, stream>>>(w.Xr, M, rb, r, rp, w.Q);
      FB_COUNT_LAUNCH(1);
      FB_LAUNCH_CHECK();
      for (int c = 0; c < r; ++c) S_host[c] = th[c];
      FB_CUDA(cudaMemcpyAsync(w.S, S_host.data(), sizeof(double) * rp, cudaMemcpyHostToDevice, stream));
      FB_CUDA(cudaStreamSynchronize(stream));
      info3[2] = std::fabs(th[r - 1]);
      break;
    }
  }
  return FB_OK;
}

}  // namespace fb

using namespace fb;

extern "C" {

size_t focusr_cpd_workspace_bytes(int n_x, int n_y, int dim, int num_eig) {
  if (n_x <= 0 || n_y <= 0 || dim <= 0) return 0;
  return cpd_layout(n_x, n_y, dim, num_eig, nullptr, nullptr, 0);
}

int focusr_cpd_affine(const double* x, int n_x, const double* y, int n_y, int dim, int max_iterations, double tolerance,
                      double outlier_w, double* b_out, double* t_out, double* ty_out, double* result_host, void* workspace,
                      size_t workspace_bytes, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_x > 0 && n_y > 0, "cpd: empty point set");
  FB_REQUIRE(dim >= 1 && dim <= CPD_MAXD, "cpd: dimension %d unsupported (1..%d)", dim, CPD_MAXD);
  FB_REQUIRE(outlier_w >= 0.0 && outlier_w < 1.0, "cpd: outlier weight must be in [0, 1)");
  const int N = n_x, M = n_y, D = dim;
  CpdWs w;
  const size_t need = cpd_layout(N, M, D, 0, &w, workspace, workspace_bytes);
  if (need > workspace_bytes) {
    set_error("cpd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return FB_ERR_WORKSPACE;
  }
  int rc = cpd_init_state(w, x, N, y, M, D, max_iterations, tolerance, outlier_w, stream);
  if (rc) return rc;
  std::vector<double> eye((size_t)D * D + D, 0.0);
  for (int i = 0; i < D; ++i) eye[(size_t)i * D + i] = 1.0;
  FB_CUDA(cudaMemcpyAsync(w.bm, eye.data(), sizeof(double) * D * D, cudaMemcpyHostToDevice, stream));
  FB_CUDA(cudaMemcpyAsync(w.tv, eye.data() + (size_t)D * D, sizeof(double) * D, cudaMemcpyHostToDevice, stream));
  FB_CUDA(cudaStreamSynchronize(stream));
  const int chm = div_up(M, CPD_MCH), chn = div_up(N, CPD_MCH);
  CpdState h;
  if ((rc = cpd_read_state(w, &h, stream))) return rc;
  while (h.active) {
    const int batch = std::min(8, h.max_iterations - h.iteration);
    for (int b = 0; b < batch; ++b) {
      if ((rc = cpd_estep(w, x, N, M, D, stream))) return rc;
#define FB_L(DD)                                                                                                         \
  k_cpd_means<DD><<<1, 1024, 0, stream>>>(w.p1, w.px, y, M, w.st);                                                       \
  k_cpd_affine_moments<DD><<<chm + chn, 256, 0, stream>>>(w.p1, w.px, y, M, w.pt1, x, N, w.st, w.mom, chm);              \
  k_cpd_affine_solve<DD><<<1, 512, 0, stream>>>(w.mom, chm + chn, w.st, w.bm, w.tv);                                     \
  k_cpd_affine_apply<DD><<<div_up(M, 256), 256, 0, stream>>>(y, M, w.bm, w.tv, w.ty, w.st)
      FB_CPD_DISPATCH(D, FB_L)
#undef FB_L
      k_cpd_finish<<<1, 1, 0, stream>>>(w.st);
      FB_COUNT_LAUNCH(5);
      FB_LAUNCH_CHECK();
    }
    if ((rc = cpd_read_state(w, &h, stream))) return rc;
  }
  FB_REQUIRE(!h.singular, "cpd affine: singular moment matrix (degenerate point set)");
  if (b_out) FB_CUDA(cudaMemcpyAsync(b_out, w.bm, sizeof(double) * D * D, cudaMemcpyDeviceToDevice, stream));
  if (t_out) FB_CUDA(cudaMemcpyAsync(t_out, w.tv, sizeof(double) * D, cudaMemcpyDeviceToDevice, stream));
  if (ty_out) FB_CUDA(cudaMemcpyAsync(ty_out, w.ty, sizeof(double) * (size_t)M * D, cudaMemcpyDeviceToDevice, stream));
  if (result_host) {
    result_host[0] = h.iteration;
    result_host[1] = h.sigma2;
    result_host[2] = h.q;
    result_host[3] = h.diff;
  }
  return FB_OK;
}

int focusr_cpd_deformable(const double* x, int n_x, const double* y, int n_y, int dim, int max_iterations, double tolerance,
                          double outlier_w, double alpha, double beta, int num_eig, double* w_out, double* ty_out,
                          double* result_host, void* workspace, size_t workspace_bytes, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_x > 0 && n_y > 0, "cpd: empty point set");
  FB_REQUIRE(dim >= 1 && dim <= CPD_MAXD, "cpd: dimension %d unsupported (1..%d)", dim, CPD_MAXD);
  FB_REQUIRE(outlier_w >= 0.0 && outlier_w < 1.0, "cpd: outlier weight must be in [0, 1)");
  FB_REQUIRE(alpha > 0.0 && beta > 0.0, "cpd: alpha and beta must be positive");
  FB_REQUIRE(num_eig >= 1, "cpd: num_eig must be positive");
  const int N = n_x, M = n_y, D = dim;
  const int r = num_eig < M ? num_eig : M;
  FB_REQUIRE((r + 7) / 8 * 8 <= CPD_MAX_RP, "cpd: num_eig %d unsupported (at most %d)", num_eig, CPD_MAX_RP);
  FB_REQUIRE(M <= 65535, "cpd: at most 65535 control points (got %d)", M);
  CpdWs w;
  const size_t need = cpd_layout(N, M, D, num_eig, &w, workspace, workspace_bytes);
  if (need > workspace_bytes) {
    set_error("cpd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return FB_ERR_WORKSPACE;
  }
  int rc = cpd_init_state(w, x, N, y, M, D, max_iterations, tolerance, outlier_w, stream);
  if (rc) return rc;
  double info3[3];
  if ((rc = cpd_low_rank(w, y, M, D, beta, num_eig, info3, stream))) return rc;
  const int rp = w.rp, dp = w.dp;
  FB_CUDA(cudaMemsetAsync(w.W, 0, sizeof(double) * (size_t)M * dp, stream));
  const size_t smem = sizeof(double) * (size_t)rp * (rp + dp + 1);
  FB_CUDA(cudaFuncSetAttribute(k_cpd_def_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // register-resident elimination when every thread's share of the matrix fits 28 registers
  const int cw = rp + D, rgroups = 512 / cw;
  const bool reg_lu = rgroups >= 1 && (rp + rgroups - 1) / rgroups <= 28;
  const size_t smem_reg = sizeof(double) * ((size_t)rp * (cw + 1) + 2 * cw + 3 * rp);
  if (reg_lu)
    FB_CUDA(cudaFuncSetAttribute(k_cpd_def_solve_reg<28>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_reg));
  CpdState h;
  if ((rc = cpd_read_state(w, &h, stream))) return rc;
  while (h.active) {
    const int batch = std::min(8, h.max_iterations - h.iteration);
    for (int b = 0; b < batch; ++b) {
      if ((rc = cpd_estep(w, x, N, M, D, stream))) return rc;
      k_cpd_def_prep<<<div_up(M * dp, 256), 256, 0, stream>>>(w.p1, w.px, y, M, D, dp, w.F, w.st);
      if ((rc = tn_gram(w.Q, rp, w.Q, rp, w.p1, M, rp, rp, w.tn_part, w.T1, w.st, stream))) return rc;
      if ((rc = tn_gram(w.Q, rp, w.F, dp, nullptr, M, rp, dp, w.tn_part, w.T2, w.st, stream))) return rc;
      if (reg_lu)
        k_cpd_def_solve_reg<28><<<1, 512, smem_reg, stream>>>(w.T1, w.T2, w.S, rp, dp, D, alpha, w.st, w.Zs);
      else
        k_cpd_def_solve<<<1, 512, smem, stream>>>(w.T1, w.T2, w.S, rp, dp, D, alpha, w.st, w.Zs);
#define FB_L(DD) k_cpd_def_w<DD><<<div_up(M, 8), 256, 0, stream>>>(w.F, w.p1, w.Q, rp, w.Zs, M, dp, w.st, w.W)
      FB_CPD_DISPATCH(D, FB_L)
#undef FB_L
      if ((rc = tn_gram(w.Q, rp, w.W, dp, nullptr, M, rp, dp, w.tn_part, w.QtW, w.st, stream))) return rc;
#define FB_L(DD)                                                                                      \
  k_cpd_def_ty<DD><<<div_up(M, 8), 256, 0, stream>>>(y, w.Q, rp, w.S, w.QtW, M, dp, w.st, w.ty);       \
  k_cpd_def_variance<DD><<<1, 1024, 0, stream>>>(x, N, w.pt1, w.ty, M, w.p1, w.px, w.st)
      FB_CPD_DISPATCH(D, FB_L)
#undef FB_L
      k_cpd_finish<<<1, 1, 0, stream>>>(w.st);
      FB_COUNT_LAUNCH(6);
      FB_LAUNCH_CHECK();
    }
    if ((rc = cpd_read_state(w, &h, stream))) return rc;
  }
  FB_REQUIRE(!h.singular, "cpd deformable: singular system in the M-step");
  if (w_out) {
    k_cpd_unpad<<<div_up(M * D, 256), 256, 0, stream>>>(w.W, M, D, dp, w_out);
    FB_COUNT_LAUNCH(1);
  }
  if (ty_out) FB_CUDA(cudaMemcpyAsync(ty_out, w.ty, sizeof(double) * (size_t)M * D, cudaMemcpyDeviceToDevice, stream));
  if (result_host) {
    result_host[0] = h.iteration;
    result_host[1] = h.sigma2;
    result_host[2] = h.diff;
    result_host[3] = info3[0];
    result_host[4] = info3[1];
    result_host[5] = info3[2];
  }
  FB_LAUNCH_CHECK();
  return FB_OK;
}

int focusr_cpd_kernel_matrix(const double* y, int n_y, int dim, double beta, double* g_out, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(dim >= 1 && dim <= CPD_MAXD, "cpd: dimension %d unsupported (1..%d)", dim, CPD_MAXD);
  FB_REQUIRE(beta > 0.0 && n_y > 0 && n_y <= 65535, "cpd: bad kernel width or control set size");
  const double ninv = -0.5 / (beta * beta);
  dim3 gg(div_up(n_y, 256), n_y);
#define FB_L(DD) k_gauss_gram<DD><<<gg, 256, 0, stream>>>(y, n_y, ninv, g_out)
  FB_CPD_DISPATCH(dim, FB_L)
#undef FB_L
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

int focusr_cpd_affine_apply(const double* pts, int n, int dim, const double* b, const double* t, double* out,
                            focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(dim >= 1 && dim <= CPD_MAXD, "cpd: dimension %d unsupported (1..%d)", dim, CPD_MAXD);
  if (n <= 0) return FB_OK;
#define FB_L(DD) k_cpd_affine_apply<DD><<<div_up(n, 256), 256, 0, stream>>>(pts, n, b, t, out, nullptr)
  FB_CPD_DISPATCH(dim, FB_L)
#undef FB_L
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

int focusr_cpd_deformable_apply(const double* pts, int n, const double* y, int n_y, int dim, const double* w, double beta,
                                double* out, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(dim >= 1 && dim <= CPD_MAXD, "cpd: dimension %d unsupported (1..%d)", dim, CPD_MAXD);
  FB_REQUIRE(beta > 0.0 && n_y > 0, "cpd: bad kernel width or empty control set");
  if (n <= 0) return FB_OK;
  const double ninv = -0.5 / (beta * beta);
#define FB_L(DD) k_cpd_def_apply<DD><<<div_up(n, CPD_T), CPD_T, 0, stream>>>(pts, n, y, n_y, w, ninv, out)
  FB_CPD_DISPATCH(dim, FB_L)
#undef FB_L
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

}  // extern "C"
