// K3: eigenvector normalisation (reference graph.py:254-257), eigsort's cost matrices
// (eigsort.py:34-41, 162-233), the flip + reorder of eigsort.py:108-122 and the spectral
// coordinates of focusr.py:492-508.
//
// The expensive part of eigsort in the reference is 2 n^2 calls of scipy.stats.wasserstein_distance
// on 5000-sample columns (each call sorts both inputs).  Only 3n distinct columns exist
// (target_i, source_j, -source_j after the log shift), so they are sorted once each -- one CTA,
// bitonic network in shared memory -- and every (i, j, flip) distance is then a merged-CDF
// integral evaluated in parallel with one binary search per sample (same definition as
// scipy/stats/_stats_py.py `_cdf_distance`, p = 1).  c_lambda, min(c, c_f), the n x n assignment
// and the flip list are O(n^2..n^3) with n <= ~70 and stay on the host (pyfocusr_b200/eigsort.py).
#include <float.h>

#include "common.cuh"
#include "eigsort_decide.h"
#include "knn.cuh"
#include "rowops.h"

namespace fb {

// ---------------------------------------------------------------------------------------------
// B2: per (mesh, column): v <- (v - min) / (max - min) - 0.5
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_normalize_cols(double* __restrict__ vecs, int ld, const int* __restrict__ mesh_off,
                 const int* __restrict__ n_cols) {
  const int mesh = blockIdx.y, j = blockIdx.x;
  if (j >= n_cols[mesh]) return;
  const int r0 = mesh_off[mesh], r1 = mesh_off[mesh + 1];
  double mn = DBL_MAX, mx = -DBL_MAX;
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const double v = vecs[(size_t)r * ld + j];
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  __shared__ double smn[256], smx[256];
  smn[threadIdx.x] = mn;
  smx[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      smn[threadIdx.x] = fmin(smn[threadIdx.x], smn[threadIdx.x + o]);
      smx[threadIdx.x] = fmax(smx[threadIdx.x], smx[threadIdx.x + o]);
    }
    __syncthreads();
  }
  mn = smn[0];
  const double ptp = FB_SUB(smx[0], mn);
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const double v = vecs[(size_t)r * ld + j];
    vecs[(size_t)r * ld + j] = FB_SUB(FB_DIV(FB_SUB(v, mn), ptp), 0.5);
  }
}

// C5: new[:, dst[t]] = sign[t] * old[:, src[t]]  (eigsort.py:108-122: flip, then fancy-index copy)
constexpr int MAX_MOVES = 96;
__global__ void __launch_bounds__(128)
k_flip_permute(double* __restrict__ vecs, int ld, const int* __restrict__ mesh_off, const int* __restrict__ dst,
               const int* __restrict__ src, const int* __restrict__ sign, int n_moves) {
  const int mesh = blockIdx.y;
  const int r = mesh_off[mesh] + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= mesh_off[mesh + 1]) return;
  double tmp[MAX_MOVES];
  double* row = vecs + (size_t)r * ld;
  for (int t = 0; t < n_moves; ++t) {
    const double v = row[src[mesh * n_moves + t]];
    tmp[t] = sign[mesh * n_moves + t] < 0 ? FB_MUL(v, -1.0) : v;
  }
  for (int t = 0; t < n_moves; ++t) row[dst[mesh * n_moves + t]] = tmp[t];
}

// D1: out[i][u] = vecs[i][u] * weights[mesh][u]
__global__ void k_spectral_coords(const double* __restrict__ vecs, int ld, const int* __restrict__ mesh_off,
                                  const double* __restrict__ weights, int ns, double* __restrict__ out,
                                  int max_rows) {
  const int mesh = blockIdx.y;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nrows = mesh_off[mesh + 1] - mesh_off[mesh];
  if (t >= (long long)nrows * ns) return;
  const int r = mesh_off[mesh] + (int)(t / ns), u = (int)(t % ns);
  out[(size_t)r * ns + u] = FB_MUL(vecs[(size_t)r * ld + u], weights[mesh * ns + u]);
}

// ---------------------------------------------------------------------------------------------
// eigsort costs
// ---------------------------------------------------------------------------------------------
// C1: sampled xyz normalised per axis to [0,1] over the sample (graph.py:269-272).
// grid (2, n_pairs): blockIdx.x = 0 target, 1 source.  out_t [pair][n_t][3], out_s [pair][n_s][3]
__global__ void __launch_bounds__(256)
k_sample_points(const double* __restrict__ points, const int* __restrict__ mesh_off,
                const int* __restrict__ t_mesh, const int* __restrict__ s_mesh,
                const long long* __restrict__ idx_t, const long long* __restrict__ idx_s, int n_t, int n_s,
                double* __restrict__ out_t, double* __restrict__ out_s) {
  const int pair = blockIdx.y, side = blockIdx.x;
  const int mesh = side == 0 ? t_mesh[pair] : s_mesh[pair];
  const long long* idx = side == 0 ? idx_t + (size_t)pair * n_t : idx_s + (size_t)pair * n_s;
  const int n = side == 0 ? n_t : n_s;
  const size_t base = mesh_off[mesh];
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (int s = threadIdx.x; s < n; s += blockDim.x)
    for (int a = 0; a < 3; ++a) {
      const double v = points[3 * (base + idx[s]) + a];
      mn[a] = fmin(mn[a], v);
      mx[a] = fmax(mx[a], v);
    }
  __shared__ double smn[3][256], smx[3][256];
  for (int a = 0; a < 3; ++a) {
    smn[a][threadIdx.x] = mn[a];
    smx[a][threadIdx.x] = mx[a];
  }
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int a = 0; a < 3; ++a) {
        smn[a][threadIdx.x] = fmin(smn[a][threadIdx.x], smn[a][threadIdx.x + o]);
        smx[a][threadIdx.x] = fmax(smx[a][threadIdx.x], smx[a][threadIdx.x + o]);
      }
    __syncthreads();
  }
  double* dstp = side == 0 ? out_t + (size_t)pair * n_t * 3 : out_s + (size_t)pair * n_s * 3;
  for (int s = threadIdx.x; s < n; s += blockDim.x)
    for (int a = 0; a < 3; ++a) {
      const double v = points[3 * (base + idx[s]) + a];
      dstp[3 * (size_t)s + a] = FB_DIV(FB_SUB(v, smn[a][0]), FB_SUB(smx[a][0], smn[a][0]));
    }
}

__global__ void k_linear_offsets(int* off, int n, int stride) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) off[p] = p * stride;
}

// C3 step 1: column c of pair p (c < n: target_i; c < 2n: source_j; else: -source_j), log-shifted
// (eigsort.py:183-188: log(v + 0.5 + eps), eps = machine epsilon) and sorted ascending.
// One CTA per column; bitonic sort of the padded power-of-two array in shared memory.
__global__ void __launch_bounds__(512)
k_sorted_logcols(const double* __restrict__ vecs, int ld, const int* __restrict__ mesh_off,
                 const int* __restrict__ t_mesh, const int* __restrict__ s_mesh,
                 const long long* __restrict__ idx_t, const long long* __restrict__ idx_s, int n_t, int n_s,
                 int n, int s_max, int p2, double* __restrict__ sorted) {
  extern __shared__ double sh[];
  const int pair = blockIdx.y, c = blockIdx.x;
  const bool is_t = c < n;
  const int col = is_t ? c : (c < 2 * n ? c - n : c - 2 * n);
  const double sgn = c >= 2 * n ? -1.0 : 1.0;
  const int mesh = is_t ? t_mesh[pair] : s_mesh[pair];
  const long long* idx = is_t ? idx_t + (size_t)pair * n_t : idx_s + (size_t)pair * n_s;
  const int cnt = is_t ? n_t : n_s;
  const size_t base = mesh_off[mesh];
  const double eps = 2.220446049250313e-16;
  for (int s = threadIdx.x; s < p2; s += blockDim.x) {
    double v = __longlong_as_double(0x7ff0000000000000LL);
    if (s < cnt) {
      const double e = vecs[(base + idx[s]) * ld + col];
      v = log(FB_ADD(FB_ADD(sgn < 0 ? -e : e, 0.5), eps));
      if (v != v) v = __longlong_as_double(0x7ff0000000000000LL);  // NaN cannot be ordered; keep it last
    }
    sh[s] = v;
  }
  __syncthreads();
  for (int k = 2; k <= p2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < p2; t += blockDim.x) {
        const int ixj = t ^ j;
        if (ixj > t) {
          const double a = sh[t], b = sh[ixj];
          const bool up = (t & k) == 0;
          if ((a > b) == up) {
            sh[t] = b;
            sh[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  double* dstp = sorted + ((size_t)pair * 3 * n + c) * s_max;
  for (int s = threadIdx.x; s < cnt; s += blockDim.x) dstp[s] = sh[s];
}

__device__ __forceinline__ int lower_bound_d(const double* a, int n, double x) {  // first a[i] >= x
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < x)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_bound_d(const double* a, int n, double x) {  // first a[i] > x
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= x)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

// C3 step 2: W1(u, v) = sum over the merged sorted values of |F_u - F_v| * gap  (scipy
// _cdf_distance with p=1).  grid (n*n, 2, n_pairs): blockIdx.x = i*n+j, blockIdx.y = flip.
__global__ void __launch_bounds__(256)
k_wasserstein(const double* __restrict__ sorted, int n, int n_t, int n_s, int s_max,
              double* __restrict__ c_hist, double* __restrict__ c_hist_f) {
  const int pair = blockIdx.z, flip = blockIdx.y, i = blockIdx.x / n, j = blockIdx.x % n;
  const double* u = sorted + ((size_t)pair * 3 * n + i) * s_max;
  const double* v = sorted + ((size_t)pair * 3 * n + (flip ? 2 * n : n) + j) * s_max;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double acc = 0.0;
  for (int t = threadIdx.x; t < n_t + n_s; t += blockDim.x) {
    double x, nxt, fu, fv;
    if (t < n_t) {
      x = u[t];
      const int lb = lower_bound_d(v, n_s, x);
      nxt = fmin(t + 1 < n_t ? u[t + 1] : inf, lb < n_s ? v[lb] : inf);
      fu = (double)(t + 1) / (double)n_t;
      fv = (double)lb / (double)n_s;
    } else {
      const int c = t - n_t;
      x = v[c];
      const int ub = upper_bound_d(u, n_t, x);
      nxt = fmin(c + 1 < n_s ? v[c + 1] : inf, ub < n_t ? u[ub] : inf);
      fu = (double)ub / (double)n_t;
      fv = (double)(c + 1) / (double)n_s;
    }
    if (nxt < inf && nxt > x) acc += fabs(fu - fv) * (nxt - x);
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) (flip ? c_hist_f : c_hist)[(size_t)pair * n * n + i * n + j] = red[0];
}

// C4: c_spatial[i][j] = sqrt(sum_r (s_j[nn[r]] - t_i[r])^2) / n_t, flipped: (-s_j[nn[r]] - t_i[r]).
__global__ void __launch_bounds__(128)
k_spatial(const double* __restrict__ vecs, int ld, const int* __restrict__ mesh_off,
          const int* __restrict__ t_mesh, const int* __restrict__ s_mesh, const long long* __restrict__ idx_t,
          const long long* __restrict__ idx_s, const long long* __restrict__ nn, int n_t, int n_s, int n,
          double* __restrict__ c_spatial, double* __restrict__ c_spatial_f) {
  const int pair = blockIdx.y, i = blockIdx.x / n, j = blockIdx.x % n;
  const size_t bt = mesh_off[t_mesh[pair]], bs = mesh_off[s_mesh[pair]];
  const long long* it = idx_t + (size_t)pair * n_t;
  const long long* is = idx_s + (size_t)pair * n_s;
  const long long* nnp = nn + (size_t)pair * n_t;
  double a = 0.0, af = 0.0;
  for (int r = threadIdx.x; r < n_t; r += blockDim.x) {
    const double tv = vecs[(bt + it[r]) * ld + i];
    const double sv = vecs[(bs + is[nnp[r]]) * ld + j];
    const double d = sv - tv, df = -sv - tv;
    a += d * d;
    af += df * df;
  }
  __shared__ double r0[128], r1[128];
  r0[threadIdx.x] = a;
  r1[threadIdx.x] = af;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      r0[threadIdx.x] += r0[threadIdx.x + o];
      r1[threadIdx.x] += r1[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    c_spatial[(size_t)pair * n * n + i * n + j] = sqrt(r0[0]) / (double)n_t;
    c_spatial_f[(size_t)pair * n * n + i * n + j] = sqrt(r1[0]) / (double)n_t;
  }
}

}  // namespace fb

using namespace fb;

extern "C" {

int focusr_normalize_columns(double* vecs, int n_points, int ld, const int* mesh_point_off, int n_meshes,
                             const int* n_cols, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_points > 0 && ld > 0 && n_meshes > 0, "normalize_columns: empty");
  dim3 grid(ld, n_meshes);
  k_normalize_cols<<<grid, 256, 0, stream>>>(vecs, ld, mesh_point_off, n_cols);
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

int focusr_flip_permute_columns(double* vecs, int n_points, int ld, const int* mesh_point_off, int n_meshes,
                                int max_mesh_points, const int* dst, const int* src, const int* sign,
                                int n_moves, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_moves >= 0 && n_moves <= MAX_MOVES, "flip_permute: at most %d columns", MAX_MOVES);
  if (n_moves == 0) return FB_OK;
  dim3 grid(div_up(max_mesh_points, 128), n_meshes);
  k_flip_permute<<<grid, 128, 0, stream>>>(vecs, ld, mesh_point_off, dst, src, sign, n_moves);
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

int focusr_spectral_coords(const double* vecs, int n_points, int ld, const int* mesh_point_off, int n_meshes,
                           int max_mesh_points, const double* weights, int ns, double* out,
                           focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(ns >= 1 && ns <= ld, "spectral_coords: need 1 <= ns <= ld");
  dim3 grid(div_up((long long)max_mesh_points * ns, 256), n_meshes);
  k_spectral_coords<<<grid, 256, 0, stream>>>(vecs, ld, mesh_point_off, weights, ns, out, max_mesh_points);
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

// C2, C5, D1 on the device: one thread per pair runs eigsort_decide_pair (eigsort_decide.h) and writes the column moves
// of BOTH meshes of the pair (identity for the reference graph) and their spectral weights.
__global__ void __launch_bounds__(64)
k_eigsort_decide(const double* __restrict__ eig_vals, int ldv, const int* __restrict__ n_found, const int* __restrict__ t_mesh,
                 const int* __restrict__ s_mesh, int n_pairs, const double* __restrict__ c_hist,
                 const double* __restrict__ c_hist_f, const double* __restrict__ c_spatial,
                 const double* __restrict__ c_spatial_f, int n, int ns, int target_as_reference, int weighted,
                 double* __restrict__ q_out, int* __restrict__ dst, int* __restrict__ src, int* __restrict__ sign,
                 double* __restrict__ weights, double* __restrict__ scratch, int* __restrict__ status) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int mt = t_mesh[p], ms = s_mesh[p];
  const size_t nn = (size_t)n * n;
  int d[LSAP_MAX], s[LSAP_MAX], g[LSAP_MAX];
  double w[LSAP_MAX];
  const int rc = eigsort_decide_pair(eig_vals + (size_t)mt * ldv, n_found[mt], eig_vals + (size_t)ms * ldv, n_found[ms],
                                     c_hist + p * nn, c_hist_f + p * nn, c_spatial + p * nn, c_spatial_f + p * nn, n, ns,
                                     target_as_reference != 0, weighted != 0, q_out + (size_t)p * n, d, s, g, w,
                                     scratch + (size_t)p * 2 * nn);
  if (rc != 0) atomicExch(status, 1);
  const int moved = target_as_reference ? ms : mt, fixed = target_as_reference ? mt : ms;
  for (int k = 0; k < n; ++k) {
    dst[(size_t)moved * n + k] = rc == 0 ? d[k] : k;
    src[(size_t)moved * n + k] = rc == 0 ? s[k] : k;
    sign[(size_t)moved * n + k] = rc == 0 ? g[k] : 1;
    dst[(size_t)fixed * n + k] = k;
    src[(size_t)fixed * n + k] = k;
    sign[(size_t)fixed * n + k] = 1;
  }
  for (int k = 0; k < ns; ++k) weights[(size_t)mt * ns + k] = weights[(size_t)ms * ns + k] = rc == 0 ? w[k] : 1.0;
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

size_t focusr_eigsort_workspace_bytes(int n_pairs, int n_samp_t, int n_samp_s, int n_features) {
  const int s_max = n_samp_t > n_samp_s ? n_samp_t : n_samp_s;
  size_t b = 0;
  b += align_up(sizeof(double) * (size_t)n_pairs * n_samp_t * 3);
  b += align_up(sizeof(double) * (size_t)n_pairs * n_samp_s * 3);
  b += align_up(sizeof(double) * (size_t)n_pairs * 3 * n_features * s_max);  // sorted columns
  b += align_up(sizeof(int) * ((size_t)n_pairs + 1)) * 2;                    // segment offsets
  b += knn_pruned_workspace_bytes((long long)n_pairs * n_samp_s, (long long)n_pairs * n_samp_t, n_pairs, 3);
  return b + 1024;
}

int focusr_eigsort_costs(const double* vecs, int ld, const double* points, const int* mesh_point_off,
                         const int* t_mesh, const int* s_mesh, int n_pairs, const long long* idx_t,
                         const long long* idx_s, int n_samp_t, int n_samp_s, int n_features, double* c_hist,
                         double* c_hist_f, double* c_spatial, double* c_spatial_f, long long* nn_idx,
                         void* workspace, size_t workspace_bytes, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_pairs > 0 && n_samp_t > 0 && n_samp_s > 0 && n_features > 0 && n_features <= ld,
             "eigsort_costs: bad sizes");
  const int n = n_features;
  const int s_max = n_samp_t > n_samp_s ? n_samp_t : n_samp_s;
  const int p2 = next_pow2(s_max);
  FB_REQUIRE(p2 <= 16384, "eigsort_costs: at most 16384 ordering samples per mesh (got %d)", s_max);
  FB_REQUIRE((long long)n * n <= 65535 && n_pairs <= 65535, "eigsort_costs: grid too large");
  Carver cv(workspace, workspace_bytes);
  double* pts_t = cv.take<double>((size_t)n_pairs * n_samp_t * 3);
  double* pts_s = cv.take<double>((size_t)n_pairs * n_samp_s * 3);
  double* sorted = cv.take<double>((size_t)n_pairs * 3 * n * s_max);
  int* q_off = cv.take<int>((size_t)n_pairs + 1);
  int* r_off = cv.take<int>((size_t)n_pairs + 1);
  const size_t knn_bytes = knn_pruned_workspace_bytes((long long)n_pairs * n_samp_s, (long long)n_pairs * n_samp_t, n_pairs, 3);
  char* knn_ws = cv.take<char>(knn_bytes);
  if (!cv.ok()) {
    set_error("eigsort_costs: workspace too small (%zu < %zu)", workspace_bytes, cv.used);
    return FB_ERR_WORKSPACE;
  }
  k_sample_points<<<dim3(2, n_pairs), 256, 0, stream>>>(points, mesh_point_off, t_mesh, s_mesh, idx_t, idx_s,
                                                        n_samp_t, n_samp_s, pts_t, pts_s);
  k_linear_offsets<<<div_up(n_pairs + 1, 256), 256, 0, stream>>>(q_off, n_pairs + 1, n_samp_t);
  k_linear_offsets<<<div_up(n_pairs + 1, 256), 256, 0, stream>>>(r_off, n_pairs + 1, n_samp_s);
  FB_COUNT_LAUNCH(3);
  FB_LAUNCH_CHECK();
  // nearest sampled source point of every sampled target point (eigsort.py:203-204)
  int rc;
  if (knn_pruned_applicable(n_samp_s, n_samp_t, 3, 1))
    rc = launch_knn_pruned(pts_s, 3, r_off, pts_t, 3, q_off, n_pairs, n_samp_s, n_samp_t, (long long)n_pairs * n_samp_s,
                           (long long)n_pairs * n_samp_t, 3, 1, nn_idx, nullptr, knn_ws, knn_bytes, stream);
  else
    rc = launch_knn(pts_s, 3, r_off, pts_t, 3, q_off, n_pairs, n_samp_t, 3, 1, nn_idx, nullptr, stream);
  if (rc) return rc;
  const size_t smem = sizeof(double) * (size_t)p2;
  // the opt-in is per device and per function: set whenever it is needed (cheap), never cached process-wide
  if (smem > 48 * 1024) FB_CUDA(cudaFuncSetAttribute(k_sorted_logcols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_sorted_logcols<<<dim3(3 * n, n_pairs), 512, smem, stream>>>(vecs, ld, mesh_point_off, t_mesh, s_mesh, idx_t,
                                                                idx_s, n_samp_t, n_samp_s, n, s_max, p2, sorted);
  k_wasserstein<<<dim3(n * n, 2, n_pairs), 256, 0, stream>>>(sorted, n, n_samp_t, n_samp_s, s_max, c_hist, c_hist_f);
  k_spatial<<<dim3(n * n, n_pairs), 128, 0, stream>>>(vecs, ld, mesh_point_off, t_mesh, s_mesh, idx_t, idx_s, nn_idx,
                                                      n_samp_t, n_samp_s, n, c_spatial, c_spatial_f);
  FB_COUNT_LAUNCH(3);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

size_t focusr_eigsort_decide_workspace_bytes(int n_pairs, int n_features) {
  return sizeof(double) * 2 * (size_t)n_pairs * n_features * n_features + 256;
}

int focusr_eigsort_decide(const double* eig_vals, int ldv, const int* n_found, const int* t_mesh, const int* s_mesh,
                          int n_pairs, const double* c_hist, const double* c_hist_f, const double* c_spatial,
                          const double* c_spatial_f, int n_features, int ns, int target_as_reference, int weighted,
                          double* q_out, int* dst, int* src, int* sign, double* weights, int* status, void* workspace,
                          size_t workspace_bytes, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_pairs > 0 && n_features >= 1 && n_features <= LSAP_MAX && ns >= 1 && ns <= n_features && n_features <= ldv,
             "eigsort_decide: need 1 <= ns <= n_features <= %d", LSAP_MAX);
  FB_REQUIRE(workspace_bytes >= focusr_eigsort_decide_workspace_bytes(n_pairs, n_features), "eigsort_decide: workspace too small");
  FB_CUDA(cudaMemsetAsync(status, 0, sizeof(int), stream));
  k_eigsort_decide<<<div_up(n_pairs, 64), 64, 0, stream>>>(eig_vals, ldv, n_found, t_mesh, s_mesh, n_pairs, c_hist, c_hist_f,
                                                           c_spatial, c_spatial_f, n_features, ns, target_as_reference, weighted,
                                                           q_out, dst, src, sign, weights, static_cast<double*>(workspace), status);
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  return FB_OK;
}
}
