// K1: mesh Laplacian assembly as count / scan / fill / per-row sort+unique / scan / compact.
//
// Replaces the reference's Python loop over cells x edges filling a scipy lil_matrix
// (graph.py:148-178), A.sum(axis=1) and (d+1e-8)**-1 (graph.py:216-219) and, on request, the
// CSR product G @ (D - A) (graph.py:221-226).  Everything is integer/byte traffic plus one
// sqrt/div per stored entry: HBM-bound, so the kernels are plain coalesced streaming passes.
// Bit-exactness with numpy comes from rowops.h: no FMA contraction
// in the edge weight, sequential ascending-column degree sums.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"
#include "rowops.h"

namespace fb {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// exclusive scan (int32), three passes: block totals, scan of totals, block scan + offset
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

size_t scan_tmp_ints(int n) { return (size_t)div_up(n, SCAN_TILE) + 1; }

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  // exclusive scan of one int per thread over a 256-thread block
  __shared__ int warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;  // inclusive over warps
  }
  __syncthreads();
  const int warp_off = wid > 0 ? warp_sums[wid - 1] : 0;
  *total = warp_sums[SCAN_THREADS / 32 - 1];
  const int r = warp_off + inc - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_totals(const int* __restrict__ in, int n,
                                                              int* __restrict__ totals) {
  const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += in[base + i];
  int total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) totals[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_of_totals(int* totals, int nb) {
  // single block: each thread owns a contiguous chunk
  __shared__ int part[1024];
  const int chunk = (nb + 1023) / 1024;
  const int b0 = threadIdx.x * chunk, b1 = min(nb, b0 + chunk);
  int s = 0;
  for (int i = b0; i < b1; ++i) s += totals[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < 1024; ++i) {
      const int t = part[i];
      part[i] = run;
      run += t;
    }
  }
  __syncthreads();
  int run = part[threadIdx.x];
  for (int i = b0; i < b1; ++i) {
    const int t = totals[i];
    totals[i] = run;
    run += t;
  }
  if (b1 == nb && b0 < b1) totals[nb] = run;  // grand total
  if (nb == 0 && threadIdx.x == 0) totals[0] = 0;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int* __restrict__ in, int n,
                                                             const int* __restrict__ totals, int nb,
                                                             int* __restrict__ out) {
  const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int total;
  int run = block_exclusive_scan(s, &total) + totals[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = totals[nb];
}

int exclusive_scan_i32(const int* in, int* out, int n, int* tmp, cudaStream_t stream) {
  const int nb = div_up(n, SCAN_TILE);
  if (n <= 0) {
    FB_CUDA(cudaMemsetAsync(out, 0, sizeof(int), stream));
    return FB_OK;
  }
  k_scan_totals<<<nb, SCAN_THREADS, 0, stream>>>(in, n, tmp);
  k_scan_of_totals<<<1, 1024, 0, stream>>>(tmp, nb);
  k_scan_apply<<<nb, SCAN_THREADS, 0, stream>>>(in, n, tmp, nb, out);
  FB_COUNT_LAUNCH(3);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

// ---------------------------------------------------------------------------------------------
// adjacency build
// ---------------------------------------------------------------------------------------------
__global__ void k_edge_count(const int* __restrict__ tris, int n_edges, int n_points,
                             int* __restrict__ cnt, int* __restrict__ bad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int f = e / 3, k = e - 3 * f;
  const int p1 = tris[3 * f + k];
  const int p2 = tris[3 * f + (k == 2 ? 0 : k + 1)];
  if ((unsigned)p1 >= (unsigned)n_points || (unsigned)p2 >= (unsigned)n_points) {
    atomicAdd(bad, 1);
    return;
  }
  atomicAdd(&cnt[p1], 1);
}

__global__ void k_edge_fill(const int* __restrict__ tris, int n_edges, int n_points,
                            const int* __restrict__ start, int* __restrict__ cursor,
                            int* __restrict__ raw_cols) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int f = e / 3, k = e - 3 * f;
  const int p1 = tris[3 * f + k];
  const int p2 = tris[3 * f + (k == 2 ? 0 : k + 1)];
  if ((unsigned)p1 >= (unsigned)n_points || (unsigned)p2 >= (unsigned)n_points) return;
  const int slot = start[p1] + atomicAdd(&cursor[p1], 1);
  raw_cols[slot] = p2;
}

// One thread per row: sort the row's raw columns (the fill order is non-deterministic, the sorted order is
// not), collapse duplicates (assignment semantics of graph.py:178: a repeated directed edge carries the identical
// weight), compute every kept edge weight ONCE and the degree as the sequential ascending-column sum of the weights
// (= scipy's A.sum(axis=1) bit for bit), and write the sorted row back in place.  Rows of up to 16 entries -- every
// row of a triangle mesh in practice -- are sorted in registers by an odd-even transposition network; longer rows
// fall back to an insertion sort in global memory.  The row keeps its slot [start[i], start[i+1]): when no row of
// the batch had a duplicate (the common case: `dups` stays 0) that slot layout already IS the final CSR.
template <int L>
__device__ __forceinline__ void sort_network(int (&c)[L]) {
#pragma unroll
  for (int round = 0; round < L; ++round) {
#pragma unroll
    for (int a = (round & 1); a + 1 < L; a += 2) {
      const int lo = min(c[a], c[a + 1]), hi = max(c[a], c[a + 1]);
      c[a] = lo;
      c[a + 1] = hi;
    }
  }
}

template <int L>
__device__ __forceinline__ int row_finish_small(const double* __restrict__ points, int pd, int i, int* __restrict__ crow,
                                                double* __restrict__ wrow, int len, double* deg_out) {
  int c[L];
#pragma unroll
  for (int a = 0; a < L; ++a) c[a] = a < len ? crow[a] : 0x7fffffff;
  sort_network<L>(c);
  const double* pi = points + (size_t)pd * i;
  int u = 0, last = -1;
  double d = 0.0;
#pragma unroll
  for (int a = 0; a < L; ++a) {
    const int v = c[a];
    if (a < len && v != last) {
      const double w = edge_weight(pi, points + (size_t)pd * v, pd);
      crow[u] = v;
      wrow[u] = w;
      d = FB_ADD(d, w);
      ++u;
      last = v;
    }
  }
  *deg_out = d;
  return u;
}

__global__ void __launch_bounds__(256)
k_row_finish(const double* __restrict__ points, int pd, int n_points, const int* __restrict__ start,
             int* __restrict__ cols, double* __restrict__ weights, int* __restrict__ ucnt, double* __restrict__ degree,
             double* __restrict__ degree_inv, int* __restrict__ dups) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const int s = start[i], len = start[i + 1] - s;
  int* c = cols + s;
  double* wr = weights + s;
  double d = 0.0;
  int u;
  if (len <= 8) {
    u = row_finish_small<8>(points, pd, i, c, wr, len, &d);
  } else if (len <= 16) {
    u = row_finish_small<16>(points, pd, i, c, wr, len, &d);
  } else {
    for (int a = 1; a < len; ++a) {
      const int v = c[a];
      int b = a - 1;
      while (b >= 0 && c[b] > v) {
        c[b + 1] = c[b];
        --b;
      }
      c[b + 1] = v;
    }
    u = 0;
    const double* pi = points + (size_t)pd * i;
    for (int a = 0; a < len; ++a) {
      const int v = c[a];
      if (a > 0 && v == c[u - 1]) continue;
      const double w = edge_weight(pi, points + (size_t)pd * v, pd);
      c[u] = v;
      wr[u] = w;
      d = FB_ADD(d, w);
      ++u;
    }
  }
  ucnt[i] = u;
  degree[i] = d;
  degree_inv[i] = degree_inverse(d);
  if (u < len) atomicAdd(dups, len - u);
}

// duplicate directed edges exist (rare: inconsistent / non-manifold input): rows move left to their final offsets
__global__ void k_copy_ints(const int* __restrict__ in, int* __restrict__ out, long long n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = in[t];
}

__global__ void k_compact(const double* __restrict__ points, int pd, int n_points,
                          const int* __restrict__ start, const int* __restrict__ raw_cols,
                          const int* __restrict__ row_ptr, int* __restrict__ cols,
                          double* __restrict__ weights) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const int s = start[i], o = row_ptr[i], u = row_ptr[i + 1] - o;
  const double* pi = points + (size_t)pd * i;
  for (int a = 0; a < u; ++a) {
    const int v = raw_cols[s + a];
    cols[o + a] = v;
    weights[o + a] = edge_weight(pi, points + (size_t)pd * v, pd);
  }
}

__device__ __forceinline__ int find_mesh(const int* __restrict__ off, int n_meshes, int row) {
  int lo = 0, hi = n_meshes - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= row)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// per-mesh structure facts the solver needs: one-way entries (A_ij without A_ji, which make L
// non-normal: SURVEY.md section 7.3-1), zero-degree rows (exact null vectors), non-finite weights.
// Rows are read in the slot layout k_row_finish leaves: row i = [row_ptr[i], row_ptr[i] + ucnt[i]).
__global__ void k_mesh_stats(const int* __restrict__ row_ptr, const int* __restrict__ ucnt, const int* __restrict__ cols,
                             const double* __restrict__ weights, int n_points,
                             const int* __restrict__ mesh_off, int n_meshes,
                             int* __restrict__ mesh_info) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const int m = find_mesh(mesh_off, n_meshes, i);
  const int s = row_ptr[i], e = s + ucnt[i];
  int oneway = 0, nonfinite = 0;
  for (int p = s; p < e; ++p) {
    const int j = cols[p];
    if (!isfinite(weights[p])) ++nonfinite;
    int lo = row_ptr[j], hi = lo + ucnt[j] - 1;
    bool found = false;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const int c = cols[mid];
      if (c == i) {
        found = true;
        break;
      }
      if (c < i)
        lo = mid + 1;
      else
        hi = mid - 1;
    }
    oneway += !found;
  }
  constexpr int S = FOCUSR_MESH_INFO_INTS;
  if (oneway) atomicAdd(&mesh_info[S * m + 1], oneway);
  if (e == s) atomicAdd(&mesh_info[S * m + 2], 1);
  if (nonfinite) atomicAdd(&mesh_info[S * m + 3], nonfinite);
  // longest row and entry count of the mesh: one atomic per warp unless the warp straddles meshes
  int len = e - s, total = e - s;
  const unsigned act = __activemask();
  bool warp_wide = false;
  if (act == 0xffffffffu) {
    const int m_lo = __shfl_sync(act, m, 0);
    warp_wide = __all_sync(act, m == m_lo);
  }
  if (warp_wide) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
      total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMax(&mesh_info[S * m + 4], len);
      atomicAdd(&mesh_info[S * m + 0], total);
    }
  } else {
    atomicMax(&mesh_info[S * m + 4], len);
    if (total) atomicAdd(&mesh_info[S * m + 0], total);
  }
}

// ---------------------------------------------------------------------------------------------
// L = D~^-1 (D - A) materialised (sorted columns, explicit zeros dropped)
// ---------------------------------------------------------------------------------------------
__global__ void k_lap_count(const int* __restrict__ row_ptr, const int* __restrict__ cols,
                            const double* __restrict__ weights, const double* __restrict__ degree,
                            const double* __restrict__ degree_inv, int n_points,
                            int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const double gi = degree_inv[i];
  double diag = degree[i];
  int c = 0;
  for (int p = row_ptr[i]; p < row_ptr[i + 1]; ++p) {
    if (cols[p] == i)
      diag = FB_SUB(diag, weights[p]);
    else
      c += FB_MUL(gi, -weights[p]) != 0.0;
  }
  c += FB_MUL(gi, diag) != 0.0;
  cnt[i] = c;
}

__global__ void k_lap_fill(const int* __restrict__ row_ptr, const int* __restrict__ cols,
                           const double* __restrict__ weights, const double* __restrict__ degree,
                           const double* __restrict__ degree_inv, int n_points,
                           const int* __restrict__ l_row_ptr, int* __restrict__ l_cols,
                           double* __restrict__ l_vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const double gi = degree_inv[i];
  double diag = degree[i];
  for (int p = row_ptr[i]; p < row_ptr[i + 1]; ++p)
    if (cols[p] == i) diag = FB_SUB(diag, weights[p]);
  const double dval = FB_MUL(gi, diag);
  int o = l_row_ptr[i];
  bool diag_done = false;
  for (int p = row_ptr[i]; p < row_ptr[i + 1]; ++p) {
    const int j = cols[p];
    if (!diag_done && j >= i) {
      if (dval != 0.0) {
        l_cols[o] = i;
        l_vals[o++] = dval;
      }
      diag_done = true;
    }
    if (j == i) continue;
    const double v = FB_MUL(gi, -weights[p]);
    if (v != 0.0) {
      l_cols[o] = j;
      l_vals[o++] = v;
    }
  }
  if (!diag_done && dval != 0.0) {
    l_cols[o] = i;
    l_vals[o++] = dval;
  }
}

}  // namespace fb

using namespace fb;

extern "C" {

const char* focusr_last_error(void) { return fb::g_err; }
int focusr_version(void) { return 100; }
unsigned long long focusr_launch_count(void) { return fb::g_launch_count.load(); }

size_t focusr_laplacian_workspace_bytes(int n_points, int n_tris) {
  size_t b = 0;
  b += align_up(sizeof(int) * ((size_t)n_points + 1)) * 4;  // cnt, start, cursor, ucnt
  b += align_up(sizeof(int) * (size_t)3 * n_tris);          // raw_cols
  b += align_up(sizeof(int) * scan_tmp_ints(n_points + 1));
  b += align_up(sizeof(int) * 4);
  return b + 1024;
}

int focusr_laplacian_build(const double* points, int point_dim, const int* tris, int n_points, int n_tris,
                           const int* mesh_point_off, int n_meshes, int* row_ptr, int* cols,
                           double* weights, double* degree, double* degree_inv, int* mesh_info,
                           void* workspace, size_t workspace_bytes, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_points > 0 && n_tris >= 0 && n_meshes > 0, "laplacian_build: empty input");
  FB_REQUIRE(point_dim >= 1 && point_dim <= 64, "laplacian_build: point_dim must be in [1, 64]");
  FB_REQUIRE((long long)3 * n_tris < 2147483647LL, "laplacian_build: too many triangles for int32 CSR");
  Carver cv(workspace, workspace_bytes);
  int* cnt = cv.take<int>((size_t)n_points + 1);
  int* start = cv.take<int>((size_t)n_points + 1);
  int* cursor = cv.take<int>((size_t)n_points + 1);
  int* ucnt = cv.take<int>((size_t)n_points + 1);
  int* raw_cols = cv.take<int>((size_t)3 * n_tris);
  int* scan_tmp = cv.take<int>(scan_tmp_ints(n_points + 1));
  int* bad = cv.take<int>(4);
  if (!cv.ok()) {
    set_error("laplacian_build: workspace too small (%zu < %zu)", workspace_bytes, cv.used);
    return FB_ERR_WORKSPACE;
  }
  const int n_edges = 3 * n_tris;
  const int T = 256;
  FB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)n_points + 1), stream));
  FB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)n_points + 1), stream));
  FB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int) * 4, stream));
  FB_CUDA(cudaMemsetAsync(mesh_info, 0, sizeof(int) * FOCUSR_MESH_INFO_INTS * (size_t)n_meshes, stream));
  if (n_edges > 0) {
    k_edge_count<<<div_up(n_edges, T), T, 0, stream>>>(tris, n_edges, n_points, cnt, bad);
    FB_COUNT_LAUNCH(1);
  }
  // the slot of row i, [row_ptr[i], row_ptr[i+1]), holds its raw directed edges; it is final unless duplicates exist
  int rc = exclusive_scan_i32(cnt, row_ptr, n_points, scan_tmp, stream);
  if (rc) return rc;
  if (n_edges > 0) {
    k_edge_fill<<<div_up(n_edges, T), T, 0, stream>>>(tris, n_edges, n_points, row_ptr, cursor, cols);
    FB_COUNT_LAUNCH(1);
  }
  k_row_finish<<<div_up(n_points, T), T, 0, stream>>>(points, point_dim, n_points, row_ptr, cols, weights, ucnt, degree,
                                                      degree_inv, bad + 1);
  k_mesh_stats<<<div_up(n_points, T), T, 0, stream>>>(row_ptr, ucnt, cols, weights, n_points, mesh_point_off, n_meshes,
                                                      mesh_info);
  FB_COUNT_LAUNCH(2);
  FB_LAUNCH_CHECK();
  int bad_host[2] = {0, 0};
  FB_CUDA(cudaMemcpyAsync(bad_host, bad, sizeof(bad_host), cudaMemcpyDeviceToHost, stream));
  FB_CUDA(cudaStreamSynchronize(stream));
  FB_REQUIRE(bad_host[0] == 0, "laplacian_build: %d triangle corner(s) index outside [0, n_points)", bad_host[0]);
  if (bad_host[1] > 0) {
    // duplicates were collapsed: close the gaps (sorted unique columns -> raw_cols at the old offsets -> final offsets,
    // weights recomputed: the same products as in k_row_finish)
    FB_CUDA(cudaMemcpyAsync(start, row_ptr, sizeof(int) * ((size_t)n_points + 1), cudaMemcpyDeviceToDevice, stream));
    k_copy_ints<<<div_up(n_edges, T), T, 0, stream>>>(cols, raw_cols, n_edges);
    rc = exclusive_scan_i32(ucnt, row_ptr, n_points, scan_tmp, stream);
    if (rc) return rc;
    k_compact<<<div_up(n_points, T), T, 0, stream>>>(points, point_dim, n_points, start, raw_cols, row_ptr, cols, weights);
    FB_COUNT_LAUNCH(2);
    FB_LAUNCH_CHECK();
  }
  return FB_OK;
}

int focusr_laplacian_csr(const int* row_ptr, const int* cols, const double* weights,
                         const double* degree, const double* degree_inv, int n_points,
                         int* l_row_ptr, int* l_cols, double* l_vals, void* workspace,
                         size_t workspace_bytes, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_points > 0, "laplacian_csr: empty input");
  Carver cv(workspace, workspace_bytes);
  int* cnt = cv.take<int>((size_t)n_points + 1);
  int* scan_tmp = cv.take<int>(scan_tmp_ints(n_points + 1));
  if (!cv.ok()) {
    set_error("laplacian_csr: workspace too small (%zu < %zu)", workspace_bytes, cv.used);
    return FB_ERR_WORKSPACE;
  }
  const int T = 256;
  k_lap_count<<<div_up(n_points, T), T, 0, stream>>>(row_ptr, cols, weights, degree, degree_inv, n_points, cnt);
  FB_COUNT_LAUNCH(1);
  int rc = exclusive_scan_i32(cnt, l_row_ptr, n_points, scan_tmp, stream);
  if (rc) return rc;
  k_lap_fill<<<div_up(n_points, T), T, 0, stream>>>(row_ptr, cols, weights, degree, degree_inv, n_points,
                                                    l_row_ptr, l_cols, l_vals);
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  return FB_OK;
}
}
