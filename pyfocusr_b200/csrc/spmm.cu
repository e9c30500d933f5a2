// CSR x dense-block kernels: the hot loop of the eigen-solver (one fused SpMM + Chebyshev
// three-term update per filter degree) and the graph smoothing of Graph.mean_filter_graph.
//
// Layout: the block of b vectors is row-major [n_points][b], so gathering neighbour j's row is one
// contiguous 8*b-byte read (128 B for b = 16 = one L2 line) and the matrix (int32 column + fp64
// weight per stored entry) is streamed exactly once per pass.  TPR threads cooperate on a row,
// each owning b/(2*TPR) double2 slices; all TPR threads read the same (col, weight) entry, which
// the LSU serves as one broadcast.  Algorithmic bytes per filter degree and mesh:
// 12*nnz + 16*N (degree, 1/degree~) + 3 * 8*b*N (read Y, read X_prev, write X_next).
//
// These paths are HBM/L2-bandwidth-bound fp64 work with ~0.3 flop/byte; there is no GEMM shape
// here, so no tensor cores (see DESIGN.md "what bounds each kernel").
#include "common.cuh"
#include "rowops.h"
#include "spmm.cuh"

namespace fb {

constexpr int SPMM_THREADS = 256;
constexpr int SPMM_ROWS_PER_BLOCK = 256;

// MODE 0: out = alpha * (L y - c y) - gamma * x_prev        (Chebyshev step)
// MODE 1: out = (D - A) y
// MODE 2: out = L y = dinv * (D - A) y
//
// Tuning record (tools/spmm_bench.py, profiles/r1_summary.md): the kernel is latency-bound on the
// dependent chain row_ptr -> (col, weight) -> gather.  What helps is MORE RESIDENT WARPS: capping
// registers at 32 (8 CTAs = 64 warps per SM) took the filter step from 4151 to 4288 GB/s.  What
// does not help on B200: staging the CSR entries in shared memory (occupancy drops to 16 warps,
// 1.8 TB/s), fetching 8 entries then issuing 8 gathers back to back (2.9 TB/s), or loading the
// warp's entry range with one coalesced load and distributing it by shuffles (3.1 TB/s), bringing the
// CTA's window of y rows into shared memory with one TMA bulk copy (k_spmm_staged below: 3.7 TB/s in the
// filter, 4 CTAs/SM), or blocking the filter over groups of meshes that fit the 126 MB L2 (eigs.cu).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int B, int TPR, int MODE>
__global__ void __launch_bounds__(SPMM_THREADS, (B / (2 * TPR) == 1) ? 8 : ((B / (2 * TPR) == 2) ? 6 : 3))
k_spmm(const int* __restrict__ row_ptr, const int* __restrict__ cols, const double* __restrict__ weights,
       const double* __restrict__ degree, const double* __restrict__ degree_inv,
       const int* __restrict__ mesh_off, const double* __restrict__ y, const double* __restrict__ x_prev,
       double* __restrict__ out, const double* __restrict__ alpha, const double* __restrict__ gamma,
       const double* __restrict__ center, int step, int n_steps, int rows_per_block) {
  constexpr int VPT = B / (2 * TPR);
  static_assert(VPT * 2 * TPR == B, "block size must be a multiple of 2*TPR");
  const int mesh = blockIdx.y;
  const int r0 = mesh_off[mesh] + blockIdx.x * rows_per_block;
  const int r1 = min(mesh_off[mesh + 1], r0 + rows_per_block);
  if (r0 >= r1) return;
  double al = 1.0, ga = 0.0, cc = 0.0;
  if (MODE == 0) {
    al = alpha[(size_t)mesh * n_steps + step];
    ga = gamma[(size_t)mesh * n_steps + step];
    cc = center[mesh];
  }
  const int g = threadIdx.x / TPR, t = threadIdx.x % TPR;
  for (int row = r0 + g; row < r1; row += SPMM_THREADS / TPR) {
    double2 acc[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) acc[v] = make_double2(0.0, 0.0);
    const int p0 = row_ptr[row], p1 = row_ptr[row + 1];
#pragma unroll 4
    for (int p = p0; p < p1; ++p) {
      const int c = cols[p];
      const double w = weights[p];
      const double2* src = reinterpret_cast<const double2*>(y + (size_t)c * B);
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const double2 a = __ldg(src + t + v * TPR);
        acc[v].x = fma(w, a.x, acc[v].x);
        acc[v].y = fma(w, a.y, acc[v].y);
      }
    }
    const double d = degree[row];
    const double di = degree_inv[row];
    const double2* yr = reinterpret_cast<const double2*>(y + (size_t)row * B);
    const double2* xr = reinterpret_cast<const double2*>(x_prev + (size_t)row * B);
    double2* o = reinterpret_cast<double2*>(out + (size_t)row * B);
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const double2 yv = __ldg(yr + t + v * TPR);
      double2 r;
      if (MODE == 1) {
        r.x = d * yv.x - acc[v].x;
        r.y = d * yv.y - acc[v].y;
      } else if (MODE == 2) {
        r.x = di * (d * yv.x - acc[v].x);
        r.y = di * (d * yv.y - acc[v].y);
      } else {
        const double lx = di * (d * yv.x - acc[v].x);
        const double ly = di * (d * yv.y - acc[v].y);
        r.x = al * (lx - cc * yv.x);
        r.y = al * (ly - cc * yv.y);
        if (ga != 0.0) {  // step 0 has no predecessor: x_prev may be uninitialised memory
          const double2 xv = __ldg(xr + t + v * TPR);
          r.x -= ga * xv.x;
          r.y -= ga * xv.y;
        }
      }
      o[t + v * TPR] = r;
    }
  }
}


// Tuning record of the CSR filter steps (B200, 256 meshes x 15 212 vertices, b = 16; variants that were measured and
// removed -- the fp32 forms of the step now live in sell.cu on a sliced-ELL copy of the matrix):
//   L2 prefetch of the CTA's streams, variants 0/1/2/3: fp64 step 4295 / 4136 / 4070 / 4059 GB/s (the prefetches compete
//   with a memory system that is already busy) -> off; fp32 step 4483 / 4521 / 4687 / 4758 GB/s (half the bytes per row:
//   more latency-bound, so asking early pays) -> variant 3 is what the fp32 kernels do.
//   TMA bulk-staged window of y rows in shared memory (cp.async.bulk + mbarrier): 3712 against 4355 GB/s (the hardware
//   L1 already serves 59% of the gathers and the 48 KB window halves the resident warps).
//   Resident CTAs per SM of the correction step 8 / 6 / 5 (32 / 40 / 48 registers): 5032 / 4978 / 4701 GB/s.
//   Streaming cache operators (ld.cs / st.cs) on the single-use streams: 0.252 against 0.2445 ms per launch.
template <int B, int TPR>
static int launch_spmm_b(int mode, const SpmmGraph& g, const double* y, const double* x_prev, double* out,
                         const double* alpha, const double* gamma, const double* center, int step,
                         int n_steps, cudaStream_t stream) {
  // A CTA normally walks 256 rows in passes of 256 / TPR; a single mesh (the drop-in Focusr(target, source) call: 15k
  // rows = 60 such CTAs on 148 SMs, 20 us per step of 16 MB) gets one pass per CTA instead, so that the whole GPU works
  // on the step.  Same arithmetic per row either way.
  int rpb = SPMM_ROWS_PER_BLOCK;
  if ((long long)div_up(g.max_mesh_rows, rpb) * g.n_meshes < 2 * sm_count()) rpb = SPMM_THREADS / TPR;
  dim3 grid(div_up(g.max_mesh_rows, rpb), g.n_meshes);
  if (mode == 0)
    k_spmm<B, TPR, 0><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                         g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps, rpb);
  else if (mode == 1)
    k_spmm<B, TPR, 1><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                         g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps, rpb);
  else
    k_spmm<B, TPR, 2><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                         g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps, rpb);
  FB_COUNT_LAUNCH(1);
  return FB_OK;
}

bool spmm_block_supported(int b) {
  switch (b) {
    case 8: case 16: case 24: case 32: case 40: case 48: case 56: case 64: case 72: case 80: case 88: case 96:
      return true;
    default:
      return false;
  }
}

int launch_spmm(int mode, int b, const SpmmGraph& g, const double* y, const double* x_prev, double* out,
                const double* alpha, const double* gamma, const double* center, int step, int n_steps,
                cudaStream_t stream) {
#define FB_CASE(BB, TT) \
  case BB:              \
    return launch_spmm_b<BB, TT>(mode, g, y, x_prev, out, alpha, gamma, center, step, n_steps, stream);
  switch (b) {
    FB_CASE(8, 4)
    FB_CASE(16, 8)
    FB_CASE(24, 4)
    FB_CASE(32, 8)
    FB_CASE(40, 4)
    FB_CASE(48, 8)
    FB_CASE(56, 4)
    FB_CASE(64, 16)
    FB_CASE(72, 4)
    FB_CASE(80, 8)
    FB_CASE(88, 4)
    FB_CASE(96, 16)
    default:
      set_error("spmm: unsupported block size %d (multiples of 8 up to 96)", b);
      return FB_ERR_UNSUPPORTED;
  }
#undef FB_CASE
}


// ---------------------------------------------------------------------------------------------
// Row-partitioned multi-GPU variant with the halo FUSED into the SpMM: a column that belongs to
// another rank is read straight from that rank's memory over NVLink (peer pointer obtained through
// CUDA IPC), so there is no pack / send / receive and no ghost copy -- the transfer of the ~1% of
// boundary rows overlaps the gathers of the local 99%.  peer_y[p] is rank p's copy of the block being
// read; ghost g = col - n_loc lives at row ghost_row[g] of rank ghost_peer[g].  Ordering between
// ranks is a flag barrier (k_peer_barrier in eigs.cu) before the launch.
// ---------------------------------------------------------------------------------------------
template <int B, int TPR, int MODE>
__global__ void __launch_bounds__(SPMM_THREADS, (B / (2 * TPR) == 1) ? 8 : ((B / (2 * TPR) == 2) ? 6 : 3))
k_spmm_p2p(const int* __restrict__ row_ptr, const int* __restrict__ cols, const double* __restrict__ weights,
           const double* __restrict__ degree, const double* __restrict__ degree_inv, int n_loc,
           const double* __restrict__ y, const double* const* __restrict__ peer_y,
           const int* __restrict__ ghost_peer, const int* __restrict__ ghost_row,
           const double* __restrict__ x_prev, double* __restrict__ out, const double* __restrict__ alpha,
           const double* __restrict__ gamma, const double* __restrict__ center, int step, int n_steps) {
  constexpr int VPT = B / (2 * TPR);
  const int r0 = blockIdx.x * SPMM_ROWS_PER_BLOCK;
  const int r1 = min(n_loc, r0 + SPMM_ROWS_PER_BLOCK);
  double al = 1.0, ga = 0.0, cc = 0.0;
  if (MODE == 0) {
    al = alpha[step];
    ga = gamma[step];
    cc = center[0];
  }
  const int g = threadIdx.x / TPR, t = threadIdx.x % TPR;
  for (int row = r0 + g; row < r1; row += SPMM_THREADS / TPR) {
    double2 acc[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) acc[v] = make_double2(0.0, 0.0);
    const int p0 = row_ptr[row], p1 = row_ptr[row + 1];
#pragma unroll 4
    for (int p = p0; p < p1; ++p) {
      const int c = cols[p];
      const double w = weights[p];
      const double2* src;
      if (c < n_loc) {
        src = reinterpret_cast<const double2*>(y + (size_t)c * B);
      } else {
        const int gi = c - n_loc;
        src = reinterpret_cast<const double2*>(peer_y[ghost_peer[gi]] + (size_t)ghost_row[gi] * B);
      }
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const double2 a = __ldg(src + t + v * TPR);
        acc[v].x = fma(w, a.x, acc[v].x);
        acc[v].y = fma(w, a.y, acc[v].y);
      }
    }
    const double d = degree[row];
    const double di = degree_inv[row];
    const double2* yr = reinterpret_cast<const double2*>(y + (size_t)row * B);
    const double2* xr = reinterpret_cast<const double2*>(x_prev + (size_t)row * B);
    double2* o = reinterpret_cast<double2*>(out + (size_t)row * B);
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const double2 yv = __ldg(yr + t + v * TPR);
      double2 r;
      if (MODE == 1) {
        r.x = d * yv.x - acc[v].x;
        r.y = d * yv.y - acc[v].y;
      } else {
        const double lx = di * (d * yv.x - acc[v].x);
        const double ly = di * (d * yv.y - acc[v].y);
        r.x = al * (lx - cc * yv.x);
        r.y = al * (ly - cc * yv.y);
        if (ga != 0.0) {
          const double2 xv = __ldg(xr + t + v * TPR);
          r.x -= ga * xv.x;
          r.y -= ga * xv.y;
        }
      }
      o[t + v * TPR] = r;
    }
  }
}

template <int B, int TPR>
static int launch_spmm_p2p_b(int mode, const SpmmGraph& g, int n_loc, const double* y, const double* const* peer_y,
                             const int* ghost_peer, const int* ghost_row, const double* x_prev, double* out,
                             const double* alpha, const double* gamma, const double* center, int step, int n_steps,
                             cudaStream_t stream) {
  dim3 grid(div_up(n_loc, SPMM_ROWS_PER_BLOCK));
  if (mode == 0)
    k_spmm_p2p<B, TPR, 0><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv, n_loc, y,
                                                             peer_y, ghost_peer, ghost_row, x_prev, out, alpha, gamma, center, step, n_steps);
  else
    k_spmm_p2p<B, TPR, 1><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv, n_loc, y,
                                                             peer_y, ghost_peer, ghost_row, x_prev, out, alpha, gamma, center, step, n_steps);
  FB_COUNT_LAUNCH(1);
  return FB_OK;
}

int launch_spmm_p2p(int mode, int b, const SpmmGraph& g, int n_loc, const double* y, const double* const* peer_y,
                    const int* ghost_peer, const int* ghost_row, const double* x_prev, double* out, const double* alpha,
                    const double* gamma, const double* center, int step, int n_steps, cudaStream_t stream) {
#define FB_CASE(BB, TT) \
  case BB:              \
    return launch_spmm_p2p_b<BB, TT>(mode, g, n_loc, y, peer_y, ghost_peer, ghost_row, x_prev, out, alpha, gamma, center, step, n_steps, stream);
  switch (b) {
    FB_CASE(8, 4)
    FB_CASE(16, 8)
    FB_CASE(24, 4)
    FB_CASE(32, 8)
    FB_CASE(40, 4)
    FB_CASE(48, 8)
    FB_CASE(56, 4)
    FB_CASE(64, 16)
    FB_CASE(72, 4)
    FB_CASE(80, 8)
    FB_CASE(88, 4)
    FB_CASE(96, 16)
    default:
      set_error("spmm (p2p): unsupported block size %d", b);
      return FB_ERR_UNSUPPORTED;
  }
#undef FB_CASE
}

__global__ void k_gather_rows(const double* __restrict__ in, const long long* __restrict__ idx,
                              const int* __restrict__ idx_base, int n_rows, int n_cols,
                              double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_rows * n_cols) return;
  const int i = (int)(t / n_cols), k = (int)(t - (long long)i * n_cols);
  const long long src = idx[i] + (idx_base ? idx_base[i] : 0);
  out[t] = in[src * n_cols + k];
}

}  // namespace fb

using namespace fb;

extern "C" {

int focusr_gather_rows(const double* in, const long long* idx, const int* idx_base, int n_rows, int n_cols,
                       double* out, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_rows > 0 && n_cols > 0, "gather_rows: empty");
  const int T = 256;
  k_gather_rows<<<div_up((long long)n_rows * n_cols, T), T, 0, stream>>>(in, idx, idx_base, n_rows, n_cols, out);
  FB_COUNT_LAUNCH(1);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

int focusr_laplacian_apply(const int* row_ptr, const int* cols, const double* weights, const double* degree,
                           const double* degree_inv, const int* mesh_point_off, int n_meshes,
                           int max_mesh_points, const double* x, double* y, int n_cols,
                           focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_meshes > 0 && max_mesh_points > 0 && spmm_block_supported(n_cols),
             "laplacian_apply: n_cols must be a multiple of 8 up to 96");
  SpmmGraph g{row_ptr, cols, weights, degree, degree_inv, mesh_point_off, n_meshes, max_mesh_points};
  int rc = launch_spmm(2, n_cols, g, x, x, y, nullptr, nullptr, nullptr, 0, 0, stream);
  if (rc) return rc;
  FB_LAUNCH_CHECK();
  return FB_OK;
}
}
