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

// PF > 0 (focusr_set_tuning(4, PF), b = 16 filter steps only): the CTA asks L2 for the lines it will stream from HBM
// before it starts walking rows: 1 = its x_prev rows, 2 = + its own y rows, 3 = + its slice of cols / weights.
template <int B, int TPR, int MODE, int PF = 0>
__global__ void __launch_bounds__(SPMM_THREADS, (B / (2 * TPR) == 1) ? 8 : ((B / (2 * TPR) == 2) ? 6 : 3))
k_spmm(const int* __restrict__ row_ptr, const int* __restrict__ cols, const double* __restrict__ weights,
       const double* __restrict__ degree, const double* __restrict__ degree_inv,
       const int* __restrict__ mesh_off, const double* __restrict__ y, const double* __restrict__ x_prev,
       double* __restrict__ out, const double* __restrict__ alpha, const double* __restrict__ gamma,
       const double* __restrict__ center, int step, int n_steps) {
  constexpr int VPT = B / (2 * TPR);
  static_assert(VPT * 2 * TPR == B, "block size must be a multiple of 2*TPR");
  const int mesh = blockIdx.y;
  const int r0 = mesh_off[mesh] + blockIdx.x * SPMM_ROWS_PER_BLOCK;
  const int r1 = min(mesh_off[mesh + 1], r0 + SPMM_ROWS_PER_BLOCK);
  if (r0 >= r1) return;
  double al = 1.0, ga = 0.0, cc = 0.0;
  if (MODE == 0) {
    al = alpha[(size_t)mesh * n_steps + step];
    ga = gamma[(size_t)mesh * n_steps + step];
    cc = center[mesh];
  }
  if (PF > 0) {
    static_assert(PF == 0 || B * 8 == 128, "prefetch variants assume one 128-byte line per row");
    const int pr = r0 + (int)threadIdx.x;  // SPMM_ROWS_PER_BLOCK == SPMM_THREADS: one row line per thread
    if (pr < r1) {
      if (ga != 0.0) prefetch_l2(x_prev + (size_t)pr * B);
      if (PF >= 2) prefetch_l2(y + (size_t)pr * B);
    }
    if (PF >= 3) {
      const int q0 = row_ptr[r0], q1 = row_ptr[r1];
      for (int q = q0 + 32 * (int)threadIdx.x; q < q1; q += 32 * SPMM_THREADS) prefetch_l2(cols + q);
      for (int q = q0 + 16 * (int)threadIdx.x; q < q1; q += 16 * SPMM_THREADS) prefetch_l2(weights + q);
    }
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


// ---------------------------------------------------------------------------------------------
// TMA-staged variant: the window of y rows the CTA's rows gather from -- its own 256 rows plus
// SPMM_HALO rows either side, which is where mesh neighbours live for locality-ordered vertices --
// is brought into shared memory by ONE bulk asynchronous copy (cp.async.bulk, SASS UBLKCP) signalled
// through an mbarrier; gathers hit shared memory, columns outside the window fall back to global.
// Selected with focusr_set_tuning(0, 1).  Measured on B200 (256 meshes x 15 212 vertices, b = 16): bit-identical
// output, 3712 GB/s in the filter against 4355 GB/s for k_spmm -- the hardware L1 already serves 59% of the
// gathers and the 48 KB window costs half the resident warps -- so it is kept as an A/B variant, not the default.
// ---------------------------------------------------------------------------------------------
constexpr int SPMM_HALO = 64;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int B, int TPR, int MODE>
__global__ void __launch_bounds__(SPMM_THREADS, 4)
k_spmm_staged(const int* __restrict__ row_ptr, const int* __restrict__ cols, const double* __restrict__ weights,
              const double* __restrict__ degree, const double* __restrict__ degree_inv,
              const int* __restrict__ mesh_off, const double* __restrict__ y, const double* __restrict__ x_prev,
              double* __restrict__ out, const double* __restrict__ alpha, const double* __restrict__ gamma,
              const double* __restrict__ center, int step, int n_steps) {
  constexpr int VPT = B / (2 * TPR);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_y = reinterpret_cast<double*>(smem_raw);  // [(ROWS + 2 HALO)][B]
  __shared__ __align__(8) unsigned long long s_bar;
  const int mesh = blockIdx.y;
  const int m_lo = mesh_off[mesh], m_hi = mesh_off[mesh + 1];
  const int r0 = m_lo + blockIdx.x * SPMM_ROWS_PER_BLOCK;
  const int r1 = min(m_hi, r0 + SPMM_ROWS_PER_BLOCK);
  if (r0 >= r1) return;
  const int w0 = max(m_lo, r0 - SPMM_HALO), w1 = min(m_hi, r1 + SPMM_HALO);
  const unsigned bar = smem_u32(&s_bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned)(w1 - w0) * B * 8u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(s_y)),
                 "l"(y + (size_t)w0 * B), "r"(bytes), "r"(bar)
                 : "memory");
  }
  double al = 1.0, ga = 0.0, cc = 0.0;
  if (MODE == 0) {
    al = alpha[(size_t)mesh * n_steps + step];
    ga = gamma[(size_t)mesh * n_steps + step];
    cc = center[mesh];
  }
  const int g = threadIdx.x / TPR, t = threadIdx.x % TPR;
  // first row's pointers are fetched while the bulk copy is in flight
  int row = r0 + g;
  int p0 = row < r1 ? row_ptr[row] : 0, p1 = row < r1 ? row_ptr[row + 1] : 0;
  {
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    }
  }
  for (; row < r1; row += SPMM_THREADS / TPR) {
    double2 acc[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) acc[v] = make_double2(0.0, 0.0);
#pragma unroll 4
    for (int p = p0; p < p1; ++p) {
      const int c = cols[p];
      const double w = weights[p];
      const bool in_win = c >= w0 && c < w1;
      const double2* src = in_win ? reinterpret_cast<const double2*>(s_y + (size_t)(c - w0) * B)
                                  : reinterpret_cast<const double2*>(y + (size_t)c * B);
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const double2 a = src[t + v * TPR];
        acc[v].x = fma(w, a.x, acc[v].x);
        acc[v].y = fma(w, a.y, acc[v].y);
      }
    }
    const double d = degree[row];
    const double di = degree_inv[row];
    const double2* yr = reinterpret_cast<const double2*>(s_y + (size_t)(row - w0) * B);
    const double2* xr = reinterpret_cast<const double2*>(x_prev + (size_t)row * B);
    double2* o = reinterpret_cast<double2*>(out + (size_t)row * B);
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const double2 yv = yr[t + v * TPR];
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
        if (ga != 0.0) {
          const double2 xv = __ldg(xr + t + v * TPR);
          r.x -= ga * xv.x;
          r.y -= ga * xv.y;
        }
      }
      o[t + v * TPR] = r;
    }
    const int nrow = row + SPMM_THREADS / TPR;
    if (nrow < r1) {
      p0 = row_ptr[nrow];
      p1 = row_ptr[nrow + 1];
    }
  }
}

// L2 prefetch variants of the b = 16 filter steps (see k_spmm).  Measured on B200, 256 meshes x 15 212 vertices, variants
// 0/1/2/3: fp64 step 4295 / 4136 / 4070 / 4059 GB/s (the prefetches compete with a memory system that is already
// busy), fp32 step 4483 / 4521 / 4687 / 4758 GB/s (half the bytes per row: more latency-bound, so asking early pays).
int g_spmm_prefetch = 0;      // focusr_set_tuning(4, v): fp64 step
int g_spmm_prefetch_f32 = 3;  // focusr_set_tuning(5, v): fp32 step
// focusr_set_tuning(7, v): resident CTAs per SM the correction step is compiled for (0 = 8 at 32 registers; 6 -> 40, 5 -> 48
// registers).  Measured on B200 (128 pairs per launch): 5032 / 4978 / 4701 GB/s for 8 / 6 / 5 -- occupancy wins again.
// focusr_set_tuning(6, 1) (streaming cache operators on the single-use streams): 0.252 ms against 0.2445 ms per launch.
int g_spmm_minb = 0;
int g_spmm_hint = 0;          // focusr_set_tuning(6, v): streaming cache operators in the fp32 correction step
int g_spmm_variant = 0;  // focusr_set_tuning(0, v): 0 = register-capped gather kernel, 1 = TMA-staged window
int g_mixed_precision = 1;  // focusr_set_tuning(3, v): 1 = early filter passes in fp32 (default), 0 = fp64 throughout

template <int B, int TPR>
static int launch_spmm_b(int mode, const SpmmGraph& g, const double* y, const double* x_prev, double* out,
                         const double* alpha, const double* gamma, const double* center, int step,
                         int n_steps, cudaStream_t stream) {
  dim3 grid(div_up(g.max_mesh_rows, SPMM_ROWS_PER_BLOCK), g.n_meshes);
  if (g_spmm_variant == 1 && B <= 32) {
    const size_t smem = sizeof(double) * (size_t)(SPMM_ROWS_PER_BLOCK + 2 * SPMM_HALO) * B;
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(k_spmm_staged<B, TPR, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_spmm_staged<B, TPR, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_spmm_staged<B, TPR, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr_done = true;
    }
    if (mode == 0)
      k_spmm_staged<B, TPR, 0><<<grid, SPMM_THREADS, smem, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                                     g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps);
    else if (mode == 1)
      k_spmm_staged<B, TPR, 1><<<grid, SPMM_THREADS, smem, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                                     g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps);
    else
      k_spmm_staged<B, TPR, 2><<<grid, SPMM_THREADS, smem, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                                     g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps);
    FB_COUNT_LAUNCH(1);
    return FB_OK;
  }
  if (mode == 0 && B == 16 && g_spmm_prefetch > 0) {
    constexpr int BB = B == 16 ? B : 16, TT = B == 16 ? TPR : 8;  // only instantiated for b = 16
    if (g_spmm_prefetch == 1)
      k_spmm<BB, TT, 0, 1><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                              g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps);
    else if (g_spmm_prefetch == 2)
      k_spmm<BB, TT, 0, 2><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                              g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps);
    else
      k_spmm<BB, TT, 0, 3><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                              g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps);
  } else if (mode == 0)
    k_spmm<B, TPR, 0><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                         g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps);
  else if (mode == 1)
    k_spmm<B, TPR, 1><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                         g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps);
  else
    k_spmm<B, TPR, 2><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights, g.degree, g.degree_inv,
                                                         g.mesh_off, y, x_prev, out, alpha, gamma, center, step, n_steps);
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
// Mixed-precision filter step: the same fused SpMM + three-term update with the vector blocks stored
// and updated in fp32 (the matrix comes from its fp32 copy -- k_matrix_f32: weights, (degree, 1/degree~) per row --,
// table values are rounded as they are loaded; fmaf throughout).  Used by the driver only for passes that are meant to land above the fp32 floor
// (chfsi_driver.hpp); what is tested and returned is always computed in fp64.  Algorithmic bytes per
// step and mesh: 8 nnz + 12 N + 12 b N (against 12 nnz + 20 N + 24 b N): 252 N instead of 476 N at b = 16.
//   IO 0: y, x_prev, out fp32.
//   IO 1: first step of a pass -- y is the fp64 block; writes out (fp32) and an fp32 copy of y, which
//         is the next step's x_prev (gamma is 0 at step 0, x_prev is not read).
//   IO 2: last step of a pass -- y, x_prev fp32, out is written as fp64.
// A thread owns B/(4*TPR) float4 slices of a row, so gathers stay 16-byte loads.
// ---------------------------------------------------------------------------------------------
template <int B, int TPR, int IO, int PF = 0>
__global__ void __launch_bounds__(SPMM_THREADS, (B / (4 * TPR) == 1) ? 8 : ((B / (4 * TPR) == 2) ? 6 : 3))
k_spmm_f32(const int* __restrict__ row_ptr, const int* __restrict__ cols, const float* __restrict__ weights,
           const float2* __restrict__ ddi,
           const int* __restrict__ mesh_off, const void* __restrict__ y_, const float* __restrict__ x_prev,
           void* __restrict__ out_, float* __restrict__ y_copy, const double* __restrict__ alpha,
           const double* __restrict__ gamma, const double* __restrict__ center, int step, int n_steps) {
  constexpr int VPT = B / (4 * TPR);
  static_assert(VPT * 4 * TPR == B, "block size must be a multiple of 4*TPR");
  const int mesh = blockIdx.y;
  const int r0 = mesh_off[mesh] + blockIdx.x * SPMM_ROWS_PER_BLOCK;
  const int r1 = min(mesh_off[mesh + 1], r0 + SPMM_ROWS_PER_BLOCK);
  if (r0 >= r1) return;
  const float al = (float)alpha[(size_t)mesh * n_steps + step];
  const float ga = (float)gamma[(size_t)mesh * n_steps + step];
  const float cc = (float)center[mesh];
  if (PF > 0 && IO == 0) {  // same L2 prefetch variants as k_spmm
    const int pr = r0 + (int)threadIdx.x;
    if (pr < r1) {
      if (ga != 0.f) prefetch_l2(x_prev + (size_t)pr * B);
      if (PF >= 2) prefetch_l2(static_cast<const float*>(y_) + (size_t)pr * B);
    }
    if (PF >= 3) {
      const int q0 = row_ptr[r0], q1 = row_ptr[r1];
      for (int q = q0 + 32 * (int)threadIdx.x; q < q1; q += 32 * SPMM_THREADS) prefetch_l2(cols + q);
      for (int q = q0 + 32 * (int)threadIdx.x; q < q1; q += 32 * SPMM_THREADS) prefetch_l2(weights + q);
    }
  }
  const int g = threadIdx.x / TPR, t = threadIdx.x % TPR;
  auto load_y = [&](int r, int slice) -> float4 {
    if (IO == 1) {
      const double2* src = reinterpret_cast<const double2*>(static_cast<const double*>(y_) + (size_t)r * B) + 2 * slice;
      const double2 a = __ldg(src), b = __ldg(src + 1);
      return make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
    } else {
      return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(y_) + (size_t)r * B) + slice);
    }
  };
  for (int row = r0 + g; row < r1; row += SPMM_THREADS / TPR) {
    float4 acc[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int p0 = row_ptr[row], p1 = row_ptr[row + 1];
#pragma unroll 4
    for (int p = p0; p < p1; ++p) {
      const int c = cols[p];
      const float w = weights[p];
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const float4 a = load_y(c, t + v * TPR);
        acc[v].x = fmaf(w, a.x, acc[v].x);
        acc[v].y = fmaf(w, a.y, acc[v].y);
        acc[v].z = fmaf(w, a.z, acc[v].z);
        acc[v].w = fmaf(w, a.w, acc[v].w);
      }
    }
    const float2 dd = ddi[row];
    const float d = dd.x, di = dd.y;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int slice = t + v * TPR;
      const float4 yv = load_y(row, slice);
      float4 r;
      r.x = al * (di * (d * yv.x - acc[v].x) - cc * yv.x);
      r.y = al * (di * (d * yv.y - acc[v].y) - cc * yv.y);
      r.z = al * (di * (d * yv.z - acc[v].z) - cc * yv.z);
      r.w = al * (di * (d * yv.w - acc[v].w) - cc * yv.w);
      if (IO != 1 && ga != 0.f) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x_prev + (size_t)row * B) + slice);
        r.x -= ga * xv.x;
        r.y -= ga * xv.y;
        r.z -= ga * xv.z;
        r.w -= ga * xv.w;
      }
      if (IO == 2) {
        double2* o = reinterpret_cast<double2*>(static_cast<double*>(out_) + (size_t)row * B) + 2 * slice;
        o[0] = make_double2((double)r.x, (double)r.y);
        o[1] = make_double2((double)r.z, (double)r.w);
      } else {
        reinterpret_cast<float4*>(static_cast<float*>(out_) + (size_t)row * B)[slice] = r;
      }
      if (IO == 1) reinterpret_cast<float4*>(y_copy + (size_t)row * B)[slice] = yv;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 CORRECTION step (chfsi_driver.hpp): z_next = alpha_j ((L - c) z + r_j) - gamma_j z_prev per column j, where
// r = L x - theta x is the fp64 residual block of the Ritz vectors rounded to fp32, the tables are per column (the
// polynomial of column j is normalised to 1 at theta_j) and z starts at 0, so that p(L) x = x + z.  z is as small as
// the error of x: fp32 rounding here is relative to that error.  LAST: the fp64 block is updated, x += z_next.
// Algorithmic bytes per step and mesh: 8 nnz + 12 N + 16 b N (z, z_prev, r read; z_next written): 316 N at b = 16.
// ---------------------------------------------------------------------------------------------
// HINT 1 (focusr_set_tuning(6, 1), b = 16): single-use streams (z_prev, r, the matrix, z_next) are accessed with the
// streaming cache operators so that the gathered rows of z, which are reused, stay in L2.
template <int B, int TPR, int LAST, int HINT = 0, int MINB = 0>
__global__ void __launch_bounds__(SPMM_THREADS, MINB ? MINB : ((B / (4 * TPR) == 1) ? 8 : ((B / (4 * TPR) == 2) ? 6 : 3)))
k_spmm_corr(const int* __restrict__ row_ptr, const int* __restrict__ cols, const float* __restrict__ weights,
            const float2* __restrict__ ddi,
            const int* __restrict__ mesh_off, const float* __restrict__ z, const float* __restrict__ z_prev,
            const float* __restrict__ r, float* __restrict__ z_next, double* __restrict__ x,
            const float* __restrict__ alpha_c, const float* __restrict__ gamma_c, const double* __restrict__ center,
            int step, int n_steps, int has_prev, int prefetch) {
  constexpr int VPT = B / (4 * TPR);
  static_assert(VPT * 4 * TPR == B, "block size must be a multiple of 4*TPR");
  const int mesh = blockIdx.y;
  const int r0 = mesh_off[mesh] + blockIdx.x * SPMM_ROWS_PER_BLOCK;
  const int r1 = min(mesh_off[mesh + 1], r0 + SPMM_ROWS_PER_BLOCK);
  if (r0 >= r1) return;
  if (prefetch) {  // ask L2 early for what this CTA streams from HBM (pays for the fp32 steps, see g_spmm_prefetch_f32)
    const int pr = r0 + (int)threadIdx.x;
    if (pr < r1) {
      if (has_prev) prefetch_l2(z_prev + (size_t)pr * B);
      prefetch_l2(z + (size_t)pr * B);
      prefetch_l2(r + (size_t)pr * B);
    }
    const int q0 = row_ptr[r0], q1 = row_ptr[r1];
    for (int q = q0 + 32 * (int)threadIdx.x; q < q1; q += 32 * SPMM_THREADS) prefetch_l2(cols + q);
    for (int q = q0 + 32 * (int)threadIdx.x; q < q1; q += 32 * SPMM_THREADS) prefetch_l2(weights + q);
  }
  const float cc = (float)center[mesh];
  const int g = threadIdx.x / TPR, t = threadIdx.x % TPR;
  const float4* al_p = reinterpret_cast<const float4*>(alpha_c + ((size_t)mesh * n_steps + step) * B);
  const float4* ga_p = reinterpret_cast<const float4*>(gamma_c + ((size_t)mesh * n_steps + step) * B);
  for (int row = r0 + g; row < r1; row += SPMM_THREADS / TPR) {
    float4 acc[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int p0 = row_ptr[row], p1 = row_ptr[row + 1];
#pragma unroll 4
    for (int p = p0; p < p1; ++p) {
      const int c = HINT ? __ldcs(cols + p) : cols[p];
      const float w = HINT ? __ldcs(weights + p) : weights[p];
      const float4* src = reinterpret_cast<const float4*>(z + (size_t)c * B);
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const float4 a = __ldg(src + t + v * TPR);
        acc[v].x = fmaf(w, a.x, acc[v].x);
        acc[v].y = fmaf(w, a.y, acc[v].y);
        acc[v].z = fmaf(w, a.z, acc[v].z);
        acc[v].w = fmaf(w, a.w, acc[v].w);
      }
    }
    const float2 dd = ddi[row];
    const float d = dd.x, di = dd.y;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int slice = t + v * TPR;
      const float4 zv = __ldg(reinterpret_cast<const float4*>(z + (size_t)row * B) + slice);
      const float4* r_p = reinterpret_cast<const float4*>(r + (size_t)row * B) + slice;
      const float4 rv = HINT ? __ldcs(r_p) : __ldg(r_p);
      const float4 al = __ldg(al_p + slice);
      float4 o;
      o.x = al.x * ((di * (d * zv.x - acc[v].x) - cc * zv.x) + rv.x);
      o.y = al.y * ((di * (d * zv.y - acc[v].y) - cc * zv.y) + rv.y);
      o.z = al.z * ((di * (d * zv.z - acc[v].z) - cc * zv.z) + rv.z);
      o.w = al.w * ((di * (d * zv.w - acc[v].w) - cc * zv.w) + rv.w);
      if (has_prev) {
        const float4 ga = __ldg(ga_p + slice);
        const float4* p_p = reinterpret_cast<const float4*>(z_prev + (size_t)row * B) + slice;
        const float4 pv = HINT ? __ldcs(p_p) : __ldg(p_p);
        o.x -= ga.x * pv.x;
        o.y -= ga.y * pv.y;
        o.z -= ga.z * pv.z;
        o.w -= ga.w * pv.w;
      }
      if (LAST) {
        double2* xo = reinterpret_cast<double2*>(x + (size_t)row * B) + 2 * slice;
        double2 a = xo[0], b = xo[1];
        a.x += (double)o.x;
        a.y += (double)o.y;
        b.x += (double)o.z;
        b.y += (double)o.w;
        xo[0] = a;
        xo[1] = b;
      } else {
        float4* o_p = reinterpret_cast<float4*>(z_next + (size_t)row * B) + slice;
        if (HINT) __stcs(o_p, o);
        else *o_p = o;
      }
    }
  }
}

template <int B, int TPR>
static int launch_spmm_corr_b(bool last, const SpmmGraph& g, const float* z, const float* z_prev, const float* r,
                              float* z_next, double* x, const float* alpha_c, const float* gamma_c, const double* center,
                              int step, int n_steps, bool has_prev, cudaStream_t stream) {
  dim3 grid(div_up(g.max_mesh_rows, SPMM_ROWS_PER_BLOCK), g.n_meshes);
  const int pf = (B == 16 && g_spmm_prefetch_f32 > 0) ? 1 : 0;
  if (!last && B == 16 && g_spmm_minb > 0) {  // occupancy A/B (focusr_set_tuning(7, v)): 6 or 5 CTAs per SM instead of 8
    constexpr int BB = B == 16 ? B : 16, TT = B == 16 ? TPR : 4;
    if (g_spmm_minb == 6)
      k_spmm_corr<BB, TT, 0, 0, 6><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off, z,
                                                                      z_prev, r, z_next, x, alpha_c, gamma_c, center, step, n_steps,
                                                                      has_prev ? 1 : 0, pf);
    else
      k_spmm_corr<BB, TT, 0, 0, 5><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off, z,
                                                                      z_prev, r, z_next, x, alpha_c, gamma_c, center, step, n_steps,
                                                                      has_prev ? 1 : 0, pf);
  } else if (!last && B == 16 && g_spmm_hint == 1) {
    constexpr int BB = B == 16 ? B : 16, TT = B == 16 ? TPR : 4;  // only instantiated for b = 16
    k_spmm_corr<BB, TT, 0, 1><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off, z,
                                                                 z_prev, r, z_next, x, alpha_c, gamma_c, center, step, n_steps,
                                                                 has_prev ? 1 : 0, pf);
  } else if (last)
    k_spmm_corr<B, TPR, 1><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off, z,
                                                              z_prev, r, z_next, x, alpha_c, gamma_c, center, step, n_steps,
                                                              has_prev ? 1 : 0, pf);
  else
    k_spmm_corr<B, TPR, 0><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off, z,
                                                              z_prev, r, z_next, x, alpha_c, gamma_c, center, step, n_steps,
                                                              has_prev ? 1 : 0, pf);
  FB_COUNT_LAUNCH(1);
  return FB_OK;
}

int launch_spmm_corr(bool last, int b, const SpmmGraph& g, const float* z, const float* z_prev, const float* r, float* z_next,
                     double* x, const float* alpha_c, const float* gamma_c, const double* center, int step, int n_steps,
                     bool has_prev, cudaStream_t stream) {
#define FB_CASE(BB, TT) \
  case BB:              \
    return launch_spmm_corr_b<BB, TT>(last, g, z, z_prev, r, z_next, x, alpha_c, gamma_c, center, step, n_steps, has_prev, stream);
  switch (b) {
    FB_CASE(8, 2)
    FB_CASE(16, 4)
    FB_CASE(24, 2)
    FB_CASE(32, 4)
    FB_CASE(40, 2)
    FB_CASE(48, 4)
    FB_CASE(56, 2)
    FB_CASE(64, 4)
    FB_CASE(72, 2)
    FB_CASE(80, 4)
    FB_CASE(88, 2)
    FB_CASE(96, 4)
    default:
      set_error("spmm (fp32 correction): unsupported block size %d (multiples of 8 up to 96)", b);
      return FB_ERR_UNSUPPORTED;
  }
#undef FB_CASE
}

__global__ void k_matrix_f32(const double* __restrict__ weights, const double* __restrict__ degree,
                             const double* __restrict__ degree_inv, long long nnz, int n_rows, float* __restrict__ wf,
                             float2* __restrict__ ddi) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nnz) wf[t] = (float)weights[t];
  if (t < n_rows) ddi[t] = make_float2((float)degree[t], (float)degree_inv[t]);
}

int launch_matrix_f32(const double* weights, const double* degree, const double* degree_inv, long long nnz, int n_rows,
                      float* wf, float2* ddi, cudaStream_t stream) {
  const long long n = nnz > n_rows ? nnz : n_rows;
  k_matrix_f32<<<div_up(n, 256), 256, 0, stream>>>(weights, degree, degree_inv, nnz, n_rows, wf, ddi);
  FB_COUNT_LAUNCH(1);
  return FB_OK;
}

template <int B, int TPR>
static int launch_spmm_f32_b(int io, const SpmmGraph& g, const void* y, const float* x_prev, void* out, float* y_copy,
                             const double* alpha, const double* gamma, const double* center, int step, int n_steps,
                             cudaStream_t stream) {
  dim3 grid(div_up(g.max_mesh_rows, SPMM_ROWS_PER_BLOCK), g.n_meshes);
  if (io == 0 && B == 16 && g_spmm_prefetch_f32 > 0) {
    constexpr int BB = B == 16 ? B : 16, TT = B == 16 ? TPR : 4;  // only instantiated for b = 16
    if (g_spmm_prefetch_f32 == 1)
      k_spmm_f32<BB, TT, 0, 1><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off,
                                                                  y, x_prev, out, y_copy, alpha, gamma, center, step, n_steps);
    else if (g_spmm_prefetch_f32 == 2)
      k_spmm_f32<BB, TT, 0, 2><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off,
                                                                  y, x_prev, out, y_copy, alpha, gamma, center, step, n_steps);
    else
      k_spmm_f32<BB, TT, 0, 3><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off,
                                                                  y, x_prev, out, y_copy, alpha, gamma, center, step, n_steps);
  } else if (io == 0)
    k_spmm_f32<B, TPR, 0><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off,
                                                             y, x_prev, out, y_copy, alpha, gamma, center, step, n_steps);
  else if (io == 1)
    k_spmm_f32<B, TPR, 1><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off,
                                                             y, x_prev, out, y_copy, alpha, gamma, center, step, n_steps);
  else
    k_spmm_f32<B, TPR, 2><<<grid, SPMM_THREADS, 0, stream>>>(g.row_ptr, g.cols, g.weights_f, g.ddi_f, g.mesh_off,
                                                             y, x_prev, out, y_copy, alpha, gamma, center, step, n_steps);
  FB_COUNT_LAUNCH(1);
  return FB_OK;
}

int launch_spmm_f32(int io, int b, const SpmmGraph& g, const void* y, const float* x_prev, void* out, float* y_copy,
                    const double* alpha, const double* gamma, const double* center, int step, int n_steps,
                    cudaStream_t stream) {
#define FB_CASE(BB, TT) \
  case BB:              \
    return launch_spmm_f32_b<BB, TT>(io, g, y, x_prev, out, y_copy, alpha, gamma, center, step, n_steps, stream);
  switch (b) {
    FB_CASE(8, 2)
    FB_CASE(16, 4)
    FB_CASE(24, 2)
    FB_CASE(32, 4)
    FB_CASE(40, 2)
    FB_CASE(48, 4)
    FB_CASE(56, 2)
    FB_CASE(64, 4)
    FB_CASE(72, 2)
    FB_CASE(80, 4)
    FB_CASE(88, 2)
    FB_CASE(96, 4)
    default:
      set_error("spmm (fp32): unsupported block size %d (multiples of 8 up to 96)", b);
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

// ---------------------------------------------------------------------------------------------
// Graph.mean_filter_graph (graph.py:349-354).  scipy stores each row of
// average_mat = diag(1/(1+d)) @ (A + I) in DESCENDING column order and `average_mat @ x`
// accumulates y += a*x in stored order (multiply, then add); one thread per row walks the row of A
// backwards and splices the diagonal in at its sorted position, which reproduces that bit for bit.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256)
k_mean_filter(const int* __restrict__ row_ptr, const int* __restrict__ cols, const double* __restrict__ weights,
              const double* __restrict__ degree, int row_begin, int row_end, const double* __restrict__ x,
              double* __restrict__ out, int n_cols_rt) {
  const int i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row_end) return;
  const int nc = C > 0 ? C : n_cols_rt;
  const double dsm = FB_DIV(1.0, FB_ADD(1.0, degree[i]));
  const int p0 = row_ptr[i], p1 = row_ptr[i + 1];
  constexpr int CMAX = C > 0 ? C : 8;
  double acc[CMAX];
#pragma unroll
  for (int k = 0; k < CMAX; ++k) acc[k] = 0.0;
  bool diag_done = false;
  for (int p = p1 - 1; p >= p0; --p) {
    const int j = cols[p];
    if (!diag_done && j < i) {
      const double* xi = x + (size_t)i * nc;
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < nc) acc[k] = FB_ADD(acc[k], FB_MUL(dsm, xi[k]));
      diag_done = true;
    }
    double val;
    if (j == i) {
      val = FB_MUL(dsm, FB_ADD(weights[p], 1.0));
      diag_done = true;
    } else {
      val = FB_MUL(dsm, weights[p]);
    }
    const double* xj = x + (size_t)j * nc;
#pragma unroll
    for (int k = 0; k < CMAX; ++k)
      if (k < nc) acc[k] = FB_ADD(acc[k], FB_MUL(val, xj[k]));
  }
  if (!diag_done) {
    const double* xi = x + (size_t)i * nc;
#pragma unroll
    for (int k = 0; k < CMAX; ++k)
      if (k < nc) acc[k] = FB_ADD(acc[k], FB_MUL(dsm, xi[k]));
  }
  double* o = out + (size_t)i * nc;
#pragma unroll
  for (int k = 0; k < CMAX; ++k)
    if (k < nc) o[k] = acc[k];
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

__global__ void k_copy_rows(const double* __restrict__ in, double* __restrict__ out, long long begin, long long end) {
  const long long t = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < end) out[t] = in[t];
}

}  // namespace fb

using namespace fb;

extern "C" {

int focusr_set_tuning(int key, int value) {
  if (key == 0) {
    fb::g_spmm_variant = value;
    return 0;
  }
  if (key == 1) {
    fb::g_l2_budget_mb = value;
    return 0;
  }
  if (key == 2) {
    fb::g_smooth_variant = value;
    return 0;
  }
  if (key == 3) {
    fb::g_mixed_precision = value;
    return 0;
  }
  if (key == 4) {
    fb::g_spmm_prefetch = value;
    return 0;
  }
  if (key == 5) {
    fb::g_spmm_prefetch_f32 = value;
    return 0;
  }
  if (key == 6) {
    fb::g_spmm_hint = value;
    return 0;
  }
  if (key == 7) {
    fb::g_spmm_minb = value;
    return 0;
  }
  fb::set_error("set_tuning: unknown key %d", key);
  return fb::FB_ERR_ARG;
}

int focusr_mean_filter(const int* row_ptr, const int* cols, const double* weights, const double* degree,
                       int row_begin, int row_end, const double* values_in, double* values_out,
                       double* scratch, int n_cols, int iterations, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(row_end > row_begin && n_cols >= 1 && n_cols <= 8 && iterations >= 0,
             "mean_filter: need rows, 1 <= n_cols <= 8, iterations >= 0");
  const int n = row_end - row_begin;
  const int T = 256;
  if (iterations == 0) {
    const long long b = (long long)row_begin * n_cols, e = (long long)row_end * n_cols;
    k_copy_rows<<<div_up(e - b, T), T, 0, stream>>>(values_in, values_out, b, e);
    FB_COUNT_LAUNCH(1);
    FB_LAUNCH_CHECK();
    return FB_OK;
  }
  // ping-pong so that the last iteration writes values_out
  const double* src = values_in;
  for (int it = 0; it < iterations; ++it) {
    double* dst = ((iterations - 1 - it) % 2 == 0) ? values_out : scratch;
    if (n_cols == 3)
      k_mean_filter<3><<<div_up(n, T), T, 0, stream>>>(row_ptr, cols, weights, degree, row_begin, row_end, src, dst, 3);
    else if (n_cols == 1)
      k_mean_filter<1><<<div_up(n, T), T, 0, stream>>>(row_ptr, cols, weights, degree, row_begin, row_end, src, dst, 1);
    else
      k_mean_filter<0><<<div_up(n, T), T, 0, stream>>>(row_ptr, cols, weights, degree, row_begin, row_end, src, dst, n_cols);
    src = dst;
  }
  FB_COUNT_LAUNCH(iterations);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

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
