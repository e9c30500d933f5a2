// The fp32 Chebyshev filter steps of the eigensolver on a sliced-ELL (SELL-64) copy of the matrix, and the smoothing
// passes of Graph.mean_filter_graph.  Both are chains of hundreds of dependent, HBM-bound launches.
//
// (1) Filter steps (chfsi_driver.hpp: plain fp32 blocks and the fp32 correction form).  Round 1 ran them on CSR: per row
//     a dependent chain row_ptr -> (col, weight) -> gather, two 4-byte loads per entry whose addresses are 6 entries
//     apart between neighbouring rows, and ncu showed the L1TEX wavefront pipe 61% busy next to DRAM at 69%.  Here the
//     fp32 copy of the matrix is stored slice by slice: 64 rows x W entries, entry k of the 64 rows contiguous, {column,
//     weight} packed in 8 bytes.  One coalesced load per k serves 8 rows of a warp, there is no row pointer, the trip
//     count W is uniform over the CTA pass, and the single-use streams (entries, z_prev, r, the stores) bypass L1 and
//     are marked evict_first in L2 so that the gathered block, which is re-used ~7 times, survives there.
//     Measured on B200, 128 pairs per launch (tools/kernel_ab.py, gpurun_out/r2_ab1.log): correction step 0.2494 ms (CSR)
//     -> 0.2793 (SELL, default cache policy) -> 0.2277 ms (SELL + stream policy) = 5.40 TB/s = 82.7% of the measured HBM
//     peak by algorithmic bytes; plain fp32 step 0.2024 -> 0.2003 ms (75%).  evict_last on the gathered block instead:
//     0.2430 / 0.2374 ms (worse for the plain step); both (the default): 0.2298 / 0.1985.  6 or 5 resident CTAs per SM
//     (40 / 46 registers): 0.2537 / 0.2593 ms; without the L2 prefetch of the CTA's streams: 0.2696 ms.  The absolute
//     numbers move by +-8% from one B200 of the pool to the next (r2_ab1..3.log); the order of the variants does not,
//     except that stream policy alone vs both is a toss-up for the correction step and both is never worse for the
//     plain step (0.1855 vs 0.2210, 0.1864 vs 0.1886, 0.1985 vs 0.2003 ms on three boxes).  The CSR kernels are gone.
// (2) Programmatic dependent launch: every step of a chain is launched with programmaticStreamSerialization; a CTA
//     first asks L2 for the data of its rows that the previous step does not write (matrix slice, r, z_prev), then
//     executes griddepcontrol.wait, then griddepcontrol.launch_dependents.  The next step's CTAs therefore become
//     resident as the last wave of this step drains and their prefetches overlap its tail; no launch gap is left.
// (3) Graph.mean_filter_graph (graph.py:349-354): one thread per row walks the CSR row of A backwards and splices the
//     diagonal in at its sorted position -- scipy's accumulation order, bit for bit.  A sliced-ELL form of the smoothing
//     matrix (entries pre-multiplied, padded [n][4] iterates fetched with one 256-bit load) was measured at 18.6 / 17.95 ms
//     against 17.8 ms for 300 passes over 128 meshes and removed: the pass is bound by its 6.4-wave launches, not by
//     the row walk (programmatic dependent launch: 17.8 -> 17.0 ms).  Marking the matrix stream evict_first in L2, which
//     pays in the filter steps, costs here (18.3 -> 20.0 ms): the 140 MB matrix of 128 targets is partly L2-resident
//     from one pass to the next.
#include <vector>

#include "common.cuh"
#include "rowops.h"
#include "sell.cuh"

namespace fb {

constexpr int FS_THREADS = 256;

// programmatic dependent launch (see header comment)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
constexpr int FS_ROWS = 256;  // rows per CTA = 4 slices

// ---------------------------------------------------------------------------------------------------------------
// cache-policy helpers (sm_100a PTX): 128-bit loads take the eviction priority through a createpolicy descriptor
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// gathered (re-used) data
template <int HINT>
__device__ __forceinline__ float4 ld_keep(const float4* p, unsigned long long pol) {
  if (HINT) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
  }
  return __ldg(p);
}
template <int HINT>
__device__ __forceinline__ double2 ld_keep(const double2* p, unsigned long long pol) {
  if (HINT) {
    double2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
  }
  return __ldg(p);
}
// single-use streams
template <int HINT>
__device__ __forceinline__ float4 ld_stream(const float4* p, unsigned long long pol) {
  if (HINT) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
  }
  return __ldg(p);
}
template <int HINT>
__device__ __forceinline__ int2 ld_stream(const int2* p, unsigned long long pol) {
  if (HINT) {
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
  }
  return __ldg(p);
}
template <int HINT>
__device__ __forceinline__ void st_stream(float4* p, float4 v, unsigned long long pol) {
  if (HINT)
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
  else
    *p = v;
}
__device__ __forceinline__ void prefetch_l2_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---------------------------------------------------------------------------------------------------------------
// SELL build (fp32 filter copy)
// ---------------------------------------------------------------------------------------------------------------
long long sell_entries_cap(const int* off, const int* info, int n_meshes, int extra_per_row) {
  long long cap = 0;
  for (int m = 0; m < n_meshes; ++m) {
    const long long rows = off[m + 1] - off[m];
    const long long slices = (rows + SELL_ROWS - 1) / SELL_ROWS;
    cap += slices * SELL_ROWS * (long long)(info[FOCUSR_MESH_INFO_INTS * m + 4] + extra_per_row);
  }
  return cap;
}

int sell_slice_count(const int* off, int n_meshes) {
  long long s = 0;
  for (int m = 0; m < n_meshes; ++m) s += (off[m + 1] - off[m] + SELL_ROWS - 1) / SELL_ROWS;
  return (int)s;
}

__device__ __forceinline__ int find_slice_mesh(const int* __restrict__ mso, int n_meshes, int s) {
  int lo = 0, hi = n_meshes - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (mso[mid] <= s)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// one warp per slice: 64 * (longest row of the slice)
__global__ void __launch_bounds__(256)
k_sell_widths(const int* __restrict__ row_ptr, const int* __restrict__ mesh_off, const int* __restrict__ mso, int n_meshes,
              int n_slices, int* __restrict__ slice_cnt) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= n_slices) return;
  const int m = find_slice_mesh(mso, n_meshes, s);
  const int r0 = mesh_off[m] + (s - mso[m]) * SELL_ROWS, r1 = min(mesh_off[m + 1], r0 + SELL_ROWS);
  int w = 0;
  for (int r = r0 + lane; r < r1; r += 32) w = max(w, row_ptr[r + 1] - row_ptr[r]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
  if (lane == 0) slice_cnt[s] = w * SELL_ROWS;
}

// one thread per (slice, row slot)
__global__ void __launch_bounds__(256)
k_sell_fill_f32(const int* __restrict__ row_ptr, const int* __restrict__ cols, const double* __restrict__ weights,
                const int* __restrict__ mesh_off, const int* __restrict__ mso, int n_meshes, int n_slices,
                const int* __restrict__ slice_ptr, int2* __restrict__ entries) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int s = (int)(t / SELL_ROWS), rl = (int)(t % SELL_ROWS);
  if (s >= n_slices) return;
  const int m = find_slice_mesh(mso, n_meshes, s);
  const int r0 = mesh_off[m] + (s - mso[m]) * SELL_ROWS, r1 = mesh_off[m + 1];
  const int e0 = slice_ptr[s], w = (slice_ptr[s + 1] - e0) / SELL_ROWS;
  const int row = r0 + rl;
  int p0 = 0, len = 0;
  if (row < r1) {
    p0 = row_ptr[row];
    len = row_ptr[row + 1] - p0;
  }
  for (int k = 0; k < w; ++k) {
    int2 e = make_int2(r0, 0);  // padding: a valid row of the same slice, weight +0.0f
    if (k < len) e = make_int2(cols[p0 + k], __float_as_int((float)weights[p0 + k]));
    entries[(size_t)e0 + (size_t)k * SELL_ROWS + rl] = e;
  }
}

__global__ void k_ddi_f32(const double* __restrict__ degree, const double* __restrict__ degree_inv, int n_rows,
                          float2* __restrict__ ddi) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_rows) ddi[t] = make_float2((float)degree[t], (float)degree_inv[t]);
}

int sell_build_f32(const int* row_ptr, const int* cols, const double* weights, const double* degree,
                   const double* degree_inv, const int* mesh_off, const int* off_host, int n_meshes, int n_rows,
                   int* mesh_slice_off, int* slice_ptr, int2* entries, float2* ddi, int* slice_cnt, int* scan_tmp,
                   cudaStream_t stream) {
  // first slice of every mesh: tiny, built on the host and staged by the runtime (pageable source)
  static thread_local std::vector<int> mso_host;
  mso_host.assign((size_t)n_meshes + 1, 0);
  for (int m = 0; m < n_meshes; ++m) mso_host[m + 1] = mso_host[m] + (off_host[m + 1] - off_host[m] + SELL_ROWS - 1) / SELL_ROWS;
  const int n_slices = mso_host[n_meshes];
  FB_CUDA(cudaMemcpyAsync(mesh_slice_off, mso_host.data(), sizeof(int) * ((size_t)n_meshes + 1), cudaMemcpyHostToDevice, stream));
  k_sell_widths<<<div_up((long long)n_slices * 32, 256), 256, 0, stream>>>(row_ptr, mesh_off, mesh_slice_off, n_meshes, n_slices,
                                                                           slice_cnt);
  FB_COUNT_LAUNCH(1);
  int rc = exclusive_scan_i32(slice_cnt, slice_ptr, n_slices, scan_tmp, stream);
  if (rc) return rc;
  k_sell_fill_f32<<<div_up((long long)n_slices * SELL_ROWS, 256), 256, 0, stream>>>(row_ptr, cols, weights, mesh_off, mesh_slice_off,
                                                                                    n_meshes, n_slices, slice_ptr, entries);
  k_ddi_f32<<<div_up(n_rows, 256), 256, 0, stream>>>(degree, degree_inv, n_rows, ddi);
  FB_COUNT_LAUNCH(2);
  FB_LAUNCH_CHECK();
  return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// filter step on the SELL copy.  TPR threads share a row, each owns B / (4 TPR) float4 slices of it.
// MODE 0: out = al (L y - c y) - ga x_prev                    (fp32 blocks; al, ga per mesh: double tables)
// MODE 1: the same with y read from the fp64 block, ga = 0; also writes the fp32 copy of y (next step's x_prev)
// MODE 2: the same, result written to the fp64 block
// MODE 3: z_next = al_j ((L - c) z + r_j) - ga_j z_prev       (correction form; al, ga per column: float tables)
// MODE 4: x += that (fp64 block), nothing else stored
// ---------------------------------------------------------------------------------------------------------------
template <int B, int TPR, int MODE, int POL, int MINB>
__global__ void __launch_bounds__(FS_THREADS, MINB)
k_filter_sell(const int2* __restrict__ entries, const int* __restrict__ slice_ptr, const int* __restrict__ mso,
              const float2* __restrict__ ddi, const int* __restrict__ mesh_off, const void* __restrict__ y_,
              const float* __restrict__ x_prev, const float* __restrict__ r, void* __restrict__ out_, float* __restrict__ y_copy,
              const void* __restrict__ alpha_, const void* __restrict__ gamma_, const double* __restrict__ center, int step,
              int n_steps, int has_prev, int prefetch, int early) {
  constexpr int VPT = B / (4 * TPR);
  static_assert(VPT * 4 * TPR == B, "block size must be a multiple of 4*TPR");
  constexpr bool CORR = MODE >= 3;
  constexpr int KEEP = POL & 1, STRM = (POL >> 1) & 1;
  const int mesh = blockIdx.y;
  const int r0 = mesh_off[mesh] + blockIdx.x * FS_ROWS;
  const int r1 = min(mesh_off[mesh + 1], r0 + FS_ROWS);
  if (r0 >= r1) return;
  const int slice0 = mso[mesh] + blockIdx.x * (FS_ROWS / SELL_ROWS);
  const unsigned long long pol_keep = KEEP ? policy_evict_last() : 0ull;
  const unsigned long long pol_strm = STRM ? policy_evict_first() : 0ull;
  const float* yf = static_cast<const float*>(y_);
  // Before the dependency on the previous step resolves: ask L2 for what that step does not write -- this CTA's
  // matrix slice, its r rows and its z_prev rows (written two steps back: complete, because the previous step released
  // this launch only after ITS wait returned).  After it: the rows of y this CTA owns.
  const int pr = r0 + (int)threadIdx.x;
  constexpr int LINES = (B * 4 + 127) / 128;  // 128-byte lines per fp32 row
  if (!early) {
    pdl_wait();
    pdl_launch_dependents();
  }
  if (prefetch && MODE != 1) {
    if (pr < r1) {
#pragma unroll
      for (int l = 0; l < LINES; ++l) {
        if (has_prev) prefetch_l2_line(x_prev + (size_t)pr * B + 32 * l);
        if (CORR) prefetch_l2_line(r + (size_t)pr * B + 32 * l);
      }
    }
    const int ns = min(FS_ROWS / SELL_ROWS, (r1 - r0 + SELL_ROWS - 1) / SELL_ROWS);
    const int q0 = slice_ptr[slice0], q1 = slice_ptr[slice0 + ns];
    for (int q = q0 + 16 * (int)threadIdx.x; q < q1; q += 16 * FS_THREADS) prefetch_l2_line(entries + q);
  }
  if (early) {
    pdl_wait();
    pdl_launch_dependents();
  }
  if (prefetch && MODE != 1 && pr < r1) {
#pragma unroll
    for (int l = 0; l < LINES; ++l) prefetch_l2_line(yf + (size_t)pr * B + 32 * l);
  }
  const float cc = (float)center[mesh];
  const int g = threadIdx.x / TPR, t = threadIdx.x % TPR;
  float al_s = 0.f, ga_s = 0.f;
  const float4* al_p = nullptr;
  const float4* ga_p = nullptr;
  if (CORR) {
    al_p = reinterpret_cast<const float4*>(static_cast<const float*>(alpha_) + ((size_t)mesh * n_steps + step) * B);
    ga_p = reinterpret_cast<const float4*>(static_cast<const float*>(gamma_) + ((size_t)mesh * n_steps + step) * B);
  } else {
    al_s = (float)static_cast<const double*>(alpha_)[(size_t)mesh * n_steps + step];
    ga_s = (float)static_cast<const double*>(gamma_)[(size_t)mesh * n_steps + step];
  }
  auto load_y = [&](int rr, int slice) -> float4 {
    if (MODE == 1) {
      const double2* src = reinterpret_cast<const double2*>(static_cast<const double*>(y_) + (size_t)rr * B) + 2 * slice;
      const double2 a = ld_keep<KEEP>(src, pol_keep), b = ld_keep<KEEP>(src + 1, pol_keep);
      return make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
    } else {
      return ld_keep<KEEP>(reinterpret_cast<const float4*>(yf + (size_t)rr * B) + slice, pol_keep);
    }
  };
  // a CTA pass covers RP = 256 / TPR rows = RP / 64 slices; a warp never straddles slices
  constexpr int RP = FS_THREADS / TPR, SPP = RP / SELL_ROWS;
  static_assert(RP % SELL_ROWS == 0, "a CTA pass must cover whole slices");
  const int gs = g / SELL_ROWS, rl = g % SELL_ROWS;
#pragma unroll 1
  for (int pass = 0; pass < FS_ROWS / RP; ++pass) {
    const int rb = r0 + pass * RP + gs * SELL_ROWS;  // first row of this thread's slice
    if (rb >= r1) break;
    const int row = rb + rl;
    const int e0 = slice_ptr[slice0 + pass * SPP + gs];
    const int w = (slice_ptr[slice0 + pass * SPP + gs + 1] - e0) / SELL_ROWS;
    const int2* ep = entries + (size_t)e0 + rl;
    float4 acc[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int k = 0; k < w; ++k) {
      const int2 e = ld_stream<STRM>(ep + (size_t)k * SELL_ROWS, pol_strm);
      const float wt = __int_as_float(e.y);
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const float4 a = load_y(e.x, t + v * TPR);
        acc[v].x = fmaf(wt, a.x, acc[v].x);
        acc[v].y = fmaf(wt, a.y, acc[v].y);
        acc[v].z = fmaf(wt, a.z, acc[v].z);
        acc[v].w = fmaf(wt, a.w, acc[v].w);
      }
    }
    if (row >= r1) continue;
    const float2 dd = ddi[row];
    const float d = dd.x, di = dd.y;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int slice = t + v * TPR;
      const float4 yv = load_y(row, slice);
      float4 o;
      if (CORR) {
        const float4 rv = ld_stream<STRM>(reinterpret_cast<const float4*>(r + (size_t)row * B) + slice, pol_strm);
        const float4 al = __ldg(al_p + slice);
        o.x = al.x * ((di * (d * yv.x - acc[v].x) - cc * yv.x) + rv.x);
        o.y = al.y * ((di * (d * yv.y - acc[v].y) - cc * yv.y) + rv.y);
        o.z = al.z * ((di * (d * yv.z - acc[v].z) - cc * yv.z) + rv.z);
        o.w = al.w * ((di * (d * yv.w - acc[v].w) - cc * yv.w) + rv.w);
        if (has_prev) {
          const float4 ga = __ldg(ga_p + slice);
          const float4 pv = ld_stream<STRM>(reinterpret_cast<const float4*>(x_prev + (size_t)row * B) + slice, pol_strm);
          o.x -= ga.x * pv.x;
          o.y -= ga.y * pv.y;
          o.z -= ga.z * pv.z;
          o.w -= ga.w * pv.w;
        }
      } else {
        o.x = al_s * (di * (d * yv.x - acc[v].x) - cc * yv.x);
        o.y = al_s * (di * (d * yv.y - acc[v].y) - cc * yv.y);
        o.z = al_s * (di * (d * yv.z - acc[v].z) - cc * yv.z);
        o.w = al_s * (di * (d * yv.w - acc[v].w) - cc * yv.w);
        if (MODE != 1 && ga_s != 0.f) {
          const float4 pv = ld_stream<STRM>(reinterpret_cast<const float4*>(x_prev + (size_t)row * B) + slice, pol_strm);
          o.x -= ga_s * pv.x;
          o.y -= ga_s * pv.y;
          o.z -= ga_s * pv.z;
          o.w -= ga_s * pv.w;
        }
      }
      if (MODE == 2) {
        double2* op = reinterpret_cast<double2*>(static_cast<double*>(out_) + (size_t)row * B) + 2 * slice;
        op[0] = make_double2((double)o.x, (double)o.y);
        op[1] = make_double2((double)o.z, (double)o.w);
      } else if (MODE == 4) {
        double2* xo = reinterpret_cast<double2*>(static_cast<double*>(out_) + (size_t)row * B) + 2 * slice;
        double2 a = xo[0], b = xo[1];
        a.x += (double)o.x;
        a.y += (double)o.y;
        b.x += (double)o.z;
        b.y += (double)o.w;
        xo[0] = a;
        xo[1] = b;
      } else {
        st_stream<STRM>(reinterpret_cast<float4*>(static_cast<float*>(out_) + (size_t)row * B) + slice, o, pol_strm);
      }
      if (MODE == 1) reinterpret_cast<float4*>(y_copy + (size_t)row * B)[slice] = yv;
    }
  }
}

template <int B, int TPR, int MODE, int MINB>
static void launch_fs_pol(int pol, dim3 grid, cudaStream_t stream, const SellF32& m, const int* mesh_off, const void* y,
                          const float* x_prev, const float* r, void* out, float* y_copy, const void* alpha, const void* gamma,
                          const double* center, int step, int n_steps, int has_prev, int prefetch, int pdl) {
#define FB_FS_GO(P)                                                                                                        \
  if (pdl)                                                                                                                 \
    launch_pdl(k_filter_sell<B, TPR, MODE, P, MINB>, grid, dim3(FS_THREADS), stream, m.entries, m.slice_ptr,               \
               m.mesh_slice_off, m.ddi, mesh_off, y, x_prev, r, out, y_copy, alpha, gamma, center, step, n_steps, has_prev, \
               prefetch, pdl == 2 ? 1 : 0);                                                                                \
  else                                                                                                                     \
    k_filter_sell<B, TPR, MODE, P, MINB><<<grid, FS_THREADS, 0, stream>>>(m.entries, m.slice_ptr, m.mesh_slice_off, m.ddi,  \
                                                                          mesh_off, y, x_prev, r, out, y_copy, alpha, gamma, \
                                                                          center, step, n_steps, has_prev, prefetch, 0)
  switch (pol & 3) {
    case 0: FB_FS_GO(0); break;
    case 1: FB_FS_GO(1); break;
    case 2: FB_FS_GO(2); break;
    default: FB_FS_GO(3); break;
  }
#undef FB_FS_GO
}

template <int B, int TPR>
static int launch_fs_b(int mode, const SellF32& m, const int* mesh_off, int n_meshes, int max_mesh_rows, const void* y,
                       const float* x_prev, const float* r, void* out, float* y_copy, const void* alpha, const void* gamma,
                       const double* center, int step, int n_steps, bool has_prev, const FilterTuning& tune,
                       cudaStream_t stream) {
  dim3 grid(div_up(max_mesh_rows, FS_ROWS), n_meshes);
  constexpr int VPT = B / (4 * TPR);
  constexpr int MB = VPT == 1 ? 8 : (VPT == 2 ? 6 : 3);
  const int hp = has_prev ? 1 : 0, pf = tune.prefetch ? 1 : 0;
#define FB_FS_MODE(MODE_, MINB_) \
  launch_fs_pol<B, TPR, MODE_, MINB_>(tune.policy, grid, stream, m, mesh_off, y, x_prev, r, out, y_copy, alpha, gamma, center, step, n_steps, hp, pf, tune.pdl)
  // the occupancy A/B (6 or 5 resident CTAs instead of 8) exists for the two steady-state b = 16 kernels only
  if (B == 16 && (mode == 0 || mode == 3) && (tune.min_blocks == 6 || tune.min_blocks == 5)) {
    constexpr int BB = B == 16 ? B : 16, TT = B == 16 ? TPR : 4;
    if (mode == 0 && tune.min_blocks == 6) launch_fs_pol<BB, TT, 0, 6>(tune.policy, grid, stream, m, mesh_off, y, x_prev, r, out, y_copy, alpha, gamma, center, step, n_steps, hp, pf, tune.pdl);
    else if (mode == 0) launch_fs_pol<BB, TT, 0, 5>(tune.policy, grid, stream, m, mesh_off, y, x_prev, r, out, y_copy, alpha, gamma, center, step, n_steps, hp, pf, tune.pdl);
    else if (tune.min_blocks == 6) launch_fs_pol<BB, TT, 3, 6>(tune.policy, grid, stream, m, mesh_off, y, x_prev, r, out, y_copy, alpha, gamma, center, step, n_steps, hp, pf, tune.pdl);
    else launch_fs_pol<BB, TT, 3, 5>(tune.policy, grid, stream, m, mesh_off, y, x_prev, r, out, y_copy, alpha, gamma, center, step, n_steps, hp, pf, tune.pdl);
  } else {
    switch (mode) {
      case 0: FB_FS_MODE(0, MB); break;
      case 1: FB_FS_MODE(1, MB); break;
      case 2: FB_FS_MODE(2, MB); break;
      case 3: FB_FS_MODE(3, MB); break;
      default: FB_FS_MODE(4, MB); break;
    }
  }
#undef FB_FS_MODE
  FB_COUNT_LAUNCH(1);
  return FB_OK;
}

int launch_filter_sell(int mode, int b, const SellF32& m, const int* mesh_off, int n_meshes, int max_mesh_rows,
                       const void* y, const float* x_prev, const float* r, void* out, float* y_copy, const void* alpha,
                       const void* gamma, const double* center, int step, int n_steps, bool has_prev,
                       const FilterTuning& tune, cudaStream_t stream) {
#define FB_CASE(BB, TT) \
  case BB:              \
    return launch_fs_b<BB, TT>(mode, m, mesh_off, n_meshes, max_mesh_rows, y, x_prev, r, out, y_copy, alpha, gamma, center, step, n_steps, has_prev, tune, stream);
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
      set_error("filter (SELL): unsupported block size %d (multiples of 8 up to 96)", b);
      return FB_ERR_UNSUPPORTED;
  }
#undef FB_CASE
}

// ===============================================================================================================
// smoothing passes (Graph.mean_filter_graph)
// ===============================================================================================================
// scipy stores each row of average_mat = diag(1/(1+d)) @ (A + I) in DESCENDING column order and `average_mat @ x`
// accumulates y += a * x in stored order (multiply, then add); one thread per row (any column count up to 8) walks the
// row of A backwards and splices the diagonal in at its sorted position, which reproduces that bit for bit.  Passes are
// chained by programmatic dependent launch: the CTA prefetches its slice of the matrix, which no pass writes, into L2
// before it waits for the previous pass.
template <int C>
__global__ void __launch_bounds__(256)
k_mean_filter(const int* __restrict__ row_ptr, const int* __restrict__ cols, const double* __restrict__ weights,
              const double* __restrict__ degree, int row_begin, int row_end, const double* __restrict__ x,
              double* __restrict__ out, int n_cols_rt) {
  const int i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
  {
    const int c0 = row_begin + blockIdx.x * blockDim.x, c1 = min(row_end, c0 + (int)blockDim.x);
    const int q0 = row_ptr[c0], q1 = row_ptr[c1];
    for (int q = q0 + 32 * (int)threadIdx.x; q < q1; q += 32 * (int)blockDim.x) prefetch_l2_line(cols + q);
    for (int q = q0 + 16 * (int)threadIdx.x; q < q1; q += 16 * (int)blockDim.x) prefetch_l2_line(weights + q);
  }
  pdl_wait();
  pdl_launch_dependents();
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
    const double wp = weights[p];
    if (!diag_done && j < i) {
      const double* xi = x + (size_t)i * nc;
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < nc) acc[k] = FB_ADD(acc[k], FB_MUL(dsm, xi[k]));
      diag_done = true;
    }
    double val;
    if (j == i) {
      val = FB_MUL(dsm, FB_ADD(wp, 1.0));
      diag_done = true;
    } else {
      val = FB_MUL(dsm, wp);
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

__global__ void k_copy_rows(const double* __restrict__ in, double* __restrict__ out, long long begin, long long end) {
  const long long t = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < end) out[t] = in[t];
}

struct SmoothLayout {
  double *buf_a, *buf_b;
  size_t bytes;
};

static SmoothLayout smooth_layout(int n_rows, int n_cols, void* ws) {
  Carver cv(ws, (size_t)-1);
  SmoothLayout l;
  l.buf_a = cv.take<double>((size_t)n_rows * n_cols);
  l.buf_b = cv.take<double>((size_t)n_rows * n_cols);
  l.bytes = cv.used + 256;
  return l;
}

}  // namespace fb

using namespace fb;

extern "C" {

long long focusr_sell_entries_cap(const int* mesh_point_off_host, const int* mesh_info_host, int n_meshes) {
  return sell_entries_cap(mesh_point_off_host, mesh_info_host, n_meshes, 0);
}

size_t focusr_mean_filter_workspace_bytes(int n_rows, int n_cols) { return smooth_layout(n_rows, n_cols, nullptr).bytes; }

int focusr_mean_filter(const int* row_ptr, const int* cols, const double* weights, const double* degree,
                       int row_begin, int row_end, const double* values_in, double* values_out, int n_cols,
                       int iterations, void* workspace, size_t workspace_bytes, focusr_stream_t stream_) {
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
  const SmoothLayout l = smooth_layout(n, n_cols, workspace);
  FB_REQUIRE(iterations == 1 || (workspace != nullptr && workspace_bytes >= l.bytes),
             "mean_filter: workspace too small (%zu < %zu)", workspace_bytes, l.bytes);
  // ping-pong between two buffers indexed by global row (shifted so that row_begin lands on their start)
  const size_t rb = (size_t)row_begin;
  double* pa = l.buf_a - rb * n_cols;
  double* pb = l.buf_b - rb * n_cols;
  const double* src = values_in;
  const dim3 grid(div_up(n, T)), block(T);
  for (int it = 0; it < iterations; ++it) {
    double* dst = it == iterations - 1 ? values_out : ((it & 1) ? pb : pa);
    cudaError_t e;
    if (n_cols == 3)
      e = launch_pdl(k_mean_filter<3>, grid, block, stream, row_ptr, cols, weights, degree, row_begin, row_end, src, dst, 3);
    else if (n_cols == 1)
      e = launch_pdl(k_mean_filter<1>, grid, block, stream, row_ptr, cols, weights, degree, row_begin, row_end, src, dst, 1);
    else
      e = launch_pdl(k_mean_filter<0>, grid, block, stream, row_ptr, cols, weights, degree, row_begin, row_end, src, dst, n_cols);
    FB_CUDA(e);
    src = dst;
  }
  FB_COUNT_LAUNCH(iterations);
  FB_LAUNCH_CHECK();
  return FB_OK;
}
}
