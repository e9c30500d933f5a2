// K5, persistent form: all smoothing passes of one mesh inside ONE thread-block cluster.
//
// Graph.mean_filter_graph (graph.py:320-354) applies M = diag(1/(1+d)) (A + I) 300 (+40) times to an [N][3]
// array.  As separate launches every pass streams the mesh's matrix (12 B per entry) and vectors from HBM / L2.
// Here a cluster of 16 CTAs (non-portable size, opt-in) owns one mesh for the whole call: each CTA keeps its
// slice of the vector (ping-pong) AND its slice of the matrix -- entries already multiplied by 1/(1+d), in
// scipy's accumulation order with the diagonal spliced in, neighbour ids packed as (owner CTA, offset) -- in
// shared memory; a pass gathers neighbour values through distributed shared memory (local slice: 38 cycles,
// remote slice: ~215 cycles) and ends with one cluster barrier.  No global-memory traffic inside the loop.
// The arithmetic per row is the one of k_mean_filter (same operands, same order, no FMA): bit-identical output.
//
// MEASURED (128 meshes x 15 212 vertices, 300 passes; tools/smooth_bench.py): 16.6 ms against 17.8 ms for one launch
// per pass -- NOT the 5-8x the on-chip residency suggests, so the per-pass kernels stay the default
// (DeviceGraph.mean_filter(cluster=False)).  Per pass a CTA spends ~3700 cycles on purely local shared-memory
// work (36 bytes of matrix + vector per entry = ~1900 cycles of 128 B/clk shared-memory wavefronts as the floor,
// 64-bit gathers issue as generic loads), +1700 cycles when ~15% of the gathers cross CTAs (distributed shared
// memory moves ~20 B/clk per SM), +850-1400 cycles per cluster barrier (UCGABAR + MEMBAR.ALL.GPU in SASS); only 9
// clusters of 16 fit the 148 SMs.  Streaming the matrix from L2 each pass, as the per-pass kernel does, costs
// about the same per SM-cycle.  Kept as the tested, bit-identical alternative and as the record of the experiment.
// The matrix slice is stored entry-major (ELL: entry q of row r at [q][r]) so that the 32 rows of a warp read 32
// consecutive words -- row-major storage costs an 8- to 16-way bank conflict per entry, measured 10x slower.
// Rows with more entries than the ELL width fall back to reading the CSR from global memory inside the same
// loop, so any degree distribution is handled.
#include <cooperative_groups.h>

#include "common.cuh"
#include "rowops.h"

namespace cg = cooperative_groups;

namespace fb {

int g_smooth_variant = 2;  // tuning key 2: 0 = 1024 threads x 4-entry gather batches, 1 = 512 x 8, 2 = 512 x 4, 3 = 1024 x 8

template <int C, int MFC_T, int MFC_U>
__global__ void __launch_bounds__(MFC_T)
k_mean_filter_cluster(const int* __restrict__ row_ptr, const int* __restrict__ cols, const double* __restrict__ weights,
                      const double* __restrict__ degree, const int* __restrict__ mesh_off, int mesh_begin,
                      const double* __restrict__ x_in, double* __restrict__ x_out, int iterations, int per_cap, int ell_w) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int mesh = mesh_begin + (int)blockIdx.x / CL;
  const int m0 = mesh_off[mesh], m1 = mesh_off[mesh + 1], n = m1 - m0;
  const int per = (n + CL - 1) / CL;
  const int r0 = min(m1, m0 + rank * per), r1 = min(m1, r0 + per);
  const int nrows = r1 - r0;
  extern __shared__ __align__(16) unsigned char smraw[];
  double* buf = reinterpret_cast<double*>(smraw);   // [2][per_cap * C]
  double* vals = buf + 2 * (size_t)per_cap * C;     // [ell_w][per_cap]
  int* ecol = reinterpret_cast<int*>(vals + (size_t)ell_w * per_cap);  // [ell_w][per_cap]   (owner << 24) | offset
  int* rlen = ecol + (size_t)ell_w * per_cap;       // [per_cap]   entries of the row, -1 = not in shared memory
  const int t = threadIdx.x;
  for (int i = t; i < nrows * C; i += MFC_T) buf[i] = x_in[(size_t)r0 * C + i];
  // entry lists in k_mean_filter's processing order (descending columns, the diagonal spliced in)
  for (int r = t; r < nrows; r += MFC_T) {
    const int i = r0 + r;
    const int p0 = row_ptr[i], p1 = row_ptr[i + 1];
    if (p1 - p0 + 1 > ell_w) {
      rlen[r] = -1;
      continue;
    }
    const double dsm = FB_DIV(1.0, FB_ADD(1.0, degree[i]));
    int w = 0;
    bool diag_done = false;
    const int self = ((i - m0) / per << 24) | ((i - m0) % per);
    for (int p = p1 - 1; p >= p0; --p) {
      const int j = cols[p];
      if (!diag_done && j < i) {
        vals[(size_t)w * per_cap + r] = dsm;
        ecol[(size_t)w * per_cap + r] = self;
        ++w;
        diag_done = true;
      }
      double v;
      if (j == i) {
        v = FB_MUL(dsm, FB_ADD(weights[p], 1.0));
        diag_done = true;
      } else {
        v = FB_MUL(dsm, weights[p]);
      }
      const int jl = j - m0;
      vals[(size_t)w * per_cap + r] = v;
      ecol[(size_t)w * per_cap + r] = (jl / per << 24) | (jl % per);
      ++w;
    }
    if (!diag_done) {
      vals[(size_t)w * per_cap + r] = dsm;
      ecol[(size_t)w * per_cap + r] = self;
      ++w;
    }
    rlen[r] = w;
  }
  cluster.sync();
  for (int it = 0; it < iterations; ++it) {
    double* src = buf + (size_t)(it & 1) * per_cap * C;
    double* dst = buf + (size_t)((it + 1) & 1) * per_cap * C;
    for (int r = t; r < nrows; r += MFC_T) {
      double acc[C];
#pragma unroll
      for (int k = 0; k < C; ++k) acc[k] = 0.0;
      const int len = rlen[r];
      if (len >= 0) {
        // gather first, accumulate after: the loads of MFC_U entries are independent and overlap their
        // (distributed) shared-memory latency; the additions then run in the stored order
        for (int q0 = 0; q0 < len; q0 += MFC_U) {
          double v[MFC_U], xv[MFC_U][C];
#pragma unroll
          for (int u = 0; u < MFC_U; ++u) {
            if (q0 + u < len) {
              const int pc = ecol[(size_t)(q0 + u) * per_cap + r];
              v[u] = vals[(size_t)(q0 + u) * per_cap + r];
              // own slice: ordinary shared-memory loads; only a neighbour in another CTA's slice goes through
              // the cluster's distributed-shared-memory path
              const int owner = pc >> 24;
              const double* xj = (owner == rank ? src : cluster.map_shared_rank(src, owner)) + (size_t)(pc & 0xffffff) * C;
#pragma unroll
              for (int k = 0; k < C; ++k) xv[u][k] = xj[k];
            }
          }
#pragma unroll
          for (int u = 0; u < MFC_U; ++u) {
            if (q0 + u < len) {
#pragma unroll
              for (int k = 0; k < C; ++k) acc[k] = FB_ADD(acc[k], FB_MUL(v[u], xv[u][k]));
            }
          }
        }
      } else {  // this row's entries did not fit: same arithmetic from the CSR in global memory
        const int i = r0 + r;
        const double dsm = FB_DIV(1.0, FB_ADD(1.0, degree[i]));
        const int p0 = row_ptr[i], p1 = row_ptr[i + 1];
        bool diag_done = false;
        const double* xi = src + (size_t)r * C;
        for (int p = p1 - 1; p >= p0; --p) {
          const int j = cols[p];
          if (!diag_done && j < i) {
#pragma unroll
            for (int k = 0; k < C; ++k) acc[k] = FB_ADD(acc[k], FB_MUL(dsm, xi[k]));
            diag_done = true;
          }
          double v;
          if (j == i) {
            v = FB_MUL(dsm, FB_ADD(weights[p], 1.0));
            diag_done = true;
          } else {
            v = FB_MUL(dsm, weights[p]);
          }
          const int jl = j - m0;
          const int owner = jl / per;
          const double* xj = (owner == rank ? src : cluster.map_shared_rank(src, owner)) + (size_t)(jl % per) * C;
#pragma unroll
          for (int k = 0; k < C; ++k) acc[k] = FB_ADD(acc[k], FB_MUL(v, xj[k]));
        }
        if (!diag_done) {
#pragma unroll
          for (int k = 0; k < C; ++k) acc[k] = FB_ADD(acc[k], FB_MUL(dsm, xi[k]));
        }
      }
#pragma unroll
      for (int k = 0; k < C; ++k) dst[(size_t)r * C + k] = acc[k];
    }
    cluster.sync();
  }
  const double* fin = buf + (size_t)(iterations & 1) * per_cap * C;
  for (int i = t; i < nrows * C; i += MFC_T) x_out[(size_t)r0 * C + i] = fin[i];
}

template <int C, int MFC_T, int MFC_U>
static int launch_cluster(const int* row_ptr, const int* cols, const double* weights, const double* degree,
                          const int* mesh_off, int mesh_begin, int n_meshes, int max_mesh_points, const double* x_in,
                          double* x_out, int iterations, cudaStream_t stream) {
  const int CL = 16;
  const int per_cap = (max_mesh_points + CL - 1) / CL;
  if (per_cap >= (1 << 24)) return 1;
  const size_t limit = 200 * 1024;
  const size_t fixed = sizeof(double) * 2 * (size_t)per_cap * C + sizeof(int) * (size_t)per_cap + 16;
  if (fixed + (size_t)12 * per_cap * 4 > limit) return 1;  // not even 4 entries per row fit: not applicable
  int ell_w = (int)((limit - fixed) / ((size_t)12 * per_cap));
  if (ell_w > 16) ell_w = 16;
  const size_t smem = fixed + (size_t)12 * per_cap * ell_w;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(k_mean_filter_cluster<C, MFC_T, MFC_U>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(k_mean_filter_cluster<C, MFC_T, MFC_U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit) != cudaSuccess) {
      cudaGetLastError();
      return 1;
    }
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(CL * n_meshes), 1, 1);
  cfg.blockDim = dim3(MFC_T, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&n_clusters, k_mean_filter_cluster<C, MFC_T, MFC_U>, &cfg) != cudaSuccess || n_clusters < 1) {
    cudaGetLastError();
    return 1;
  }
  if (cudaLaunchKernelEx(&cfg, k_mean_filter_cluster<C, MFC_T, MFC_U>, row_ptr, cols, weights, degree, mesh_off, mesh_begin, x_in, x_out,
                         iterations, per_cap, ell_w) != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  FB_COUNT_LAUNCH(1);
  return 0;
}

}  // namespace fb

using namespace fb;

extern "C" {

// returns FB_OK, or FB_ERR_UNSUPPORTED when the cluster form does not apply (caller uses focusr_mean_filter)
int focusr_mean_filter_meshes(const int* row_ptr, const int* cols, const double* weights, const double* degree,
                              const int* mesh_point_off, int mesh_begin, int mesh_end, int max_mesh_points,
                              const double* values_in, double* values_out, int n_cols, int iterations,
                              focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(mesh_end > mesh_begin && max_mesh_points > 0 && iterations >= 1, "mean_filter_meshes: bad arguments");
  int rc = 1;
#define FB_MFC(CC, TT, UU)                                                                                         \
  rc = launch_cluster<CC, TT, UU>(row_ptr, cols, weights, degree, mesh_point_off, mesh_begin, mesh_end - mesh_begin, \
                                  max_mesh_points, values_in, values_out, iterations, stream)
  const int var = g_smooth_variant;
  if (n_cols == 3) {
    if (var == 1) FB_MFC(3, 512, 8);
    else if (var == 2) FB_MFC(3, 512, 4);
    else if (var == 3) FB_MFC(3, 1024, 8);
    else FB_MFC(3, 1024, 4);
  } else if (n_cols == 1) {
    if (var == 1 || var == 2) FB_MFC(1, 512, 8);
    else FB_MFC(1, 1024, 8);
  }
#undef FB_MFC
  if (rc != 0) {
    set_error("mean_filter_meshes: the cluster form does not apply (n_cols %d, %d points per mesh)", n_cols, max_mesh_points);
    return FB_ERR_UNSUPPORTED;
  }
  return FB_OK;
}

}  // extern "C"
