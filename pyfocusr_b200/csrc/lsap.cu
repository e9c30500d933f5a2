// scipy.optimize.linear_sum_assignment on the GPU, for the `hungarian` correspondence of the reference
// (focusr.py:340-349: cdist, then linear_sum_assignment of the N x N matrix -- minutes of one host core at 15k vertices).
//
// Algorithm: exactly scipy's (rectangular_lsap.cpp: the shortest augmenting path method of Crouse, IEEE TAES 2016) --
// rows are added one by one; for a row a Dijkstra-like search scans the unvisited columns, `r = min_val + cost[i][j] -
// u[i] - v[j]` in that order of operations, takes the column with the lowest tentative cost (ties: an unassigned column
// wins, the LAST one in scan order among unassigned ones, otherwise the FIRST in scan order; the scan order is that of
// scipy's `remaining` array with its swap-with-last removal), follows it to the row it is assigned to, and so on until
// an unassigned column (the sink) is reached; then the duals are updated and the path is flipped.  Same floating-point
// operations, same tie rules => the same assignment as scipy, bit for bit, also where the optimum is not unique
// (tests/test_host_logic.py checks the sequential form of the same rules, eigsort_decide.h; tests/test_gpu_parity.py
// this kernel against scipy).
//
// Parallel form: ONE thread-block cluster of 8 CTAs x 1024 threads owns the problem.  CTA c keeps the state of the
// columns [c W, (c+1) W) in its shared memory (v, tentative cost, predecessor row, assigned row, position in scipy's
// `remaining` order, visited flag: 32 bytes per column, up to ~56 000 columns); the row duals and the assignment live in
// global memory.  One search step = every thread updates its columns from the current row of the cost matrix (read
// once, coalesced, 1/8 per CTA) and proposes its best (value, tie key, column, assigned row); warp shuffles, one
// __syncthreads, one store per peer into the peers' shared memory (DSMEM) and ONE cluster barrier later every CTA holds
// the 8 proposals and takes the same decision.  The `remaining` array itself never exists: removing position p moves the
// last position to p, which every thread applies to its own columns.  A step costs one DRAM row read + one cluster
// barrier (~2 us) instead of an O(N) scan on one host core.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace fb {

constexpr int LS_CLUSTER = 8;
constexpr int LS_THREADS = 1024;

struct LsCand {
  double val;
  unsigned key;  // unassigned column: 0x40000000 + position, assigned: 0x3fffffff - position (larger wins on equal val)
  int col;
  int row;       // row the column is assigned to (-1: none)
  int pad;
};

__device__ __forceinline__ bool ls_better(double av, unsigned ak, double bv, unsigned bk) {
  return av < bv || (av == bv && ak > bk);
}

__device__ __forceinline__ double ld_cg_f64(const double* p) { return __ldcg(p); }

__global__ void __launch_bounds__(LS_THREADS, 1)
k_lsap(const double* __restrict__ cost, int nr, int nc, int W, double* __restrict__ u, int* __restrict__ col4row,
       int* __restrict__ status) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  extern __shared__ __align__(16) unsigned char ls_smem[];
  double* v = reinterpret_cast<double*>(ls_smem);
  double* spc = v + W;
  int* path = reinterpret_cast<int*>(spc + W);
  int* r4c = path + W;
  int* pos = r4c + W;
  int* sc = pos + W;
  __shared__ LsCand s_warp[LS_THREADS / 32];
  __shared__ LsCand s_slot[2][LS_CLUSTER];
  const int c0 = rank * W;
  const int ncol = max(0, min(W, nc - c0));
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  for (int l = tid; l < ncol; l += LS_THREADS) {
    v[l] = 0.0;
    r4c[l] = -1;
    path[l] = -1;
  }
  cluster.sync();
  unsigned it = 0;  // search steps so far (parity selects the proposal slots)
  for (int cur = 0; cur < nr; ++cur) {
    for (int l = tid; l < ncol; l += LS_THREADS) {
      spc[l] = inf;
      sc[l] = 0;
      pos[l] = nc - 1 - (c0 + l);
    }
    double min_val = 0.0;
    int i = cur, nrem = nc, sink = -1, prev_last = -1, prev_index = -1;
    bool fail = false;
    while (sink < 0) {
      const double ui = ld_cg_f64(u + i);
      const double* crow = cost + (size_t)i * nc + c0;
      double bv = inf;
      unsigned bk = 0u;
      int bc = -1, br = -1;
      for (int l = tid; l < ncol; l += LS_THREADS) {
        if (sc[l]) continue;
        int p = pos[l];
        if (p == prev_last) {  // scipy: remaining[index] = remaining[--num_remaining]
          p = prev_index;
          pos[l] = p;
        }
        const double r = ((min_val + __ldg(crow + l)) - ui) - v[l];
        double s = spc[l];
        if (r < s) {
          s = r;
          spc[l] = r;
          path[l] = i;
        }
        const int row = r4c[l];
        const unsigned key = row < 0 ? 0x40000000u + (unsigned)p : 0x3fffffffu - (unsigned)p;
        if (bc < 0 || ls_better(s, key, bv, bk)) {
          bv = s;
          bk = key;
          bc = c0 + l;
          br = row;
        }
      }
      // warp, CTA, cluster
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const unsigned ok = __shfl_xor_sync(0xffffffffu, bk, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        const int orow = __shfl_xor_sync(0xffffffffu, br, o);
        if (oc >= 0 && (bc < 0 || ls_better(ov, ok, bv, bk))) {
          bv = ov;
          bk = ok;
          bc = oc;
          br = orow;
        }
      }
      if (lane == 0) {
        s_warp[warp].val = bv;
        s_warp[warp].key = bk;
        s_warp[warp].col = bc;
        s_warp[warp].row = br;
      }
      __syncthreads();
      if (warp == 0) {
        LsCand c = s_warp[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, c.val, o);
          const unsigned ok = __shfl_xor_sync(0xffffffffu, c.key, o);
          const int oc = __shfl_xor_sync(0xffffffffu, c.col, o);
          const int orow = __shfl_xor_sync(0xffffffffu, c.row, o);
          if (oc >= 0 && (c.col < 0 || ls_better(ov, ok, c.val, c.key))) {
            c.val = ov;
            c.key = ok;
            c.col = oc;
            c.row = orow;
          }
        }
        if (lane < LS_CLUSTER) {  // lane p posts this CTA's proposal into CTA p's slot
          LsCand* remote = cluster.map_shared_rank(&s_slot[it & 1u][rank], lane);
          *remote = c;
        }
      }
      cluster.sync();
      LsCand best = s_slot[it & 1u][0];
#pragma unroll
      for (int p = 1; p < LS_CLUSTER; ++p) {
        const LsCand c = s_slot[it & 1u][p];
        if (c.col >= 0 && (best.col < 0 || ls_better(c.val, c.key, best.val, best.key))) best = c;
      }
      ++it;
      if (best.col < 0 || !(best.val < inf)) {  // no column left / infeasible (non-finite costs): same verdict everywhere
        fail = true;
        break;
      }
      min_val = best.val;
      const bool un = best.key >= 0x40000000u;
      const int index = un ? (int)(best.key - 0x40000000u) : (int)(0x3fffffffu - best.key);
      const int jl = best.col - c0;
      if (jl >= 0 && jl < ncol && (jl % LS_THREADS) == tid) sc[jl] = 1;  // the thread that owns the column
      prev_last = nrem - 1;
      prev_index = index;
      --nrem;
      if (best.row < 0)
        sink = best.col;
      else
        i = best.row;
    }
    if (fail) {
      if (rank == 0 && tid == 0) *status = -1;
      break;
    }
    // duals (scipy updates u over the visited rows and v over the visited columns; a visited row other than `cur` is the
    // row a visited column is assigned to, so both are column-owned updates)
    if (rank == 0 && tid == 0) __stcg(u + cur, ld_cg_f64(u + cur) + min_val);
    for (int l = tid; l < ncol; l += LS_THREADS)
      if (sc[l]) {
        const double d = min_val - spc[l];
        const int row = r4c[l];
        if (row >= 0) __stcg(u + row, ld_cg_f64(u + row) + d);
        v[l] = v[l] - d;
      }
    cluster.sync();  // the dual updates have read the old assignment
    if (rank == 0 && tid == 0) {  // flip the path from the sink back to `cur`
      int j = sink;
      for (int hops = 0;; ++hops) {
        if (hops > nr || j < 0) {  // cannot happen (the path leads back to `cur`); a bug must not become a hang
          *status = -2;
          break;
        }
        const int owner = j / W, l = j - owner * W;
        const int ii = *cluster.map_shared_rank(path + l, owner);
        *cluster.map_shared_rank(r4c + l, owner) = ii;
        const int jn = __ldcg(col4row + ii);
        __stcg(col4row + ii, j);
        j = jn;
        if (ii == cur) break;
      }
    }
    cluster.sync();
  }
  cluster.sync();  // nobody leaves while a peer may still address its shared memory
}

static size_t lsap_smem_bytes(int W) { return (size_t)W * (8 + 8 + 4 + 4 + 4 + 4); }

}  // namespace fb

using namespace fb;

extern "C" {

size_t focusr_lsap_workspace_bytes(int n_rows) { return align_up(sizeof(double) * (size_t)n_rows) + 256; }

// cost: device, n_rows x n_cols row-major, finite, n_rows <= n_cols.  col4row: device int32 [n_rows] (column assigned to
// every row, what scipy returns as col_ind).  status_host: 0, or -1 for an infeasible matrix (synchronises the stream).
int focusr_lsap(const double* cost, int n_rows, int n_cols, int* col4row, int* status_host, void* workspace,
                size_t workspace_bytes, focusr_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FB_REQUIRE(n_rows >= 1 && n_cols >= n_rows, "lsap: need 1 <= n_rows <= n_cols (transpose the problem otherwise)");
  const int W = (div_up(n_cols, LS_CLUSTER) + 1) & ~1;  // even: the int arrays behind the doubles stay 8-byte aligned
  const size_t smem = lsap_smem_bytes(W);
  FB_REQUIRE(smem <= 200 * 1024, "lsap: %d columns exceed the %d the cluster's shared memory holds", n_cols,
             (int)(200 * 1024 / 32) * LS_CLUSTER);
  FB_REQUIRE(workspace != nullptr && workspace_bytes >= focusr_lsap_workspace_bytes(n_rows), "lsap: workspace too small");
  double* u = static_cast<double*>(workspace);
  int* status_dev = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + align_up(sizeof(double) * (size_t)n_rows));
  FB_CUDA(cudaMemsetAsync(u, 0, sizeof(double) * (size_t)n_rows, stream));
  FB_CUDA(cudaMemsetAsync(status_dev, 0, sizeof(int), stream));
  FB_CUDA(cudaMemsetAsync(col4row, 0xff, sizeof(int) * (size_t)n_rows, stream));
  FB_CUDA(cudaFuncSetAttribute(k_lsap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(LS_CLUSTER);
  cfg.blockDim = dim3(LS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = LS_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FB_CUDA(cudaLaunchKernelEx(&cfg, k_lsap, cost, n_rows, n_cols, W, u, col4row, status_dev));
  FB_COUNT_LAUNCH(1);
  int st = 0;
  FB_CUDA(cudaMemcpyAsync(&st, status_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
  FB_CUDA(cudaStreamSynchronize(stream));
  if (status_host) *status_host = st;
  return FB_OK;
}
}
