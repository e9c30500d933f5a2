// internal interface of spmm.cu
#pragma once
#include <cuda_runtime.h>

namespace fb {

struct SpmmGraph {
  const int* row_ptr;
  const int* cols;
  const double* weights;
  const double* degree;
  const double* degree_inv;
  const int* mesh_off;  // device [n_meshes + 1]
  int n_meshes;
  int max_mesh_rows;
  // fp32 copy of the matrix for the fp32 filter steps (k_matrix_f32): weights, and (degree, 1/degree~) per row
  const float* weights_f = nullptr;
  const float2* ddi_f = nullptr;
};

bool spmm_block_supported(int b);

// focusr_set_tuning(1, MB): L2 budget for blocking the filter over groups of meshes (0 = off)
extern int g_l2_budget_mb;
extern int g_smooth_variant;  // smooth_cluster.cu
extern int g_mixed_precision;  // focusr_set_tuning(3, v): fp32 early filter passes (default on)

// mode 0: out = alpha[mesh][step] * (L y - center[mesh] y) - gamma[mesh][step] * x_prev
// mode 1: out = (D - A) y          mode 2: out = L y
int launch_spmm(int mode, int b, const SpmmGraph& g, const double* y, const double* x_prev, double* out,
                const double* alpha, const double* gamma, const double* center, int step, int n_steps,
                cudaStream_t stream);

// fp32 copy of the matrix: wf[p] = (float)weights[p], ddi[i] = ((float)degree[i], (float)degree_inv[i])
int launch_matrix_f32(const double* weights, const double* degree, const double* degree_inv, long long nnz, int n_rows,
                      float* wf, float2* ddi, cudaStream_t stream);

// filter step on fp32 blocks (io 0), entering from (io 1) / returning to (io 2) the fp64 block; see spmm.cu
int launch_spmm_f32(int io, int b, const SpmmGraph& g, const void* y, const float* x_prev, void* out, float* y_copy,
                    const double* alpha, const double* gamma, const double* center, int step, int n_steps,
                    cudaStream_t stream);

// fp32 correction step: z_next = alpha_j ((L - c) z + r_j) - gamma_j z_prev with per-column tables [mesh][step][b];
// last: x (fp64) += z_next instead of storing it
int launch_spmm_corr(bool last, int b, const SpmmGraph& g, const float* z, const float* z_prev, const float* r, float* z_next,
                     double* x, const float* alpha_c, const float* gamma_c, const double* center, int step, int n_steps,
                     bool has_prev, cudaStream_t stream);

// row-partitioned multi-GPU: remote columns are gathered from the owning rank's memory (NVLink P2P)
int launch_spmm_p2p(int mode, int b, const SpmmGraph& g, int n_loc, const double* y, const double* const* peer_y,
                    const int* ghost_peer, const int* ghost_row, const double* x_prev, double* out, const double* alpha,
                    const double* gamma, const double* center, int step, int n_steps, cudaStream_t stream);

}  // namespace fb
