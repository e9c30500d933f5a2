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
};

bool spmm_block_supported(int b);

// mode 0: out = alpha[mesh][step] * (L y - center[mesh] y) - gamma[mesh][step] * x_prev
// mode 1: out = (D - A) y          mode 2: out = L y
int launch_spmm(int mode, int b, const SpmmGraph& g, const double* y, const double* x_prev, double* out,
                const double* alpha, const double* gamma, const double* center, int step, int n_steps,
                cudaStream_t stream);

// row-partitioned multi-GPU: remote columns are gathered from the owning rank's memory (NVLink P2P)
int launch_spmm_p2p(int mode, int b, const SpmmGraph& g, int n_loc, const double* y, const double* const* peer_y,
                    const int* ghost_peer, const int* ghost_row, const double* x_prev, double* out, const double* alpha,
                    const double* gamma, const double* center, int step, int n_steps, cudaStream_t stream);

}  // namespace fb
