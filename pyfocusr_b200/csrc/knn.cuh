// internal interface of knn.cu
#pragma once
#include <cuda_runtime.h>

namespace fb {
int launch_knn(const double* refs, int ld_refs, const int* ref_off, const double* queries, int ld_queries,
               const int* query_off, int n_segments, int max_queries, int dim, int k, long long* idx,
               double* dist, cudaStream_t stream);
}
