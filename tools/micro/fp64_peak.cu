// Measured FP64 denominators of this pool's B200 (SURVEY.md section 7.3-8 asks for a DGEMM-class peak; cuBLAS DGEMM is
// timed from Python in tools/dense_evidence.py, this is the hand-written instruction-issue peak):
//   (1) mma.sync.aligned.m8n8k4.f64 (SASS DMMA) issued back to back from registers, 8 independent accumulators per warp,
//       every SM full of warps: 512 flop per warp instruction;
//   (2) fma.rn.f64 (DFMA) likewise: 64 flop per warp instruction.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak     (prints one JSON line)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
  double c[8];
  for (int i = 0; i < 8; ++i) c[i] = i;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
  for (int i = 0; i < 8; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 8, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * grid * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best_mma = 1e30f, best_fma = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    float ms;
    cudaEventRecord(e0);
    k_dmma<<<grid, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_mma) best_mma = ms;
    cudaEventRecord(e0);
    k_dfma<<<grid, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_fma) best_fma = ms;
  }
  if (cudaGetLastError() != cudaSuccess) return 1;
  const double warps = (double)grid * 8;
  const double mma_flops = warps * iters * 8.0 * 512.0, fma_flops = warps * iters * 8.0 * 64.0;
  printf("{\"sms\": %d, \"dmma_m8n8k4_tflops\": %.2f, \"dfma_tflops\": %.2f, \"dfma_ginstr_per_s\": %.1f, \"how\": \"back-to-back from "
         "registers, 8 independent accumulators per warp, %d warps per SM, best of 4 after warm-up, CUDA events\"}\n",
         sms, mma_flops / best_mma / 1e9, fma_flops / best_fma / 1e9, fma_flops / 2.0 / 32.0 / best_fma / 1e6 * 32.0, 64);
  return 0;
}
