// Micro-benchmark: cost per step of the synchronisation schemes considered for the persistent filter kernel of the
// row-partitioned solve (one GPU; the cross-GPU part adds the NVLink flag latency).  Each CTA owns 64 rows of a [rows][24]
// fp32 block; a step reads the rows of the two neighbouring CTAs and writes its own rows of the other buffer.
//   mode 0: no synchronisation (wrong results, lower bound)   mode 1: grid barrier (counter + gate, as step_barrier)
//   mode 2: neighbour flags (wait for CTA c-1, c+1 to have finished the previous step)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu && ./barrier_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__global__ void __launch_bounds__(256, 8)
k(int mode, int steps, float* a, float* b, unsigned* counter, unsigned long long* gate, unsigned long long* flags) {
  const int c = blockIdx.x, G = gridDim.x;
  const int row = c * 64 + threadIdx.x / 4, sl = threadIdx.x % 4;
  for (int s = 0; s < steps; ++s) {
    float* src = (s & 1) ? b : a;
    float* dst = (s & 1) ? a : b;
    if (mode == 1) {
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned old;
        fence_gpu();
        asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
        const unsigned long long epoch = s + 1;
        if (old + 1u == (unsigned)(s + 1) * G) {
          fence_gpu();
          asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(gate), "l"(epoch) : "memory");
        } else {
          unsigned long long v;
          do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(gate) : "memory"); } while (v < epoch);
        }
        fence_gpu();
      }
      __syncthreads();
    } else if (mode == 2) {
      __syncthreads();
      if (threadIdx.x < 2) {
        const int nb = threadIdx.x == 0 ? (c + G - 1) % G : (c + 1) % G;
        unsigned long long v;
        do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + 16 * nb) : "memory"); } while (v < (unsigned long long)s);
      }
      if (threadIdx.x == 0) fence_gpu();
      __syncthreads();
    }
    const int l = ((c + G - 1) % G) * 64 + threadIdx.x / 4, r = ((c + 1) % G) * 64 + threadIdx.x / 4;
    float4 x, y, z;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "l"(reinterpret_cast<float4*>(src + (size_t)l * 16) + sl) : "memory");
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w) : "l"(reinterpret_cast<float4*>(src + (size_t)r * 16) + sl) : "memory");
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(z.x), "=f"(z.y), "=f"(z.z), "=f"(z.w) : "l"(reinterpret_cast<float4*>(src + (size_t)row * 16) + sl) : "memory");
    float4 o = make_float4(0.25f * (x.x + y.x) + 0.5f * z.x, 0.25f * (x.y + y.y) + 0.5f * z.y, 0.25f * (x.z + y.z) + 0.5f * z.z, 0.25f * (x.w + y.w) + 0.5f * z.w);
    reinterpret_cast<float4*>(dst + (size_t)row * 16)[sl] = o;
    if (mode == 2) {
      __syncthreads();
      if (threadIdx.x == 0) {
        fence_gpu();
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flags + 16 * c), "l"((unsigned long long)(s + 1)) : "memory");
      }
    }
  }
}

int main() {
  for (int G : {282, 976, 1184}) {
    float *a, *b;
    unsigned* counter;
    unsigned long long *gate, *flags;
    cudaMalloc(&a, (size_t)G * 64 * 16 * 4);
    cudaMalloc(&b, (size_t)G * 64 * 16 * 4);
    cudaMalloc(&counter, 256);
    cudaMalloc(&gate, 256);
    cudaMalloc(&flags, (size_t)G * 128);
    cudaMemset(a, 0, (size_t)G * 64 * 16 * 4);
    cudaMemset(b, 0, (size_t)G * 64 * 16 * 4);
    for (int mode = 0; mode < 3; ++mode) {
      const int steps = 2000;
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(counter, 0, 256);
        cudaMemset(gate, 0, 256);
        cudaMemset(flags, 0, (size_t)G * 128);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        int m = mode, st = steps;
        void* args[] = {&m, &st, &a, &b, &counter, &gate, &flags};
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchCooperativeKernel((const void*)k, dim3(G), dim3(256), args, 0, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(err)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      printf("G=%4d mode %d: %.3f us per step\n", G, mode, best * 1e3f / steps);
    }
  }
  return 0;
}
