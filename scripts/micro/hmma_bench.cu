// Microbenchmark (developer tool): latency and issue interval of the legacy mma.sync.m16n8k16 (HMMA.16816.F32) on sm_100a.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ab/hmma_bench scripts/micro/hmma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void bench(long long* out, float* sink, int iters) {
  float c[CHAINS][4];
  uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
  uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = 0x3c003c00u;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[k][j] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) mma16816(c[k], a, b0, b1);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CHAINS>
void run(int warps, long long* d_out, float* d_sink) {
  const int iters = 2000;
  bench<CHAINS><<<1, 32 * warps>>>(d_out, d_sink, iters);
  bench<CHAINS><<<1, 32 * warps>>>(d_out, d_sink, iters);
  long long h;
  cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double per = (double)h / iters / CHAINS;
  printf("warps/SM %2d (per sub-partition %4.1f)  independent chains %d : %7.2f cycles per HMMA per warp -> %6.2f cycles per HMMA per sub-partition\n",
         warps, warps / 4.0, CHAINS, per, per / (warps / 4.0 < 1 ? 1 : warps / 4.0));
}

int main() {
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 4 * 1024 * 64);
  for (int warps : {1, 4, 8, 16}) {
    run<1>(warps, d_out, d_sink);
    run<2>(warps, d_out, d_sink);
    run<4>(warps, d_out, d_sink);
    run<8>(warps, d_out, d_sink);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
