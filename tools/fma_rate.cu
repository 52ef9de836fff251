// Micro-benchmark: achievable fp32 FMA rate per SM for the register-tile patterns of the GEMM microkernels (4x4 outer
// product per k, operands in registers), scalar FFMA vs packed FFMA2, at 8 / 16 / 32 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate tools/fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma2_bcast(float& d0, float& d1, float a, float b0, float b1) {
  unsigned long long B, C;
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(d0), "f"(d1));
  asm("{\n\t.reg .b64 aa;\n\tmov.b64 aa, {%1, %1};\n\tfma.rn.f32x2 %0, aa, %2, %0;\n\t}" : "+l"(C) : "f"(a), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(C));
}

template <int MODE>
__global__ void k(float* out, int iters, long long* clk) {
  float acc[4][4];
  float a[4], b[4];
  for (int i = 0; i < 4; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = threadIdx.x * 2e-3f - i; for (int j = 0; j < 4; ++j) acc[i][j] = 0.f; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        } else {
          fma2_bcast(acc[i][0], acc[i][1], a[i], b[0], b[1]);
          fma2_bcast(acc[i][2], acc[i][3], a[i], b[2], b[3]);
        }
      }
      // rotate operands so the compiler cannot hoist (cheap: 2 FADDs on 8 FMAs... keep it off the fma pipe count)
      a[kk & 3] += 1e-7f; b[(kk + 1) & 3] -= 1e-7f;
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int threads = 256; threads <= 1024; threads *= 2) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, iters, clk); else k<1><<<148, threads>>>(out, iters, clk);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      const double fma = (double)iters * 8 * 16 * threads;
      printf("%s warps/SM=%2d: %.1f FMA/clk/SM (%lld clk)\n", mode ? "FFMA2" : "FFMA ", threads / 32, fma / (double)c, c);
    }
  return 0;
}
