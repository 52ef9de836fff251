// Standalone check of the tcgen05 building blocks the large-minibatch path uses (csrc/umma.cuh), before they go into a learn kernel:
//   * "layout P" = the no-swizzle canonical UMMA shared-memory layout of a [R x C] fp32 matrix,
//       float offset(r, c) = (r / 8) * 8C + (c / 4) * 32 + (r % 8) * 4 + (c % 4)
//     read by tcgen05.mma BOTH as a K-major operand (MN = r, K = c) and as an MN-major operand (MN = c, K = r);
//   * 3xTF32 error compensation (hi*hi + hi*lo + lo*hi) with fp32 accumulators in TMEM;
//   * the three GEMM shapes of an MLP layer: forward (A, B K-major), backward-dX (A K-major, B MN-major), dW (A, B MN-major).
// Prints max relative error against a float64 host product (1xTF32 and 3xTF32) and the issue->commit time of a long MMA chain.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o variants/umma_test tools/umma_test.cu && variants/umma_test
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

#include "../freerl_b200/csrc/umma.cuh"

// mode 0: D[m][n] = sum_k A[m][k] B[n][k]     SS: A, B in shared memory, layout Q (K-major)               (forward, operands staged)
// mode 1: same product                         TS: A written to TMEM with tcgen05.st (hi | lo), B layout Q (forward / dX of the learn kernel)
// mode 2: D[m][n] = sum_k A[k][m] B[k][n]     SS: A stored [K x M], B stored [K x N], layout S (MN-major)  (dW = dZ^T . H over tile rows)
// M = 128, N = 128, K = 64.  terms: 1 = TF32, 3 = 3xTF32.
__global__ void __launch_bounds__(128, 1) umma_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int mode,
                                                      int terms, int reps, long long* clk, int n_issue = 128) {
  constexpr int M = 128, N = 128, K = 64;
  extern __shared__ __align__(1024) unsigned char smraw[];
  float* a_hi = (float*)smraw;
  float* a_lo = a_hi + M * K;
  float* b_hi = a_lo + M * K;
  float* b_lo = b_hi + N * K;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (mode != 1)
    for (int i = t; i < M * K; i += 128) {
      float x = A[i];
      int o = mode == 2 ? um_s_off(i / M, i % M, M) : um_q_off(i / K, i % K, M);
      a_hi[o] = um_hi(x);
      a_lo[o] = x - um_hi(x);
    }
  for (int i = t; i < N * K; i += 128) {
    float x = B[i];
    int o = mode == 2 ? um_s_off(i / N, i % N, N) : um_q_off(i / K, i % K, N);
    b_hi[o] = um_hi(x);
    b_lo[o] = x - um_hi(x);
  }
  if (t == 0) { um_mbar_init(&bar, 1); um_fence_mbar_init(); }
  um_fence_proxy_async();
  if (warp == 0) um_tmem_alloc<256>(&tmem_slot);
  um_fence_before();
  __syncthreads();
  um_fence_after();
  const uint32_t tmem = tmem_slot, lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t t_ah = tmem + 128, t_al = tmem + 192;          // A hi | lo, 64 columns each
  if (mode == 1) {
    const int row = warp * 32 + (t & 31);
    for (int c0 = 0; c0 < K; c0 += 16) {
      float hi[16], lo[16];
      for (int i = 0; i < 16; ++i) { float x = A[row * K + c0 + i]; hi[i] = um_hi(x); lo[i] = x - hi[i]; }
      um_st16(t_ah + lane_base + c0, hi);
      um_st16(t_al + lane_base + c0, lo);
    }
    um_wait_st();
    um_fence_before();
    __syncthreads();
    um_fence_after();
  }
  if (t == 0) {
    const int mn = mode == 2;
    const uint32_t idesc = um_idesc_tf32(M, n_issue, mn, mn);      // n_issue < N: timing of narrow MMAs (results unused)
    const uint32_t a_step = mn ? 32 * M : 32 * M, b_step = mn ? 32 * N : 32 * N;   // bytes per K = 8 step (both layouts: 32 x rows-or-cols)
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t ah = mn ? um_desc_s(um_smem_u32(a_hi) + ks * a_step, M) : um_desc_q(um_smem_u32(a_hi) + ks * a_step, M);
        const uint64_t al = mn ? um_desc_s(um_smem_u32(a_lo) + ks * a_step, M) : um_desc_q(um_smem_u32(a_lo) + ks * a_step, M);
        const uint64_t bh = mn ? um_desc_s(um_smem_u32(b_hi) + ks * b_step, N) : um_desc_q(um_smem_u32(b_hi) + ks * b_step, N);
        const uint64_t bl = mn ? um_desc_s(um_smem_u32(b_lo) + ks * b_step, N) : um_desc_q(um_smem_u32(b_lo) + ks * b_step, N);
        const uint32_t first = (rep | ks) == 0 ? 0u : 1u;
        if (mode == 1) {
          if (terms == 3) {
            um_mma_ts(tmem, t_al + ks * 8, bh, idesc, first);
            um_mma_ts(tmem, t_ah + ks * 8, bl, idesc, 1u);
            um_mma_ts(tmem, t_ah + ks * 8, bh, idesc, 1u);
          } else {
            um_mma_ts(tmem, t_ah + ks * 8, bh, idesc, first);
          }
        } else if (terms == 3) {
          um_mma_ss(tmem, al, bh, idesc, first);
          um_mma_ss(tmem, ah, bl, idesc, 1u);
          um_mma_ss(tmem, ah, bh, idesc, 1u);
        } else {
          um_mma_ss(tmem, ah, bh, idesc, first);
        }
      }
    um_commit(&bar);
    um_mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (clk) *clk = t1 - t0;
  }
  __syncthreads();
  um_mbar_wait(&bar, 0);
  um_fence_after();
  const int row = warp * 32 + (t & 31);
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    um_ld16(tmem + lane_base + c0, v);
    for (int i = 0; i < 16; ++i) D[row * N + c0 + i] = v[i];
  }
  um_fence_before();
  __syncthreads();
  if (warp == 0) um_tmem_dealloc<256>(tmem);
}

// ---- address probe: one 128x128x8 MMA, A = K-major selector (A[m][k] = [k == m % 8]), B region filled with a position code ----
__global__ void __launch_bounds__(128, 1) umma_probe(float* __restrict__ D, uint32_t lbo, uint32_t sbo, uint32_t layout_type, int b_mn, int code_mode) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  float* a = (float*)smraw;                 // [128 x 8] layout P (C = 8): 4 KB
  float* b = a + 1024;                      // 64 KB region, position coded
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 128 * 8; i += 128) { int r = i / 8, c = i % 8; a[um_q_off(r, c, 128)] = (c == (r & 7)) ? 1.f : 0.f; }
  for (int i = t; i < 16384; i += 128) b[i] = code_mode == 0 ? (float)((i >> 2) & 2047) : code_mode == 1 ? (float)(i & 3) : (float)(i >> 13);
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(um_smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(um_smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    const uint32_t idesc = um_idesc_tf32(128, 128, 0, b_mn);
    uint64_t ad = um_desc_q(um_smem_u32(a), 128);
    uint64_t bd = um_desc(um_smem_u32(b), lbo, sbo, layout_type);
    um_mma_ss(tmem, ad, bd, idesc, 0u);
    um_commit(&bar);
  }
  __syncthreads();
  um_mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + (t & 31);
  for (int c0 = 0; c0 < 128; c0 += 8) {
    float v[16];
    if (c0 & 8) continue;
    um_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) D[row * 128 + c0 + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

static void run_probe(float* dD, uint32_t lbo, uint32_t sbo, uint32_t lt, int b_mn) {
  std::vector<float> c0(128 * 128), c1(128 * 128), c2(128 * 128);
  const int smem = 4096 + 65536;
  CK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float>* outs[3] = {&c0, &c1, &c2};
  for (int cm = 0; cm < 3; ++cm) {
    CK(cudaMemset(dD, 0xff, 128 * 128 * 4));
    umma_probe<<<1, 128, smem>>>(dD, lbo, sbo, lt, b_mn, cm);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe lbo %u sbo %u lt %u mn %d: CUDA error %s\n", lbo, sbo, lt, b_mn, cudaGetErrorString(e)); exit(1); }
    CK(cudaMemcpy(outs[cm]->data(), dD, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  }
  printf("probe lbo %u sbo %u layout_type %u b_mn_major %d: float offset read for (k, n):\n", lbo, sbo, lt, b_mn);
  const int ns[] = {0, 1, 2, 3, 4, 5, 8, 12, 16, 32, 64, 127};
  for (int k = 0; k < 8; ++k) {
    printf("  k=%d:", k);
    for (int n : ns) printf(" n%d->%d", n, ((int)c2[k * 128 + n] * 2048 + (int)c0[k * 128 + n]) * 4 + (int)c1[k * 128 + n]);
    printf("\n");
  }
}

int main(int argc, char** argv) {
  constexpr int M = 128, N = 128, K = 64;
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  long long* dclk;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMalloc(&dclk, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const int smem = (2 * M * K + 2 * N * K) * 4;
  CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (argc > 1) {      // probe mode: umma_test probe
    run_probe(dD, 2048, 128, 0, 0);      // layout Q of a [128 x 8] B operand (K-major)
    run_probe(dD, 512, 2048, 1, 1);      // layout S of a [8 x 128] B operand (MN-major, 128B swizzle / 32B atom)
    return 0;
  }
  int bad = 0;
  for (int mode = 0; mode < 3; ++mode)
    for (int terms = 1; terms <= 3; terms += 2) {
      CK(cudaMemset(dD, 0, D.size() * 4));
      umma_kernel<<<1, 128, smem>>>(dA, dB, dD, mode, terms, 1, dclk);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0, maxref = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double s = 0;
          for (int k = 0; k < K; ++k) {
            double a = mode == 2 ? A[k * M + m] : A[m * K + k];
            double b = mode == 2 ? B[k * N + n] : B[n * K + k];
            s += a * b;
          }
          maxerr = fmax(maxerr, fabs(s - D[m * N + n]));
          maxref = fmax(maxref, fabs(s));
        }
      printf("mode %d terms %d: max abs err %.3e (max |ref| %.2f) -> rel %.3e\n", mode, terms, maxerr, maxref, maxerr / maxref);
      if (maxerr / maxref > (terms == 3 ? 2e-6 : 5e-3)) bad++;
    }
  // timing: chain of reps x (K/8) x terms MMAs of 128x128x8
  for (int terms = 1; terms <= 3; terms += 2)
    for (int mode = 0; mode < 3; ++mode) {
      const int reps = 64;
      umma_kernel<<<1, 128, smem>>>(dA, dB, dD, mode, terms, reps, dclk);
      CK(cudaDeviceSynchronize());
      long long c;
      CK(cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost));
      int n = reps * (K / 8) * terms;
      printf("timing mode %d terms %d: %d MMAs (128x128x8 tf32) in %lld clk = %.1f clk / MMA = %.2f TFLOP/s/SM-equivalent at 1.965 GHz\n", mode, terms, n, c,
             (double)c / n, 2.0 * 128 * 128 * 8 / ((double)c / n) * 1.965e9 / 1e12);
    }
  for (int nn = 16; nn <= 64; nn *= 2)
    for (int mode = 0; mode < 3; ++mode) {
      const int reps = 64;
      umma_kernel<<<1, 128, smem>>>(dA, dB, dD, mode, 3, reps, dclk, nn);
      CK(cudaDeviceSynchronize());
      long long c;
      CK(cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost));
      int n = reps * (K / 8) * 3;
      printf("timing N = %d mode %d: %d MMAs (128x%dx8 tf32) in %lld clk = %.1f clk / MMA\n", nn, mode, n, nn, c, (double)c / n);
    }
  printf(bad ? "FAILED (%d)\n" : "ALL OK\n", bad);
  return bad != 0;
}
