// Advantage normalisation with the data RESIDENT in shared memory between the statistics and the normalise pass (sm_100a, GPU only).
//
// (x - mean) / (std + eps) needs the grid-wide mean / std before the first output element can be written; the two-launch AdvNormAlgo
// therefore reads x twice (12 B moved per element for 8 B algorithmic).  Up to 148 x 48 K elements (28 MB of fp32) fit in the shared
// memory of the chip, so here a CTA per SM pulls its contiguous slice in with the copy engine (cp.async.bulk, no registers, the whole
// slice in flight at once), sums it in float64 out of shared memory, meets the other CTAs at ONE grid barrier (cooperative launch),
// normalises the slice in place and pushes it out with one bulk store: x is read once, out is written once, one launch instead of two.
// Statistics: per-thread float64 partial sums over the slice in a fixed order, a fixed-order block tree, and every CTA folds the
// per-CTA partials in CTA order — the result does not depend on scheduling.  Needs n % 4 == 0 and 16-byte aligned tensors.
// Measured (B200, tools/membench.py): 7 - 8 us against 11 us of the two-launch kernel up to the MAPPO sizes (393 K elements), but 27 us
// against 24 us at the largest resident size (7.3 M): load, statistics, barrier, normalise and store are serial per SM, so the chip's
// 29 MB of shared memory is drained and refilled with DRAM idle in between.  The kernel is therefore taken up to 2 M elements; larger
// or unaligned inputs take the two-launch kernel.
#pragma once
#include "umma.cuh"

#define ANR_CAP_FLOATS 49152            // 192 KB slice per CTA
#define ANR_CHUNK_BYTES 32768u
#define ANR_MAX_N (1 << 21)

__device__ __forceinline__ double anr_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
// fixed-order block sum of two doubles per thread; the result is valid in every thread
__device__ __forceinline__ void anr_block_sum(double& s, double& q, double* red, int t) {
  s = anr_warp_sum(s);
  q = anr_warp_sum(q);
  __syncthreads();                                  // red may still be read from the previous call
  if ((t & 31) == 0) { red[t >> 5] = s; red[8 + (t >> 5)] = q; }
  __syncthreads();
  s = 0.0; q = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { s += red[w]; q += red[8 + w]; }
}

__global__ void __launch_bounds__(256, 1) frl_adv_norm_resident_kernel(const __grid_constant__ AdvNormArgs a) {
  extern __shared__ __align__(128) float anr_sm[];
  float* slice = anr_sm;
  double* red = reinterpret_cast<double*>(anr_sm + ANR_CAP_FLOATS);
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 16);
  const int t = (int)threadIdx.x, cta = (int)blockIdx.x, ncta = (int)gridDim.x;
  const int n4 = a.n >> 2, per = (n4 + ncta - 1) / ncta;                 // quads per CTA
  const int q0 = cta * per, q1 = (q0 + per < n4) ? q0 + per : n4, nq = q1 > q0 ? q1 - q0 : 0;
  const uint32_t bytes = (uint32_t)nq * 16u;
  if (t == 0) {
    um_mbar_init(bar, 1);
    um_fence_mbar_init();
    if (bytes) {
      um_mbar_expect_tx(bar, bytes);
      for (uint32_t o = 0; o < bytes; o += ANR_CHUNK_BYTES)
        um_bulk_g2s(reinterpret_cast<char*>(slice) + o, reinterpret_cast<const char*>(a.x + 4 * (size_t)q0) + o,
                    (bytes - o < ANR_CHUNK_BYTES) ? bytes - o : ANR_CHUNK_BYTES, bar);
    }
  }
  __syncthreads();
  if (bytes) um_mbar_wait(bar, 0);
  double s = 0.0, q = 0.0;
  for (int i = t; i < nq; i += 256) {
    const float4 v = *reinterpret_cast<const float4*>(slice + 4 * i);
    s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    q += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
  }
  anr_block_sum(s, q, red, t);
  if (t == 0) { a.part[2 * cta] = s; a.part[2 * cta + 1] = q; }
  __threadfence();
  cooperative_groups::this_grid().sync();
  s = 0.0; q = 0.0;
  for (int i = t; i < ncta; i += 256) { s += __ldcg(a.part + 2 * i); q += __ldcg(a.part + 2 * i + 1); }
  anr_block_sum(s, q, red, t);
  const double mean = s / (double)a.n;
  double var = (q - s * mean) / (double)(a.n - 1);                       // torch.std(): unbiased
  if (var < 0.0) var = 0.0;
  const float mf = (float)mean, den = (float)sqrt(var) + a.eps;
  for (int i = t; i < nq; i += 256) {
    float4 v = *reinterpret_cast<float4*>(slice + 4 * i);
    v.x = fdiv(v.x - mf, den); v.y = fdiv(v.y - mf, den); v.z = fdiv(v.z - mf, den); v.w = fdiv(v.w - mf, den);
    *reinterpret_cast<float4*>(slice + 4 * i) = v;
  }
  um_fence_proxy_async();
  __syncthreads();
  if (t == 0 && bytes) {
    for (uint32_t o = 0; o < bytes; o += ANR_CHUNK_BYTES)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<char*>(a.out + 4 * (size_t)q0) + o),
                   "r"(um_smem_u32(reinterpret_cast<char*>(slice) + o)), "r"((bytes - o < ANR_CHUNK_BYTES) ? bytes - o : ANR_CHUNK_BYTES)
                   : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// 0 = launched, 1 = not applicable (the caller takes the two-launch kernel), < 0 = error
static int adv_norm_resident_launch(const AdvNormArgs& a0, cudaStream_t s) {
  static const bool off = getenv("FREERL_B200_ADVNORM_TWO_PASS") != nullptr;   // A/B switch
  const int sms = frl_device_max_ctas();
  if (off || a0.n % 4 || ((((size_t)a0.x) | ((size_t)a0.out)) & 15) || (long)a0.n > (long)sms * ANR_CAP_FLOATS || a0.n > ANR_MAX_N) return 1;
  const int n4 = a0.n >> 2;
  int ncta = (n4 + 255) / 256;                       // at least one quad per thread before the grid grows
  if (ncta > sms) ncta = sms;
  const int smem = ANR_CAP_FLOATS * 4 + 16 * 8 + 64;
  FRL_SMEM_OPT_IN(frl_adv_norm_resident_kernel, smem, 48 * 1024);
  AdvNormArgs a = a0;
  a.ncta = ncta;
  void* kargs[] = {(void*)&a};
  FRL_CUDA_OK(cudaLaunchCooperativeKernel((void*)frl_adv_norm_resident_kernel, dim3(ncta), dim3(256), kargs, (size_t)smem, s));
  ++frl_launch_counter;
  return 0;
}
