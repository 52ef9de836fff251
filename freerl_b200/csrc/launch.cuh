// Persistent cooperative launcher shared by all fused learn() kernels.
//
// An "Algo" provides:   typedef Args;  static const int NSTAGES;
//                       static int  wbuf_floats(const Args&);     // size of one weight staging buffer
//                       static int  user_floats(const Args&);     // smem floats after the engine's own
//                       static int  grid(const Args&, int max);   // CTAs wanted
//                       static int  n_updates(const Args&);
//                       static FRL_DEV void stage(int s, int u, Cta&, float* user, const Args&);
//                       static bool writes_params(int s);         // stage s rewrites weights a later TMA copy reads
//                       static bool stage_enabled(int s, int u, const Args&);   // false: skip stage s of update u and its barrier
// The kernel runs  for u: for s: stage(s,u); grid.sync()  — one launch performs n_updates sequential
// learn() steps (stages are separated by grid-wide barriers because every optimiser step needs the
// global gradient norm and the next phase needs the updated weights).
#pragma once
#include "engine.cuh"

#ifndef FRL_EMUL
namespace cg = cooperative_groups;

template <class A>
__global__ void __launch_bounds__(FRL_NT, 1) frl_persistent_kernel(const __grid_constant__ typename A::Args a) {
  extern __shared__ __align__(1024) float frl_smem[];
  Cta c;
  float* user = cta_init(c, (int)blockIdx.x, (int)gridDim.x, frl_smem, A::wbuf_floats(a));
  cg::grid_group grid = cg::this_grid();
  const int U = A::n_updates(a);
  for (int u = 0; u < U; ++u) {
    for (int s = 0; s < A::NSTAGES; ++s) {
      if (!A::stage_enabled(s, u, a)) continue;      // block-uniform AND grid-uniform: stage and its barrier are skipped together
      trace(1000 + s);
      A::stage(s, u, c, user, a);
      stage_reset(c);
      trace(1100 + s);
      stamp(c, 100 + s);
      if (A::writes_params(s)) fence_proxy_async();   // generic-proxy weight writes -> later TMA (async-proxy) reads
      grid.sync();
      stamp(c, 200 + s);
    }
  }
  res_drain(c);      // no bulk copy may be in flight when the CTA retires
}

int frl_device_max_ctas();   // SM count of the current device (1 CTA / SM for the persistent kernels)
extern long long frl_launch_counter;       // kernels launched by this library so far (frl_launch_count(), read by bench.py)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember the opt-in per device ordinal
#define FRL_SMEM_OPT_IN(kernel, smem_bytes, floor_bytes)                                                        \
  do {                                                                                                          \
    static int configured_bytes[64] = {0};                                                                      \
    int dev_ = 0;                                                                                               \
    FRL_CUDA_OK(cudaGetDevice(&dev_));                                                                          \
    if (dev_ < 0 || dev_ >= 64) { frl_set_error("device ordinal %d out of range", dev_); return -3; }           \
    if ((smem_bytes) > (floor_bytes) && (smem_bytes) > configured_bytes[dev_]) {                                \
      FRL_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (smem_bytes)));     \
      configured_bytes[dev_] = (smem_bytes);                                                                    \
    }                                                                                                           \
  } while (0)

template <class A>
int frl_launch(const typename A::Args& a, cudaStream_t stream) {
  const int smem_bytes = (cta_base_floats(A::wbuf_floats(a)) + A::user_floats(a)) * 4 + 64;
  if (smem_bytes > 227 * 1024) {
    frl_set_error("kernel needs %d B of shared memory (> 227 KB)", smem_bytes);
    return -3;
  }
  FRL_SMEM_OPT_IN(frl_persistent_kernel<A>, smem_bytes, 0);
  const int grid = A::grid(a, frl_device_max_ctas());
  int per_sm = 0;      // a cooperative grid must be co-resident
  FRL_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frl_persistent_kernel<A>, FRL_NT, (size_t)smem_bytes));
  if (per_sm * frl_device_max_ctas() < grid) {
    frl_set_error("cooperative launch of %d CTAs does not fit the device (%d per SM)", grid, per_sm);
    return -3;
  }
  typename A::Args args = a;
  void* kargs[] = {(void*)&args};
  FRL_CUDA_OK(cudaLaunchCooperativeKernel((void*)frl_persistent_kernel<A>, dim3(grid), dim3(FRL_NT), kargs, (size_t)smem_bytes, stream));
  ++frl_launch_counter;
  return 0;
}

// plain (non-cooperative) launch of a single-stage Algo over an arbitrary grid
template <class A>
__global__ void __launch_bounds__(FRL_NT, 1) frl_tile_kernel(const __grid_constant__ typename A::Args a) {
  extern __shared__ __align__(1024) float frl_smem[];
  Cta c;
  float* user = cta_init(c, (int)blockIdx.x, (int)gridDim.x, frl_smem, A::wbuf_floats(a));
  A::stage(0, 0, c, user, a);
}

template <class A>
int frl_launch_tiles(const typename A::Args& a, cudaStream_t stream) {
  const int smem_bytes = (cta_base_floats(A::wbuf_floats(a)) + A::user_floats(a)) * 4 + 64;
  FRL_SMEM_OPT_IN(frl_tile_kernel<A>, smem_bytes, 0);
  const int grid = A::grid(a, 1 << 30);
  frl_tile_kernel<A><<<grid, FRL_NT, smem_bytes, stream>>>(a);
  FRL_CUDA_OK(cudaGetLastError());
  ++frl_launch_counter;
  return 0;
}

#else  // ---------------------------------------------------------------- host emulation (tests only)
#include <stdlib.h>
#include <vector>

inline int frl_device_max_ctas() { return 148; }

template <class A>
int frl_launch(const typename A::Args& a, cudaStream_t) {
  const int floats = cta_base_floats(A::wbuf_floats(a)) + A::user_floats(a) + 16;
  const int grid = A::grid(a, frl_device_max_ctas());
  std::vector<float*> mem(grid);
  std::vector<Cta> ctas(grid);
  std::vector<float*> user(grid);
  for (int g = 0; g < grid; ++g) {
    mem[g] = (float*)aligned_alloc(1024, ((size_t)floats * 4 + 1023) / 1024 * 1024);
    for (int i = 0; i < floats; ++i) mem[g][i] = NAN;     // poison: reads of unwritten smem surface as NaN
    user[g] = cta_init(ctas[g], g, grid, mem[g], A::wbuf_floats(a));
  }
  const int U = A::n_updates(a);
  for (int u = 0; u < U; ++u)
    for (int s = 0; s < A::NSTAGES; ++s) {
      if (!A::stage_enabled(s, u, a)) continue;
      for (int g = 0; g < grid; ++g) {
        A::stage(s, u, ctas[g], user[g], a);
        stage_reset(ctas[g]);
      }
    }
  for (int g = 0; g < grid; ++g) { res_drain(ctas[g]); free(mem[g]); }
  return 0;
}

template <class A>
int frl_launch_tiles(const typename A::Args& a, cudaStream_t s) {
  struct One : A { };
  const int floats = cta_base_floats(A::wbuf_floats(a)) + A::user_floats(a) + 16;
  const int grid = A::grid(a, 1 << 30);
  float* mem = (float*)aligned_alloc(1024, ((size_t)floats * 4 + 1023) / 1024 * 1024);
  for (int g = 0; g < grid; ++g) {
    for (int i = 0; i < floats; ++i) mem[i] = NAN;
    Cta c;
    float* user = cta_init(c, g, grid, mem, A::wbuf_floats(a));
    A::stage(0, 0, c, user, a);
  }
  free(mem);
  (void)s;
  return 0;
}
#endif

// ------------------------------------------------------------------------------------------------
// tiny bump allocator over the user part of shared memory (16-B granularity)
// ------------------------------------------------------------------------------------------------
struct SmemBump {
  float* p;
  FRL_DEVM float* take(int floats) {
    float* r = p;
    p += (floats + 3) & ~3;
    return r;
  }
};
FRL_HD int pad4(int x) { return (x + 3) & ~3; }
