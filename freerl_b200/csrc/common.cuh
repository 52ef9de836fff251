// freerl_b200 — common device/host helpers (sm_100a).
//
// All kernel bodies are written in a "phase" style:
//
//     FRL_PAR(t) { ...work of thread t, no barrier inside... }
//     FRL_SYNC();
//
// On the GPU FRL_PAR runs its body once per thread and FRL_SYNC is __syncthreads().  When the same
// source is compiled with -DFRL_EMUL by g++ (tests/emul, TEST-ONLY, never shipped, never loaded by the
// product path) FRL_PAR is a loop over the block's threads, so indexing and arithmetic of every kernel
// can be checked against the oracle on a machine without a GPU.  Variables declared outside FRL_PAR are
// block-uniform by construction.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#ifndef FRL_EMUL
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define FRL_DEV __device__ __forceinline__
#define FRL_HD __host__ __device__ __forceinline__
#define FRL_DEVM __device__ __forceinline__
#define FRL_NOINL __device__ __noinline__
#define FRL_INLINE_ALT __device__ __forceinline__
#define FRL_SHD static __host__ __device__ __forceinline__
#define FRL_HDM __host__ __device__ __forceinline__
#define FRL_SDEV static __device__ __forceinline__
#define FRL_PAR(t) for (int t = (int)threadIdx.x, _frl_once = 0; _frl_once < 1; ++_frl_once)
#define FRL_SYNC() __syncthreads()
#else
#define FRL_DEV static inline
#define FRL_HD static inline
#define FRL_DEVM inline
#define FRL_NOINL static
#define FRL_INLINE_ALT static inline
#define FRL_SHD static inline
#define FRL_HDM inline
#define FRL_SDEV static inline
#define FRL_PAR(t) for (int t = 0; t < FRL_NT; ++t)
#define FRL_SYNC() ((void)0)
typedef void* cudaStream_t;
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline float __ldg(const float* p) { return *p; }
#endif

#ifndef FRL_NT
#define FRL_NT 256          // threads per CTA for every engine kernel
#endif
#include "../../include/freerl_b200.h"   // frl_layer_t / frl_net_t / argument structs (the C ABI)

// Segment table entry for per-tensor optimiser semantics (cautious AdamW mask mean, c_adamw.py:116).
struct frl_seg_t { int off, len, numel; };

enum { FRL_ACT_NONE = 0, FRL_ACT_RELU = 1, FRL_ACT_TANH = 2 };

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (fast mode: on-device sampling / exploration noise).
// ------------------------------------------------------------------------------------------------
FRL_HD void frl_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

FRL_HD float frl_u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)

// two standard normals from two 32-bit words (Box-Muller)
FRL_HD void frl_boxmuller(uint32_t a, uint32_t b, float* n0, float* n1) {
  float u1 = frl_u01(a), u2 = frl_u01(b);
  float r = sqrtf(-2.0f * logf(u1));
#if defined(__CUDA_ARCH__)
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);     // exact-range reduction in units of pi: none of cosf()'s large-argument slow path
  *n0 = r * cs;
  *n1 = r * sn;
#else
  float th = 6.28318530717958647692f * u2;
  *n0 = r * cosf(th);
  *n1 = r * sinf(th);
#endif
}

// N(0,1) for element (stream, idx) of draw `ctr`
FRL_HD float frl_randn(uint64_t seed, uint32_t stream, uint32_t ctr, uint32_t idx) {
  uint32_t o[4];
  frl_philox((uint32_t)seed, (uint32_t)(seed >> 32), idx >> 1, ctr, stream, 0x5eed5eedu, o);
  float a, b;
  frl_boxmuller(o[0], o[1], &a, &b);
  return (idx & 1) ? b : a;
}

// ------------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------------
#ifdef __cplusplus
extern "C" {
#endif
void frl_set_error(const char* fmt, ...);
#ifdef __cplusplus
}
#endif

#ifndef FRL_EMUL
#define FRL_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      frl_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                            \
    }                                                                                       \
  } while (0)
#endif
