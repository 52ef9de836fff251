// freerl_b200 — CTA-level training engine for small (hidden 128) MLPs on sm_100a.
//
// Building blocks used by every fused learn() kernel (DQN / SAC / TD3 / DDPG / PPO / multi-agent):
//   * Stager      : double-buffered TMA (cp.async.bulk + mbarrier) staging of one weight matrix at a time
//                   from HBM/L2 into shared memory, prefetching the next matrix while the current is used.
//   * gemm_rk     : C[R][N] = epi(A[R][K] * B[K][N] + bias)   (forward with B = W^T, backward-dX with B = W)
//                   4x4 register tiles, K split across threads, fixed-order shared-memory reduction.
//   * gemm_outer  : dW[M][N] = sum_r dY[r][m] X[r][n]  written to this CTA's gradient partial in global.
//   * reduce_grads / adam_update / polyak_update : cross-CTA fixed-order (deterministic, no float atomics)
//                   gradient reduction, global-norm clip, torch-exact Adam, target Polyak; keep the
//                   transposed weight mirrors in sync.
// Activations live in shared memory as row-major [R][width_pad] tiles (width_pad % 4 == 0, pads zero).
#pragma once
#include "common.cuh"

// The product build runs the GEMMs as fp32 FFMA on the CUDA cores (8 x 128 x 128 per op: far below the "genuine dense
// >= 256 x 256" bar the north star sets for tensor cores, and fp32 keeps the 1e-5 parity budget with room to spare).
// -DFRL_MMA_3XTF32 builds the experimental tensor-core flavour instead (mma.sync m16n8k8, error-compensated 3xTF32).
// Measured on B200 (fused SAC learn, B = 256, 64 CTAs): 92.7 us / learn vs 90.1 us for FFMA — the legacy mma.sync path
// issues slowly on sm_100a and the hi/lo operand split costs as many instructions as it saves — with 2-4x the rounding
// noise (one Adam-conditioned parameter in 2944 left the 1e-5 band in the TD3 parity test).  Kept as a recorded negative
// result, not shipped.
#if !defined(FRL_EMUL) && defined(FRL_MMA_3XTF32)
#define FRL_MMA 1
#endif

// bisect switches (debug): -DFRL_INL_GEMM / -DFRL_INL_MISC / -DFRL_INL_OPT force-inline a group again
#ifdef FRL_INL_GEMM
#define FRL_NI_GEMM FRL_INLINE_ALT
#else
#define FRL_NI_GEMM FRL_NOINL
#endif
#ifdef FRL_INL_MISC
#define FRL_NI_MISC FRL_INLINE_ALT
#else
#define FRL_NI_MISC FRL_NOINL
#endif
// Optimiser-stage helpers and the MLP drivers are inlined by default: measured on B200 (fused SAC learn, 64 CTAs)
//   all four groups out of line 108 us / learn, only OPT inlined 103.5, only MLP inlined 97.1, both inlined 93.6.
// (Out of line the stager context `Cta` is passed by reference and lives in local memory, and the net descriptors stop
//  being constant-bank operands.)  The GEMM microkernels and the misc helpers stay out of line: fully inlined, the kernel
//  was 490 KB of SASS and instruction-fetch bound.  -DFRL_NOINL_OPT / -DFRL_NOINL_MLP switch back for experiments.
#ifndef FRL_NOINL_OPT
#define FRL_NI_OPT FRL_INLINE_ALT
#else
#define FRL_NI_OPT FRL_NOINL
#endif
#ifndef FRL_NOINL_MLP
#define FRL_NI_MLP FRL_INLINE_ALT
#else
#define FRL_NI_MLP FRL_NOINL
#endif

// ------------------------------------------------------------------------------------------------
// exact (non-contracted) fp32 helpers so optimiser math rounds like torch's op-by-op kernels
// ------------------------------------------------------------------------------------------------
#ifndef FRL_EMUL
FRL_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
FRL_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
FRL_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }
FRL_DEV float fsqrt(float a) { return __fsqrt_rn(a); }
FRL_DEV double frl_dmul(double a, double b) { return __dmul_rn(a, b); }
FRL_DEV double frl_dadd(double a, double b) { return __dadd_rn(a, b); }
#else
static inline float fmul(float a, float b) { volatile float r = a * b; return r; }
static inline float fadd(float a, float b) { volatile float r = a + b; return r; }
static inline float fdiv(float a, float b) { return a / b; }
static inline float fsqrt(float a) { return sqrtf(a); }
static inline double frl_dmul(double a, double b) { volatile double r = a * b; return r; }
static inline double frl_dadd(double a, double b) { volatile double r = a + b; return r; }
#endif
FRL_DEV float xmul(float a, float b) { return fmul(a, b); }       // dtype-generic spellings for templated numpy-order arithmetic
FRL_DEV double xmul(double a, double b) { return frl_dmul(a, b); }

// ------------------------------------------------------------------------------------------------
// Packed fp32x2 FMA (Blackwell `fma.rn.f32x2`, SASS FFMA2): two IEEE fp32 FMAs per issued instruction — the same rounding
// as two FFMAs, so results are bit-identical to the scalar loops they replace.  The GEMM microkernels are FMA-issue bound
// (3-register FFMA issues every 2nd cycle per SM sub-partition), so halving the instruction count is what buys time.
//   fma2_bcast: (d0, d1) += a * (b0, b1)        scalar multiplicand broadcast by the instruction itself (Ra.F32 form)
//   fma2_elem : (d0, d1) += (a0, a1) * (b0, b1) element-wise pairs (both operands natural register pairs of an LDS.128)
// -DFRL_NO_FFMA2 keeps the scalar FFMA loops (A/B timing); the host emulation is scalar.
// ------------------------------------------------------------------------------------------------
#if !defined(FRL_EMUL) && !defined(FRL_NO_FFMA2)
FRL_DEV void fma2_bcast(float& d0, float& d1, float a, float b0, float b1) {
  unsigned long long B, C;
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(d0), "f"(d1));
  asm("{\n\t.reg .b64 aa;\n\tmov.b64 aa, {%1, %1};\n\tfma.rn.f32x2 %0, aa, %2, %0;\n\t}" : "+l"(C) : "f"(a), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(C));
}
FRL_DEV void fma2_elem(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long A, B, C;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(d0), "f"(d1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(C) : "l"(A), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(C));
}
#else
FRL_DEV void fma2_bcast(float& d0, float& d1, float a, float b0, float b1) { d0 += a * b0; d1 += a * b1; }
FRL_DEV void fma2_elem(float& d0, float& d1, float a0, float a1, float b0, float b1) { d0 += a0 * b0; d1 += a1 * b1; }
#endif

// ------------------------------------------------------------------------------------------------
// Transposed-mirror layout (`pt`): per layer WT[in_pad][wt_ld] followed by bias[out_pad]; ONE staged copy of a layer
// serves forward (Y = X WT) and backward (dX = dY WT^T, read "transposed").
//   FFMA build: the row stride is padded so that wt_ld / 4 is odd — rows k, k+1, ... start in different 16-B bank
//     groups and the backward float4 reads of consecutive rows are free of bank conflicts; activation tiles are dense.
//   3xTF32 build: wt_ld = 8 (mod 32) and activation strides = 4 (mod 32) make the scalar fragment loads (lane 4g + t
//     reads row t / column g style patterns) conflict free.
// The host asks the library for the stride (frl_wt_ld), so both layouts work with the same Python code.
// ------------------------------------------------------------------------------------------------
#ifdef FRL_MMA
FRL_HD int wt_ld_of(int out_pad) { return out_pad + ((40 - (out_pad & 31)) & 31); }
FRL_HD int act_ld(int width) { return width + ((36 - (width & 31)) & 31); }
#else
FRL_HD int wt_ld_of(int out_pad) { return ((out_pad >> 2) & 1) ? out_pad : out_pad + 4; }
FRL_HD int act_ld(int width) { return width; }
#endif
FRL_HD int wt_ld(const frl_layer_t& L) { return wt_ld_of(L.out_pad); }
FRL_HD int wt_bias(const frl_layer_t& L) { return L.in_pad * wt_ld(L); }               // offset of the bias inside the layer image
FRL_HD int wt_floats(const frl_layer_t& L) { return L.in_pad * wt_ld(L) + L.out_pad; }

#ifndef FRL_TMA_CHUNK
#define FRL_TMA_CHUNK 16384
#endif

// ------------------------------------------------------------------------------------------------
// CTA context
// ------------------------------------------------------------------------------------------------
struct Cta {
  int cta, ncta;            // this CTA's index / grid size
  float* smem;              // dynamic shared memory base (1024-B aligned)
  // ---- stager (block-uniform state, replicated in every thread's registers) ----
  float* wbuf0;             // two weight staging buffers in smem (scalars: no dynamically indexed
  float* wbuf1;             //  arrays, they would push the whole context into local memory)
  uint64_t* bar;            // two mbarriers in smem
  uint32_t phase0, phase1;
  const float* pend_ptr;    // global source of the outstanding prefetch (or nullptr)
  int pend_buf;
  int next_buf;
  // ---- resident slots (fused actor-critic kernel): wbuf0 / wbuf1 each hold a whole 3-layer head ----
  uint32_t sphase, spend;   // per (slot, layer) mbarrier parity / copy-outstanding bits (bit = slot * 3 + layer)
  const float* stag0;       // global source (pt + wt_off of the head's first layer) resident in slot 0 / 1, or nullptr
  const float* stag1;
  int mode;                 // algorithm-private flag (AcAlgo: 1 = resident slots)
  float* red;               // [FRL_NT*16] reduction scratch in smem
  long long* dbg;           // optional timestamp sink (frl_debug_set_timing), CTA 0 / thread 0 only
  int dbg_n;
};

// Debug timestamps: CTA 0 / thread 0 appends (id, clock) pairs.  Disabled (nullptr) in normal runs.
#if !defined(FRL_EMUL) && defined(FRL_TRACE)
__device__ long long* frl_dbg_ptr = nullptr;
__device__ int frl_trace_cta = 0;           // which CTA writes the op trace (frl_debug_set_trace_cta)
FRL_DEV void stamp(Cta&, int) {}            // the op trace owns the sink in -DFRL_TRACE builds
#elif !defined(FRL_EMUL)
__device__ long long* frl_dbg_ptr = nullptr;
FRL_DEV void stamp(Cta& c, int id) {
  if (c.dbg && c.cta == 0 && threadIdx.x == 0 && c.dbg_n < 2000) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    c.dbg[2 * c.dbg_n] = id; c.dbg[2 * c.dbg_n + 1] = t;
  }
  if (c.dbg) c.dbg_n++;
}
#else
FRL_DEV void stamp(Cta&, int) {}
#endif

// fine-grained op tracing (debug builds only: -DFRL_TRACE): (id, clock64) pairs from CTA 0 / thread `who`
#if defined(FRL_TRACE) && !defined(FRL_EMUL)
__shared__ int frl_trace_n;         // per-CTA event counter in smem: a probe costs ~50 clk (a global atomic cost ~1000)
FRL_DEV void trace(int id, int who = 0) {
  if (frl_dbg_ptr && (int)blockIdx.x == frl_trace_cta && (int)threadIdx.x == who) {
    const int k = frl_trace_n;
    frl_trace_n = k + 1;
    if (k < 2000) { frl_dbg_ptr[2 * k] = id; frl_dbg_ptr[2 * k + 1] = clock64(); }
  }
}
#else
#define trace(...) ((void)0)
#endif

// ------------------------------------------------------------------------------------------------
// TMA staging
// ------------------------------------------------------------------------------------------------
#ifndef FRL_EMUL
FRL_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

FRL_DEV void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
FRL_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
FRL_DEV void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
FRL_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "FRL_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra FRL_DONE_%=;\n\t"
      "bra FRL_WAIT_%=;\n\t"
      "FRL_DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
FRL_DEV void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#endif

// Carve the stager + scratch out of dynamic smem.  Returns the first free float after them.
FRL_DEV float* cta_init(Cta& c, int cta, int ncta, float* smem, int wbuf_floats) {
  c.cta = cta; c.ncta = ncta; c.smem = smem;
  c.wbuf0 = smem;
  c.wbuf1 = smem + wbuf_floats;
  c.red = smem + 2 * wbuf_floats;
  c.bar = (uint64_t*)(c.red + FRL_NT * 16);
  c.phase0 = c.phase1 = 0;
  c.pend_ptr = nullptr; c.pend_buf = 0; c.next_buf = 0;
  c.sphase = c.spend = 0; c.stag0 = c.stag1 = nullptr; c.mode = 0;
  c.dbg = nullptr; c.dbg_n = 0;
#ifndef FRL_EMUL
  c.dbg = frl_dbg_ptr;
#ifdef FRL_TRACE
  if (threadIdx.x == 0) frl_trace_n = 0;
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&c.bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();
#endif
  return (float*)(c.bar + 8);       // eight 8-B mbarriers (2 streaming + 6 resident), alignment preserved
}

// smem floats consumed by cta_init
FRL_HD int cta_base_floats(int wbuf_floats) { return 2 * wbuf_floats + FRL_NT * 16 + 16; }

// thread 0 only; out of line (every acquire / prefetch site would otherwise carry an unrolled copy of the chunk loop)
FRL_NI_MISC void stage_issue_t0(float* dstf, uint64_t* bar, const float* src, int bytes) {
#ifndef FRL_EMUL
  fence_proxy_async();
  mbar_expect_tx(bar, (uint32_t)bytes);
  char* dst = (char*)dstf;
  const char* sp = (const char*)src;
#pragma unroll 1
  for (int off = 0; off < bytes; off += FRL_TMA_CHUNK) {      // several bulk copies in flight on one mbarrier
    const int nb = (bytes - off) < FRL_TMA_CHUNK ? (bytes - off) : FRL_TMA_CHUNK;
    tma_bulk_g2s(dst + off, sp + off, (uint32_t)nb, bar);
  }
#else
  (void)bar;
  memcpy(dstf, src, (size_t)bytes);
#endif
}
FRL_DEV void stage_issue(Cta& c, int buf, const float* src, int bytes) {
#ifndef FRL_EMUL
  if (threadIdx.x == 0)
#endif
    stage_issue_t0(buf ? c.wbuf1 : c.wbuf0, c.bar + buf, src, bytes);
}

FRL_DEV void stage_wait(Cta& c, int buf) {
#ifndef FRL_EMUL
  mbar_wait(c.bar + buf, buf ? c.phase1 : c.phase0);
#endif
  if (buf) c.phase1 ^= 1u; else c.phase0 ^= 1u;
}

// Start fetching `src` into the idle buffer.  Precondition: every thread is past its last read of that
// buffer (all engine ops end with FRL_SYNC).  At most one prefetch is outstanding.
FRL_DEV void stage_prefetch(Cta& c, const float* src, int bytes) {
  if (src == nullptr || c.pend_ptr != nullptr) return;
  stage_issue(c, c.next_buf, src, bytes);
  c.pend_ptr = src;
  c.pend_buf = c.next_buf;
}

// Returns the smem copy of `src` (waits for the matching prefetch, or fetches now).
FRL_DEV const float* stage_acquire(Cta& c, const float* src, int bytes) {
  if (c.pend_ptr != nullptr && c.pend_ptr != src) {   // stale hint: drain it to keep barrier phases aligned
    stage_wait(c, c.pend_buf);
    c.pend_ptr = nullptr;
    c.next_buf = c.pend_buf ^ 1;
    FRL_SYNC();
  }
  if (c.pend_ptr == nullptr) {
    stage_issue(c, c.next_buf, src, bytes);
    c.pend_buf = c.next_buf;
  }
  stage_wait(c, c.pend_buf);
  c.pend_ptr = nullptr;
  c.next_buf = c.pend_buf ^ 1;
  return c.pend_buf ? c.wbuf1 : c.wbuf0;
}

// Must be called when weights may have changed under a buffer we would otherwise trust (after grid sync).
FRL_DEV void stage_reset(Cta& c) {
  if (c.pend_ptr != nullptr) {
    stage_wait(c, c.pend_buf);
    c.pend_ptr = nullptr;
    c.next_buf = c.pend_buf ^ 1;
  }
}

// ------------------------------------------------------------------------------------------------
// float4 helpers
// ------------------------------------------------------------------------------------------------
FRL_DEV float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
FRL_DEV void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// L2 residency hints for streaming kernels that read an array twice (frl_adv_norm): the first pass loads with evict_last so that the
// array stays in the 126 MB L2, the second pass loads and stores with evict_first so that the output does not push the unread part out
#ifndef FRL_EMUL
FRL_DEV unsigned long long l2_policy_evict_last() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
FRL_DEV unsigned long long l2_policy_evict_first() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
FRL_DEV float4 ld4_hint(const float* p, unsigned long long pol) {
  if (pol == 0ull) return ld4(p);
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
FRL_DEV void st4_hint(float* p, float4 v, unsigned long long pol) {
  if (pol == 0ull) { st4(p, v); return; }
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
#else
FRL_DEV unsigned long long l2_policy_evict_last() { return 0; }
FRL_DEV unsigned long long l2_policy_evict_first() { return 0; }
FRL_DEV float4 ld4_hint(const float* p, unsigned long long) { return ld4(p); }
FRL_DEV void st4_hint(float* p, float4 v, unsigned long long) { st4(p, v); }
#endif
// shared-memory flavours: explicit ld.shared / st.shared (the address-space is not always inferable once the
// hot loops live in non-inlined functions; `__builtin_assume(__isShared(p))` proved fragile — the optimiser used it
// to delete the K loop of the de-inlined gemm — so the space is spelled out in PTX instead).
#ifndef FRL_EMUL
FRL_DEV float4 lds4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
FRL_DEV void sts4(float* p, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
FRL_DEV float lds1(const float* p) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
  return v;
}
#else
FRL_DEV float4 lds4(const float* p) { return ld4(p); }
FRL_DEV void sts4(float* p, float4 v) { st4(p, v); }
FRL_DEV float lds1(const float* p) { return *p; }
#endif
FRL_DEV float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// (The GEMMs below are templated on HM only to give the tanh variants of the PPO update / inference their OWN instantiations: in
// the HM = 0 ones no caller ever passes FRL_ACT_TANH, so the compiler's whole-program constant propagation drops the tanh branch
// from the shared epilogues exactly as it did before the `tanh` switch existed.)
FRL_DEV float apply_act(float v, int act) {
  if (act == FRL_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == FRL_ACT_TANH) return tanhf(v);
  return v;
}

enum { EPI_BIAS_ACT = 0, EPI_RELU_MASK = 1 };

// ------------------------------------------------------------------------------------------------
// Shared-memory "word address" helpers.  The GEMM microkernels do all smem pointer arithmetic in 32-bit word
// offsets relative to their operand bases (no 64-bit generic pointer math + cvta per access in the hot loop).
// On the GPU an `sptr` is a 32-bit shared-space BYTE address; in the emulation it is a host pointer.
// ------------------------------------------------------------------------------------------------
#ifndef FRL_EMUL
typedef uint32_t sptr;
FRL_DEV sptr sp_of(const float* p) { return smem_u32(p); }
FRL_DEV float4 sp_ld4(sptr base, int word) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + 4u * (uint32_t)word));
  return v;
}
FRL_DEV float sp_ld1(sptr base, int word) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + 4u * (uint32_t)word));
  return v;
}
FRL_DEV void sp_st4(sptr base, int word, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + 4u * (uint32_t)word), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
FRL_DEV void sp_st1(sptr base, int word, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + 4u * (uint32_t)word), "f"(v) : "memory");
}
#else
typedef const float* sptr;
FRL_DEV sptr sp_of(const float* p) { return p; }
FRL_DEV float4 sp_ld4(sptr base, int word) { return ld4(base + word); }
FRL_DEV float sp_ld1(sptr base, int word) { return base[word]; }
FRL_DEV void sp_st4(sptr base, int word, float4 v) { st4(const_cast<float*>(base) + word, v); }
FRL_DEV void sp_st1(sptr base, int word, float v) { const_cast<float*>(base)[word] = v; }
#endif

// integer helpers — the GEMM index decode deliberately avoids runtime integer division: on the measured critical path a
// handful of MUFU.RCP-based divisions per thread cost ~0.5 us per op (28 ops per learn).
FRL_DEV unsigned floor_log2(unsigned x) {
#ifndef FRL_EMUL
  return 31u - (unsigned)__clz((int)x);
#else
  unsigned s = 0; while ((2u << s) <= x) ++s; return s;
#endif
}
FRL_DEV unsigned ceil_pow2_log(unsigned x) { const unsigned f = floor_log2(x); return ((1u << f) == x) ? f : f + 1; }
// log2 of the reduction split: the largest s <= 3 with (2^sh_tiles << s) <= FRL_NT and 2^s <= nchunk (closed form, no loop)
template <unsigned N> struct FrlLog2 { static const unsigned value = 1 + FrlLog2<N / 2>::value; };
template <> struct FrlLog2<1> { static const unsigned value = 0; };
FRL_DEV unsigned split_log(unsigned sh_tiles, unsigned nchunk) {
  constexpr unsigned NTL = FrlLog2<FRL_NT>::value;
  unsigned s = sh_tiles >= NTL ? 0u : NTL - sh_tiles;
  const unsigned fk = nchunk ? floor_log2(nchunk) : 0u;
  if (fk < s) s = fk;
  return s < 3u ? s : 3u;
}
FRL_DEV unsigned div_fast(unsigned t, unsigned d) { return t / d; }   // (a power-of-two shift fast path measured 0.5 us / learn slower)

// ------------------------------------------------------------------------------------------------
// gemm_rk:  C[r][n] = epi( sum_k A[r][k] * Bs[k][n] (+ bias[n]) ),  r < R, n < N_pad, k < K_pad
//   A  : smem, row-major, leading dim lda (multiple of 4), columns [0,K_pad) readable & finite
//   Bs : smem, row-major [K_pad][N_pad] (rows >= logical K are zero)
//   epi: EPI_BIAS_ACT  -> act(acc + bias[n])            (bias in smem, may be null)
//        EPI_RELU_MASK -> acc * (mask[r][n] > 0)        (backward through ReLU; mask = stored activation)
//   C  : smem, leading dim ldc.  Must not alias A.
// Work item = (k-split ks, tile); the tile count is padded to a power of two so item -> (ks, tile) is shift/mask, and a
// tile is (n-tile, row-tile) with the row-tile in the low bit(s).  4x4 register tiles, K split `ksplit` (power of two)
// ways across the CTA, fixed-order reduction of the split partials through shared memory.
// (NOT inlined: one copy of the hot loop keeps the persistent kernels' instruction footprint inside the I-cache;
//  the fully inlined build was 490 KB of SASS and spent most cycles in `no_instruction` stalls.)
// ------------------------------------------------------------------------------------------------
// fixed-order reduction of the K-split partials red[ks][R][N_pad] + epilogue -> C   (second phase of gemm_rk / gemm_nt)
// HM (compile time, default 0): what a MASK epilogue means — 0: relu'(mask) = mask > 0; 1: tanh'(mask) = 1 - mask^2 (the `tanh`
// switch of PPO_file/PPO_with_tricks.py).  A template parameter, not a run-time one: the HM = 0 instantiations — every kernel
// except the tanh variants of the PPO update / inference — compile to exactly the code they had before the switch existed.
template <int R, int HM = 0>
FRL_DEV void gemm_finish(sptr sR, unsigned ksplit, int N_pad, int epi, int act, bool has_bias, sptr sBias, sptr sM, int ldm,
                         sptr sC, int ldc) {
  const unsigned nt = (unsigned)N_pad >> 2;
  FRL_PAR(t) {
    // element e -> (row r, column group j): e = r * nt + j; one division per thread, then an incremental walk
    const unsigned ne = R * nt;
    if ((unsigned)t < ne) {
      unsigned r = div_fast((unsigned)t, nt), j = (unsigned)t - r * nt;
      const unsigned dr = div_fast(FRL_NT, nt), dj = FRL_NT - dr * nt;
      for (unsigned e = (unsigned)t; e < ne; e += FRL_NT) {
        const int n0 = (int)j * 4;
        int w = (int)r * N_pad + n0;
        float4 s = sp_ld4(sR, w);
        for (unsigned ks = 1; ks < ksplit; ++ks) { w += R * N_pad; s = f4add(s, sp_ld4(sR, w)); }
        float o[4] = {s.x, s.y, s.z, s.w};
        if (epi == EPI_BIAS_ACT) {
          if (has_bias) { const float4 bv = sp_ld4(sBias, n0); o[0] += bv.x; o[1] += bv.y; o[2] += bv.z; o[3] += bv.w; }
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = apply_act(o[q], act);
        } else if (HM) {
          const float4 mv = sp_ld4(sM, (int)r * ldm + n0);
          o[0] = o[0] * (1.f - mv.x * mv.x); o[1] = o[1] * (1.f - mv.y * mv.y);
          o[2] = o[2] * (1.f - mv.z * mv.z); o[3] = o[3] * (1.f - mv.w * mv.w);
        } else {
          const float4 mv = sp_ld4(sM, (int)r * ldm + n0);
          o[0] = mv.x > 0.f ? o[0] : 0.f; o[1] = mv.y > 0.f ? o[1] : 0.f;
          o[2] = mv.z > 0.f ? o[2] : 0.f; o[3] = mv.w > 0.f ? o[3] : 0.f;
        }
        sp_st4(sC, (int)r * ldc + n0, make_float4(o[0], o[1], o[2], o[3]));
        r += dr; j += dj;
        if (j >= nt) { j -= nt; ++r; }
      }
    }
  }
  FRL_SYNC();
}

#ifndef FRL_MMA
template <int R, int HM = 0>
FRL_NI_GEMM void gemm_rk(float* red, const float* A, int lda, int K_pad, const float* Bs, int ldb, int N_pad, const float* bias,
                         int epi, int act, const float* mask, int ldm, float* C, int ldc) {
  constexpr unsigned RT = R / 4, RT_SH = (RT == 1 ? 0 : (RT == 2 ? 1 : 2));
  const unsigned nt = (unsigned)N_pad >> 2;
  const unsigned tiles = nt * RT;
  const unsigned sh_t = ceil_pow2_log(tiles), tiles_p2 = 1u << sh_t;
  const unsigned nchunk = (unsigned)K_pad >> 2;
  const unsigned sh_k = split_log(sh_t, nchunk);                // ksplit = 2^sh_k <= min(NT / tiles_p2, nchunk, 8)
  const unsigned ksplit = 1u << sh_k;
  const unsigned items = tiles_p2 << sh_k;
  const sptr sA = sp_of(A), sB = sp_of(Bs), sR = sp_of(red), sC = sp_of(C);
  const bool has_bias = bias != nullptr;
  const sptr sBias = sp_of(has_bias ? bias : Bs), sM = sp_of(mask ? mask : Bs);
  // (A dedicated one-warp-per-row + shuffle-tree path for the narrow N_pad <= 8 heads was measured 2.7 us / learn SLOWER
  //  than this generic tiling on B200, inlined or not, and was dropped.)
  trace(50);
  FRL_PAR(t) {
    for (unsigned item = (unsigned)t; item < items; item += FRL_NT) {
      const unsigned ks = item >> sh_t, tile = item & (tiles_p2 - 1);
      if (tile >= tiles) continue;
      const int n0 = (int)(tile >> RT_SH) * 4, r0 = (int)(tile & (RT - 1)) * 4;
      const int kc0 = (int)((ks * nchunk) >> sh_k), kc1 = (int)(((ks + 1) * nchunk) >> sh_k);
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      int wb = kc0 * 4 * ldb + n0;            // word offset of Bs[k][n0]
      int wa = r0 * lda + kc0 * 4;            // word offset of A[r0][k]
      const int wb_step = 4 * ldb;
#pragma unroll 2
      for (int kc = kc0; kc < kc1; ++kc) {
        const float4 b0 = sp_ld4(sB, wb);
        const float4 b1 = sp_ld4(sB, wb + ldb);
        const float4 b2 = sp_ld4(sB, wb + 2 * ldb);
        const float4 b3 = sp_ld4(sB, wb + 3 * ldb);
        float4 av[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = sp_ld4(sA, wa + i * lda);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = av[i];
          fma2_bcast(acc[i][0], acc[i][1], a.x, b0.x, b0.y); fma2_bcast(acc[i][2], acc[i][3], a.x, b0.z, b0.w);
          fma2_bcast(acc[i][0], acc[i][1], a.y, b1.x, b1.y); fma2_bcast(acc[i][2], acc[i][3], a.y, b1.z, b1.w);
          fma2_bcast(acc[i][0], acc[i][1], a.z, b2.x, b2.y); fma2_bcast(acc[i][2], acc[i][3], a.z, b2.z, b2.w);
          fma2_bcast(acc[i][0], acc[i][1], a.w, b3.x, b3.y); fma2_bcast(acc[i][2], acc[i][3], a.w, b3.z, b3.w);
        }
        wb += wb_step;
        wa += 4;
      }
      if (ksplit == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v = acc[i][j];
            if (epi == EPI_BIAS_ACT) v = apply_act(v + (has_bias ? sp_ld1(sBias, n0 + j) : 0.f), act);
            else v = (sp_ld1(sM, (r0 + i) * ldm + n0 + j) > 0.f) ? v : 0.f;
            o[j] = v;
          }
          sp_st4(sC, (r0 + i) * ldc + n0, make_float4(o[0], o[1], o[2], o[3]));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          sp_st4(sR, ((int)ks * R + r0 + i) * N_pad + n0, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
      }
    }
  }
  trace(51);
  FRL_SYNC();
  trace(52);
  if (ksplit > 1) gemm_finish<R, HM>(sR, ksplit, N_pad, epi, act, has_bias, sBias, sM, ldm, sC, ldc);
  trace(53);
}

// ------------------------------------------------------------------------------------------------
// gemm_nt:  C[r][k] = epi( sum_n A[r][n] * Bs[k][n] ),  r < R, k < K_out, n < N_red     (backward dX = dY * W from the
//   forward image Bs = WT[in_pad][ldb]: both operands are contiguous along the reduction index n)
//   Each thread owns rows r0..r0+3 and the four k's {kl, kl+KL, kl+2KL, kl+3KL} (KL = K_out/4): consecutive lanes read
//   consecutive rows of Bs, which the padded row stride (wt_ld) spreads over all bank groups.  N_red is split
//   2^sh_n ways across the CTA; the partials take the same fixed-order reduction + epilogue as gemm_rk.
//   epi: EPI_RELU_MASK (mask = stored activation) or EPI_BIAS_ACT with act NONE / no bias (plain product).
// ------------------------------------------------------------------------------------------------
template <int R, int HM = 0>
FRL_NI_GEMM void gemm_nt(float* red, const float* A, int lda, int N_red, const float* Bs, int ldb, int K_out, int epi,
                         const float* mask, int ldm, float* C, int ldc) {
  constexpr unsigned RT = R / 4, RT_SH = (RT == 1 ? 0 : (RT == 2 ? 1 : 2));
  const unsigned KL = (unsigned)K_out >> 2;
  const unsigned sh_kl = ceil_pow2_log(KL), klp2 = 1u << sh_kl;
  const unsigned tiles_p2 = klp2 << RT_SH;
  const unsigned nchunk = (unsigned)N_red >> 2;
  const unsigned sh_n = split_log(sh_kl + RT_SH, nchunk);
  const unsigned nsplit = 1u << sh_n;
  const unsigned items = tiles_p2 << sh_n;
  const sptr sA = sp_of(A), sB = sp_of(Bs), sR = sp_of(red), sC = sp_of(C);
  const sptr sM = sp_of(mask ? mask : Bs);
  trace(60);
  FRL_PAR(t) {
    for (unsigned item = (unsigned)t; item < items; item += FRL_NT) {
      const unsigned kl = item & (klp2 - 1), rt = (item >> sh_kl) & (RT - 1), ns = item >> (sh_kl + RT_SH);
      if (kl >= KL) continue;
      const int r0 = (int)rt * 4;
      const int nc0 = (int)((ns * nchunk) >> sh_n), nc1 = (int)(((ns + 1) * nchunk) >> sh_n);
      // two interleaved partial sums per output (even / odd reduction index): the pairs (a.x, a.y) * (b.x, b.y) are natural
      // register pairs of the LDS.128 operands, so each FFMA2 retires two products; folded as lo + hi at the end
      float acc[4][4], acc_hi[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; acc_hi[i][j] = 0.f; }
      int wa = r0 * lda + nc0 * 4;                  // A[r0][n]
      int wb = (int)kl * ldb + nc0 * 4;             // Bs[kl][n]
      const int kstep = (int)KL * ldb;
#pragma unroll 2
      for (int nc = nc0; nc < nc1; ++nc) {
        float4 av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = sp_ld4(sA, wa + i * lda);
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = sp_ld4(sB, wb + j * kstep);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            fma2_elem(acc[i][j], acc_hi[i][j], av[i].x, av[i].y, bv[j].x, bv[j].y);
            fma2_elem(acc[i][j], acc_hi[i][j], av[i].z, av[i].w, bv[j].z, bv[j].w);
          }
        wa += 4; wb += 4;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += acc_hi[i][j];
      if (nsplit == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = (int)kl + j * (int)KL;
            float v = acc[i][j];
            if (epi == EPI_RELU_MASK) {
              if (HM) { const float m = sp_ld1(sM, (r0 + i) * ldm + k); v = v * (1.f - m * m); }
              else v = (sp_ld1(sM, (r0 + i) * ldm + k) > 0.f) ? v : 0.f;
            }
            sp_st1(sC, (r0 + i) * ldc + k, v);
          }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) sp_st1(sR, ((int)ns * R + r0 + i) * K_out + (int)kl + j * (int)KL, acc[i][j]);
      }
    }
  }
  trace(61);
  FRL_SYNC();
  trace(62);
  if (nsplit > 1) gemm_finish<R, HM>(sR, nsplit, K_out, epi, FRL_ACT_NONE, false, sB, sM, ldm, sC, ldc);
  trace(63);
}

// ------------------------------------------------------------------------------------------------
// gemm_outer: G[m][n] (+)= sum_{r<R} dY[r][m] * X[r][n]   for m < M_pad, n < N_pad, written to GLOBAL
//   G has leading dim N_pad (the trainable W layout [out_pad][in_pad]).  Columns n >= N_real are forced
//   to zero (X pad columns may alias neighbouring fields).  Also the bias gradient gb[m] (+)= sum_r dY[r][m].
// ------------------------------------------------------------------------------------------------
template <int R>
FRL_NI_GEMM void gemm_outer(const float* dY, int ldy, int M_pad, const float* X, int ldx, int N_pad, int N_real,
                            float* G, float* gb, bool accumulate) {
  const unsigned mt = (unsigned)M_pad >> 2, nt = (unsigned)N_pad >> 2;
  const unsigned tiles = mt * nt;
  const sptr sY = sp_of(dY), sX = sp_of(X);
  trace(70);
  FRL_PAR(t) {
    if ((unsigned)t < tiles) {
      // tile -> (mi, ni): one division per thread, then an incremental walk (tile += FRL_NT)
      unsigned mi = div_fast((unsigned)t, nt), ni = (unsigned)t - mi * nt;
      const unsigned dm = div_fast(FRL_NT, nt), dn = FRL_NT - dm * nt;
      for (unsigned tile = (unsigned)t; tile < tiles; tile += FRL_NT) {
        const int m0 = (int)mi * 4, n0 = (int)ni * 4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 a = sp_ld4(sY, r * ldy + m0);
          const float4 b = sp_ld4(sX, r * ldx + n0);
          fma2_bcast(acc[0][0], acc[0][1], a.x, b.x, b.y); fma2_bcast(acc[0][2], acc[0][3], a.x, b.z, b.w);
          fma2_bcast(acc[1][0], acc[1][1], a.y, b.x, b.y); fma2_bcast(acc[1][2], acc[1][3], a.y, b.z, b.w);
          fma2_bcast(acc[2][0], acc[2][1], a.z, b.x, b.y); fma2_bcast(acc[2][2], acc[2][3], a.z, b.z, b.w);
          fma2_bcast(acc[3][0], acc[3][1], a.w, b.x, b.y); fma2_bcast(acc[3][2], acc[3][3], a.w, b.z, b.w);
        }
        float* gp = G + m0 * N_pad + n0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n0 + j >= N_real) acc[i][j] = 0.f;
          float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          if (accumulate) v = f4add(v, ld4(gp));
          st4(gp, v);
          gp += N_pad;
        }
        mi += dm; ni += dn;
        if (ni >= nt) { ni -= nt; ++mi; }
      }
    }
    // bias gradient
    for (int m = t; m < M_pad; m += FRL_NT) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) s += sp_ld1(sY, r * ldy + m);
      if (accumulate) s += gb[m];
      gb[m] = s;
    }
  }
  trace(71);
  FRL_SYNC();
  trace(72);
}

#endif   // !FRL_MMA

// ------------------------------------------------------------------------------------------------
// Tensor-core flavours of the three microkernels (CUDA build; -DFRL_NO_MMA keeps the FFMA ones above).
// The batch tile has R = 8 rows, so every product is shaped  [16 x 8] += [16 x 8k] * [8k x 8]  with the 128-wide feature
// dimension as M, the 8 batch rows as N and the reduction as K: exactly mma.sync m16n8k8 with NO padding waste, one
// m-tile (16 output features) per warp, the whole reduction inside the warp — no K-split, no cross-warp reduction pass,
// one barrier per op.  Operands are fp32 in shared memory; each is split on the fly into tf32 hi + lo parts and the
// product is formed as hi*hi + (hi*lo + lo*hi) in three fp32 accumulators ("3xTF32": per-product error ~2^-21, the same
// order as the reordering noise of an fp32 FFMA chain; a plain tf32 product would miss the 1e-5 parity budget).
//   A fragment (row-major 16x8): a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);   g = lane / 4, t = lane % 4
//   B fragment (8x8):            b0 (k = t, n = g)   b1 (k = t+4, n = g)
//   C fragment (16x8):           c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
// ------------------------------------------------------------------------------------------------
#ifdef FRL_MMA
FRL_DEV void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  // hi = x rounded to tf32's 10 mantissa bits (round half away, on the integer pipe: cvt.rna.tf32.f32 runs at 1/8 rate and
  // was the bottleneck of the whole k-step); lo = x - hi is exact, the tensor core reads its top 10 mantissa bits.
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// d += a * b  (tf32 operands, fp32 accumulate).  Not volatile: a pure function of its operands, free to be scheduled.
FRL_DEV void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// Shared-memory load WITHOUT volatile / memory clobber: inside the (non-inlined) GEMM microkernels nothing writes the
// operands between function entry and the closing barrier, so the compiler may hoist and software-pipeline these freely
// (the volatile flavour pins every load behind the previous k-step's math: 200+ clk per k-step instead of ~60).
FRL_DEV float lds_op(uint32_t byte_addr) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(byte_addr));
  return v;
}
// One k-step of the 3xTF32 product.  The tensor core adds into its fp32 accumulator with truncation; chained over the
// 16 k-steps of a 128-deep reduction that bias reaches ~1e-6 relative.  So the dominant hi*hi term (exact products) is
// produced against a ZERO accumulator and folded into `acc` with round-to-nearest FADDs, like an FFMA chain; only the
// 2^-11-times-smaller cross terms are chained inside the tensor core (two chains, to halve the dependent latency).
FRL_DEV void mma3(float (&acc)[4], float (&s1)[4], float (&s2)[4], const float (&af)[4], const float (&bf)[2]) {
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) tf32_split(af[i], ah[i], al[i]);
#pragma unroll
  for (int i = 0; i < 2; ++i) tf32_split(bf[i], bh[i], bl[i]);
  // every MMA starts from a zero accumulator: no HMMA -> HMMA dependency (its latency is ~100 clk on this path; chained,
  // a lone warp spent 120 clk per k-step), the sums are carried by round-to-nearest FADDs
  float d[4] = {0.f, 0.f, 0.f, 0.f}, e1[4] = {0.f, 0.f, 0.f, 0.f}, e2[4] = {0.f, 0.f, 0.f, 0.f};
  mma_tf32(d, ah, bh);
  mma_tf32(e1, ah, bl);
  mma_tf32(e2, al, bh);
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i] += d[i]; s1[i] += e1[i]; s2[i] += e2[i]; }
}

// out[q] for the m-tile `mt` of C[r][m] = sum_k W(m, k) X[r][k]: rows m0 = 16 mt + g, m1 = m0 + 8 (clamped to M - 1 for the
// loads), batch rows 2t, 2t + 1.  W(m, k) lives at byte wB + 4 (m wsm + k wsk)  (forward: wsm = 1, wsk = ldb; backward on
// the same image: wsm = ldb, wsk = 1), X[r][k] at xB + 4 (r ldx + k).  Kred % 4 == 0; a trailing half step is zero filled.
FRL_DEV void mma_tile(uint32_t wB, int wsm, int wsk, uint32_t xB, int ldx, int M, int Kred, int mt, int g, int t, float (&out)[4]) {
  const int m0 = mt * 16 + g, m1 = m0 + 8;
  const int m0c = m0 < M ? m0 : M - 1, m1c = m1 < M ? m1 : M - 1;
  uint32_t pa0 = wB + 4u * (uint32_t)(m0c * wsm + t * wsk), pa1 = wB + 4u * (uint32_t)(m1c * wsm + t * wsk);
  uint32_t pb = xB + 4u * (uint32_t)(g * ldx + t);
  const uint32_t k4 = 16u * (uint32_t)wsk, k8 = 32u * (uint32_t)wsk;
  const int nfull = Kred >> 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int ks = 0; ks < nfull; ++ks) {
    const float af[4] = {lds_op(pa0), lds_op(pa1), lds_op(pa0 + k4), lds_op(pa1 + k4)};
    const float bf[2] = {lds_op(pb), lds_op(pb + 16u)};
    pa0 += k8; pa1 += k8; pb += 32u;
    mma3(acc, s1, s2, af, bf);
  }
  if (Kred & 4) {
    const float af[4] = {lds_op(pa0), lds_op(pa1), 0.f, 0.f};
    const float bf[2] = {lds_op(pb), 0.f};
    mma3(acc, s1, s2, af, bf);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) out[q] = acc[q] + (s1[q] + s2[q]);
}

// C[r][n] = epi( sum_k A[r][k] * Bs[k][n] (+ bias[n]) )    (same contract as the FFMA gemm_rk; `red` unused)
template <int R, int HM = 0>
FRL_NI_GEMM void gemm_rk(float* red, const float* A, int lda, int K_pad, const float* Bs, int ldb, int N_pad, const float* bias,
                         int epi, int act, const float* mask, int ldm, float* C, int ldc) {
  static_assert(R == 8, "the tensor-core path maps the batch tile onto the n = 8 side of m16n8k8");
  (void)red;
  const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const sptr sA = sp_of(A), sB = sp_of(Bs), sC = sp_of(C);
  const bool has_bias = bias != nullptr;
  const sptr sBias = sp_of(has_bias ? bias : Bs), sM = sp_of(mask ? mask : Bs);
  const int mtiles = (N_pad + 15) >> 4;
  trace(50);
  for (int mt = warp; mt < mtiles; mt += FRL_NT / 32) {
    float v4[4];
    mma_tile(sB, 1, ldb, sA, lda, N_pad, K_pad, mt, g, t, v4);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int o = mt * 16 + g + ((q & 2) ? 8 : 0), r = 2 * t + (q & 1);
      if (o < N_pad) {
        float v = v4[q];
        if (epi == EPI_BIAS_ACT) v = apply_act(v + (has_bias ? sp_ld1(sBias, o) : 0.f), act);
        else v = (sp_ld1(sM, r * ldm + o) > 0.f) ? v : 0.f;
        sp_st1(sC, r * ldc + o, v);
      }
    }
  }
  trace(51);
  FRL_SYNC();
  trace(53);
}

// C[r][k] = epi( sum_n A[r][n] * Bs[k][n] )   (backward dX on the forward image; same contract as the FFMA gemm_nt)
template <int R, int HM = 0>
FRL_NI_GEMM void gemm_nt(float* red, const float* A, int lda, int N_red, const float* Bs, int ldb, int K_out, int epi,
                         const float* mask, int ldm, float* C, int ldc) {
  static_assert(R == 8, "tensor-core path: R == 8");
  (void)red;
  const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const sptr sA = sp_of(A), sB = sp_of(Bs), sC = sp_of(C), sM = sp_of(mask ? mask : Bs);
  const int mtiles = (K_out + 15) >> 4;
  trace(60);
  for (int mt = warp; mt < mtiles; mt += FRL_NT / 32) {
    float v4[4];
    mma_tile(sB, ldb, 1, sA, lda, K_out, N_red, mt, g, t, v4);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = mt * 16 + g + ((q & 2) ? 8 : 0), r = 2 * t + (q & 1);
      if (i < K_out) {
        float v = v4[q];
        if (epi == EPI_RELU_MASK) {
          if (HM) { const float m = sp_ld1(sM, r * ldm + i); v = v * (1.f - m * m); }
          else v = (sp_ld1(sM, r * ldm + i) > 0.f) ? v : 0.f;
        }
        sp_st1(sC, r * ldc + i, v);
      }
    }
  }
  trace(61);
  FRL_SYNC();
  trace(63);
}

// G[m][n] (+)= sum_{r<8} dY[r][m] * X[r][n]  to GLOBAL (one k-step per 16x8 tile), bias gradient gb[m] (+)= sum_r dY[r][m]
template <int R>
FRL_NI_GEMM void gemm_outer(const float* dY, int ldy, int M_pad, const float* X, int ldx, int N_pad, int N_real,
                            float* G, float* gb, bool accumulate) {
  static_assert(R == 8, "tensor-core path: R == 8");
  const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const sptr sY = sp_of(dY), sX = sp_of(X);
  const int mtiles = (M_pad + 15) >> 4, ntiles = (N_pad + 7) >> 3;
  // a warp owns an (m-tile, n-tile parity) strip so that narrow layers (one or two m-tiles) still use all 8 warps
  const int nstrips = mtiles >= FRL_NT / 32 ? 1 : ((FRL_NT / 32) / mtiles);
  trace(70);
  for (int u = warp; u < mtiles * nstrips; u += FRL_NT / 32) {
    const int mt = u / nstrips, strip = u - mt * nstrips;
    const int m0 = mt * 16 + g, m1 = m0 + 8;
    const int m0c = m0 < M_pad ? m0 : M_pad - 1, m1c = m1 < M_pad ? m1 : M_pad - 1;
    uint32_t ah[4], al[4];
    tf32_split(lds_op(sY + 4u * (uint32_t)(t * ldy + m0c)), ah[0], al[0]);
    tf32_split(lds_op(sY + 4u * (uint32_t)(t * ldy + m1c)), ah[1], al[1]);
    tf32_split(lds_op(sY + 4u * (uint32_t)((t + 4) * ldy + m0c)), ah[2], al[2]);
    tf32_split(lds_op(sY + 4u * (uint32_t)((t + 4) * ldy + m1c)), ah[3], al[3]);
    float xn0, xn1;
    {
      const int n = strip * 8 + g, nc = n < N_pad ? n : N_pad - 1;
      xn0 = lds_op(sX + 4u * (uint32_t)(t * ldx + nc)); xn1 = lds_op(sX + 4u * (uint32_t)((t + 4) * ldx + nc));
    }
#pragma unroll 2
    for (int nt = strip; nt < ntiles; nt += nstrips) {
      const float x0 = xn0, x1 = xn1;
      if (nt + nstrips < ntiles) {            // next tile's fragment before this tile's math
        const int n = (nt + nstrips) * 8 + g, nc = n < N_pad ? n : N_pad - 1;
        xn0 = lds_op(sX + 4u * (uint32_t)(t * ldx + nc)); xn1 = lds_op(sX + 4u * (uint32_t)((t + 4) * ldx + nc));
      }
      uint32_t bh[2], bl[2];
      tf32_split(x0, bh[0], bl[0]);
      tf32_split(x1, bh[1], bl[1]);
      float hh[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
      mma_tf32(hh, ah, bh);
      mma_tf32(lo, ah, bl);
      mma_tf32(lo, al, bh);
      const int c0 = nt * 8 + 2 * t;            // columns c0, c0 + 1 of rows m0 (hh[0..1]) and m1 (hh[2..3])
      if (c0 < N_pad) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = h ? m1 : m0;
          if (m < M_pad) {
            float2 v = make_float2(c0 < N_real ? hh[2 * h] + lo[2 * h] : 0.f, c0 + 1 < N_real ? hh[2 * h + 1] + lo[2 * h + 1] : 0.f);
            float2* dst = reinterpret_cast<float2*>(G + (size_t)m * N_pad + c0);
            if (accumulate) { const float2 o = *dst; v.x += o.x; v.y += o.y; }
            *dst = v;
          }
        }
      }
    }
  }
  // bias gradient (exact fp32 row sums)
  for (int m = (int)threadIdx.x; m < M_pad; m += FRL_NT) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) s += sp_ld1(sY, r * ldy + m);
    if (accumulate) s += gb[m];
    gb[m] = s;
  }
  trace(71);
  FRL_SYNC();
  trace(72);
}
#endif   // FRL_MMA

// ------------------------------------------------------------------------------------------------
// layer / MLP passes
// ------------------------------------------------------------------------------------------------
FRL_HD int layer_fwd_bytes(const frl_layer_t& L) { return wt_floats(L) * 4; }
FRL_DEV const float* layer_fwd_src(const frl_net_t& n, int li) { return n.pt + n.L[li].wt_off; }

struct Hint { const float* ptr; int bytes; };
FRL_DEV Hint no_hint() { Hint h; h.ptr = nullptr; h.bytes = 0; return h; }
FRL_DEV Hint fwd_hint(const frl_net_t& n, int li) { Hint h; h.ptr = layer_fwd_src(n, li); h.bytes = layer_fwd_bytes(n.L[li]); return h; }
FRL_DEV Hint bwd_hint(const frl_net_t& n, int li) { return fwd_hint(n, li); }   // backward reads the same (forward) image

// ------------------------------------------------------------------------------------------------
// Resident slots: wbuf0 / wbuf1 each sized for a whole 3-layer head (all layers of a head are contiguous in `pt`).
// A slot is tagged with the global source it holds; res_fetch is a no-op when the head is already resident, so a stage
// can be written as "I need head H in slot s" and weights survive across stages / learns until the optimiser stage
// that rewrites them calls res_invalidate.  Every layer has its own mbarrier: layer 0 can be consumed while layers 1, 2
// are still in flight, and a fetch issued in one stage may complete (unawaited) during later ones.
// ------------------------------------------------------------------------------------------------
FRL_DEV const float* res_key(const frl_net_t& n, int l0) { return n.pt + n.L[l0].wt_off; }
FRL_DEV int res_find(const Cta& c, const frl_net_t& n, int l0) {
  const float* key = res_key(n, l0);
  return c.stag0 == key ? 0 : (c.stag1 == key ? 1 : -1);
}
FRL_DEV void res_drain_slot(Cta& c, int s) {
  bool any = false;
  for (int k = 0; k < 3; ++k) {
    const uint32_t bit = 1u << (s * 3 + k);
    if (c.spend & bit) {
#ifndef FRL_EMUL
      mbar_wait(c.bar + 2 + s * 3 + k, (c.sphase >> (s * 3 + k)) & 1u);
#endif
      c.sphase ^= bit; c.spend &= ~bit; any = true;
    }
  }
  if (any) FRL_SYNC();      // nobody may still be polling the old phase when the barrier is re-armed
}
// thread 0: one bulk copy + mbarrier per layer (kept out of line: eight call sites)
FRL_NI_MISC void res_issue(float* dst, uint64_t* bars, const frl_net_t& n, int l0, int nl) {
#ifndef FRL_EMUL
  trace(20);
  fence_proxy_async();
  trace(21);
  for (int k = 0; k < nl; ++k) {
    const frl_layer_t& L = n.L[l0 + k];
    const uint32_t bytes = (uint32_t)wt_floats(L) * 4u;
    mbar_expect_tx(bars + k, bytes);
    trace(22);
    tma_bulk_g2s(dst + (L.wt_off - n.L[l0].wt_off), n.pt + L.wt_off, bytes, bars + k);
    trace(23);
  }
#else
  (void)bars;
  for (int k = 0; k < nl; ++k) {
    const frl_layer_t& L = n.L[l0 + k];
    memcpy(dst + (L.wt_off - n.L[l0].wt_off), n.pt + L.wt_off, (size_t)wt_floats(L) * 4);
  }
#endif
}
// Start fetching layers [l0, l0+nl) of net n into slot s unless that head is already there.
// Precondition: every thread is past its last read of slot s (all engine ops end with FRL_SYNC).
FRL_DEV void res_fetch(Cta& c, int s, const frl_net_t& n, int l0, int nl) {
  const float* key = res_key(n, l0);
  if ((s ? c.stag1 : c.stag0) == key) return;
  res_drain_slot(c, s);
  if (s) c.stag1 = key; else c.stag0 = key;
#ifndef FRL_EMUL
  if (threadIdx.x == 0)
#endif
    res_issue(s ? c.wbuf1 : c.wbuf0, c.bar + 2 + s * 3, n, l0, nl);
  c.spend |= ((1u << nl) - 1u) << (s * 3);
}
// smem image of layer l0+k of the head in slot s (waits for its copy if still outstanding)
FRL_DEV const float* res_layer(Cta& c, int s, const frl_net_t& n, int l0, int k) {
  const uint32_t bit = 1u << (s * 3 + k);
  if (c.spend & bit) {
#ifndef FRL_EMUL
    mbar_wait(c.bar + 2 + s * 3 + k, (c.sphase >> (s * 3 + k)) & 1u);
#endif
    c.sphase ^= bit; c.spend &= ~bit;
  }
  return (s ? c.wbuf1 : c.wbuf0) + (n.L[l0 + k].wt_off - n.L[l0].wt_off);
}
// the parameters of n were rewritten: forget resident copies (outstanding copies are drained by the next res_fetch)
FRL_DEV void res_invalidate(Cta& c, const frl_net_t& n) {
  if (c.stag0 >= n.pt && c.stag0 < n.pt + n.n_pt) c.stag0 = nullptr;
  if (c.stag1 >= n.pt && c.stag1 < n.pt + n.n_pt) c.stag1 = nullptr;
}
FRL_DEV void res_drain(Cta& c) { res_drain_slot(c, 0); res_drain_slot(c, 1); }

// Layer image for an op: resident slot `slot` (>= 0) or, with slot < 0, the streaming double buffer (fetch now unless
// prefetched, then start prefetching `next`).
FRL_DEV const float* wt_acquire(Cta& c, int slot, const frl_net_t& n, int l0, int k, Hint next) {
  if (slot >= 0) return res_layer(c, slot, n, l0, k);
  const float* Bs = stage_acquire(c, layer_fwd_src(n, l0 + k), layer_fwd_bytes(n.L[l0 + k]));
  stage_prefetch(c, next.ptr, next.bytes);
  return Bs;
}

// Y = act(X W^T + b) on a staged layer image
template <int R, int HM = 0>
FRL_DEV void layer_fwd_img(Cta& c, const frl_layer_t& L, const float* Bs, const float* X, int ldx, float* Y, int ldy, int act) {
  gemm_rk<R, HM>(c.red, X, ldx, L.in_pad, Bs, wt_ld(L), L.out_pad, Bs + wt_bias(L), EPI_BIAS_ACT, act, nullptr, 0, Y, ldy);
}
// dX = (dY W) * relu'(mask)   (mask == nullptr: no activation derivative) on the same forward image
template <int R, int HM = 0>
FRL_DEV void layer_bwd_img(Cta& c, const frl_layer_t& L, const float* Bs, const float* dY, int ldy, const float* mask, int ldm,
                           float* dX, int ldx) {
  gemm_nt<R, HM>(c.red, dY, ldy, L.out_pad, Bs, wt_ld(L), L.in_pad, mask ? EPI_RELU_MASK : EPI_BIAS_ACT, mask, ldm, dX, ldx);
}

// Streaming flavours.  `next` is what the caller will need after this layer (prefetched during the math).
template <int R>
FRL_DEV void layer_fwd(Cta& c, const frl_net_t& n, int li, const float* X, int ldx, float* Y, int ldy, int act, Hint next) {
  layer_fwd_img<R>(c, n.L[li], wt_acquire(c, -1, n, li, 0, next), X, ldx, Y, ldy, act);
}
template <int R>
FRL_DEV void layer_bwd_dx(Cta& c, const frl_net_t& n, int li, const float* dY, int ldy, const float* mask, int ldm,
                          float* dX, int ldx, Hint next) {
  layer_bwd_img<R>(c, n.L[li], wt_acquire(c, -1, n, li, 0, next), dY, ldy, mask, ldm, dX, ldx);
}

// MLP forward over layers [l0, l0+nl): hidden layers ReLU (HM = 1: tanh), last layer `act_out`.
//   nl == 3: H1 = relu(l0 X), H2 = relu(l1 H1), OUT = act(l2 H2);   nl == 2: H1 = relu(l0 X), OUT = act(l1 H1).
//   slot >= 0: the head is resident in that slot (res_fetch was issued by the caller); slot < 0: streaming.
template <int R, int HM = 0>
FRL_NI_MLP void mlp_fwd(Cta& c, const frl_net_t& n, int l0, int nl, const float* X, int ldx, float* H1, float* H2, int ldh,
                     float* OUT, int ldo, int act_out, Hint next, int slot = -1) {
  if (nl == 3) {
    layer_fwd_img<R, HM>(c, n.L[l0], wt_acquire(c, slot, n, l0, 0, fwd_hint(n, l0 + 1)), X, ldx, H1, ldh, HM ? FRL_ACT_TANH : FRL_ACT_RELU);
    layer_fwd_img<R, HM>(c, n.L[l0 + 1], wt_acquire(c, slot, n, l0, 1, fwd_hint(n, l0 + 2)), H1, ldh, H2, ldh, HM ? FRL_ACT_TANH : FRL_ACT_RELU);
    layer_fwd_img<R>(c, n.L[l0 + 2], wt_acquire(c, slot, n, l0, 2, next), H2, ldh, OUT, ldo, act_out);
  } else {
    layer_fwd_img<R, HM>(c, n.L[l0], wt_acquire(c, slot, n, l0, 0, fwd_hint(n, l0 + 1)), X, ldx, H1, ldh, HM ? FRL_ACT_TANH : FRL_ACT_RELU);
    layer_fwd_img<R>(c, n.L[l0 + 1], wt_acquire(c, slot, n, l0, 1, next), H1, ldh, OUT, ldo, act_out);
  }
}

// Backward of mlp_fwd given dOUT (gradient w.r.t. the last layer's pre-activation output).
//   gp  : this CTA's gradient partial for net n (n_p floats, same layout as n.p) or nullptr (no dW wanted)
//   dXo : if non-null receives dL/dX [R][in_pad of layer l0]
//   D1/D2: scratch [R][ldh] for the hidden-layer gradients.
template <int R, int HM = 0>
FRL_NI_MLP void mlp_bwd(Cta& c, const frl_net_t& n, int l0, int nl, const float* X, int ldx, const float* H1, const float* H2,
                     int ldh, const float* dOUT, int ldo, float* D1, float* D2, float* dXo, int lddx, float* gp,
                     bool accumulate, Hint next, int slot = -1) {
  const float* dcur = dOUT;
  int ldc = ldo;
  for (int k = nl - 1; k >= 0; --k) {
    const int li = l0 + k;
    const frl_layer_t& L = n.L[li];
    const float* Xin = (k == 0) ? X : (k == 1 ? H1 : H2);
    const int ldin = (k == 0) ? ldx : ldh;
    if (gp) gemm_outer<R>(dcur, ldc, L.out_pad, Xin, ldin, L.in_pad, L.in, gp + L.w_off, gp + L.b_off, accumulate);
    if (k > 0) {
      float* dn = (k == 2) ? D2 : D1;   // gradient wrt H2 (k==2) or H1 (k==1)
      Hint h = (k - 1 > 0 || dXo) ? bwd_hint(n, li - 1) : next;
      layer_bwd_img<R, HM>(c, L, wt_acquire(c, slot, n, l0, k, h), dcur, ldc, Xin, ldin, dn, ldh);
      dcur = dn; ldc = ldh;
    } else if (dXo) {
      layer_bwd_img<R>(c, L, wt_acquire(c, slot, n, l0, k, next), dcur, ldc, nullptr, 0, dXo, lddx);
    }
  }
}

// fixed-order sum of n values at src[i*stride], 16 independent loads in flight (the naive loop serialises one L2 round
// trip per element: ~10 us for 32 partials when a single thread does it while its CTA waits at the next barrier)
FRL_DEV float strided_sum(const float* src, int stride, int n) {
  float tot = 0.f;
  int i = 0;
  for (; i + 16 <= n; i += 16) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = src[(size_t)(i + k) * stride];
#pragma unroll
    for (int k = 0; k < 16; ++k) tot += v[k];
  }
  for (; i < n; ++i) tot += src[(size_t)i * stride];
  return tot;
}

// Sums of up to three strided series of n <= FRL_NT values each, loaded by n threads IN PARALLEL (one L2 round trip; a
// single thread walking 64 partials costs ~0.7 us per dependent round trip) and added in fixed order.  sh: >= 3n+4 floats.
FRL_DEV void cta_sums(float* sh, const float* p0, int s0, const float* p1, int s1, const float* p2, int s2, int n, float out[3]) {
  FRL_PAR(t) {
    if (t < n) {
      sh[t] = p0[(size_t)t * s0];
      sh[n + t] = p1 ? p1[(size_t)t * s1] : 0.f;
      sh[2 * n + t] = p2 ? p2[(size_t)t * s2] : 0.f;
    }
  }
  FRL_SYNC();
  FRL_PAR(t) {
    if (t < 3) {
      float acc = 0.f;
      for (int i = 0; i < n; ++i) acc += sh[t * n + i];
      sh[3 * n + t] = acc;
    }
  }
  FRL_SYNC();
  out[0] = sh[3 * n]; out[1] = sh[3 * n + 1]; out[2] = sh[3 * n + 2];
}

// ------------------------------------------------------------------------------------------------
// block-wide fixed-order sum of one float per thread (result broadcast to all threads via smem)
// ------------------------------------------------------------------------------------------------
// Fixed association: a shuffle tree inside each warp (lane i += lane i+off, off = 16..1), then the warp totals in warp
// order.  3 barriers instead of the 9 of a shared-memory tree; the emulation applies the same tree to the array.
FRL_NI_MISC float block_sum(float* slot /*[FRL_NT] smem*/) {
#ifndef FRL_EMUL
  float v = slot[threadIdx.x];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < FRL_NT / 32; ++w) tot += slot[w];
  __syncthreads();
  return tot;
#else
  float tot = 0.f;
  for (int w = 0; w < FRL_NT / 32; ++w) {
    for (int off = 16; off > 0; off >>= 1)
      for (int i = 0; i < off; ++i) slot[w * 32 + i] += slot[w * 32 + i + off];
    tot += slot[w * 32];
  }
  return tot;
#endif
}

// ------------------------------------------------------------------------------------------------
// cross-CTA gradient reduction (fixed order) + sum of squares partial per CTA
//   gpart: [ncontrib][stride];  n.g <- sum_c gpart[c];  sumsq_part[cta] <- sum over this CTA's slice of g^2
// ------------------------------------------------------------------------------------------------
FRL_NI_OPT void reduce_grads(int cta, int ncta, float* slot, const frl_net_t& n, const float* gpart, int stride, int ncontrib,
                            float* sumsq_part) {
  FRL_PAR(t) {
    float local = 0.f;
    // n_p and stride are multiples of 4: each thread owns groups of 4 consecutive parameters (16-B loads)
    for (int p = (cta * FRL_NT + t) * 4; p < n.n_p; p += ncta * FRL_NT * 4) {
      float4 s = ld4(gpart + p);
      int cc = 1;
      for (; cc + 8 <= ncontrib; cc += 8) {     // 8 independent loads in flight, summed in fixed order
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ld4(gpart + (size_t)(cc + i) * stride + p);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = f4add(s, v[i]);
      }
      for (; cc < ncontrib; ++cc) s = f4add(s, ld4(gpart + (size_t)cc * stride + p));
      st4(n.g + p, s);
      local += s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;
    }
    slot[t] = local;
  }
  FRL_SYNC();
  float tot = block_sum(slot);
  FRL_PAR(t) { if (t == 0 && sumsq_part) sumsq_part[cta] = tot; }
  FRL_SYNC();
}

struct AdamHP {
  float lr_over_bc1_neg;   // -lr / (1 - b1^step)          (torch: value = -step_size)
  float bc2_sqrt;          // sqrt(1 - b2^step)
  float one_minus_b1;      // float(1 - b1)
  float b2, one_minus_b2;
  float eps;
  float weight_decay;      // L2 (added to grad), 0 = off
  float max_norm;          // clip_grad_norm_ max (<= 0: no clipping)
};

FRL_HD AdamHP make_adam_hp(double lr, double b1, double b2, double eps, double wd, double max_norm, long step) {
  AdamHP h;
  // b^step by squaring (two independent ~20-deep DMUL chains): pow() / expm1(log()) on the single thread that computes
  // this cost ~0.7 us per optimiser stage.  The few-ulp double error disappears in the float conversions below.
  double p1 = 1.0, p2 = 1.0, q1 = b1, q2 = b2;
  for (unsigned long e = (unsigned long)step; e; e >>= 1) {
    if (e & 1) { p1 *= q1; p2 *= q2; }
    q1 *= q1; q2 *= q2;
  }
  double bc1 = 1.0 - p1;
  double bc2 = 1.0 - p2;
  h.lr_over_bc1_neg = (float)(-(lr / bc1));
  h.bc2_sqrt = (float)sqrt(bc2);
  h.one_minus_b1 = (float)(1.0 - b1);
  h.b2 = (float)b2;
  h.one_minus_b2 = (float)(1.0 - b2);
  h.eps = (float)eps;
  h.weight_decay = (float)wd;
  h.max_norm = (float)max_norm;
  return h;
}

FRL_NOINL AdamHP adam_hp_ni(double lr, double b1, double b2, double eps, double wd, double max_norm, long step) {
  return make_adam_hp(lr, b1, b2, eps, wd, max_norm, step);
}
FRL_NOINL float randn_ni(uint64_t seed, uint32_t stream, uint32_t ctr, uint32_t idx) { return frl_randn(seed, stream, ctr, idx); }

// where does parameter index p live in the transposed mirror?  (-1: no mirror, e.g. log_std)
FRL_DEV int mirror_index(const frl_net_t& n, int p) {
  for (int li = 0; li < n.n_layers; ++li) {
    const frl_layer_t& L = n.L[li];
    const int wsz = L.out_pad * L.in_pad;
    if (p >= L.w_off && p < L.w_off + wsz) {
      const int e = p - L.w_off, j = e / L.in_pad, k = e % L.in_pad;
      return L.wt_off + k * wt_ld(L) + j;
    }
    if (p >= L.b_off && p < L.b_off + L.out_pad) return L.wt_off + wt_bias(L) + (p - L.b_off);
  }
  return -1;
}

// distance in the mirror between parameter p and p+1 (same tensor): W row (k, k+1) -> wt_ld, bias -> 1
FRL_DEV int mirror_stride(const frl_net_t& n, int p) {
  for (int li = 0; li < n.n_layers; ++li) {
    const frl_layer_t& L = n.L[li];
    if (p >= L.w_off && p < L.w_off + L.out_pad * L.in_pad) return wt_ld(L);
  }
  return 1;
}

// torch.optim.Adam single-tensor math on this CTA's slice of the parameters, preceded by the global-norm
// clip (coef from the per-CTA sum-of-squares partials written by reduce_grads), optionally followed by the
// Polyak update of a target net with identical layout.  Keeps p / pt (and target p / pt) in sync.
struct AdamSpec { double lr, b1, b2, eps, wd, max_norm; long step; };

FRL_NI_OPT void adam_update(int cta, int ncta, float* sh, const frl_net_t& n, const float* sumsq_part, int nparts,
                           const AdamSpec sp, const frl_net_t* tgt, float tau) {
  // (1) per-CTA scalars: bias corrections (double pow) by one thread, norm partials summed in fixed order from smem
  FRL_PAR(t) {
    if (sumsq_part) { for (int i = t; i < nparts; i += FRL_NT) sh[64 + i] = sumsq_part[i]; }
    if (t == 0) {
      const AdamHP h = adam_hp_ni(sp.lr, sp.b1, sp.b2, sp.eps, sp.wd, sp.max_norm, sp.step);
      sh[0] = h.lr_over_bc1_neg; sh[1] = h.bc2_sqrt; sh[2] = h.one_minus_b1; sh[3] = h.b2; sh[4] = h.one_minus_b2;
      sh[5] = h.eps; sh[6] = h.weight_decay; sh[7] = h.max_norm;
    }
  }
  FRL_SYNC();
  FRL_PAR(t) {
    if (t == 0) {
      float coef = 1.f;
      if (sh[7] > 0.f && sumsq_part) {
        float tot = 0.f;
        for (int i = 0; i < nparts; ++i) tot += sh[64 + i];
        coef = sh[7] / (sqrtf(tot) + 1e-6f);
        if (coef > 1.f) coef = 1.f;
      }
      sh[8] = coef;
    }
  }
  FRL_SYNC();
  AdamHP hp;
  hp.lr_over_bc1_neg = sh[0]; hp.bc2_sqrt = sh[1]; hp.one_minus_b1 = sh[2]; hp.b2 = sh[3]; hp.one_minus_b2 = sh[4];
  hp.eps = sh[5]; hp.weight_decay = sh[6]; hp.max_norm = sh[7];
  const float coef = sh[8];
  const float omt = (float)(1.0 - (double)tau);
  FRL_PAR(t) {
    // 4 consecutive parameters per thread (16-B loads/stores of p, m, v, g and the target); mirrors are scattered
    for (int p4 = (cta * FRL_NT + t) * 4; p4 < n.n_p; p4 += ncta * FRL_NT * 4) {
      const float4 g4 = ld4(n.g + p4), w4 = ld4(n.p + p4), m4 = ld4(n.m + p4), v4 = ld4(n.v + p4);
      float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tgt) t4 = ld4(tgt->p + p4);
      float gg[4] = {g4.x, g4.y, g4.z, g4.w}, ww[4] = {w4.x, w4.y, w4.z, w4.w};
      float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float g = gg[i] * coef;
        float w = ww[i];
        if (hp.weight_decay != 0.f) g = fmaf(w, hp.weight_decay, g);
        float m = mm[i], v = vv[i];
        m = fmaf(hp.one_minus_b1, g - m, m);                           // exp_avg.lerp_(grad, 1-b1)
        v = fadd(fmul(v, hp.b2), fmul(fmul(hp.one_minus_b2, g), g));   // mul_(b2).addcmul_(g, g, 1-b2)
        const float denom = fadd(fdiv(fsqrt(v), hp.bc2_sqrt), hp.eps);
        w = fadd(w, fdiv(fmul(hp.lr_over_bc1_neg, m), denom));         // addcdiv_(m, denom, value=-step_size)
        mm[i] = m; vv[i] = v; ww[i] = w;
        if (tgt) tt[i] = fadd(fmul(tt[i], omt), fmul(w, tau));
      }
      st4(n.m + p4, make_float4(mm[0], mm[1], mm[2], mm[3]));
      st4(n.v + p4, make_float4(vv[0], vv[1], vv[2], vv[3]));
      st4(n.p + p4, make_float4(ww[0], ww[1], ww[2], ww[3]));
      if (tgt) st4(tgt->p + p4, make_float4(tt[0], tt[1], tt[2], tt[3]));
      const int mi0 = mirror_index(n, p4);    // a group of 4 never straddles tensors (all offsets are multiples of 4)
      if (mi0 >= 0) {
        const int stride = mirror_stride(n, p4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          n.pt[mi0 + i * stride] = ww[i];
          if (tgt) tgt->pt[mi0 + i * stride] = tt[i];
        }
      }
    }
  }
  FRL_SYNC();
}

FRL_DEV void polyak_update(Cta& c, const frl_net_t& src, const frl_net_t& tgt, float tau) {
  const float omt = (float)(1.0 - (double)tau);
  FRL_PAR(t) {
    for (int p = c.cta * FRL_NT + t; p < src.n_p; p += c.ncta * FRL_NT) {
      float tw = fadd(fmul(tgt.p[p], omt), fmul(src.p[p], tau));
      tgt.p[p] = tw;
      const int mi = mirror_index(src, p);
      if (mi >= 0) tgt.pt[mi] = tw;
    }
  }
  FRL_SYNC();
}

// zero-fill a smem tile [R][ld]
FRL_DEV void tile_zero(float* T, int n) {
  FRL_PAR(t) { for (int i = t; i < n; i += FRL_NT) T[i] = 0.f; }
  FRL_SYNC();
}
