// tcgen05 / TMEM building blocks of the large-minibatch learn path (sm_100a only; inline PTX, no CUTLASS).
//
// Numerics: every GEMM is 3xTF32 — an fp32 operand x is split into hi = x with the low 13 mantissa bits cleared (exactly a
// TF32 number) and lo = x - hi (exact in fp32; the tensor core truncates it to TF32 again), and
//     D += A_lo * B_hi;  D += A_hi * B_lo;  D += A_hi * B_hi        (fp32 accumulators in TMEM)
// which leaves a relative error of ~2^-21 per product (measured on B200 with tools/umma_test.cu: 1e-6 of max |D| at K = 64,
// against 7e-4 for plain TF32) — the budget of the 1e-5 loss tolerance.
//
// Shared / global operand layouts (both measured with the address probe of tools/umma_test.cu, profiles/r2j_umma_probe.log):
//   layout Q  (K-major operand, no swizzle)   [R rows (M or N index) x C cols (K index)], C % 8 == 0, R % 8 == 0
//       float offset(r, c) = (c / 4) * 4R + r * 4 + (c % 4)                 -> core matrix = 8 rows x 16 B, contiguous 128 B
//       descriptor: layout_type 0, SBO = 128 B (next 8 rows), LBO = 16R B (next 4 columns); one K = 8 MMA step = +32R B
//       a K-chunk of 32 columns is one contiguous 128R-byte block (bulk-copyable)
//   layout S  (MN-major operand, "128B swizzle with 32B atomicity", the only MN-major layout the tensor core reads for TF32)
//       [K rows (reduction index) x C cols (M or N index)], C % 32 == 0, rows % 8 == 0
//       float offset(k, n) = (k / 4) * 4C + (n / 32) * 128 + (k % 4) * 32 + ((((n % 32) / 8) ^ (k % 4)) * 8) + (n % 8)
//       descriptor: layout_type 1, LBO = 512 B (next 32 columns), SBO = 16C B (next 4 rows); one K = 8 MMA step = +32C B
//       a chunk of 32 rows is one contiguous 128C-byte block; bases must be 512-B aligned
// TMEM: lane = row of the 128-row tile, column = 32-bit element; a warp reads / writes lanes 32 * (warp % 4) .. + 31.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define UM_DEV __device__ __forceinline__

__host__ __device__ __forceinline__ int um_q_off(int r, int c, int R) { return (c >> 2) * (4 * R) + r * 4 + (c & 3); }
__host__ __device__ __forceinline__ int um_s_off(int k, int n, int C) {
  return (k >> 2) * (4 * C) + (n >> 5) * 128 + (k & 3) * 32 + (((((n & 31) >> 3) ^ (k & 3))) << 3) + (n & 7);
}
__host__ __device__ __forceinline__ float um_hi(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
#else
  union { float f; uint32_t u; } v; v.f = x; v.u &= 0xffffe000u; return v.f;
#endif
}

UM_DEV uint32_t um_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

UM_DEV uint64_t um_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;                     // descriptor version of sm_100
  d |= (uint64_t)layout_type << 61;
  return d;
}
UM_DEV uint64_t um_desc_q(uint32_t saddr, int R) { return um_desc(saddr, 16u * (uint32_t)R, 128u, 0u); }   // layout Q, K-major
UM_DEV uint64_t um_desc_s(uint32_t saddr, int C) { return um_desc(saddr, 512u, 16u * (uint32_t)C, 1u); }   // layout S, MN-major

// kind::tf32, fp32 accumulate; M in {64, 128}, N % 16 == 0 (M = 128), a_mn / b_mn: 1 = MN-major operand
UM_DEV uint32_t um_idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

UM_DEV void um_mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// A operand from TMEM (lane = row, column = K index), B from shared memory
UM_DEV void um_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// mbarrier arrive once every MMA issued so far by this thread has completed (implies tcgen05.fence::before_thread_sync)
UM_DEV void um_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(um_smem_u32(bar)) : "memory");
}
UM_DEV void um_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
UM_DEV void um_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
UM_DEV void um_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
UM_DEV void um_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one warp: allocate ncols (power of two >= 32) TMEM columns, base address -> *slot (shared memory).  RELINQUISH = true gives up
// the CTA's right to allocate again (lets another CTA of the SM allocate); a persistent kernel that allocates once per update
// (one CTA per SM) must keep the permit.
template <int NCOLS, bool RELINQUISH = false>
UM_DEV void um_tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem_u32(slot)), "n"(NCOLS) : "memory");
  if (RELINQUISH) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
UM_DEV void um_tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// TMEM <-> registers, 32 lanes x 32 bit, 16 consecutive columns (the caller adds the warp's lane base << 16 to taddr)
UM_DEV void um_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  um_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
UM_DEV void um_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
      "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])),
      "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
      "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}

UM_DEV void um_mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(um_smem_u32(bar)), "r"(count)); }
UM_DEV void um_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(um_smem_u32(bar)), "r"(bytes) : "memory");
}
UM_DEV void um_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(um_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
UM_DEV void um_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(um_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(um_smem_u32(bar))
               : "memory");
}
UM_DEV void um_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
UM_DEV void um_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
