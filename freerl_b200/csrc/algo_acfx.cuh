// Small-batch fused actor-critic learn() — the latency-optimised schedule for the reference's own batch sizes (B <= 256).
//   SAC   SAC_file/SAC.py:222-271      TD3   TD3_file/TD3.py:189-244      DDPG   DDPG_file/DDPG.py:203-233
// Same arithmetic as algo_ac.cuh (single agent, hidden 128-128, no Batch_ObsNorm); what changes is WHERE the work runs.  The
// r1 trace of the generic kernel (profiles/r2a_trace_generic.txt) showed a learn as a chain of 28 layer ops of ~3.4 k clk each
// — 1.9 k in the FMA loop, 0.9 k in the shared-memory K-split pass, 0.6 k of scalar index math between ops — plus 12 k clk of
// per-tile outer products, 2 x 11 k clk of reduce / Adam stages and 7 cooperative-groups grid barriers of 2.5 k clk.  Here:
//   * CTA sets.  For every 8-row tile t and critic head h there is a TARGET CTA T(t,h) and an ONLINE CTA C(t,h); weights stay
//     resident in shared memory (T: actor_target + target head h; C: actor + online head h) and are re-fetched by TMA only
//     after the optimiser stage that rewrites them.  T and C run concurrently in phase A: T computes a' and Q'_h(s',a') and
//     publishes them with a release flag, C meanwhile runs the online critic forward on (s,a) and the actor forward pi(s),
//     then picks up the targets (acquire spin), forms y and backpropagates.  The chain per learn drops from 28 ops to 16.
//   * dW out of the chain.  Worker CTAs only write the layer inputs X_l and pre-activation gradients dY_l of their tile to a
//     blocked exchange buffer in L2; the weight gradient dW_l = dY_l^T X_l over ALL rows is one batched GEMM stage whose
//     16x16 / 16x32 output tiles are spread over all 148 SMs (operands arrive as contiguous TMA bulk copies).  The CTA that
//     produced a tile keeps it in shared memory across the norm barrier and applies clip + Adam + Polyak + mirror refresh to
//     exactly those elements: no per-CTA partial gradients, no separate cross-CTA reduce stage, fixed summation order.
//   * GEMM microkernels with static 128-wide shapes: each warp owns 16 output columns, lanes split K 8 ways and combine with a
//     3-step shuffle reduce-scatter (no shared-memory K-split pass, one barrier per op).
//   * Hand-rolled grid barrier (red.release + ld.acquire spin on one L2 word) instead of cooperative-groups grid.sync.
// Stages per learn:  0 phase A (targets || online critic + actor forward, critic backward)   1 dW(critic) + sum of squares
//   2 clip + Adam(critic) + Polyak   3 phase C (Q(s,pi(s)), dQ/da, actor backward)   4 dW(actor)   5 Adam(actor) + Polyak + alpha
#pragma once
#include "algo_ac.cuh"
#ifdef FRL_EMUL
#include <stdlib.h>
#include <stdio.h>
#endif

#define FX_LDW 132            // wt_ld_of(128): row stride of a 128-wide layer image
#define FX_MAXB 256           // exchange operands of one dW job must fit one weight slot
#define FX_SPIN_LIMIT (1u << 24)
#define FX_SYNC_WORDS 4096    // uint32 words of frl_ac_args_t.sync: [0] grid-barrier counter, [64, 2048) 3 x 64-bit hand-off packets per
                              // batch row, [2048, 4096) one 64-bit gradient-norm packet per CTA for the critic and the actor

// ------------------------------------------------------------------------------------------------------------------------
// cross-CTA signalling (GPU: release / acquire on L2 words; emulation: CTAs of a stage run in index order)
// ------------------------------------------------------------------------------------------------------------------------
#ifndef FRL_EMUL
FRL_DEV void fx_flag_set(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
FRL_DEV unsigned fx_flag_ld(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
FRL_DEV void fx_flag_wait(const unsigned* p, unsigned v) {
  unsigned spins = 0;
  while ((int)(fx_flag_ld(p) - v) < 0) {
    if (++spins > FX_SPIN_LIMIT) __trap();        // never hang the device on a protocol bug
  }
}
FRL_DEV float fx_ldcg(const float* p) { return __ldcg(p); }
// Hand-off packets: one 64-bit word = (epoch << 32 | float bits).  An aligned 8-byte store is single-copy atomic, so a reader that
// sees the epoch it waits for has the value of THAT store — no release / acquire pair, no separate flag word, and a consumer can
// poll all the packets it needs with independent loads in flight (one L2 round trip when the producer is already done).
FRL_DEV void fx_pk_put(unsigned long long* p, unsigned epoch, float v) {
  const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
FRL_DEV unsigned long long fx_pk_ld(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
FRL_DEV float fx_pk_wait1(const unsigned long long* p, unsigned epoch) {
  unsigned spins = 0;
  for (;;) {
    const unsigned long long a = fx_pk_ld(p);
    if ((unsigned)(a >> 32) == epoch) return __uint_as_float((unsigned)a);
    if (++spins > FX_SPIN_LIMIT) __trap();
  }
}
FRL_DEV void fx_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// wait until the (up to three) packets carry `epoch`; returns their values.  Unused slots: pass nullptr.
FRL_DEV void fx_pk_wait3(const unsigned long long* p0, const unsigned long long* p1, const unsigned long long* p2, unsigned epoch,
                         float* v0, float* v1, float* v2) {
  unsigned spins = 0;
  for (;;) {
    const unsigned long long a = fx_pk_ld(p0), b = p1 ? fx_pk_ld(p1) : a, d = p2 ? fx_pk_ld(p2) : a;
    if ((unsigned)(a >> 32) == epoch && (unsigned)(b >> 32) == epoch && (unsigned)(d >> 32) == epoch) {
      *v0 = __uint_as_float((unsigned)a); *v1 = __uint_as_float((unsigned)b); *v2 = __uint_as_float((unsigned)d);
      return;
    }
    if (++spins > FX_SPIN_LIMIT) __trap();
  }
}
FRL_DEV float shx(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
// grid-wide barrier: every CTA adds 1 to *ctr (zeroed by the host before the launch) and waits for `target` arrivals
FRL_DEV void fx_grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  trace(90);
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    trace(91);
    unsigned spins = 0;
    while ((int)(fx_flag_ld(ctr) - target) < 0) {
      if (++spins > FX_SPIN_LIMIT) __trap();
    }
    trace(92);
  }
  __syncthreads();
  trace(93);
}
#else
static inline void fx_flag_set(unsigned* p, unsigned v) { *p = v; }
static inline void fx_flag_wait(const unsigned* p, unsigned v) {
  if ((int)(*p - v) < 0) { fprintf(stderr, "frl emulation: AcFx flag protocol violated (%u < %u)\n", *p, v); abort(); }
}
static inline float fx_ldcg(const float* p) { return *p; }
static inline void fx_pk_put(unsigned long long* p, unsigned epoch, float v) {
  unsigned b; memcpy(&b, &v, 4);
  *p = ((unsigned long long)epoch << 32) | b;
}
static inline float fx_pk_wait1(const unsigned long long* p, unsigned epoch) {
  if ((unsigned)(*p >> 32) != epoch) { fprintf(stderr, "frl emulation: AcFx packet protocol violated\n"); abort(); }
  const unsigned b = (unsigned)*p; float v; memcpy(&v, &b, 4);
  return v;
}
static inline void fx_prefetch_l1(const void*) {}
static inline void fx_pk_wait3(const unsigned long long* p0, const unsigned long long* p1, const unsigned long long* p2, unsigned epoch,
                               float* v0, float* v1, float* v2) {
  const unsigned long long* ps[3] = {p0, p1 ? p1 : p0, p2 ? p2 : p0};
  float* vs[3] = {v0, v1, v2};
  for (int i = 0; i < 3; ++i) {
    if ((unsigned)(*ps[i] >> 32) != epoch) { fprintf(stderr, "frl emulation: AcFx packet protocol violated\n"); abort(); }
    const unsigned b = (unsigned)*ps[i]; memcpy(vs[i], &b, 4);
  }
}
// per-thread scratch standing in for the register files of one CTA when a warp shuffle has to be emulated
static thread_local float fx_emu_a[FRL_NT][64];
static thread_local float fx_emu_b[FRL_NT][64];
// one reduce-scatter step on NV values per thread: thread t keeps half (t & m ? 1 : 0) and adds its partner's copy of that half
static inline void fx_emu_rs(float (*src)[64], float (*dst)[64], int NV, int m) {
  for (int t = 0; t < FRL_NT; ++t) {
    const int h = (t & m) ? NV / 2 : 0;
    for (int i = 0; i < NV / 2; ++i) dst[t][i] = fadd(src[t][h + i], src[t ^ m][h + i]);
  }
}
// butterfly sum step: every thread adds its partner's value
static inline void fx_emu_bf(float (*src)[64], float (*dst)[64], int NV, int m) {
  for (int t = 0; t < FRL_NT; ++t)
    for (int i = 0; i < NV; ++i) dst[t][i] = fadd(src[t][i], src[t ^ m][i]);
}
#endif

// ------------------------------------------------------------------------------------------------------------------------
// microkernels on one 8-row tile, 256 threads.  Activation tiles are [8][128] (hidden) or [8][ld] (inputs), zero padded.
// ------------------------------------------------------------------------------------------------------------------------

// Y[8][128] = act(X[8][4 nch] * WT[4 nch][128] + bias)         (layer 0: nch = in_pad / 4 <= 8; layer 1: nch = 32)
//   warp w owns columns 16w..16w+15; lane (kq = lane / 4, cg = lane % 4) accumulates all 8 rows x 4 columns over the K chunks
//   kq, kq + 8, ...; the 8 K-partials are combined by a reduce-scatter over lane bits 4, 3, 2 that leaves row kq with lane kq.
//   Bank behaviour: a quarter warp reads two consecutive 16-B chunks of X (broadcast) and rows 4c+q of two chunks 16 B x 4
//   wide whose row offset differs by 4 * 132 = 16 (mod 32) words — conflict free.
//   xg != nullptr: the output tile is also written to the exchange blocks it belongs to (xg = region + (blk0 Rmax + row0) 16,
//   rm16 = 16 Rmax): element (r, col) -> xg[(col / 16) rm16 + 16 r + col % 16] — the fire-and-forget global store of the
//   epilogue replaces a separate smem -> global pass.
template <int NCH_STATIC>
FRL_NI_GEMM void fx_fwd(const float* X, int ldx, int nch_rt, const float* W, const float* bias, int relu, float* Y, float* xg = nullptr,
                        int rm16 = 0) {
  const int nch = NCH_STATIC > 0 ? NCH_STATIC : nch_rt;
  trace(50);
  FRL_PAR(t) {
    const int w = t >> 5, l = t & 31, cg = l & 3, kq = l >> 2, col = 16 * w + 4 * cg;
    const sptr sX = sp_of(X), sW = sp_of(W);
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    constexpr int NIT = NCH_STATIC > 0 ? (NCH_STATIC + 7) / 8 : 1;      // layer 0 has in_pad <= 32: one pass
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int c = kq + 8 * it;
      if (NCH_STATIC <= 0 && c >= nch) break;
      const int wb = 4 * c * FX_LDW + col;
      const float4 b0 = sp_ld4(sW, wb), b1 = sp_ld4(sW, wb + FX_LDW), b2 = sp_ld4(sW, wb + 2 * FX_LDW), b3 = sp_ld4(sW, wb + 3 * FX_LDW);
      float4 av[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) av[r] = sp_ld4(sX, r * ldx + 4 * c);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        fma2_bcast(acc[r][0], acc[r][1], av[r].x, b0.x, b0.y); fma2_bcast(acc[r][2], acc[r][3], av[r].x, b0.z, b0.w);
        fma2_bcast(acc[r][0], acc[r][1], av[r].y, b1.x, b1.y); fma2_bcast(acc[r][2], acc[r][3], av[r].y, b1.z, b1.w);
        fma2_bcast(acc[r][0], acc[r][1], av[r].z, b2.x, b2.y); fma2_bcast(acc[r][2], acc[r][3], av[r].z, b2.z, b2.w);
        fma2_bcast(acc[r][0], acc[r][1], av[r].w, b3.x, b3.y); fma2_bcast(acc[r][2], acc[r][3], av[r].w, b3.z, b3.w);
      }
    }
#ifndef FRL_EMUL
    const bool h4 = (l & 16) != 0, h3 = (l & 8) != 0, h2 = (l & 4) != 0;
    float v1[4][4], v2[2][4], v3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float keep = h4 ? acc[i + 4][j] : acc[i][j], send = h4 ? acc[i][j] : acc[i + 4][j];
        v1[i][j] = keep + shx(send, 16);
      }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float keep = h3 ? v1[i + 2][j] : v1[i][j], send = h3 ? v1[i][j] : v1[i + 2][j];
        v2[i][j] = keep + shx(send, 8);
      }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float keep = h2 ? v2[1][j] : v2[0][j], send = h2 ? v2[0][j] : v2[1][j];
      v3[j] = keep + shx(send, 4);
    }
    const float4 bv = sp_ld4(sp_of(bias), col);
    float o[4] = {v3[0] + bv.x, v3[1] + bv.y, v3[2] + bv.z, v3[3] + bv.w};
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = o[j] > 0.f ? o[j] : 0.f;
    }
    sp_st4(sp_of(Y), kq * 128 + col, make_float4(o[0], o[1], o[2], o[3]));
    if (xg) st4(xg + (size_t)w * rm16 + kq * 16 + 4 * cg, make_float4(o[0], o[1], o[2], o[3]));
#else
    for (int r = 0; r < 8; ++r)
      for (int j = 0; j < 4; ++j) fx_emu_a[t][r * 4 + j] = acc[r][j];
#endif
  }
#ifdef FRL_EMUL
  fx_emu_rs(fx_emu_a, fx_emu_b, 32, 16);
  fx_emu_rs(fx_emu_b, fx_emu_a, 16, 8);
  fx_emu_rs(fx_emu_a, fx_emu_b, 8, 4);
  FRL_PAR(t) {
    const int w = t >> 5, l = t & 31, cg = l & 3, kq = l >> 2, col = 16 * w + 4 * cg;
    for (int j = 0; j < 4; ++j) {
      float o = fx_emu_b[t][j] + bias[col + j];
      if (relu) o = o > 0.f ? o : 0.f;
      Y[kq * 128 + col + j] = o;
      if (xg) xg[(size_t)w * rm16 + kq * 16 + 4 * cg + j] = o;
    }
  }
#endif
  trace(51);
  FRL_SYNC();
  trace(53);
}

// dX[8][128] = (dY[8][128] * W[128][128]) * relu'(mask)  on the forward image WT[k][n] (row k holds column k of W):
//   dX[r][k] = sum_n dY[r][n] WT[k][n].  Warp w owns outputs k = 16w..16w+15; lane (cg = lane / 8, nq = lane % 8) accumulates
//   8 rows x 4 outputs (k = 16w + 4cg + i) over the n chunks nq, nq + 8, ... with even / odd n in separate accumulators (both
//   operands are natural register pairs of the 16-B loads -> FFMA2); reduce-scatter over lane bits 2, 1, 0 leaves row nq.
//   A quarter warp reads one 128-B row segment of WT and one of dY: conflict free.
FRL_NI_GEMM void fx_bwd(const float* dY, const float* W, const float* mask, float* dX, float* xg = nullptr, int rm16 = 0) {
  trace(60);
  FRL_PAR(t) {
    const int w = t >> 5, l = t & 31, nq = l & 7, cg = l >> 3, k0 = 16 * w + 4 * cg;
    const sptr sY = sp_of(dY), sW = sp_of(W);
    float acc[8][4], ach[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[r][i] = 0.f; ach[r][i] = 0.f; }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int c = nq + 8 * it;
      float4 bv[4], av[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) bv[i] = sp_ld4(sW, (k0 + i) * FX_LDW + 4 * c);
#pragma unroll
      for (int r = 0; r < 8; ++r) av[r] = sp_ld4(sY, r * 128 + 4 * c);
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          fma2_elem(acc[r][i], ach[r][i], av[r].x, av[r].y, bv[i].x, bv[i].y);
          fma2_elem(acc[r][i], ach[r][i], av[r].z, av[r].w, bv[i].z, bv[i].w);
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][i] += ach[r][i];
#ifndef FRL_EMUL
    const bool h2 = (l & 4) != 0, h1 = (l & 2) != 0, h0 = (l & 1) != 0;
    float v1[4][4], v2[2][4], v3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float keep = h2 ? acc[i + 4][j] : acc[i][j], send = h2 ? acc[i][j] : acc[i + 4][j];
        v1[i][j] = keep + shx(send, 4);
      }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float keep = h1 ? v1[i + 2][j] : v1[i][j], send = h1 ? v1[i][j] : v1[i + 2][j];
        v2[i][j] = keep + shx(send, 2);
      }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float keep = h0 ? v2[1][j] : v2[0][j], send = h0 ? v2[0][j] : v2[1][j];
      v3[j] = keep + shx(send, 1);
    }
    const float4 mv = sp_ld4(sp_of(mask), nq * 128 + k0);
    const float4 ov = make_float4(mv.x > 0.f ? v3[0] : 0.f, mv.y > 0.f ? v3[1] : 0.f, mv.z > 0.f ? v3[2] : 0.f, mv.w > 0.f ? v3[3] : 0.f);
    sp_st4(sp_of(dX), nq * 128 + k0, ov);
    if (xg) st4(xg + (size_t)w * rm16 + nq * 16 + 4 * cg, ov);
#else
    for (int r = 0; r < 8; ++r)
      for (int i = 0; i < 4; ++i) fx_emu_a[t][r * 4 + i] = acc[r][i];
#endif
  }
#ifdef FRL_EMUL
  fx_emu_rs(fx_emu_a, fx_emu_b, 32, 4);
  fx_emu_rs(fx_emu_b, fx_emu_a, 16, 2);
  fx_emu_rs(fx_emu_a, fx_emu_b, 8, 1);
  FRL_PAR(t) {
    const int w = t >> 5, l = t & 31, nq = l & 7, cg = l >> 3, k0 = 16 * w + 4 * cg;
    for (int j = 0; j < 4; ++j) {
      const float o = mask[nq * 128 + k0 + j] > 0.f ? fx_emu_b[t][j] : 0.f;
      dX[nq * 128 + k0 + j] = o;
      if (xg) xg[(size_t)w * rm16 + nq * 16 + 4 * cg + j] = o;
    }
  }
#endif
  trace(61);
  FRL_SYNC();
  trace(63);
}

// narrow head forward:  Y[8][ldy] (columns < NO) = X[8][128] * WT[128][ld] + bias,  NO = 4 or 8 (the layer's out_pad).
//   warp r = row, lane l owns k = l, l + 32, l + 64, l + 96 (consecutive lanes read consecutive 16-B / 48-B rows of WT);
//   the 32 K-partials of each output are summed by a 5-step butterfly (every lane ends with the same total).
FRL_NI_GEMM void fx_fwd_narrow(const float* X, const float* W, const float* bias, int NO, int ld, float* Y, int ldy) {
  trace(54);
  FRL_PAR(t) {
    const int r = t >> 5, l = t & 31;
    const sptr sX = sp_of(X), sW = sp_of(W);
    float p[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) p[n] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = l + 32 * q;
      const float a = sp_ld1(sX, r * 128 + k);
      const float4 b0 = sp_ld4(sW, k * ld);
      p[0] += a * b0.x; p[1] += a * b0.y; p[2] += a * b0.z; p[3] += a * b0.w;
      if (NO > 4) {
        const float4 b1 = sp_ld4(sW, k * ld + 4);
        p[4] += a * b1.x; p[5] += a * b1.y; p[6] += a * b1.z; p[7] += a * b1.w;
      }
    }
#ifndef FRL_EMUL
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
#pragma unroll
      for (int n = 0; n < 8; ++n) p[n] += shx(p[n], m);
    if (l < NO) {
      float v = p[0];
#pragma unroll
      for (int n = 1; n < 8; ++n) v = (l == n) ? p[n] : v;
      Y[r * ldy + l] = v + bias[l];
    }
#else
    for (int n = 0; n < 8; ++n) fx_emu_a[t][n] = p[n];
#endif
  }
#ifdef FRL_EMUL
  fx_emu_bf(fx_emu_a, fx_emu_b, 8, 16);
  fx_emu_bf(fx_emu_b, fx_emu_a, 8, 8);
  fx_emu_bf(fx_emu_a, fx_emu_b, 8, 4);
  fx_emu_bf(fx_emu_b, fx_emu_a, 8, 2);
  fx_emu_bf(fx_emu_a, fx_emu_b, 8, 1);
  FRL_PAR(t) {
    const int r = t >> 5, l = t & 31;
    if (l < NO) Y[r * ldy + l] = fx_emu_b[t][l] + bias[l];
  }
#endif
  trace(55);
  FRL_SYNC();
}

// narrow head backward:  dH[8][128] = (dOut[8][ldo](columns < NO) * W) * relu'(mask):  dH[r][k] = sum_n dOut[r][n] WT[k][n]
FRL_NI_GEMM void fx_bwd_narrow(const float* dOut, int ldo, const float* W, int NO, int ld, const float* mask, float* dH,
                               float* xg = nullptr, int rm16 = 0) {
  trace(64);
  // thread = (output column k, rows 4 hi .. 4 hi + 3): consecutive lanes read consecutive 16-B (NO = 4) / 48-B (NO = 8) rows of WT
  // and write consecutive words of dH — no bank conflicts; the dOut rows are warp-wide broadcasts
  FRL_PAR(t) {
    const int k = t & 127, hi = t >> 7;
    const sptr sW = sp_of(W), sO = sp_of(dOut), sM = sp_of(mask), sH = sp_of(dH);
    const float4 w0 = sp_ld4(sW, k * ld);
    float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (NO > 4) w1 = sp_ld4(sW, k * ld + 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 4 * hi + i;
      const float4 d0 = sp_ld4(sO, r * ldo);
      float s = 0.f;
      s += d0.x * w0.x; s += d0.y * w0.y; s += d0.z * w0.z; s += d0.w * w0.w;
      if (NO > 4) {
        const float4 d1 = sp_ld4(sO, r * ldo + 4);
        s += d1.x * w1.x; s += d1.y * w1.y; s += d1.z * w1.z; s += d1.w * w1.w;
      }
      const float o = sp_ld1(sM, r * 128 + k) > 0.f ? s : 0.f;
      sp_st1(sH, r * 128 + k, o);
      if (xg) xg[(size_t)(k >> 4) * rm16 + r * 16 + (k & 15)] = o;
    }
  }
  trace(65);
  FRL_SYNC();
}

// gradient wrt NA (<= 8) input columns c0.. of layer 0 (dQ/da):  dXa[r][j] = sum_n D1[r][n] WT0[c0 + j][n]
FRL_NI_GEMM void fx_bwd_cols(const float* D1, const float* W0, int c0, int NA, float* dXa /*[8][8]*/) {
  trace(66);
  FRL_PAR(t) {
    const int r = t >> 5, l = t & 31, j = l & 7, nq = l >> 3;
    float p = 0.f;
    if (j < NA) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = ld4(D1 + r * 128 + 32 * nq + 4 * i), b = ld4(W0 + (c0 + j) * FX_LDW + 32 * nq + 4 * i);
        p += a.x * b.x; p += a.y * b.y; p += a.z * b.z; p += a.w * b.w;
      }
    }
#ifndef FRL_EMUL
    p += shx(p, 8);
    p += shx(p, 16);
    if (nq == 0) dXa[r * 8 + j] = p;
#else
    fx_emu_a[t][0] = p;
#endif
  }
#ifdef FRL_EMUL
  fx_emu_bf(fx_emu_a, fx_emu_b, 1, 8);
  fx_emu_bf(fx_emu_b, fx_emu_a, 1, 16);
  FRL_PAR(t) {
    const int r = t >> 5, l = t & 31;
    if ((l >> 3) == 0) dXa[r * 8 + (l & 7)] = fx_emu_a[t][0];
  }
#endif
  trace(67);
  FRL_SYNC();
}

// ------------------------------------------------------------------------------------------------------------------------
// exchange buffer: layer inputs X and pre-activation gradients dY of every tile, blocked [block of 16 columns][row][16] so
// that the operands of one dW job are contiguous (1-D TMA bulk copies)
// ------------------------------------------------------------------------------------------------------------------------
FRL_HD int fx_xb(const frl_net_t& n, int li) { return (n.L[li].in_pad + 15) >> 4; }
FRL_HD int fx_yb(const frl_net_t& n, int li) { return (n.L[li].out_pad + 15) >> 4; }
FRL_HD int fx_xblocks(const frl_net_t& n, int upto) { int b = 0; for (int i = 0; i < upto; ++i) b += fx_xb(n, i); return b; }
FRL_HD int fx_yblocks(const frl_net_t& n, int upto) { int b = 0; for (int i = 0; i < upto; ++i) b += fx_yb(n, i); return b; }

// copy a tile T[8][ld] (columns [0, 16 nblk)) into blocks blk0.. of an exchange region: rows row0..row0+7
FRL_DEV void fx_put_blocks(float* region, int blk0, int nblk, int Rmax, int row0, const float* T, int ld, int width) {
  FRL_PAR(t) {
    const int sh = nblk == 8 ? 5 : (nblk == 2 ? 3 : (nblk == 1 ? 2 : -1));        // log2(4 nblk) for the block counts in use
    for (int e = t; e < 8 * nblk * 4; e += FRL_NT) {
      const int r = sh >= 0 ? (e >> sh) : e / (nblk * 4), q = e - r * (nblk * 4), b = q >> 2, c4 = (q & 3) * 4, col = b * 16 + c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col + 3 < width) v = ld4(T + r * ld + col);
      else {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < 4; ++i) if (col + i < width) o[i] = T[r * ld + col + i];
        v = make_float4(o[0], o[1], o[2], o[3]);
      }
      st4(region + ((size_t)(blk0 + b) * Rmax + row0 + r) * 16 + c4, v);
    }
  }
}

struct FxJob {
  int li;         // absolute layer index (-1: the extras job, e.g. SAC log_std)
  int n0, k0;     // first output row / input column
  int NB, KB;     // tile height / width (NB * KB in {256, 512})
};

// Sampled rows of one tile -> raw[8][row_floats] without passing through registers (LDGSTS): issued one optimiser stage ahead of
// the phase that consumes them, so the index load, the DRAM latency of the random rows and the smem write are off the chain.
FRL_DEV void fx_gather_async(const float* storage, int row_floats, const int64_t* idx, int nvalid, float* raw) {
  const int q = row_floats >> 2;
  FRL_PAR(t) {
    for (int e = t; e < 8 * q; e += FRL_NT) {
      const int r = e / q, j = e - r * q;
      float* dst = raw + r * row_floats + 4 * j;
      if (r < nvalid) {
        const float* src = storage + (size_t)idx[r] * row_floats + 4 * j;
#ifndef FRL_EMUL
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
#else
        memcpy(dst, src, 16);
#endif
      } else {
        st4(dst, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
#ifndef FRL_EMUL
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
  }
}
FRL_DEV void fx_gather_wait() {
#ifndef FRL_EMUL
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
  FRL_SYNC();
}

// what stage() needs from the argument block, derived once per launch (AcFx::build_plan) and kept in shared memory
struct FxJobP { int li, n0, k0, NB, KB, ksh, yb0, xb0, xbn, pad; };      // li = -2: no job in this slot; -1: the extras job
struct FxPlan {
  int role, wi, tile, h, l0, row0, nvalid, njobs_c, njobs_a;
  int cxb[3], cyb[3], axb[3], ayb[3];                                    // first exchange block of this CTA's critic head / the actor layers
  int w_cx, w_cy, w_ax, w_ay0, w_ay1, w_lsg, w_stats, w_sumsq;
  int coff[3], cboff[3], aoff[3], aboff[3];                              // layer image / bias offsets inside a resident head slot (floats)
  FxJobP jc[2], ja[2];
};
#define FX_PLAN_FLOATS 96
static_assert(sizeof(FxPlan) <= FX_PLAN_FLOATS * 4, "FxPlan outgrew its shared-memory reservation");

// Start fetching the 3-layer head (l0..l0+2 of net n, contiguous in the mirror) into slot s unless it is already resident:
// layer 0 on the slot's first mbarrier (small: the first op can start early), layers 1 + 2 as ONE copy on the second.  The two
// copies are issued by the leaders of two different warps (`lead`, `lead + 32`) — each issue is a proxy fence (~1 k clk) plus
// ~0.7 k clk of expect_tx / UBLKCP latency, and a stage that needs two heads gives them four different warps.
FRL_DEV void fx_fetch(Cta& c, int s, const frl_net_t& n, int l0, int lead) {
  const float* key = res_key(n, l0);
  if ((s ? c.stag1 : c.stag0) == key) return;
  res_drain_slot(c, s);
  if (s) c.stag1 = key; else c.stag0 = key;
  float* dst = s ? c.wbuf1 : c.wbuf0;
  const int f0 = wt_floats(n.L[l0]), f12 = wt_floats(n.L[l0 + 1]) + wt_floats(n.L[l0 + 2]);
#ifndef FRL_EMUL
  uint64_t* bars = c.bar + 2 + s * 3;
  if ((int)threadIdx.x == lead) {
    fence_proxy_async();
    mbar_expect_tx(bars, (uint32_t)f0 * 4u);
    tma_bulk_g2s(dst, key, (uint32_t)f0 * 4u, bars);
  } else if ((int)threadIdx.x == lead + 32) {
    fence_proxy_async();
    mbar_expect_tx(bars + 1, (uint32_t)f12 * 4u);
    tma_bulk_g2s(dst + f0, key + f0, (uint32_t)f12 * 4u, bars + 1);
  }
#else
  (void)lead;
  memcpy(dst, key, (size_t)(f0 + f12) * 4);
#endif
  c.spend |= 3u << (s * 3);
}
// smem image of layer l0 + k of the head fetched into slot s by fx_fetch (waits for its copy if still outstanding)
FRL_DEV const float* fx_layer(Cta& c, int s, const frl_net_t& n, int l0, int k) {
  const uint32_t bit = 1u << (s * 3 + (k ? 1 : 0));
  if (c.spend & bit) {
#ifndef FRL_EMUL
    mbar_wait(c.bar + 2 + s * 3 + (k ? 1 : 0), (c.sphase >> (s * 3 + (k ? 1 : 0))) & 1u);
#endif
    c.sphase ^= bit; c.spend &= ~bit;
  }
  return (s ? c.wbuf1 : c.wbuf0) + (n.L[l0 + k].wt_off - n.L[l0].wt_off);
}

// wait (if still outstanding) for layer k of the head fx_fetch put into slot s
FRL_DEV void fx_wait(Cta& c, int s, int k) {
  const uint32_t bit = 1u << (s * 3 + (k ? 1 : 0));
  if (c.spend & bit) {
#ifndef FRL_EMUL
    mbar_wait(c.bar + 2 + s * 3 + (k ? 1 : 0), (c.sphase >> (s * 3 + (k ? 1 : 0))) & 1u);
#endif
    c.sphase ^= bit; c.spend &= ~bit;
  }
}

// Sum of slot[0..31], written by the lanes of warp 0 in the phase just before (no block barrier needed): the shuffle tree of
// block_sum's first level, so a phase whose other warps contribute zeros gets block_sum's bits.  GPU: valid in thread 0 only.
FRL_DEV float fx_tree32(float* slot) {
#ifndef FRL_EMUL
  float v = 0.f;
  if (threadIdx.x < 32) {
    __syncwarp();
    v = slot[threadIdx.x];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  }
  return v;
#else
  for (int off = 16; off > 0; off >>= 1)
    for (int i = 0; i < off; ++i) slot[i] += slot[i + off];
  return slot[0];
#endif
}

struct AcFx {
  typedef frl_ac_args_t Args;
  static const int NSTAGES = 6;
  FRL_SHD int heads(const Args& a) { return a.n_heads; }
  FRL_SHD bool is_sac(const Args& a) { return a.actor_kind == FRL_ACTOR_SAC; }
  FRL_SHD int ntile(const Args& a) { return (a.B + 7) >> 3; }
  FRL_SHD int rmax(const Args& a) { return ntile(a) * 8; }
  // (the 64-bit modulo is a ~500-clk subroutine on the GPU: not paid when every step is a policy step)
  FRL_SHD bool is_policy_step(const Args& a, int u) { return a.policy_freq <= 1 ? true : ((a.total_it0 + u + 1) % a.policy_freq) == 0; }
  FRL_SHD bool stage_enabled(int s, int u, const Args& a) { return s >= 3 ? is_policy_step(a, u) : true; }
  FRL_SHD bool writes_params(int) { return true; }     // every stage publishes something a later TMA copy reads
  FRL_SHD int n_updates(const Args& a) { return a.n_updates; }
  // the optimiser stages open with an all-gather of norm packets, which orders them after every CTA's dW stage by itself
  FRL_SHD bool barrier_after(int s) { return s != 1 && s != 4; }
  FRL_SHD int heads_used(const Args& a) { return is_sac(a) ? a.n_heads : 1; }

  FRL_SHD int head_floats(const frl_net_t& n, int l0) { return wt_floats(n.L[l0]) + wt_floats(n.L[l0 + 1]) + wt_floats(n.L[l0 + 2]); }
  FRL_SHD int wbuf_floats(const Args& a) {
    int x = head_floats(a.actor, 0), y = head_floats(a.critic, 0);
    return ((x > y ? x : y) + 31) & ~31;
  }
  // shared memory after the engine's own (two slots + c.red + mbarriers): see the SmemBump sequence in stage()
  static const int GST_SLOTS = 2;                    // dW jobs one CTA may own (jobs <= GST_SLOTS * grid)
  FRL_SHD int user_floats(const Args& a) {
    return FX_PLAN_FLOATS + 8 * a.replay.row_floats + 3 * 8 * 32 + 64 + 6 * 1024 + 5 * 64 + 2 * 32 + 2 * FRL_NT + GST_SLOTS * 528 + 128 +
           GST_SLOTS * 16 + 16 + 4 + 8 + 64 + 64;
  }
  FRL_SHD int grid(const Args&, int max_ctas) { return max_ctas; }

  // ---- exchange / workspace layout (floats) ----
  struct Ws {
    size_t cx, cy, ax, ay0, ay1, lsg, stats, sumsq, total;
  };
  FRL_SHD Ws ws_layout(const Args& a, int ncta) {
    Ws w;
    const size_t R16 = (size_t)rmax(a) * 16;
    size_t o = 0;
    w.cx = o; o += (size_t)fx_xblocks(a.critic, a.critic.n_layers) * R16;
    w.cy = o; o += (size_t)fx_yblocks(a.critic, a.critic.n_layers) * R16;
    w.ax = o; o += (size_t)fx_xblocks(a.actor, 3) * R16;
    w.ay0 = o; o += (size_t)fx_yblocks(a.actor, 3) * R16;
    w.ay1 = o; o += (size_t)fx_yblocks(a.actor, 3) * R16;
    w.lsg = o; o += (size_t)ncta * 8;
    w.stats = o; o += (size_t)ncta * 8;
    w.sumsq = o; o += 512;
    w.total = (o + 3) & ~(size_t)3;
    return w;
  }

  // ---- dW jobs of a net: layer li % 3 == 0 -> 16 x 32 tiles, 1 -> 16 x 16, 2 -> out_pad x (256 / out_pad); + extras ----
  FRL_SHD void job_shape(const frl_net_t& n, int li, int* NB, int* KB) {
    const int k = li % 3;
    if (k == 0) { *NB = 16; *KB = 32; }
    else if (k == 1) { *NB = 16; *KB = 16; }
    else { *NB = n.L[li].out_pad >= 16 ? 16 : (n.L[li].out_pad >= 8 ? 8 : 4); *KB = *NB == 16 ? 16 : (*NB == 8 ? 32 : 64); }
  }
  FRL_SHD int lg2(int x) { return x >= 64 ? 6 : (x >= 32 ? 5 : (x >= 16 ? 4 : (x >= 8 ? 3 : 2))); }      // tile sides are 4 .. 64
  FRL_SHD int layer_jobs(const frl_net_t& n, int li) {
    int NB, KB;
    job_shape(n, li, &NB, &KB);
    return ((n.L[li].out_pad + NB - 1) >> lg2(NB)) * ((n.L[li].in_pad + KB - 1) >> lg2(KB));
  }
  FRL_SHD int net_jobs(const frl_net_t& n) {
    int j = 0;
    for (int li = 0; li < n.n_layers; ++li) j += layer_jobs(n, li);
    return j + (n.x_len > 0 ? 1 : 0);
  }
  FRL_SHD FxJob job_of(const frl_net_t& n, int j) {
    FxJob J;
    for (int li = 0; li < n.n_layers; ++li) {
      const int cnt = layer_jobs(n, li);
      if (j < cnt) {
        job_shape(n, li, &J.NB, &J.KB);
        const int kbc = (n.L[li].in_pad + J.KB - 1) >> lg2(J.KB);
        const int jn = kbc == 1 ? j : (kbc == 8 ? j >> 3 : (kbc == 2 ? j >> 1 : j / kbc));
        J.li = li; J.n0 = jn * J.NB; J.k0 = (j - jn * kbc) * J.KB;
        return J;
      }
      j -= cnt;
    }
    J.li = -1; J.n0 = J.k0 = 0; J.NB = J.KB = 0;
    return J;
  }

  // host-side eligibility (capi.cu): everything else takes the generic kernel of algo_ac.cuh
  FRL_SHD bool shape_ok(const frl_net_t& n, int l0, int max_out) {
    // (fx_fetch copies layers l0+1, l0+2 as one block: the head's layer images must be contiguous in the mirror)
    if (n.L[l0 + 1].wt_off != n.L[l0].wt_off + wt_floats(n.L[l0]) || n.L[l0 + 2].wt_off != n.L[l0 + 1].wt_off + wt_floats(n.L[l0 + 1])) return false;
    return n.L[l0].out_pad == 128 && n.L[l0 + 1].in_pad == 128 && n.L[l0 + 1].out_pad == 128 && n.L[l0 + 2].in_pad == 128 &&
           n.L[l0].in_pad <= 32 && n.L[l0 + 2].out_pad <= max_out;
  }
  static bool eligible(const Args& a, int max_ctas) {
    if (a.n_agents > 1 || a.obs_norm[0] || !a.ws || !a.sync || a.defer_polyak) return false;
    if (wt_ld_of(128) != FX_LDW) return false;
    if (a.B > FX_MAXB || a.replay.act_dim > 8 || a.n_heads < 1 || a.n_heads > 2) return false;
    if (a.actor.n_layers != 3 || a.critic.n_layers != 3 * a.n_heads) return false;
    if (!shape_ok(a.actor, 0, 8) || !shape_ok(a.actor_target, 0, 8)) return false;
    for (int h = 0; h < a.n_heads; ++h)
      if (!shape_ok(a.critic, 3 * h, 4) || !shape_ok(a.critic_target, 3 * h, 4)) return false;
    if (a.actor.L[2].out_pad < 4 || a.critic.L[2].out_pad != 4) return false;
    for (int k = 0; k < 3; ++k) {       // the plan's slot offsets serve the nets and their targets alike
      if (a.actor_target.L[k].wt_off - a.actor_target.L[0].wt_off != a.actor.L[k].wt_off - a.actor.L[0].wt_off) return false;
      for (int h = 0; h < a.n_heads; ++h)
        if (a.critic_target.L[3 * h + k].wt_off - a.critic_target.L[3 * h].wt_off != a.critic.L[3 * h + k].wt_off - a.critic.L[3 * h].wt_off) return false;
    }
    if (64 + 6 * rmax(a) > 2048 || max_ctas > 512) return false;
    if (is_sac(a) && a.n_heads != 2) return false;
    if (2 * ntile(a) * a.n_heads > max_ctas) return false;
    if (net_jobs(a.critic) > 512 || net_jobs(a.actor) > 512) return false;
    if (net_jobs(a.critic) > GST_SLOTS * max_ctas || net_jobs(a.actor) > GST_SLOTS * max_ctas) return false;
    const size_t smem = (size_t)(cta_base_floats(wbuf_floats(a)) + user_floats(a)) * 4 + 64 + sizeof(Args) + 64;
    return smem <= (size_t)227 * 1024;
  }

  FRL_SDEV float noise_at(const float* ptr, const Args& a, int u, int row, int j, uint32_t stream) {
    if (ptr) return ptr[((size_t)u * a.B + row) * a.replay.act_dim + j];
    return randn_ni(a.seed, stream, (uint32_t)(a.total_it0 + u), (uint32_t)(row * a.replay.act_dim + j));
  }

  // ---- per-CTA plan: what stage() derives from the argument block (roles, workspace offsets, exchange block numbers, the dW
  //      jobs this CTA owns), computed ONCE by thread 0 and kept in shared memory.  Recomputing it per stage cost 2 - 5 k clk of
  //      redundant scalar code at the head of each of the 6 stages of every learn (profiles/r2e_trace_fx_cta64.txt). ----
  FRL_SDEV void plan_jobs(FxJobP* out, const frl_net_t& n, int njobs, const Cta& c) {
    for (int slot = 0; slot < GST_SLOTS; ++slot) {
      FxJobP& Q = out[slot];
      const int j = c.cta + slot * c.ncta;
      Q.li = -2; Q.n0 = Q.k0 = Q.NB = Q.KB = Q.ksh = Q.yb0 = Q.xb0 = Q.xbn = Q.pad = 0;
      if (j >= njobs) continue;
      const FxJob J = job_of(n, j);
      Q.li = J.li; Q.n0 = J.n0; Q.k0 = J.k0; Q.NB = J.NB; Q.KB = J.KB;
      if (J.li < 0) continue;
      Q.ksh = J.KB == 16 ? 4 : (J.KB == 32 ? 5 : 6);
      Q.yb0 = fx_yblocks(n, J.li) + (J.n0 >> 4);
      Q.xb0 = fx_xblocks(n, J.li) + (J.k0 >> 4);
      const int xbn_all = fx_xb(n, J.li) - (J.k0 >> 4);
      Q.xbn = (J.KB >> 4) < xbn_all ? (J.KB >> 4) : xbn_all;
    }
  }
  FRL_SDEV void build_plan(FxPlan& P, const Cta& c, const Args& a) {
    const int NH = a.n_heads, nwork = ntile(a) * NH;
    P.role = c.cta < nwork ? 0 : (c.cta < 2 * nwork ? 1 : 2);
    P.wi = P.role == 0 ? c.cta : c.cta - nwork;
    P.tile = P.wi / NH; P.h = P.wi - P.tile * NH; P.l0 = 3 * P.h; P.row0 = P.tile * 8;
    P.nvalid = (a.B - P.row0) < 8 ? (a.B - P.row0) : 8;
    P.njobs_c = net_jobs(a.critic); P.njobs_a = net_jobs(a.actor);
    for (int k = 0; k < 3; ++k) {
      P.cxb[k] = fx_xblocks(a.critic, P.l0 + k); P.cyb[k] = fx_yblocks(a.critic, P.l0 + k);
      P.axb[k] = fx_xblocks(a.actor, k); P.ayb[k] = fx_yblocks(a.actor, k);
    }
    const Ws W = ws_layout(a, c.ncta);
    P.w_cx = (int)W.cx; P.w_cy = (int)W.cy; P.w_ax = (int)W.ax; P.w_ay0 = (int)W.ay0; P.w_ay1 = (int)W.ay1;
    P.w_lsg = (int)W.lsg; P.w_stats = (int)W.stats; P.w_sumsq = (int)W.sumsq;
    for (int k = 0; k < 3; ++k) {
      P.coff[k] = a.critic.L[P.l0 + k].wt_off - a.critic.L[P.l0].wt_off; P.cboff[k] = P.coff[k] + wt_bias(a.critic.L[P.l0 + k]);
      P.aoff[k] = a.actor.L[k].wt_off - a.actor.L[0].wt_off; P.aboff[k] = P.aoff[k] + wt_bias(a.actor.L[k]);
    }
    plan_jobs(P.jc, a.critic, P.njobs_c, c);
    plan_jobs(P.ja, a.actor, P.njobs_a, c);
  }

  // Operands of one dW job -> the scratch slot, as 1-D bulk copies of whole exchange blocks ([R][16] each) on mbarrier bar[0]:
  // copy 0 = dY block of yreg0, [copy 1 = the same block of yreg1], then the job's X blocks.  Warp q's leader issues copy q
  // (each issue is ~300 clk of uniform-datapath latency plus the proxy fence; spread over warps they overlap).
  FRL_SDEV void job_issue(Cta& c, const FxJobP& J, float* scr, const float* xreg, const float* yreg0, const float* yreg1, int Rm) {
    const int nyb = yreg1 ? 2 : 1, ncopies = nyb + J.xbn;
    const size_t blk = (size_t)Rm * 16;
#ifndef FRL_EMUL
    const int tid = (int)threadIdx.x, q = tid >> 5;
    if ((tid & 31) == 0 && q < ncopies) {
      fence_proxy_async();
      if (q == 0) mbar_expect_tx(c.bar, (uint32_t)ncopies * (uint32_t)Rm * 64u);
      const float* src = q == 0 ? yreg0 + (size_t)J.yb0 * blk : (q < nyb ? yreg1 + (size_t)J.yb0 * blk : xreg + (size_t)(J.xb0 + q - nyb) * blk);
      tma_bulk_g2s(scr + (size_t)q * blk, src, (uint32_t)Rm * 64u, c.bar);
    }
#else
    for (int q = 0; q < ncopies; ++q) {
      const float* src = q == 0 ? yreg0 + (size_t)J.yb0 * blk : (q < nyb ? yreg1 + (size_t)J.yb0 * blk : xreg + (size_t)(J.xb0 + q - nyb) * blk);
      memcpy(scr + (size_t)q * blk, src, blk * 4);
    }
    (void)c;
#endif
  }

  // dW tile of one job:  G[n][k] = sum_R DY[R][n0 + n] X[R][k0 + k]   (+ second DY operand added element-wise; + bias sums)
  //   operands in shared memory: DY0 / DY1 as [R][16] (columns n0 % 16 ..), X as KB / 16 arrays [R][16]
  //   lanes: TL = NB KB / 16 register tiles of 4 x 4 per warp pass, 32 / TL row sub-splits inside the warp, 8 warps split the
  //   rows further (row = split, split + S, ...); warp partials are combined in fixed order through c.red.
  //   Returns (every thread) this thread's share of the tile's sum of squares over the real (unpadded) elements.
  FRL_SDEV float job_compute(Cta& c, const FxJobP& J, const frl_layer_t& L, int R, const float* DY0, const float* DY1, const float* Xs,
                             int ncol0, float* gst, float* gbst, float* redb, float* sq /*[FRL_NT]*/) {
    const int TL = (J.NB * J.KB) >> 4, ktl = J.KB >> 2, RSW = 32 / TL, S = 8 * RSW;
    trace(70);
    FRL_PAR(t) {
      const int w = t >> 5, l = t & 31, tl = l % TL, rs = l / TL, nt = tl / ktl, kt = tl - nt * ktl;
      const sptr xa = sp_of(Xs + (size_t)(kt >> 2) * R * 16 + 4 * (kt & 3));
      const sptr ya = sp_of(DY0 + ncol0 + 4 * nt);
      const bool two = DY1 != nullptr;
      const sptr yb = sp_of((two ? DY1 : DY0) + ncol0 + 4 * nt);
      float acc[4][4], ab[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ab[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      }
#pragma unroll 4
      for (int r = w * RSW + rs; r < R; r += S) {
        float4 av = sp_ld4(ya, r * 16);
        if (two) av = f4add(av, sp_ld4(yb, r * 16));
        const float4 bv = sp_ld4(xa, r * 16);
        fma2_bcast(acc[0][0], acc[0][1], av.x, bv.x, bv.y); fma2_bcast(acc[0][2], acc[0][3], av.x, bv.z, bv.w);
        fma2_bcast(acc[1][0], acc[1][1], av.y, bv.x, bv.y); fma2_bcast(acc[1][2], acc[1][3], av.y, bv.z, bv.w);
        fma2_bcast(acc[2][0], acc[2][1], av.z, bv.x, bv.y); fma2_bcast(acc[2][2], acc[2][3], av.z, bv.z, bv.w);
        fma2_bcast(acc[3][0], acc[3][1], av.w, bv.x, bv.y); fma2_bcast(acc[3][2], acc[3][3], av.w, bv.z, bv.w);
        ab[0] += av.x; ab[1] += av.y; ab[2] += av.z; ab[3] += av.w;
      }
#ifndef FRL_EMUL
      if (RSW == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ab[i] += shx(ab[i], 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += shx(acc[i][j], 16);
        }
      }
      if (rs == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) st4(c.red + w * 512 + (4 * nt + i) * J.KB + 4 * kt, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        if (kt == 0) st4(redb + w * 16 + 4 * nt, make_float4(ab[0], ab[1], ab[2], ab[3]));
      }
#else
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) fx_emu_a[t][i * 4 + j] = acc[i][j];
        fx_emu_a[t][16 + i] = ab[i];
      }
#endif
    }
#ifdef FRL_EMUL
    if (RSW == 2) fx_emu_bf(fx_emu_a, fx_emu_b, 20, 16);
    FRL_PAR(t) {
      const int w = t >> 5, l = t & 31, tl = l % TL, rs = l / TL, nt = tl / ktl, kt = tl - nt * ktl;
      float (*src)[64] = RSW == 2 ? fx_emu_b : fx_emu_a;
      if (rs == 0) {
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j) c.red[w * 512 + (4 * nt + i) * J.KB + 4 * kt + j] = src[t][i * 4 + j];
        if (kt == 0) for (int i = 0; i < 4; ++i) redb[w * 16 + 4 * nt + i] = src[t][16 + i];
      }
    }
#endif
    trace(71);
    FRL_SYNC();
    // warp partials in warp order -> the gradient tile (kept in shared memory for the optimiser stage) + this thread's squares
    FRL_PAR(t) {
      float l = 0.f;
      for (int e = t; e < J.NB * J.KB; e += FRL_NT) {
        float s = c.red[e];
#pragma unroll
        for (int w = 1; w < 8; ++w) s += c.red[w * 512 + e];
        gst[e] = s;
        const int nn = e >> J.ksh, kk = e - (nn << J.ksh);
        if (J.n0 + nn < L.out_pad && J.k0 + kk < L.in_pad) l += s * s;
      }
      if (t < J.NB) {
        float s = redb[t];
#pragma unroll
        for (int w = 1; w < 8; ++w) s += redb[w * 16 + t];
        gbst[t] = s;
        if (J.k0 == 0 && J.n0 + t < L.out_pad) l += s * s;
      }
      sq[t] = l;
    }
    trace(72);
    FRL_SYNC();
    return block_sum(sq);
  }

  // ---- one optimiser element (torch.optim.Adam single-tensor math, same rounding order as engine.cuh::adam_update) ----
  FRL_SDEV void adam_val(const AdamHP& hp, float coef, float gin, float& w, float& mm, float& vv) {
    float g = gin * coef;
    if (hp.weight_decay != 0.f) g = fmaf(w, hp.weight_decay, g);
    mm = fmaf(hp.one_minus_b1, g - mm, mm);
    vv = fadd(fmul(vv, hp.b2), fmul(fmul(hp.one_minus_b2, g), g));
    const float denom = fadd(fdiv(fsqrt(vv), hp.bc2_sqrt), hp.eps);
    w = fadd(w, fdiv(fmul(hp.lr_over_bc1_neg, mm), denom));
  }

  // dW + sum of squares of this CTA's jobs of net n (stage 1 / 4).  `scr` = the weight slot that is dead in this stage.
  FRL_SDEV void dw_stage(Cta& c, const Args& a, const frl_net_t& n, const FxJobP* jobs, const float* xreg, const float* yreg0,
                         const float* yreg1, float* scr, float* gst, float* gbst, float* redb, float* red0,
                         unsigned long long* ss_pk, unsigned epoch, const float* lsg, int nls, const AdamSpec sp, float* hpst) {
    const int Rm = rmax(a);
    // bias corrections of the optimiser stage that follows (double-precision powers, ~1.5 k clk on one thread): computed here,
    // by a thread that has nothing to do while the operands are in flight, and kept in shared memory
    FRL_PAR(t) {
      if (t == FRL_NT - 1) {
        const AdamHP h = adam_hp_ni(sp.lr, sp.b1, sp.b2, sp.eps, sp.wd, sp.max_norm, sp.step);
        hpst[0] = h.lr_over_bc1_neg; hpst[1] = h.bc2_sqrt; hpst[2] = h.one_minus_b1; hpst[3] = h.b2; hpst[4] = h.one_minus_b2;
        hpst[5] = h.eps; hpst[6] = h.weight_decay; hpst[7] = h.max_norm;
      }
    }
    float ss_cta = 0.f;
    for (int slot = 0; slot < GST_SLOTS; ++slot) {
      const FxJobP J = jobs[slot];
      if (J.li == -2) break;
      float ss = 0.f;
      if (J.li < 0) {
        // extras (SAC log_std): fixed-order sum of the per-CTA partials written in phase C
        FRL_PAR(t) {
          float l = 0.f;
          if (t < n.x_len) {
            float s = 0.f;
            for (int k = 0; k < nls; ++k) s += fx_ldcg(lsg + (size_t)k * 8 + t);
            gst[slot * 528 + t] = s;
            l = s * s;
          }
          red0[t] = l;
        }
        FRL_SYNC();
        ss = block_sum(red0);
      } else {
        const int nyb = yreg1 ? 2 : 1;
        float* DY0 = scr;
        float* DY1 = yreg1 ? scr + (size_t)Rm * 16 : nullptr;
        float* Xs = scr + (size_t)nyb * Rm * 16;
        // (blocks beyond the layer's last X block are not loaded; they are zeroed below so no stale NaN enters the sums)
        trace(24);
        job_issue(c, J, scr, xreg, yreg0, yreg1, Rm);
        trace(73);
        if (J.xbn < (J.KB >> 4)) {
          FRL_PAR(t) { for (int e = t; e < ((J.KB >> 4) - J.xbn) * Rm * 16; e += FRL_NT) Xs[(size_t)J.xbn * Rm * 16 + e] = 0.f; }
          FRL_SYNC();
        }
        stage_wait(c, 0);
        trace(74);
        ss = job_compute(c, J, n.L[J.li], Rm, DY0, DY1, Xs, J.n0 & 15, gst + slot * 528, gbst + slot * 16, redb, red0);
      }
      ss_cta += ss;
    }
    // this CTA's share of the squared gradient norm, as an epoch-tagged packet: publishing it IS the arrival at the norm
    // all-gather the optimiser stage opens with (no grid barrier between the two stages on the GPU)
    FRL_PAR(t) { if (t == 0) fx_pk_put(ss_pk, epoch, ss_cta); }
    FRL_SYNC();
  }

  // clip + Adam (+ Polyak of tgt) on the elements of this CTA's jobs (stage 2 / 5).  Returns the total squared norm.
  FRL_SDEV float opt_stage(Cta& c, const frl_net_t& n, const frl_net_t* tgt, float tau, const float* hpst, const float* gst,
                           const float* gbst, const unsigned long long* ss_pk, unsigned epoch, float* sh, const FxJobP* jobs) {
    trace(80);
    // the optimiser operands of this CTA's first job start moving towards L1 while the norm packets are polled
    {
      const FxJobP J = jobs[0];
      if (J.li >= 0) {
        const frl_layer_t& L = n.L[J.li];
        FRL_PAR(t) {
          // one 16-element row segment of the tile per thread (NB KB / 16 <= 32 segments), for p, m, v and the target
          const int segs = (J.NB * J.KB) >> 4, seg = t & 31, which = t >> 5;
          if (seg < segs && which < 4) {
            const int nn = (seg << 4) >> J.ksh, kk = (seg << 4) - (nn << J.ksh), row = J.n0 + nn, col = J.k0 + kk;
            if (row < L.out_pad && col < L.in_pad) {
              const int pi = L.w_off + row * L.in_pad + col;
              const float* base = which == 0 ? n.p : (which == 1 ? n.m : (which == 2 ? n.v : (tgt ? tgt->p : nullptr)));
              if (base) fx_prefetch_l1(base + pi);
            }
          }
        }
      }
    }
    // total squared gradient norm: one packet per CTA (written at the end of its dW stage), folded in a fixed order (strided
    // per-thread sums, then the block tree), so every CTA computes the same number
    FRL_PAR(t) {
      float v = 0.f;
      for (int i = t; i < c.ncta; i += FRL_NT) v += fx_pk_wait1(ss_pk + i, epoch);
      sh[t] = v;
    }
    FRL_SYNC();
    const float total = block_sum(sh);
    trace(82);
    AdamHP hp;
    hp.lr_over_bc1_neg = hpst[0]; hp.bc2_sqrt = hpst[1]; hp.one_minus_b1 = hpst[2]; hp.b2 = hpst[3]; hp.one_minus_b2 = hpst[4];
    hp.eps = hpst[5]; hp.weight_decay = hpst[6]; hp.max_norm = hpst[7];
    float coef = 1.f;
    if (hp.max_norm > 0.f) {
      coef = hp.max_norm / (sqrtf(total) + 1e-6f);
      if (coef > 1.f) coef = 1.f;
    }
    const float omt = (float)(1.0 - (double)tau);
    for (int slot = 0; slot < GST_SLOTS; ++slot) {
      const FxJobP J = jobs[slot];
      if (J.li == -2) break;
      FRL_PAR(t) {
        if (J.li < 0) {
          if (t < n.x_len) {
            float w = n.p[n.x_off + t], mm = n.m[n.x_off + t], vv = n.v[n.x_off + t];
            adam_val(hp, coef, gst[slot * 528 + t], w, mm, vv);
            n.p[n.x_off + t] = w; n.m[n.x_off + t] = mm; n.v[n.x_off + t] = vv;
            if (tgt) tgt->p[n.x_off + t] = fadd(fmul(tgt->p[n.x_off + t], omt), fmul(w, tau));
          }
        } else {
          // up to two tile elements + one bias element per thread; every load is issued before the first dependent use so the
          // L2 round trips of p / m / v / target overlap (they were serialised behind the stores of the previous element)
          const frl_layer_t& L = n.L[J.li];
          const int ldw = wt_ld(L);
          int pi[3], mi[3];
          float gg[3], pw[3], pm[3], pv[3], tp[3];
          bool ok[3];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int e = t + q * FRL_NT;
            ok[q] = false; pi[q] = mi[q] = 0; gg[q] = 0.f;
            if (e < J.NB * J.KB) {
              const int nn = e >> J.ksh, kk = e - (nn << J.ksh), row = J.n0 + nn, col = J.k0 + kk;
              if (row < L.out_pad && col < L.in_pad) {
                ok[q] = true; pi[q] = L.w_off + row * L.in_pad + col; mi[q] = L.wt_off + col * ldw + row; gg[q] = gst[slot * 528 + e];
              }
            }
          }
          ok[2] = J.k0 == 0 && t < J.NB && J.n0 + t < L.out_pad;
          pi[2] = L.b_off + J.n0 + t; mi[2] = L.wt_off + wt_bias(L) + J.n0 + t; gg[2] = ok[2] ? gbst[slot * 16 + t] : 0.f;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            pw[q] = pm[q] = pv[q] = tp[q] = 0.f;
            if (ok[q]) {
              pw[q] = n.p[pi[q]]; pm[q] = n.m[pi[q]]; pv[q] = n.v[pi[q]];
              if (tgt) tp[q] = tgt->p[pi[q]];
            }
          }
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (ok[q]) {
              adam_val(hp, coef, gg[q], pw[q], pm[q], pv[q]);
              n.p[pi[q]] = pw[q]; n.m[pi[q]] = pm[q]; n.v[pi[q]] = pv[q];
              n.pt[mi[q]] = pw[q];
              if (tgt) {
                const float tw = fadd(fmul(tp[q], omt), fmul(pw[q], tau));
                tgt->p[pi[q]] = tw; tgt->pt[mi[q]] = tw;
              }
            }
          }
        }
      }
    }
    trace(83);
    FRL_SYNC();
    return total;
  }

  FRL_SDEV void stage(int s, int u, Cta& c, float* user, const Args& a) {
    const frl_net_t& A = a.actor;
    const frl_net_t& C = a.critic;
    const frl_replay_t& rb = a.replay;
    trace(1);
    SmemBump sb; sb.p = user;
    FxPlan* Pp = reinterpret_cast<FxPlan*>(sb.take(FX_PLAN_FLOATS));
    float* raw = sb.take(8 * rb.row_floats);
    float* XA = sb.take(8 * 32);             // actor input  [8][aip]
    float* XS = sb.take(8 * 32);             // critic input [8][cip] = [obs | act]
    float* XN = sb.take(8 * 32);             // [next_obs | a'] (target set) or [obs | pi(obs)] (phase C)
    float* dXa = sb.take(64);                // dQ/da [8][8]
    float* H1 = sb.take(1024);
    float* H2 = sb.take(1024);
    float* A1 = sb.take(1024);               // actor activations: live from phase A to phase C on the online set
    float* A2 = sb.take(1024);
    float* D1 = sb.take(1024);
    float* D2 = sb.take(1024);
    float* MU = sb.take(64);
    float* UU = sb.take(64);
    float* AC = sb.take(64);
    float* dMU = sb.take(64);
    float* EPS = sb.take(64);
    float* QA = sb.take(32);
    float* dQA = sb.take(32);
    float* red0 = sb.take(FRL_NT);
    float* red1 = sb.take(FRL_NT);
    float* gst = sb.take(GST_SLOTS * 528);   // gradient tile(s) of this CTA's dW job(s): kept from the dW stage to the optimiser stage
    float* redb = sb.take(128);
    float* gbst = sb.take(GST_SLOTS * 16);
    float* hpst = sb.take(16);               // optimiser scalars: written in the dW stage, read in the optimiser stage
    float* alpha_s = sb.take(4);             // exp(log_alpha) of this learn: read in phase A (y), reused in phase C
    float* LS = sb.take(8);                  // the actor's log_std as phase A read it (it changes in stage 5 only): reused in phase C
    float* NZ = sb.take(64);                 // this tile's sampling noise [8][8], drawn / loaded at the head of phase A
    if (s == 0 && u == 0) {
      FRL_PAR(t) { if (t == 0) build_plan(*Pp, c, a); }
      FRL_SYNC();
    }
    const FxPlan& P = *Pp;
    const int od = rb.obs_dim, ad = rb.act_dim, rf = rb.row_floats;
    const int NH = a.n_heads, Rm = rmax(a), rm16 = Rm * 16, nwork = ntile(a) * NH;
    const bool sac = is_sac(a);
    const int hu = heads_used(a);
    const bool policy_step = is_policy_step(a, u);
    const int role = P.role, wi = P.wi, h = P.h, l0 = P.l0, row0 = P.row0, nvalid = P.nvalid;
    const float invB = 1.0f / (float)a.B;
    const int aip = A.L[0].in_pad, cip = C.L[0].in_pad, ap = A.L[2].out_pad, ald = wt_ld(A.L[2]), cld = wt_ld(C.L[2]);
    float* ws = a.ws;
    // hand-off packets of the target set (zeroed with the barrier word at every launch): per batch row {Q'_0, Q'_1, log pi(a'|s')}
    unsigned long long* pk = reinterpret_cast<unsigned long long*>(a.sync + 64);
    unsigned long long* sspk = reinterpret_cast<unsigned long long*>(a.sync + 2048);      // norm packets: [0, 512) critic, [512, 1024) actor
    const unsigned epoch = (unsigned)u + 1u;
    float* const slot0 = c.wbuf0;
    float* const slot1 = c.wbuf1;
    const bool metrics_cta = c.cta == c.ncta - 1;        // the last CTA owns the lightest jobs: it also folds the metrics

    if (s == 0) {
      if (role == 2) return;
      const int64_t* idx = a.indices + (size_t)u * a.B + row0;
      if (role == 0) {
        // ------------------------------- target set: a' = actor_target(s'), Q'_h(s', a') -------------------------------
        const frl_net_t& AT = a.actor_target;
        const frl_net_t& CT = a.critic_target;
        fx_fetch(c, 0, AT, 0, 128);                // (issued by warps 4 - 7: warps 0 - 2 have the row gather to do)
        fx_fetch(c, 1, CT, l0, 192);
        trace(10);
        if (u == 0) fx_gather_async(rb.storage, rf, idx, nvalid, raw);      // later learns: prefetched in stage 2 of the previous one
        const bool smoothing = !sac && a.target_smoothing;
        // sampling noise and log_std do not depend on the forward passes: warps 4, 5 and 7 fetch / draw them now (a Philox +
        // Box-Muller draw or an L2 round trip each) instead of inside the sampling phase on the critical chain
        FRL_PAR(t) {
          const int q = t - 128;
          if (q >= 0 && q < 8 * ad) {
            const int r = q / ad, jj = q - r * ad;
            NZ[r * 8 + jj] = ((sac || smoothing) && r < nvalid) ? noise_at(a.noise_next, a, u, row0 + r, jj, 1u) : 0.f;
          }
          if (sac && t >= 224 && t < 224 + ad) LS[t - 224] = fx_ldcg(AT.p + AT.x_off + t - 224);
        }
        fx_gather_wait();
        trace(11);
        put_cols<8>(XA, aip, 0, raw, rf, rb_col_nobs(rb), od, aip);
        put_cols<8>(XN, cip, 0, raw, rf, rb_col_nobs(rb), od, cip);
        FRL_SYNC();
        trace(12);
        fx_wait(c, 0, 0);
        fx_fwd<0>(XA, aip, aip >> 2, slot0 + P.aoff[0], slot0 + P.aboff[0], 1, H1);
        fx_wait(c, 0, 1);
        fx_fwd<32>(H1, 128, 32, slot0 + P.aoff[1], slot0 + P.aboff[1], 1, H2);
        fx_fwd_narrow(H2, slot0 + P.aoff[2], slot0 + P.aboff[2], ap, ald, MU, 8);
        FRL_PAR(t) {
          if (t < 8 * ad) {
            const int r = t / ad, jj = t - r * ad;
            const float mean = MU[r * 8 + jj];
            float act;
            if (sac) {
              const float ls = fminf(fmaxf(LS[jj], -20.f), 2.f);
              const float sd = expf(ls);
              const float e = NZ[r * 8 + jj];
              const float uu = fadd(mean, fmul(e, sd));
              const float diff = uu - mean;
              float lp = -(diff * diff) / (2.f * (sd * sd)) - logf(sd) - FRL_LOG_SQRT_2PI;
              lp -= 2.f * (FRL_LOG2F - uu - softplus_t(-2.f * uu));
              UU[r * 8 + jj] = lp;
              act = tanhf(uu);
            } else if (smoothing) {
              const float e = NZ[r * 8 + jj];
              float nz = fmul(a.policy_noise_scale, fmul(e, a.policy_noise));
              nz = fminf(fmaxf(nz, -a.noise_clip), a.noise_clip);
              float v = fadd(fmul(tanhf(mean), a.max_action), nz);
              v = fminf(fmaxf(v, -a.max_action), a.max_action);
              act = fdiv(v, a.max_action);
            } else {
              act = tanhf(mean);
            }
            XN[r * cip + od + jj] = act;
          }
        }
        FRL_SYNC();
        trace(13);
        fx_wait(c, 1, 0);
        fx_fwd<0>(XN, cip, cip >> 2, slot1 + P.coff[0], slot1 + P.cboff[0], 1, H1);
        fx_wait(c, 1, 1);
        fx_fwd<32>(H1, 128, 32, slot1 + P.coff[1], slot1 + P.cboff[1], 1, H2);
        fx_fwd_narrow(H2, slot1 + P.coff[2], slot1 + P.cboff[2], 4, cld, QA, 4);
        FRL_PAR(t) {
          if (t < 8) {
            fx_pk_put(pk + (size_t)(row0 + t) * 3 + h, epoch, QA[t * 4]);
            if (sac && h == 0) {
              float lp = 0.f;
              for (int j = 0; j < ad; ++j) lp += UU[t * 8 + j];
              fx_pk_put(pk + (size_t)(row0 + t) * 3 + 2, epoch, lp);
            }
          }
        }
        trace(15);
        return;
      }
      // ------------------------------- online set: Q_h(s, a), pi(s); then y, loss, critic backward -------------------------------
      float* cx = ws + P.w_cx;
      float* cy = ws + P.w_cy;
      fx_fetch(c, 1, C, l0, 128);
      const bool do_actor = policy_step && h < hu;
      if (do_actor) fx_fetch(c, 0, A, 0, 192);
      trace(10);
      if (u == 0) fx_gather_async(rb.storage, rf, idx, nvalid, raw);
      FRL_PAR(t) {
        const int q = t - 128;
        if (do_actor && sac && q >= 0 && q < 64) {
          const int r = q >> 3, j = q & 7;
          NZ[q] = (j < ad && r < nvalid) ? noise_at(a.noise_new, a, u, row0 + r, j, 2u) : 0.f;
        }
        if (do_actor && sac && t >= 224 && t < 224 + ad) LS[t - 224] = fx_ldcg(A.p + A.x_off + t - 224);
        if (t == 255) alpha_s[0] = sac ? expf(fx_ldcg(a.alpha_state)) : 0.f;      // rewritten in stage 5 of the previous learn: read through L2
      }
      fx_gather_wait();
      trace(11);
      put_cols<8>(XS, cip, 0, raw, rf, 0, od + ad, cip);        // [obs | act] are adjacent in a replay row
      put_cols<8>(XA, aip, 0, raw, rf, 0, od, aip);
      FRL_SYNC();
      trace(12);
      // layer inputs X and pre-activation gradients dY of this tile go to the exchange blocks of head h's layers l0..l0+2 as they
      // are produced (the 128-wide ones from the GEMM epilogues)
      fx_put_blocks(cx, P.cxb[0], fx_xb(C, l0), Rm, row0, XS, cip, cip);
      const float* w0 = slot1 + P.coff[0];
      const float* w1 = slot1 + P.coff[1];
      const float* w2 = slot1 + P.coff[2];
      fx_wait(c, 1, 0);
      fx_fwd<0>(XS, cip, cip >> 2, w0, slot1 + P.cboff[0], 1, H1, cx + ((size_t)P.cxb[1] * Rm + row0) * 16, rm16);
      fx_wait(c, 1, 1);
      fx_fwd<32>(H1, 128, 32, w1, slot1 + P.cboff[1], 1, H2, cx + ((size_t)P.cxb[2] * Rm + row0) * 16, rm16);
      fx_fwd_narrow(H2, w2, slot1 + P.cboff[2], 4, cld, QA, 4);
      if (do_actor) {
        trace(16);
        float* ax = (h == 0) ? ws + P.w_ax : nullptr;           // head 0's CTA publishes the actor's layer inputs
        if (ax) fx_put_blocks(ax, P.axb[0], fx_xb(A, 0), Rm, row0, XA, aip, aip);
        fx_wait(c, 0, 0);
        fx_fwd<0>(XA, aip, aip >> 2, slot0 + P.aoff[0], slot0 + P.aboff[0], 1, A1, ax ? ax + ((size_t)P.axb[1] * Rm + row0) * 16 : nullptr, rm16);
        fx_wait(c, 0, 1);
        fx_fwd<32>(A1, 128, 32, slot0 + P.aoff[1], slot0 + P.aboff[1], 1, A2, ax ? ax + ((size_t)P.axb[2] * Rm + row0) * 16 : nullptr, rm16);
        fx_fwd_narrow(A2, slot0 + P.aoff[2], slot0 + P.aboff[2], ap, ald, MU, 8);
        FRL_PAR(t) {
          if (t < 64) {
            const int r = t >> 3, j = t & 7;
            float act = 0.f, e = 0.f, lp = 0.f;
            if (j < ad) {
              const float mean = MU[r * 8 + j];
              if (sac) {
                const float ls = fminf(fmaxf(LS[j], -20.f), 2.f);
                const float sd = expf(ls);
                e = NZ[t];
                const float uu = fadd(mean, fmul(e, sd));
                const float diff = uu - mean;
                lp = -(diff * diff) / (2.f * (sd * sd)) - logf(sd) - FRL_LOG_SQRT_2PI;
                lp -= 2.f * (FRL_LOG2F - uu - softplus_t(-2.f * uu));
                act = tanhf(uu);
              } else {
                act = tanhf(mean);
              }
            }
            AC[t] = act; UU[t] = lp; EPS[t] = e;
          }
        }
        // (no barrier: the next phase touches none of AC / UU / EPS, and several barriers precede their first reader in phase C)
      }
      // targets of this tile from the target set (both heads) -> y, loss, dL/dQ in one phase (threads 0..7 = the tile's rows)
      trace(17);
      FRL_PAR(t) {
        float l = 0.f;
        if (t < 8) {
          const int r = t;
          const float al = alpha_s[0];
          float y = 0.f;
          if (r < nvalid) {
            const unsigned long long* pr = pk + (size_t)(row0 + r) * 3;
            float nq, nq1, lp;
            fx_pk_wait3(pr, NH == 2 ? pr + 1 : nullptr, sac ? pr + 2 : nullptr, epoch, &nq, &nq1, &lp);
            if (NH == 2) nq = fminf(nq, nq1);
            const float rew = raw[r * rf + rb_col_rew(rb)], dn = raw[r * rf + rb_col_done(rb)];
            if (sac) {
              y = fadd(rew, fmul(fmul(a.gamma, fadd(1.f, -dn)), fadd(nq, fmul(al, -lp))));      // SAC.py:235
            } else {
              y = fadd(rew, fmul(fmul(a.gamma, nq), fadd(1.f, -dn)));                            // TD3.py:209 / DDPG.py:212
            }
          }
          for (int j = 0; j < 4; ++j) dQA[t * 4 + j] = 0.f;
          if (r < nvalid) {
            const float d0 = QA[t * 4] - y;
            dQA[t * 4] = 2.f * d0 * invB;
            l = d0 * d0;
          }
        }
        if (t < 32) red0[t] = l;
      }
      const float loss_c = fx_tree32(red0);
      FRL_SYNC();
      trace(18);
      fx_put_blocks(cy, P.cyb[2], 1, Rm, row0, dQA, 4, 4);
      fx_bwd_narrow(dQA, 4, w2, 4, cld, H2, D2, cy + ((size_t)P.cyb[1] * Rm + row0) * 16, rm16);
      fx_bwd(D2, w1, H1, D1, cy + ((size_t)P.cyb[0] * Rm + row0) * 16, rm16);
      FRL_PAR(t) { if (t == 0) ws[P.w_stats + (size_t)wi * 8 + 0] = loss_c; }
      trace(19);
    } else if (s == 1 || s == 4) {
      // dW of the critic (1) / the actor (4): ONE inlined copy of the stage body serves both (the kernel is ~180 KB of SASS and a
      // learn walks most of it once: instruction fetch is a visible part of every short phase)
      const bool cr = s == 1;
      const frl_net_t& N = cr ? C : A;
      res_invalidate(c, N);
      res_invalidate(c, cr ? a.critic_target : a.actor_target);
      res_drain_slot(c, cr ? 1 : 0);
      if (cr) c.stag1 = nullptr; else c.stag0 = nullptr;      // the slot is scratch in this stage and the next whatever it held
      long step;
      if (cr) step = (long)(a.step_critic0 + u + 1);
      else {
        const long n_policy_before = (a.policy_freq > 1) ? (long)((a.total_it0 + u) / a.policy_freq - a.total_it0 / a.policy_freq) : (long)u;
        step = (long)(a.step_actor0 + n_policy_before + 1);
      }
      const AdamSpec hp = {cr ? a.lr_critic : a.lr_actor, a.beta1, a.beta2, a.eps, cr ? a.wd_critic : 0.0, (double)a.max_norm, step};
      dw_stage(c, a, N, cr ? P.jc : P.ja, ws + (cr ? P.w_cx : P.w_ax), ws + (cr ? P.w_cy : P.w_ay0), (!cr && hu == 2) ? ws + P.w_ay1 : nullptr,
               cr ? c.wbuf1 : c.wbuf0, gst, gbst, redb, red0, sspk + (cr ? 0 : 512) + c.cta, epoch, ws + P.w_lsg, cr ? 0 : nwork, hp, hpst);
    } else if (s == 2 || s == 5) {
      const bool cr = s == 2;
      if (cr && role != 2 && u + 1 < a.n_updates)          // rows of the NEXT learn's tile (raw is dead after phase A)
        fx_gather_async(rb.storage, rf, a.indices + (size_t)(u + 1) * a.B + row0, nvalid, raw);
      const float tot = opt_stage(c, cr ? C : A, cr ? (policy_step ? &a.critic_target : nullptr) : &a.actor_target, a.tau, hpst, gst, gbst,
                                  sspk + (cr ? 0 : 512), epoch, c.red, cr ? P.jc : P.ja);
      if (metrics_cta && cr) {
        FRL_PAR(t) { red0[t] = t < nwork ? fx_ldcg(ws + P.w_stats + (size_t)t * 8) : 0.f; }
        FRL_SYNC();
        const float ls = block_sum(red0);
        FRL_PAR(t) {
          if (t == 0) {
            a.out[u * 8 + 0] = ls * invB; a.out[u * 8 + 4] = sqrtf(tot);
            a.out[u * 8 + 2] = sac ? expf(fx_ldcg(a.alpha_state)) : 0.f;
          }
        }
        FRL_SYNC();
      }
      if (metrics_cta && !cr) {
        FRL_PAR(t) {
          const bool on = t < nwork && (t % NH) < hu;
          red0[t] = on ? fx_ldcg(ws + P.w_stats + (size_t)t * 8 + 1) : 0.f;
          red1[t] = on ? fx_ldcg(ws + P.w_stats + (size_t)t * 8 + 2) : 0.f;
        }
        FRL_SYNC();
        const float l = block_sum(red0);
        const float en = block_sum(red1);
        FRL_PAR(t) {
          if (t == 0) {
            a.out[u * 8 + 1] = l * invB;
            a.out[u * 8 + 5] = sqrtf(tot);
            a.out[u * 8 + 6] = en * invB;
            if (sac && a.adaptive_alpha) {
              // alpha_loss = (exp(log_alpha) * (entropy - target_entropy).detach()).mean();  Adam(lr alpha_lr) on log_alpha
              const float mean_term = en * invB - a.target_entropy;
              const float al = expf(a.alpha_state[0]);
              const float g = al * mean_term;
              a.out[u * 8 + 3] = g;
              const AdamHP ha = adam_hp_ni(a.alpha_lr, a.beta1, a.beta2, a.eps, 0.0, 0.0, (long)(a.step_alpha0 + u + 1));
              float m = a.alpha_state[1], v = a.alpha_state[2], w = a.alpha_state[0];
              m = fmaf(ha.one_minus_b1, g - m, m);
              v = fadd(fmul(v, ha.b2), fmul(fmul(ha.one_minus_b2, g), g));
              const float denom = fadd(fdiv(fsqrt(v), ha.bc2_sqrt), ha.eps);
              w = fadd(w, fdiv(fmul(ha.lr_over_bc1_neg, m), denom));
              a.alpha_state[0] = w; a.alpha_state[1] = m; a.alpha_state[2] = v;
            }
          }
        }
        FRL_SYNC();
      }
    } else if (s == 3) {
      // ------------------------------- phase C: Q_h(s, pi(s)) with the updated critic, dQ/da, actor backward -------------------------------
      if (role == 0) { fx_fetch(c, 1, a.critic_target, l0, 128); return; }      // prefetch for the next learn
      if (role != 1) return;
      fx_fetch(c, 1, C, l0, 128);
      if (h >= hu) return;
      const float alpha = sac ? alpha_s[0] : 0.f;
      put_cols<8>(XN, cip, 0, XS, cip, 0, od, od);
      FRL_PAR(t) {
        if (t < 8 * (cip - od)) {
          const int r = t / (cip - od), j = t - r * (cip - od);
          XN[r * cip + od + j] = j < ad ? AC[r * 8 + j] : 0.f;
        }
      }
      FRL_SYNC();
      trace(30);
      const float* w0 = slot1 + P.coff[0];
      const float* w1 = slot1 + P.coff[1];
      const float* w2 = slot1 + P.coff[2];
      fx_wait(c, 1, 0);
      fx_fwd<0>(XN, cip, cip >> 2, w0, slot1 + P.cboff[0], 1, H1);
      fx_wait(c, 1, 1);
      fx_fwd<32>(H1, 128, 32, w1, slot1 + P.cboff[1], 1, H2);
      fx_fwd_narrow(H2, w2, slot1 + P.cboff[2], 4, cld, QA, 4);
      const float dq = -invB / (float)hu;
      FRL_PAR(t) {
        float v = 0.f;
        if (t < 8) {
          for (int j = 0; j < 4; ++j) dQA[t * 4 + j] = 0.f;
          if (t < nvalid) { dQA[t * 4] = dq; v = QA[t * 4]; }
        }
        if (t < 32) red0[t] = v;
      }
      const float qsum = fx_tree32(red0);
      FRL_SYNC();
      trace(31);
      fx_bwd_narrow(dQA, 4, w2, 4, cld, H2, D2);
      fx_bwd(D2, w1, H1, D1);
      fx_bwd_cols(D1, w0, od, ad, dXa);
      trace(32);
      FRL_PAR(t) {
        float lsum = 0.f, esum = 0.f;
        if (t < 8) {
          const int r = t;
          float lp = 0.f;
          for (int j = 0; j < 8; ++j) {
            float g = 0.f;
            if (j < ad && r < nvalid) {
              const float act = AC[r * 8 + j];
              g = dXa[r * 8 + j] * (1.f - act * act);
              if (sac && h == 0) { g += alpha * invB * 2.f * act; lp += UU[r * 8 + j]; }
            }
            dMU[r * 8 + j] = g;
          }
          if (r < nvalid && h == 0) { esum = -lp; lsum = alpha * lp; }        // actor_loss = mean(-Q_pi - alpha * entropy)
        }
        if (t < 32) { red0[t] = lsum; red1[t] = esum; }
      }
      const float loss_a = fx_tree32(red0) - qsum / (float)hu;
      const float ent = fx_tree32(red1);
      FRL_SYNC();
      if (sac) {
        // d/dlog_std_j = sum_r dL/du std eps (+ head 0: -alpha / B per row); zero outside the clamp range
        FRL_PAR(t) {
          if (t < 8) {
            const float lsr = (t < ad) ? LS[t] : 1e30f;
            float g = 0.f;
            if (lsr >= -20.f && lsr <= 2.f) {
              const float sd = expf(lsr);
              for (int r = 0; r < nvalid; ++r) g += dMU[r * 8 + t] * sd * EPS[r * 8 + t] - (h == 0 ? alpha * invB : 0.f);
            }
            ws[P.w_lsg + (size_t)wi * 8 + t] = g;
          }
        }
      }
      trace(33);
      float* ay = ws + (h == 0 ? P.w_ay0 : P.w_ay1);
      fx_put_blocks(ay, P.ayb[2], 1, Rm, row0, dMU, 8, ap);
      const float* p2 = slot0 + P.aoff[2];
      const float* p1 = slot0 + P.aoff[1];
      fx_bwd_narrow(dMU, 8, p2, ap, ald, A2, D2, ay + ((size_t)P.ayb[1] * Rm + row0) * 16, rm16);
      fx_bwd(D2, p1, A1, D1, ay + ((size_t)P.ayb[0] * Rm + row0) * 16, rm16);
      trace(34);
      FRL_PAR(t) { if (t == 0) { ws[P.w_stats + (size_t)wi * 8 + 1] = loss_a; ws[P.w_stats + (size_t)wi * 8 + 2] = ent; } }
      trace(39);
    }
  }
};

// ------------------------------------------------------------------------------------------------------------------------
// launcher: cooperative launch (co-residency guarantee), hand-rolled grid barrier
// ------------------------------------------------------------------------------------------------------------------------
#ifndef FRL_EMUL
template <class A>
__global__ void __launch_bounds__(FRL_NT, 1) frl_fx_kernel(const __grid_constant__ typename A::Args a) {
  extern __shared__ __align__(1024) float frl_smem[];
  // The argument block (network / replay descriptors, ~3 KB) is copied to shared memory once: read from the constant bank, its
  // scattered fields cost a constant-cache miss each at the head of every stage (measured: 2 - 5 k clk of scalar preamble per
  // stage, profiles/r2d_trace_64.txt), and every stage of every learn re-reads them.
  __shared__ __align__(16) typename A::Args sa;
  {
    const int* src = reinterpret_cast<const int*>(&a);
    int* dst = reinterpret_cast<int*>(&sa);
    for (int i = (int)threadIdx.x; i < (int)(sizeof(typename A::Args) / 4); i += FRL_NT) dst[i] = src[i];
  }
  Cta c;
  float* user = cta_init(c, (int)blockIdx.x, (int)gridDim.x, frl_smem, A::wbuf_floats(a));      // ends with __syncthreads
  const int U = A::n_updates(a);
  unsigned* const ctr = a.sync;
  unsigned target = 0;
  for (int u = 0; u < U; ++u) {
    for (int s = 0; s < A::NSTAGES; ++s) {
      if (!A::stage_enabled(s, u, sa)) continue;
      trace(1000 + s);
      A::stage(s, u, c, user, sa);
      trace(1100 + s);
      stamp(c, 100 + s);
      // (no proxy fence here: the thread that issues a TMA copy of data other CTAs wrote runs fence.proxy.async after this
      //  barrier's acquire, which puts the fence on the causality path between the generic writes and the bulk read)
      if (A::barrier_after(s)) {
        target += gridDim.x;
        fx_grid_barrier(ctr, target);
      }
      stamp(c, 200 + s);
    }
  }
  res_drain(c);
}

template <class A>
int frl_launch_fx(const typename A::Args& a, cudaStream_t stream) {
  const int smem_bytes = (cta_base_floats(A::wbuf_floats(a)) + A::user_floats(a)) * 4 + 64;      // + the static copy of the arguments
  int dev = 0;
  FRL_CUDA_OK(cudaGetDevice(&dev));
  static int configured_bytes[64] = {0};
  if (dev < 0 || dev >= 64) { frl_set_error("device ordinal %d out of range", dev); return -3; }
  if (smem_bytes > configured_bytes[dev]) {
    FRL_CUDA_OK(cudaFuncSetAttribute(frl_fx_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured_bytes[dev] = smem_bytes;
  }
  const int grid = A::grid(a, frl_device_max_ctas());
  int per_sm = 0;
  FRL_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frl_fx_kernel<A>, FRL_NT, (size_t)smem_bytes));
  if (per_sm * frl_device_max_ctas() < grid) {
    frl_set_error("cooperative launch of %d CTAs does not fit the device (%d per SM)", grid, per_sm);
    return -3;
  }
  FRL_CUDA_OK(cudaMemsetAsync(a.sync, 0, FX_SYNC_WORDS * 4, stream));       // barrier counter + hand-off packets
  typename A::Args args = a;
  void* kargs[] = {(void*)&args};
  FRL_CUDA_OK(cudaLaunchCooperativeKernel((void*)frl_fx_kernel<A>, dim3(grid), dim3(FRL_NT), kargs, (size_t)smem_bytes, stream));
  ++frl_launch_counter;
  return 0;
}
#else
template <class A>
int frl_launch_fx(const typename A::Args& a, cudaStream_t s) {
  memset(a.sync, 0, FX_SYNC_WORDS * 4);
  return frl_launch<A>(a, s);
}
#endif
