// Large-minibatch PPO / MAPPO forward + backward on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 error
// compensation, fp32 accumulators in TMEM) — stage "fwd/bwd" of frl_ppo_update when a minibatch has >= 1024 rows and both
// networks are in -> 128 -> 128 -> out MLPs (C3 PPO LunarLander: 8192-row minibatches; C5 MAPPO: 131 072-row full batches).
//   PPO.learn minibatch loop   PPO_file/PPO.py:245-283     Agent.update_ac_   PPO_file/PPO.py:145-152
//   MAPPO.learn                MAPPO_file/MAPPO.py:392-436 (LayerNorm nets: MAPPO.py:127-218)
//
// Work split: the first half of the grid owns the actor, the second half the critic; a CTA walks 128-row tiles of the
// minibatch.  Per tile ("phase A") the activations never touch shared memory:
//     x (registers) -> TMEM (hi | lo columns) --MMA(W1)--> TMEM acc --tcgen05.ld, +b, ReLU [, LayerNorm]--> TMEM (hi | lo) --MMA(W2)--> ...
// the A operand of every forward / backward-dX GEMM is read from TMEM, the B operand (pre-split hi / lo weights in UMMA
// layout Q, written once per update by the split stage) streams through a 4 x 32 KB shared-memory ring of bulk copies that
// runs one GEMM ahead of the epilogues.  The epilogues also leave hi / lo copies of x, h1, h2 and of the three dZ in a
// per-CTA global scratch (L2-resident, UMMA layout S); "phase B" streams them back in 32-row chunks and forms
//     dW2 += dZ2^T h1 (kept in TMEM across all tiles of the CTA),  dW1 += dZ1^T x,  dW3^T += h2^T dZ3,  db1 / db2 = dZ^T 1
// with MN-major operands.  One gradient partial per (net, CTA) goes to gpart; the reduce / clip / optimiser stages of
// algo_ppo.cuh are unchanged.  Algorithmic bytes per 128-row tile and net: ~0.3 MB of weights + ~0.6 MB of scratch written
// and read back, all L2 hits; HBM sees the gathered rollout rows only.
#pragma once
#ifndef FRL_EMUL
#include "umma.cuh"

#define UM_SLOT_F 8192                  // floats per ring slot: hi 16 KB | lo 16 KB
#define UM_WS_LAYER 65536               // split-weight block per layer: fwd hi | fwd lo | bwd hi | bwd lo, 16384 floats each
#define UM_WS_CTA 155648                // per-CTA activation scratch (floats): X 2x8192, H1 H2 DZ1 DZ2 2x16384 each, DZ3 2x4096
#define UM_USER_FLOATS 50432            // shared memory the fwd/bwd stage carves from `user` (incl. 1 KB alignment slack)

FRL_HD int um_pad16(int x) { return (x + 15) & ~15; }
FRL_HD int um_pad32(int x) { return (x + 31) & ~31; }

// host + device: can this update take the tensor-core path?
FRL_HD bool um_eligible(const frl_ppo_args_t& a) {
  if (!a.umma_ws || a.mb < 1024 || a.hidden_tanh || a.net.n_layers != 6 || a.continuous == 2) return false;
  for (int r = 0; r < 2; ++r) {
    const frl_layer_t &L0 = a.net.L[3 * r], &L1 = a.net.L[3 * r + 1], &L2 = a.net.L[3 * r + 2];
    if (L0.out != 128 || L1.in != 128 || L1.out != 128 || L2.in != 128 || L0.in > 64 || L2.out > 16) return false;
    if (L1.in_pad != 128 || L2.in_pad != 128) return false;
  }
  return a.n_adv <= 16 && a.act_cols <= 16 && a.logp_cols <= 16;
}
FRL_HD int um_grid(const frl_ppo_args_t& a, int max_ctas) {
  const int want = 2 * ((a.mb + 127) / 128), cap = max_ctas & ~1;
  return want < cap ? want : cap;
}
// gradient partials (and loss partials) that stage "reduce" has to sum for a minibatch of `rows` rows
FRL_HD int um_ncontrib(int rows, int ncta) {
  const int ntile = (rows + 127) / 128, nper = ncta >> 1;
  return ntile < nper ? ntile : nper;
}

// ---- stage "split": fp32 weights -> hi / lo TF32 pairs in UMMA layout Q, forward (W[n][k]: rows n, K = k) and backward-dX
// (W^T[k][n]: rows k, K = n) operands of every layer -------------------------------------------------------------------
__device__ __noinline__ void um_split(const Cta& c, const frl_ppo_args_t& a) {
  const frl_net_t& N = a.net;
  const int tid = (int)threadIdx.x;
  for (int li = 0; li < 6; ++li) {
    const frl_layer_t& L = N.L[li];
    const int j = li % 3;
    float* ws = a.umma_ws + (size_t)li * UM_WS_LAYER;
    const int Rf = j == 2 ? 16 : 128, Cf = j == 0 ? um_pad16(L.in) : 128;
    for (int e = c.cta * FRL_NT + tid; e < Rf * Cf; e += c.ncta * FRL_NT) {
      const int n = e / Cf, k = e % Cf;
      const float x = (n < L.out && k < L.in) ? N.p[L.w_off + n * L.in_pad + k] : 0.f;
      const float hi = um_hi(x);
      const int o = um_q_off(n, k, Rf);
      ws[o] = hi;
      ws[16384 + o] = x - hi;
    }
    if (j != 0) {
      const int Cb = j == 2 ? 16 : 128;            // rows k = 128 input features, K = n (output features, padded)
      for (int e = c.cta * FRL_NT + tid; e < 128 * Cb; e += c.ncta * FRL_NT) {
        const int n = e / 128, k = e % 128;
        const float x = (n < L.out && k < L.in) ? N.p[L.w_off + n * L.in_pad + k] : 0.f;
        const float hi = um_hi(x);
        const int o = um_q_off(k, n, 128);
        ws[32768 + o] = hi;
        ws[49152 + o] = x - hi;
      }
    }
  }
}

// one updated parameter -> its hi / lo copies in the split-weight blocks (called by the optimiser stages, so only the first update of
// a launch needs the split stage)
UM_DEV void um_split_one(const frl_ppo_args_t& a, int p, float w) {
  const frl_net_t& N = a.net;
  for (int li = 0; li < 6; ++li) {
    const frl_layer_t& L = N.L[li];
    if (p >= L.w_off && p < L.w_off + L.out_pad * L.in_pad) {
      const int e = p - L.w_off, n = e / L.in_pad, k = e % L.in_pad;
      if (n >= L.out || k >= L.in) return;
      const int j = li % 3;
      float* ws = a.umma_ws + (size_t)li * UM_WS_LAYER;
      const float hi = um_hi(w), lo = w - hi;
      const int o = um_q_off(n, k, j == 2 ? 16 : 128);
      ws[o] = hi;
      ws[16384 + o] = lo;
      if (j != 0) {
        const int ob = um_q_off(k, n, 128);
        ws[32768 + ob] = hi;
        ws[49152 + ob] = lo;
      }
      return;
    }
  }
}

// 4 x 4 transpose of float4 pieces among the four lanes of an aligned lane group (butterfly, an involution): afterwards
// v[i] of lane l = the former v[l & 3] of lane (l & ~3) + i
UM_DEV void um_xpose4(float4* v, int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
#pragma unroll
  for (int p = 0; p < 4; p += 2) {
    float4 s_ = b0 ? v[p] : v[p + 1], r_;
    r_.x = __shfl_xor_sync(0xffffffffu, s_.x, 1); r_.y = __shfl_xor_sync(0xffffffffu, s_.y, 1);
    r_.z = __shfl_xor_sync(0xffffffffu, s_.z, 1); r_.w = __shfl_xor_sync(0xffffffffu, s_.w, 1);
    if (b0) v[p] = r_; else v[p + 1] = r_;
  }
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    float4 s_ = b1 ? v[p] : v[p + 2], r_;
    r_.x = __shfl_xor_sync(0xffffffffu, s_.x, 2); r_.y = __shfl_xor_sync(0xffffffffu, s_.y, 2);
    r_.z = __shfl_xor_sync(0xffffffffu, s_.z, 2); r_.w = __shfl_xor_sync(0xffffffffu, s_.w, 2);
    if (b1) v[p] = r_; else v[p + 2] = r_;
  }
}

// 16 consecutive columns (c0 ..) of this thread's row -> hi / lo copies in the layout-S scratch.  A thread owns a ROW (TMEM lane), and
// rows are 128 B apart in layout S: storing its own 16-byte pieces would touch 32 lines per warp instruction (measured: the stores,
// not the MMAs, bounded the epilogues).  The four lanes of a row group first transpose their pieces, so that every instruction
// writes, per group, the 64 contiguous bytes of ONE row (8 lines per instruction).  All 32 lanes of the warp must call it.
UM_DEV void um_s_store16(float* bh, float* bl, int row, int c0, int C, const float* x) {
  const int lane = (int)threadIdx.x & 31, l = lane & 3, rb = row & ~3;
  float4 v[4] = {make_float4(x[0], x[1], x[2], x[3]), make_float4(x[4], x[5], x[6], x[7]), make_float4(x[8], x[9], x[10], x[11]),
                 make_float4(x[12], x[13], x[14], x[15])};
  um_xpose4(v, lane);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = um_s_off(rb + i, c0 + 4 * l, C);
    const float4 h = make_float4(um_hi(v[i].x), um_hi(v[i].y), um_hi(v[i].z), um_hi(v[i].w));
    st4(bh + o, h);
    st4(bl + o, make_float4(v[i].x - h.x, v[i].y - h.y, v[i].z - h.z, v[i].w - h.w));
  }
}
// the inverse: 16 columns of this thread's row, hi + lo (= the exact fp32 value), read back with the same coalescing
UM_DEV void um_s_load16(const float* bh, const float* bl, int row, int c0, int C, float* y) {
  const int lane = (int)threadIdx.x & 31, l = lane & 3, rb = row & ~3;
  float4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = um_s_off(rb + i, c0 + 4 * l, C);
    const float4 h = ld4(bh + o), lo = ld4(bl + o);
    v[i] = make_float4(h.x + lo.x, h.y + lo.y, h.z + lo.z, h.w + lo.w);
  }
  um_xpose4(v, lane);
#pragma unroll
  for (int i = 0; i < 4; ++i) { y[4 * i] = v[i].x; y[4 * i + 1] = v[i].y; y[4 * i + 2] = v[i].z; y[4 * i + 3] = v[i].w; }
}

// ring + barriers, driven by thread 0 only
struct UmPipe {
  float* ring;
  uint64_t *full, *empty, *acc;
  uint32_t tm;
  uint32_t head, tail;
};
UM_DEV void um_fill(UmPipe& p, const float* hi, const float* lo, uint32_t bytes) {
  const uint32_t s = p.head & 3u, use = p.head >> 2;
  if (use > 0) um_mbar_wait(&p.empty[s], (use - 1) & 1u);
  um_mbar_expect_tx(&p.full[s], lo ? 2 * bytes : bytes);
  um_bulk_g2s(p.ring + s * UM_SLOT_F, hi, bytes, &p.full[s]);
  if (lo) um_bulk_g2s(p.ring + s * UM_SLOT_F + 4096, lo, bytes, &p.full[s]);
  p.head++;
}
// D[128 x R] (TMEM column dcol) (+)= A (TMEM hi at column 128 + acol, lo at 256 + acol; kc K-columns) . B (next ring slot, layout Q with R rows)
UM_DEV void um_consume_ts(UmPipe& p, uint32_t dcol, uint32_t acol, int R, int kc, bool first) {
  const uint32_t s = p.tail & 3u, use = p.tail >> 2;
  um_mbar_wait(&p.full[s], use & 1u);
  um_fence_after();
  const uint32_t sb = um_smem_u32(p.ring + s * UM_SLOT_F);
  const uint64_t bh = um_desc_q(sb, R), bl = um_desc_q(sb + 16384u, R);
  const uint32_t idesc = um_idesc_tf32(128, R, 0, 0);
  for (int ks = 0; ks < kc / 8; ++ks) {
    const uint64_t adv = (uint64_t)((uint32_t)(ks * 32 * R) >> 4);
    um_mma_ts(p.tm + dcol, p.tm + 256 + acol + ks * 8, bh + adv, idesc, (first && ks == 0) ? 0u : 1u);
    um_mma_ts(p.tm + dcol, p.tm + 128 + acol + ks * 8, bl + adv, idesc, 1u);
    um_mma_ts(p.tm + dcol, p.tm + 128 + acol + ks * 8, bh + adv, idesc, 1u);
  }
  um_commit(&p.empty[s]);
  p.tail++;
}
// one 32-row chunk of a dW GEMM: A = ring slot (layout S, CA columns, M = 128 of them), B = next slot (layout S, CB columns,
// first NB used) -> D[128 x NB] at TMEM column dcol
UM_DEV void um_consume_ss(UmPipe& p, uint32_t dcol, int CA, int CB, int NB, bool first) {
  const uint32_t sa = p.tail & 3u, ua = p.tail >> 2, sb_ = (p.tail + 1) & 3u, ub = (p.tail + 1) >> 2;
  um_mbar_wait(&p.full[sa], ua & 1u);
  um_mbar_wait(&p.full[sb_], ub & 1u);
  um_fence_after();
  const uint32_t aa = um_smem_u32(p.ring + sa * UM_SLOT_F), ba = um_smem_u32(p.ring + sb_ * UM_SLOT_F);
  const uint64_t ah = um_desc_s(aa, CA), al = um_desc_s(aa + 16384u, CA), bh = um_desc_s(ba, CB), bl = um_desc_s(ba + 16384u, CB);
  const uint32_t idesc = um_idesc_tf32(128, NB, 1, 1);
  for (int ks = 0; ks < 4; ++ks) {
    const uint64_t aadv = (uint64_t)((uint32_t)(ks * 32 * CA) >> 4), badv = (uint64_t)((uint32_t)(ks * 32 * CB) >> 4);
    um_mma_ss(p.tm + dcol, al + aadv, bh + badv, idesc, (first && ks == 0) ? 0u : 1u);
    um_mma_ss(p.tm + dcol, ah + aadv, bl + badv, idesc, 1u);
    um_mma_ss(p.tm + dcol, ah + aadv, bh + badv, idesc, 1u);
  }
  um_commit(&p.empty[sa]);
  um_commit(&p.empty[sb_]);
  p.tail += 2;
}

// per-thread view of the stage: which TMEM lanes / columns the thread owns, the row-statistics exchange buffer
struct UmThr {
  uint32_t tm, lane_base;
  int row, hh, tid;
  float* xs;            // [4][2][128] partial row sums of the two half-row threads
};

// Forward epilogue of a hidden layer: acc -> +bias -> ReLU [-> F.layer_norm over the 128 features, MAPPO.py:145-151] -> TMEM A
// operand (hi | lo) + layout-S scratch.  Returns the ReLU mask of the thread's 64 columns; *rstd = 1 / sqrt(var + 1e-5) of the row.
template <int NO>      // NO = 0: no output layer; 4 / 8 / 16: register accumulators for that many output columns
UM_DEV uint64_t um_hidden_fwd(const UmThr& th, const float* bias, bool ln, float* Sh, float* Sl, float* rstd, const float* W3s = nullptr,
                              float* lgp = nullptr) {
  uint64_t mask = 0;
  float mean = 0.f, rs = 1.f;
  if (ln) {
    float s = 0.f;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int c0 = th.hh * 64 + j * 16;
      float v[16];
      um_ld16(th.tm + th.lane_base + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) s += fmaxf(v[i] + bias[c0 + i], 0.f);
    }
    th.xs[th.hh * 128 + th.row] = s;
    __syncthreads();
    mean = (th.xs[th.row] + th.xs[128 + th.row]) * (1.0f / 128.f);
    float q = 0.f;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int c0 = th.hh * 64 + j * 16;
      float v[16];
      um_ld16(th.tm + th.lane_base + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float d = fmaxf(v[i] + bias[c0 + i], 0.f) - mean; q += d * d; }
    }
    th.xs[256 + th.hh * 128 + th.row] = q;
    __syncthreads();
    rs = 1.0f / sqrtf((th.xs[256 + th.row] + th.xs[384 + th.row]) * (1.0f / 128.f) + 1e-5f);
  }
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    const int c0 = th.hh * 64 + j * 16;
    float v[16], hi[16], lo[16];
    trace(2100 + j, 32);
    um_ld16(th.tm + th.lane_base + c0, v);
    trace(2110 + j, 32);
#pragma unroll
    uint32_t m16 = 0;
    for (int i = 0; i < 16; ++i) {
      float x = v[i] + bias[c0 + i];
      if (x > 0.f) m16 |= 1u << i; else x = 0.f;
      if (ln) x = (x - mean) * rs;
      hi[i] = um_hi(x);
      lo[i] = x - hi[i];
      v[i] = x;
    }
    mask |= (uint64_t)m16 << (j * 16);
    trace(2120 + j, 32);
    um_st16(th.tm + th.lane_base + 128 + c0, hi);
    um_st16(th.tm + th.lane_base + 256 + c0, lo);
    trace(2130 + j, 32);
    um_s_store16(Sh, Sl, th.row, c0, 128, v);
    trace(2140 + j, 32);
    if (NO > 0) {                              // output layer on the CUDA cores (out <= 16): partial dot products over this thread's columns
#pragma unroll
      for (int n = 0; n < NO; ++n) {
        float acc = lgp[n];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc = fmaf(v[i], W3s[n * 128 + c0 + i], acc);
        lgp[n] = acc;
      }
    }
  }
  *rstd = rs;
  return mask;
}

// Backward epilogue of a hidden layer: acc = dL/d(layer output) -> [layer-norm backward with Y = the normalised activations read
// back from the scratch (hi + lo is the exact fp32 value): dX = rstd (dY - mean(dY) - Y mean(dY Y))] -> ReLU mask -> dZ, written
// to the layout-S scratch and (to_tmem) as the next GEMM's A operand.
// Column sums of a [32 lanes x 16] register tile by recursive halving (16 shuffles): afterwards every lane holds the sum over the
// warp's 32 rows of column um_colsum_col(lane).
UM_DEV float um_colsum16(const float* x, int lane) {
  float a8[8], a4[4], a2[2], a1;
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) a8[i] = (b4 ? x[8 + i] : x[i]) + __shfl_xor_sync(0xffffffffu, b4 ? x[i] : x[8 + i], 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) a4[i] = (b3 ? a8[4 + i] : a8[i]) + __shfl_xor_sync(0xffffffffu, b3 ? a8[i] : a8[4 + i], 8);
#pragma unroll
  for (int i = 0; i < 2; ++i) a2[i] = (b2 ? a4[2 + i] : a4[i]) + __shfl_xor_sync(0xffffffffu, b2 ? a4[i] : a4[2 + i], 4);
  a1 = (b1 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, b1 ? a2[0] : a2[1], 2);
  return a1 + __shfl_xor_sync(0xffffffffu, a1, 1);
}
UM_DEV int um_colsum_col(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

// dY of 16 columns: from the TMEM accumulator, or (W3s given) = dz . W3 on the CUDA cores
UM_DEV void um_dy16(const UmThr& th, int c0, const float* W3s, const float* dz, int nout, float* v) {
  if (!W3s) { um_ld16(th.tm + th.lane_base + c0, v); return; }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
#pragma unroll 1
  for (int n = 0; n < nout; ++n) {           // dz: this row's dZ3 in shared memory
    const float d = dz[n];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaf(d, W3s[n * 128 + c0 + i], v[i]);
  }
}

UM_DEV void um_hidden_bwd(const UmThr& th, uint64_t mask, bool ln, float rs, const float* Yh, const float* Yl, float* Dh, float* Dl,
                          bool to_tmem, float* dbp, const float* W3s = nullptr, const float* dz = nullptr, int nout = 0) {
  float m1 = 0.f, m2 = 0.f;
  if (ln) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int c0 = th.hh * 64 + j * 16;
      float v[16];
      um_dy16(th, c0, W3s, dz, nout, v);
      float y[16];
      um_s_load16(Yh, Yl, th.row, c0, 128, y);
#pragma unroll
      for (int i = 0; i < 16; ++i) { s1 += v[i]; s2 += v[i] * y[i]; }
    }
    th.xs[512 + th.hh * 128 + th.row] = s1;
    th.xs[768 + th.hh * 128 + th.row] = s2;
    __syncthreads();
    m1 = (th.xs[512 + th.row] + th.xs[640 + th.row]) * (1.0f / 128.f);
    m2 = (th.xs[768 + th.row] + th.xs[896 + th.row]) * (1.0f / 128.f);
  }
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    const int c0 = th.hh * 64 + j * 16;
    float v[16], hi[16], lo[16];
    um_dy16(th, c0, W3s, dz, nout, v);
    if (ln) {
      float y[16];
      um_s_load16(Yh, Yl, th.row, c0, 128, y);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = rs * (v[i] - m1 - y[i] * m2);
    }
#pragma unroll
    const uint32_t m16 = (uint32_t)(mask >> (j * 16));
    for (int i = 0; i < 16; ++i) {
      const float x = ((m16 >> i) & 1u) ? v[i] : 0.f;
      hi[i] = um_hi(x);
      lo[i] = x - hi[i];
      v[i] = x;
    }
    {                                          // bias gradient: column sums over the warp's 32 rows, accumulated per TMEM lane quarter
      const int lane = th.tid & 31;
      const float cs_ = um_colsum16(v, lane);
      if ((lane & 1) == 0) dbp[((th.tid >> 5) & 3) * 128 + c0 + um_colsum_col(lane)] += cs_;
    }
    if (to_tmem) {
      um_st16(th.tm + th.lane_base + 128 + c0, hi);
      um_st16(th.tm + th.lane_base + 256 + c0, lo);
    }
    um_s_store16(Dh, Dl, th.row, c0, 128, v);
  }
}

// ---- stage "fwd/bwd" -----------------------------------------------------------------------------------------------
__device__ __noinline__ void ppo_umma_stage(Cta& c, float* user, const frl_ppo_args_t& a, int u) {
  const frl_net_t& N = a.net;
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31, q = warp & 3, hh = warp >> 2;
  const int row = q * 32 + lane;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const int rows = a.mb_rows[u];
  const int ntile = (rows + 127) / 128, nper = c.ncta >> 1;
  const int role = c.cta >= nper ? 1 : 0;                  // 0 = actor CTA, 1 = critic CTA
  const int ci = role ? c.cta - nper : c.cta;
  if (ci >= um_ncontrib(rows, c.ncta) || c.cta >= 2 * nper) return;
  const int l0 = 3 * role;
  const frl_layer_t &L0 = N.L[l0], &L1 = N.L[l0 + 1], &L2 = N.L[l0 + 2];
  const int in = L0.in, nout = L2.out;
  const int K0 = um_pad16(in), Cx = um_pad32(in);

  float* ring = (float*)(((uintptr_t)user + 1023) & ~(uintptr_t)1023);
  float* W3s = ring + 4 * UM_SLOT_F;           // [16][128] fp32 output-layer weights (rows >= out are zero)
  float* accW1 = W3s + 2048;                   // [64][128]  dW1[n][k] at [k][n]
  float* accW3 = accW1 + 64 * 128;             // [16][128]  dW3[n][k] at [n][k]
  float* accB = accW3 + 16 * 128;              // [2][4][128] db1 | db2 partial sums per TMEM lane quarter
  float* accX = accB + 1024;                   // [32]       db3[16] | dlog_std[16]
  float* bias = accX + 32;                     // [3][128]
  float* cs = bias + 384;                      // [128][17] column-sum scratch
  float* red = cs + 128 * 17;                  // [256]
  float* xs = red + 256;                       // [4][2][128] row-statistics exchange (LayerNorm)
  uint64_t* bars = (uint64_t*)(xs + 1024);     // full[4] | empty[4] | acc
  uint32_t* tslot = (uint32_t*)(bars + 9);

  for (int i = tid; i < 64 * 128 + 16 * 128 + 1024 + 32; i += FRL_NT) accW1[i] = 0.f;
  for (int i = tid; i < 2048; i += FRL_NT) W3s[i] = (i >> 7) < nout ? N.p[L2.w_off + (i >> 7) * L2.in_pad + (i & 127)] : 0.f;
  for (int i = tid; i < 384; i += FRL_NT) {
    const frl_layer_t& L = N.L[l0 + i / 128];
    bias[i] = (i % 128) < L.out ? N.p[L.b_off + (i % 128)] : 0.f;
  }
  float* dzs = c.red;                          // [128][17] dZ3 rows (engine scratch, unused by this stage otherwise)
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) um_mbar_init(&bars[i], 1);
    um_fence_mbar_init();
  }
  um_fence_proxy_async();
  if (warp == 0) um_tmem_alloc<512>(tslot);
  um_fence_before();
  __syncthreads();
  um_fence_after();

  stamp(c, 300);
  UmPipe p;
  p.ring = ring; p.full = bars; p.empty = bars + 4; p.acc = bars + 8; p.tm = *tslot; p.head = p.tail = 0;
  const uint32_t tm = p.tm;
  uint32_t accn = 0;                           // accumulator-barrier uses so far (all threads keep the count)
  const bool ln = a.layer_norm != 0;
  UmThr th;
  th.tm = tm; th.lane_base = lane_base; th.row = row; th.hh = hh; th.tid = tid; th.xs = xs;

  const float* wsl = a.umma_ws;
  float* act = a.umma_ws + (size_t)6 * UM_WS_LAYER + (size_t)c.cta * UM_WS_CTA;
  float *Xh = act, *Xl = act + 8192, *H1h = act + 16384, *H1l = act + 32768, *H2h = act + 49152, *H2l = act + 65536;
  float *D1h = act + 81920, *D1l = act + 98304, *D2h = act + 114688, *D2l = act + 131072, *D3h = act + 147456, *D3l = act + 151552;
#define UM_W(li, w) (wsl + (size_t)(li) * UM_WS_LAYER + (size_t)(w) * 16384)

  float la = 0.f, lc = 0.f, le = 0.f;
  const float inv_rows = 1.0f / (float)rows, inv_rn = 1.0f / (float)(rows * a.n_adv);
  int it = 0;
  for (int tile = ci; tile < ntile; tile += nper, ++it) {
    const int row0 = tile * 128;
    const int nvalid = (rows - row0) < 128 ? (rows - row0) : 128;
    const bool valid = row < nvalid;
    const int64_t gi = valid ? a.indices[(size_t)u * a.mb + row0 + row] : 0;
    if (tid == 0)
      for (int c0 = 0; c0 < K0; c0 += 32) {
        const int kc = (K0 - c0) < 32 ? (K0 - c0) : 32;
        um_fill(p, UM_W(l0, 0) + c0 * 128, UM_W(l0, 1) + c0 * 128, (uint32_t)kc * 512u);
      }
    // ---- inputs: x -> TMEM A columns + scratch X ----
    if (hh == 0) {
      const float* src = (role && a.critic_obs) ? a.critic_obs + (size_t)gi * a.critic_obs_dim : a.obs + (size_t)gi * a.obs_dim;
      float mean = 0.f, rs = 1.f;
      if (ln && valid) {                       // feature_norm: F.layer_norm over the `in` input features (MAPPO.py:145)
        float s_ = 0.f, q_ = 0.f;
        for (int k = 0; k < in; ++k) s_ += src[k];
        mean = s_ / (float)in;
        for (int k = 0; k < in; ++k) { const float d = src[k] - mean; q_ += d * d; }
        rs = 1.0f / sqrtf(q_ / (float)in + 1e-5f);
      }
      for (int c0 = 0; c0 < K0; c0 += 16) {
        float xv[16], hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = (valid && c0 + i < in) ? src[c0 + i] : 0.f;
          if (ln && valid && c0 + i < in) x = (x - mean) * rs;
          xv[i] = x;
          hi[i] = um_hi(x);
          lo[i] = x - hi[i];
        }
        um_st16(tm + lane_base + 128 + c0, hi);
        um_st16(tm + lane_base + 256 + c0, lo);
        um_s_store16(Xh, Xl, row, c0, Cx, xv);
      }
    }
    um_wait_st();
    um_fence_before();
    __syncthreads();
    stamp(c, 301);
    // ---- layer 1: acc = x W1^T ----
    trace(2001, 32);
    if (tid == 0) {
      trace(2200, 0);
      um_fence_after();
      for (int c0 = 0; c0 < K0; c0 += 32) um_consume_ts(p, 0, c0, 128, (K0 - c0) < 32 ? (K0 - c0) : 32, c0 == 0);
      um_commit(p.acc);
      trace(2201, 0);
    }
    if (warp == 0) um_mbar_wait(p.acc, accn & 1u);      // one polling warp: seven more would contend with the operand reads of the running MMAs
    ++accn;
    __syncthreads();
    um_fence_after();
    trace(2002, 32);
    trace(2202, 0);
    if (tid == 0)
      for (int c0 = 0; c0 < 128; c0 += 32) um_fill(p, UM_W(l0 + 1, 0) + c0 * 128, UM_W(l0 + 1, 1) + c0 * 128, 16384u);
    float rs1 = 1.f, rs2 = 1.f;
    const uint64_t mask1 = um_hidden_fwd<0>(th, bias, ln, H1h, H1l, &rs1);
    um_wait_st();
    um_fence_before();
    __syncthreads();
    stamp(c, 302);
    // ---- layer 2: acc = h1 W2^T ----
    trace(2003, 32);
    if (tid == 0) {
      trace(2210, 0);
      um_fence_after();
      for (int c0 = 0; c0 < 128; c0 += 32) um_consume_ts(p, 0, c0, 128, 32, c0 == 0);
      trace(2211, 0);
      um_commit(p.acc);
    }
    if (warp == 0) um_mbar_wait(p.acc, accn & 1u);      // one polling warp: seven more would contend with the operand reads of the running MMAs
    ++accn;
    __syncthreads();
    um_fence_after();
    trace(2004, 32);
    trace(2212, 0);
    if (tid == 0)                              // W2 backward operand for dH1, one GEMM ahead
      for (int c0 = 0; c0 < 128; c0 += 32) um_fill(p, UM_W(l0 + 1, 2) + c0 * 128, UM_W(l0 + 1, 3) + c0 * 128, 16384u);
    // layer-2 epilogue; the output layer (out <= 16 columns) runs on the CUDA cores from the registers of the same pass
    float lgp[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) lgp[i] = 0.f;
    const uint64_t mask2 = nout <= 4 ? um_hidden_fwd<4>(th, bias + 128, ln, H2h, H2l, &rs2, W3s, lgp)
                         : nout <= 8 ? um_hidden_fwd<8>(th, bias + 128, ln, H2h, H2l, &rs2, W3s, lgp)
                                     : um_hidden_fwd<16>(th, bias + 128, ln, H2h, H2l, &rs2, W3s, lgp);
    if (hh == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) cs[row * 17 + i] = lgp[i];
    }
    trace(2005, 32);
    um_wait_st();
    trace(2006, 32);
    __syncthreads();
    trace(2007, 32);
    stamp(c, 303);
    // ---- heads: losses and dL/d(output) per row (PPO.py:256-279); the row's logits / dZ3 live in shared memory (runtime-bounded
    //      loops over out <= 16 columns: the 16-wide predicated register version was 20 k clk of mostly dead code per tile) ----
    if (hh == 0) {
      float* lg = cs + row * 17;               // logits, then (continuous actor) d/dlog_std per action dim
      float* dz = dzs + row * 17;
#pragma unroll
      for (int j = 0; j < 16; ++j) { lg[j] = j < nout ? (lgp[j] + lg[j]) + bias[256 + j] : 0.f; dz[j] = 0.f; }
      if (valid && role == 0) {
        float lp_now = 0.f, lp_old = 0.f, ent = 0.f, lse = 0.f;
        int act = 0;
        if (a.continuous) {
          for (int j = 0; j < nout; ++j) {
            const float mean = tanhf(lg[j]);
            const float ls = fminf(fmaxf(N.p[N.x_off + j], -20.f), 2.f);
            const float sd = expf(ls);
            const float diff = a.action[(size_t)gi * a.act_cols + j] - mean;
            lp_now += -(diff * diff) / (2.f * (sd * sd)) - logf(sd) - FRL_HALF_LOG_2PI;
            ent += 0.5f + FRL_HALF_LOG_2PI + logf(sd);
          }
          for (int j = 0; j < a.logp_cols; ++j) lp_old += a.logp_old[(size_t)gi * a.logp_cols + j];
        } else {
          float mx = lg[0];
          for (int j = 1; j < nout; ++j) mx = fmaxf(mx, lg[j]);
          float se_ = 0.f;
          for (int j = 0; j < nout; ++j) se_ += expf(lg[j] - mx);
          lse = mx + logf(se_);
          act = (int)a.action[(size_t)gi * a.act_cols];
          for (int j = 0; j < nout; ++j) { const float l_ = lg[j] - lse; ent -= expf(l_) * l_; }
          lp_now = lg[act] - lse;
          lp_old = a.logp_old[(size_t)gi * a.logp_cols];
        }
        const float ratio = expf(lp_now - lp_old);
        float dlp = 0.f, surr = 0.f;
        const float lo_ = 1.f - a.clip_param, hi_ = 1.f + a.clip_param;
        const float rc = fminf(fmaxf(ratio, lo_), hi_);
        for (int k = 0; k < a.n_adv; ++k) {
          const float A = a.adv[(size_t)gi * a.n_adv + k];
          const float s1 = ratio * A, s2 = rc * A;
          surr += fminf(s1, s2);
          const bool inside = ratio >= lo_ && ratio <= hi_;
          if (inside || s1 < s2) dlp += -A * ratio * inv_rn;
        }
        la += -surr * inv_rn;
        le += ent;
        if (a.continuous) {
          for (int j = 0; j < nout; ++j) {
            const float mean = tanhf(lg[j]);
            const float lsr = N.p[N.x_off + j];
            const float ls = fminf(fmaxf(lsr, -20.f), 2.f);
            const float sd = expf(ls);
            const float diff = a.action[(size_t)gi * a.act_cols + j] - mean;
            dz[j] = dlp * (diff / (sd * sd)) * (1.f - mean * mean);
            lg[j] = (lsr >= -20.f && lsr <= 2.f) ? dlp * ((diff * diff) / (sd * sd) - 1.f) - a.entropy_coef * inv_rows : 0.f;
          }
        } else {
          for (int j = 0; j < nout; ++j) {
            const float l_ = lg[j] - lse, pj = expf(l_);
            float g = dlp * ((j == act ? 1.f : 0.f) - pj);
            g += -a.entropy_coef * inv_rows * (-pj * (l_ + ent));
            dz[j] = g;
          }
        }
      } else if (valid) {
        float g = 0.f, l = 0.f;
        for (int k = 0; k < a.n_adv; ++k) {
          const float d = lg[0] - a.v_target[(size_t)gi * a.n_adv + k];
          if (a.value_loss == 1) {
            const float e = -d, ae = fabsf(e), dl = a.huber_delta;
            l += (ae <= dl) ? 0.5f * e * e : dl * (ae - 0.5f * dl);
            g += -((ae <= dl) ? e : (e > 0.f ? dl : -dl)) * inv_rn;
          } else if (a.value_loss == 2) {                 // clipped value loss, as in algo_ppo.cuh (MAPPO_discrete.py:350-357)
            const float vo = a.v_old[(size_t)gi * a.n_adv + k], dv = lg[0] - vo;
            const float ec = fminf(fmaxf(dv, -a.clip_param), a.clip_param) + vo - a.v_target[(size_t)gi * a.n_adv + k];
            const bool inside = dv >= -a.clip_param && dv <= a.clip_param;
            const float qo = d * d, qc = ec * ec;
            if (qo >= qc) { l += qo; g += 2.f * d * inv_rn; }
            else { l += qc; if (inside) g += 2.f * ec * inv_rn; }
          } else {
            g += 2.f * d * inv_rn;
            l += d * d;
          }
        }
        lc += l;
        dz[0] = g;
      } else if (role == 0 && a.continuous) {
        for (int j = 0; j < 16; ++j) lg[j] = 0.f;      // padded rows add nothing to d/dlog_std
      }
      float d16[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) d16[i] = dz[i];
      um_s_store16(D3h, D3l, row, 0, 32, d16);
    }
    trace(2008, 32);
    __syncthreads();
    trace(2009, 32);
    stamp(c, 304);
    {                                          // db3: column sums of dZ3 — 16 segments of 8 rows, then the segments in order
      float s = 0.f;
      for (int r = 0; r < 8; ++r) s += dzs[((tid >> 4) * 8 + r) * 17 + (tid & 15)];
      red[tid] = s;
    }
    __syncthreads();
    if (tid < 16) {
      float s = 0.f;
      for (int g = 0; g < 16; ++g) s += red[g * 16 + tid];
      accX[tid] += s;
    }
    if (role == 0 && a.continuous) {
      __syncthreads();
      {                                        // cs rows hold d/dlog_std per action dim since the head
        float s = 0.f;
        for (int r = 0; r < 8; ++r) s += cs[((tid >> 4) * 8 + r) * 17 + (tid & 15)];
        red[tid] = s;
      }
      __syncthreads();
      if (tid < 16) {
        float s = 0.f;
        for (int g = 0; g < 16; ++g) s += red[g * 16 + tid];
        accX[16 + tid] += s;
      }
    }
    // dH2 = dZ3 W3 on the CUDA cores, straight into the layer-2 backward epilogue
    um_hidden_bwd(th, mask2, ln, rs2, H2h, H2l, D2h, D2l, true, accB + 512, W3s, dzs + row * 17, nout);
    um_wait_st();
    um_fence_before();
    um_fence_proxy_async();
    __syncthreads();
    stamp(c, 305);
    // ---- dH1 = dZ2 W2 ----
    if (tid == 0) {
      um_fence_after();
      for (int c0 = 0; c0 < 128; c0 += 32) um_consume_ts(p, 0, c0, 128, 32, c0 == 0);
      um_commit(p.acc);
    }
    if (warp == 0) um_mbar_wait(p.acc, accn & 1u);      // one polling warp: seven more would contend with the operand reads of the running MMAs
    ++accn;
    __syncthreads();
    um_fence_after();
    um_hidden_bwd(th, mask1, ln, rs1, H1h, H1l, D1h, D1l, false, accB);
    um_fence_before();
    um_fence_proxy_async();                    // scratch written by the generic proxy -> bulk-copy (async proxy) reads
    __syncthreads();
    stamp(c, 306);
    // ---- phase B: dW3^T = h2^T dZ3 | dW2 = dZ2^T h1, db2 | dW1 = dZ1^T x, db1   (12 steps of 32 rows, two slots each) ----
    if (tid == 0) {
      um_fence_after();
      for (int st = 0; st <= 12; ++st) {
        if (st < 12) {
          const int g = st >> 2, j = st & 3;
          const float *ah, *al, *bh, *bl;
          uint32_t ab = 16384u, bb;
          if (g == 0) { ah = H2h; al = H2l; bh = D3h + j * 1024; bl = D3l + j * 1024; bb = 4096u; }
          else if (g == 1) { ah = D2h; al = D2l; bh = H1h + j * 4096; bl = H1l + j * 4096; bb = 16384u; }
          else { ah = D1h; al = D1l; bh = Xh + j * 32 * Cx; bl = Xl + j * 32 * Cx; bb = (uint32_t)(128 * Cx); }
          um_fill(p, ah + j * 4096, al + j * 4096, ab);
          um_fill(p, bh, bl, bb);
        }
        if (st > 0) {
          const int g = (st - 1) >> 2, j = (st - 1) & 3;
          if (g == 0) um_consume_ss(p, 0, 128, 32, 16, j == 0);
          else if (g == 1) um_consume_ss(p, 384, 128, 128, 128, j == 0 && it == 0);
          else um_consume_ss(p, 64, 128, Cx, K0, j == 0);
        }
      }
      um_commit(p.acc);
    }
    if (warp == 0) um_mbar_wait(p.acc, accn & 1u);      // one polling warp: seven more would contend with the operand reads of the running MMAs
    ++accn;
    __syncthreads();
    um_fence_after();
    if (hh == 0) {
      float v[16];
      um_ld16(tm + lane_base, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) accW3[i * 128 + row] += v[i];
      for (int c0 = 0; c0 < K0; c0 += 16) {
        um_ld16(tm + lane_base + 64 + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) accW1[(c0 + i) * 128 + row] += v[i];
      }
    }
    um_fence_before();
    __syncthreads();
    stamp(c, 307);
  }

  // ---- gradient partial of this (net, CTA) -> gpart slot ci; padded entries are written as zeros ----
  um_fence_after();
  float* gp = a.gpart + (size_t)ci * N.n_p;
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    const int c0 = hh * 64 + j * 16;
    float v[16];
    um_ld16(tm + lane_base + 384 + c0, v);
    if (row < L1.out_pad) {
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        st4(gp + L1.w_off + row * 128 + c0 + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    }
  }
  for (int e = tid; e < L0.out_pad * L0.in_pad; e += FRL_NT) {
    const int n = e / L0.in_pad, k = e % L0.in_pad;
    gp[L0.w_off + e] = (n < L0.out && k < L0.in) ? accW1[k * 128 + n] : 0.f;
  }
  for (int e = tid; e < L2.out_pad * L2.in_pad; e += FRL_NT) {
    const int n = e / L2.in_pad, k = e % L2.in_pad;
    gp[L2.w_off + e] = (n < L2.out && k < L2.in) ? accW3[n * 128 + k] : 0.f;
  }
  __syncthreads();
  if (tid < L0.out_pad) gp[L0.b_off + tid] = tid < L0.out ? ((accB[tid] + accB[128 + tid]) + accB[256 + tid]) + accB[384 + tid] : 0.f;
  if (tid < L1.out_pad) gp[L1.b_off + tid] = tid < L1.out ? ((accB[512 + tid] + accB[640 + tid]) + accB[768 + tid]) + accB[896 + tid] : 0.f;
  if (tid < L2.out_pad) gp[L2.b_off + tid] = tid < L2.out ? accX[tid] : 0.f;
  if (role == 0 && a.continuous && tid < L2.out_pad) gp[N.x_off + tid] = tid < nout ? accX[16 + tid] : 0.f;
  // loss partials (fixed shuffle-tree association of block_sum)
  __syncthreads();
  red[tid] = role ? lc : la;
  __syncthreads();
  const float l0s = block_sum(red);
  red[tid] = le;
  __syncthreads();
  const float l2s = block_sum(red);
  if (tid == 0) {
    a.stats[ci * 8 + (role ? 1 : 0)] = l0s;
    if (role == 0) a.stats[ci * 8 + 2] = l2s;
  }
  um_fence_before();
  __syncthreads();
  stamp(c, 308);
  if (warp == 0) um_tmem_dealloc<512>(tm);
  if (tid == 0)
    for (int i = 0; i < 9; ++i) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(um_smem_u32(&bars[i])) : "memory");
#undef UM_W
}
#endif  // !FRL_EMUL
