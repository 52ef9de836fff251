// Fused PPO minibatch update + GAE.
//   PPO.learn minibatch loop      PPO_file/PPO.py:245-283   (clipped surrogate, entropy bonus, value MSE)
//   Agent.update_ac_              PPO_file/PPO.py:145-152   (two backward()s, clip 0.5 actor / 0.5 critic, ONE optimiser step)
//   c_adamw.AdamW.step            PPO_file/c_adamw.py:65-122 (cautious mask with PER-TENSOR mean, eps after sqrt)
//   GAE                           PPO_file/PPO.py:222-233   (float64 reverse scan, zero tail, v_target = adv + V)
// Actor (layers 0-2 [+ log_std]) and critic (layers 3-5) share ONE parameter block / optimiser state, like the
// reference's merged `ac_optimizer`.  One persistent launch runs every minibatch of every epoch (n_updates steps).
#pragma once
#include "algo_ac.cuh"

#define FRL_NSEG (2 * FRL_MAX_LAYERS + 1)
#define FRL_HALF_LOG_2PI 0.91893853320467274178f
#include "algo_ppo_umma.cuh"

// digamma / trigamma for x >= 1 (recurrence up to x >= 6, then the asymptotic series; evaluated in double: the Beta head calls them
// a few times per row) — torch.distributions.Beta.entropy / log_prob gradients, PPO_with_tricks.py:120-150, 327-333
FRL_HD double frl_digamma(double x) {
  double r = 0.0;
  while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + log(x) - 0.5 / x - f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f * (1.0 / 132.0)))));
}
FRL_HD double frl_trigamma(double x) {
  double r = 0.0;
  while (x < 6.0) { r += 1.0 / (x * x); x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + 1.0 / x + 0.5 * f + (1.0 / x) * f * (1.0 / 6.0 - f * (1.0 / 30.0 - f * (1.0 / 42.0 - f * (1.0 / 30.0))));
}
FRL_HD float frl_softplus(float z) { return z > 20.f ? z : log1pf(expf(z)); }       // F.softplus (beta 1, threshold 20)

// tensor (segment) id of parameter index p and that tensor's logical element count
FRL_DEV int seg_of(const frl_net_t& n, int p, int* numel) {
  for (int li = 0; li < n.n_layers; ++li) {
    const frl_layer_t& L = n.L[li];
    if (p >= L.w_off && p < L.w_off + L.out_pad * L.in_pad) { *numel = L.out * L.in; return 2 * li; }
    if (p >= L.b_off && p < L.b_off + L.out_pad) { *numel = L.out; return 2 * li + 1; }
  }
  *numel = n.x_len;
  return 2 * FRL_MAX_LAYERS;
}

// ---- F.layer_norm(x, x.size()[1:]) without affine, eps 1e-5 (MAPPO.py:145-151), rows of a [R][ld] smem tile ----------
template <int R>
FRL_NI_MISC void ln_fwd(const float* X, int ld, int n, float* Y, float* rstd, float* scratch /*[R*32*2]*/) {
  // one warp per row; R > FRL_NT / 32 rows (the 16-row inference tile) take the rows in rounds of FRL_NT / 32
  FRL_PAR(t) {
    for (int r = t >> 5; r < R; r += FRL_NT / 32) {
      const int l = t & 31;
      float s = 0.f;
      for (int k = l; k < n; k += 32) s += X[r * ld + k];
      scratch[r * 32 + l] = s;
    }
  }
  FRL_SYNC();
  FRL_PAR(t) {
    for (int r = t >> 5; r < R; r += FRL_NT / 32) {
      const int l = t & 31;
      float mean = 0.f;
      for (int i = 0; i < 32; ++i) mean += scratch[r * 32 + i];
      mean = mean / (float)n;
      float s = 0.f;
      for (int k = l; k < n; k += 32) { const float d = X[r * ld + k] - mean; s += d * d; }
      scratch[R * 32 + r * 32 + l] = s;
    }
  }
  FRL_SYNC();
  FRL_PAR(t) {
    for (int r = t >> 5; r < R; r += FRL_NT / 32) {
      const int l = t & 31;
      float mean = 0.f, var = 0.f;
      for (int i = 0; i < 32; ++i) { mean += scratch[r * 32 + i]; var += scratch[R * 32 + r * 32 + i]; }
      mean = mean / (float)n;
      var = var / (float)n;
      const float rs = 1.0f / sqrtf(var + 1e-5f);
      for (int k = l; k < ld; k += 32) Y[r * ld + k] = (k < n) ? (X[r * ld + k] - mean) * rs : 0.f;
      if (l == 0) rstd[r] = rs;
    }
  }
  FRL_SYNC();
}

// dX = rstd * (dY - mean(dY) - Y * mean(dY*Y)), then optionally masked by relu'(H) (H = the pre-LN activation)
template <int R>
FRL_NI_MISC void ln_bwd(const float* dY, const float* Y, const float* rstd, const float* H, int ld, int n, float* dX,
                        float* scratch /*[R*32*2]*/) {
  FRL_PAR(t) {
    if (t < R * 32) {
      const int r = t >> 5, l = t & 31;
      float s1 = 0.f, s2 = 0.f;
      for (int k = l; k < n; k += 32) { const float g = dY[r * ld + k]; s1 += g; s2 += g * Y[r * ld + k]; }
      scratch[t] = s1; scratch[R * 32 + t] = s2;
    }
  }
  FRL_SYNC();
  FRL_PAR(t) {
    if (t < R * 32) {
      const int r = t >> 5, l = t & 31;
      float m1 = 0.f, m2 = 0.f;
      for (int i = 0; i < 32; ++i) { m1 += scratch[r * 32 + i]; m2 += scratch[R * 32 + r * 32 + i]; }
      m1 = m1 / (float)n; m2 = m2 / (float)n;
      for (int k = l; k < ld; k += 32) {
        float v = 0.f;
        if (k < n) {
          v = rstd[r] * (dY[r * ld + k] - m1 - Y[r * ld + k] * m2);
          if (H && !(H[r * ld + k] > 0.f)) v = 0.f;
        }
        dX[r * ld + k] = v;
      }
    }
  }
  FRL_SYNC();
}

// activations kept by one 3-layer net pass (layer-norm variant keeps both the ReLU outputs and their normalised copies)
struct NetBufs { float *X0, *H1, *H1n, *H2, *H2n, *rs1, *rs2, *scratch; };

template <int R>
FRL_NI_MISC void tile_copy(const float* X, float* Y, int n) {
  FRL_PAR(t) { for (int e = t; e < n; e += FRL_NT) Y[e] = X[e]; }
  FRL_SYNC();
}
// ln: 0 none, 1 input + hidden (MAPPO.py), 2 hidden only, 3 input only (the acting networks of MAPPO_discrete.py: its discrete actor
// drops the normalised input, and LayerNorm / feature_norm are separate switches); the backward (net_bwd) exists for 0 / 1
template <int R, int HM = 0>
FRL_DEV void net_fwd(Cta& c, const frl_net_t& N, int l0, int ln, const float* Xin, int ip, int n_in, const NetBufs& b, int ldh,
                     float* OUT, int ldo, Hint next) {
  if (!ln) { mlp_fwd<R, HM>(c, N, l0, 3, Xin, ip, b.H1, b.H2, ldh, OUT, ldo, FRL_ACT_NONE, next); return; }
  if (ln != 2) ln_fwd<R>(Xin, ip, n_in, b.X0, b.rs1, b.scratch);          // feature_norm (rstd not needed later)
  else tile_copy<R>(Xin, b.X0, R * ip);
  layer_fwd<R>(c, N, l0, b.X0, ip, b.H1, ldh, FRL_ACT_RELU, fwd_hint(N, l0 + 1));
  if (ln != 3) ln_fwd<R>(b.H1, ldh, N.L[l0].out, b.H1n, b.rs1, b.scratch);
  else tile_copy<R>(b.H1, b.H1n, R * ldh);
  layer_fwd<R>(c, N, l0 + 1, b.H1n, ldh, b.H2, ldh, FRL_ACT_RELU, fwd_hint(N, l0 + 2));
  if (ln != 3) ln_fwd<R>(b.H2, ldh, N.L[l0 + 1].out, b.H2n, b.rs2, b.scratch);
  else tile_copy<R>(b.H2, b.H2n, R * ldh);
  layer_fwd<R>(c, N, l0 + 2, b.H2n, ldh, OUT, ldo, FRL_ACT_NONE, next);
}

template <int R, int HM = 0>
FRL_DEV void net_bwd(Cta& c, const frl_net_t& N, int l0, bool ln, const float* Xin, int ip, const NetBufs& b, int ldh, const float* dOUT,
                     int ldo, float* D1, float* D2, float* gp, bool accumulate, Hint next) {
  if (!ln) { mlp_bwd<R, HM>(c, N, l0, 3, Xin, ip, b.H1, b.H2, ldh, dOUT, ldo, D1, D2, nullptr, 0, gp, accumulate, next); return; }
  const frl_layer_t &L0 = N.L[l0], &L1 = N.L[l0 + 1], &L2 = N.L[l0 + 2];
  gemm_outer<R>(dOUT, ldo, L2.out_pad, b.H2n, ldh, L2.in_pad, L2.in, gp + L2.w_off, gp + L2.b_off, accumulate);
  layer_bwd_dx<R>(c, N, l0 + 2, dOUT, ldo, nullptr, 0, D2, ldh, bwd_hint(N, l0 + 1));
  ln_bwd<R>(D2, b.H2n, b.rs2, b.H2, ldh, L1.out, D1, b.scratch);          // D1 = dL/d(pre-ReLU of layer l0+1)
  gemm_outer<R>(D1, ldh, L1.out_pad, b.H1n, ldh, L1.in_pad, L1.in, gp + L1.w_off, gp + L1.b_off, accumulate);
  layer_bwd_dx<R>(c, N, l0 + 1, D1, ldh, nullptr, 0, D2, ldh, next);
  ln_bwd<R>(D2, b.H1n, b.rs1, b.H1, ldh, L0.out, D1, b.scratch);
  gemm_outer<R>(D1, ldh, L0.out_pad, b.X0, ip, L0.in_pad, L0.in, gp + L0.w_off, gp + L0.b_off, accumulate);
}

#ifndef FRL_EMUL
// Exchange stage of the data-parallel peers (frl_dp_peers_t): publish the epoch to every peer, wait for every peer's epoch,
// sum the world's gradient blocks in rank order into net.g.  All peers run the same launch; a peer that never arrives is
// reported after ~2 s through out[u][7] = -1 instead of hanging the cooperative grid.
FRL_DEV void dp_exchange(Cta& c, const frl_ppo_args_t& a, int u) {
  const frl_net_t& N = a.net;
  const unsigned epoch = a.dp.epoch0 + (unsigned)u + 1u;
  const int world = a.dp.world, rank = a.dp.rank, tid = (int)threadIdx.x;
  if (c.cta == 0 && tid < world && tid != rank) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.dp.flags[tid] + rank), "r"(epoch) : "memory");
  }
  if (tid < world && tid != rank) {
    const unsigned* f = a.dp.flags[rank] + tid;
    unsigned* dead = a.dp.flags[rank] + 32;        // sticky: a peer timed out once -> later exchanges do not wait again
    long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - epoch) >= 0) break;
      if (*(volatile unsigned*)dead) { if (c.cta == 0) a.out[u * 8 + 7] = -1.f; break; }
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 2000000000ll) { *(volatile unsigned*)dead = 1u; if (c.cta == 0) a.out[u * 8 + 7] = -1.f; break; }
    }
  }
  __syncthreads();
  const size_t par = (size_t)(epoch & 1u) * N.n_p;
  for (int p = (c.cta * FRL_NT + tid) * 4; p < N.n_p; p += c.ncta * FRL_NT * 4) {
    float4 v[FRL_DP_MAX_RANKS];                    // all peers' loads in flight together (NVLink latency once, not world times)
#pragma unroll
    for (int r = 0; r < FRL_DP_MAX_RANKS; ++r)
      if (r < world)
        asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[r].x), "=f"(v[r].y), "=f"(v[r].z), "=f"(v[r].w) : "l"(a.dp.g[r] + par + p) : "memory");
    float4 sgm = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < FRL_DP_MAX_RANKS; ++r)
      if (r < world) sgm = f4add(sgm, v[r]);       // rank order: bit-identical on every rank
    st4(N.g + p, sgm);
  }
  __syncthreads();
}
#endif

// R = batch rows per CTA tile: 8 for the reference-sized minibatches (more CTAs per minibatch), 16 for large minibatches
// (>= 1024 rows: twice the FMAs per staged weight and per barrier; chosen by frl_ppo_update when the tile fits in shared memory).
// HM (compile time): the `tanh` switch of PPO_with_tricks.py — bit 0: actor, bit 1: critic hidden layers use tanh.  HM = 0 is the
// kernel every other PPO-family class launches; the tanh instantiations exist for 8-row tiles only.
// UM = 1 (GPU only): the tensor-core variant of algo_ppo_umma.cuh — one more stage in front (split the weights into hi / lo TF32
// operands), stage "fwd/bwd" runs on tcgen05 over 128-row tiles, the reduce / clip / optimiser stages below are shared.
// GRP = 1: group mode (algo_ppo_group.cuh, MAPPO_discrete.py's episode-wide LayerNorm): stage "fwd/bwd" walks whole groups per CTA.
#include "algo_ppo_group.cuh"
template <int R, int HM = 0, int UM = 0, int GRP = 0>
struct PpoAlgoT {
  typedef frl_ppo_args_t Args;
  static const int NSTAGES = UM ? 7 : 6;      // [split,] fwd/bwd, reduce, exchange (data-parallel peers only), norms, optimiser x 2
  static const int XCHG = UM ? 3 : 2;         // physical index of the exchange stage
  FRL_SHD bool writes_params(int) { return true; }
  // the exchange stage exists for data-parallel peers only; the weight split of the tensor-core variant runs on the first update of
  // a launch (the optimiser stages keep the split copies current afterwards)
  // single rank, gradient used as reduced: the reduce stage also leaves the per-CTA sums of squares, the norms stage is skipped
  FRL_SHD bool norms_in_reduce(const Args& a) { return a.dp.world <= 1 && a.stage_hi == 0 && (a.grad_scale == 0.f || a.grad_scale == 1.f); }
  FRL_SHD bool stage_enabled(int s, int u, const Args& a) {
    return (s != XCHG || a.dp.world > 1) && !(UM && s == 0 && u > 0) && !(s == XCHG + 1 && norms_in_reduce(a)) &&
           !(s == XCHG + 3 && a.optimizer == FRL_OPT_ADAM);      // plain Adam is one sweep: no second optimiser stage, no barrier for it
  }
  FRL_SHD int wbuf_floats(const Args& a) { return UM ? 32 : ((AcAlgo::max_layer_floats(a.net) + 31) & ~31); }      // UM: no weight stager, its ring lives in `user`
  FRL_SHD int user_floats(const Args& a) {
#ifndef FRL_EMUL
    if (UM) return UM_USER_FLOATS;
#endif
    const int ldh = act_ld(a.net.L[0].out_pad), ip = a.net.L[0].in_pad, cip = a.net.L[3].in_pad, ap = a.net.L[2].out_pad;
    // the LayerNorm variant keeps the normalised copies of the input and of both hidden activations, per net
    const int ln_extra = a.layer_norm ? (ip + cip + 4 * ldh) : 0;
    const int std_floats = R * (ip + cip + ln_extra + 6 * ldh + 4 * ap + 8 + 4 * a.n_adv + 8 + 4 + 64) + 2 * FRL_NT + 64 + FRL_NSEG + 3;
    if (GRP) { const int gf = ppo_group_user_floats(a); return gf > std_floats ? gf : std_floats; }
    return std_floats;
  }
  FRL_SHD int grid(const Args& a, int max_ctas) {
#ifndef FRL_EMUL
    if (UM) return um_grid(a, max_ctas);
#endif
    int tiles = GRP ? a.mb / a.group_rows : (a.mb + R - 1) / R;
    return tiles < max_ctas ? tiles : max_ctas;
  }
  FRL_SHD int n_updates(const Args& a) { return a.n_updates; }

  FRL_SDEV bool is_critic(const frl_net_t& n, int p) { return p >= n.L[3].w_off && p < n.x_off; }

  FRL_SDEV void stage(int s, int u, Cta& c, float* user, const Args& a) {
    const frl_net_t& N = a.net;
    const int ldh = act_ld(N.L[0].out_pad), ip = N.L[0].in_pad, ap = N.L[2].out_pad, nout = N.L[2].out;
    // physical -> logical stage (stage_lo / stage_hi of the host-driven data-parallel split count the logical ones):
    //   [split +] fwd/bwd = 0, reduce = 1, (exchange), norms = 2, optimiser = 3, 4
    if (s == XCHG) {
#ifndef FRL_EMUL
      dp_exchange(c, a, u);
#endif
      return;
    }
    const int phys = s;
    s = phys > XCHG ? phys - 1 : phys;
    if (UM) s = s == 0 ? 0 : s - 1;
    if (a.stage_hi > 0 && (s < a.stage_lo || s >= a.stage_hi)) return;
#ifndef FRL_EMUL
    if (UM && phys == 0) { um_split(c, a); return; }
    if (UM && phys == 1) { ppo_umma_stage(c, user, a, u); return; }
#endif
    const int rows = a.mb_rows[u];
    const int ntile = GRP ? rows / a.group_rows : (rows + R - 1) / R;        // work items of stage 0 = gradient partials to reduce
#ifndef FRL_EMUL
    const int ncontrib = UM ? um_ncontrib(rows, c.ncta) : (ntile < c.ncta ? ntile : c.ncta);
#else
    const int ncontrib = ntile < c.ncta ? ntile : c.ncta;
#endif
    SmemBump sb; sb.p = user;
    const int cip = N.L[3].in_pad;
    const bool ln = a.layer_norm != 0;
    float* X = sb.take(R * ip);
    float* XC = sb.take(R * cip);       // critic input tile (joint obs for MAPPO, else a copy of X)
    NetBufs ba, bc;
    ba.H1 = sb.take(R * ldh); ba.H2 = sb.take(R * ldh);
    bc.H1 = sb.take(R * ldh); bc.H2 = sb.take(R * ldh);
    ba.X0 = bc.X0 = ba.H1n = ba.H2n = bc.H1n = bc.H2n = nullptr;
    if (a.layer_norm) {
      ba.X0 = sb.take(R * ip);  bc.X0 = sb.take(R * cip);
      ba.H1n = sb.take(R * ldh); ba.H2n = sb.take(R * ldh);
      bc.H1n = sb.take(R * ldh); bc.H2n = sb.take(R * ldh);
    }
    ba.rs1 = sb.take(R); ba.rs2 = sb.take(R); bc.rs1 = sb.take(R); bc.rs2 = sb.take(R);
    ba.scratch = bc.scratch = sb.take(R * 64);
    float* D1 = sb.take(R * ldh);
    float* D2 = sb.take(R * ldh);
    float* OA = sb.take(R * ap);
    float* dOA = sb.take(R * ap);
    float* ACT = sb.take(R * ap);
    float* LPO = sb.take(R * ap);
    float* V = sb.take(R * 4);
    float* dV = sb.take(R * 4);
    float* ADV = sb.take(R * a.n_adv);
    float* VT = sb.take(R * a.n_adv);
    float* VO = sb.take(R * a.n_adv);   // rollout values of the clipped value loss (value_loss 2); user_floats counts 4 n_adv columns
    float* red0 = sb.take(FRL_NT);
    float* red1 = sb.take(FRL_NT);
    int* segc = (int*)sb.take(FRL_NSEG + 3);
    float* gp = a.gpart + (size_t)c.cta * N.n_p;

    if (GRP && s == 0) { ppo_group_stage(c, user, a, u); return; }
    if (!UM && !GRP && s == 0) {        // (compile-time guard: the tensor-core and group kernels never reach this stage body — keep it out of their code)
      if (c.cta >= ntile) return;
      float la = 0.f, lc = 0.f, le = 0.f;
      bool first = true;
      const float inv_rows = 1.0f / (float)rows, inv_rn = 1.0f / (float)(rows * a.n_adv);
      for (int tile = c.cta; tile < ntile; tile += c.ncta) {
        const int row0 = tile * R;
        const int nvalid = (rows - row0) < R ? (rows - row0) : R;
        const int64_t* idx = a.indices + (size_t)u * a.mb + row0;
        stage_prefetch(c, layer_fwd_src(N, 0), layer_fwd_bytes(N.L[0]));
        FRL_PAR(t) {
          for (int e = t; e < R * ip; e += FRL_NT) {
            const int r = e / ip, j = e % ip;
            X[e] = (r < nvalid && j < a.obs_dim) ? a.obs[(size_t)idx[r] * a.obs_dim + j] : 0.f;
          }
          for (int e = t; e < R * cip; e += FRL_NT) {
            const int r = e / cip, j = e % cip;
            float v = 0.f;
            if (r < nvalid) {
              if (a.critic_obs) { if (j < a.critic_obs_dim) v = a.critic_obs[(size_t)idx[r] * a.critic_obs_dim + j]; }
              else if (j < a.obs_dim) v = a.obs[(size_t)idx[r] * a.obs_dim + j];
            }
            XC[e] = v;
          }
          for (int e = t; e < R * ap; e += FRL_NT) {
            const int r = e / ap, j = e % ap;
            ACT[e] = (r < nvalid && j < a.act_cols) ? a.action[(size_t)idx[r] * a.act_cols + j] : 0.f;
            LPO[e] = (r < nvalid && j < a.logp_cols) ? a.logp_old[(size_t)idx[r] * a.logp_cols + j] : 0.f;
          }
          for (int e = t; e < R * a.n_adv; e += FRL_NT) {
            const int r = e / a.n_adv, j = e % a.n_adv;
            ADV[e] = (r < nvalid) ? a.adv[(size_t)idx[r] * a.n_adv + j] : 0.f;
            VT[e] = (r < nvalid) ? a.v_target[(size_t)idx[r] * a.n_adv + j] : 0.f;
            VO[e] = (r < nvalid && a.value_loss == 2) ? a.v_old[(size_t)idx[r] * a.n_adv + j] : 0.f;
          }
        }
        FRL_SYNC();
        const int c_in = a.critic_obs ? a.critic_obs_dim : a.obs_dim;
        net_fwd<R, (HM & 1)>(c, N, 0, ln ? 1 : 0, X, ip, a.obs_dim, ba, ldh, OA, ap, fwd_hint(N, 3));
        net_fwd<R, ((HM >> 1) & 1)>(c, N, 3, ln ? 1 : 0, XC, cip, c_in, bc, ldh, V, 4, bwd_hint(N, 2));
        // policy head: log-prob, entropy, ratio, clipped surrogate and its gradient w.r.t. the actor output
        FRL_PAR(t) {
          float sa = 0.f, se = 0.f;
          if (t < R) {
            const int r = t;
            for (int j = 0; j < ap; ++j) dOA[r * ap + j] = 0.f;
            if (r < nvalid) {
              float lp_now = 0.f, lp_old = 0.f, ent = 0.f;
              if (a.continuous == 2) {
                // Beta head (Actor_Beta, PPO_with_tricks.py:120-150): output columns [alpha logits (A) | beta logits (A)],
                // alpha = softplus(.) + 1, beta likewise; Beta(alpha, beta).log_prob(x), .entropy() as torch's Dirichlet computes them
                const int A = nout >> 1;
                for (int j = 0; j < A; ++j) {
                  const double al = (double)frl_softplus(OA[r * ap + j]) + 1.0, be = (double)frl_softplus(OA[r * ap + A + j]) + 1.0;
                  const double x = (double)ACT[r * ap + j];
                  const double lnB = lgamma(al) + lgamma(be) - lgamma(al + be);
                  lp_now += (float)((al - 1.0) * log(x) + (be - 1.0) * log(1.0 - x) - lnB);
                  ent += (float)(lnB + (al + be - 2.0) * frl_digamma(al + be) - (al - 1.0) * frl_digamma(al) - (be - 1.0) * frl_digamma(be));
                }
                for (int j = 0; j < a.logp_cols; ++j) lp_old += LPO[r * ap + j];
              } else if (a.continuous) {
                for (int j = 0; j < nout; ++j) {
                  const float mean = tanhf(OA[r * ap + j]);
                  const float ls = fminf(fmaxf(N.p[N.x_off + j], -20.f), 2.f);
                  const float sd = expf(ls);
                  const float diff = ACT[r * ap + j] - mean;
                  lp_now += -(diff * diff) / (2.f * (sd * sd)) - logf(sd) - FRL_HALF_LOG_2PI;
                  ent += 0.5f + FRL_HALF_LOG_2PI + logf(sd);
                }
                for (int j = 0; j < a.logp_cols; ++j) lp_old += LPO[r * ap + j];
              } else {
                float mx = OA[r * ap];
                for (int j = 1; j < nout; ++j) mx = fmaxf(mx, OA[r * ap + j]);
                float se_ = 0.f;
                for (int j = 0; j < nout; ++j) se_ += expf(OA[r * ap + j] - mx);
                const float lse = mx + logf(se_);
                const int act = (int)ACT[r * ap];
                lp_now = OA[r * ap + act] - lse;
                for (int j = 0; j < nout; ++j) { const float lg = OA[r * ap + j] - lse; ent -= expf(lg) * lg; }
                lp_old = LPO[r * ap];
              }
              const float ratio = expf(lp_now - lp_old);
              // -min(r*A, clamp(r)*A) averaged over rows and advantage columns; gradient through the active branch
              float dlp = 0.f, surr = 0.f;
              const float lo = 1.f - a.clip_param, hi = 1.f + a.clip_param;
              const float rc = fminf(fmaxf(ratio, lo), hi);
              for (int k = 0; k < a.n_adv; ++k) {
                const float A = ADV[r * a.n_adv + k];
                const float s1 = ratio * A, s2 = rc * A;
                surr += fminf(s1, s2);
                const bool inside = ratio >= lo && ratio <= hi;
                if (inside || s1 < s2) dlp += -A * ratio * inv_rn;      // ties (inside the clip range) pass the full gradient
              }
              sa = -surr * inv_rn;
              se = ent;
              // chain to the actor output
              if (a.continuous == 2) {
                const int A = nout >> 1;
                const float dent = -a.entropy_coef * inv_rows;                       // d(loss) / d(entropy of this row)
                for (int j = 0; j < A; ++j) {
                  const float za = OA[r * ap + j], zb = OA[r * ap + A + j];
                  const double al = (double)frl_softplus(za) + 1.0, be = (double)frl_softplus(zb) + 1.0, x = (double)ACT[r * ap + j];
                  const double psi0 = frl_digamma(al + be), tri0 = frl_trigamma(al + be);
                  const double dlp_a = log(x) + psi0 - frl_digamma(al), dlp_b = log(1.0 - x) + psi0 - frl_digamma(be);
                  const double dh_a = -(al - 1.0) * frl_trigamma(al) + (al + be - 2.0) * tri0;
                  const double dh_b = -(be - 1.0) * frl_trigamma(be) + (al + be - 2.0) * tri0;
                  const float sa_ = 1.f / (1.f + expf(-za)), sb_ = 1.f / (1.f + expf(-zb));        // d softplus / dz (1 beyond the threshold)
                  dOA[r * ap + j] = (float)((double)dlp * dlp_a + (double)dent * dh_a) * (za > 20.f ? 1.f : sa_);
                  dOA[r * ap + A + j] = (float)((double)dlp * dlp_b + (double)dent * dh_b) * (zb > 20.f ? 1.f : sb_);
                }
              } else if (a.continuous) {
                for (int j = 0; j < nout; ++j) {
                  const float mean = tanhf(OA[r * ap + j]);
                  const float ls = fminf(fmaxf(N.p[N.x_off + j], -20.f), 2.f);
                  const float sd = expf(ls);
                  const float diff = ACT[r * ap + j] - mean;
                  dOA[r * ap + j] = dlp * (diff / (sd * sd)) * (1.f - mean * mean);
                  LPO[r * ap + j] = dlp * ((diff * diff) / (sd * sd) - 1.f);      // d/dlog_std via log-prob (reuse LPO)
                }
              } else {
                float mx = OA[r * ap];
                for (int j = 1; j < nout; ++j) mx = fmaxf(mx, OA[r * ap + j]);
                float se_ = 0.f;
                for (int j = 0; j < nout; ++j) se_ += expf(OA[r * ap + j] - mx);
                const float lse = mx + logf(se_);
                const int act = (int)ACT[r * ap];
                for (int j = 0; j < nout; ++j) {
                  const float lg = OA[r * ap + j] - lse, pj = expf(lg);
                  // d logp_a/dz_j = [j==a] - p_j ;  dH/dz_j = -p_j (lg + H)
                  float g = dlp * ((j == act ? 1.f : 0.f) - pj);
                  g += -a.entropy_coef * inv_rows * (-pj * (lg + ent));
                  dOA[r * ap + j] = g;
                }
              }
            } else if (a.continuous == 1) {
              for (int j = 0; j < ap; ++j) LPO[r * ap + j] = 0.f;
            }
          }
          red0[t] = sa; red1[t] = se;
        }
        FRL_SYNC();
        la += block_sum(red0);
        le += block_sum(red1);
        // value loss  mse(v_target, V)  (mean over rows x advantage columns)
        FRL_PAR(t) {
          float l = 0.f;
          if (t < R) {
            for (int j = 0; j < 4; ++j) dV[t * 4 + j] = 0.f;
            if (t < nvalid) {
              float g = 0.f;
              for (int k = 0; k < a.n_adv; ++k) {
                const float d = V[t * 4] - VT[t * a.n_adv + k];
                if (a.value_loss == 1) {
                  // huber(e = v_target - V, delta): e^2/2 inside, delta(|e| - delta/2) outside  (MAPPO.py:273-276)
                  const float e = -d, ae = fabsf(e), dl = a.huber_delta;
                  l += (ae <= dl) ? 0.5f * e * e : dl * (ae - 0.5f * dl);
                  g += -((ae <= dl) ? e : (e > 0.f ? dl : -dl)) * inv_rn;
                } else if (a.value_loss == 2) {
                  // max(e_clip^2, e_orig^2), e_clip = clamp(V - v_old, +-c) + v_old - v_target (MAPPO_discrete.py:350-357); inside the
                  // clip range both are the same number and torch.max's tie shares the gradient between two equal halves
                  const float vo = VO[t * a.n_adv + k], dv = V[t * 4] - vo;
                  const float ec = fminf(fmaxf(dv, -a.clip_param), a.clip_param) + vo - VT[t * a.n_adv + k];
                  const bool inside = dv >= -a.clip_param && dv <= a.clip_param;
                  const float qo = d * d, qc = ec * ec;
                  if (qo >= qc) { l += qo; g += 2.f * d * inv_rn; }
                  else { l += qc; if (inside) g += 2.f * ec * inv_rn; }
                } else {
                  g += 2.f * d * inv_rn;
                  l += d * d;
                }
              }
              dV[t * 4] = g;
            }
          }
          red0[t] = l;
        }
        FRL_SYNC();
        lc += block_sum(red0);
        if (a.continuous == 1) {
          // log_std gradient: log-prob path (stored in LPO) + entropy bonus  -c * mean_rows(sum_j 1)
          FRL_PAR(t) {
            if (t < ap) {
              float g = 0.f;
              if (t < nout) {
                const float lsr = N.p[N.x_off + t];
                if (lsr >= -20.f && lsr <= 2.f) {
                  for (int r = 0; r < nvalid; ++r) g += LPO[r * ap + t] - a.entropy_coef * inv_rows;
                }
              }
              gp[N.x_off + t] = first ? g : gp[N.x_off + t] + g;
            }
          }
          FRL_SYNC();
        }
        net_bwd<R, (HM & 1)>(c, N, 0, ln, X, ip, ba, ldh, dOA, ap, D1, D2, gp, !first, bwd_hint(N, 5));
        net_bwd<R, ((HM >> 1) & 1)>(c, N, 3, ln, XC, cip, bc, ldh, dV, 4, D1, D2, gp, !first, no_hint());
        first = false;
      }
      FRL_PAR(t) {
        if (t == 0) { a.stats[c.cta * 8 + 0] = la; a.stats[c.cta * 8 + 1] = lc; a.stats[c.cta * 8 + 2] = le; }
      }
      FRL_SYNC();
    } else if (s == 1) {
      // fixed-order cross-CTA reduction of the gradient partials -> net.g.  Eight threads share one float4 column: each folds a
      // contiguous eighth of the partials (independent L2 loads in flight instead of one thread walking up to 148 of them), then
      // the eight sub-sums are folded in lane order through shared memory — the association is fixed by (ncontrib, 8) alone.
      const int nq = N.n_p >> 2, per = (ncontrib + 7) >> 3;
      float* r4 = c.red;                                     // [FRL_NT] float4
      // data-parallel peers: the rank's sum goes to its exchange block g[epoch & 1]; the exchange stage writes net.g
      float* gdst = a.dp.world > 1 ? a.dp.g[a.dp.rank] + (size_t)((a.dp.epoch0 + (unsigned)u + 1u) & 1u) * N.n_p : N.g;
      // the clip norms ride along (single rank, unscaled gradient): the thread that writes a quad of net.g also squares it, so the
      // separate "norms" stage and its grid barrier are skipped (stage_enabled)
      const bool fold_norms = norms_in_reduce(a);
      FRL_PAR(t) { red0[t] = 0.f; red1[t] = 0.f; }
      FRL_SYNC();
      for (int q0 = c.cta * (FRL_NT / 8); q0 < nq; q0 += c.ncta * (FRL_NT / 8)) {
        FRL_PAR(t) {
          const int sub = t & 7, q = q0 + (t >> 3);
          float4 sgm = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < nq) {
            const int c0 = sub * per, c1 = (c0 + per) < ncontrib ? (c0 + per) : ncontrib;
            const float* src = a.gpart + (size_t)c0 * N.n_p + 4 * q;
#pragma unroll 4
            for (int cc = c0; cc < c1; ++cc, src += N.n_p) sgm = f4add(sgm, ld4(src));
          }
          st4(r4 + 4 * t, sgm);
        }
        FRL_SYNC();
        FRL_PAR(t) {
          const int q = q0 + (t >> 3);
          if ((t & 7) == 0 && q < nq) {
            float4 sgm = ld4(r4 + 4 * t);
            for (int l = 1; l < 8; ++l) sgm = f4add(sgm, ld4(r4 + 4 * (t + l)));
            st4(gdst + 4 * q, sgm);
            if (fold_norms) {
              const float qq = sgm.x * sgm.x + sgm.y * sgm.y + sgm.z * sgm.z + sgm.w * sgm.w;
              if (is_critic(N, 4 * q)) red1[t] += qq; else red0[t] += qq;
            }
          }
        }
        FRL_SYNC();
      }
      if (fold_norms) {
        const float ta = block_sum(red0), tc = block_sum(red1);
        FRL_PAR(t) { if (t == 0) { a.sumsq[c.cta * 2] = ta; a.sumsq[c.cta * 2 + 1] = tc; } }
        FRL_SYNC();
      }
#ifndef FRL_EMUL
      if (a.dp.world > 1) __threadfence_system();            // the peers read this block over NVLink after the exchange flag
#endif
    } else if (s == 2) {
      // (data-parallel: net.g now holds the all-reduced SUM over ranks) scale, then separate actor / critic sum-of-squares
      const float gs = a.grad_scale > 0.f ? a.grad_scale : 1.f;
      FRL_PAR(t) {
        float la = 0.f, lc = 0.f;
        for (int p = (c.cta * FRL_NT + t) * 4; p < N.n_p; p += c.ncta * FRL_NT * 4) {
          float4 sgm = ld4(N.g + p);
          if (gs != 1.f) { sgm.x *= gs; sgm.y *= gs; sgm.z *= gs; sgm.w *= gs; st4(N.g + p, sgm); }
          const float q = sgm.x * sgm.x + sgm.y * sgm.y + sgm.z * sgm.z + sgm.w * sgm.w;
          if (is_critic(N, p)) lc += q; else la += q;
        }
        red0[t] = la; red1[t] = lc;
      }
      FRL_SYNC();
      const float ta = block_sum(red0), tc = block_sum(red1);
      FRL_PAR(t) { if (t == 0) { a.sumsq[c.cta * 2] = ta; a.sumsq[c.cta * 2 + 1] = tc; } }
      FRL_SYNC();
    } else {
      // optimiser scalars (broadcast through smem).  The per-CTA partials are fetched by one thread EACH and folded in CTA
      // order from shared memory (same association as a serial walk; a single thread walking 148 partials in global memory
      // cost 7 us per sum while its CTA, and with it the whole grid, waited at the next barrier).
      float* sh = c.red;
      float nrm[3] = {0.f, 0.f, 0.f}, met[3] = {0.f, 0.f, 0.f};
      const int rep = (a.optimizer == FRL_OPT_ADAM && a.opt_repeat > 1) ? 2 : 1;
      // pass 2 of the cautious optimiser runs right after pass 1 in the same launch: the scalars pass 1 left in sh[0..8] are still there
      const bool reuse = s == 4 && (a.stage_hi == 0 || a.stage_lo <= 3);
      if (!reuse) {
        // the per-CTA sums of squares, folded by 2 x 8 threads (contiguous eighths in CTA order, then the eighths in order)
        float* stg2 = c.red + 64;
        FRL_PAR(t) { for (int i = t; i < 2 * c.ncta; i += FRL_NT) stg2[i] = a.sumsq[i]; }
        FRL_SYNC();
        FRL_PAR(t) {
          if (t < 16) {
            const int k = t >> 3, sub = t & 7, per8 = (c.ncta + 7) >> 3;
            const int i0 = sub * per8, i1 = (i0 + per8) < c.ncta ? (i0 + per8) : c.ncta;
            float acc = 0.f;
            for (int i = i0; i < i1; ++i) acc += stg2[2 * i + k];
            sh[16 + t] = acc;
          }
        }
        FRL_SYNC();
        FRL_PAR(t) {
          if (t < 2) {
            float acc = 0.f;
            for (int l = 0; l < 8; ++l) acc += sh[16 + t * 8 + l];
            sh[32 + t] = acc;
          }
        }
        FRL_SYNC();
        nrm[0] = sh[32]; nrm[1] = sh[33];
        FRL_SYNC();
      }
      if (s == 3 && c.cta == 0) cta_sums(sh + 64, a.stats, 8, a.stats + 1, 8, a.stats + 2, 8, ncontrib, met);
      FRL_PAR(t) {
        if (t == 0 && !reuse) {
          const float ta = nrm[0], tc = nrm[1];
          float ca = 1.f, cc = 1.f;
          if (a.max_norm_actor > 0.f) ca = fminf(a.max_norm_actor / (sqrtf(ta) + 1e-6f), 1.f);
          if (a.max_norm_critic > 0.f) cc = fminf(a.max_norm_critic / (sqrtf(tc) + 1e-6f), 1.f);
          if (a.max_norm_joint > 0.f) ca = cc = fminf(a.max_norm_joint / (sqrtf(ta + tc) + 1e-6f), 1.f);   // one clip over both nets
          sh[0] = ca; sh[1] = cc;
          // optimiser step k of this update is step number step0 + u * rep + k + 1 (rep = 2: MAPPO_discrete's second step())
          for (int k = 0; k < rep; ++k) {
            double p1 = 1.0, p2 = 1.0, q1 = a.beta1, q2 = a.beta2;          // beta^step by squaring (as make_adam_hp)
            for (unsigned long e = (unsigned long)(a.step0 + (int64_t)u * rep + k + 1); e; e >>= 1) {
              if (e & 1) { p1 *= q1; p2 *= q2; }
              q1 *= q1; q2 *= q2;
            }
            const double bc1 = 1.0 - p1, bc2 = 1.0 - p2;
            if (k == 0) sh[2] = (float)(a.lr * sqrt(bc2) / bc1);         // c_adamw step_size
            sh[3 + 3 * k] = (float)(-(a.lr / bc1));                   // torch Adam: -lr/bc1
            sh[4 + 3 * k] = (float)sqrt(bc2);
            sh[5 + 3 * k] = (float)(-((a.lr_critic > 0.0 ? a.lr_critic : a.lr) / bc1));
          }
          if (s == 3 && c.cta == 0) {
            const float l0 = met[0], l1 = met[1], l2 = met[2];
            const float ent_mean = l2 / (float)rows;
            a.out[u * 8 + 0] = l0 - a.entropy_coef * ent_mean;
            a.out[u * 8 + 1] = l1 / (float)(rows * a.n_adv);
            a.out[u * 8 + 2] = ent_mean;
            a.out[u * 8 + 3] = sqrtf(ta);
            a.out[u * 8 + 4] = sqrtf(tc);
          }
        }
        if (t < FRL_NSEG) segc[t] = 0;
      }
      FRL_SYNC();
      const float coef_a = sh[0], coef_c = sh[1], step_size = sh[2], adam_step = sh[3], bc2s = sh[4], adam_step_c = sh[5];
      const float adam_step1 = sh[6], bc2s1 = sh[7], adam_step_c1 = sh[8];      // second step of opt_repeat 2 (unset otherwise, unused)
      const float b1 = (float)a.beta1, b2 = (float)a.beta2, omb1 = (float)(1.0 - a.beta1), omb2 = (float)(1.0 - a.beta2);
      const float eps = (float)a.eps;
      if (a.optimizer == FRL_OPT_ADAM) {
        if (s == 4) return;
        FRL_PAR(t) {
          for (int p = c.cta * FRL_NT + t; p < N.n_p; p += c.ncta * FRL_NT) {
            const bool crit = is_critic(N, p);
            const float g = N.g[p] * (crit ? coef_c : coef_a);
            float m = N.m[p], v = N.v[p], w = N.p[p];
            for (int k = 0; k < rep; ++k) {
              m = fmaf(omb1, g - m, m);
              v = fadd(fmul(v, b2), fmul(fmul(omb2, g), g));
              const float denom = fadd(fdiv(fsqrt(v), k ? bc2s1 : bc2s), eps);
              w = fadd(w, fdiv(fmul(crit ? (k ? adam_step_c1 : adam_step_c) : (k ? adam_step1 : adam_step), m), denom));
            }
            N.m[p] = m; N.v[p] = v; N.p[p] = w;
            const int mi = mirror_index(N, p);
            if (mi >= 0) N.pt[mi] = w;
#ifndef FRL_EMUL
            if (UM) um_split_one(a, p, w);
#endif
          }
        }
        FRL_SYNC();
        return;
      }
      if (s == 3) {
        // cautious AdamW, pass 1: moments + per-tensor count of (exp_avg * grad > 0)
        FRL_PAR(t) {
          for (int p = c.cta * FRL_NT + t; p < N.n_p; p += c.ncta * FRL_NT) {
            const float g = N.g[p] * (is_critic(N, p) ? coef_c : coef_a);
            float m = N.m[p], v = N.v[p];
            m = fmaf(g, omb1, fmul(m, b1));                               // mul_(b1).add_(g, alpha=1-b1)
            v = fadd(fmul(v, b2), fmul(fmul(omb2, g), g));                // mul_(b2).addcmul_(g, g, 1-b2)
            N.m[p] = m; N.v[p] = v;
            if (m * g > 0.f) {
              int numel;
              const int sg = seg_of(N, p, &numel);
#ifndef FRL_EMUL
              atomicAdd(&segc[sg], 1);
#else
              segc[sg] += 1;
#endif
            }
          }
        }
        FRL_SYNC();
        FRL_PAR(t) { if (t < FRL_NSEG) a.segcnt[c.cta * FRL_NSEG + t] = (float)segc[t]; }
        FRL_SYNC();
      } else {
        // pass 2: p += -step * (m * mask / max(mean(mask), 1e-3)) / (sqrt(v) + eps)
        // per-tensor mask counts of all CTAs: staged into shared memory by the whole CTA (one coalesced sweep instead of 13 threads
        // each waiting on ncta / 16 rounds of L2 latency), then folded in CTA order — the same association as before
        float* segmean = red0;
        float* stg = c.red + 64;                 // c.red[0..5] still hold the optimiser scalars
        FRL_PAR(t) { for (int i = t; i < c.ncta * FRL_NSEG; i += FRL_NT) stg[i] = a.segcnt[i]; }
        FRL_SYNC();
        // the counts are small integers held in floats: any association gives the same sum, so eight threads share a tensor
        float* part8 = red1;
        FRL_PAR(t) {
          if (t < FRL_NSEG * 8) {
            const int sg = t >> 3, sub = t & 7, per8 = (c.ncta + 7) >> 3;
            const int i0 = sub * per8, i1 = (i0 + per8) < c.ncta ? (i0 + per8) : c.ncta;
            float tot = 0.f;
            for (int cc = i0; cc < i1; ++cc) tot += stg[cc * FRL_NSEG + sg];
            part8[t] = tot;
          }
        }
        FRL_SYNC();
        FRL_PAR(t) {
          if (t < FRL_NSEG) {
            float tot = 0.f;
            for (int l = 0; l < 8; ++l) tot += part8[t * 8 + l];
            segmean[t] = tot;
          }
        }
        FRL_SYNC();
        FRL_PAR(t) {
          for (int p = c.cta * FRL_NT + t; p < N.n_p; p += c.ncta * FRL_NT) {
            const float g = N.g[p] * (is_critic(N, p) ? coef_c : coef_a);
            const float m = N.m[p], v = N.v[p];
            float w = N.p[p];
            int numel;
            const int sg = seg_of(N, p, &numel);
            if (numel > 0) {
              const float mean = fmaxf(fdiv(segmean[sg], (float)numel), 1e-3f);
              const float mask = (m * g > 0.f) ? fdiv(1.f, mean) : 0.f;
              const float denom = fadd(fsqrt(v), eps);
              const float ng = fdiv(fmul(m, mask), denom);
              w = fmaf(ng, -step_size, w);
            }
            N.p[p] = w;
            const int mi = mirror_index(N, p);
            if (mi >= 0) N.pt[mi] = w;
#ifndef FRL_EMUL
            if (UM) um_split_one(a, p, w);
#endif
          }
        }
        FRL_SYNC();
      }
    }
  }
};
typedef PpoAlgoT<8> PpoAlgo;

// ------------------------------------------------------------------------------------------------
// GAE: one warp per env column, lanes own contiguous time chunks, float64 composition of the affine maps
// A_t = b_t + a_t * A_{t+1}  via warp shuffles (segmented by construction: a_t = 0 where adv_done).
// ------------------------------------------------------------------------------------------------
struct GaeArgs {
  const float *reward, *done, *adv_done, *vs, *vs_next;
  int T, N;
  double gamma, lmbda;
  float *adv, *v_target;
};

struct GaeAlgo {
  typedef GaeArgs Args;
  static const int NSTAGES = 1;
  FRL_SHD int wbuf_floats(const Args&) { return 32; }
  FRL_SHD int user_floats(const Args&) { return 6 * (FRL_NT) + 64; }
  FRL_SHD int grid(const Args& a, int) { const int wpb = FRL_NT / 32; return (a.N + wpb - 1) / wpb; }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV void stage(int, int, Cta& c, float* user, const Args& a) {
    const int wpb = FRL_NT / 32;
    double* sa = (double*)user;            // [FRL_NT] per-lane chunk product
    double* sbv = sa + FRL_NT;             // [FRL_NT] per-lane chunk offset
    const int chunk = (a.T + 31) / 32;
    const float g32 = (float)a.gamma;
    // pass 1: every lane folds its chunk [t0,t1) into (P, S):  A_{t0} = S + P * A_{t1}
    FRL_PAR(t) {
      const int col = c.cta * wpb + (t >> 5), lane = t & 31;
      double P = 1.0, S = 0.0;
      if (col < a.N) {
        const int t0 = lane * chunk, t1 = (t0 + chunk < a.T) ? t0 + chunk : a.T;
        for (int k = t1 - 1; k >= t0; --k) {
          const size_t i = (size_t)k * a.N + col;
          const float td = fadd(fadd(a.reward[i], fmul(fmul(g32, fadd(1.f, -a.done[i])), a.vs_next[i])), -a.vs[i]);
          const double ak = a.gamma * a.lmbda * (1.0 - (double)a.adv_done[i]);
          S = (double)td + ak * S;          // A_k = td_k + a_k * A_{k+1}
          P = ak * P;
        }
      }
      sa[t] = P; sbv[t] = S;
    }
    FRL_SYNC();
    // pass 2: suffix composition of the per-lane affine maps across the 32 lanes of a column.
    // GPU: Hillis-Steele suffix scan with warp shuffles ((P,S) o (P',S') = (P*P', S + P*S')), then the incoming value
    // of lane l is the inclusive result of lane l+1.  (The emulation build walks the lanes sequentially.)
    double* tailv = (double*)user + 2 * FRL_NT;
#ifndef FRL_EMUL
    {
      const int t = (int)threadIdx.x, lane = t & 31;
      double P = sa[t], S = sbv[t];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const double P2 = __shfl_down_sync(0xffffffffu, P, off), S2 = __shfl_down_sync(0xffffffffu, S, off);
        if (lane + off < 32) { S = S + P * S2; P = P * P2; }
      }
      const double nxt = __shfl_down_sync(0xffffffffu, S, 1);
      tailv[t] = (lane == 31) ? 0.0 : nxt;
    }
#else
    FRL_PAR(t) {
      const int lane = t & 31, w0 = t & ~31;
      double tail = 0.0;                    // A at the start of the chunk after this lane's
      for (int l = 31; l > lane; --l) tail = sbv[w0 + l] + sa[w0 + l] * tail;
      tailv[t] = tail;
    }
#endif
    FRL_SYNC();
    // pass 3: re-walk the chunk with the true incoming value and write adv / v_target
    FRL_PAR(t) {
      const int col = c.cta * wpb + (t >> 5), lane = t & 31;
      if (col < a.N) {
        const int t0 = lane * chunk, t1 = (t0 + chunk < a.T) ? t0 + chunk : a.T;
        double A = tailv[t];
        for (int k = t1 - 1; k >= t0; --k) {
          const size_t i = (size_t)k * a.N + col;
          const float td = fadd(fadd(a.reward[i], fmul(fmul(g32, fadd(1.f, -a.done[i])), a.vs_next[i])), -a.vs[i]);
          A = (double)td + a.gamma * a.lmbda * A * (1.0 - (double)a.adv_done[i]);
          const float af = (float)A;
          a.adv[i] = af;
          a.v_target[i] = fadd(af, a.vs[i]);
        }
      }
    }
    FRL_SYNC();
  }
};

// ------------------------------------------------------------------------------------------------
// GAE for N >= 32 env columns: a CTA owns 32 adjacent columns (lane = column: every row access is one coalesced 128-B
// line) and walks the time axis from the end in ROUNDS of 8 chunks (one warp each, <= GAE_LC steps per chunk).  In a round,
// pass 1 reads the five inputs ONCE, folds each chunk into its affine map A_{t0} = S + P * A_{t1} (float64) and stashes
// (td, 1 - adv_done, vs) in shared memory; the 8 maps of a column are composed through shared memory on top of the carry
// from the previous (later-in-time) round; pass 2 re-walks the chunk from the stash and writes adv / v_target; the value
// at the round's first step is the next round's carry.  The stash is 48 KB whatever T is, so four CTAs fit on an SM and the
// algorithmic traffic (20 B read + 8 B written per element) is also the DRAM traffic.
// ------------------------------------------------------------------------------------------------
#define GAE_LC 16
struct GaeTileAlgo {
  typedef GaeArgs Args;
  static const int MIN_CTAS = 4;
  static const int NCHUNK = FRL_NT / 32;
  FRL_SHD int chunk_len(const Args& a) { const int l = (a.T + NCHUNK - 1) / NCHUNK; return l < GAE_LC ? l : GAE_LC; }
  FRL_SHD int smem_floats(const Args& a) { return 4 * FRL_NT + 128 + 3 * chunk_len(a) * FRL_NT; }
  FRL_SHD int grid(const Args& a) { return (a.N + 31) / 32; }
  FRL_SDEV void run(int cta, int, float* sm, const Args& a) {
    double* sa = (double*)sm;              // [FRL_NT] chunk product
    double* sbv = sa + FRL_NT;             // [FRL_NT] chunk offset
    double* carry = sbv + FRL_NT;          // [2][32] value entering the round from later time steps (double-buffered)
    float* st = sm + 4 * FRL_NT + 128;      // [3][Lc][FRL_NT] stash
    const int Lc = chunk_len(a), R = Lc * NCHUNK;
    const float g32 = (float)a.gamma;
    const double gl = a.gamma * a.lmbda;
    FRL_PAR(t) { if (t < 32) carry[t] = 0.0; }               // zero tail at t = T
    const int rounds = (a.T + R - 1) / R;
    for (int r = 0; r < rounds; ++r) {
      // rounds are aligned to the END of the rollout: round r covers [hi - R, hi), the first one may be short at the front
      const int hi = a.T - r * R, lo = hi - R > 0 ? hi - R : 0;
      FRL_PAR(t) {
        const int col = cta * 32 + (t & 31), w = t >> 5;
        double P = 1.0, S = 0.0;
        const int t0 = lo + w * Lc, t1 = (t0 + Lc < hi) ? t0 + Lc : hi;
        if (col < a.N) {
#pragma unroll 4
          for (int k = t1 - 1; k >= t0; --k) {
            const size_t i = (size_t)k * a.N + col;
            const float v = a.vs[i];
            const float td = fadd(fadd(a.reward[i], fmul(fmul(g32, fadd(1.f, -a.done[i])), a.vs_next[i])), -v);
            const float om = 1.f - a.adv_done[i];
            const double ak = gl * (double)om;
            S = (double)td + ak * S;
            P = ak * P;
            float* q = st + (size_t)(k - t0) * FRL_NT + t;
            q[0] = td; q[(size_t)Lc * FRL_NT] = om; q[(size_t)2 * Lc * FRL_NT] = v;
          }
        }
        sa[t] = P; sbv[t] = S;
      }
      FRL_SYNC();
      FRL_PAR(t) {
        const int col = cta * 32 + (t & 31), lane = t & 31, w = t >> 5;
        const int t0 = lo + w * Lc, t1 = (t0 + Lc < hi) ? t0 + Lc : hi;
        if (col < a.N && t0 < t1) {
          double A = carry[(r & 1) * 32 + lane];
          for (int cc = NCHUNK - 1; cc > w; --cc) A = sbv[cc * 32 + lane] + sa[cc * 32 + lane] * A;
#pragma unroll 4
          for (int k = t1 - 1; k >= t0; --k) {
            const size_t i = (size_t)k * a.N + col;
            const float* q = st + (size_t)(k - t0) * FRL_NT + t;
            const float td = q[0], om = q[(size_t)Lc * FRL_NT], v = q[(size_t)2 * Lc * FRL_NT];
            A = (double)td + gl * (double)om * A;
            const float af = (float)A;
            a.adv[i] = af;
            a.v_target[i] = fadd(af, v);
          }
          if (w == 0) carry[((r + 1) & 1) * 32 + lane] = A;     // value at the round's first step -> next round's carry
        }
      }
      FRL_SYNC();
    }
  }
};
