// Fused actor-critic learn() for the off-policy family:
//   SAC    SAC_file/SAC.py:222-271 (Actor :60-97, Critic :103-127, Alpha :154-169, Agent.update_* :141-151)
//   TD3    TD3_file/TD3.py:189-244          DDPG   DDPG_file/DDPG.py:203-233
//   MADDPG MADDPG_file/MADDPG.py:186-237    (n_agents > 1: centralised critic on cat(all obs, all actions))
// One persistent cooperative launch runs n_updates sequential learns; 7 grid-wide stages per learn (the actor stages 4-6
// and their barriers are skipped on TD3's off steps):
//   0 targets        [Batch_ObsNorm: fold the batch mean of obs into the running statistics, normalise obs / next_obs]
//            gather rows, a' = actor_target(s') (every agent), target critic head[role] -> exchange buffer
//   1 critic         y from both heads' target values, critic head[role] forward + backward -> gradient partials
//   2 reduce         fixed-order cross-CTA reduction (per head: only that role's CTAs) + sum of squares
//   3 optimiser      clip_grad_norm_ + Adam (+ critic-target Polyak)
//   4 actor          actor forward, critic head[role] forward + dQ/da with the UPDATED critic, actor backward
//   5 reduce         actor gradient
//   6 optimiser      clip + Adam + actor-target Polyak + temperature step
// CTA roles: a twin critic is split head-per-CTA (cta = tile_slot * n_heads + role), which doubles the SMs in use and
// cuts the sequential layer ops per CTA from 45 to 28; the target action / actor forward are recomputed per role.
#pragma once
#include "algo_dqn.cuh"

#define FRL_LOG_SQRT_2PI 0.91893853320467274178f
#define FRL_LOG2F 0.69314718055994530942f

FRL_DEV float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // F.softplus(beta=1, threshold=20)

// Cross-CTA reduction of gradient partials.  Partial of CTA k lives at gpart + k*stride.  With split_heads the critic
// block's head h (layers 3h..3h+2) only receives the partials of the CTAs with role h (k = h, h+nrole, ...); otherwise
// every parameter sums `cnt` partials starting at CTA 0 with the given CTA step.
FRL_NI_OPT void reduce_grads_roles(int cta, int ncta, float* slot, const frl_net_t& n, const float* gpart, int stride, int nslots,
                                   int nrole, int split_heads, int cta_step, float* sumsq_part) {
  // Two threads per 16-B parameter group, each summing half of the partials with all its loads in flight (the partials
  // were just written by other SMs: every dependent L2 round trip costs ~0.7 us), halves combined in fixed order.
  float* pair = slot + FRL_NT;                       // [FRL_NT][4]
  const int ngroups = n.n_p >> 2;
  FRL_PAR(t) { slot[t] = 0.f; }
  FRL_SYNC();
  for (int base = cta * (FRL_NT / 2); base < ngroups; base += ncta * (FRL_NT / 2)) {
    FRL_PAR(t) {
      const int g = base + (t >> 1), half = t & 1, p = g * 4;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < ngroups) {
        int first = 0, step = cta_step, cnt = nslots * (nrole / cta_step);
        if (split_heads) {
          int h = 0;
          for (int li = 3; li < n.n_layers; li += 3) if (p >= n.L[li].w_off) h = li / 3;
          first = h; step = nrole; cnt = nslots;
        }
        const int mid = (cnt + 1) >> 1;
        int k = half ? mid : 0;
        const int k1 = half ? cnt : mid;
        for (; k + 16 <= k1; k += 16) {
          float4 v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = ld4(gpart + (size_t)(first + (k + i) * step) * stride + p);
#pragma unroll
          for (int i = 0; i < 16; ++i) s = f4add(s, v[i]);
        }
        for (; k + 4 <= k1; k += 4) {
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = ld4(gpart + (size_t)(first + (k + i) * step) * stride + p);
#pragma unroll
          for (int i = 0; i < 4; ++i) s = f4add(s, v[i]);
        }
        for (; k < k1; ++k) s = f4add(s, ld4(gpart + (size_t)(first + k * step) * stride + p));
      }
      st4(pair + 4 * t, s);
    }
    FRL_SYNC();
    FRL_PAR(t) {
      const int g = base + (t >> 1);
      if (!(t & 1) && g < ngroups) {
        const float4 s = f4add(ld4(pair + 4 * t), ld4(pair + 4 * t + 4));
        st4(n.g + g * 4, s);
        slot[t] += s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;
      }
    }
    FRL_SYNC();
  }
  float tot = block_sum(slot);
  FRL_PAR(t) { if (t == 0 && sumsq_part) sumsq_part[cta] = tot; }
  FRL_SYNC();
}

// SPEC selects a compile-time specialisation of the same source (the generic kernel is 320 KB of SASS — more than the
// instruction cache holds, and a learn walks most of it; the specialisations drop the multi-agent loops, Batch_ObsNorm and
// the branches of the other actor kind):  0 generic (multi-agent / Batch_ObsNorm / anything) · 1 single-agent SAC (twin
// critic, no Batch_ObsNorm) · 2 single-agent deterministic actor (TD3 / DDPG, no Batch_ObsNorm).  Same arithmetic, same order.
template <int SPEC>
struct AcAlgoT {
  typedef frl_ac_args_t Args;
  static const int NSTAGES = 7;
  FRL_SHD bool multi(const Args& a) { return SPEC == 0 && a.n_agents > 1; }
  FRL_SHD bool has_bon(const Args& a) { return SPEC == 0 && a.obs_norm[0] != nullptr; }
  FRL_SHD bool is_sac(const Args& a) { return SPEC == 1 ? true : (SPEC == 2 ? false : a.actor_kind == FRL_ACTOR_SAC); }
  FRL_SHD int heads(const Args& a) { return SPEC == 1 ? 2 : a.n_heads; }
  FRL_SHD bool writes_params(int s) { return s == 3 || s == 6; }   // the two Adam / Polyak stages
  FRL_SHD bool is_policy_step(const Args& a, int u) { return ((a.total_it0 + u + 1) % (a.policy_freq > 0 ? a.policy_freq : 1)) == 0; }
  FRL_SHD bool stage_enabled(int s, int u, const Args& a) {
    if (s >= 4) return is_policy_step(a, u);                // TD3 twin_delay: no actor stages (nor their barriers) on off steps
    return true;
  }
  FRL_SHD int norm_floats(const Args& a) { return has_bon(a) ? ((3 * obs_off(a, ma_n(a)) + 3) & ~3) : 0; }

  FRL_SHD int max_layer_floats(const frl_net_t& n) {
    int mx = 0;
    for (int i = 0; i < n.n_layers; ++i) {
      int f = wt_floats(n.L[i]);
      if (f > mx) mx = f;
    }
    return mx;
  }
  // floats of the largest 3-layer head image (heads are contiguous in the mirror)
  FRL_SHD int max_head_floats(const frl_net_t& n) {
    int mx = 0;
    for (int h = 0; h + 3 <= n.n_layers; h += 3) {
      int f = wt_floats(n.L[h]) + wt_floats(n.L[h + 1]) + wt_floats(n.L[h + 2]);
      if (f > mx) mx = f;
    }
    return mx;
  }
  FRL_SHD int stream_floats(const Args& a) {
    int x = max_layer_floats(a.actor), y = max_layer_floats(a.critic);
    for (int j = 0; j < ma_n(a); ++j) { int v = max_layer_floats(tnet(a, j)); if (v > x) x = v; }
    return ((x > y ? x : y) + 31) & ~31;
  }
  FRL_SHD int head_floats(const Args& a) {
    int x = max_head_floats(a.actor), y = max_head_floats(a.critic);
    for (int j = 0; j < ma_n(a); ++j) { int v = max_head_floats(tnet(a, j)); if (v > x) x = v; }
    return ((x > y ? x : y) + 31) & ~31;
  }
  // Resident mode (two whole heads in smem, weights kept across stages) whenever it fits; else per-layer streaming.
  FRL_SHD bool resident(const Args& a) {
    return (size_t)(cta_base_floats(head_floats(a)) + user_floats(a)) * 4 + 64 <= (size_t)227 * 1024;
  }
  FRL_SHD int wbuf_floats(const Args& a) { return resident(a) ? head_floats(a) : stream_floats(a); }
  // ---- multi-agent helpers: where agent j's obs / action sit inside the joint critic input [obs_1..obs_N | act_1..act_N]
  FRL_SHD int ma_n(const Args& a) { return multi(a) ? a.n_agents : 1; }
  FRL_SHD const frl_replay_t& rep(const Args& a, int j) { return multi(a) ? a.ma_replay[j] : a.replay; }
  FRL_SHD const frl_net_t& tnet(const Args& a, int j) { return multi(a) ? a.ma_actor_target[j] : a.actor_target; }
  FRL_SHD int obs_off(const Args& a, int j) { int o = 0; for (int k = 0; k < j; ++k) o += rep(a, k).obs_dim; return o; }
  FRL_SHD int act_off(const Args& a, int j) { int o = obs_off(a, ma_n(a)); for (int k = 0; k < j; ++k) o += rep(a, k).act_dim; return o; }
  FRL_SHD int raw_off(const Args& a, int j) { int o = 0; for (int k = 0; k < j; ++k) o += FRL_R * rep(a, k).row_floats; return o; }
  FRL_SHD int max_aip(const Args& a) { int m = a.actor.L[0].in_pad; for (int j = 0; j < ma_n(a); ++j) { int v = tnet(a, j).L[0].in_pad; if (v > m) m = v; } return m; }
  FRL_SHD int max_ap(const Args& a) { int m = a.actor.L[2].out_pad; for (int j = 0; j < ma_n(a); ++j) { int v = tnet(a, j).L[2].out_pad; if (v > m) m = v; } return m; }

  FRL_SHD int user_floats(const Args& a) {
    const int ldh = act_ld(a.critic.L[0].out_pad), sa = act_ld(a.critic.L[0].in_pad), ap = max_ap(a);
    return raw_off(a, ma_n(a)) + FRL_R * (max_aip(a) + 3 * sa + 6 * ldh + 6 * ap + 3 * 4 + 16) + 2 * FRL_NT + 128 + PLAN_FLOATS + norm_floats(a);
  }
  FRL_SHD int nslots_of(const Args& a, int max_ctas) {
    const int tiles = (a.B + FRL_R - 1) / FRL_R, cap = max_ctas / (heads(a) > 0 ? heads(a) : 1);
    return tiles < cap ? tiles : (cap > 0 ? cap : 1);
  }
  // Worker CTAs (row tile x critic head) run the GEMM stages; with FRL_AC_GRID > 0 the grid is padded with HELPER CTAs
  // (up to FRL_AC_GRID, capped by the SM count) that only take part in the cross-CTA reduce / Adam / Polyak stages, whose
  // parameter sweep is partitioned over ALL CTAs.
#ifndef FRL_AC_GRID
#define FRL_AC_GRID 96
#endif
  FRL_SHD int grid(const Args& a, int max_ctas) {
    const int w = nslots_of(a, max_ctas) * heads(a), cap = FRL_AC_GRID < max_ctas ? FRL_AC_GRID : max_ctas;
    return w > cap ? w : cap;
  }
  FRL_SHD int n_updates(const Args& a) { return a.n_updates; }

  FRL_SDEV float noise_at(const float* ptr, const Args& a, int u, int row, int j, uint32_t stream) {
    if (ptr) return ptr[((size_t)u * a.B + row) * a.replay.act_dim + j];
    return randn_ni(a.seed, stream, (uint32_t)(a.total_it0 + u), (uint32_t)(row * a.replay.act_dim + j));
  }

  // Launch-invariant geometry, computed once (thread 0, first stage call) into shared memory: evaluating the multi-agent
  // offset helpers inline at every use had grown to ~1/4 of the kernel's instructions.
  struct Plan {
    int NA, ai, ap, aip, res, gstride;
    int obs_off[FRL_MAX_AGENTS + 1], act_off[FRL_MAX_AGENTS + 1], raw_off[FRL_MAX_AGENTS + 1];
  };
  static const int PLAN_FLOATS = 32;
  FRL_SDEV void fill_plan(Plan& P, const Args& a) {
    P.NA = ma_n(a); P.ai = multi(a) ? a.agent_index : 0;
    P.ap = max_ap(a); P.aip = max_aip(a); P.res = resident(a) ? 1 : 0;
    P.gstride = a.critic.n_p > a.actor.n_p ? a.critic.n_p : a.actor.n_p;
    int o = 0, r = 0;
    for (int j = 0; j < P.NA; ++j) { P.obs_off[j] = o; P.raw_off[j] = r; o += rep(a, j).obs_dim; r += FRL_R * rep(a, j).row_floats; }
    P.obs_off[P.NA] = o; P.raw_off[P.NA] = r;
    for (int j = 0; j < P.NA; ++j) { P.act_off[j] = o; o += rep(a, j).act_dim; }
    P.act_off[P.NA] = o;
  }

  // Batch_ObsNorm: (x - mean) / (std + 1e-8) on the obs and next_obs columns of the gathered rows, in place.
  FRL_SDEV void norm_rows(const Args& a, const Plan& P, const float* NORM, float* raw0, int nvalid) {
    const int tot = P.obs_off[P.NA];
    FRL_PAR(t) {
      for (int e = t; e < 2 * tot; e += FRL_NT) {
        const int col = e < tot ? e : e - tot;
        int j = 0;
        while (j + 1 < P.NA && col >= P.obs_off[j + 1]) ++j;
        const frl_replay_t& rj = rep(a, j);
        const int k = col - P.obs_off[j];
        const float mean = NORM[3 * P.obs_off[j] + k], den = fadd(NORM[3 * P.obs_off[j] + 2 * rj.obs_dim + k], 1e-8f);
        float* x = raw0 + P.raw_off[j] + (e < tot ? 0 : rb_col_nobs(rj)) + k;
        for (int r = 0; r < nvalid; ++r) x[r * rj.row_floats] = fdiv(fadd(x[r * rj.row_floats], -mean), den);
      }
    }
    FRL_SYNC();
  }

  // Batch_ObsNorm statistics of one learn: x_bar = mean over the B sampled rows of every obs column, folded into the running
  // {mean, S, std} (RunningMeanStd_batch_size.update, SAC.py:399-411).  Every CTA computes the same numbers redundantly from
  // global memory (the rows are L2 hits: the batch gather touches them in the same stage); CTA 0 publishes the new state in
  // stage 1, after the barrier that ends everyone's reads of the old one.
  // Summation ORDER follows torch's CPU `x.mean(dim=0)` = sum_out / B (aten/src/ATen/native/cpu/SumKernel.cpp cascade_sum,
  // AVX2 build as run in this container, SURVEY 7.3-3 "parity is defined against the oracle as run here"):
  //   * columns inside a full group of 32 (obs_dim >= 8) or of 4 (obs_dim < 8) take `multi_row_sum`: rows in order, 16-row
  //     blocks cascading into levels every 16 / 256 blocks, remainder rows in level 0, levels added low to high;
  //   * the other columns take `row_sum`: four interleaved lanes of floor(B/4) rows (row = k + 4 i), each lane summed like
  //     above, then the B % 4 leftover rows added to lane 0, then the lanes in order.
  // The normalisation divides by std = sqrt(S/n) of batch-MEAN differences, so a different association shows up as 1e-3
  // relative loss differences at n = 2; with the same association the statistics are bit-identical.
  FRL_SDEV bool bon_col_seq(int od, int k) { return od >= 8 ? k < (od & ~31) : k < (od & ~3); }
  FRL_SDEV void bon_update(Cta& c, const Args& a, const Plan& P, float* NORM, float* scratch, int u) {
    const int NA = P.NA, tot = P.obs_off[NA], npair = tot * 4, B = a.B;
    const int64_t* idx = a.indices + (size_t)u * B;
    // scratch: part[npair][nbp] block sums | lvl[npair][4] cascade levels      (6 * FRL_R * ldh floats available)
    const int avail = 6 * FRL_R * act_ld(a.critic.L[0].out_pad) - npair * 4;
    int nbp = avail / npair;
    if (nbp > 16) nbp = 16;
    float* part = scratch;
    float* lvl = scratch + npair * nbp;
    FRL_PAR(t) { for (int i = t; i < npair * 4; i += FRL_NT) lvl[i] = 0.f; }
    const int max_blocks = B >> 4;
    for (int b0 = 0; b0 < max_blocks; b0 += nbp) {
      FRL_SYNC();
      FRL_PAR(t) {
        for (int it = t; it < npair * nbp; it += FRL_NT) {
          const int pair = it / nbp, bl = b0 + it % nbp, col = pair >> 2, k = pair & 3;
          int j = 0;
          while (j + 1 < NA && col >= P.obs_off[j + 1]) ++j;
          const frl_replay_t& rj = rep(a, j);
          const bool seq = bon_col_seq(rj.obs_dim, col - P.obs_off[j]);
          const int lane_rows = seq ? (k == 0 ? B : 0) : (B >> 2), first = seq ? 0 : k, stride = seq ? 1 : 4;
          if (bl >= (lane_rows >> 4)) continue;
          const float* base = rj.storage + (col - P.obs_off[j]);
          int64_t ix[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) ix[q] = idx[first + stride * (16 * bl + q)];
          float v[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = base[(size_t)ix[q] * rj.row_floats];
          float sm = v[0];
#pragma unroll
          for (int q = 1; q < 16; ++q) sm += v[q];
          part[pair * nbp + it % nbp] = sm;
        }
      }
      FRL_SYNC();
      FRL_PAR(t) {
        for (int pair = t; pair < npair; pair += FRL_NT) {
          const int col = pair >> 2, k = pair & 3;
          int j = 0;
          while (j + 1 < NA && col >= P.obs_off[j + 1]) ++j;
          const bool seq = bon_col_seq(rep(a, j).obs_dim, col - P.obs_off[j]);
          const int nblk = (seq ? (k == 0 ? B : 0) : (B >> 2)) >> 4;
          float a1 = lvl[pair * 4 + 1], a2 = lvl[pair * 4 + 2], a3 = lvl[pair * 4 + 3];
          for (int q = 0; q < nbp && b0 + q < nblk; ++q) {
            const int i = (b0 + q + 1) << 4;               // lane rows consumed so far
            a1 += part[pair * nbp + q];
            if ((i & 0xF0) == 0) { a2 += a1; a1 = 0.f; if ((i & 0xF00) == 0) { a3 += a2; a2 = 0.f; } }
          }
          lvl[pair * 4 + 1] = a1; lvl[pair * 4 + 2] = a2; lvl[pair * 4 + 3] = a3;
        }
      }
    }
    FRL_SYNC();
    FRL_PAR(t) {                // remainder rows of each lane (< 16), the levels low to high, row_sum's leftover rows
      for (int pair = t; pair < npair; pair += FRL_NT) {
        const int col = pair >> 2, k = pair & 3;
        int j = 0;
        while (j + 1 < NA && col >= P.obs_off[j + 1]) ++j;
        const frl_replay_t& rj = rep(a, j);
        const bool seq = bon_col_seq(rj.obs_dim, col - P.obs_off[j]);
        const int lane_rows = seq ? (k == 0 ? B : 0) : (B >> 2), first = seq ? 0 : k, stride = seq ? 1 : 4;
        const float* base = rj.storage + (col - P.obs_off[j]);
        float a0 = 0.f;
        for (int r = lane_rows & ~15; r < lane_rows; ++r) a0 += base[(size_t)idx[first + stride * r] * rj.row_floats];
        float res = ((a0 + lvl[pair * 4 + 1]) + lvl[pair * 4 + 2]) + lvl[pair * 4 + 3];
        if (!seq && k == 0) for (int r = B & ~3; r < B; ++r) res += base[(size_t)idx[r] * rj.row_floats];
        lvl[pair * 4] = res;
      }
    }
    FRL_SYNC();
    const long n = (long)(a.obs_norm_n0 + u + 1);
    FRL_PAR(t) {
      for (int col = t; col < tot; col += FRL_NT) {
        int j = 0;
        while (j + 1 < NA && col >= P.obs_off[j + 1]) ++j;
        const int k = col - P.obs_off[j], odj = rep(a, j).obs_dim;
        const float sum = ((lvl[col * 16] + lvl[col * 16 + 4]) + lvl[col * 16 + 8]) + lvl[col * 16 + 12];
        const float xb = fdiv(sum, (float)B);
        const float* st = a.obs_norm[j];
        float mean, S = st[odj + k], sd;
        if (n == 1) { mean = xb; sd = xb; }                    // first call: mean = std = x_bar (reference quirk)
        else {
          const float old = st[k];
          mean = fadd(old, fdiv(fadd(xb, -old), (float)n));
          S = fadd(S, fmul(fadd(xb, -old), fadd(xb, -mean)));
          sd = fsqrt(fdiv(S, (float)n));
        }
        float* o = NORM + 3 * P.obs_off[j];
        o[k] = mean; o[odj + k] = S; o[2 * odj + k] = sd;
      }
    }
    FRL_SYNC();
    (void)c;
  }

  FRL_SDEV void stage(int s, int u, Cta& c, float* user, const Args& a) {
    const bool bon = has_bon(a);
    const frl_net_t& A = a.actor;
    const frl_net_t& C = a.critic;
    const frl_replay_t& rb = a.replay;
    const int od = rb.obs_dim, ad = rb.act_dim, rf = rb.row_floats;
    Plan& P = *reinterpret_cast<Plan*>(user);
    user += PLAN_FLOATS;
    if (s == 0 && u == 0) {
      FRL_PAR(t) { if (t == 0) fill_plan(P, a); }
      FRL_SYNC();
    }
    float* NORM = user;                        // [agent j at 3*obs_off[j]] {mean, S, std} of this learn (Batch_ObsNorm)
    user += norm_floats(a);
    const int ldh = act_ld(C.L[0].out_pad), sa = act_ld(C.L[0].in_pad), ap = P.ap;   // strides (bank-conflict free), not widths
    const int NA = SPEC != 0 ? 1 : P.NA, ai = SPEC != 0 ? 0 : P.ai;
    const int aip = P.aip;
    const int ntile = (a.B + FRL_R - 1) / FRL_R;
    const int nrole = heads(a), role = c.cta % nrole, nslots = nslots_of(a, c.ncta);
    const bool helper = c.cta >= nslots * nrole;               // reduce / optimiser stages only
    const int slot = helper ? ntile : c.cta / nrole;           // helpers own no row tile: every tile loop is empty
    const int l0 = 3 * role;                                   // this CTA's critic head = layers l0..l0+2
    const bool one_tile = ntile <= nslots;
    const float invB = 1.0f / (float)a.B;
    const bool sac = is_sac(a);
    const bool smoothing = SPEC != 1 && a.target_smoothing;
    const bool policy_step = is_policy_step(a, u);
    const long n_policy_before = (a.policy_freq > 1) ? (long)((a.total_it0 + u) / a.policy_freq - a.total_it0 / a.policy_freq) : (long)u;
    const int heads_used = sac ? nrole : 1;                // actor loss: SAC mean of both heads, TD3 Q1 only, DDPG single
    float alpha = 0.f;
    if (sac) alpha = expf(a.alpha_state[0]);
    const bool res = P.res != 0;

    SmemBump sb; sb.p = user;
    float* raw0 = sb.take(P.raw_off[NA]);   // gathered rows of every agent's replay (same indices)
    float* raw = raw0 + P.raw_off[ai];      // this agent's rows (reward / done come from here)
    float* XA = sb.take(FRL_R * aip);        // actor input (one agent's obs / next_obs)
    float* XS = sb.take(FRL_R * sa);         // [obs | act]
    float* XN = sb.take(FRL_R * sa);         // [next_obs | a']   (stage 4: [obs | pi(obs)])
    float* dXP = sb.take(FRL_R * sa);        // dQ/d[obs|a]
    float* H1 = sb.take(FRL_R * ldh);        // critic head activations
    float* H2 = sb.take(FRL_R * ldh);
    float* A1 = sb.take(FRL_R * ldh);        // actor activations
    float* A2 = sb.take(FRL_R * ldh);
    float* D1 = sb.take(FRL_R * ldh);
    float* D2 = sb.take(FRL_R * ldh);
    float* MU = sb.take(FRL_R * ap);         // actor output (pre-tanh mean)
    float* UU = sb.take(FRL_R * ap);         // SAC per-dim log-prob terms
    float* AC = sb.take(FRL_R * ap);         // squashed action
    float* dMU = sb.take(FRL_R * ap);
    float* EPS = sb.take(FRL_R * ap);
    float* QA = sb.take(FRL_R * 4);          // head output
    float* dQA = sb.take(FRL_R * 4);
    float* rowv = sb.take(FRL_R * 4);        // per-row scalars
    float* red0 = sb.take(FRL_NT);
    float* red1 = sb.take(FRL_NT);
    const int gstride = P.gstride;
    float* gp = a.gpart + (size_t)c.cta * gstride;

    if (s == 0) {
      if (bon && !helper) {
        bon_update(c, a, P, NORM, H1, u);    // H1..D2 (6 contiguous [R][ldh] tiles) are free scratch here
      }
      // ---------------- target action(s) + this role's target critic head ----------------
      for (int tile = slot; tile < ntile; tile += nslots) {
        const int row0 = tile * FRL_R;
        const int nvalid = (a.B - row0) < FRL_R ? (a.B - row0) : FRL_R;
        // resident plan: actor targets alternate between the slots (next one prefetched behind the current forward), the
        // target critic head goes to the slot the last actor does not use; once that actor is done its slot takes the
        // online critic head stage 1 needs.  Heads still resident from the previous learn / tile are not fetched again.
        const bool last_tile = tile + nslots >= ntile;
        int at0 = -1, cts = -1;
        if (res) {
          const int ctk = res_find(c, a.critic_target, l0);
          at0 = ctk >= 0 ? (ctk ^ 1) : 1;
          res_fetch(c, at0, tnet(a, 0), 0, 3);
          if (NA == 1) res_fetch(c, at0 ^ 1, a.critic_target, l0, 3);
        } else {
          stage_prefetch(c, layer_fwd_src(tnet(a, 0), 0), layer_fwd_bytes(tnet(a, 0).L[0]));
        }
        trace(10);
        for (int j = 0; j < NA; ++j) gather_rows<FRL_R>(rep(a, j).storage, rep(a, j).row_floats, a.indices + (size_t)u * a.B + row0, nvalid, raw0 + P.raw_off[j]);
        if (bon) norm_rows(a, P, NORM, raw0, nvalid);
        trace(11);
        put_cols<FRL_R>(XN, sa, P.act_off[0], raw0, 1, 0, 0, sa - P.act_off[0]);   // zero the action + pad columns of XN
        for (int j = 0; j < NA; ++j) {
          const frl_replay_t& rj = rep(a, j);
          put_cols<FRL_R>(XN, sa, P.obs_off[j], raw0 + P.raw_off[j], rj.row_floats, rb_col_nobs(rj), rj.obs_dim, rj.obs_dim);
        }
        {
          const frl_replay_t& r0 = rep(a, 0);
          const int tip0 = tnet(a, 0).L[0].in_pad;
          put_cols<FRL_R>(XA, tip0, 0, raw0, r0.row_floats, rb_col_nobs(r0), r0.obs_dim, tip0);
        }
        FRL_SYNC();
        for (int j = 0; j < NA; ++j) {                                               // a'_j = actor_target_j(next_obs_j)
          const frl_replay_t& rj = rep(a, j);
          const frl_net_t& T = tnet(a, j);
          const int adj = rj.act_dim, tip = T.L[0].in_pad, aoff = P.act_off[j];
          int sl = -1;
          if (res) {
            sl = at0 ^ (j & 1);
            res_fetch(c, sl, T, 0, 3);
            if (j + 1 < NA) res_fetch(c, sl ^ 1, tnet(a, j + 1), 0, 3);
            else if (NA > 1) res_fetch(c, sl ^ 1, a.critic_target, l0, 3);
            cts = sl ^ 1;
          }
          if (j > 0) {
            put_cols<FRL_R>(XA, tip, 0, raw0 + P.raw_off[j], rj.row_floats, rb_col_nobs(rj), rj.obs_dim, tip);
            FRL_SYNC();
          }
          trace(12);
          mlp_fwd<FRL_R>(c, T, 0, 3, XA, tip, A1, A2, ldh, MU, ap, FRL_ACT_NONE,
                         j + 1 < NA ? fwd_hint(tnet(a, j + 1), 0) : fwd_hint(a.critic_target, l0), sl);
          FRL_PAR(t) {
            if (t < FRL_R * adj) {
              const int r = t / adj, jj = t % adj;
              const float mean = MU[r * ap + jj];
              float act;
              if (sac) {
                const float* ls_p = T.p + T.x_off;
                const float ls = fminf(fmaxf(ls_p[jj], -20.f), 2.f);
                const float sd = expf(ls);
                const float e = (r < nvalid) ? noise_at(a.noise_next, a, u, row0 + r, jj, 1u) : 0.f;
                const float uu = fadd(mean, fmul(e, sd));
                const float diff = uu - mean;
                float lp = -(diff * diff) / (2.f * (sd * sd)) - logf(sd) - FRL_LOG_SQRT_2PI;
                lp -= 2.f * (FRL_LOG2F - uu - softplus_t(-2.f * uu));
                UU[r * ap + jj] = lp;                      // per-dim log-prob contribution
                act = tanhf(uu);
              } else if (smoothing) {
                float e = 0.f;
                if (r < nvalid) {
                  if (multi(a))              // MATD3: every agent's target action has its own randn_like draw
                    e = a.ma_noise_next[j] ? a.ma_noise_next[j][(size_t)(row0 + r) * adj + jj]
                                           : randn_ni(a.seed, 16u + (uint32_t)j, (uint32_t)(a.total_it0 * NA + ai), (uint32_t)((row0 + r) * adj + jj));
                  else
                    e = noise_at(a.noise_next, a, u, row0 + r, jj, 1u);
                }
                float nz = fmul(a.policy_noise_scale, fmul(e, a.policy_noise));
                nz = fminf(fmaxf(nz, -a.noise_clip), a.noise_clip);
                float v = fadd(fmul(tanhf(mean), a.max_action), nz);
                v = fminf(fmaxf(v, -a.max_action), a.max_action);
                act = fdiv(v, a.max_action);
              } else {
                act = tanhf(mean);
              }
              XN[r * sa + aoff + jj] = act;
            }
          }
          trace(13);
          FRL_SYNC();
          trace(14);
        }
        if (res && last_tile) res_fetch(c, cts ^ 1, C, l0, 3);
        mlp_fwd<FRL_R>(c, a.critic_target, l0, 3, XN, sa, H1, H2, ldh, QA, 4, FRL_ACT_NONE, no_hint(), cts);
        // exchange: xchg[row][role] = Q'_role ; xchg[B*nrole + row] = sum_j log pi(a'|s')  (SAC, role 0)
        FRL_PAR(t) {
          if (t < nvalid) {
            a.xchg[(size_t)(row0 + t) * nrole + role] = QA[t * 4];
            if (sac && role == 0) {
              float lp = 0.f;
              for (int j = 0; j < ad; ++j) lp += UU[t * ap + j];
              a.xchg[(size_t)a.B * nrole + row0 + t] = lp;
            }
          }
        }
        trace(15);
        FRL_SYNC();
      }
    } else if (s == 1) {
      // ---------------- TD target, this role's critic head forward + backward ----------------
      float loss_acc = 0.f;
      bool first = true;
      if (bon && c.cta == 0) {
        FRL_PAR(t) {
          if (t < P.obs_off[NA]) {
            int j = 0;
            while (j + 1 < NA && t >= P.obs_off[j + 1]) ++j;
            const int k = t - P.obs_off[j], odj = rep(a, j).obs_dim;
            const float* o = NORM + 3 * P.obs_off[j];
            float* st = a.obs_norm[j];
            st[k] = o[k]; st[odj + k] = o[odj + k]; st[2 * odj + k] = o[2 * odj + k];
          }
        }
      }
      for (int tile = slot; tile < ntile; tile += nslots) {
        const int row0 = tile * FRL_R;
        const int nvalid = (a.B - row0) < FRL_R ? (a.B - row0) : FRL_R;
        int cs = -1;
        if (res) {
          cs = res_find(c, C, l0);
          if (cs < 0) cs = 0;
          res_fetch(c, cs, C, l0, 3);
          if (policy_step && role < heads_used && tile + nslots >= ntile) res_fetch(c, cs ^ 1, A, 0, 3);   // stage 4's actor
        } else {
          stage_prefetch(c, layer_fwd_src(C, l0), layer_fwd_bytes(C.L[l0]));
        }
        // one tile per CTA: the rows gathered by stage 0 of this learn are still in shared memory
        if (!one_tile) {
          for (int j = 0; j < NA; ++j) gather_rows<FRL_R>(rep(a, j).storage, rep(a, j).row_floats, a.indices + (size_t)u * a.B + row0, nvalid, raw0 + P.raw_off[j]);
          if (bon) norm_rows(a, P, NORM, raw0, nvalid);
        }
        for (int j = 0; j < NA; ++j) {
          const frl_replay_t& rj = rep(a, j);
          const float* rw = raw0 + P.raw_off[j];
          put_cols<FRL_R>(XS, sa, P.obs_off[j], rw, rj.row_floats, 0, rj.obs_dim, rj.obs_dim);
          put_cols<FRL_R>(XS, sa, P.act_off[j], rw, rj.row_floats, rb_col_act(rj), rj.act_dim, j == NA - 1 ? sa - P.act_off[j] : rj.act_dim);
        }
        FRL_PAR(t) {
          if (t < FRL_R) {
            const int r = t;
            float y = 0.f;
            if (r < nvalid) {
              float nq = a.xchg[(size_t)(row0 + r) * nrole];
              if (nrole == 2) nq = fminf(nq, a.xchg[(size_t)(row0 + r) * nrole + 1]);
              const float rew = raw[r * rf + rb_col_rew(rb)], dn = raw[r * rf + rb_col_done(rb)];
              if (sac) {
                const float lp = a.xchg[(size_t)a.B * nrole + row0 + r];
                // target = r + gamma*(1-d)*(minQ' + alpha*(-logp'))      (SAC.py:235)
                y = fadd(rew, fmul(fmul(a.gamma, fadd(1.f, -dn)), fadd(nq, fmul(alpha, -lp))));
              } else {
                y = fadd(rew, fmul(fmul(a.gamma, nq), fadd(1.f, -dn)));     // TD3.py:209 / DDPG.py:212 / MADDPG.py:214
              }
            }
            rowv[r * 4 + 0] = y;
          }
        }
        FRL_SYNC();
        mlp_fwd<FRL_R>(c, C, l0, 3, XS, sa, H1, H2, ldh, QA, 4, FRL_ACT_NONE, bwd_hint(C, l0 + 2), cs);
        FRL_PAR(t) {
          float l = 0.f;
          if (t < FRL_R) {
            const int r = t;
            for (int j = 0; j < 4; ++j) dQA[r * 4 + j] = 0.f;
            if (r < nvalid) {
              const float d0 = QA[r * 4] - rowv[r * 4];
              dQA[r * 4] = 2.f * d0 * invB;
              l = d0 * d0;
            }
          }
          red0[t] = l;
        }
        FRL_SYNC();
        loss_acc += block_sum(red0);
        mlp_bwd<FRL_R>(c, C, l0, 3, XS, sa, H1, H2, ldh, dQA, 4, D1, D2, nullptr, 0, gp, !first, no_hint(), cs);
        first = false;
      }
      FRL_PAR(t) { if (t == 0) a.stats[c.cta * 8 + 0] = loss_acc; }
      FRL_SYNC();
    } else if (s == 2) {
      reduce_grads_roles(c.cta, c.ncta, c.red, C, a.gpart, gstride, nslots, nrole, 1, 1, a.sumsq);
    } else if (s == 3) {
      const AdamSpec hp = {a.lr_critic, a.beta1, a.beta2, a.eps, a.wd_critic, (double)a.max_norm, (long)(a.step_critic0 + u + 1)};
      adam_update(c.cta, c.ncta, c.red, C, a.sumsq, c.ncta, hp, (policy_step && !a.defer_polyak) ? &a.critic_target : nullptr, a.tau);
      res_invalidate(c, C);
      if (policy_step && !a.defer_polyak) res_invalidate(c, a.critic_target);
      if (c.cta == 0) {                      // metrics (block-uniform branch)
        float o[3];
        cta_sums(c.red, a.stats, 8, a.sumsq, 1, nullptr, 0, c.ncta, o);
        FRL_PAR(t) {
          if (t == 0) { a.out[u * 8 + 0] = o[0] * invB; a.out[u * 8 + 4] = sqrtf(o[1]); a.out[u * 8 + 2] = alpha; }
        }
        FRL_SYNC();
      }
    } else if (s == 4) {
      if (!policy_step) return;
      // ---------------- actor forward, this role's critic head forward + dQ/da (updated critic), actor backward ----------------
      float loss_acc = 0.f, ent_acc = 0.f;
      bool first = true;
      const bool active = role < heads_used;
      for (int tile = slot; tile < ntile && active; tile += nslots) {
        const int row0 = tile * FRL_R;
        const int nvalid = (a.B - row0) < FRL_R ? (a.B - row0) : FRL_R;
        int as = -1, cs = -1;
        if (res) {
          as = res_find(c, A, 0);
          if (as < 0) as = 0;
          res_fetch(c, as, A, 0, 3);
          cs = as ^ 1;
          res_fetch(c, cs, C, l0, 3);            // the critic head as updated by stage 3, behind the actor forward
        } else {
          stage_prefetch(c, layer_fwd_src(A, 0), layer_fwd_bytes(A.L[0]));
        }
        if (!one_tile) {
          for (int j = 0; j < NA; ++j) gather_rows<FRL_R>(rep(a, j).storage, rep(a, j).row_floats, a.indices + (size_t)u * a.B + row0, nvalid, raw0 + P.raw_off[j]);
          if (bon) norm_rows(a, P, NORM, raw0, nvalid);
        }
        for (int j = 0; j < NA; ++j) {
          const frl_replay_t& rj = rep(a, j);
          const float* rw = raw0 + P.raw_off[j];
          put_cols<FRL_R>(XN, sa, P.obs_off[j], rw, rj.row_floats, 0, rj.obs_dim, rj.obs_dim);
          put_cols<FRL_R>(XN, sa, P.act_off[j], rw, rj.row_floats, rb_col_act(rj), rj.act_dim, j == NA - 1 ? sa - P.act_off[j] : rj.act_dim);
        }
        const int aoff_i = P.act_off[ai], aipi = A.L[0].in_pad;
        put_cols<FRL_R>(XA, aipi, 0, raw, rf, 0, od, aipi);
        FRL_SYNC();
        mlp_fwd<FRL_R>(c, A, 0, 3, XA, aipi, A1, A2, ldh, MU, ap, FRL_ACT_NONE, fwd_hint(C, l0), as);
        FRL_PAR(t) {
          if (t < FRL_R * ap) {
            const int r = t / ap, j = t % ap;
            float act = 0.f, e = 0.f, lp = 0.f;
            if (j < ad) {
              const float mean = MU[r * ap + j];
              if (sac) {
                const float ls = fminf(fmaxf(A.p[A.x_off + j], -20.f), 2.f);
                const float sd = expf(ls);
                e = (r < nvalid) ? noise_at(a.noise_new, a, u, row0 + r, j, 2u) : 0.f;
                const float uu = fadd(mean, fmul(e, sd));
                const float diff = uu - mean;
                lp = -(diff * diff) / (2.f * (sd * sd)) - logf(sd) - FRL_LOG_SQRT_2PI;
                lp -= 2.f * (FRL_LOG2F - uu - softplus_t(-2.f * uu));
                act = tanhf(uu);
              } else {
                act = tanhf(mean);
              }
              XN[r * sa + aoff_i + j] = act;
            }
            AC[r * ap + j] = act; UU[r * ap + j] = lp; EPS[r * ap + j] = e;
          }
        }
        FRL_SYNC();
        // Q_role(s, pi(s)); dL/dQ = -(1/heads_used)/B
        const float dq = -invB / (float)heads_used;
        mlp_fwd<FRL_R>(c, C, l0, 3, XN, sa, H1, H2, ldh, QA, 4, FRL_ACT_NONE, bwd_hint(C, l0 + 2), cs);
        FRL_PAR(t) {
          float v = 0.f;
          if (t < FRL_R) {
            for (int j = 0; j < 4; ++j) dQA[t * 4 + j] = 0.f;
            if (t < nvalid) { dQA[t * 4] = dq; v = QA[t * 4]; }
          }
          red0[t] = v;
        }
        FRL_SYNC();
        const float qsum_tile = block_sum(red0);
        mlp_bwd<FRL_R>(c, C, l0, 3, XN, sa, H1, H2, ldh, dQA, 4, D1, D2, dXP, sa, nullptr, false, bwd_hint(A, 2), cs);
        // the critic slot is free again: start on the target head the next learn's stage 0 opens with (Polyak'd in stage 3)
        if (res && tile + nslots >= ntile && u + 1 < a.n_updates) res_fetch(c, cs, a.critic_target, l0, 3);
        // actor head backward (the log-prob / entropy terms are added once, by role 0)
        FRL_PAR(t) {
          float lsum = 0.f, esum = 0.f;
          if (t < FRL_R) {
            const int r = t;
            float lp = 0.f;
            for (int j = 0; j < ap; ++j) {
              float g = 0.f;
              if (j < ad && r < nvalid) {
                const float act = AC[r * ap + j];
                g = dXP[r * sa + aoff_i + j] * (1.f - act * act);
                if (sac && role == 0) { g += alpha * invB * 2.f * act; lp += UU[r * ap + j]; }
              }
              dMU[r * ap + j] = g;
            }
            if (r < nvalid && role == 0) { esum = -lp; lsum = alpha * lp; }     // actor_loss = mean(-Q_pi - alpha*entropy)
          }
          red0[t] = lsum; red1[t] = esum;
        }
        FRL_SYNC();
        loss_acc += block_sum(red0) - qsum_tile / (float)heads_used;
        ent_acc += block_sum(red1);
        if (sac) {
          // d/dlog_std_j = sum_r [ dL/du * std*eps ] (+ role 0: -alpha/B per row); zero outside the clamp range
          FRL_PAR(t) {
            if (t < ap) {       // also clears the 16-B padding of the log_std slot in the shared partial buffer
              const float lsr = (t < ad) ? A.p[A.x_off + t] : 1e30f;
              float g = 0.f;
              if (lsr >= -20.f && lsr <= 2.f) {
                const float sd = expf(lsr);
                for (int r = 0; r < nvalid; ++r) g += dMU[r * ap + t] * sd * EPS[r * ap + t] - (role == 0 ? alpha * invB : 0.f);
              }
              gp[A.x_off + t] = first ? g : gp[A.x_off + t] + g;
            }
          }
          FRL_SYNC();
        }
        mlp_bwd<FRL_R>(c, A, 0, 3, XA, aipi, A1, A2, ldh, dMU, ap, D1, D2, nullptr, 0, gp, !first, no_hint(), as);
        first = false;
      }
      FRL_PAR(t) { if (t == 0) { a.stats[c.cta * 8 + 1] = loss_acc; a.stats[c.cta * 8 + 2] = ent_acc; } }
      FRL_SYNC();
    } else if (s == 5) {
      if (!policy_step) return;
      // actor gradient: every ACTIVE role contributes to every actor parameter (TD3's idle head-2 CTAs are skipped)
      reduce_grads_roles(c.cta, c.ncta, c.red, A, a.gpart, gstride, nslots, nrole, 0, heads_used == nrole ? 1 : nrole, a.sumsq);
    } else {
      if (!policy_step) return;
      const AdamSpec hp = {a.lr_actor, a.beta1, a.beta2, a.eps, 0.0, (double)a.max_norm, (long)(a.step_actor0 + n_policy_before + 1)};
      adam_update(c.cta, c.ncta, c.red, A, a.sumsq, c.ncta, hp, a.defer_polyak ? nullptr : &a.actor_target, a.tau);
      res_invalidate(c, A);
      if (!a.defer_polyak) res_invalidate(c, a.actor_target);
      float o3[3] = {0.f, 0.f, 0.f};
      if (c.cta == 0) cta_sums(c.red, a.stats + 1, 8, a.stats + 2, 8, a.sumsq, 1, c.ncta, o3);
      FRL_PAR(t) {
        if (c.cta == 0 && t == 0) {
          const float l = o3[0], en = o3[1], ss = o3[2];
          a.out[u * 8 + 1] = l * invB;
          a.out[u * 8 + 5] = sqrtf(ss);
          a.out[u * 8 + 6] = en * invB;
          if (sac && a.adaptive_alpha) {
            // alpha_loss = (exp(log_alpha) * (entropy - target_entropy).detach()).mean();  Adam(lr alpha_lr) on log_alpha
            const float mean_term = en * invB - a.target_entropy;
            const float al = expf(a.alpha_state[0]);
            const float g = al * mean_term;
            a.out[u * 8 + 3] = g;      // == alpha_loss value (alpha * mean(entropy - target))
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
  }
};
typedef AcAlgoT<0> AcAlgo;            // generic instantiation (also provides the static sizing helpers other kernels borrow)
