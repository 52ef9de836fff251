// Fused discrete-action SAC learn() — the `hands_on` variant of the reference:
//   SAC_file/SAC_add_discrete.py:137-177 (Actor_discrete_hands_on: softmax policy; Critic_discrete_hands_on: twin heads obs -> Q(s, .)),
//   :299-348 (learn), :350-360 (update_target), Alpha :206-223 (target entropy 0.6 * log(n_actions)).
//     p' = actor(s')  [the ONLINE actor, like upstream]        y = r + gamma (1 - d) ( sum_a p'_a min_h Qt_h(s')_a + alpha H(p') )
//     critic loss = mse(Q1(s)[a], y) + mse(Q2(s)[a], y)         clip_grad_norm_(0.5), Adam
//     actor loss  = mean( - sum_a p_a min_h Q_h(s)_a - alpha H(p) )  with the UPDATED critic, H(p) = - sum p log(p + 1e-8)
//     Polyak of critic_target and actor_target, alpha loss = alpha * mean(H(p) - target_entropy)
// One persistent cooperative launch runs n_updates learns; six stages per learn (fwd/bwd critic · reduce · Adam + Polyak ·
// fwd/bwd actor · reduce · Adam + Polyak + alpha), built from the engine's tile primitives (8 batch rows per CTA tile).
#pragma once
#include "algo_dqn.cuh"

struct SacdAlgo {
  typedef frl_sacd_args_t Args;
  static const int NSTAGES = 6;
  FRL_SHD bool writes_params(int) { return true; }
  FRL_SHD bool stage_enabled(int, int, const Args&) { return true; }
  FRL_SHD int wbuf_floats(const Args& a) {
    int mx = 0;
    for (int i = 0; i < a.critic.n_layers; ++i) { const int f = wt_floats(a.critic.L[i]); if (f > mx) mx = f; }
    for (int i = 0; i < a.actor.n_layers; ++i) { const int f = wt_floats(a.actor.L[i]); if (f > mx) mx = f; }
    return (mx + 31) & ~31;
  }
  FRL_SHD int user_floats(const Args& a) {
    const int ldh = act_ld(a.critic.L[0].out_pad), ip = a.critic.L[0].in_pad, op = a.critic.L[2].out_pad;
    return FRL_R * (a.replay.row_floats + 2 * ip + 4 * ldh + 8 * op + 8) + 2 * FRL_NT + 64;
  }
  FRL_SHD int grid(const Args& a, int max_ctas) {
    const int tiles = (a.B + FRL_R - 1) / FRL_R;
    return tiles < max_ctas ? tiles : max_ctas;
  }
  FRL_SHD int n_updates(const Args& a) { return a.n_updates; }

  FRL_SDEV void stage(int s, int u, Cta& c, float* user, const Args& a) {
    const frl_net_t &A = a.actor, &C = a.critic;
    const int ldh = act_ld(C.L[0].out_pad), ip = C.L[0].in_pad, op = C.L[2].out_pad, nact = C.L[2].out;
    const int RF = a.replay.row_floats;
    const int ntile = (a.B + FRL_R - 1) / FRL_R;
    const int ncontrib = ntile < c.ncta ? ntile : c.ncta;
    const float invB = 1.0f / (float)a.B;
    const float alpha = expf(a.alpha_state[0]);
    SmemBump sb; sb.p = user;
    float* raw = sb.take(FRL_R * RF);
    float* Xo = sb.take(FRL_R * ip);
    float* Xn = sb.take(FRL_R * ip);
    float* H1 = sb.take(FRL_R * ldh);
    float* H2 = sb.take(FRL_R * ldh);
    float* D1 = sb.take(FRL_R * ldh);
    float* D2 = sb.take(FRL_R * ldh);
    float* P = sb.take(FRL_R * op);           // policy logits -> probabilities
    float* Q1 = sb.take(FRL_R * op);
    float* Q2 = sb.take(FRL_R * op);
    float* T1 = sb.take(FRL_R * op);          // target heads / scratch
    float* T2 = sb.take(FRL_R * op);
    float* dA = sb.take(FRL_R * op);
    float* dB = sb.take(FRL_R * op);
    float* Y = sb.take(FRL_R * 4);
    float* red0 = sb.take(FRL_NT);
    float* red1 = sb.take(FRL_NT);
    const size_t gstride = (size_t)(A.n_p > C.n_p ? A.n_p : C.n_p);
    float* gp = a.gpart + (size_t)c.cta * gstride;

    if (s == 0) {
      // ---- targets, twin critic forward / backward ----
      if (c.cta >= ntile) return;
      float loss_acc = 0.f;
      bool first = true;
      for (int tile = c.cta; tile < ntile; tile += c.ncta) {
        const int row0 = tile * FRL_R, nvalid = (a.B - row0) < FRL_R ? (a.B - row0) : FRL_R;
        stage_prefetch(c, layer_fwd_src(A, 0), layer_fwd_bytes(A.L[0]));
        gather_rows<FRL_R>(a.replay.storage, RF, a.indices + (size_t)u * a.B + row0, nvalid, raw);
        copy_cols<FRL_R>(Xo, ip, 0, raw, RF, 0, a.replay.obs_dim, ip);
        copy_cols<FRL_R>(Xn, ip, 0, raw, RF, rb_col_nobs(a.replay), a.replay.obs_dim, ip);
        mlp_fwd<FRL_R>(c, A, 0, 3, Xn, ip, H1, H2, ldh, P, op, FRL_ACT_NONE, fwd_hint(a.critic_target, 0));
        mlp_fwd<FRL_R>(c, a.critic_target, 0, 3, Xn, ip, H1, H2, ldh, T1, op, FRL_ACT_NONE, fwd_hint(a.critic_target, 3));
        mlp_fwd<FRL_R>(c, a.critic_target, 3, 3, Xn, ip, H1, H2, ldh, T2, op, FRL_ACT_NONE, fwd_hint(C, 0));
        FRL_PAR(t) {
          if (t < FRL_R) {
            const int r = t;
            float y = 0.f;
            if (r < nvalid) {
              float mx = P[r * op];
              for (int j = 1; j < nact; ++j) mx = fmaxf(mx, P[r * op + j]);
              float se = 0.f;
              for (int j = 0; j < nact; ++j) se += expf(P[r * op + j] - mx);
              float nq = 0.f, ent = 0.f;
              for (int j = 0; j < nact; ++j) {
                const float pj = expf(P[r * op + j] - mx) / se;
                nq += pj * fminf(T1[r * op + j], T2[r * op + j]);
                ent -= pj * logf(pj + 1e-8f);
              }
              const float rew = raw[r * RF + rb_col_rew(a.replay)], dn = raw[r * RF + rb_col_done(a.replay)];
              y = rew + a.gamma * (1.f - dn) * (nq + alpha * ent);
            }
            Y[r] = y;
          }
        }
        FRL_SYNC();
        // head 1 forward -> loss -> backward, then head 2 (activations of one head at a time)
        for (int h = 0; h < 2; ++h) {
          float* Q = h ? Q2 : Q1;
          mlp_fwd<FRL_R>(c, C, 3 * h, 3, Xo, ip, H1, H2, ldh, Q, op, FRL_ACT_NONE, bwd_hint(C, 3 * h + 2));
          FRL_PAR(t) {
            float l = 0.f;
            if (t < FRL_R) {
              const int r = t;
              for (int j = 0; j < op; ++j) dA[r * op + j] = 0.f;
              if (r < nvalid) {
                const int act = (int)raw[r * RF + rb_col_act(a.replay)];
                const float diff = Q[r * op + act] - Y[r];
                dA[r * op + act] = 2.f * diff * invB;
                l = diff * diff;
              }
            }
            red0[t] = l;
          }
          FRL_SYNC();
          loss_acc += block_sum(red0);
          mlp_bwd<FRL_R>(c, C, 3 * h, 3, Xo, ip, H1, H2, ldh, dA, op, D1, D2, nullptr, 0, gp, !first, h == 0 ? fwd_hint(C, 3) : no_hint());
        }
        first = false;
      }
      FRL_PAR(t) { if (t == 0) a.stats[c.cta * 8 + 0] = loss_acc; }
      FRL_SYNC();
    } else if (s == 1) {
      reduce_grads(c.cta, c.ncta, c.red, C, a.gpart, (int)gstride, ncontrib, a.sumsq);
    } else if (s == 2) {
      const AdamSpec hp = {a.lr_critic, a.beta1, a.beta2, a.eps, 0.0, (double)a.max_norm, (long)(a.step_critic0 + u + 1)};
      adam_update(c.cta, c.ncta, c.red, C, a.sumsq, c.ncta, hp, &a.critic_target, a.tau);
      if (c.cta == 0) {
        float o[3];
        cta_sums(c.red, a.stats, 8, a.sumsq, 1, nullptr, 0, ncontrib, o);
        float nrm[3];
        cta_sums(c.red, a.sumsq, 1, nullptr, 0, nullptr, 0, c.ncta, nrm);
        FRL_PAR(t) { if (t == 0) { a.out[u * 8 + 0] = o[0] * invB; a.out[u * 8 + 4] = sqrtf(nrm[0]); a.out[u * 8 + 2] = alpha; } }
        FRL_SYNC();
      }
    } else if (s == 3) {
      // ---- actor forward / backward against the UPDATED critic ----
      if (c.cta >= ntile) return;
      float loss_acc = 0.f, ent_acc = 0.f;
      bool first = true;
      for (int tile = c.cta; tile < ntile; tile += c.ncta) {
        const int row0 = tile * FRL_R, nvalid = (a.B - row0) < FRL_R ? (a.B - row0) : FRL_R;
        stage_prefetch(c, layer_fwd_src(C, 0), layer_fwd_bytes(C.L[0]));
        gather_rows<FRL_R>(a.replay.storage, RF, a.indices + (size_t)u * a.B + row0, nvalid, raw);
        copy_cols<FRL_R>(Xo, ip, 0, raw, RF, 0, a.replay.obs_dim, ip);
        mlp_fwd<FRL_R>(c, C, 0, 3, Xo, ip, D1, D2, ldh, Q1, op, FRL_ACT_NONE, fwd_hint(C, 3));
        mlp_fwd<FRL_R>(c, C, 3, 3, Xo, ip, D1, D2, ldh, Q2, op, FRL_ACT_NONE, fwd_hint(A, 0));
        mlp_fwd<FRL_R>(c, A, 0, 3, Xo, ip, H1, H2, ldh, P, op, FRL_ACT_NONE, bwd_hint(A, 2));
        FRL_PAR(t) {
          float l = 0.f, en = 0.f;
          if (t < FRL_R) {
            const int r = t;
            for (int j = 0; j < op; ++j) dA[r * op + j] = 0.f;
            if (r < nvalid) {
              float mx = P[r * op];
              for (int j = 1; j < nact; ++j) mx = fmaxf(mx, P[r * op + j]);
              float se = 0.f;
              for (int j = 0; j < nact; ++j) se += expf(P[r * op + j] - mx);
              // g_j = d(loss_row)/dp_j = -min Q_j + alpha (log(p_j + eps) + p_j / (p_j + eps));  dz_j = p_j (g_j - sum_k p_k g_k) / B
              float qpi = 0.f, ent = 0.f, pg = 0.f;
              for (int j = 0; j < nact; ++j) {
                const float pj = expf(P[r * op + j] - mx) / se, lg = logf(pj + 1e-8f), mq = fminf(Q1[r * op + j], Q2[r * op + j]);
                const float g = -mq + alpha * (lg + pj / (pj + 1e-8f));
                qpi += pj * mq;
                ent -= pj * lg;
                pg += pj * g;
                T1[r * op + j] = pj;
                T2[r * op + j] = g;
              }
              for (int j = 0; j < nact; ++j) dA[r * op + j] = T1[r * op + j] * (T2[r * op + j] - pg) * invB;
              l = -qpi - alpha * ent;
              en = ent;
            }
          }
          red0[t] = l; red1[t] = en;
        }
        FRL_SYNC();
        loss_acc += block_sum(red0);
        ent_acc += block_sum(red1);
        mlp_bwd<FRL_R>(c, A, 0, 3, Xo, ip, H1, H2, ldh, dA, op, D1, D2, nullptr, 0, gp, !first, no_hint());
        first = false;
      }
      FRL_PAR(t) { if (t == 0) { a.stats[c.cta * 8 + 1] = loss_acc; a.stats[c.cta * 8 + 2] = ent_acc; } }
      FRL_SYNC();
    } else if (s == 4) {
      reduce_grads(c.cta, c.ncta, c.red, A, a.gpart, (int)gstride, ncontrib, a.sumsq);
    } else {
      const AdamSpec hp = {a.lr_actor, a.beta1, a.beta2, a.eps, 0.0, (double)a.max_norm, (long)(a.step_actor0 + u + 1)};
      adam_update(c.cta, c.ncta, c.red, A, a.sumsq, c.ncta, hp, &a.actor_target, a.tau);
      if (c.cta == 0) {
        float o[3], nrm[3];
        cta_sums(c.red, a.stats + 1, 8, a.stats + 2, 8, nullptr, 0, ncontrib, o);
        cta_sums(c.red, a.sumsq, 1, nullptr, 0, nullptr, 0, c.ncta, nrm);
        FRL_PAR(t) {
          if (t == 0) {
            a.out[u * 8 + 1] = o[0] * invB;
            a.out[u * 8 + 5] = sqrtf(nrm[0]);
            a.out[u * 8 + 6] = o[1] * invB;
            if (a.adaptive_alpha) {
              // alpha_loss = (exp(log_alpha) * (entropy - target_entropy).detach()).mean();  Adam(lr alpha_lr) on log_alpha
              const float g = alpha * (o[1] * invB - a.target_entropy);
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
    }
  }
};
