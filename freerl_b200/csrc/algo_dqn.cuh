// Fused DQN learn():  Buffer.sample gather -> Q_target(s') max -> TD target -> Q(s)[a] -> MSE -> backward -> Adam -> Polyak
// Reference: DQN_file/DQN.py:104-128 (learn + update_target), Agent.update_Qnet :56-59 (no grad clipping).
// One persistent cooperative launch runs n_updates sequential learn() steps; 2 grid barriers per step.
#pragma once
#include "launch.cuh"

#ifndef FRL_R
#define FRL_R 8   // batch rows per CTA tile
#endif

// ---- replay row access ----------------------------------------------------------------------------------
FRL_DEV int rb_col_act(const frl_replay_t& rb) { return rb.obs_dim; }
FRL_DEV int rb_col_rew(const frl_replay_t& rb) { return rb.obs_dim + rb.act_dim; }
FRL_DEV int rb_col_done(const frl_replay_t& rb) { return rb.obs_dim + rb.act_dim + 1; }
FRL_DEV int rb_col_nobs(const frl_replay_t& rb) { return rb.obs_dim + rb.act_dim + 2; }

// Gather R sampled rows (vectorised 16-B loads, one row = row_floats/4 lanes) into smem raw[R][row_floats].
// Rows >= nvalid are zero-filled.
template <int R>
FRL_NI_MISC void gather_rows(const float* storage, int row_floats, const int64_t* idx, int nvalid, float* raw) {
  const int q = row_floats >> 2;
  FRL_PAR(t) {
    for (int e = t; e < R * q; e += FRL_NT) {
      const int r = e / q, j = e % q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nvalid) {
#ifndef FRL_EMUL
        v = __ldg(reinterpret_cast<const float4*>(storage + (size_t)idx[r] * row_floats) + j);
#else
        v = ld4(storage + (size_t)idx[r] * row_floats + 4 * j);
#endif
      }
      st4(raw + r * row_floats + 4 * j, v);
    }
  }
  FRL_SYNC();
}

// dst[r][dcol0 + j] = src[r][scol0 + j] for j < n, then zero up to ncols_total_pad (if > 0)
template <int R>
FRL_NI_MISC void copy_cols(float* dst, int ldd, int dcol0, const float* src, int lds, int scol0, int n, int zero_to) {
  FRL_PAR(t) {
    const int w = (zero_to > dcol0 + n ? zero_to : dcol0 + n) - dcol0;
    for (int e = t; e < R * w; e += FRL_NT) {
      const int r = e / w, j = e % w;
      dst[r * ldd + dcol0 + j] = (j < n) ? src[r * lds + scol0 + j] : 0.f;
    }
  }
  FRL_SYNC();
}

// put_cols: dst[r][dcol0 + j] = (j < n) ? src[r][scol0 + j] : 0  for j < w, r < R.  One thread per column (no index
// division) and NO barrier: several of these fill disjoint tiles from the gathered rows inside one phase.
template <int R>
FRL_DEV void put_cols(float* dst, int ldd, int dcol0, const float* src, int lds, int scol0, int n, int w) {
  FRL_PAR(t) {
    for (int j = t; j < w; j += FRL_NT) {
#pragma unroll
      for (int r = 0; r < R; ++r) dst[r * ldd + dcol0 + j] = (j < n) ? src[r * lds + scol0 + j] : 0.f;
    }
  }
}

struct DqnAlgo {
  typedef frl_dqn_args_t Args;
  static const int NSTAGES = 2;
  FRL_SHD bool writes_params(int) { return true; }
  FRL_SHD bool stage_enabled(int, int, const Args&) { return true; }
  FRL_SHD int wbuf_floats(const Args& a) {
    int mx = 0;
    for (int i = 0; i < a.q.n_layers; ++i) {
      int f = wt_floats(a.q.L[i]);
      if (f > mx) mx = f;
    }
    return (mx + 31) & ~31;
  }
  FRL_SHD int user_floats(const Args& a) {
    const int ldh = act_ld(a.q.L[0].out_pad), in_pad = a.q.L[0].in_pad, op = a.q.L[a.q.n_layers - 1].out_pad;
    return FRL_R * (a.replay.row_floats + 2 * in_pad + 3 * ldh + 4 * op + 8) + FRL_NT + 64;
  }
  FRL_SHD int grid(const Args& a, int max_ctas) {
    int tiles = (a.B + FRL_R - 1) / FRL_R;
    return tiles < max_ctas ? tiles : max_ctas;
  }
  FRL_SHD int n_updates(const Args& a) { return a.n_updates; }

  FRL_SDEV void stage(int s, int u, Cta& c, float* user, const Args& a) {
    const frl_net_t& q = a.q;
    const int nl = q.n_layers;
    const int ldh = act_ld(q.L[0].out_pad), in_pad = q.L[0].in_pad, op = q.L[nl - 1].out_pad;
    const int c0 = a.dueling ? 1 : 0;                          // first action column of the head ([V | A] when dueling)
    const int nact = q.L[nl - 1].out - c0;
    if (s == 0) {
      SmemBump sb; sb.p = user;
      float* raw = sb.take(FRL_R * a.replay.row_floats);
      float* Xo = sb.take(FRL_R * in_pad);
      float* Xn = sb.take(FRL_R * in_pad);
      float* H1 = sb.take(FRL_R * ldh);
      float* Ht = sb.take(FRL_R * ldh);
      float* D1 = sb.take(FRL_R * ldh);
      float* Q = sb.take(FRL_R * op);
      float* Qt = sb.take(FRL_R * op);
      float* dQ = sb.take(FRL_R * op);
      float* Qn = sb.take(FRL_R * op);               // Double: online net on next_obs
      float* lossr = sb.take(FRL_NT);
      float* gp = a.gpart + (size_t)c.cta * q.n_p;
      const int ntile = (a.B + FRL_R - 1) / FRL_R;
      const float invB = 1.0f / (float)a.B;
      float loss_acc = 0.f;   // block-uniform (only thread 0's copy is used on the GPU)
      bool first = true;
      // PER: the reference multiplies is_weight [B] with td_error^2 [B,1] -> mean over [B,B] = mean(w) * mean(td^2)
      float wmean = 1.f;
      if (a.is_weight) {
        FRL_PAR(t) {
          float sw = 0.f;
          for (int i = t; i < a.B; i += FRL_NT) sw += a.is_weight[i];
          lossr[t] = sw;
        }
        FRL_SYNC();
        wmean = block_sum(lossr) * invB;
      }
      for (int tile = c.cta; tile < ntile; tile += c.ncta) {
        const int row0 = tile * FRL_R;
        const int nvalid = (a.B - row0) < FRL_R ? (a.B - row0) : FRL_R;
        stage_prefetch(c, layer_fwd_src(a.q_target, 0), layer_fwd_bytes(a.q_target.L[0]));
        gather_rows<FRL_R>(a.replay.storage, a.replay.row_floats, a.indices + (size_t)u * a.B + row0, nvalid, raw);
        copy_cols<FRL_R>(Xo, in_pad, 0, raw, a.replay.row_floats, 0, a.replay.obs_dim, in_pad);
        copy_cols<FRL_R>(Xn, in_pad, 0, raw, a.replay.row_floats, rb_col_nobs(a.replay), a.replay.obs_dim, in_pad);
        // target net on next_obs, online net on obs
        mlp_fwd<FRL_R>(c, a.q_target, 0, nl, Xn, in_pad, Ht, Ht, ldh, Qt, op, FRL_ACT_NONE, fwd_hint(q, 0));
        if (a.double_q) mlp_fwd<FRL_R>(c, q, 0, nl, Xn, in_pad, Ht, Ht, ldh, Qn, op, FRL_ACT_NONE, fwd_hint(q, 0));
        mlp_fwd<FRL_R>(c, q, 0, nl, Xo, in_pad, H1, H1, ldh, Q, op, FRL_ACT_NONE, nl > 1 ? bwd_hint(q, nl - 1) : no_hint());
        // TD target, loss, dL/dQ  (DQN.py:110-116)
        FRL_PAR(t) {
          float l = 0.f;
          if (t < FRL_R) {
            const int r = t;
            for (int j = 0; j < op; ++j) dQ[r * op + j] = 0.f;
            if (r < nvalid) {
              // Dueling: Q_j = (V + A_j) - mean(A)  (DQN_with_tricks.py:79); plain head: Q_j = head_j
              float mt = 0.f, mn = 0.f, mo = 0.f;
              if (a.dueling) {
                float st = 0.f, sn = 0.f, so = 0.f;
                for (int j = 0; j < nact; ++j) {
                  st = fadd(st, Qt[r * op + 1 + j]);
                  so = fadd(so, Q[r * op + 1 + j]);
                  if (a.double_q) sn = fadd(sn, Qn[r * op + 1 + j]);
                }
                mt = fdiv(st, (float)nact); mo = fdiv(so, (float)nact); mn = fdiv(sn, (float)nact);
              }
              const float vt = a.dueling ? Qt[r * op] : 0.f, vo = a.dueling ? Q[r * op] : 0.f, vn = (a.dueling && a.double_q) ? Qn[r * op] : 0.f;
              float nq;
              if (a.double_q) {                       // first max of the online net's Q(s'), like torch.argmax
                int best = 0;
                float bv = a.dueling ? fadd(fadd(vn, Qn[r * op + 1]), -mn) : Qn[r * op];
                for (int j = 1; j < nact; ++j) {
                  const float qv = a.dueling ? fadd(fadd(vn, Qn[r * op + 1 + j]), -mn) : Qn[r * op + j];
                  if (qv > bv) { bv = qv; best = j; }
                }
                nq = a.dueling ? fadd(fadd(vt, Qt[r * op + 1 + best]), -mt) : Qt[r * op + best];
              } else {
                nq = a.dueling ? fadd(fadd(vt, Qt[r * op + 1]), -mt) : Qt[r * op];
                for (int j = 1; j < nact; ++j) nq = fmaxf(nq, a.dueling ? fadd(fadd(vt, Qt[r * op + 1 + j]), -mt) : Qt[r * op + j]);
              }
              const float rew = raw[r * a.replay.row_floats + rb_col_rew(a.replay)];
              const float dn = raw[r * a.replay.row_floats + rb_col_done(a.replay)];
              const float y = fadd(rew, fmul(fmul(a.gamma, nq), fadd(1.f, -dn)));
              const int act = (int)raw[r * a.replay.row_floats + rb_col_act(a.replay)];
              const float cur = a.dueling ? fadd(fadd(vo, Q[r * op + 1 + act]), -mo) : Q[r * op + act];
              const float diff = cur - y;
              const float gq = 2.f * diff * invB * wmean;
              if (a.dueling) {                        // dV = g, dA_j = g * (delta_ja - 1/n)
                dQ[r * op] = gq;
                const float gm = gq / (float)nact;
                for (int j = 0; j < nact; ++j) dQ[r * op + 1 + j] = (j == act ? gq : 0.f) - gm;
              } else {
                dQ[r * op + act] = gq;
              }
              if (a.td_error) a.td_error[(size_t)u * a.B + row0 + r] = diff;
              l = diff * diff;
            }
          }
          lossr[t] = l;
        }
        FRL_SYNC();
        loss_acc += block_sum(lossr);
        mlp_bwd<FRL_R>(c, q, 0, nl, Xo, in_pad, H1, H1, ldh, dQ, op, D1, D1, nullptr, 0, gp, !first, no_hint());
        first = false;
      }
      FRL_PAR(t) { if (t == 0) { a.stats[c.cta * 8 + 0] = loss_acc; a.stats[c.cta * 8 + 1] = wmean; } }
      FRL_SYNC();
    } else {
      // stage 1: cross-CTA reduce + Adam + Polyak on this CTA's parameter slice (no global norm needed)
      const int ncontrib = grid(a, c.ncta);
      reduce_grads(c.cta, c.ncta, c.red, q, a.gpart, q.n_p, ncontrib, nullptr);
      const AdamSpec hp = {a.lr, a.beta1, a.beta2, a.eps, 0.0, 0.0, (long)(a.step0 + u + 1)};
      adam_update(c.cta, c.ncta, c.red, q, nullptr, 0, hp, &a.q_target, a.tau);
      if (c.cta == 0) {                     // loss metric: partials fetched in parallel, folded in CTA order (block-uniform branch)
        float o[3];
        cta_sums(c.red, a.stats, 8, nullptr, 0, nullptr, 0, ncontrib, o);
        FRL_PAR(t) { if (t == 0) a.out[u * 8 + 0] = o[0] / (float)a.B * a.stats[1]; }
        FRL_SYNC();
      }
    }
  }
};
