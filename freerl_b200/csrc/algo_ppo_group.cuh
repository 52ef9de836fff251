// Group mode of the fused PPO update: MAPPO_discrete.py's networks inside learn().
//   Actor_discrete / Critic forward    MAPPO_file/MAPPO_discrete.py:95-153   F.layer_norm(x, x.size()[1:]) of [minibatch, T, N, features]
//   losses                             MAPPO_file/MAPPO_discrete.py:333-361  (Categorical surrogate + entropy; value loss 0 / 2 / 3)
// `x.size()[1:]` of a 4-D tensor makes every LayerNorm run jointly over the T*N rows of one EPISODE (group) and the feature axis, so
// the mean / variance of a hidden layer depend on ALL rows of the group before the next layer can start, and the backward needs two
// more group-wide means per LayerNorm.  A CTA owns whole groups and sweeps the group's 8-row tiles once per statistic, recomputing
// the forward pass each time (nothing larger than one tile is kept; the statistics are block-uniform scalars):
//   sweep 0  (critic, feature_norm)  sum x, sum x^2 of the inputs                         -> mu0, rstd0
//   sweep 1  H1 = relu(L0 x^)        sum, sum of squares                                  -> mu1, rstd1
//   sweep 2  H2 = relu(L1 Y1)        ...                                                  -> mu2, rstd2
//   sweep 3  OUT, loss, dOUT, dY2 = dOUT W3        mean dY2, mean dY2 * Y2                -> LayerNorm-2 backward terms
//   sweep 4  dH2 = LN2'(dY2) relu', dW3, dW2, dY1 = dH2 W2     mean dY1, mean dY1 * Y1    -> LayerNorm-1 backward terms
//   sweep 5  dH1 = LN1'(dY1) relu', dW1
// (about 4x the FLOPs of the plain tile kernel; this is a sibling variant, not the bench path).  The reduce / clip / optimiser stages
// are the shared ones of PpoAlgoT.  value_loss 3 (ValueClip + huber_loss: the squared maximum of two batch-mean scalars) needs the
// whole minibatch's critic outputs before any gradient: a pre-pass launch (group_prepass) leaves per-CTA sums, the update launch folds them.
#pragma once

struct GrpStat { float mu, rs; };
FRL_DEV GrpStat grp_stat(double s, double q, double n) {
  GrpStat r;
  const double mu = s / n;
  double var = q / n - mu * mu;                       // biased, like F.layer_norm
  if (var < 0.0) var = 0.0;
  r.mu = (float)mu;
  r.rs = (float)(1.0 / sqrt(var + 1e-5));
  return r;
}
// s += sum A, q += sum A * B over nvalid rows x n columns (B = A: sum of squares)
FRL_NI_MISC void grp_sums(const float* A, const float* B, int ld, int nvalid, int n, float* red0, float* red1, double* s, double* q) {
  FRL_PAR(t) {
    float x = 0.f, y = 0.f;
    for (int e = t; e < nvalid * n; e += FRL_NT) {
      const int r = e / n, k = e - r * n;
      const float v = A[r * ld + k];
      x += v; y += v * B[r * ld + k];
    }
    red0[t] = x; red1[t] = y;
  }
  FRL_SYNC();
  *s += (double)block_sum(red0);
  *q += (double)block_sum(red1);
}
// Y = (X - mu) * rs on the valid rows / real columns, 0 elsewhere ([rows][ld] tile)
FRL_NI_MISC void grp_norm(const float* X, int ld, int rows, int nvalid, int n, GrpStat st, float* Y) {
  FRL_PAR(t) {
    for (int e = t; e < rows * ld; e += FRL_NT) {
      const int r = e / ld, k = e - r * ld;
      Y[e] = (r < nvalid && k < n) ? (X[e] - st.mu) * st.rs : 0.f;
    }
  }
  FRL_SYNC();
}
// dX = rs * (dY - m1 - Y * m2)  [ln]  or  dY, then masked by relu'(H); 0 on padded rows / columns
FRL_NI_MISC void grp_ln_bwd(const float* dY, const float* Y, const float* H, int ld, int rows, int nvalid, int n, bool ln, float rs, float m1,
                            float m2, float* dX) {
  FRL_PAR(t) {
    for (int e = t; e < rows * ld; e += FRL_NT) {
      const int r = e / ld, k = e - r * ld;
      float v = 0.f;
      if (r < nvalid && k < n && H[e] > 0.f) v = ln ? rs * (dY[e] - m1 - Y[e] * m2) : dY[e];
      dX[e] = v;
    }
  }
  FRL_SYNC();
}
FRL_DEV float grp_huber(float e, float d) { const float ae = fabsf(e); return ae <= d ? 0.5f * e * e : d * (ae - 0.5f * d); }
FRL_DEV float grp_huber_d(float e, float d) { return fabsf(e) <= d ? e : (e > 0.f ? d : -d); }

FRL_HD int ppo_group_user_floats(const frl_ppo_args_t& a) {
  const int ldh = act_ld(a.net.L[0].out_pad), ip = a.net.L[0].in_pad, cip = a.net.L[3].in_pad, ap = a.net.L[2].out_pad;
  const int ipm = ip > cip ? ip : cip;
  return 8 * (2 * ipm + 6 * ldh + 2 * ap + 8 + 8) + 2 * FRL_NT + 64;
}

FRL_DEV void ppo_group_stage(Cta& c, float* user, const frl_ppo_args_t& a, int u) {
  const int R = 8;
  const frl_net_t& N = a.net;
  const int G = a.group_rows, rows = a.mb_rows[u], ngroups = rows / G, nst = (G + R - 1) / R;
  if (c.cta >= ngroups) return;
  const int ncontrib = ngroups < c.ncta ? ngroups : c.ncta;
  const int ldh = act_ld(N.L[0].out_pad), ip = N.L[0].in_pad, cip = N.L[3].in_pad, ap = N.L[2].out_pad, nout = N.L[2].out;
  const int ipm = ip > cip ? ip : cip;
  SmemBump sb; sb.p = user;
  float* XR = sb.take(R * ipm);
  float* XN = sb.take(R * ipm);
  float* H1 = sb.take(R * ldh); float* Y1b = sb.take(R * ldh);
  float* H2 = sb.take(R * ldh); float* Y2b = sb.take(R * ldh);
  float* D1 = sb.take(R * ldh); float* D2 = sb.take(R * ldh);
  float* OUT = sb.take(R * ap);
  float* dOUT = sb.take(R * ap);
  float* ROW = sb.take(R * 8);          // per row: action, old log-prob, advantage, v_target, v_old
  float* red0 = sb.take(FRL_NT);
  float* red1 = sb.take(FRL_NT);
  float* gp = a.gpart + (size_t)c.cta * N.n_p;
  const bool hid = (a.group_norm & 1) != 0, cin = (a.group_norm & 2) != 0, prepass = a.group_prepass != 0;
  const float inv_rows = 1.0f / (float)rows;
  const int c_in = a.critic_obs ? a.critic_obs_dim : a.obs_dim;
  float la = 0.f, lc = 0.f, le = 0.f;
  double hc = 0.0, ho = 0.0;            // pre-pass: sums of huber(e_clip), huber(e_orig) over this CTA's rows
  // value_loss 3: the two batch means, folded from the pre-pass launch's per-CTA sums in CTA order (identical on every CTA)
  float vco = 0.f, vcc = 0.f, vloss = 0.f;
  if (a.value_loss == 3 && !prepass) {
    FRL_PAR(t) {
      if (t == 0) {
        double A = 0.0, B = 0.0;
        for (int i = 0; i < ncontrib; ++i) { A += (double)a.stats[i * 8 + 5]; B += (double)a.stats[i * 8 + 6]; }
        const float am = (float)(A / (double)rows), bm = (float)(B / (double)rows);
        // torch.max(a^2, b^2): the larger branch takes the gradient (an exact tie is shared; both halves then differ by rounding only)
        red0[0] = (bm * bm >= am * am) ? 2.f * bm * inv_rows : 0.f;
        red0[1] = (bm * bm >= am * am) ? 0.f : 2.f * am * inv_rows;
        red0[2] = fmaxf(am * am, bm * bm);
      }
    }
    FRL_SYNC();
    vco = red0[0]; vcc = red0[1]; vloss = red0[2];
    FRL_SYNC();
  }
  bool first_group = true;
  for (int g = c.cta; g < ngroups; g += c.ncta) {
    const int64_t* gidx = a.indices + (size_t)u * a.mb + (size_t)g * G;
    for (int net = prepass ? 1 : 0; net < 2; ++net) {
      const int l0 = 3 * net, ipn = net ? cip : ip, in_dim = net ? c_in : a.obs_dim, ldo = net ? 4 : ap;
      const frl_layer_t &L0 = N.L[l0], &L1 = N.L[l0 + 1], &L2 = N.L[l0 + 2];
      const bool in_norm = net == 1 && cin;
      const float* src = (net && a.critic_obs) ? a.critic_obs : a.obs;
      double s0 = 0.0, q0 = 0.0, s1 = 0.0, q1 = 0.0, s2 = 0.0, q2 = 0.0, s3 = 0.0, q3 = 0.0, s4 = 0.0, q4 = 0.0;
      GrpStat st0, st1, st2;
      st0.mu = st1.mu = st2.mu = 0.f; st0.rs = st1.rs = st2.rs = 1.f;
      float m3a = 0.f, m3b = 0.f, m4a = 0.f, m4b = 0.f;
      for (int pass = in_norm ? 0 : (hid ? 1 : 3); pass <= (prepass ? 3 : 5); ++pass) {
        if (!hid && (pass == 1 || pass == 2)) continue;
        for (int st = 0; st < nst; ++st) {
          const int row0 = st * R, nvalid = (G - row0) < R ? (G - row0) : R;
          const int64_t* idx = gidx + row0;
          FRL_PAR(t) {
            for (int e = t; e < R * ipn; e += FRL_NT) {
              const int r = e / ipn, j = e - r * ipn;
              XR[e] = (r < nvalid && j < in_dim) ? src[(size_t)idx[r] * in_dim + j] : 0.f;
            }
            if (t < R) {
              const bool v = t < nvalid;
              const size_t i = v ? (size_t)idx[t] : 0;
              ROW[t * 8 + 0] = v ? a.action[i * a.act_cols] : 0.f;
              ROW[t * 8 + 1] = v ? a.logp_old[i * a.logp_cols] : 0.f;
              ROW[t * 8 + 2] = v ? a.adv[i * a.n_adv] : 0.f;
              ROW[t * 8 + 3] = v ? a.v_target[i * a.n_adv] : 0.f;
              ROW[t * 8 + 4] = (v && a.v_old) ? a.v_old[i * a.n_adv] : 0.f;
            }
          }
          FRL_SYNC();
          if (pass == 0) { grp_sums(XR, XR, ipn, nvalid, in_dim, red0, red1, &s0, &q0); continue; }
          const float* X0 = XR;
          if (in_norm) { grp_norm(XR, ipn, R, nvalid, in_dim, st0, XN); X0 = XN; }
          layer_fwd<R>(c, N, l0, X0, ipn, H1, ldh, FRL_ACT_RELU, fwd_hint(N, pass == 1 ? l0 : l0 + 1));
          if (pass == 1) { grp_sums(H1, H1, ldh, nvalid, L0.out, red0, red1, &s1, &q1); continue; }
          const float* Y1 = H1;
          if (hid) { grp_norm(H1, ldh, R, nvalid, L0.out, st1, Y1b); Y1 = Y1b; }
          layer_fwd<R>(c, N, l0 + 1, Y1, ldh, H2, ldh, FRL_ACT_RELU, fwd_hint(N, pass == 2 ? l0 : l0 + 2));
          if (pass == 2) { grp_sums(H2, H2, ldh, nvalid, L1.out, red0, red1, &s2, &q2); continue; }
          const float* Y2 = H2;
          if (hid) { grp_norm(H2, ldh, R, nvalid, L1.out, st2, Y2b); Y2 = Y2b; }
          layer_fwd<R>(c, N, l0 + 2, Y2, ldh, OUT, ldo, FRL_ACT_NONE, fwd_hint(N, prepass ? l0 : l0 + 2));
          // ---- heads: loss terms (counted in sweep 3 only) and the gradient w.r.t. the network output ----
          FRL_PAR(t) {
            float x0 = 0.f, x1 = 0.f;
            if (t < R) {
              const int r = t;
              for (int j = 0; j < ldo; ++j) dOUT[r * ldo + j] = 0.f;
              if (r < nvalid && net == 0) {
                float mx = OUT[r * ap];
                for (int j = 1; j < nout; ++j) mx = fmaxf(mx, OUT[r * ap + j]);
                float se_ = 0.f;
                for (int j = 0; j < nout; ++j) se_ += expf(OUT[r * ap + j] - mx);
                const float lse = mx + logf(se_);
                const int act = (int)ROW[r * 8 + 0];
                const float lp_now = OUT[r * ap + act] - lse, lp_old = ROW[r * 8 + 1], A = ROW[r * 8 + 2];
                float ent = 0.f;
                for (int j = 0; j < nout; ++j) { const float lg = OUT[r * ap + j] - lse; ent -= expf(lg) * lg; }
                const float ratio = expf(lp_now - lp_old);
                const float lo = 1.f - a.clip_param, hi = 1.f + a.clip_param;
                const float rc = fminf(fmaxf(ratio, lo), hi);
                const float s1_ = ratio * A, s2_ = rc * A;
                const bool inside = ratio >= lo && ratio <= hi;
                const float dlp = (inside || s1_ < s2_) ? -A * ratio * inv_rows : 0.f;
                x0 = -fminf(s1_, s2_) * inv_rows;
                x1 = ent;
                for (int j = 0; j < nout; ++j) {
                  const float lg = OUT[r * ap + j] - lse, pj = expf(lg);
                  dOUT[r * ap + j] = dlp * ((j == act ? 1.f : 0.f) - pj) + a.entropy_coef * inv_rows * pj * (lg + ent);
                }
              } else if (r < nvalid) {
                const float V = OUT[r * 4], vt = ROW[r * 8 + 3], vo = ROW[r * 8 + 4];
                const float eo = V - vt, dv = V - vo;
                const float ec = fminf(fmaxf(dv, -a.clip_param), a.clip_param) + vo - vt;
                const bool inside = dv >= -a.clip_param && dv <= a.clip_param;
                if (a.value_loss == 3) {
                  x0 = grp_huber(ec, a.huber_delta);          // pre-pass sums
                  x1 = grp_huber(eo, a.huber_delta);
                  dOUT[r * 4] = vco * grp_huber_d(eo, a.huber_delta) + (inside ? vcc * grp_huber_d(ec, a.huber_delta) : 0.f);
                } else if (a.value_loss == 2) {
                  const float qo = eo * eo, qc = ec * ec;
                  if (qo >= qc) { x0 = qo; dOUT[r * 4] = 2.f * eo * inv_rows; }
                  else { x0 = qc; dOUT[r * 4] = inside ? 2.f * ec * inv_rows : 0.f; }
                } else {
                  x0 = eo * eo;
                  dOUT[r * 4] = 2.f * eo * inv_rows;
                }
              }
            }
            red0[t] = x0; red1[t] = x1;
          }
          FRL_SYNC();
          if (pass == 3) {
            const float t0 = block_sum(red0), t1 = block_sum(red1);
            if (net == 0) { la += t0; le += t1; }
            else if (a.value_loss == 3) { hc += (double)t0; ho += (double)t1; }
            else lc += t0;
          }
          if (prepass) continue;
          layer_bwd_dx<R>(c, N, l0 + 2, dOUT, ldo, nullptr, 0, D2, ldh, bwd_hint(N, pass == 3 ? l0 : l0 + 1));       // dY2
          if (pass == 3) { if (hid) grp_sums(D2, Y2, ldh, nvalid, L1.out, red0, red1, &s3, &q3); continue; }
          const bool acc = !(first_group && st == 0);
          grp_ln_bwd(D2, Y2, H2, ldh, R, nvalid, L1.out, hid, st2.rs, m3a, m3b, D1);                                      // dH2 (pre-ReLU)
          if (pass == 4) {
            gemm_outer<R>(dOUT, ldo, L2.out_pad, Y2, ldh, L2.in_pad, L2.in, gp + L2.w_off, gp + L2.b_off, acc);
            gemm_outer<R>(D1, ldh, L1.out_pad, Y1, ldh, L1.in_pad, L1.in, gp + L1.w_off, gp + L1.b_off, acc);
          }
          layer_bwd_dx<R>(c, N, l0 + 1, D1, ldh, nullptr, 0, D2, ldh, bwd_hint(N, l0));                                   // dY1
          if (pass == 4) { if (hid) grp_sums(D2, Y1, ldh, nvalid, L0.out, red0, red1, &s4, &q4); continue; }
          grp_ln_bwd(D2, Y1, H1, ldh, R, nvalid, L0.out, hid, st1.rs, m4a, m4b, D1);                                      // dH1 (pre-ReLU)
          gemm_outer<R>(D1, ldh, L0.out_pad, X0, ipn, L0.in_pad, L0.in, gp + L0.w_off, gp + L0.b_off, acc);
        }
        if (pass == 0) st0 = grp_stat(s0, q0, (double)G * in_dim);
        else if (pass == 1) st1 = grp_stat(s1, q1, (double)G * L0.out);
        else if (pass == 2) st2 = grp_stat(s2, q2, (double)G * L1.out);
        else if (pass == 3) { m3a = (float)(s3 / ((double)G * L1.out)); m3b = (float)(q3 / ((double)G * L1.out)); }
        else if (pass == 4) { m4a = (float)(s4 / ((double)G * L0.out)); m4b = (float)(q4 / ((double)G * L0.out)); }
      }
    }
    first_group = false;
  }
  if (a.value_loss == 3 && !prepass) lc = c.cta == 0 ? vloss * (float)rows : 0.f;       // out[u][1] = sum of the CTAs' lc / rows
  FRL_PAR(t) {
    if (t == 0) {
      if (prepass) { a.stats[c.cta * 8 + 5] = (float)hc; a.stats[c.cta * 8 + 6] = (float)ho; }
      else { a.stats[c.cta * 8 + 0] = la; a.stats[c.cta * 8 + 1] = lc; a.stats[c.cta * 8 + 2] = le; }
    }
  }
  FRL_SYNC();
}
