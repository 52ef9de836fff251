// Fused Rainbow learn(): Categorical (C51) + Dueling + NoisyLinear heads, Double-DQN action selection, PER weights,
// n-step gamma.  Reference: DQN_file/DQN_with_tricks.py:81-160 (Categorical._predict / forward / projection_dist),
// :242-284 (learn), DQN_file/Noisy_net.py:17-76 (factorised noise, resampled on EVERY forward).
//
// The trainable block holds the torch tensors (l1.weight/bias, V/A.{weight,bias}_{mu,sigma}); for each of the three
// forwards of one learn() (online on s', target on s', online on s) a noise-applied "effective" linear net
//   W = mu + sigma * (eps_out x eps_in),  b = bias_mu + bias_sigma * eps_out
// is materialised in HBM (stage 0) in the engine's layouts, so the forward/backward passes are the ordinary TMA-staged
// GEMMs.  The A head (n_actions*n_atoms outputs) is split into column blocks of <= 128 outputs.
#pragma once
#include "algo_ppo.cuh"

struct RainbowAlgo {
  typedef frl_rainbow_args_t Args;
  static const int NSTAGES = 3;
  FRL_SHD bool writes_params(int) { return true; }
  FRL_SHD bool stage_enabled(int, int, const Args&) { return true; }
  FRL_SHD int wbuf_floats(const Args& a) { return (AcAlgo::max_layer_floats(a.eff[2]) + 31) & ~31; }
  FRL_SHD int natot(const Args& a) { return (a.n_actions * a.n_atoms + 3) & ~3; }
  FRL_SHD int user_floats(const Args& a) {
    const int ldh = act_ld(a.eff[2].L[0].out_pad), ip = a.eff[2].L[0].in_pad, zp = (a.n_atoms + 3) & ~3;
    return FRL_R * (a.replay.row_floats + 2 * ip + 3 * ldh + 3 * zp + 2 * natot(a) + 7 * zp + 8) + 3 * FRL_NT + 64;
  }
  FRL_SHD int grid(const Args& a, int max_ctas) {
    int tiles = (a.B + FRL_R - 1) / FRL_R;
    return tiles < max_ctas ? tiles : max_ctas;
  }
  FRL_SHD int n_updates(const Args&) { return 1; }

  // effective parameter e (index into eff-net block) of forward f <- trainable block + noise
  FRL_SDEV void noisy_apply(const Args& a, int f, int cta, int ncta, const float* eps_all = nullptr) {
    const frl_net_t& E = a.eff[f];
    const float* P = (f == 1) ? a.p_target : a.p;
    const float* eps = (eps_all ? eps_all : a.eps) + (size_t)f * a.eps_len;
    FRL_PAR(t) {
      for (int e = cta * FRL_NT + t; e < E.n_p; e += ncta * FRL_NT) {
        float val = 0.f;
        int mi = -1;
        for (int li = 0; li < E.n_layers; ++li) {
          const frl_layer_t& L = E.L[li];
          const frl_noisy_map_t& M = a.map[li];
          const int wsz = L.out_pad * L.in_pad;
          if (e >= L.w_off && e < L.w_off + wsz) {
            const int j = (e - L.w_off) / L.in_pad, k = (e - L.w_off) % L.in_pad;
            if (j < L.out && k < L.in) {
              const int src = (M.row0 + j) * L.in_pad + k;       // torch tensors are [out][in_pad] row-major in the block
              val = P[M.mu_w + src];
              if (M.sg_w >= 0) val = fadd(val, fmul(P[M.sg_w + src], fmul(eps[M.eps_out + M.row0 + j], eps[M.eps_in + k])));
            }
            mi = L.wt_off + k * wt_ld(L) + j;
            break;
          }
          if (e >= L.b_off && e < L.b_off + L.out_pad) {
            const int j = e - L.b_off;
            if (j < L.out) {
              val = P[M.mu_b + M.row0 + j];
              if (M.sg_b >= 0) val = fadd(val, fmul(P[M.sg_b + M.row0 + j], eps[M.eps_out + M.row0 + j]));
            }
            mi = L.wt_off + wt_bias(L) + j;
            break;
          }
        }
        E.p[e] = val;
        if (mi >= 0) E.pt[mi] = val;
      }
    }
    FRL_SYNC();
  }

  // ---- per-row distribution helpers ------------------------------------------------------------------------------------
  // The distributional head works on R rows x 51 atoms (x n_actions).  Every step below is spread over (row, atom) items or
  // over `lanes` partial reductions per row — the first version walked each row with ONE thread (8 of 256 threads busy for
  // ~5 us per softmax) and those serial walks were a third of the learn.
  // Row i of a "row set" starts at X[(i / inner) * outer + (i % inner) * istride]  (dist of one action: inner 1, outer zp;
  // all actions of the dueling head: inner n_actions, outer nat, istride n_atoms).
  FRL_SDEV int row_base(int i, int inner, int outer, int istride) { return (i / inner) * outer + (i % inner) * istride; }
  FRL_SDEV int row_lanes(int nrows) { int l = FRL_NT / (nrows > 0 ? nrows : 1); return l > 32 ? 32 : (l < 1 ? 1 : l); }

  // in-place softmax over n entries of each row; sc: [FRL_NT + nrows] floats of scratch.  Rows beyond FRL_NT are not supported
  // (n_actions * R <= FRL_NT is checked by the host wrapper through the layer-count limit).
  FRL_SDEV void softmax_rows(float* X, int nrows, int n, int inner, int outer, int istride, float* sc) {
    const int lanes = row_lanes(nrows);
    FRL_PAR(t) {
      if (t < nrows * lanes) {
        const int i = t / lanes, l = t % lanes, b = row_base(i, inner, outer, istride);
        float m = -1e30f;
        for (int z = l; z < n; z += lanes) m = fmaxf(m, X[b + z]);
        sc[t] = m;
      }
    }
    FRL_SYNC();
    FRL_PAR(t) {
      if (t < nrows) { float m = sc[t * lanes]; for (int l = 1; l < lanes; ++l) m = fmaxf(m, sc[t * lanes + l]); sc[FRL_NT + t] = m; }
    }
    FRL_SYNC();
    FRL_PAR(t) {
      if (t < nrows * lanes) {
        const int i = t / lanes, l = t % lanes, b = row_base(i, inner, outer, istride);
        const float mx = sc[FRL_NT + i];
        float se = 0.f;
        for (int z = l; z < n; z += lanes) { const float e = expf(X[b + z] - mx); X[b + z] = e; se += e; }
        sc[t] = se;
      }
    }
    FRL_SYNC();
    FRL_PAR(t) {
      if (t < nrows) { float se = sc[t * lanes]; for (int l = 1; l < lanes; ++l) se += sc[t * lanes + l]; sc[FRL_NT + t] = se; }
    }
    FRL_SYNC();
    FRL_PAR(t) {
      if (t < nrows * lanes) {
        const int i = t / lanes, l = t % lanes, b = row_base(i, inner, outer, istride);
        const float se = sc[FRL_NT + i];
        for (int z = l; z < n; z += lanes) X[b + z] = X[b + z] / se;
      }
    }
    FRL_SYNC();
  }

  // logits -> per-action expected value q[r][ac] = sum_z softmax_z(V + A[ac] - mean_a A)[z] * z_atom
  //   mean: [R][zp] scratch, L: [R][nat] scratch (all-action logits / probabilities), sc: [2*FRL_NT] scratch
  FRL_SDEV void head_probs(const Args& a, const float* V, int zp, const float* A, int nat, float* qv /*[R][n_actions]*/, float* mean,
                           float* L, float* sc) {
    const int nA = a.n_actions, nZ = a.n_atoms;
    FRL_PAR(t) {
      for (int e = t; e < FRL_R * nZ; e += FRL_NT) {
        const int r = e / nZ, z = e % nZ;
        float m = 0.f;
        for (int b = 0; b < nA; ++b) m += A[r * nat + b * nZ + z];
        mean[r * zp + z] = m / (float)nA;
      }
    }
    FRL_SYNC();
    FRL_PAR(t) {
      for (int e = t; e < FRL_R * nA * nZ; e += FRL_NT) {
        const int r = e / (nA * nZ), j = e % (nA * nZ), z = j % nZ;
        L[r * nat + j] = V[r * zp + z] + A[r * nat + j] - mean[r * zp + z];
      }
    }
    FRL_SYNC();
    softmax_rows(L, FRL_R * nA, nZ, nA, nat, nZ, sc);
    const int nrows = FRL_R * nA, lanes = row_lanes(nrows);
    FRL_PAR(t) {
      if (t < nrows * lanes) {
        const int i = t / lanes, l = t % lanes, b = row_base(i, nA, nat, nZ);
        float q = 0.f;
        for (int z = l; z < nZ; z += lanes) q += L[b + z] * a.z[z];
        sc[t] = q;
      }
    }
    FRL_SYNC();
    FRL_PAR(t) {
      if (t < nrows) { float q = sc[t * lanes]; for (int l = 1; l < lanes; ++l) q += sc[t * lanes + l]; qv[t] = q; }     // t = r * nA + ac
    }
    FRL_SYNC();
  }

  // distribution of one chosen action per row -> P[r][z]   (sc: [2*FRL_NT] scratch)
  FRL_SDEV void dist_of(const Args& a, const float* V, int zp, const float* A, int nat, const int* act, float* P, float* sc) {
    const int nA = a.n_actions, nZ = a.n_atoms;
    FRL_PAR(t) {
      for (int e = t; e < FRL_R * nZ; e += FRL_NT) {
        const int r = e / nZ, z = e % nZ;
        float m = 0.f;
        for (int b = 0; b < nA; ++b) m += A[r * nat + b * nZ + z];
        m = m / (float)nA;
        P[r * zp + z] = V[r * zp + z] + A[r * nat + act[r] * nZ + z] - m;
      }
    }
    FRL_SYNC();
    softmax_rows(P, FRL_R, nZ, 1, zp, 0, sc);
  }

  // hidden -> V and A head outputs through effective net E (layers 1.. )
  FRL_SDEV void heads_fwd(Cta& c, const frl_net_t& E, const float* H, int ldh, float* V, int zp, float* A, int nat, Hint next) {
    layer_fwd<FRL_R>(c, E, 1, H, ldh, V, zp, FRL_ACT_NONE, fwd_hint(E, 2));
    int col = 0;
    for (int li = 2; li < E.n_layers; ++li) {
      layer_fwd<FRL_R>(c, E, li, H, ldh, A + col, nat, FRL_ACT_NONE, li + 1 < E.n_layers ? fwd_hint(E, li + 1) : next);
      col += E.L[li].out;
    }
  }

  FRL_SDEV void stage(int s, int u, Cta& c, float* user, const Args& a) {
    const frl_net_t& E3 = a.eff[2];
    const int ldh = act_ld(E3.L[0].out_pad), ip = E3.L[0].in_pad, zp = (a.n_atoms + 3) & ~3, nat = natot(a);
    const int nA = a.n_actions, nZ = a.n_atoms, rf = a.replay.row_floats;
    const int ntile = (a.B + FRL_R - 1) / FRL_R;
    const int ncontrib = ntile < c.ncta ? ntile : c.ncta;
    if (s == 0) {
      const float* eps_all = nullptr;
      if (a.noise_gen) {
        // fast mode: every CTA derives the same 3 x eps_len factorised-noise vector in shared memory (counter-based, so no
        // exchange and no extra barrier); CTA 0 publishes it for the host-side weight_epsilon bookkeeping
        float* se = user;
        FRL_PAR(t) {
          for (int e = t; e < 3 * a.eps_len; e += FRL_NT) {
            const int f = e / a.eps_len, i = e - f * a.eps_len;
            const float x = frl_randn(a.noise_seed, 40u + (uint32_t)f, (uint32_t)a.noise_counter, (uint32_t)i);
            const float fe = (x < 0.f ? -1.f : (x > 0.f ? 1.f : 0.f)) * sqrtf(fabsf(x));
            se[e] = fe;
            if (c.cta == 0) const_cast<float*>(a.eps)[e] = fe;
          }
        }
        FRL_SYNC();
        eps_all = se;
      }
      for (int f = 0; f < 3; ++f) noisy_apply(a, f, c.cta, c.ncta, eps_all);
      return;
    }
    if (s == 1) {
      SmemBump sb; sb.p = user;
      float* raw = sb.take(FRL_R * rf);
      float* Xo = sb.take(FRL_R * ip);
      float* Xn = sb.take(FRL_R * ip);
      float* H = sb.take(FRL_R * ldh);      // hidden of the gradient-carrying forward
      float* Hn = sb.take(FRL_R * ldh);     // scratch hidden (next_obs passes)
      float* dH = sb.take(FRL_R * ldh);
      float* V = sb.take(FRL_R * zp);
      float* dV = sb.take(FRL_R * zp);
      float* Pm = sb.take(FRL_R * zp);      // projected target distribution m
      float* A = sb.take(FRL_R * nat);
      float* dA = sb.take(FRL_R * nat);
      float* Pn = sb.take(FRL_R * zp);      // next_dist / current dist
      float* Dt = sb.take(FRL_R * zp);      // scratch per-row
      float* qv = sb.take(FRL_R * ((nA + 3) & ~3) + 4);
      int* act = (int*)sb.take(FRL_R + 4);
      float* red0 = sb.take(FRL_NT);
      float* sc2 = sb.take(2 * FRL_NT);        // row-reduction scratch of the distribution helpers
      float* proj = sb.take(3 * FRL_R * zp);   // projection scratch: l index, u index, l-weight per (row, source atom)
      float* gp = a.gpart + (size_t)c.cta * E3.n_p;
      float loss_acc = 0.f;
      bool first = true;
      if (c.cta >= ntile) return;
      for (int tile = c.cta; tile < ntile; tile += c.ncta) {
        const int row0 = tile * FRL_R;
        const int nvalid = (a.B - row0) < FRL_R ? (a.B - row0) : FRL_R;
        const frl_net_t& En = a.double_q ? a.eff[0] : a.eff[1];
        stage_prefetch(c, layer_fwd_src(En, 0), layer_fwd_bytes(En.L[0]));
        gather_rows<FRL_R>(a.replay.storage, a.replay.row_floats, a.indices + (size_t)u * a.B + row0, nvalid, raw);
        copy_cols<FRL_R>(Xo, ip, 0, raw, rf, 0, a.replay.obs_dim, ip);
        copy_cols<FRL_R>(Xn, ip, 0, raw, rf, rb_col_nobs(a.replay), a.replay.obs_dim, ip);
        // (1) next action: Double -> online net (forward #1), else the target's own argmax
        if (a.double_q) {
          layer_fwd<FRL_R>(c, a.eff[0], 0, Xn, ip, Hn, ldh, FRL_ACT_RELU, fwd_hint(a.eff[0], 1));
          heads_fwd(c, a.eff[0], Hn, ldh, V, zp, A, nat, fwd_hint(a.eff[1], 0));
          head_probs(a, V, zp, A, nat, qv, Dt, dA, sc2);
          FRL_PAR(t) {
            if (t < FRL_R) { int best = 0; for (int b = 1; b < nA; ++b) if (qv[t * nA + b] > qv[t * nA + best]) best = b; act[t] = best; }
          }
          FRL_SYNC();
        }
        // (2) target net on next_obs (forward #2) -> next_dist of the chosen action
        layer_fwd<FRL_R>(c, a.eff[1], 0, Xn, ip, Hn, ldh, FRL_ACT_RELU, fwd_hint(a.eff[1], 1));
        heads_fwd(c, a.eff[1], Hn, ldh, V, zp, A, nat, fwd_hint(E3, 0));
        if (!a.double_q) {
          head_probs(a, V, zp, A, nat, qv, Dt, dA, sc2);
          FRL_PAR(t) {
            if (t < FRL_R) { int best = 0; for (int b = 1; b < nA; ++b) if (qv[t * nA + b] > qv[t * nA + best]) best = b; act[t] = best; }
          }
          FRL_SYNC();
        }
        dist_of(a, V, zp, A, nat, act, Pn, sc2);
        // (3) projection of the target distribution  (projection_dist, DQN_with_tricks.py:135-160), two parallel phases:
        //     per source atom (r, z): l, u and the two weighted contributions; per target atom (r, j): gather, in the
        //     reference's index_add_ order (all l-contributions with z ascending, then all u-contributions) — bit-identical
        //     to the sequential scatter.  Scratch: lidx / uidx / wl in `proj`, wu in Dt (free until step 5).
        {
          int* lidx = (int*)proj;
          int* uidx = lidx + FRL_R * zp;
          float* wl = proj + 2 * FRL_R * zp;
          float* wu = Dt;
          FRL_PAR(t) {
            for (int e = t; e < FRL_R * nZ; e += FRL_NT) {
              const int r = e / nZ, z = e % nZ;
              int l = -1, uu = -1;
              float cl = 0.f, cu = 0.f;
              if (r < nvalid) {
                const float rew = raw[r * rf + rb_col_rew(a.replay)], dn = raw[r * rf + rb_col_done(a.replay)];
                float tz = fadd(rew, fmul(fmul(a.gamma, a.z[z]), fadd(1.f, -dn)));
                tz = fminf(fmaxf(tz, a.v_min), a.v_max);
                const float b = fdiv(fadd(tz, -a.v_min), a.delta_z);
                l = (int)floorf(b); uu = (int)ceilf(b);
                const float nd = Pn[r * zp + z];
                cl = fmul(fadd((float)(uu + (l == uu ? 1 : 0)), -b), nd);
                cu = fmul(fadd(b, -(float)l), nd);
              }
              lidx[r * zp + z] = l; uidx[r * zp + z] = uu; wl[r * zp + z] = cl; wu[r * zp + z] = cu;
            }
          }
          FRL_SYNC();
          FRL_PAR(t) {
            for (int e = t; e < FRL_R * zp; e += FRL_NT) {
              const int r = e / zp, jz = e % zp;
              float m = 0.f;
              if (jz < nZ) {
                for (int z = 0; z < nZ; ++z) if (lidx[r * zp + z] == jz) m = fadd(m, wl[r * zp + z]);
                for (int z = 0; z < nZ; ++z) if (uidx[r * zp + z] == jz) m = fadd(m, wu[r * zp + z]);
              }
              Pm[e] = m;
            }
          }
          FRL_SYNC();
        }
        // (4) online net on obs with the stored actions (forward #3, carries the gradient)
        layer_fwd<FRL_R>(c, E3, 0, Xo, ip, H, ldh, FRL_ACT_RELU, fwd_hint(E3, 1));
        heads_fwd(c, E3, H, ldh, V, zp, A, nat, bwd_hint(E3, 1));
        FRL_PAR(t) { if (t < FRL_R) act[t] = (int)raw[t * rf + rb_col_act(a.replay)]; }
        FRL_SYNC();
        dist_of(a, V, zp, A, nat, act, Pn, sc2);
        // (5) loss, PER error and d(loss)/d(logits of the taken action): per (r, z) terms, `lanes` partial sums per row
        {
          const int lanes = row_lanes(FRL_R);
          float* perr = sc2;                       // [R * lanes] partial sum_z m log p      (then [FRL_NT + r] = row total)
          float* pdot = dA;                        // [R * lanes] partial sum_z p dL/dp      (then [R*lanes + r] = row total)
          FRL_PAR(t) {
            if (t < FRL_R * lanes) {
              const int r = t / lanes, l = t % lanes;
              float err = 0.f, dot = 0.f;
              if (r < nvalid) {
                const float w = a.is_weight ? a.is_weight[row0 + r] : 1.f;
                for (int z = l; z < nZ; z += lanes) {
                  const float p = Pn[r * zp + z];
                  const float pc = fminf(fmaxf(p, 1e-5f), 1.f - 1e-5f);
                  err += Pm[r * zp + z] * logf(pc);
                  // dL/dp = -(m * w / B) / p inside the clamp range, 0 outside
                  const float gpz = (p >= 1e-5f && p <= 1.f - 1e-5f) ? -(Pm[r * zp + z] * w / (float)a.B) / pc : 0.f;
                  Dt[r * zp + z] = gpz;
                  dot += p * gpz;
                }
              }
              perr[t] = err; pdot[t] = dot;
            }
          }
          FRL_SYNC();
          FRL_PAR(t) {
            float l = 0.f;
            if (t < FRL_R) {
              float err = 0.f, dot = 0.f;
              for (int k = 0; k < lanes; ++k) { err += perr[t * lanes + k]; dot += pdot[t * lanes + k]; }
              pdot[FRL_R * lanes + t] = dot;
              if (t < nvalid) {
                const float w = a.is_weight ? a.is_weight[row0 + t] : 1.f;
                if (a.error_out) a.error_out[row0 + t] = err;
                l = -err * w;
              }
            }
            red0[t] = l;
          }
          FRL_SYNC();
          loss_acc += block_sum(red0);
          FRL_PAR(t) {
            for (int e = t; e < FRL_R * zp; e += FRL_NT) {
              const int r = e / zp, z = e % zp;
              dV[e] = (r < nvalid && z < nZ) ? Pn[e] * (Dt[e] - pdot[FRL_R * lanes + r]) : 0.f;      // softmax backward
            }
          }
          FRL_SYNC();
        }
        // dueling backward: dV[z] = dlogit[z];  dA[a'][z] = ([a'==a] - 1/nA) * dlogit[z]
        FRL_PAR(t) {
          for (int e = t; e < FRL_R * nat; e += FRL_NT) {
            const int r = e / nat, j = e % nat;
            float g = 0.f;
            if (j < nA * nZ) {
              const int ac = j / nZ, z = j % nZ;
              g = ((ac == act[r] ? 1.f : 0.f) - 1.f / (float)nA) * dV[r * zp + z];
            }
            dA[e] = g;
          }
        }
        FRL_SYNC();
        // (6) backward through the heads into the hidden layer, then l1
        gemm_outer<FRL_R>(dV, zp, E3.L[1].out_pad, H, ldh, E3.L[1].in_pad, E3.L[1].in, gp + E3.L[1].w_off, gp + E3.L[1].b_off, !first);
        layer_bwd_dx<FRL_R>(c, E3, 1, dV, zp, nullptr, 0, dH, ldh, bwd_hint(E3, 2));
        int col = 0;
        for (int li = 2; li < E3.n_layers; ++li) {
          const frl_layer_t& L = E3.L[li];
          gemm_outer<FRL_R>(dA + col, nat, L.out_pad, H, ldh, L.in_pad, L.in, gp + L.w_off, gp + L.b_off, !first);
          layer_bwd_dx<FRL_R>(c, E3, li, dA + col, nat, nullptr, 0, Hn, ldh, li + 1 < E3.n_layers ? bwd_hint(E3, li + 1) : no_hint());
          FRL_PAR(t) { for (int e = t; e < FRL_R * ldh; e += FRL_NT) dH[e] += Hn[e]; }
          FRL_SYNC();
          col += L.out;
        }
        FRL_PAR(t) { for (int e = t; e < FRL_R * ldh; e += FRL_NT) dH[e] = H[e] > 0.f ? dH[e] : 0.f; }
        FRL_SYNC();
        gemm_outer<FRL_R>(dH, ldh, E3.L[0].out_pad, Xo, ip, E3.L[0].in_pad, E3.L[0].in, gp + E3.L[0].w_off, gp + E3.L[0].b_off, !first);
        first = false;
      }
      FRL_PAR(t) { if (t == 0) a.stats[c.cta * 8] = loss_acc; }
      FRL_SYNC();
      return;
    }
    // stage 2: reduce effective-net gradients over CTAs, map to (mu, sigma), Adam + Polyak on the trainable block
    float* sh = c.red;
    if (c.cta == 0) {                       // loss metric: partials fetched in parallel, folded in CTA order (block-uniform branch)
      float o[3];
      cta_sums(sh, a.stats, 8, nullptr, 0, nullptr, 0, ncontrib, o);
      FRL_PAR(t) { if (t == 0) a.out[u * 8] = o[0] / (float)a.B; }
      FRL_SYNC();
    }
    FRL_PAR(t) {
      if (t == 0) {
        const AdamHP h = make_adam_hp(a.lr, a.beta1, a.beta2, a.eps_adam, 0.0, 0.0, (long)(a.step0 + u + 1));
        sh[0] = h.lr_over_bc1_neg; sh[1] = h.bc2_sqrt; sh[2] = h.one_minus_b1; sh[3] = h.b2; sh[4] = h.one_minus_b2; sh[5] = h.eps;
      }
    }
    FRL_SYNC();
    const float* eps3 = a.eps + (size_t)2 * a.eps_len;
    const float omt = (float)(1.0 - (double)a.tau);
    FRL_PAR(t) {
      for (int e = c.cta * FRL_NT + t; e < E3.n_p; e += c.ncta * FRL_NT) {
        int dst_mu = -1, dst_sg = -1;
        float noise = 0.f;
        for (int li = 0; li < E3.n_layers; ++li) {
          const frl_layer_t& L = E3.L[li];
          const frl_noisy_map_t& M = a.map[li];
          if (e >= L.w_off && e < L.w_off + L.out_pad * L.in_pad) {
            const int j = (e - L.w_off) / L.in_pad, k = (e - L.w_off) % L.in_pad;
            if (j < L.out && k < L.in) {
              dst_mu = M.mu_w + (M.row0 + j) * L.in_pad + k;
              if (M.sg_w >= 0) { dst_sg = M.sg_w + (M.row0 + j) * L.in_pad + k; noise = fmul(eps3[M.eps_out + M.row0 + j], eps3[M.eps_in + k]); }
            }
            break;
          }
          if (e >= L.b_off && e < L.b_off + L.out_pad) {
            const int j = e - L.b_off;
            if (j < L.out) {
              dst_mu = M.mu_b + M.row0 + j;
              if (M.sg_b >= 0) { dst_sg = M.sg_b + M.row0 + j; noise = eps3[M.eps_out + M.row0 + j]; }
            }
            break;
          }
        }
        if (dst_mu < 0) continue;
        float g = a.gpart[e];
        for (int cc = 1; cc < ncontrib; ++cc) g += a.gpart[(size_t)cc * E3.n_p + e];
        for (int which = 0; which < 2; ++which) {
          const int d = which == 0 ? dst_mu : dst_sg;
          if (d < 0) continue;
          const float gg = which == 0 ? g : fmul(g, noise);
          float m = a.m[d], v = a.v[d], w = a.p[d];
          m = fmaf(sh[2], gg - m, m);
          v = fadd(fmul(v, sh[3]), fmul(fmul(sh[4], gg), gg));
          const float denom = fadd(fdiv(fsqrt(v), sh[1]), sh[5]);
          w = fadd(w, fdiv(fmul(sh[0], m), denom));
          a.m[d] = m; a.v[d] = v; a.p[d] = w;
          a.p_target[d] = fadd(fmul(a.p_target[d], omt), fmul(w, a.tau));
        }
      }
    }
    FRL_SYNC();
  }
};

// standalone noise application + greedy action (select_action of the Rainbow agent): effective net = eff[0]
struct RainbowInferAlgo {
  struct Args { frl_rainbow_args_t r; const float* obs; int n; float* out; };
  static const int NSTAGES = 1;
  FRL_SHD int wbuf_floats(const Args& a) { return RainbowAlgo::wbuf_floats(a.r); }
  FRL_SHD int user_floats(const Args& a) { return RainbowAlgo::user_floats(a.r); }
  FRL_SHD int grid(const Args& a, int) { return (a.n + FRL_R - 1) / FRL_R; }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV void stage(int, int, Cta& c, float* user, const Args& a) {
    const frl_net_t& E = a.r.eff[0];
    const int ldh = act_ld(E.L[0].out_pad), ip = E.L[0].in_pad, zp = (a.r.n_atoms + 3) & ~3, nat = RainbowAlgo::natot(a.r), nA = a.r.n_actions;
    SmemBump sb; sb.p = user;
    float* X = sb.take(FRL_R * ip);
    float* H = sb.take(FRL_R * ldh);
    float* V = sb.take(FRL_R * zp);
    float* A = sb.take(FRL_R * nat);
    float* qv = sb.take(FRL_R * ((nA + 3) & ~3) + 4);
    float* mean = sb.take(FRL_R * zp);
    float* L = sb.take(FRL_R * nat);
    float* sc2 = sb.take(2 * FRL_NT);
    const int row0 = c.cta * FRL_R;
    const int nvalid = (a.n - row0) < FRL_R ? (a.n - row0) : FRL_R;
    stage_prefetch(c, layer_fwd_src(E, 0), layer_fwd_bytes(E.L[0]));
    FRL_PAR(t) {
      for (int e = t; e < FRL_R * ip; e += FRL_NT) {
        const int r = e / ip, j = e % ip;
        X[e] = (r < nvalid && j < a.r.replay.obs_dim) ? a.obs[(size_t)(row0 + r) * a.r.replay.obs_dim + j] : 0.f;
      }
    }
    FRL_SYNC();
    layer_fwd<FRL_R>(c, E, 0, X, ip, H, ldh, FRL_ACT_RELU, fwd_hint(E, 1));
    RainbowAlgo::heads_fwd(c, E, H, ldh, V, zp, A, nat, no_hint());
    RainbowAlgo::head_probs(a.r, V, zp, A, nat, qv, mean, L, sc2);
    FRL_PAR(t) {
      if (t < nvalid) {
        int best = 0;
        for (int b = 1; b < nA; ++b) if (qv[t * nA + b] > qv[t * nA + best]) best = b;
        a.out[row0 + t] = (float)best;
      }
    }
    FRL_SYNC();
  }
};

struct NoisyApplyAlgo {        // one effective net refresh as its own launch (select_action path)
  struct Args { frl_rainbow_args_t r; int f; };
  static const int NSTAGES = 1;
  FRL_SHD int wbuf_floats(const Args&) { return 32; }
  FRL_SHD int user_floats(const Args&) { return 64; }
  FRL_SHD int grid(const Args& a, int) { return (a.r.eff[a.f].n_p + FRL_NT - 1) / FRL_NT; }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV void stage(int, int, Cta& c, float*, const Args& a) { RainbowAlgo::noisy_apply(a.r, a.f, c.cta, c.ncta); }
};
