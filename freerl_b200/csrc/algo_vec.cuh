// Per-step host work of the reference train loops, as device ops over N vectorised envs:
//   Normalization / RunningMeanStd      PPO_file/normalization.py:17-49 (== MAPPO_file/normalization.py), DDPG_file/DDPG.py:358-388
//   RewardScaling                       PPO_file/normalization.py:87-101
//   OUNoise / Gaussian exploration      SAC_file/SAC.py:334-355, DDPG_file/DDPG.py:334-355,519-522
//
// The reference keeps ONE running-statistics object and feeds it one observation per env step.  With N envs stepped
// together the rows of a vector step are folded IN ENV ORDER (row 0, 1, ... N-1), each row normalised with the statistics
// that include it — exactly what the reference object returns when it is called once per row — so the statistics and the
// outputs are bit-identical to the reference classes, including their dtype behaviour under NumPy 2 (float32 observations
// keep a float32 mean and a float64 S / std; the n == 1 call sets mean = std = x).  The fold is a dependent float64 chain
// per feature column (one thread per column, rows prefetched by the unrolled loop); `update == 0` (evaluation) is a plain
// elementwise kernel.
#pragma once
#include "engine.cuh"

// state layout (doubles): [0, D) mean · [D, 2D) S · [2D, 3D) std.  n lives on the host (Python int, like the reference).
struct VecNormArgs {
  double* state; int64_t n0;            // updates folded before this call
  const void* x; int x_is_f64;          // [N][D] float32 or float64
  int N, D, update;
  float* out;                           // [N][D] fp32 (what Buffer.sample would hand to the networks), may be NULL
  double* out64;                        // optional float64 copy of the normalised rows (the reference's return dtype), may be NULL
};

template <class T>
FRL_DEVM void vecnorm_column(const VecNormArgs& a, int d) {
  const T* x = (const T*)a.x;
  double mean = a.state[d], S = a.state[a.D + d], sd = a.state[2 * a.D + d];
  int64_t n = a.n0;
#pragma unroll 4
  for (int i = 0; i < a.N; ++i) {
    const T xv = x[(size_t)i * a.D + d];
    ++n;
    double y;
    if (n == 1) {                        // normalization.py:28-30: mean = x, std = x (both keep x's dtype)
      mean = (double)xv; sd = (double)xv;
      const T den = (T)sd + (T)1e-8;
      y = (double)(T)((xv - (T)mean) / den);
    } else {
      const T old = (T)mean;
      const T dlt = xv - old;
      const T m2 = old + dlt / (T)n;     // python int n is a weak scalar: the division happens in x's dtype
      S = frl_dadd(S, (double)xmul(dlt, (T)(xv - m2)));
      mean = (double)m2;
      sd = sqrt(S / (double)n);
      y = (double)(T)(xv - m2) / (sd + 1e-8);
    }
    if (a.out) a.out[(size_t)i * a.D + d] = (float)y;
    if (a.out64) a.out64[(size_t)i * a.D + d] = y;
  }
  a.state[d] = mean; a.state[a.D + d] = S; a.state[2 * a.D + d] = sd;
}

struct VecNormFold {
  typedef VecNormArgs Args;
  static const int MIN_CTAS = 1;
  FRL_SDEV void run(int cta, int, float*, const Args& a) {
    FRL_PAR(t) {
      const int d = cta * FRL_NT + t;
      if (d < a.D) {
        if (a.x_is_f64) vecnorm_column<double>(a, d); else vecnorm_column<float>(a, d);
      }
    }
  }
};

// update == False: (x - mean) / (std + 1e-8) with frozen statistics.  `n0 == 1` keeps the reference's dtype quirk (std is
// still the first observation in x's dtype); n0 == 0 divides by 1e-8 like the reference's zero-initialised std.
struct VecNormApply {
  typedef VecNormArgs Args;
  static const int MIN_CTAS = 4;
  FRL_SDEV void run(int cta, int ncta, float*, const Args& a) {
    FRL_PAR(t) {
      const long total = (long)a.N * a.D;
      for (long e = (long)cta * FRL_NT + t; e < total; e += (long)ncta * FRL_NT) {
        const int d = (int)(e % a.D);
        const double mean = a.state[d], sd = a.state[2 * a.D + d];
        double y;
        if (a.x_is_f64) {
          y = (((const double*)a.x)[e] - mean) / (sd + 1e-8);
        } else {
          const float xv = ((const float*)a.x)[e];
          if (a.n0 == 1) y = (double)((xv - (float)mean) / ((float)sd + (float)1e-8));
          else if (a.n0 == 0) y = ((double)xv - mean) / (sd + 1e-8);
          else y = (double)(xv - (float)mean) / (sd + 1e-8);
        }
        if (a.out) a.out[e] = (float)y;
        if (a.out64) a.out64[e] = y;
      }
    }
  }
};

// RewardScaling over N envs: R_i = gamma * R_i + x_i per env, then the shared RunningMeanStd (shape 1) is fed R_0 .. R_{N-1}
// in env order and x_i is divided by the std that includes R_i.  state = {mean, S, std}; all float64 like the reference
// (np.zeros(1) R, python-float rewards).  One thread: the fold is a scalar dependent chain.
struct RewardScaleArgs {
  double* state; int64_t n0;
  double* R;                            // [N] discounted return per env (in/out)
  const void* x; int x_is_f64;          // [N] rewards
  double gamma; int N;
  float* out; double* out64;
};
struct RewardScaleFold {
  typedef RewardScaleArgs Args;
  static const int MIN_CTAS = 1;
  FRL_SDEV void run(int, int, float*, const Args& a) {
    FRL_PAR(t) {
      if (t == 0) {
        double mean = a.state[0], S = a.state[1], sd = a.state[2];
        int64_t n = a.n0;
#pragma unroll 4
        for (int i = 0; i < a.N; ++i) {
          const double xv = a.x_is_f64 ? ((const double*)a.x)[i] : (double)((const float*)a.x)[i];
          const double R = frl_dadd(frl_dmul(a.gamma, a.R[i]), xv);
          a.R[i] = R;
          ++n;
          if (n == 1) { mean = R; sd = R; }
          else {
            const double old = mean;
            mean = old + (R - old) / (double)n;
            S = frl_dadd(S, frl_dmul(R - old, R - mean));
            sd = sqrt(S / (double)n);
          }
          const double y = xv / (sd + 1e-8);
          if (a.out) a.out[i] = (float)y;
          if (a.out64) a.out64[i] = y;
        }
        a.state[0] = mean; a.state[1] = S; a.state[2] = sd;
      }
    }
  }
};

// Exploration noise for N envs x A action dims, then the reference's clip:
//   kind 0 (OUNoise.noise, SAC.py:347-355):  dx = theta (mu - x) + sqrt(dt) sigma z;  x += dx;  noise = x * scale (scale < 0: none)
//                                            action_ = clip(action * max_action + noise * max_action, -max_action, max_action)
//   kind 1 (Gaussian, DDPG.py:522):          action_ = clip(action * max_action + gauss_scale * (sigma * max_action * z), ...)
// z = randn[N][A] float64 drawn by the caller in the reference's order (parity mode) or Philox(seed, counter) when NULL.
typedef frl_explore_args_t ExploreArgs;      // public struct (include/freerl_b200.h)
struct ExploreAlgo {
  typedef ExploreArgs Args;
  static const int MIN_CTAS = 4;
  FRL_SDEV void run(int cta, int ncta, float*, const Args& a) {
    FRL_PAR(t) {
      const int total = a.N * a.A;
      for (int e = cta * FRL_NT + t; e < total; e += ncta * FRL_NT) {
        const double z = a.z ? a.z[e] : (double)frl_randn(a.seed, 0x0e5cu, (uint32_t)a.counter, (uint32_t)e);
        // numpy: float32 action * python float stays float32
        const float am = fmul(a.action[e], (float)a.max_action);
        double noise;
        if (a.kind == 0) {
          const double x = a.ou_state[e];
          const double dx = frl_dadd(frl_dmul(a.theta, a.mu - x), frl_dmul(frl_dmul(sqrt(a.dt), a.sigma), z));
          const double xn = x + dx;
          a.ou_state[e] = xn;
          noise = frl_dmul(a.scale < 0.0 ? xn : frl_dmul(xn, a.scale), a.max_action);
        } else {
          noise = frl_dmul(a.gauss_scale, 0.0 + frl_dmul(frl_dmul(a.gauss_sigma, a.max_action), z));
        }
        double v = frl_dadd((double)am, noise);
        if (a.clip) v = v < -a.max_action ? -a.max_action : (v > a.max_action ? a.max_action : v);
        if (a.out64) a.out64[e] = v;
        if (a.out) a.out[e] = (float)v;
      }
    }
  }
};

// reset rows of a [N][W] float64 state where mask[i] != 0 (episode ends: OUNoise.reset -> mu, RewardScaling.reset -> 0)
struct MaskedResetArgs { double* state; const uint8_t* mask; int N, W; double value; };
struct MaskedReset {
  typedef MaskedResetArgs Args;
  static const int MIN_CTAS = 4;
  FRL_SDEV void run(int cta, int ncta, float*, const Args& a) {
    FRL_PAR(t) {
      for (int e = cta * FRL_NT + t; e < a.N * a.W; e += ncta * FRL_NT)
        if (a.mask[e / a.W]) a.state[e] = a.value;
    }
  }
};

// epsilon-greedy of the DQN mains (DQN_file/DQN.py:307-310) for N envs: out[i] = u[i] < epsilon ? random action : greedy[i].
// Parity mode passes the host-drawn u / random actions (the legacy stream consumes a randint only where u < epsilon, so the
// host draws them in env order); fast mode (u == NULL) draws both from Philox(seed, counter, i).
struct EpsGreedyArgs {
  const int64_t* greedy; int N, n_actions; double epsilon;
  const double* u; const int64_t* rnd; uint64_t seed, counter;
  int64_t* out;
};
struct EpsGreedyAlgo {
  typedef EpsGreedyArgs Args;
  static const int MIN_CTAS = 4;
  FRL_SDEV void run(int cta, int ncta, float*, const Args& a) {
    FRL_PAR(t) {
      for (int i = cta * FRL_NT + t; i < a.N; i += ncta * FRL_NT) {
        double u;
        int64_t r;
        if (a.u) { u = a.u[i]; r = a.rnd ? a.rnd[i] : 0; }
        else {
          uint32_t o[4];
          frl_philox((uint32_t)a.seed, (uint32_t)(a.seed >> 32), (uint32_t)i, (uint32_t)a.counter, (uint32_t)(a.counter >> 32), 0xe9511eedu, o);
          u = ((double)(((uint64_t)o[0] << 21) ^ (uint64_t)(o[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);    // 53-bit (0,1)
          r = (int64_t)(((uint64_t)o[2] * (uint64_t)a.n_actions) >> 32);
        }
        a.out[i] = (u < a.epsilon) ? r : a.greedy[i];
      }
    }
  }
};

// dis_to_con (DQN_file/DQN.py:195-217) for N discrete actions -> [N][shape] continuous env actions.  float32 bounds, float64
// arithmetic (np.int64 / int -> float64; * the float32 span -> float64), one-dimensional and per-dimension digit forms.
struct DisToConArgs {
  const int64_t* action; int N, n_actions, shape, per;      // per = int(n_actions ** (1 / shape)) computed by the host like the reference
  const float *low, *high;                                  // dev [shape]
  double* out64; float* out;
};
struct DisToConAlgo {
  typedef DisToConArgs Args;
  static const int MIN_CTAS = 4;
  FRL_SDEV void run(int cta, int ncta, float*, const Args& a) {
    FRL_PAR(t) {
      for (int e = cta * FRL_NT + t; e < a.N * a.shape; e += ncta * FRL_NT) {
        const int i = e / a.shape, j = e - i * a.shape;
        const float span = a.high[j] - a.low[j];
        double frac;
        if (a.shape == 1) frac = (double)a.action[i] / (double)(a.n_actions - 1);
        else {
          int64_t p = 1;
          for (int k = 0; k < j; ++k) p *= a.per;
          int64_t q = a.action[i] / p;                       // floor division: actions are >= 0
          frac = (double)(q % a.per) / (double)(a.per - 1);
        }
        const double v = frl_dadd((double)a.low[j], frl_dmul(frac, (double)span));
        if (a.out64) a.out64[e] = v;
        if (a.out) a.out[e] = (float)v;
      }
    }
  }
};
