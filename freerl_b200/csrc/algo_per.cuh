// Device sum-tree for prioritized replay, bit-exact with the reference's float64 array heap.
//   SumTree            DQN_file/Buffer.py:134-194   (tree[2*cap-1] f64, leaf i at i+cap-1, NO power-of-two padding,
//                                                   ancestors maintained by `+= change`, never re-summed)
//   PER_Buffer         DQN_file/Buffer.py:66-132    (stratified sample, IS weights, update_priorities)
// Bit-exactness: ancestors receive their `change`s in batch order.  Different nodes are independent, so the batch is
// applied level by level: within a level every distinct node has one "leader" lane that adds the changes of all
// batch items mapping to it IN BATCH ORDER (fp64 addition is not associative — order is what makes the last ulp match).
#pragma once
#include "launch.cuh"

#ifndef FRL_EMUL
FRL_DEV double f64mul(double a, double b) { return __dmul_rn(a, b); }
FRL_DEV double f64add(double a, double b) { return __dadd_rn(a, b); }
FRL_DEV double f64div(double a, double b) { return __ddiv_rn(a, b); }
#else
static inline double f64mul(double a, double b) { volatile double r = a * b; return r; }
static inline double f64add(double a, double b) { volatile double r = a + b; return r; }
static inline double f64div(double a, double b) { return a / b; }
#endif

#define FRL_PER_MAXB 1024

// ---- ordered batch update ------------------------------------------------------------------------------------
struct TreeUpdateArgs {
  double* tree; int64_t cap;
  const int64_t* idx;        // [B] buffer indices
  const float* pri32;        // [B] new priorities (fp32, widened exactly) or nullptr
  const double* pri64;       // [1] a single fp64 priority for every item (PER add: max priority) or nullptr
  double pri_const;          // used when both are null
  int64_t idx0; int idx_is_range;   // idx_is_range: item i targets (idx0 + i) % cap (batched ring add)
  int B;
};

struct TreeUpdateAlgo {
  typedef TreeUpdateArgs Args;
  static const int NSTAGES = 1;
  FRL_SHD int wbuf_floats(const Args&) { return 32; }
  FRL_SHD int user_floats(const Args& a) { return 2 * (2 * a.B + 2 * a.B) + 64; }     // node[B] (i64) + change[B] (f64)
  FRL_SHD int grid(const Args&, int) { return 1; }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV void stage(int, int, Cta&, float* user, const Args& a) {
    int64_t* node = (int64_t*)user;                 // current node of item i (-1: finished)
    double* change = (double*)(node + a.B);
    const int B = a.B;
    // leaf phase: change_i = p_i - (value the leaf holds when item i is applied)
    FRL_PAR(t) {
      for (int i = t; i < B; i += FRL_NT) {
        const int64_t bi = a.idx_is_range ? (a.idx0 + i) % a.cap : a.idx[i];
        node[i] = bi + a.cap - 1;
      }
    }
    FRL_SYNC();
    FRL_PAR(t) {
      for (int i = t; i < B; i += FRL_NT) {
        const double p = a.pri32 ? (double)a.pri32[i] : (a.pri64 ? a.pri64[0] : a.pri_const);
        int prev = -1;
        for (int j = i - 1; j >= 0; --j) if (node[j] == node[i]) { prev = j; break; }
        const double before = prev >= 0 ? (a.pri32 ? (double)a.pri32[prev] : p) : a.tree[node[i]];
        change[i] = f64add(p, -before);
      }
    }
    FRL_SYNC();
    FRL_PAR(t) {
      for (int i = t; i < B; i += FRL_NT) {
        bool last = true;
        for (int j = i + 1; j < B; ++j) if (node[j] == node[i]) { last = false; break; }
        if (last) a.tree[node[i]] = a.pri32 ? (double)a.pri32[i] : (a.pri64 ? a.pri64[0] : a.pri_const);
      }
    }
    FRL_SYNC();
    // Ancestors.  With key = heap index + 1 the parent of a node is key >> 1, so the ancestor of item i at level L is
    // (key_i >> L) - 1 and two items meet at level L iff their keys agree after the shift.  A node's new value depends only
    // on its old value and the batch-ordered changes of the items below it, never on other levels: every (item, level)
    // pair is processed independently in ONE phase (no per-level barrier, all tree loads in flight together).  The pair is
    // the node's "leader" when no earlier item shares the node; the leader adds the changes of all its items IN BATCH
    // ORDER (fp64 addition is not associative: the order is what makes the last ulp match the sequential reference).
    FRL_PAR(t) {
      for (int i = t; i < B; i += FRL_NT) node[i] = node[i] + 1;       // node[] now holds the leaf KEY
    }
    FRL_SYNC();
    int maxlev = 0;
    { uint64_t k = (uint64_t)(2 * a.cap - 1); while (k > 1) { k >>= 1; ++maxlev; } }   // depth of the deepest leaf
    FRL_PAR(t) {
      for (int it = t; it < B * maxlev; it += FRL_NT) {
        const int L = it / B + 1, i = it - (L - 1) * B;               // consecutive threads: consecutive items of one level
        const int64_t k = node[i] >> L;
        if (k < 1) continue;
        bool leader = true;
        for (int j = 0; j < i; ++j) if ((node[j] >> L) == k) { leader = false; break; }
        if (!leader) continue;
        double v = f64add(a.tree[k - 1], change[i]);
        for (int j = i + 1; j < B; ++j) if ((node[j] >> L) == k) v = f64add(v, change[j]);
        a.tree[k - 1] = v;
      }
    }
    FRL_SYNC();
  }
};

// ---- stratified sampling + importance weights --------------------------------------------------------------------
struct TreeSampleArgs {
  const double* tree; int64_t cap;
  const double* u;           // [B] uniforms in [0,1) (numpy legacy random_sample in parity mode) or nullptr -> Philox
  uint64_t seed, counter;
  int B;
  int64_t size;              // len(buffer)
  double beta, prob_floor;
  int64_t* out_idx;          // [B]
  float* out_pri;            // [B] leaf priorities (fp32 container, Buffer.py:104)
  float* out_w;              // [B] importance weights (fp32)
};

struct TreeSampleAlgo {
  typedef TreeSampleArgs Args;
  static const int NSTAGES = 1;
  FRL_SHD int wbuf_floats(const Args&) { return 32; }
  FRL_SHD int user_floats(const Args& a) { return 2 * a.B + 2 * FRL_NT + 64; }
  FRL_SHD int grid(const Args&, int) { return 1; }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV void stage(int, int, Cta&, float* user, const Args& a) {
    double* w = (double*)user;                      // [B]
    double* redm = w + a.B;                         // [FRL_NT]
    const int64_t nnodes = 2 * a.cap - 1;
    const double total = a.tree[0];
    const double seg = f64div(total, (double)a.B);    // segment = sumtree.sum() / batch_size
    FRL_PAR(t) {
      double mx = 0.0;
      for (int i = t; i < a.B; i += FRL_NT) {
        double ui;
        if (a.u) ui = a.u[i];
        else {
          uint32_t o[4];
          frl_philox((uint32_t)a.seed, (uint32_t)(a.seed >> 32), (uint32_t)i, (uint32_t)a.counter, (uint32_t)(a.counter >> 32), 0x9e3779b9u, o);
          ui = ((double)(((uint64_t)(o[0] >> 5) << 26) | (o[1] >> 6))) * (1.0 / 9007199254740992.0);
        }
        const double lo = f64mul(seg, (double)i), hi = f64mul(seg, (double)(i + 1));
        double s = f64add(lo, f64mul(f64add(hi, -lo), ui));       // np.random.uniform(a, b) = a + (b-a)*random_sample()
        int64_t nd = 0;
        while (2 * nd + 1 < nnodes) {                        // SumTree.get (Buffer.py:168-188)
          const int64_t left = 2 * nd + 1;
          const double tl = a.tree[left];
          if (s <= tl) nd = left;
          else { s = f64add(s, -tl); nd = left + 1; }
        }
        const float p32 = (float)a.tree[nd];
        a.out_idx[i] = nd - a.cap + 1;
        a.out_pri[i] = p32;
        double prob = f64div((double)p32, total);
        if (prob < a.prob_floor) prob = a.prob_floor;
        const double wi = pow(f64mul((double)a.size, prob), -a.beta);
        w[i] = wi;
        mx = wi > mx ? wi : mx;
      }
      redm[t] = mx;
    }
    FRL_SYNC();
    for (int s2 = FRL_NT / 2; s2 > 0; s2 >>= 1) {
      FRL_PAR(t) { if (t < s2) redm[t] = redm[t] > redm[t + s2] ? redm[t] : redm[t + s2]; }
      FRL_SYNC();
    }
    FRL_PAR(t) { for (int i = t; i < a.B; i += FRL_NT) a.out_w[i] = (float)f64div(w[i], redm[0]); }
    FRL_SYNC();
  }
};

// ---- max over leaves (np.max(tree[-cap:])) ------------------------------------------------------------------------
// phase 0: CTA-strided coalesced sweep over the leaves -> part[cta];  phase 1 (one CTA): max of the partials -> out.
// (max is order independent, so any partition gives np.max's result exactly.)
struct TreeMaxArgs { const double* tree; int64_t cap; double* part; int nblk; double* out; int phase; };
struct TreeMaxAlgo {
  typedef TreeMaxArgs Args;
  static const int NSTAGES = 1;
  FRL_SHD int wbuf_floats(const Args&) { return 32; }
  FRL_SHD int user_floats(const Args&) { return 2 * FRL_NT + 64; }
  FRL_SHD int grid(const Args& a, int) { return a.phase == 0 ? a.nblk : 1; }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV void stage(int, int, Cta& c, float* user, const Args& a) {
    double* red = (double*)user;
    FRL_PAR(t) {
      double m = -1e300;
      if (a.phase == 0) {
        const double* leaves = a.tree + (a.cap - 1);
        for (int64_t i = (int64_t)c.cta * FRL_NT + t; i < a.cap; i += (int64_t)a.nblk * FRL_NT) { const double v = leaves[i]; m = v > m ? v : m; }
      } else {
        for (int i = t; i < a.nblk; i += FRL_NT) { const double v = a.part[i]; m = v > m ? v : m; }
      }
      red[t] = m;
    }
    FRL_SYNC();
    for (int s2 = FRL_NT / 2; s2 > 0; s2 >>= 1) {
      FRL_PAR(t) { if (t < s2) red[t] = red[t] > red[t + s2] ? red[t] : red[t + s2]; }
      FRL_SYNC();
    }
    FRL_PAR(t) { if (t == 0) { if (a.phase == 0) a.part[c.cta] = red[0]; else a.out[0] = red[0]; } }
    FRL_SYNC();
  }
};

// ---- priorities  (|td| + eps) ** alpha  on fp32 (numpy computes x ** 0.5 as sqrt — bit-exact for the default alpha) ----
struct PriBody {
  const float* td; float eps; float alpha; float* out;
  FRL_HDM void operator()(long i) const {
    const float x = fabsf(td[i]) + eps;
    out[i] = (alpha == 0.5f) ? sqrtf(x) : (alpha == 1.0f ? x : (float)pow((double)x, (double)alpha));
  }
};
