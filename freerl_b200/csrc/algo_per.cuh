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
// Two stages of one cooperative launch over `depth` CTAs.  Stage 0 (CTA 0): leaf phase — change_i = p_i - (value the leaf
// holds when item i is applied), leaves written, keys + changes published in `scratch`.  Stage 1: one CTA per tree DEPTH.
// With key = heap index + 1 the parent of a node is key >> 1 and a node's depth is floor(log2(key)), so the ancestor of item
// i at depth d is (key_i >> (depth_i - d)) - 1; two items meet at depth d iff these agree.  (Grouping by DEPTH, not by
// distance from the leaf: with a non-power-of-two capacity the leaves sit at two depths and e.g. the root is 16 levels above
// one leaf and 17 above another — every node must have exactly one owner.)  A node's new value depends only on its old
// value and the batch-ordered changes of the items below it, never on other depths, so the depths run concurrently on
// different SMs.  Inside a depth, item i is the node's "leader" when no earlier item shares the node; the leader adds the
// changes of all its items IN BATCH ORDER (fp64 addition is not associative: the order is what makes the last ulp match
// the sequential reference).  Keys are 32-bit (capacity < 2^30) and scanned four at a time from shared memory.
struct TreeUpdateArgs {
  double* tree; int64_t cap;
  const int64_t* idx;        // [B] buffer indices
  const float* pri32;        // [B] new priorities (fp32, widened exactly) or nullptr
  const double* pri64;       // [1] a single fp64 priority for every item (PER add: max priority) or nullptr
  double pri_const;          // used when both are null
  int64_t idx0; int idx_is_range;   // idx_is_range: item i targets (idx0 + i) % cap (batched ring add)
  int B;
  double* scratch;           // dev [FRL_PER_MAXB] changes (f64) followed by [FRL_PER_MAXB] keys (u32)
  const float* td;           // [B] TD errors or nullptr: priority = (|td| + td_eps) ^ td_alpha in fp32 (PER_Buffer.update_priorities)
  float td_eps, td_alpha;
};

// (|td| + eps) ^ alpha as the reference computes it: float32 array arithmetic (DQN_file/Buffer.py:126-129)
FRL_HD float frl_td_priority(float td, float eps, float alpha) {
  const float x = fabsf(td) + eps;
  return (alpha == 0.5f) ? sqrtf(x) : (alpha == 1.0f ? x : (float)pow((double)x, (double)alpha));
}

struct TreeUpdateAlgo {
  typedef TreeUpdateArgs Args;
  static const int NSTAGES = 2;
  FRL_SHD bool writes_params(int) { return false; }
  FRL_SHD bool stage_enabled(int, int, const Args&) { return true; }
  FRL_SHD int wbuf_floats(const Args&) { return 32; }
  FRL_SHD int bp(const Args& a) { return (a.B + 3) & ~3; }
  FRL_SHD int user_floats(const Args& a) { return 4 * bp(a) + 64; }                   // keys (u32) + shifted keys (u32) + changes (f64)
  FRL_SHD int depth(const Args& a) { int m = 0; uint64_t k = (uint64_t)(2 * a.cap - 1); while (k > 1) { k >>= 1; ++m; } return m; }
  FRL_SHD int grid(const Args& a, int max_ctas) { const int m = depth(a); return m < 1 ? 1 : (m < max_ctas ? m : max_ctas); }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV int ilog2(uint32_t x) {
#ifndef FRL_EMUL
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
  }
  FRL_SDEV double pri_of(const Args& a, int i) {
    if (a.td) return (double)frl_td_priority(a.td[i], a.td_eps, a.td_alpha);
    return a.pri32 ? (double)a.pri32[i] : (a.pri64 ? a.pri64[0] : a.pri_const);
  }
  // any j in [lo, hi) with ks[j] == k ?   (hi - lo may be anything; ks is padded to a multiple of 4 with never-matching zeros)
  FRL_SDEV bool any_equal(const uint32_t* ks, int lo, int hi, uint32_t k) {
    bool hit = false;
    int j = lo;
    for (; j < hi && (j & 3); ++j) hit |= (ks[j] == k);
    for (; j + 4 <= hi; j += 4) {
      const uint4 q = *reinterpret_cast<const uint4*>(ks + j);
      hit |= (q.x == k) | (q.y == k) | (q.z == k) | (q.w == k);
    }
    for (; j < hi; ++j) hit |= (ks[j] == k);
    return hit;
  }
  FRL_SDEV void stage(int s, int, Cta& c, float* user, const Args& a) {
    const int B = a.B, Bp = bp(a);
    uint32_t* key = (uint32_t*)user;                // leaf KEY (= heap index + 1) of item i
    uint32_t* ks = key + Bp;                        // key >> L of the level in flight
    double* change = (double*)(ks + Bp);
    uint32_t* gkey = (uint32_t*)(a.scratch + FRL_PER_MAXB);
    if (s == 0) {
      if (c.cta != 0) return;
      FRL_PAR(t) {
        for (int i = t; i < Bp; i += FRL_NT) {
          const int64_t bi = a.idx_is_range ? (a.idx0 + i) % a.cap : (i < B ? a.idx[i] : 0);
          key[i] = i < B ? (uint32_t)(bi + a.cap) : 0u;
        }
      }
      FRL_SYNC();
      // change_i = p_i - before_i; before_i = priority of the previous item on the same leaf, else the stored leaf
      FRL_PAR(t) {
        for (int i = t; i < B; i += FRL_NT) {
          const uint32_t k = key[i];
          int prev = -1;
          for (int j = 0; j < i; ++j) prev = (key[j] == k) ? j : prev;
          const double p = pri_of(a, i);
          const double before = prev >= 0 ? pri_of(a, prev) : a.tree[(int64_t)k - 1];
          const double ch = f64add(p, -before);
          a.scratch[i] = ch;
          gkey[i] = k;
        }
      }
      FRL_SYNC();                                   // every `before` read precedes every leaf write
      FRL_PAR(t) {
        for (int i = t; i < B; i += FRL_NT)
          if (!any_equal(key, i + 1, B, key[i])) a.tree[(int64_t)key[i] - 1] = pri_of(a, i);      // the last write wins
      }
      FRL_SYNC();
      return;
    }
    const int maxdepth = depth(a);                  // depth of the deepest leaf; ancestors live at depths 0 .. maxdepth - 1
    for (int d = c.cta; d < maxdepth; d += c.ncta) {
      FRL_PAR(t) {
        for (int i = t; i < Bp; i += FRL_NT) {
          uint32_t k = 0u;
          if (i < B) {
            const uint32_t key_i = gkey[i];
            const int di = ilog2(key_i);
            if (di > d) k = key_i >> (di - d);     // 0: this item's leaf is at or above depth d (never matches)
            change[i] = a.scratch[i];
          }
          ks[i] = k;
        }
      }
      FRL_SYNC();
      FRL_PAR(t) {
        for (int i = t; i < B; i += FRL_NT) {
          const uint32_t k = ks[i];
          if (k < 1u) continue;                     // this item has no ancestor at depth d
          if (any_equal(ks, 0, i, k)) continue;     // an earlier item leads this node
          double v = f64add(a.tree[(int64_t)k - 1], change[i]);
          for (int j = i + 1; j < B; ++j) if (ks[j] == k) v = f64add(v, change[j]);
          a.tree[(int64_t)k - 1] = v;
        }
      }
      FRL_SYNC();
    }
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
  FRL_HDM void operator()(long i) const { out[i] = frl_td_priority(td[i], eps, alpha); }
};
