// freerl_b200 — C ABI implementation (see include/freerl_b200.h).
// Compiled by nvcc for sm_100a (product: libfreerl_b200.so).  The same translation unit compiles with
// `g++ -x c++ -DFRL_EMUL` into tests/emul/libfreerl_emul.so — a TEST-ONLY host emulation used by the CPU test
// suite to check kernel indexing/arithmetic against the oracle; frl_is_emulation() tells them apart and the
// product Python path refuses a library that reports 1.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "algo_ppo.cuh"
#include "algo_sacd.cuh"
#include "algo_acfx.cuh"
#include "algo_per.cuh"
#include "algo_rainbow.cuh"
#include "algo_vec.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

extern "C" void frl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* frl_last_error(void) { return g_err; }
extern "C" int frl_abi_version(void) { return 10; }
// sizeof() of the argument structs, so a binding can verify its mirror of the layout before the first call
extern "C" int frl_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(frl_layer_t);
    case 1: return (int)sizeof(frl_net_t);
    case 2: return (int)sizeof(frl_replay_t);
    case 3: return (int)sizeof(frl_dqn_args_t);
    case 4: return (int)sizeof(frl_ac_args_t);
    case 5: return (int)sizeof(frl_infer_args_t);
    case 6: return (int)sizeof(frl_ppo_args_t);
    case 7: return (int)sizeof(frl_noisy_map_t);
    case 8: return (int)sizeof(frl_rainbow_args_t);
    case 10: return (int)sizeof(frl_sacd_args_t);
    case 9: return (int)sizeof(frl_explore_args_t);
    case 11: return (int)sizeof(frl_replica_avg_args_t);
    default: return -1;
  }
}

#ifndef FRL_EMUL
extern "C" int frl_is_emulation(void) { return 0; }
int frl_device_max_ctas() {
  static int cached = 0;
  if (!cached) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
    cached = n;
  }
  return cached;
}
#else
extern "C" int frl_is_emulation(void) { return 1; }
#endif
long long frl_launch_counter = 0;
// kernels launched by the library since it was loaded (host-side count at the five launch sites; 0 forever = nothing ran on the GPU)
extern "C" long long frl_launch_count(void) { return frl_launch_counter; }
extern "C" int frl_device_sm_count(void) { return frl_device_max_ctas(); }
extern "C" int frl_wt_ld(int out_pad) { return wt_ld_of(out_pad); }

// debug: timestamps (id, globaltimer ns) from CTA 0 of the persistent kernels into a device int64 buffer [2*2000]
extern "C" int frl_debug_set_timing(void* dev_buf) {
#ifndef FRL_EMUL
  long long* p = (long long*)dev_buf;
  FRL_CUDA_OK(cudaMemcpyToSymbol(frl_dbg_ptr, &p, sizeof(p)));
#else
  (void)dev_buf;
#endif
  return 0;
}

// debug (-DFRL_TRACE builds only): which CTA of the persistent kernels writes the op trace
extern "C" int frl_debug_set_trace_cta(int cta) {
#if !defined(FRL_EMUL) && defined(FRL_TRACE)
  FRL_CUDA_OK(cudaMemcpyToSymbol(frl_trace_cta, &cta, sizeof(cta)));
#else
  (void)cta;
#endif
  return 0;
}

// ------------------------------------------------------------------------------------------------
// generic 1-D elementwise launcher (functor bodies are shared by the CUDA and the emulation build)
// ------------------------------------------------------------------------------------------------
#ifndef FRL_EMUL
template <class F>
__global__ void frl_for_kernel(long n, F f) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) f(i);
}
template <class F>
static int frl_for(long n, const F& f, cudaStream_t s) {
  if (n <= 0) return 0;
  long blocks = (n + 255) / 256;
  const long cap = (long)frl_device_max_ctas() * 16;     // grid-stride beyond 16 CTAs / SM
  if (blocks > cap) blocks = cap;
  frl_for_kernel<F><<<(unsigned)blocks, 256, 0, s>>>(n, f);
  FRL_CUDA_OK(cudaGetLastError());
  ++frl_launch_counter;
  return 0;
}
#else
template <class F>
static int frl_for(long n, const F& f, cudaStream_t) {
  for (long i = 0; i < n; ++i) f(i);
  return 0;
}
#endif

// ------------------------------------------------------------------------------------------------
// plain tile launcher for the streaming helpers: A::run(cta, ncta, smem, args) with a caller-chosen grid and a small dynamic
// shared-memory tile (no engine state, several CTAs per SM)
// ------------------------------------------------------------------------------------------------
#ifndef FRL_EMUL
template <class A>
__global__ void __launch_bounds__(FRL_NT, A::MIN_CTAS) frl_simple_kernel(const __grid_constant__ typename A::Args a) {
  extern __shared__ __align__(16) float frl_smem_tile[];
  float* frl_smem = frl_smem_tile;
  A::run((int)blockIdx.x, (int)gridDim.x, frl_smem, a);
}
template <class A>
static int frl_launch_simple(const typename A::Args& a, int grid, int smem_floats, cudaStream_t s) {
  const int smem_bytes = smem_floats * 4;
  if (smem_bytes > 227 * 1024) { frl_set_error("tile of %d B does not fit in shared memory", smem_bytes); return -3; }
  FRL_SMEM_OPT_IN(frl_simple_kernel<A>, smem_bytes, 48 * 1024);
  frl_simple_kernel<A><<<grid, FRL_NT, smem_bytes, s>>>(a);
  FRL_CUDA_OK(cudaGetLastError());
  ++frl_launch_counter;
  return 0;
}
#else
template <class A>
static int frl_launch_simple(const typename A::Args& a, int grid, int smem_floats, cudaStream_t) {
  std::vector<float> sm((size_t)smem_floats + 4);
  for (int g = 0; g < grid; ++g) {
    for (size_t i = 0; i < sm.size(); ++i) sm[i] = NAN;
    A::run(g, grid, sm.data(), a);
  }
  return 0;
}
#endif

// ------------------------------------------------------------------------------------------------
// replay: batched add (SoA fields -> AoS rows at ring slots) and gather (AoS rows -> 5 dense tensors)
// ------------------------------------------------------------------------------------------------
// A CTA moves a tile of FRL_RT_ROWS transitions through shared memory, so that BOTH sides of the AoS <-> SoA transpose are
// 16-byte coalesced accesses: the five dense tensors (odd widths such as obs_dim 17) are contiguous over a tile of rows and
// go through float4 loads / stores whenever the tile start is 16-byte aligned (always for 64-row tiles of an aligned
// tensor); the ring rows (16-byte multiples) are float4 on the storage side.  Every thread issues all loads of its share
// of the tile before the barrier (~10 independent 16-byte requests per thread in flight), the grid is a multiple of the
// SM count and tiles are taken grid-stride.
#define FRL_RT_ROWS 64
struct ReplayTileArgs {
  frl_replay_t rb; int64_t index; const int64_t* idx;
  const float *obs, *act, *rew, *nobs, *done;      // gather writes through these (cast away in the body)
  int n;
};
static int replay_tile_grid(int n) {
  const long tiles = ((long)n + FRL_RT_ROWS - 1) / FRL_RT_ROWS;
  const long cap = (long)frl_device_max_ctas() * 8;
  return (int)(tiles < cap ? tiles : cap);
}
// field f of a row: {obs, act, rew, done, next_obs} -> (width, column offset inside the ring row)
FRL_HD void replay_field(const frl_replay_t& rb, int f, int* width, int* off) {
  const int od = rb.obs_dim, ad = rb.act_dim;
  if (f == 0) { *width = od; *off = 0; }
  else if (f == 1) { *width = ad; *off = od; }
  else if (f == 2) { *width = 1; *off = od + ad; }
  else if (f == 3) { *width = 1; *off = od + ad + 1; }
  else { *width = od; *off = od + ad + 2; }
}
// Index arithmetic of the tile walkers: a thread visits elements e = e0, e0 + S, e0 + 2S, ... of a [rows][w] field; (row, col) advance by
// the constant (S / w, S % w) with one carry — two divisions per field and tile instead of one per visited element (ncu of the first tile
// kernels: issue-active 62 % / 57 % for what is a copy; the per-element divisions and address products were most of it).
struct FieldWalk {
  unsigned row, col, drow, dcol, w;
  FRL_DEVM void init(unsigned e0, unsigned step, unsigned w_) {
    w = w_; row = e0 / w_; col = e0 - row * w_; drow = step / w_; dcol = step - drow * w_;
  }
  FRL_DEVM void next() { row += drow; col += dcol; if (col >= w) { col -= w; ++row; } }
};

struct ReplayAddTiles {
  typedef ReplayTileArgs Args;
  static const int MIN_CTAS = 8;
  FRL_SDEV void run(int cta, int ncta, float* sm, const Args& a) {
    const int RF = a.rb.row_floats, nq = RF >> 2, used = 2 * a.rb.obs_dim + a.rb.act_dim + 2;
    const float* src[5] = {a.obs, a.act, a.rew, a.done, a.nobs};
    const int ntiles = (a.n + FRL_RT_ROWS - 1) / FRL_RT_ROWS;
    for (int tile = cta; tile < ntiles; tile += ncta) {
      const int r0 = tile * FRL_RT_ROWS, nr = (a.n - r0 < FRL_RT_ROWS) ? a.n - r0 : FRL_RT_ROWS;
      FRL_PAR(t) {
        for (int f = 0; f < 5; ++f) {
          int w, off;
          replay_field(a.rb, f, &w, &off);
          const float* p = src[f] + (size_t)r0 * w;
          const int ne = nr * w, n4 = (((size_t)p & 15) == 0) ? (ne >> 2) : 0;
          FieldWalk fw;
          fw.init(4u * (unsigned)t, 4u * FRL_NT, (unsigned)w);
          for (int q = t; q < n4; q += FRL_NT, fw.next()) {
            const float4 v = ld4(p + 4 * q);
            const float e4[4] = {v.x, v.y, v.z, v.w};
            unsigned col = fw.col, addr = fw.row * RF + off + fw.col;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              sm[addr] = e4[k];
              ++addr;
              if (++col == (unsigned)w) { col = 0; addr += RF - w; }
            }
          }
          for (int e = 4 * n4 + t; e < ne; e += FRL_NT) {
            const unsigned row = (unsigned)e / (unsigned)w, col = (unsigned)e - row * (unsigned)w;
            sm[row * RF + off + col] = p[e];
          }
        }
        for (int e = t; e < nr * (RF - used); e += FRL_NT) {        // zero the padding columns of the ring row
          const int row = e / (RF - used), col = e - row * (RF - used);
          sm[row * RF + used + col] = 0.f;
        }
      }
      FRL_SYNC();
      FRL_PAR(t) {
        FieldWalk rw;                                              // (row, 16-byte column) of the ring rows, nq quads per row
        rw.init((unsigned)t, FRL_NT, (unsigned)nq);
        int64_t slot0 = a.index + r0;
        for (int q = t; q < nr * nq; q += FRL_NT, rw.next()) {
          int64_t slot = slot0 + rw.row;
          if (slot >= a.rb.capacity) slot %= a.rb.capacity;
          st4(a.rb.storage + slot * RF + 4 * rw.col, lds4(sm + q * 4));
        }
      }
      FRL_SYNC();
    }
  }
};
struct ReplayGatherTiles {
  typedef ReplayTileArgs Args;
  static const int MIN_CTAS = 8;
  FRL_SDEV void run(int cta, int ncta, float* sm, const Args& a) {
    const int RF = a.rb.row_floats, nq = RF >> 2;
    float* dst[5] = {(float*)a.obs, (float*)a.act, (float*)a.rew, (float*)a.done, (float*)a.nobs};
    const int ntiles = (a.n + FRL_RT_ROWS - 1) / FRL_RT_ROWS;
    for (int tile = cta; tile < ntiles; tile += ncta) {
      const int r0 = tile * FRL_RT_ROWS, nr = (a.n - r0 < FRL_RT_ROWS) ? a.n - r0 : FRL_RT_ROWS;
      FRL_PAR(t) {
        FieldWalk rw;
        rw.init((unsigned)t, FRL_NT, (unsigned)nq);
        for (int q = t; q < nr * nq; q += FRL_NT, rw.next())
          sts4(sm + q * 4, ld4(a.rb.storage + a.idx[r0 + rw.row] * RF + 4 * rw.col));
      }
      FRL_SYNC();
      FRL_PAR(t) {
        for (int f = 0; f < 5; ++f) {
          int w, off;
          replay_field(a.rb, f, &w, &off);
          float* p = dst[f] + (size_t)r0 * w;
          const int ne = nr * w, n4 = (((size_t)p & 15) == 0) ? (ne >> 2) : 0;
          FieldWalk fw;
          fw.init(4u * (unsigned)t, 4u * FRL_NT, (unsigned)w);
          for (int q = t; q < n4; q += FRL_NT, fw.next()) {
            float e4[4];
            unsigned col = fw.col, addr = fw.row * RF + off + fw.col;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              e4[k] = sm[addr];
              ++addr;
              if (++col == (unsigned)w) { col = 0; addr += RF - w; }
            }
            st4(p + 4 * q, make_float4(e4[0], e4[1], e4[2], e4[3]));
          }
          for (int e = 4 * n4 + t; e < ne; e += FRL_NT) {
            const unsigned row = (unsigned)e / (unsigned)w, col = (unsigned)e - row * (unsigned)w;
            p[e] = sm[row * RF + off + col];
          }
        }
      }
      FRL_SYNC();
    }
  }
};

#ifndef FRL_EMUL
#include "replay_bulk.cuh"
#endif

extern "C" int frl_replay_add_batch(const frl_replay_t* rb, int64_t index, const float* obs, const float* act, const float* rew,
                                    const float* next_obs, const float* done, int n, void* stream) {
  if (!rb || !rb->storage || n < 0 || rb->row_floats % 4 || rb->row_floats < 2 * rb->obs_dim + rb->act_dim + 2) {
    frl_set_error("frl_replay_add_batch: bad arguments");
    return -1;
  }
  if (n == 0) return 0;
  ReplayTileArgs a = {*rb, index, nullptr, obs, act, rew, next_obs, done, n};
#ifndef FRL_EMUL
  if (rb_bulk_ok(a)) {                                   // copy-engine path (replay_bulk.cuh); 1 = rows too wide for its two stages
    const int rc = rb_launch_add(a, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
#endif
  return frl_launch_simple<ReplayAddTiles>(a, replay_tile_grid(n), FRL_RT_ROWS * rb->row_floats, (cudaStream_t)stream);
}

extern "C" int frl_replay_gather(const frl_replay_t* rb, const int64_t* indices, int B, float* obs, float* act, float* rew,
                                 float* next_obs, float* done, void* stream) {
  if (!rb || !rb->storage || B < 0) { frl_set_error("frl_replay_gather: bad arguments"); return -1; }
  if (B == 0) return 0;
  ReplayTileArgs a = {*rb, 0, indices, obs, act, rew, next_obs, done, B};
  return frl_launch_simple<ReplayGatherTiles>(a, replay_tile_grid(B), FRL_RT_ROWS * rb->row_floats, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// uniform sampling without replacement on the device (fast mode; NOT the numpy legacy stream)
// One CTA per update.  Sparse regime (size >= 2B, every replay that has outgrown its first batches): draw B indices, then
// redraw any element that collides with a lower-indexed one until all are distinct (expected < 2 rounds).  Dense regime
// (size < 2B, e.g. the reference's first learns where B = min(len, batch_size) == len): rejection would need
// coupon-collector many rounds, so a partial Fisher-Yates shuffle of [0, size) in shared memory is used instead.
// ------------------------------------------------------------------------------------------------
struct SampleAlgo {
  struct Args { int64_t* out; int64_t size; int B, n_updates; uint64_t seed, counter; };
  static const int NSTAGES = 1;
  FRL_SHD int wbuf_floats(const Args&) { return 32; }
  FRL_SHD bool dense(const Args& a) { return a.size < 2 * (int64_t)a.B; }
  FRL_SHD int user_floats(const Args& a) { return (dense(a) ? (int)a.size : 2 * a.B) + 64; }
  FRL_SHD int grid(const Args& a, int) { return a.n_updates; }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV int64_t draw(const Args& a, int u, int i, int round) { return draw_below(a, u, i, round, (uint64_t)a.size); }
  FRL_SDEV int64_t draw_below(const Args& a, int u, int i, int round, uint64_t range) {
    uint32_t o[4];
    frl_philox((uint32_t)a.seed, (uint32_t)(a.seed >> 32), (uint32_t)i, (uint32_t)round, (uint32_t)(a.counter + u),
               (uint32_t)((a.counter + u) >> 32) ^ 0x1d8e4e27u, o);
    const uint64_t x = ((uint64_t)o[0] << 32) | o[1];
#ifndef FRL_EMUL
    return (int64_t)__umul64hi(x, range);
#else
    return (int64_t)(((unsigned __int128)x * (unsigned __int128)range) >> 64);
#endif
  }
  FRL_SDEV void stage(int, int, Cta& c, float* user, const Args& a) {
    const int u = c.cta;
    int* idx = (int*)user;            // sizes < 2^31 (checked by the host wrapper)
    if (dense(a)) {
      const int n = (int)a.size;
      FRL_PAR(t) { for (int i = t; i < n; i += FRL_NT) idx[i] = i; }
      FRL_SYNC();
      FRL_PAR(t) {
        if (t == 0) {
          for (int i = 0; i < a.B; ++i) {
            const int j = i + (int)draw_below(a, u, i, 0, (uint64_t)(n - i));
            const int vi = idx[i];
            idx[i] = idx[j];
            idx[j] = vi;
          }
        }
      }
      FRL_SYNC();
      FRL_PAR(t) { for (int i = t; i < a.B; i += FRL_NT) a.out[(size_t)u * a.B + i] = idx[i]; }
      FRL_SYNC();
      return;
    }
    int* dup = idx + a.B;
    FRL_PAR(t) { for (int i = t; i < a.B; i += FRL_NT) { idx[i] = (int)draw(a, u, i, 0); dup[i] = 0; } }
    FRL_SYNC();
    for (int round = 1; round < 64; ++round) {
      FRL_PAR(t) {
        for (int i = t; i < a.B; i += FRL_NT) {
          int d = 0;
          const int v = idx[i];
          for (int j = 0; j < i; ++j) d |= (idx[j] == v);
          dup[i] = d;
        }
      }
      FRL_SYNC();
      int any = 0;   // block-uniform after the reduction below
      FRL_PAR(t) { if (t == 0) { int s = 0; for (int i = 0; i < a.B; ++i) s |= dup[i]; dup[a.B] = s; } }
      FRL_SYNC();
      any = dup[a.B];
      if (!any) break;
      FRL_PAR(t) { for (int i = t; i < a.B; i += FRL_NT) if (dup[i]) idx[i] = (int)draw(a, u, i, round); }
      FRL_SYNC();
    }
    FRL_PAR(t) { for (int i = t; i < a.B; i += FRL_NT) a.out[(size_t)u * a.B + i] = idx[i]; }
    FRL_SYNC();
  }
};

// Large batches (B > 8192, e.g. one 65 536-row update per vector step): out[u][i] = P_u(i), i < B, where P_u is a keyed
// BIJECTION of [0, size) — a 6-round balanced Feistel network on 2h >= log2(size) bits with cycle walking (re-encrypt until the
// value falls below `size`; < 4 expected rounds).  Distinct by construction, O(1) per element, one thread per element.
FRL_HD uint64_t frl_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct FeistelBody {
  int64_t* out; uint64_t size; int B; uint64_t seed, counter; int half_bits;
  FRL_HDM void operator()(long e) const {
    const long u = e / B;
    const uint64_t i = (uint64_t)(e - u * B);
    const uint64_t key = frl_mix64(seed ^ frl_mix64(counter + (uint64_t)u + 0x9E3779B97F4A7C15ull));
    const uint64_t mask = (1ull << half_bits) - 1;
    uint64_t x = i;
    do {
      uint64_t L = x >> half_bits, R = x & mask;
      for (int r = 0; r < 6; ++r) {
        const uint64_t f = frl_mix64(R ^ (key + (uint64_t)r * 0xD1B54A32D192ED03ull)) & mask;
        const uint64_t nl = R;
        R = L ^ f;
        L = nl;
      }
      x = (L << half_bits) | R;
    } while (x >= size);
    out[e] = (int64_t)x;
  }
};

extern "C" int frl_sample_uniform(int64_t* indices_out, int64_t size, int B, int n_updates, uint64_t seed, uint64_t counter,
                                  void* stream) {
  if (!indices_out || size <= 0 || size >= ((int64_t)1 << 31) || B <= 0 || B > size || n_updates <= 0) {
    frl_set_error("frl_sample_uniform: need 0 < B <= size < 2^31 (got size=%lld B=%d)", (long long)size, B);
    return -1;
  }
  if (B > 8192) {
    int bits = 1;
    while (((int64_t)1 << bits) < size) ++bits;
    FeistelBody b = {indices_out, (uint64_t)size, B, seed, counter, (bits + 1) / 2};
    return frl_for((long)B * n_updates, b, (cudaStream_t)stream);
  }
  SampleAlgo::Args a = {indices_out, size, B, n_updates, seed, counter};
  return frl_launch_tiles<SampleAlgo>(a, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// refresh the transposed mirror block from the trainable block (after load_state_dict / init)
// ------------------------------------------------------------------------------------------------
struct MirrorBody {
  frl_net_t n;
  FRL_HDM void operator()(long p) const {
    for (int li = 0; li < n.n_layers; ++li) {
      const frl_layer_t& L = n.L[li];
      const int wsz = L.out_pad * L.in_pad;
      if (p >= L.w_off && p < L.w_off + wsz) {
        const int e = (int)p - L.w_off, j = e / L.in_pad, k = e % L.in_pad;
        n.pt[L.wt_off + k * wt_ld(L) + j] = n.p[p];
        return;
      }
      if (p >= L.b_off && p < L.b_off + L.out_pad) { n.pt[L.wt_off + wt_bias(L) + ((int)p - L.b_off)] = n.p[p]; return; }
    }
  }
};

extern "C" int frl_net_sync_mirror(const frl_net_t* net, void* stream) {
  if (!net || !net->p || !net->pt) { frl_set_error("frl_net_sync_mirror: null net"); return -1; }
  MirrorBody b = {*net};
  return frl_for(net->n_p, b, (cudaStream_t)stream);
}

// audit hook for fast mode: the N(0,1) values the learn kernels draw for (seed, stream, counter), element idx = 0 .. n-1 — the same
// randn_ni call, so a test can hand the exact noise of a fast-mode learn to the oracle (tests/test_parity_ac.py)
struct RandnBody {
  uint64_t seed; uint32_t stream, ctr; float* out;
  FRL_DEVM void operator()(long i) const { out[i] = frl_randn(seed, stream, ctr, (uint32_t)i); }
};
extern "C" int frl_debug_randn(uint64_t seed, uint32_t stream, uint32_t ctr, long long n, float* out, void* cuda_stream) {
  if (!out || n <= 0) { frl_set_error("frl_debug_randn: bad arguments"); return -1; }
  RandnBody b = {seed, stream, ctr, out};
  return frl_for((long)n, b, (cudaStream_t)cuda_stream);
}

struct PolyakBody {
  frl_net_t src, tgt; float tau, omt;
  FRL_DEVM void operator()(long p) const {
    const float tw = fadd(fmul(tgt.p[p], omt), fmul(src.p[p], tau));
    tgt.p[p] = tw;
    for (int li = 0; li < src.n_layers; ++li) {
      const frl_layer_t& L = src.L[li];
      const int wsz = L.out_pad * L.in_pad;
      if (p >= L.w_off && p < L.w_off + wsz) {
        const int e = (int)p - L.w_off, j = e / L.in_pad, k = e % L.in_pad;
        tgt.pt[L.wt_off + k * wt_ld(L) + j] = tw;
        return;
      }
      if (p >= L.b_off && p < L.b_off + L.out_pad) { tgt.pt[L.wt_off + wt_bias(L) + ((int)p - L.b_off)] = tw; return; }
    }
  }
};

extern "C" int frl_polyak(const frl_net_t* src, const frl_net_t* target, float tau, void* stream) {
  if (!src || !target || !src->p || !target->p || !target->pt || src->n_p != target->n_p) { frl_set_error("frl_polyak: bad arguments"); return -1; }
  PolyakBody b = {*src, *target, tau, (float)(1.0 - (double)tau)};
  return frl_for(src->n_p, b, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// batched policy inference (select_action / evaluate_action for N vectorised envs)
// ------------------------------------------------------------------------------------------------
// HM = 1: tanh hidden activations (frl_infer_args_t.hidden_tanh); 0: the kernel every other caller launches.  RT = rows per CTA tile:
// 8, or 16 for batches of thousands of rows (every tile streams the whole network out of L2 — 16-row tiles halve that traffic; taken
// for n >= 2048 when the head is at most 16 wide, one output element per thread)
template <int HM, int RT = FRL_R>
struct InferAlgoT {
  typedef frl_infer_args_t Args;
  static const int NSTAGES = 1;
  FRL_SHD int nl_of(const Args& a) { return a.nl > 0 ? a.nl : a.net.n_layers; }
  FRL_SHD int wbuf_floats(const Args& a) { return (AcAlgo::max_layer_floats(a.net) + 31) & ~31; }
  FRL_SHD int user_floats(const Args& a) {
    return RT * (2 * a.net.L[a.l0].in_pad + 4 * act_ld(a.net.L[a.l0].out_pad) + a.net.L[a.l0 + nl_of(a) - 1].out_pad + 2 + 64) + 64;
  }
  FRL_SHD int grid(const Args& a, int) { return (a.n + RT - 1) / RT; }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV void stage(int, int, Cta& c, float* user, const Args& a) {
    const frl_net_t& n = a.net;
    const int nl = nl_of(a), l0 = a.l0;
    const int in_pad = n.L[l0].in_pad, ldh = act_ld(n.L[l0].out_pad), op = n.L[l0 + nl - 1].out_pad, nout = n.L[l0 + nl - 1].out;
    SmemBump sb; sb.p = user;
    float* X = sb.take(RT * in_pad);
    float* H1 = sb.take(RT * ldh);
    float* H2 = sb.take(RT * ldh);
    float* O = sb.take(RT * op);
    NetBufs nb;
    nb.H1 = H1; nb.H2 = H2;
    nb.X0 = sb.take(RT * in_pad); nb.H1n = sb.take(RT * ldh); nb.H2n = sb.take(RT * ldh);
    nb.rs1 = sb.take(RT); nb.rs2 = sb.take(RT); nb.scratch = sb.take(RT * 64);
    const int row0 = c.cta * RT;
    const int nvalid = (a.n - row0) < RT ? (a.n - row0) : RT;
    stage_prefetch(c, layer_fwd_src(n, l0), layer_fwd_bytes(n.L[l0]));
    FRL_PAR(t) {
      for (int e = t; e < RT * in_pad; e += FRL_NT) {
        const int r = e / in_pad, j = e % in_pad;
        float x = (r < nvalid && j < a.obs_dim) ? a.obs[(size_t)(row0 + r) * a.obs_dim + j] : 0.f;
        if (a.obs_norm && r < nvalid && j < a.obs_dim)       // Batch_ObsNorm with update=False (e.g. DDPG.py:168-169)
          x = fdiv(fadd(x, -a.obs_norm[j]), fadd(a.obs_norm[2 * a.obs_dim + j], 1e-8f));
        X[e] = x;
      }
    }
    FRL_SYNC();
    if (a.layer_norm && nl == 3) net_fwd<RT>(c, n, l0, a.layer_norm, X, in_pad, a.obs_dim, nb, ldh, O, op, no_hint());
    else mlp_fwd<RT, HM>(c, n, l0, nl, X, in_pad, H1, H2, ldh, O, op, FRL_ACT_NONE, no_hint());
    FRL_PAR(t) {
      if (a.mode == FRL_INFER_ARGMAX) {
        if (t < nvalid) {
          int best = 0; float bv = O[t * op];
          for (int j = 1; j < nout; ++j) if (O[t * op + j] > bv) { bv = O[t * op + j]; best = j; }   // first max, like torch.argmax
          a.out[(size_t)(row0 + t) * a.out_cols] = (float)best;
        }
      } else if (a.mode == FRL_INFER_ARGMAX_DUELING) {
        if (t < nvalid) {
          const int na = nout - 1;
          float sA = 0.f;
          for (int j = 0; j < na; ++j) sA = fadd(sA, O[t * op + 1 + j]);
          const float mean = fdiv(sA, (float)na), V = O[t * op];
          int best = 0; float bv = fadd(fadd(V, O[t * op + 1]), -mean);
          for (int j = 1; j < na; ++j) {
            const float qv = fadd(fadd(V, O[t * op + 1 + j]), -mean);
            if (qv > bv) { bv = qv; best = j; }
          }
          a.out[(size_t)(row0 + t) * a.out_cols] = (float)best;
        }
      } else if (a.mode == FRL_INFER_PPO_CAT) {
        // Categorical(logits).sample() == argmax(softmax(logits) / q), q ~ Exp(1)  (torch.multinomial, 1 draw)
        if (t < nvalid) {
          float mx = O[t * op];
          for (int j = 1; j < nout; ++j) mx = fmaxf(mx, O[t * op + j]);
          float se = 0.f;
          for (int j = 0; j < nout; ++j) se += expf(O[t * op + j] - mx);
          const float lse = mx + logf(se);
          int best = 0; float bv = -1.f;
          for (int j = 0; j < nout; ++j) {
            const float pj = expf(O[t * op + j] - lse);
            float q = a.noise ? a.noise[(size_t)(row0 + t) * nout + j] : 0.f;
            if (!a.noise) {
              uint32_t o4[4];
              frl_philox((uint32_t)a.seed, (uint32_t)(a.seed >> 32), (uint32_t)((row0 + t) * nout + j), a.counter, 4u, 0x5eed5eedu, o4);
              q = -logf(frl_u01(o4[0]));
            }
            const float v = pj / q;
            if (v > bv) { bv = v; best = j; }
          }
          a.out[(size_t)(row0 + t) * a.out_cols] = (float)best;
          a.out[(size_t)(row0 + t) * a.out_cols + 1] = O[t * op + best] - lse;
        }
      } else if (t < RT * nout) {
        const int r = t / nout, j = t % nout;
        if (r < nvalid) {
          float v = O[r * op + j];
          if (a.mode == FRL_INFER_TANH || a.mode == FRL_INFER_SAC_MEAN) v = tanhf(v);
          else if (a.mode == FRL_INFER_SAC_SAMPLE || a.mode == FRL_INFER_PPO_GAUSS) {
            const float ls = fminf(fmaxf(n.p[n.x_off + j], -20.f), 2.f);
            const float sd = expf(ls);
            const float e = a.noise ? a.noise[(size_t)(row0 + r) * nout + j]
                                    : frl_randn(a.seed, 3u, a.counter, (uint32_t)((row0 + r) * nout + j));
            if (a.mode == FRL_INFER_SAC_SAMPLE) v = tanhf(fadd(v, fmul(e, sd)));
            else {
              const float mean = tanhf(v);
              v = fadd(fmul(e, sd), mean);                                   // Normal(mean, std).sample()
              const float diff = v - mean;
              a.out[(size_t)(row0 + r) * a.out_cols + nout + j] = -(diff * diff) / (2.f * (sd * sd)) - logf(sd) - FRL_HALF_LOG_2PI;
            }
          }
          a.out[(size_t)(row0 + r) * a.out_cols + j] = v;
        }
      }
    }
    FRL_SYNC();
  }
};

typedef InferAlgoT<0> InferAlgo;

extern "C" int frl_policy_infer(const frl_infer_args_t* a, void* stream) {
  if (!a || !a->obs || !a->out || a->n <= 0 || a->l0 < 0 || a->l0 + (a->nl > 0 ? a->nl : a->net.n_layers) > a->net.n_layers) {
    frl_set_error("frl_policy_infer: bad arguments");
    return -1;
  }
  if (a->hidden_tanh) {
    if (a->layer_norm) { frl_set_error("frl_policy_infer: hidden_tanh is not available with layer_norm"); return -1; }
    return frl_launch_tiles<InferAlgoT<1> >(*a, (cudaStream_t)stream);
  }
  {
    const int nl = a->nl > 0 ? a->nl : a->net.n_layers;
    if (a->n >= 2048 && a->net.L[a->l0 + nl - 1].out <= 16) return frl_launch_tiles<InferAlgoT<0, 16> >(*a, (cudaStream_t)stream);
  }
  return frl_launch_tiles<InferAlgo>(*a, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// fused learn() entry points
// ------------------------------------------------------------------------------------------------
static int check_net(const frl_net_t& n, bool trainable, const char* who) {
  if (!n.p || !n.pt || n.n_layers < 1 || n.n_layers > FRL_MAX_LAYERS || (trainable && (!n.m || !n.v || !n.g))) {
    frl_set_error("%s: malformed net descriptor", who);
    return -1;
  }
  for (int i = 0; i < n.n_layers; ++i)
    if (n.L[i].in_pad % 4 || n.L[i].out_pad % 4 || n.L[i].w_off % 4 || n.L[i].b_off % 4 || n.L[i].wt_off % 4) {
      frl_set_error("%s: layer %d is not 16-byte aligned", who, i);
      return -1;
    }
  return 0;
}

extern "C" int frl_dqn_learn(const frl_dqn_args_t* a, void* stream) {
  if (!a || a->B <= 0 || a->n_updates <= 0 || !a->indices || !a->gpart || !a->stats || !a->out) {
    frl_set_error("frl_dqn_learn: bad arguments");
    return -1;
  }
  if (check_net(a->q, true, "frl_dqn_learn(q)") || check_net(a->q_target, false, "frl_dqn_learn(q_target)")) return -1;
  return frl_launch<DqnAlgo>(*a, (cudaStream_t)stream);
}

extern "C" int frl_sacd_learn(const frl_sacd_args_t* a, void* stream) {
  if (!a || a->B <= 0 || a->n_updates <= 0 || !a->indices || !a->gpart || !a->sumsq || !a->stats || !a->out || !a->alpha_state) {
    frl_set_error("frl_sacd_learn: bad arguments");
    return -1;
  }
  if (check_net(a->actor, true, "frl_sacd_learn(actor)") || check_net(a->critic, true, "frl_sacd_learn(critic)") ||
      check_net(a->actor_target, false, "frl_sacd_learn(actor_target)") || check_net(a->critic_target, false, "frl_sacd_learn(critic_target)"))
    return -1;
  if (a->actor.n_layers != 3 || a->critic.n_layers != 6 || a->actor.L[2].out != a->critic.L[2].out || a->critic.L[2].out != a->critic.L[5].out ||
      a->replay.act_dim != 1) {
    frl_set_error("frl_sacd_learn: actor obs->h->h->n_actions, twin critic heads obs->h->h->n_actions, 1-column action index expected");
    return -1;
  }
  return frl_launch<SacdAlgo>(*a, (cudaStream_t)stream);
}

// FREERL_B200_AC_PATH=generic forces the generic kernel (A/B timing, tests of the fallback)
extern "C" int frl_ac_path(const frl_ac_args_t* a) {
  if (!a) return 0;
  const char* e = getenv("FREERL_B200_AC_PATH");
  if (e && !strcmp(e, "generic")) return 0;
  return AcFx::eligible(*a, frl_device_max_ctas()) ? 1 : 0;
}
extern "C" long long frl_ac_ws_floats(const frl_ac_args_t* a) {
  if (!a) return 0;
  frl_ac_args_t b = *a;                 // eligibility of the shapes, whatever the scratch pointers are right now
  float dummy_ws = 0.f;
  unsigned dummy_sync = 0;
  b.ws = &dummy_ws; b.sync = &dummy_sync;
  if (!AcFx::eligible(b, frl_device_max_ctas())) return 0;
  return (long long)AcFx::ws_layout(b, frl_device_max_ctas()).total;
}

extern "C" int frl_ac_learn(const frl_ac_args_t* a, void* stream) {
  if (!a || a->B <= 0 || a->n_updates <= 0 || !a->indices || !a->gpart || !a->sumsq || !a->stats || !a->out || !a->xchg ||
      a->n_heads < 1 || a->n_heads > 2) {
    frl_set_error("frl_ac_learn: bad arguments");
    return -1;
  }
  if (check_net(a->actor, true, "frl_ac_learn(actor)") || check_net(a->critic, true, "frl_ac_learn(critic)") ||
      check_net(a->actor_target, false, "frl_ac_learn(actor_target)") || check_net(a->critic_target, false, "frl_ac_learn(critic_target)"))
    return -1;
  if (a->n_agents > FRL_MAX_AGENTS || (a->n_agents > 1 && (a->agent_index < 0 || a->agent_index >= a->n_agents))) {
    frl_set_error("frl_ac_learn: n_agents must be <= %d", FRL_MAX_AGENTS);
    return -1;
  }
  if (a->actor.n_layers != 3 || a->critic.n_layers != 3 * a->n_heads || (a->actor_kind == FRL_ACTOR_SAC && !a->alpha_state)) {
    frl_set_error("frl_ac_learn: unsupported network shape / missing alpha state");
    return -1;
  }
  // the reference's own batch sizes (B <= 256), single agent, hidden 128-128: the small-batch schedule of algo_acfx.cuh
  if (frl_ac_path(a)) return frl_launch_fx<AcFx>(*a, (cudaStream_t)stream);
  // compile-time specialisations of the same kernel source for the two single-agent families (smaller instruction footprint)
  if (a->n_agents <= 1 && !a->obs_norm[0]) {
    if (a->actor_kind == FRL_ACTOR_SAC && a->n_heads == 2) return frl_launch<AcAlgoT<1> >(*a, (cudaStream_t)stream);
    if (a->actor_kind == FRL_ACTOR_TANH) return frl_launch<AcAlgoT<2> >(*a, (cudaStream_t)stream);
  }
  return frl_launch<AcAlgo>(*a, (cudaStream_t)stream);
}

#ifndef FRL_EMUL
#include "gae_stream.cuh"
#endif

extern "C" int frl_gae(const float* reward, const float* done, const float* adv_done, const float* vs, const float* vs_next, int T,
                       int N, double gamma, double lmbda, float* adv_out, float* v_target_out, void* stream) {
  if (!reward || !done || !adv_done || !vs || !vs_next || !adv_out || !v_target_out || T <= 0 || N <= 0) {
    frl_set_error("frl_gae: bad arguments");
    return -1;
  }
  GaeArgs a = {reward, done, adv_done, vs, vs_next, T, N, gamma, lmbda, adv_out, v_target_out};
#ifndef FRL_EMUL
  if (gae_stream_ok(a)) return gae_stream_launch(a, (cudaStream_t)stream);       // inputs streamed through shared memory (gae_stream.cuh)
#endif
  if (N >= 32)                                                                     // vectorised envs: coalesced column tiles
    return frl_launch_simple<GaeTileAlgo>(a, GaeTileAlgo::grid(a), GaeTileAlgo::smem_floats(a), (cudaStream_t)stream);
  return frl_launch_tiles<GaeAlgo>(a, (cudaStream_t)stream);                       // few columns: one warp per column

}

// ---- peer-memory blocks for the in-kernel data-parallel gradient exchange (frl_dp_peers_t) ----
extern "C" int frl_dp_alloc(long long bytes, void** dev_ptr, unsigned char* ipc_handle_64) {
#ifndef FRL_EMUL
  if (bytes <= 0 || !dev_ptr || !ipc_handle_64) { frl_set_error("frl_dp_alloc: bad arguments"); return -1; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  FRL_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  FRL_CUDA_OK(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  FRL_CUDA_OK(cudaIpcGetMemHandle(&h, p));
  memcpy(ipc_handle_64, &h, 64);
  *dev_ptr = p;
  return 0;
#else
  (void)bytes; (void)dev_ptr; (void)ipc_handle_64;
  frl_set_error("frl_dp_alloc: peer memory needs the CUDA library");
  return -1;
#endif
}
extern "C" int frl_dp_open(const unsigned char* ipc_handle_64, void** dev_ptr) {
#ifndef FRL_EMUL
  if (!ipc_handle_64 || !dev_ptr) { frl_set_error("frl_dp_open: bad arguments"); return -1; }
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, 64);
  FRL_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
#else
  (void)ipc_handle_64; (void)dev_ptr;
  frl_set_error("frl_dp_open: peer memory needs the CUDA library");
  return -1;
#endif
}
extern "C" int frl_dp_close(void* peer_ptr) {
#ifndef FRL_EMUL
  if (peer_ptr) FRL_CUDA_OK(cudaIpcCloseMemHandle(peer_ptr));
#else
  (void)peer_ptr;
#endif
  return 0;
}
extern "C" int frl_dp_free(void* dev_ptr) {
#ifndef FRL_EMUL
  if (dev_ptr) FRL_CUDA_OK(cudaFree(dev_ptr));
#else
  (void)dev_ptr;
#endif
  return 0;
}

#include "replica_avg.cuh"
extern "C" int frl_replica_average(const frl_replica_avg_args_t* a, void* stream) {
  if (!a || a->n_tensors < 1 || a->n_tensors > FRL_RA_MAX_TENSORS) { frl_set_error("frl_replica_average: 1..%d tensors", FRL_RA_MAX_TENSORS); return -1; }
  if (a->dp.world < 2 || a->dp.world > FRL_DP_MAX_RANKS || a->dp.rank < 0 || a->dp.rank >= a->dp.world) { frl_set_error("frl_replica_average: bad rank / world"); return -1; }
  long long tot = 0;
  for (int i = 0; i < a->n_tensors; ++i) {
    if (!a->tensor[i] || a->n[i] <= 0) { frl_set_error("frl_replica_average: tensor %d missing", i); return -1; }
    tot += (a->n[i] + 3) & ~3;
  }
  if (tot > a->block_floats) { frl_set_error("frl_replica_average: tensors (%lld floats) exceed the exchange block (%lld)", tot, a->block_floats); return -1; }
  for (int r = 0; r < a->dp.world; ++r)
    if (!a->dp.g[r] || !a->dp.flags[r]) { frl_set_error("frl_replica_average: peer block %d missing", r); return -1; }
#ifdef FRL_EMUL
  frl_set_error("frl_replica_average: the peer-memory exchange needs the CUDA library");
  return -1;
#else
  return frl_launch<ReplicaAvgAlgo>(*a, (cudaStream_t)stream);
#endif
}

// floats of frl_ppo_args_t.umma_ws: 6 split-weight blocks + one activation scratch per CTA of the largest grid
extern "C" long long frl_ppo_umma_ws_floats(void) {
#ifndef FRL_EMUL
  return (long long)6 * UM_WS_LAYER + (long long)frl_device_max_ctas() * UM_WS_CTA;
#else
  return 0;
#endif
}

extern "C" int frl_ppo_update(const frl_ppo_args_t* a, void* stream) {
  if (!a || a->mb <= 0 || a->n_updates <= 0 || !a->indices || !a->mb_rows || !a->gpart || !a->sumsq || !a->segcnt || !a->stats ||
      !a->out || a->n_adv < 1) {
    frl_set_error("frl_ppo_update: bad arguments");
    return -1;
  }
  if (check_net(a->net, true, "frl_ppo_update(net)")) return -1;
  if (a->value_loss == 2 && !a->v_old) { frl_set_error("frl_ppo_update: value_loss 2 needs v_old"); return -1; }
  if (a->opt_repeat > 1 && a->optimizer != FRL_OPT_ADAM) { frl_set_error("frl_ppo_update: opt_repeat is for FRL_OPT_ADAM"); return -1; }
  if (a->opt_repeat > 2) { frl_set_error("frl_ppo_update: opt_repeat is 0, 1 or 2"); return -1; }
  if (a->hidden_tanh && a->layer_norm) { frl_set_error("frl_ppo_update: hidden_tanh is not available with layer_norm"); return -1; }
  if (a->dp.world > 1) {
    if (a->dp.world > FRL_DP_MAX_RANKS || a->dp.rank < 0 || a->dp.rank >= a->dp.world) { frl_set_error("frl_ppo_update: bad dp rank / world"); return -1; }
    for (int r = 0; r < a->dp.world; ++r)
      if (!a->dp.g[r] || !a->dp.flags[r]) { frl_set_error("frl_ppo_update: dp peer block %d missing", r); return -1; }
    if (a->stage_hi > 0) { frl_set_error("frl_ppo_update: the peer-memory exchange runs whole updates (stage_lo / stage_hi must be 0)"); return -1; }
#ifdef FRL_EMUL
    frl_set_error("frl_ppo_update: the peer-memory exchange needs the CUDA library");
    return -1;
#endif
  }
  if (a->net.n_layers != 6 || (a->continuous == 1 && a->net.x_len <= 0) || (a->continuous == 2 && (a->net.L[2].out & 1))) {
    frl_set_error("frl_ppo_update: net must hold actor (layers 0-2) + critic (layers 3-5)");
    return -1;
  }
  if (a->group_rows > 0) {       // MAPPO_discrete.py's episode-wide LayerNorm / scalar value loss: group mode (csrc/algo_ppo_group.cuh)
    if (a->continuous != 0 || a->n_adv != 1 || a->hidden_tanh || a->layer_norm || a->value_loss == 1 || a->dp.world > 1 ||
        a->mb % a->group_rows != 0 || a->act_cols < 1 || a->logp_cols < 1) {
      frl_set_error("frl_ppo_update: group mode is Categorical, n_adv 1, value_loss 0 / 2 / 3, single rank, minibatches of whole groups");
      return -1;
    }
    if (a->value_loss == 3 && (a->n_updates != 1 || !a->v_old)) { frl_set_error("frl_ppo_update: value_loss 3 runs one update per launch and needs v_old"); return -1; }
    if (a->group_prepass && (a->value_loss != 3 || a->stage_lo != 0 || a->stage_hi != 1)) { frl_set_error("frl_ppo_update: group_prepass needs value_loss 3 and stages [0, 1)"); return -1; }
    return frl_launch<PpoAlgoT<8, 0, 0, 1> >(*a, (cudaStream_t)stream);
  }
  if (a->value_loss == 3 || a->group_prepass) { frl_set_error("frl_ppo_update: value_loss 3 / group_prepass need group mode (group_rows)"); return -1; }
  if (a->hidden_tanh) {          // tanh hidden activations (PPO_with_tricks): compile-time variants of the 8-row-tile kernel
    if ((a->hidden_tanh & 3) == 3) return frl_launch<PpoAlgoT<8, 3> >(*a, (cudaStream_t)stream);
    if (a->hidden_tanh & 2) return frl_launch<PpoAlgoT<8, 2> >(*a, (cudaStream_t)stream);
    return frl_launch<PpoAlgoT<8, 1> >(*a, (cudaStream_t)stream);
  }
#ifndef FRL_EMUL
  // tensor-core path (tcgen05, 128-row tiles) for large minibatches over in->128->128->out networks when the caller gave the scratch
  if (um_eligible(*a)) {
    if (((uintptr_t)a->umma_ws & 511) != 0) { frl_set_error("frl_ppo_update: umma_ws must be 512-B aligned"); return -1; }
    return frl_launch<PpoAlgoT<8, 0, 1> >(*a, (cudaStream_t)stream);
  }
#endif
  // 16-row tiles for large minibatches when they fit in shared memory (checked with the launcher's own formula)
  if (a->mb >= 1024) {
    typedef PpoAlgoT<16> P16;
    const int smem_bytes = (cta_base_floats(P16::wbuf_floats(*a)) + P16::user_floats(*a)) * 4 + 64;
    if (smem_bytes <= 227 * 1024) return frl_launch<P16>(*a, (cudaStream_t)stream);
  }
  return frl_launch<PpoAlgo>(*a, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// prioritized replay: float64 sum-tree on the device
// ------------------------------------------------------------------------------------------------
extern "C" int frl_sumtree_update(double* tree, int64_t cap, const int64_t* idx, const float* pri32, const double* pri64_scalar,
                                  double pri_const, int64_t idx0, int idx_is_range, int B, double* scratch, void* stream) {
  if (!tree || cap <= 0 || cap >= ((int64_t)1 << 30) || B <= 0 || B > FRL_PER_MAXB || (!idx && !idx_is_range) || !scratch) {
    frl_set_error("frl_sumtree_update: bad arguments (B <= %d, capacity < 2^30, scratch of %d doubles)", FRL_PER_MAXB, 2 * FRL_PER_MAXB);
    return -1;
  }
  TreeUpdateArgs a = {tree, cap, idx, pri32, pri64_scalar, pri_const, idx0, idx_is_range, B, scratch, nullptr, 0.f, 0.f};
  return frl_launch<TreeUpdateAlgo>(a, (cudaStream_t)stream);
}

// PER_Buffer.update_priorities in ONE launch: priority_i = (|td_i| + eps) ^ alpha (fp32, like the reference's array arithmetic),
// then the ordered leaf / ancestor update of frl_sumtree_update.
extern "C" int frl_sumtree_update_td(double* tree, int64_t cap, const int64_t* idx, const float* td, float eps, float alpha, int B,
                                     double* scratch, void* stream) {
  if (!tree || cap <= 0 || cap >= ((int64_t)1 << 30) || B <= 0 || B > FRL_PER_MAXB || !idx || !td || !scratch) {
    frl_set_error("frl_sumtree_update_td: bad arguments (B <= %d, capacity < 2^30, scratch of %d doubles)", FRL_PER_MAXB, 2 * FRL_PER_MAXB);
    return -1;
  }
  TreeUpdateArgs a = {tree, cap, idx, nullptr, nullptr, 0.0, 0, 0, B, scratch, td, eps, alpha};
  return frl_launch<TreeUpdateAlgo>(a, (cudaStream_t)stream);
}

extern "C" int frl_sumtree_sample(const double* tree, int64_t cap, const double* u, uint64_t seed, uint64_t counter, int B, int64_t size,
                                  double beta, double prob_floor, int64_t* out_idx, float* out_pri, float* out_w, void* stream) {
  if (!tree || cap <= 0 || B <= 0 || B > 8192 || !out_idx || !out_pri || !out_w) {
    frl_set_error("frl_sumtree_sample: bad arguments");
    return -1;
  }
  TreeSampleArgs a = {tree, cap, u, seed, counter, B, size, beta, prob_floor, out_idx, out_pri, out_w};
  return frl_launch_tiles<TreeSampleAlgo>(a, (cudaStream_t)stream);
}

extern "C" int frl_sumtree_max(const double* tree, int64_t cap, double* scratch, int nscratch, double* out, void* stream) {
  if (!tree || cap <= 0 || !scratch || nscratch <= 0 || !out) { frl_set_error("frl_sumtree_max: bad arguments"); return -1; }
  int64_t want = (cap + FRL_NT * 8 - 1) / (FRL_NT * 8);            // ~8 leaves per thread
  const int64_t lim = (int64_t)4 * frl_device_max_ctas() < nscratch ? (int64_t)4 * frl_device_max_ctas() : nscratch;
  if (want > lim) want = lim;
  if (want < 1) want = 1;
  TreeMaxArgs a = {tree, cap, scratch, (int)want, out, 0};
  int rc = frl_launch_tiles<TreeMaxAlgo>(a, (cudaStream_t)stream);
  if (rc) return rc;
  a.phase = 1;
  return frl_launch_tiles<TreeMaxAlgo>(a, (cudaStream_t)stream);
}

extern "C" int frl_per_priorities(const float* td, int B, float eps, float alpha, float* out, void* stream) {
  if (!td || !out || B <= 0) { frl_set_error("frl_per_priorities: bad arguments"); return -1; }
  PriBody b = {td, eps, alpha, out};
  return frl_for(B, b, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// Rainbow
// ------------------------------------------------------------------------------------------------
static int check_rainbow(const frl_rainbow_args_t* a, const char* who) {
  if (!a || !a->p || !a->p_target || !a->eps || !a->z || a->n_actions <= 0 || a->n_atoms <= 0) {
    frl_set_error("%s: bad arguments", who);
    return -1;
  }
  for (int f = 0; f < 3; ++f)
    if (check_net(a->eff[f], false, who)) return -1;
  return 0;
}

extern "C" int frl_rainbow_learn(const frl_rainbow_args_t* a, void* stream) {
  if (check_rainbow(a, "frl_rainbow_learn")) return -1;
  if (!a->m || !a->v || !a->indices || a->B <= 0 || !a->gpart || !a->stats || !a->out) {
    frl_set_error("frl_rainbow_learn: bad arguments");
    return -1;
  }
  return frl_launch<RainbowAlgo>(*a, (cudaStream_t)stream);
}

extern "C" int frl_rainbow_act(const frl_rainbow_args_t* a, const float* obs, int n, float* out, void* stream) {
  if (check_rainbow(a, "frl_rainbow_act") || !obs || !out || n <= 0) { if (a) frl_set_error("frl_rainbow_act: bad arguments"); return -1; }
  NoisyApplyAlgo::Args na = {*a, 0};
  int rc = frl_launch_tiles<NoisyApplyAlgo>(na, (cudaStream_t)stream);
  if (rc) return rc;
  RainbowInferAlgo::Args ia = {*a, obs, n, out};
  return frl_launch_tiles<RainbowInferAlgo>(ia, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// joint advantage normalisation: out = (x - mean) / (std_unbiased + eps).  Two grid-wide launches: float64 partial
// (sum, sum of squares) per CTA over float4 grid-stride loads, then every CTA folds the partials in fixed order and
// normalises its slice.  (float64 accumulation like torch's CPU mean/std accumulators; deterministic.)
// ------------------------------------------------------------------------------------------------
FRL_NI_MISC double block_sum_f64(double* slot /*[FRL_NT] smem*/) {
  for (int s2 = FRL_NT / 2; s2 > 0; s2 >>= 1) {
    FRL_PAR(t) { if (t < s2) slot[t] += slot[t + s2]; }
    FRL_SYNC();
  }
  const double r = slot[0];
  FRL_SYNC();
  return r;
}
struct AdvNormArgs { const float* x; int n; float eps; float* out; double* part; int phase; int ncta; int nohint; };
// Two launches (a grid-wide mean / std sits between them).  Launch 0: every thread streams float4 quads with four
// independent loads in flight and accumulates sum / sum of squares in float64; per-CTA partials go to `part`.  Launch 1:
// every CTA folds the partials in the same order (identical statistics on all CTAs), then streams the input a second time
// — from the END, so that the part of `x` that launch 0 read last is still in L2 — and writes (x - mean) / (std + eps).
struct AdvNormAlgo {
  typedef AdvNormArgs Args;
  static const int MIN_CTAS = 6;
  FRL_SHD int smem_floats(const Args&) { return 4 * FRL_NT + 64; }
  FRL_SDEV void run(int cta, int ncta, float* sm, const Args& a) {
    double* s1 = (double*)sm;
    double* s2 = s1 + FRL_NT;
    const int n4 = ((((size_t)a.x | (size_t)a.out) & 15) == 0) ? (a.n >> 2) : 0;       // float4 body when 16-B aligned
    const int stride = ncta * FRL_NT;
    if (a.phase == 0) {
      FRL_PAR(t) {
        double s = 0.0, q = 0.0;
        int i = cta * FRL_NT + t;
        const unsigned long long keep = a.nohint ? 0ull : l2_policy_evict_last();          // launch 1 reads x again: ask L2 to hold on to it
        for (; i + 3 * stride < n4; i += 4 * stride) {
          const float4 v0 = ld4_hint(a.x + 4 * (size_t)i, keep), v1 = ld4_hint(a.x + 4 * (size_t)(i + stride), keep);
          const float4 v2 = ld4_hint(a.x + 4 * (size_t)(i + 2 * stride), keep), v3 = ld4_hint(a.x + 4 * (size_t)(i + 3 * stride), keep);
          const float4 vv[4] = {v0, v1, v2, v3};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 v = vv[k];
            s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
            q += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
          }
        }
        for (; i < n4; i += stride) {
          const float4 v = ld4(a.x + 4 * (size_t)i);
          s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
          q += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
        }
        for (int j = 4 * n4 + cta * FRL_NT + t; j < a.n; j += stride) { const double v = a.x[j]; s += v; q += v * v; }
        s1[t] = s; s2[t] = q;
      }
      FRL_SYNC();
      const double S = block_sum_f64(s1), Q = block_sum_f64(s2);
      FRL_PAR(t) { if (t == 0) { a.part[2 * cta] = S; a.part[2 * cta + 1] = Q; } }
      FRL_SYNC();
    } else {
      FRL_PAR(t) {
        double s = 0.0, q = 0.0;
        for (int i = t; i < ncta; i += FRL_NT) { s += a.part[2 * i]; q += a.part[2 * i + 1]; }
        s1[t] = s; s2[t] = q;
      }
      FRL_SYNC();
      const double S = block_sum_f64(s1), Q = block_sum_f64(s2);
      const double mean = S / (double)a.n;
      double var = (Q - S * mean) / (double)(a.n - 1);                                 // torch.std(): unbiased
      if (var < 0.0) var = 0.0;
      const float mf = (float)mean, den = (float)sqrt(var) + a.eps;
      FRL_PAR(t) {
        for (int j = 4 * n4 + cta * FRL_NT + t; j < a.n; j += stride) a.out[j] = fdiv(a.x[j] - mf, den);
        int i = n4 - 1 - (cta * FRL_NT + t);
        const unsigned long long drop = a.nohint ? 0ull : l2_policy_evict_first();         // last use of x; the output must not evict what is still unread
        for (; i - 3 * stride >= 0; i -= 4 * stride) {
          float4 vv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) vv[k] = ld4_hint(a.x + 4 * (size_t)(i - k * stride), drop);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float4 v = vv[k];
            v.x = fdiv(v.x - mf, den); v.y = fdiv(v.y - mf, den); v.z = fdiv(v.z - mf, den); v.w = fdiv(v.w - mf, den);
            st4_hint(a.out + 4 * (size_t)(i - k * stride), v, drop);
          }
        }
        for (; i >= 0; i -= stride) {
          float4 v = ld4(a.x + 4 * (size_t)i);
          v.x = fdiv(v.x - mf, den); v.y = fdiv(v.y - mf, den); v.z = fdiv(v.z - mf, den); v.w = fdiv(v.w - mf, den);
          st4(a.out + 4 * (size_t)i, v);
        }
      }
      FRL_SYNC();
    }
  }
};

// per-device scratch for the two-launch reductions (allocated once; 16 B per CTA)
static double* frl_reduce_scratch(int doubles) {
  static double* buf = nullptr;
  static int cap = 0;
  if (doubles > cap) {
#ifndef FRL_EMUL
    if (buf) cudaFree(buf);
    if (cudaMalloc((void**)&buf, (size_t)doubles * sizeof(double)) != cudaSuccess) { buf = nullptr; cap = 0; return nullptr; }
#else
    free(buf);
    buf = (double*)malloc((size_t)doubles * sizeof(double));
#endif
    cap = doubles;
  }
  return buf;
}

#ifndef FRL_EMUL
#include "adv_norm_resident.cuh"
#endif

extern "C" int frl_adv_norm(const float* x, int n, float eps, float* out, void* stream) {
  if (!x || !out || n < 2) { frl_set_error("frl_adv_norm: bad arguments"); return -1; }
#ifndef FRL_EMUL
  {                                                  // data resident in shared memory across the grid-wide mean / std: one read, one launch
    double* part1 = frl_reduce_scratch(4096);
    if (!part1) { frl_set_error("frl_adv_norm: scratch allocation failed"); return -2; }
    AdvNormArgs a1 = {x, n, eps, out, part1, 0, 0};
    const int rc1 = adv_norm_resident_launch(a1, (cudaStream_t)stream);
    if (rc1 <= 0) return rc1;
  }
#endif
  int ncta = (n / 16 + FRL_NT - 1) / FRL_NT;      // >= 4 quads per thread before the grid grows
  const int cap = 6 * frl_device_max_ctas();
  if (ncta > cap) ncta = cap;
  if (ncta < 1) ncta = 1;
  double* part = frl_reduce_scratch(2 * cap > 4096 ? 2 * cap : 4096);
  if (!part) { frl_set_error("frl_adv_norm: scratch allocation failed"); return -2; }
  static const bool nohint = getenv("FREERL_B200_ADVNORM_NO_HINT") != nullptr;      // A/B switch: plain loads / stores
  AdvNormArgs a = {x, n, eps, out, part, 0, ncta, nohint ? 1 : 0};
  int rc = frl_launch_simple<AdvNormAlgo>(a, ncta, AdvNormAlgo::smem_floats(a), (cudaStream_t)stream);
  if (rc) return rc;
  a.phase = 1;
  return frl_launch_simple<AdvNormAlgo>(a, ncta, AdvNormAlgo::smem_floats(a), (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// per-step train-loop work over N vectorised envs (algo_vec.cuh)
// ------------------------------------------------------------------------------------------------
static int elementwise_grid(long n) {
  long g = (n + FRL_NT - 1) / FRL_NT;
  const long cap = 4L * frl_device_max_ctas();
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
extern "C" int frl_vecnorm(double* state, int64_t n0, const void* x, int x_is_f64, int N, int D, int update, float* out, double* out64,
                           void* stream) {
  if (!state || !x || N < 0 || D <= 0 || n0 < 0 || (!out && !out64)) { frl_set_error("frl_vecnorm: bad arguments"); return -1; }
  if (N == 0) return 0;
  VecNormArgs a = {state, n0, x, x_is_f64, N, D, update, out, out64};
  if (update) return frl_launch_simple<VecNormFold>(a, (D + FRL_NT - 1) / FRL_NT, 4, (cudaStream_t)stream);
  return frl_launch_simple<VecNormApply>(a, elementwise_grid((long)N * D), 4, (cudaStream_t)stream);
}
extern "C" int frl_reward_scaling(double* state, int64_t n0, double* R, const void* x, int x_is_f64, double gamma, int N, float* out,
                                  double* out64, void* stream) {
  if (!state || !R || !x || N < 0 || n0 < 0 || (!out && !out64)) { frl_set_error("frl_reward_scaling: bad arguments"); return -1; }
  if (N == 0) return 0;
  RewardScaleArgs a = {state, n0, R, x, x_is_f64, gamma, N, out, out64};
  return frl_launch_simple<RewardScaleFold>(a, 1, 4, (cudaStream_t)stream);
}
extern "C" int frl_explore(const frl_explore_args_t* a, void* stream) {
  if (!a || !a->action || a->N < 0 || a->A <= 0 || (a->kind != 0 && a->kind != 1) || (a->kind == 0 && !a->ou_state) ||
      (!a->out && !a->out64)) {
    frl_set_error("frl_explore: bad arguments");
    return -1;
  }
  if (a->N == 0) return 0;
  return frl_launch_simple<ExploreAlgo>(*a, elementwise_grid((long)a->N * a->A), 4, (cudaStream_t)stream);
}
extern "C" int frl_masked_reset(double* state, const uint8_t* mask, int N, int W, double value, void* stream) {
  if (!state || !mask || N < 0 || W <= 0) { frl_set_error("frl_masked_reset: bad arguments"); return -1; }
  if (N == 0) return 0;
  MaskedResetArgs a = {state, mask, N, W, value};
  return frl_launch_simple<MaskedReset>(a, elementwise_grid((long)N * W), 4, (cudaStream_t)stream);
}

extern "C" int frl_epsilon_greedy(const int64_t* greedy, int N, int n_actions, double epsilon, const double* u, const int64_t* rnd,
                                  uint64_t seed, uint64_t counter, int64_t* out, void* stream) {
  if (!greedy || !out || N < 0 || n_actions <= 0 || (u && !rnd)) { frl_set_error("frl_epsilon_greedy: bad arguments"); return -1; }
  if (N == 0) return 0;
  EpsGreedyArgs a = {greedy, N, n_actions, epsilon, u, rnd, seed, counter, out};
  return frl_launch_simple<EpsGreedyAlgo>(a, elementwise_grid(N), 4, (cudaStream_t)stream);
}
extern "C" int frl_dis_to_con(const int64_t* action, int N, int n_actions, int shape, int per, const float* low, const float* high,
                              double* out64, float* out, void* stream) {
  if (!action || !low || !high || N < 0 || shape <= 0 || (!out && !out64) || (shape == 1 ? n_actions < 2 : per < 2)) {
    frl_set_error("frl_dis_to_con: bad arguments (needs >= 2 actions per dimension)");
    return -1;
  }
  if (N == 0) return 0;
  DisToConArgs a = {action, N, n_actions, shape, per, low, high, out64, out};
  return frl_launch_simple<DisToConAlgo>(a, elementwise_grid((long)N * shape), 4, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// debug micro-benchmark of one layer op (not part of the product API; used by tools/opbench.py)
//   mode 0: gemm_rk on already-staged weights   1: layer_fwd with TMA every iteration, no prefetch
//   mode 2: layer_fwd with prefetch of the next iteration's weights (alternating layers li and li2)
// ------------------------------------------------------------------------------------------------
struct OpBenchAlgo {
  struct Args { frl_net_t net; int li, li2, iters, mode, ncta; float* sink; };
  static const int NSTAGES = 1;
  FRL_SHD int wbuf_floats(const Args& a) { return (AcAlgo::max_layer_floats(a.net) + 31) & ~31; }
  FRL_SHD int user_floats(const Args& a) { return FRL_R * 3 * 256 + 64; }
  FRL_SHD int grid(const Args& a, int) { return a.ncta; }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV void stage(int, int, Cta& c, float* user, const Args& a) {
    const frl_layer_t& L = a.net.L[a.li];
    float* X = user;
    float* Y = user + FRL_R * 256;
    FRL_PAR(t) { for (int e = t; e < FRL_R * 256; e += FRL_NT) { X[e] = 0.001f * (float)(e % 97); Y[e] = 0.f; } }
    FRL_SYNC();
    if (a.mode == 0) {
      const float* Bs = stage_acquire(c, layer_fwd_src(a.net, a.li), layer_fwd_bytes(L));
      for (int it = 0; it < a.iters; ++it)
        gemm_rk<FRL_R>(c.red, X, L.in_pad, L.in_pad, Bs, wt_ld(L), L.out_pad, Bs + wt_bias(L), EPI_BIAS_ACT, FRL_ACT_RELU, nullptr, 0, Y, L.out_pad);
    } else if (a.mode == 1) {
      for (int it = 0; it < a.iters; ++it)
        layer_fwd<FRL_R>(c, a.net, a.li, X, L.in_pad, Y, L.out_pad, FRL_ACT_RELU, no_hint());
    } else {
      for (int it = 0; it < a.iters; ++it) {
        const int cur = (it & 1) ? a.li2 : a.li, nxt = (it & 1) ? a.li : a.li2;
        layer_fwd<FRL_R>(c, a.net, cur, X, a.net.L[cur].in_pad, Y, a.net.L[cur].out_pad, FRL_ACT_RELU,
                         it + 1 < a.iters ? fwd_hint(a.net, nxt) : no_hint());
      }
    }
    FRL_PAR(t) { if (t == 0) a.sink[c.cta] = Y[0]; }
    FRL_SYNC();
  }
};

extern "C" int frl_debug_opbench(const frl_net_t* net, int li, int li2, int iters, int mode, int ncta, float* sink, void* stream) {
  OpBenchAlgo::Args a = {*net, li, li2, iters, mode, ncta, sink};
  return frl_launch_tiles<OpBenchAlgo>(a, (cudaStream_t)stream);
}
