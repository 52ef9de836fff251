// Replay add with the copy engine (cp.async.bulk, SASS UBLKCP) on both sides of the AoS <-> SoA transpose (sm_100a, GPU only).
//
// The five dense tensors are contiguous over a tile of rows, and so are the ring rows of an add; a tile therefore moves as
//     add:     5 bulk loads (field chunks) -> linear staging -> [transpose in shared memory] -> row block -> 1 bulk store (2 at the wrap)
// and the SM only executes the shared-memory transpose.  Two stages per CTA: the loads of tile k + 1 are in flight while tile k is
// transposed, and the store of tile k drains while tile k + 1 is transposed (cp.async.bulk.wait_group.read before a stage is reused).
// Bulk copies need 16-byte aligned addresses and sizes: full 64-row tiles of 16-byte aligned tensors always are (64 * w * 4 bytes);
// the last, partial tile moves its dense side with plain loads / stores.  Unaligned tensors take the tile kernels of capi.cu.
#pragma once
#include "umma.cuh"

#define FRL_RB_ROWS 64          // rows per tile
#define FRL_RB_NT 256

UM_DEV void rb_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(um_smem_u32(src_smem)), "r"(bytes) : "memory");
}
UM_DEV void rb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
UM_DEV void rb_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

struct RbGeom {
  int RF, used, w[5], off[5], fb[5];      // ring-row floats, payload floats, field widths / row offsets / staging bases (floats, x rows)
};
UM_DEV RbGeom rb_geom(const frl_replay_t& rb) {
  RbGeom g;
  g.RF = rb.row_floats;
  g.used = 2 * rb.obs_dim + rb.act_dim + 2;
  int b = 0;
#pragma unroll
  for (int f = 0; f < 5; ++f) {
    replay_field(rb, f, &g.w[f], &g.off[f]);
    g.fb[f] = b;
    b += g.w[f] * FRL_RB_ROWS;
  }
  return g;
}

// shared memory of one CTA: the dense-side staging blocks [64 x used], the row blocks [64 x RF], 4 mbarriers, 64 row indices
#define FRL_RB_ADD_STG 4        // add: staging stages (bulk loads in flight: 3 tiles per CTA), 2 row blocks
static inline int rb_stage_floats(const frl_replay_t& rb) { return (FRL_RB_ROWS * (2 * rb.obs_dim + rb.act_dim + 2) + 3) & ~3; }
static inline int rb_map_words(const frl_replay_t& rb) { return ((FRL_RB_ROWS * rb.row_floats + 511) / 512) * 256; }
static inline int rb_smem_bytes(const frl_replay_t& rb) {
  const int st = rb_stage_floats(rb), rw = FRL_RB_ROWS * rb.row_floats;
  return (FRL_RB_ADD_STG * st + 2 * rw) * 4 + 64 + rb_map_words(rb) * 4;
}

// The transpose of a tile is the same permutation for every tile, so a CTA computes it once: map word m = t + 256 i of thread t holds
// the staging offsets (or 0xFFFF for the padding columns of a ring row) of row-block elements e0 = t + 512 i and e1 = e0 + 256.  Per
// tile a thread then issues 1 map load + 2 loads + 2 stores per two elements, all of them conflict-free across a warp (consecutive
// lanes, consecutive words) except at field and row boundaries.
UM_DEV uint32_t rb_partner(const RbGeom& g, int e) {
  if (e >= FRL_RB_ROWS * g.RF) return 0xFFFFu;
  const int row = e / g.RF, c = e - row * g.RF;
  if (c >= g.used) return 0xFFFFu;
  int f = 0;
#pragma unroll
  for (int k = 1; k < 5; ++k) f += (c >= g.off[k]) ? 1 : 0;
  int fb = 0, w = 0, off = 0;
#pragma unroll
  for (int k = 0; k < 5; ++k) if (k == f) { fb = g.fb[k]; w = g.w[k]; off = g.off[k]; }
  return (uint32_t)(fb + row * w + c - off);
}
UM_DEV void rb_build_map(uint32_t* map, const RbGeom& g, int t) {
  const int nwords = ((FRL_RB_ROWS * g.RF + 511) / 512) * 256;
  for (int m = t, e0 = t; m < nwords; m += FRL_RB_NT, e0 += 2 * FRL_RB_NT) map[m] = rb_partner(g, e0) | (rb_partner(g, e0 + FRL_RB_NT) << 16);
}
UM_DEV void rb_transpose(const float* stg, float* rows, const uint32_t* map, const RbGeom& g, int t) {
  const int nrow = FRL_RB_ROWS * g.RF;                         // a multiple of 256: e0 is always inside the row block
  const int nwords = (nrow + 511) / 512 * 256;
  for (int m = t, e0 = t; m < nwords; m += FRL_RB_NT, e0 += 2 * FRL_RB_NT) {
    const uint32_t u = map[m], s0 = u & 0xFFFFu, s1 = u >> 16;
    const float v0 = (s0 != 0xFFFFu) ? stg[s0] : 0.f, v1 = (s1 != 0xFFFFu) ? stg[s1] : 0.f;
    rows[e0] = v0;
    if (e0 + FRL_RB_NT < nrow) rows[e0 + FRL_RB_NT] = v1;
  }
}

__global__ void __launch_bounds__(FRL_RB_NT, 3) frl_replay_add_bulk_kernel(const __grid_constant__ ReplayTileArgs a) {
  extern __shared__ __align__(128) float rb_sm[];
  const RbGeom g = rb_geom(a.rb);
  const int t = (int)threadIdx.x, G = (int)gridDim.x;
  const int stf = (FRL_RB_ROWS * g.used + 3) & ~3, rwf = FRL_RB_ROWS * g.RF;
  float* const stg0 = rb_sm;                                   // FRL_RB_ADD_STG staging blocks
  float* const rows0 = rb_sm + FRL_RB_ADD_STG * stf;           // 2 row blocks
  uint64_t* full = reinterpret_cast<uint64_t*>(rows0 + 2 * rwf);
  uint32_t* map = reinterpret_cast<uint32_t*>(rows0 + 2 * rwf + 16);
  const float* src[5] = {a.obs, a.act, a.rew, a.done, a.nobs};
  const int ntiles = (a.n + FRL_RB_ROWS - 1) / FRL_RB_ROWS;
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < FRL_RB_ADD_STG; ++i) um_mbar_init(&full[i], 1);
    um_fence_mbar_init();
  }
  rb_build_map(map, g, t);
  __syncthreads();

  auto request = [&](int tile, int s) {                 // thread 0: bulk loads of a FULL tile (a partial tile is read with plain loads)
    if (tile >= ntiles || a.n - tile * FRL_RB_ROWS < FRL_RB_ROWS) return;
    const size_t r0 = (size_t)tile * FRL_RB_ROWS;
    um_mbar_expect_tx(&full[s], (uint32_t)(FRL_RB_ROWS * g.used * 4));
#pragma unroll
    for (int f = 0; f < 5; ++f) um_bulk_g2s(stg0 + s * stf + g.fb[f], src[f] + r0 * g.w[f], (uint32_t)(FRL_RB_ROWS * g.w[f] * 4), &full[s]);
  };

  int tile = (int)blockIdx.x;
  if (t == 0)
    for (int i = 0; i < FRL_RB_ADD_STG - 1; ++i) request(tile + i * G, i);
  for (int k = 0; tile < ntiles; ++k, tile += G) {
    const int s = k % FRL_RB_ADD_STG, r0 = tile * FRL_RB_ROWS, nr = (a.n - r0 < FRL_RB_ROWS) ? a.n - r0 : FRL_RB_ROWS;
    float* stg = stg0 + s * stf;
    float* rows = rows0 + (k & 1) * rwf;
    if (t == 0) {
      request(tile + (FRL_RB_ADD_STG - 1) * G, (k + FRL_RB_ADD_STG - 1) % FRL_RB_ADD_STG);   // that stage was released by the barrier that ended tile k-1
      rb_bulk_wait_read<1>();                                                                // the store that read this row block (tile k-2) is done
    }
    if (nr == FRL_RB_ROWS) {
      um_mbar_wait(&full[s], (uint32_t)((k / FRL_RB_ADD_STG) & 1));
    } else {
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        const float* p = src[f] + (size_t)r0 * g.w[f];
        for (int e = t; e < nr * g.w[f]; e += FRL_RB_NT) stg[g.fb[f] + e] = p[e];
      }
    }
    __syncthreads();
    rb_transpose(stg, rows, map, g, t);
    um_fence_proxy_async();
    __syncthreads();
    if (t == 0) {
      int64_t slot = a.index + r0;
      if (slot >= a.rb.capacity) slot %= a.rb.capacity;
      const int first = (int)((a.rb.capacity - slot < nr) ? a.rb.capacity - slot : nr);
      rb_bulk_s2g(a.rb.storage + slot * g.RF, rows, (uint32_t)(first * g.RF * 4));
      if (first < nr) rb_bulk_s2g(a.rb.storage, rows + first * g.RF, (uint32_t)((nr - first) * g.RF * 4));
      rb_bulk_commit();
    }
  }
  if (t == 0) rb_bulk_wait_read<0>();
}

// The gather stays with the tile kernel of capi.cu.  Two copy-engine variants were built and measured at 2^20 rows of a 704 MB ring
// (profiles/r3_replay_bulk_ncu.json): one bulk load per sampled row is bound by the copy engine's issue rate (118 us against 92 us), and
// LDGSTS row loads + bulk stores of the dense side run at 93 - 98 us — no better, because the gather is bound by DRAM, not by the SM:
// a random 176-byte row costs 288 bytes of DRAM reads (the fills are 128-byte lines: 2.25 lines per 16-byte aligned row on average),
// so the launch moves 464 MB for 369 MB of algorithmic bytes at 4.7 - 5.0 TB/s.

static inline bool rb_aligned16(const void* p) { return ((size_t)p & 15) == 0; }
static inline bool rb_bulk_ok(const ReplayTileArgs& a) {
  static const bool off = getenv("FREERL_B200_REPLAY_TILES") != nullptr;      // A/B switch: force the tile kernels of capi.cu
  if (off) return false;
  return rb_aligned16(a.obs) && rb_aligned16(a.act) && rb_aligned16(a.rew) && rb_aligned16(a.nobs) && rb_aligned16(a.done) &&
         rb_aligned16(a.rb.storage) && a.rb.row_floats % 4 == 0;
}
static int rb_launch_add(const ReplayTileArgs& a, cudaStream_t s) {
  void (*kern)(const ReplayTileArgs) = frl_replay_add_bulk_kernel;
  const int smem = rb_smem_bytes(a.rb);
  if (smem > 227 * 1024 || a.rb.row_floats > 1020) return 1;                      // rows too wide for two stages: the caller falls back to the tile kernels
  FRL_SMEM_OPT_IN(kern, smem, 48 * 1024);
  const long tiles = ((long)a.n + FRL_RB_ROWS - 1) / FRL_RB_ROWS;
  int per_sm = (228 * 1024) / (smem + 1024);
  if (per_sm > 3) per_sm = 3;
  if (per_sm < 1) per_sm = 1;
  const long cap = (long)frl_device_max_ctas() * per_sm;
  kern<<<(int)(tiles < cap ? tiles : cap), FRL_RB_NT, smem, s>>>(a);
  FRL_CUDA_OK(cudaGetLastError());
  ++frl_launch_counter;
  return 0;
}
