// Parameter average of the off-policy replicas over peer memory (SURVEY 8e "replicas + sharded replay"; one process per GPU).
// Replaces, per sync, the all-reduce + divide pairs of torch.distributed (SAC: 5 collectives + 5 eager kernels) by ONE cooperative
// launch: stage 0 packs the rank's parameter blocks into its exchange block g[epoch & 1] (frl_dp_peers_t, the blocks of the on-policy
// gradient exchange); after the grid barrier stage 1 publishes the epoch into every peer's flag slot, waits for every peer's epoch
// and writes (sum over the ranks in rank order) / world back into the rank's own tensors with all peers' P2P loads in flight together
// — bit-identical on every rank.  The double buffer makes one flag hand-off per sync enough (a rank can only write epoch e + 2 after
// every peer published e + 1, i.e. finished reading e).
#pragma once

struct ReplicaAvgAlgo {
  typedef frl_replica_avg_args_t Args;
  static const int NSTAGES = 2;
  FRL_SHD bool writes_params(int) { return false; }
  FRL_SHD bool stage_enabled(int, int, const Args&) { return true; }
  FRL_SHD int wbuf_floats(const Args&) { return 32; }
  FRL_SHD int user_floats(const Args&) { return 64; }
  FRL_SHD long total(const Args& a) { long t = 0; for (int i = 0; i < a.n_tensors; ++i) t += (a.n[i] + 3) & ~3; return t; }
  FRL_SHD int grid(const Args& a, int max_ctas) {
    const long want = (total(a) + FRL_NT * 4 - 1) / (FRL_NT * 4);
    return want < 1 ? 1 : (want < max_ctas ? (int)want : max_ctas);
  }
  FRL_SHD int n_updates(const Args&) { return 1; }
  FRL_SDEV void stage(int s, int, Cta& c, float*, const Args& a) {
#ifndef FRL_EMUL
    const unsigned epoch = a.dp.epoch0 + 1u;
    const int world = a.dp.world, rank = a.dp.rank, tid = (int)threadIdx.x;
    const size_t par = (size_t)(epoch & 1u) * (size_t)a.block_floats;
    if (s == 0) {
      long off = 0;
      for (int i = 0; i < a.n_tensors; ++i) {
        float* dst = a.dp.g[rank] + par + off;
        for (int p = c.cta * FRL_NT + tid; p < a.n[i]; p += c.ncta * FRL_NT) dst[p] = a.tensor[i][p];
        off += (a.n[i] + 3) & ~3;
      }
      __threadfence_system();                    // the peers read this block over NVLink after the flag
      return;
    }
    if (c.cta == 0 && tid < world && tid != rank) {
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.dp.flags[tid] + rank), "r"(epoch) : "memory");
    }
    if (tid < world && tid != rank) {
      const unsigned* f = a.dp.flags[rank] + tid;
      unsigned* dead = a.dp.flags[rank] + 32;    // sticky: a peer timed out once -> later syncs do not wait again
      long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (*(volatile unsigned*)dead) { if (c.cta == 0 && a.status) a.status[0] = -1.f; break; }
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        // replicas step their own envs on the host, so they drift apart more than lock-stepped gradient exchanges do: 30 s
        if (t1 - t0 > 30000000000ll) { *(volatile unsigned*)dead = 1u; if (c.cta == 0 && a.status) a.status[0] = -1.f; break; }
      }
    }
    __syncthreads();
    const float inv = 1.0f / (float)world;
    long off = 0;
    for (int i = 0; i < a.n_tensors; ++i) {
      for (int p = c.cta * FRL_NT + tid; p < a.n[i]; p += c.ncta * FRL_NT) {
        float v[FRL_DP_MAX_RANKS];               // all peers' loads in flight together (NVLink latency once, not world times)
#pragma unroll
        for (int r = 0; r < FRL_DP_MAX_RANKS; ++r)
          if (r < world) asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v[r]) : "l"(a.dp.g[r] + par + off + p) : "memory");
        float sgm = 0.f;
#pragma unroll
        for (int r = 0; r < FRL_DP_MAX_RANKS; ++r)
          if (r < world) sgm = fadd(sgm, v[r]);  // rank order: bit-identical on every rank
        a.tensor[i][p] = fmul(sgm, inv);
      }
      off += (a.n[i] + 3) & ~3;
    }
#else
    (void)s; (void)c; (void)a;
#endif
  }
};
