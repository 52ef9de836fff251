// GAE over [T, N] rollouts with the inputs streamed through shared memory by LDGSTS (cp.async, 16 B), sm_100a, GPU only.
//
// Same scan as GaeTileAlgo (algo_ppo.cuh): a CTA owns 32 adjacent env columns and walks the time axis from the end in rounds; a round
// is 8 chunks (one warp each) whose affine maps A_{t0} = S + P * A_{t1} (float64) are composed through shared memory on top of the carry
// of the later round.  What changes is how the five inputs arrive: every thread copies ONE 16-byte quad per input and round straight
// into shared memory, and the copies of round r + 1 are issued before round r is processed — each CTA keeps 20 KB of loads in flight
// while it computes, five CTAs per SM (GaeTileAlgo: loads only in flight during pass 1, ncu long_scoreboard 45 %).  A round is 32 time
// steps (4 per warp), the raw block 2 x 5 x 32 x 32 floats = 40 KB; pass 1 turns reward into the TD residual and adv_done into
// 1 - adv_done in place, pass 2 replays the chunk from there and writes adv / v_target (one 128-byte line per warp and step).
// Needs N % 4 == 0 and 16-byte aligned tensors (quads never straddle the end of a row); anything else takes GaeTileAlgo.
#pragma once

#define GAE_S_LC 4
#define GAE_S_R 32

__device__ __forceinline__ void gae_ldgsts16(float* dst_smem, const float* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem) : "memory");
}

__global__ void __launch_bounds__(256, 5) frl_gae_stream_kernel(const __grid_constant__ GaeArgs a) {
  __shared__ __align__(16) float raw[2][5][GAE_S_R][32];
  __shared__ double sP[224], sS[224], carry[2][32];       // maps of warps 1 .. 7 (nobody composes warp 0's): 5 CTAs fit on an SM
  const int t = (int)threadIdx.x, lane = t & 31, w = t >> 5;
  const int col0 = (int)blockIdx.x * 32, col = col0 + lane;
  const float* src[5] = {a.reward, a.done, a.adv_done, a.vs, a.vs_next};
  const int lrow = t >> 3, qc = t & 7;                         // this thread's quad of every input in a round
  const bool qok = col0 + 4 * qc < a.N;
  const int rounds = (a.T + GAE_S_R - 1) / GAE_S_R;
  const float g32 = (float)a.gamma;
  const double gl = a.gamma * a.lmbda;

  auto request = [&](int r) {                                  // block row j of round r = time step T - (r + 1) * 32 + j
    const long row = (long)a.T - (long)(r + 1) * GAE_S_R + lrow;
    if (r < rounds && row >= 0 && qok) {
#pragma unroll
      for (int f = 0; f < 5; ++f) gae_ldgsts16(&raw[r & 1][f][lrow][4 * qc], src[f] + (size_t)row * a.N + col0 + 4 * qc);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (t < 32) carry[0][t] = 0.0;                               // zero tail at t = T
  request(0);
  for (int r = 0; r < rounds; ++r) {
    request(r + 1);                                            // block (r + 1) & 1 held round r - 1, released by the barrier that ended it
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    float(*blk)[GAE_S_R][32] = raw[r & 1];
    const long base = (long)a.T - (long)(r + 1) * GAE_S_R;       // time step of block row 0 (negative rows: before the rollout)
    double P = 1.0, S = 0.0;
    if (col < a.N) {
#pragma unroll
      for (int k = GAE_S_LC - 1; k >= 0; --k) {
        const int j = w * GAE_S_LC + k;
        if (base + j >= 0) {
          const float v = blk[3][j][lane];
          const float td = fadd(fadd(blk[0][j][lane], fmul(fmul(g32, fadd(1.f, -blk[1][j][lane])), blk[4][j][lane])), -v);
          const float om = 1.f - blk[2][j][lane];
          const double ak = gl * (double)om;
          S = (double)td + ak * S;
          P = ak * P;
          blk[0][j][lane] = td;
          blk[2][j][lane] = om;
        }
      }
    }
    if (w > 0) { sP[t - 32] = P; sS[t - 32] = S; }
    __syncthreads();
    if (col < a.N) {
      double A = carry[r & 1][lane];
      for (int cc = 7; cc > w; --cc) A = sS[(cc - 1) * 32 + lane] + sP[(cc - 1) * 32 + lane] * A;
#pragma unroll
      for (int k = GAE_S_LC - 1; k >= 0; --k) {
        const int j = w * GAE_S_LC + k;
        if (base + j >= 0) {
          A = (double)blk[0][j][lane] + gl * (double)blk[2][j][lane] * A;
          const float af = (float)A;
          const size_t i = (size_t)(base + j) * a.N + col;
          a.adv[i] = af;
          a.v_target[i] = fadd(af, blk[3][j][lane]);
        }
      }
      if (w == 0) carry[(r + 1) & 1][lane] = A;                // value at the round's first step -> carry of the next (earlier) round
    }
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

static inline bool gae_stream_ok(const GaeArgs& a) {
  const size_t p = (size_t)a.reward | (size_t)a.done | (size_t)a.adv_done | (size_t)a.vs | (size_t)a.vs_next | (size_t)a.adv | (size_t)a.v_target;
  static const bool off = getenv("FREERL_B200_GAE_TILES") != nullptr;          // A/B switch: force GaeTileAlgo
  // below one CTA per SM the scan is latency bound and the tile kernel's longer rounds win (T 256 x N 1536: 9 us against 12 us)
  static const bool force = getenv("FREERL_B200_GAE_STREAM") != nullptr;       // tests: take this kernel at any N
  return !off && (force || (a.N + 31) / 32 >= frl_device_max_ctas()) && a.N >= 32 && a.N % 4 == 0 && (p & 15) == 0;
}
static int gae_stream_launch(const GaeArgs& a, cudaStream_t s) {
  frl_gae_stream_kernel<<<(a.N + 31) / 32, 256, 0, s>>>(a);
  FRL_CUDA_OK(cudaGetLastError());
  ++frl_launch_counter;
  return 0;
}
