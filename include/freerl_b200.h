/* freerl_b200 — C ABI of the B200-native FreeRL training hot path.
 *
 * Drop-in boundary: the reference (wild-firefox/FreeRL) is pure Python/PyTorch; its "plugin interface" for this
 * path is the set of Python classes/methods listed below.  A maintainer binds these entry points from Python with
 * ctypes; freerl_b200/_lib.py is exactly that binding.  Every pointer argument marked "dev"
 * is a raw CUDA device pointer (e.g. torch.Tensor.data_ptr()), borrowed for the duration of the stream operation.
 * All functions return 0 on success, <0 on error (message via frl_last_error()); none throws across the boundary.
 * No CPU fallback exists: every compute entry point launches sm_100a kernels on the given stream.
 *
 *   entry point               replaces (reference file:line)
 *   ------------------------  ---------------------------------------------------------------------------------
 *   frl_replay_add_batch      Buffer.add                     SAC_file/Buffer.py:28-38 (=DQN/TD3/DDPG/MADDPG copies)
 *   frl_replay_gather         Buffer.sample(indices)         SAC_file/Buffer.py:40-57
 *   frl_sample_uniform        np.random.choice(N,B,False)    DQN_file/DQN.py:97, SAC_file/SAC.py:213, TD3.py:183, DDPG.py:194
 *   frl_dqn_learn             DQN.learn + update_target      DQN_file/DQN.py:104-128
 *   frl_ac_learn              SAC.learn / TD3.learn / DDPG.learn   SAC_file/SAC.py:222-271, TD3_file/TD3.py:189-244,
 *                                                            DDPG_file/DDPG.py:203-233 (+ Agent.update_*, Alpha)
 *   frl_policy_infer          select_action / evaluate_action      DQN.py:70-88, SAC.py:192-204, TD3.py:163-174, DDPG.py:166-185
 *   frl_net_sync_mirror       (state_dict load -> refresh transposed weight mirrors; no reference counterpart)
 *   frl_sumtree_update/_sample/_max, frl_per_priorities
 *                             SumTree / PER_Buffer           DQN_file/Buffer.py:66-194
 *   frl_rainbow_learn/_act    Rainbow DQN.learn / select_action   DQN_file/DQN_with_tricks.py:81-160,198-284; Noisy_net.py:17-76
 *   frl_vecnorm / frl_reward_scaling / frl_explore / frl_masked_reset
 *                             Normalization, RewardScaling, OUNoise / Gaussian exploration of the train loops, for N envs
 *                                                            PPO_file/normalization.py:17-101; SAC_file/SAC.py:334-355; DDPG_file/DDPG.py:519-522
 *   frl_gae                   PPO.learn GAE loop             PPO_file/PPO.py:222-233 (MAPPO_file/MAPPO.py:362-383)
 *   frl_ppo_update            PPO.learn minibatch loop + Agent.update_ac_ + c_adamw.AdamW.step
 *                                                            PPO_file/PPO.py:245-283,145-152; PPO_file/c_adamw.py:65-122
 */
#ifndef FREERL_B200_H
#define FREERL_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRL_MAX_LAYERS 6
#define FRL_MAX_AGENTS 6

/* One linear layer inside a parameter block.  W is stored [out_pad][in_pad] row-major (torch layout, zero padded to
 * multiples of 4), bias [out_pad]; the transposed mirror WT [in_pad][out_pad] followed by a bias copy lives in `pt`. */
typedef struct {
  int in, out;
  int in_pad, out_pad;
  int w_off;   /* float offset of W   in p  */
  int b_off;   /* float offset of b   in p  */
  int wt_off;  /* float offset of WT (then bias copy) in pt */
} frl_layer_t;

/* A network = one parameter block (all tensors of one torch module, one optimiser) + its mirror.
 * Twin critics are ONE net with 6 layers (head h uses layers 3h..3h+2), like the reference's Critic module. */
typedef struct {
  float* p;    /* dev [n_p]  parameters */
  float* pt;   /* dev [n_pt] transposed mirrors (forward operand) */
  float* m;    /* dev [n_p]  Adam exp_avg      (NULL for target nets) */
  float* v;    /* dev [n_p]  Adam exp_avg_sq   (NULL for target nets) */
  float* g;    /* dev [n_p]  reduced gradient  (NULL for target nets) */
  int n_p, n_pt;
  int n_layers;
  int x_off, x_len;   /* extra vector parameter (SAC/PPO log_std); x_len = 0 if none */
  frl_layer_t L[FRL_MAX_LAYERS];
} frl_net_t;

/* Device ring replay: `storage` is [capacity][row_floats] fp32, one row per transition laid out
 * [obs(obs_dim) | action(act_dim) | reward | done | next_obs(obs_dim) | zero pad to a multiple of 4 floats]. */
typedef struct {
  float* storage;   /* dev */
  int64_t capacity;
  int row_floats, obs_dim, act_dim;
} frl_replay_t;

typedef struct {
  frl_net_t q, q_target;
  frl_replay_t replay;
  const int64_t* indices;   /* dev [n_updates][B] */
  int B, n_updates;
  float gamma, tau;
  double lr, beta1, beta2, eps;
  int64_t step0;            /* optimiser steps taken before this call */
  float* gpart;             /* dev scratch [frl_device_sm_count()][q.n_p] */
  float* stats;             /* dev scratch [frl_device_sm_count()][8] */
  float* out;               /* dev [n_updates][8]: out[u][0] = loss */
  /* non-distributional tricks of DQN_file/DQN_with_tricks.py:261-283 (all 0 / NULL = plain DQN.py) */
  int double_q;             /* Double: a* = argmax_a Q(s') of the ONLINE net, evaluated by the target net (:263-265) */
  int dueling;              /* Dueling (:60-79): last layer rows = [V | A_0..A_{n-1}] (out = 1 + n_actions), Q = V + A - mean(A) */
  const float* is_weight;   /* PER (:276-278): dev [B] IS weights; loss = mean over the [B,B] product w_j * td_i^2 */
  float* td_error;          /* dev [n_updates][B] or NULL: Q(s,a) - y per sampled row (the PER priorities' input) */
} frl_dqn_args_t;

enum { FRL_ACTOR_TANH = 0, FRL_ACTOR_SAC = 1 };

typedef struct {
  frl_net_t actor, actor_target, critic, critic_target;
  int n_heads;              /* critic heads: 1 (DDPG, TD3 w/o clip_double) or 2 */
  int actor_kind;           /* FRL_ACTOR_TANH | FRL_ACTOR_SAC */
  frl_replay_t replay;
  const int64_t* indices;   /* dev [n_updates][B] */
  int B, n_updates;
  const float* noise_next;  /* dev [n_updates][B][act_dim]: SAC eps of a', TD3 randn; NULL -> Philox(seed) */
  const float* noise_new;   /* dev [n_updates][B][act_dim]: SAC eps of the new action; NULL -> Philox(seed) */
  uint64_t seed;
  float gamma, tau;
  double lr_actor, lr_critic, beta1, beta2, eps, wd_critic;
  float max_norm;           /* clip_grad_norm_ max (0.5 in the reference), <=0 disables */
  int64_t step_actor0, step_critic0, total_it0;
  int policy_freq;          /* TD3 twin_delay: actor + Polyak when (total_it0+u+1) % policy_freq == 0; 1 = always */
  int target_smoothing;     /* TD3 policy_noise */
  float policy_noise, noise_clip, max_action, policy_noise_scale;
  /* SAC temperature: alpha_state = {log_alpha, exp_avg, exp_avg_sq, unused} on the device */
  float* alpha_state;
  int adaptive_alpha;
  double alpha_lr;
  float target_entropy;
  int64_t step_alpha0;
  float* gpart;             /* dev scratch [sm_count][max(actor.n_p, critic.n_p)] */
  float* sumsq;             /* dev scratch [sm_count] */
  float* stats;             /* dev scratch [sm_count][8] */
  float* out;               /* dev [n_updates][8]: critic_loss, actor_loss, alpha, alpha_loss, critic_gnorm, actor_gnorm, mean_entropy, 0 */
  /* ---- multi-agent (MADDPG, MADDPG_file/MADDPG.py:186-228): n_agents > 1 ----
   * the critic sees cat(obs_1..obs_N, act_1..act_N); next actions come from every agent's target actor on its own
   * next_obs; `replay`/`actor`/`actor_target` above belong to agent `agent_index`; all replays share the sampled indices. */
  int n_agents, agent_index;
  frl_replay_t ma_replay[FRL_MAX_AGENTS];
  frl_net_t ma_actor_target[FRL_MAX_AGENTS];
  int defer_polyak;         /* 1: do not touch the targets (MADDPG updates all targets after the agent loop) */
  float* xchg;              /* dev scratch [3*B]: target-Q exchange between the per-head CTA roles */
  /* ---- Batch_ObsNorm (SAC.py:390-421, DDPG.py:372-403, MADDPG.py:366-397): running statistics over BATCH MEANS.
   * obs_norm[j] = dev [3][obs_dim_j] floats {mean, S, std} of agent j (single-agent: j = 0); NULL = trick off.
   * Every learn first folds mean_rows(obs) into the state (n = obs_norm_n0 + u + 1; n == 1 sets mean = std = x_bar like
   * the reference), then feeds (x - mean) / (std + 1e-8) for obs AND next_obs to every network. */
  float* obs_norm[FRL_MAX_AGENTS];
  int64_t obs_norm_n0;      /* updates folded in before this call */
  const float* ma_noise_next[FRL_MAX_AGENTS];   /* n_agents > 1 with target_smoothing (MATD3_simple.py:203-205): dev [B][act_dim_j]
                                                 * randn of agent j's target action for THIS agent's sample; NULL -> Philox(seed) */
  /* ---- small-batch schedule (csrc/algo_acfx.cuh; single agent, hidden 128-128, B <= 256): taken when both are given ----
   * ws   = dev scratch of frl_ac_ws_floats(args) floats (activation / gradient exchange blocks, partial sums);
   * sync = dev scratch of 4096 uint32 (grid-barrier counter + hand-off packets; the library zeroes it at every launch). */
  float* ws;
  unsigned* sync;
} frl_ac_args_t;

/* Discrete-action SAC, the `hands_on` variant of SAC_file/SAC_add_discrete.py:137-177, 299-360: softmax actor (obs -> n_actions), twin
 * critic heads obs -> Q(s, .) in ONE 6-layer net (head h = layers 3h .. 3h+2), replay rows with a 1-column action index. */
typedef struct {
  frl_net_t actor, actor_target, critic, critic_target;
  frl_replay_t replay;
  const int64_t* indices;   /* dev [n_updates][B] */
  int B, n_updates;
  float gamma, tau;
  double lr_actor, lr_critic, beta1, beta2, eps;
  float max_norm;           /* clip_grad_norm_ max (0.5 upstream) */
  int64_t step_actor0, step_critic0;
  float* alpha_state;       /* dev {log_alpha, exp_avg, exp_avg_sq, unused} */
  int adaptive_alpha;
  double alpha_lr;
  float target_entropy;     /* 0.6 * log(n_actions)  (SAC_add_discrete.py:214) */
  int64_t step_alpha0;
  float* gpart;             /* dev scratch [sm_count][max(actor.n_p, critic.n_p)] */
  float* sumsq;             /* dev scratch [sm_count] */
  float* stats;             /* dev scratch [sm_count][8] */
  float* out;               /* dev [n_updates][8]: critic_loss, actor_loss, alpha, alpha_loss, critic_gnorm, actor_gnorm, mean_entropy, 0 */
} frl_sacd_args_t;
int frl_sacd_learn(const frl_sacd_args_t* args, void* cuda_stream);

/* FRL_INFER_ARGMAX_DUELING (7): argmax_a of V + A_a - mean(A) for a [V | A] head (Dueling.forward, DQN_with_tricks.py:75-79) */
enum { FRL_INFER_ARGMAX_DUELING = 7 };
enum { FRL_INFER_ARGMAX = 0, FRL_INFER_TANH = 1, FRL_INFER_SAC_SAMPLE = 2, FRL_INFER_SAC_MEAN = 3, FRL_INFER_RAW = 4,
       FRL_INFER_PPO_GAUSS = 5,   /* out = [action(act) | log_prob per dim(act)], noise = N(0,1) [n][act] */
       FRL_INFER_PPO_CAT = 6 };   /* out = [action index | log_prob], noise = Exp(1) [n][n_actions] (torch.multinomial trick) */

typedef struct {
  frl_net_t net;
  int l0, nl;               /* layers [l0, l0+nl) of `net` form the MLP to evaluate (nl = 0: all layers) */
  const float* obs;         /* dev [n][obs_dim] */
  int n, obs_dim;
  int mode;                 /* FRL_INFER_* */
  const float* noise;       /* dev [n][out] for SAC_SAMPLE (NULL -> Philox(seed, counter)) */
  uint64_t seed;
  uint32_t counter;
  float* out;               /* dev [n][out_cols]: actions (ARGMAX: 1 column holding the index as float; RAW: net output) */
  int out_cols;
  int layer_norm;           /* 1: MAPPO nets, F.layer_norm on the input and after each hidden ReLU; 2: hidden only; 3: input only (MAPPO_discrete.py's acting nets) */
  const float* obs_norm;    /* dev [3][obs_dim] {mean, S, std} Batch_ObsNorm state applied with update=False, or NULL */
  int hidden_tanh;          /* 1: tanh hidden activations (the `tanh` trick of PPO_file/PPO_with_tricks.py:95,172); 0: ReLU */
} frl_infer_args_t;

/* Data-parallel gradient exchange over peer memory (one process per GPU, NVLink): rank r owns a device block
 *   [ flags: FRL_DP_MAX_RANKS x uint32 (padded to 256 B) | g: 2 x n_p floats ]   allocated by frl_dp_alloc and opened by every peer with
 * frl_dp_open (CUDA IPC).  After the in-kernel cross-CTA reduction of update number e (= epoch0 + u + 1) a rank writes its gradient to
 * g[e & 1], publishes e into slot `rank` of every peer's flags, waits until its own flags hold e for every rank and sums the world's
 * gradients in rank order (bit-identical on every rank) — inside the one persistent launch, no NCCL call per step. */
#define FRL_DP_MAX_RANKS 8
typedef struct {
  float* g[FRL_DP_MAX_RANKS];           /* dev: rank r's g block (peer-mapped for r != rank) */
  unsigned* flags[FRL_DP_MAX_RANKS];    /* dev: rank r's flag block */
  int rank, world;                      /* world <= 1: exchange disabled */
  unsigned epoch0;                      /* exchanges completed before this launch (identical on all ranks) */
} frl_dp_peers_t;

/* Parameter average of the off-policy replicas over the same peer blocks (csrc/replica_avg.cuh): tensor[i] (n[i] floats, this rank's
 * parameter blocks) <- (sum over the ranks, in rank order) / world, one cooperative launch per sync instead of an all-reduce + divide
 * per tensor.  The exchange block of every rank holds 2 x block_floats floats behind its flags (block_floats >= sum of n[i] rounded up
 * to 4 each).  status (dev [1] or NULL) is set to -1 if a peer did not arrive within 30 s. */
#define FRL_RA_MAX_TENSORS 8
typedef struct {
  frl_dp_peers_t dp;                    /* epoch0 = syncs completed before this one (identical on all ranks) */
  float* tensor[FRL_RA_MAX_TENSORS];
  int n[FRL_RA_MAX_TENSORS];
  int n_tensors;
  long long block_floats;
  float* status;
} frl_replica_avg_args_t;

/* On-policy (PPO.py) minibatch update.  `net` holds actor (layers 0-2, + log_std extra when continuous) and critic
 * (layers 3-5) in ONE parameter block with ONE optimiser, like the reference's merged `ac_optimizer`. */
enum { FRL_OPT_CAUTIOUS_ADAMW = 0, FRL_OPT_ADAM = 1 };
typedef struct {
  frl_net_t net;
  int continuous;           /* 1: Normal(tanh(mean), exp(clamp(log_std))), 0: Categorical(logits), 2: Beta(softplus + 1) with the actor's last layer
                             *    = [alpha logits (A) | beta logits (A)] (Actor_Beta, PPO_with_tricks.py:120-150; FFMA tile kernels only) */
  const float* obs;         /* dev [M][obs_dim] rollout arrays (Buffer_for_PPO.all()) */
  const float* action;      /* dev [M][act_cols]  (discrete: 1 column holding the index) */
  const float* logp_old;    /* dev [M][logp_cols] */
  const float* adv;         /* dev [M][n_adv] */
  const float* v_target;    /* dev [M][n_adv] */
  int M, obs_dim, act_cols, logp_cols, n_adv;
  const int64_t* indices;   /* dev [n_updates][mb]: row indices of every minibatch (np.random.permutation slices) */
  const int* mb_rows;       /* dev [n_updates]: valid rows of each minibatch (<= mb) */
  int mb, n_updates;
  float clip_param, entropy_coef;
  float max_norm_actor, max_norm_critic;   /* <=0: no clipping */
  int optimizer;            /* FRL_OPT_* */
  double lr, beta1, beta2, eps;
  int64_t step0;
  /* ---- MAPPO options (MAPPO_file/MAPPO.py:127-218, 357-482) ---- */
  int layer_norm;           /* F.layer_norm (no affine, eps 1e-5) on the input and after each hidden ReLU, actor and critic (per row; group mode below ignores it) */
  const float* critic_obs;  /* dev [M][critic_obs_dim] joint observation for the centralised critic (NULL: critic sees `obs`) */
  int critic_obs_dim;
  int value_loss;           /* 0: mse(v_target, V);  1: huber(v_target - V, huber_delta).mean()  (MAPPO.py:273-276,426-433);  2: clipped, see v_old */
  float huber_delta;
  /* ---- data-parallel split (one process per GPU): stages are 0 fwd/bwd, 1 cross-CTA reduce -> net.g, 2 grad scale + norms,
   * 3/4 optimiser.  A DP step launches [0,2), all-reduces net.g over NCCL, then launches [2,5) with grad_scale = 1/world. */
  int stage_lo, stage_hi;   /* run stages in [stage_lo, stage_hi); 0,0 = all */
  float grad_scale;         /* multiplies the reduced gradient before clipping (0 = 1.0) */
  float* gpart;             /* dev scratch [sm_count][net.n_p] */
  float* sumsq;             /* dev scratch [sm_count][2] */
  float* segcnt;            /* dev scratch [sm_count][2*FRL_MAX_LAYERS+1] cautious-mask counts per tensor */
  float* stats;             /* dev scratch [sm_count][8] */
  float* out;               /* dev [n_updates][8]: actor_loss, critic_loss, entropy, actor_gnorm, critic_gnorm */
  double lr_critic;         /* FRL_OPT_ADAM only: learning rate of the critic layers (separate actor / critic Adams of
                             * PPO_advance/PPO.py:118-119); 0 = use `lr` for both (merged optimiser) */
  int hidden_tanh;          /* bit 0: actor, bit 1: critic use tanh hidden activations (PPO_with_tricks.py:95,172; not with layer_norm) */
  /* ---- tensor-core path (csrc/algo_ppo_umma.cuh): taken for minibatches of >= 1024 rows over in->128->128->out networks when
   * umma_ws = dev scratch of frl_ppo_umma_ws_floats() floats (512-B aligned; split weights + per-CTA activation scratch) is given */
  float* umma_ws;
  frl_dp_peers_t dp;        /* in-kernel data-parallel gradient exchange (world > 1); then grad_scale should be 1 / world */
  /* ---- MAPPO_discrete options (MAPPO_file/MAPPO_discrete.py:188-192, 336-371): shared actor / critic under ONE Adam ---- */
  float max_norm_joint;     /* > 0: clip_grad_norm_ over actor AND critic gradients together (update_ac, :191); replaces the two above */
  int opt_repeat;           /* 2: the optimiser steps twice on the same clipped gradient (update_ac's step() then :371's second step());
                             *    the step counter advances by opt_repeat per update (FRL_OPT_ADAM only); 0 / 1: once */
  const float* v_old;       /* dev [M][n_adv] rollout values for value_loss 2: max((clamp(V - v_old, +-clip_param) + v_old - v_target)^2,
                             *    (V - v_target)^2) element-wise (the ValueClip branch without huber_loss, :350-357) */
  /* ---- group mode (csrc/algo_ppo_group.cuh; Categorical actor, FFMA tiles): MAPPO_discrete.py's networks normalise
   * F.layer_norm(x, x.size()[1:]) of a 4-D [minibatch, T, N, features] tensor inside learn (:81-87,114-120,141-149), i.e. jointly over
   * the T*N rows of an EPISODE.  group_rows = T*N rows that form one group; the rows of a group are consecutive in `indices` and every
   * minibatch holds whole groups (mb and mb_rows multiples of group_rows).  A CTA owns whole groups and recomputes the forward pass once
   * per statistic it needs (5 sweeps over the group's 8-row tiles), so no activation of more than one tile is ever kept. */
  int group_rows;
  int group_norm;           /* bit 0: layer_norm after both hidden ReLUs (actor and critic); bit 1: the critic's input (feature_norm; the
                             *    reference's Actor_discrete overwrites its normalised input, :113-115, so the actor has none) */
  /* value_loss 3 (group mode, n_updates = 1): ValueClip + huber_loss, critic loss = max(a, b)^2 with the batch-mean SCALARS
   * a = mean huber(e_clip), b = mean huber(e_orig) (:350-357).  Two launches per update: group_prepass = 1 with stage_lo, stage_hi =
   * 0, 1 leaves each CTA's sums of huber(e_clip) / huber(e_orig) in stats[cta][5..6]; the update launch (group_prepass = 0) folds them. */
  int group_prepass;
} frl_ppo_args_t;

/* Rainbow (DQN_with_tricks.py): Categorical + Dueling + Noisy net.  The trainable block holds the torch tensors;
 * eff[0..2] are noise-applied effective linear nets for the three forwards of a learn (online(s'), target(s'),
 * online(s)); layer 0 = l1, layer 1 = V head, layers 2.. = column blocks (<=128 outputs) of the A head.
 * map[li] tells where layer li's mu/sigma tensors live in the trainable block and which noise entries it uses. */
typedef struct {
  int mu_w, sg_w, mu_b, sg_b;   /* float offsets in the trainable block (sg_* = -1: plain nn.Linear) */
  int row0;                     /* first output row of this block inside its torch tensor */
  int eps_in, eps_out;          /* offsets of f(eps_in) / f(eps_out) inside one forward's noise vector */
} frl_noisy_map_t;

typedef struct {
  float* p; float* m; float* v;      /* dev [n_train] online parameters + Adam state */
  float* p_target;                   /* dev [n_train] target parameters */
  int n_train;
  frl_net_t eff[3];
  frl_noisy_map_t map[FRL_MAX_LAYERS];
  const float* eps;                  /* dev [3][eps_len]: transformed noise f(x)=sign(x)sqrt|x| per forward */
  int eps_len;
  int n_actions, n_atoms;
  const float* z;                    /* dev [n_atoms] support (torch.linspace(v_min, v_max, n_atoms)) */
  float v_min, v_max, delta_z;
  int double_q;
  frl_replay_t replay;
  const int64_t* indices;            /* dev [B] */
  const float* is_weight;            /* dev [B] PER importance weights or NULL */
  int B;
  float gamma, tau;                  /* gamma = n_step_gamma when N_Step */
  double lr, beta1, beta2, eps_adam;
  int64_t step0;
  float* gpart;                      /* dev scratch [sm_count][eff[2].n_p] */
  float* stats;                      /* dev scratch [sm_count][8] */
  float* error_out;                  /* dev [B]: (m * log p).sum(1) per row (PER priorities) or NULL */
  float* out;                        /* dev [8]: out[0] = loss */
  /* fast mode: when noise_gen != 0 the kernel draws the factorised noise itself — eps[f][i] = sign(x) sqrt|x| with
   * x = Philox-normal(noise_seed, stream f, noise_counter, i) — for the three forwards of this learn and writes it to `eps`
   * (the NoisyLinear epsilon bookkeeping reads it back); otherwise `eps` is an input (reference order, torch CPU generator). */
  int noise_gen;
  uint64_t noise_seed, noise_counter;
} frl_rainbow_args_t;

/* Exploration noise of the reference train loops for N vectorised envs (SAC_file/SAC.py:334-355 OUNoise; DDPG_file/DDPG.py:519-522):
 *   kind 0: OU  dx = theta (mu - x) + sqrt(dt) sigma z; x += dx; action_ = clip(action*max_action + x*scale*max_action, +-max_action)
 *   kind 1: Gaussian  action_ = clip(action*max_action + gauss_scale * (gauss_sigma*max_action*z), +-max_action)
 * z: dev [N][A] float64 standard normals drawn by the caller in the reference's order, or NULL -> Philox(seed, counter). */
typedef struct {
  int kind, N, A;
  const float* action;               /* dev [N][A] policy output in (-1, 1) */
  double* ou_state;                  /* dev [N][A] OU state (kind 0), updated in place */
  const double* z;
  uint64_t seed, counter;
  double mu, theta, sigma, dt, scale;/* OU; scale < 0 = the reference's scale=None */
  double gauss_scale, gauss_sigma;
  double max_action;
  int clip;                          /* 0: no clip (OUNoise.noise() alone: action = 0, max_action = 1) */
  double* out64;                     /* dev [N][A] float64 env action (numpy's result dtype) or NULL */
  float* out;                        /* dev [N][A] fp32 copy or NULL */
} frl_explore_args_t;

const char* frl_last_error(void);
int frl_is_emulation(void);          /* 0 for the CUDA library (the only one the product path accepts) */
int frl_device_sm_count(void);
/* audit of fast mode: out[i] = the standard normal the learn kernels draw for (seed, stream, counter, element i) */
int frl_debug_randn(uint64_t seed, uint32_t stream, uint32_t counter, long long n, float* out, void* cuda_stream);
long long frl_launch_count(void);    /* kernels launched by the library since load (bench.py's gpu_launches) */
/* peer-memory blocks of frl_dp_peers_t: cudaMalloc + zero + IPC handle (64 bytes) / open a peer's handle / release */
int frl_dp_alloc(long long bytes, void** dev_ptr, unsigned char* ipc_handle_64);
int frl_dp_open(const unsigned char* ipc_handle_64, void** dev_ptr);
int frl_dp_close(void* peer_ptr);
int frl_dp_free(void* dev_ptr);
int frl_replica_average(const frl_replica_avg_args_t* args, void* stream);
long long frl_ppo_umma_ws_floats(void);   /* size of frl_ppo_args_t.umma_ws (0 from the test-only emulation) */
int frl_wt_ld(int out_pad);          /* row stride (floats) of a transposed-mirror layer image with this padded width */
int frl_abi_version(void);
/* sizeof the argument structs as compiled: 0 frl_layer_t, 1 frl_net_t, 2 frl_replay_t, 3 frl_dqn_args_t, 4 frl_ac_args_t,
 * 5 frl_infer_args_t, 6 frl_ppo_args_t, 7 frl_noisy_map_t, 8 frl_rainbow_args_t, 9 frl_explore_args_t, 10 frl_sacd_args_t,
 * 11 frl_replica_avg_args_t; -1 otherwise.  A binding checks its mirror
 * of the layout against these before the first call (freerl_b200/_lib.py does, at load time). */
int frl_struct_size(int which);

int frl_replay_add_batch(const frl_replay_t* rb, int64_t index, const float* obs, const float* act, const float* rew,
                         const float* next_obs, const float* done, int n, void* stream);
int frl_replay_gather(const frl_replay_t* rb, const int64_t* indices, int B, float* obs, float* act, float* rew,
                      float* next_obs, float* done, void* stream);
int frl_sample_uniform(int64_t* indices_out, int64_t size, int B, int n_updates, uint64_t seed, uint64_t counter,
                       void* stream);
int frl_net_sync_mirror(const frl_net_t* net, void* stream);
/* theta' <- theta'(1-tau) + theta*tau on parameter block + mirror (update_target, e.g. MADDPG.py:230-237) */
int frl_polyak(const frl_net_t* src, const frl_net_t* target, float tau, void* stream);
int frl_dqn_learn(const frl_dqn_args_t* args, void* stream);
int frl_ac_learn(const frl_ac_args_t* args, void* stream);
/* floats of `ws` the small-batch schedule needs for these shapes (0: not eligible, frl_ac_learn takes the generic kernel);
 * frl_ac_path: 1 if frl_ac_learn(args) would run the small-batch schedule, 0 for the generic kernel */
long long frl_ac_ws_floats(const frl_ac_args_t* args);
int frl_ac_path(const frl_ac_args_t* args);
int frl_policy_infer(const frl_infer_args_t* args, void* stream);
/* GAE over a [T][N] rollout (N env columns): td = r + gamma(1-done)v' - v (fp32); A_t = td_t + gamma*lmbda*(1-adv_done_t)A_{t+1}
 * accumulated in float64 per column from a zero tail; adv (fp32) and v_target = adv + v. */
int frl_gae(const float* reward, const float* done, const float* adv_done, const float* vs, const float* vs_next, int T, int N,
            double gamma, double lmbda, float* adv_out, float* v_target_out, void* stream);
int frl_ppo_update(const frl_ppo_args_t* args, void* stream);
/* joint advantage normalisation (MAPPO.py:385-386): out = (x - mean(x)) / (std_unbiased(x) + eps) over n floats (in place ok) */
int frl_adv_norm(const float* x, int n, float eps, float* out, void* stream);

/* Prioritized replay (DQN_file/Buffer.py:66-194): `tree` is the reference's float64 array heap [2*cap-1] on the device.
 * frl_sumtree_update applies B leaf writes IN ORDER (ancestors get `+= change` in batch order => bit-exact with the
 * sequential reference); targets are idx[i] or, with idx_is_range, (idx0+i) % cap; the new priority of item i is
 * pri32[i] | *pri64_scalar | pri_const; B <= 1024 per call, capacity < 2^30; one cooperative launch, one CTA per tree level.  frl_sumtree_sample = PER_Buffer.sample (stratified descent on u[i] in [0,1),
 * float32 priority container, float64 importance weights cast to fp32).  frl_sumtree_max = np.max(tree[-cap:]). */
int frl_sumtree_update(double* tree, int64_t cap, const int64_t* idx, const float* pri32, const double* pri64_scalar,
                       double pri_const, int64_t idx0, int idx_is_range, int B, double* scratch /* dev, >= 2048 doubles */,
                       void* stream);
/* PER_Buffer.update_priorities (DQN_file/Buffer.py:126-129) in one launch: p_i = (|td_i| + eps)^alpha in fp32, then the ordered update */
int frl_sumtree_update_td(double* tree, int64_t cap, const int64_t* idx, const float* td, float eps, float alpha, int B,
                          double* scratch /* dev, >= 2048 doubles */, void* stream);
int frl_sumtree_sample(const double* tree, int64_t cap, const double* u, uint64_t seed, uint64_t counter, int B, int64_t size,
                       double beta, double prob_floor, int64_t* out_idx, float* out_pri, float* out_w, void* stream);
int frl_sumtree_max(const double* tree, int64_t cap, double* scratch, int nscratch, double* out, void* stream);
int frl_per_priorities(const float* td, int B, float eps, float alpha, float* out, void* stream);
int frl_rainbow_learn(const frl_rainbow_args_t* args, void* stream);
/* select_action: refresh eff[0] from (p, eps[0]) and return argmax_a sum_z z*p(z|s,a) for n observations (out: [n] floats) */
int frl_rainbow_act(const frl_rainbow_args_t* args, const float* obs, int n, float* out, void* stream);


/* ---- per-step work of the reference train loops over N vectorised envs ----
 * frl_vecnorm = Normalization.__call__ (PPO_file/normalization.py:17-49): `state` = dev [3][D] float64 {mean, S, std}, n0 = updates
 * folded so far (the caller adds N after an update call).  update != 0 folds rows 0..N-1 IN ORDER, each row normalised with the
 * statistics that include it (bit-identical to calling the reference object once per row, float32 or float64 rows);
 * update == 0 applies frozen statistics.  out (fp32) / out64 (float64): [N][D], either may be NULL. */
int frl_vecnorm(double* state, int64_t n0, const void* x, int x_is_f64, int N, int D, int update, float* out, double* out64,
                void* stream);
/* RewardScaling.__call__ (normalization.py:87-97) for N envs: R[i] = gamma R[i] + x[i]; the shared running std is fed R[0..N-1]
 * in order; out[i] = x[i] / (std + 1e-8).  state = dev [3] float64 {mean, S, std}. */
int frl_reward_scaling(double* state, int64_t n0, double* R, const void* x, int x_is_f64, double gamma, int N, float* out,
                       double* out64, void* stream);
int frl_explore(const frl_explore_args_t* args, void* stream);
/* state[i][:] = value where mask[i] != 0 (OUNoise.reset / RewardScaling.reset of the envs whose episode ended); state dev [N][W] float64 */
int frl_masked_reset(double* state, const uint8_t* mask, int N, int W, double value, void* stream);
/* epsilon-greedy of the DQN mains (DQN_file/DQN.py:307-310) for N envs: out[i] = u[i] < epsilon ? rnd[i] : greedy[i]; u == NULL draws
 * u and the random action from Philox(seed, counter, i). */
int frl_epsilon_greedy(const int64_t* greedy, int N, int n_actions, double epsilon, const double* u, const int64_t* rnd,
                       uint64_t seed, uint64_t counter, int64_t* out, void* stream);
/* dis_to_con (DQN_file/DQN.py:195-217): N discrete actions -> [N][shape] continuous actions between float32 bounds low/high (dev [shape]);
 * per = int(n_actions ** (1 / shape)) as the reference computes it (unused for shape == 1). */
int frl_dis_to_con(const int64_t* action, int N, int n_actions, int shape, int per, const float* low, const float* high,
                   double* out64, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
