/*
 * include/mgrl.h — C ABI of the B200-native batched MiniGrid simulator + PPO rollout kernels.
 *
 * The reference (Idokorro/MiniGrid-RL) is pure Python and has no FFI: the seam this library
 * sits behind is the SB3 VecEnv protocol that /root/reference/src/ppo.py:118-126 constructs
 * (make_vec_env -> VecTransposeImage -> VecFrameStack(4,'first')) and ppo.py:159 / :210 /
 * :242 drive (reset / step).  Each entry point below names the reference interface it
 * replaces.  INTEGRATION.md shows the ctypes binding and the VecEnv class that a
 * maintainer of the reference would drop into ppo.py.
 *
 * Conventions
 *  - every function returns 0 (MGRL_OK) or a negative mgrl_status; mgrl_last_error() holds
 *    the message of the last failure on the calling thread.  No exception crosses the ABI.
 *  - "_dev" pointers are device pointers owned by the caller (e.g. torch.Tensor.data_ptr());
 *    "_host" pointers are host memory (pinned memory from mgrl_host_alloc is fastest).
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *    stream) unless documented otherwise.  A handle is bound to one device and is not
 *    thread-safe; distinct handles are independent (one per GPU / rank).
 *  - there is NO CPU implementation behind this ABI: without a CUDA device mgrl_create fails.
 *
 * State layout (140 bytes per environment, also the unit of get/set_state):
 *   grid[121]  one "kind" byte per cell, grid[y*size+x]
 *       0 empty (1,0,0)  1 wall (2,5,0)  2 goal (8,1,0)  3 lava (9,0,0)
 *       8+c key  16+c ball  24+8*s+c door (s: 0 open 1 closed 2 locked)
 *       64+8*m+c box (m: 0 empty, 1..6 holds a key of colour m-1)
 *       colours c: red 0 green 1 blue 2 purple 3 yellow 4 grey 5
 *   agent_x, agent_y, agent_dir, carrying(kind, 0 = nothing), step_count,
 *   target_x, target_y (0xFF = none), target_action (0 = none), mission_id,
 *   mission_done, latch_step, episode(u32), reset_draws(u16), error(u8), pad(u8)
 * Mission ids: group*24 + type*6 + colour, type: key 0 ball 1 box 2 door 3;
 *   group 0 "go to", 1 "toggle", 2 "pick up"; 72 "go to goal"; 73 "drop".
 */
#ifndef MGRL_H
#define MGRL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGRL_ABI_VERSION 1
#define MGRL_STATE_BYTES 140
#define MGRL_OBS_BYTES 147       /* 7*7*3 */
#define MGRL_N_MISSIONS 74
#define MGRL_MISSION_TOKENS 32   /* environment.py:82 msn_len */
#define MGRL_FRAMES 4            /* ppo.yaml:5 n_frames_stack */

typedef enum {
    MGRL_OK = 0,
    MGRL_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
    MGRL_ERR_CUDA = -2,      /* CUDA runtime error (message in mgrl_last_error) */
    MGRL_ERR_NO_DEVICE = -3, /* no CUDA device: this library has no CPU path */
    MGRL_ERR_ENV = -4        /* an environment flagged an error (invalid action, rejection cap) */
} mgrl_status;

/* image record layouts: HWC = image[vx][vy][c] (MiniGrid native), CHW = image[c][vx][vy] (after SB3
 * VecTransposeImage), both 147 bytes per environment; HWC148 = HWC with one zero pad byte, i.e. a
 * 148-byte record pitch, which lets the kernel emit aligned 32-bit stores (fastest). */
typedef enum { MGRL_OBS_HWC = 0, MGRL_OBS_CHW = 1, MGRL_OBS_HWC148 = 2 } mgrl_obs_layout;
/* cfg.env.problem, custom_env.py:134-152.  Mission ids: group * 24 + type * 6 + colour (group 0 'go to', 1 'toggle',
 * 2 'pick up'; type key 0, ball 1, box 2, door 3), 72 'go to goal', 73 'drop'; 'move left|right|up|down' (problems mov / full,
 * custom_env.py:216-256) take ids 24..27 - 'toggle <colour> key' is never generated - so the table keeps MGRL_N_MISSIONS rows.
 * A 'move' episode keeps its target_range in the state bytes target_x | target_y << 8 | target_action << 16 | pad << 24 as
 * decimal digits (digit k = coordinate of the cell in row / column k+1, 0 = none). */
typedef enum { MGRL_MULTI = 0, MGRL_GTO = 1, MGRL_GTG = 2, MGRL_OPN = 3, MGRL_PKP = 4, MGRL_DRP = 5, MGRL_MOV = 6,
               MGRL_FULL = 7 } mgrl_problem;

/* mirrors cfg.env of hydra_configs/single.yaml:20-28 (+ batch placement) */
typedef struct {
    int32_t size;              /* env.size, 5..11 */
    int32_t num_objects;       /* env.num_objects */
    int32_t problem;           /* mgrl_problem */
    int32_t mission;           /* env.mission: 0 go-to, 1 toggle, 2 pick-up, 5 go-to-goal, -1 = null (mixed) */
    int32_t all_doors_open;    /* env.all_doors_open */
    int32_t see_through_walls; /* env.see_through_walls */
    int32_t max_steps;         /* size*size (custom_env.py:114); 0 = derive */
    int32_t num_obstacles;     /* floor((size-2)^2 * percent_obstacles) if env.obstacles else 0 */
    int32_t num_envs;          /* environments held by this handle (the local shard) */
    int32_t obs_layout;        /* mgrl_obs_layout of every image this handle writes */
    uint64_t env_id_base;      /* global id of local env 0 (rank * num_envs): RNG key, so results
                                  do not depend on how envs are sharded across GPUs */
} mgrl_config;

typedef struct mgrl_env mgrl_env;

int mgrl_abi_version(void);
const char *mgrl_last_error(void);

/* replaces make_vec_env(make_env, n_envs, ...) (ppo.py:118-122): allocates the device state */
int mgrl_create(const mgrl_config *cfg, int device, mgrl_env **out);
int mgrl_destroy(mgrl_env *env);

/* pinned host memory for the *_host entry points */
int mgrl_host_alloc(void **ptr, size_t bytes);
int mgrl_host_free(void *ptr);

/* ---- device-resident fast path ------------------------------------------------------ */

/* replaces VecEnv.seed(seed) + VecEnv.reset() (ppo.py:134,210): fresh environments
 * (latch cleared, episode 0) from RNG keys (seed, env_id_base+i); writes the first
 * observation.  image_dev [N,pitch] in obs_layout (pitch 147 or 148), dir_dev [N], mission_dev [N]
 * (mission id). */
int mgrl_reset(mgrl_env *env, uint64_t seed, uint8_t *image_dev, uint8_t *dir_dev,
               uint8_t *mission_dev, void *stream);

/* replaces VecEnv.step(actions) for the un-stacked observation (SubprocVecEnv/DummyVecEnv
 * step_wait over Monitor(Discrete2BoxWrapper(TokenizeVocabWrapper(PlaygroundEnv)))):
 * PlaygroundEnv.step (custom_env.py:269-330) + auto-reset on done.
 *   actions_dev [N] u8 in 0..6;  image_dev [N,pitch];  dir_dev [N];  mission_dev [N];
 *   reward_dev [N] f32;  term_dev / trunc_dev [N] u8 (done = term|trunc,
 *   info['TimeLimit.truncated'] = trunc & !term);
 *   ep_len_dev [N] u8 or NULL: Monitor's episode length on done steps, else 0
 *     (Monitor's episode return equals the reward of the done step: all others are 0);
 *   term_image_dev [N,pitch] / term_dir_dev [N] or NULL: image and direction of
 *     info['terminal_observation'], written only for environments that finished on this
 *     step (its mission is the previous step's mission). */
int mgrl_step(mgrl_env *env, const uint8_t *actions_dev, uint8_t *image_dev, uint8_t *dir_dev,
              uint8_t *mission_dev, float *reward_dev, uint8_t *term_dev, uint8_t *trunc_dev,
              uint8_t *ep_len_dev, uint8_t *term_image_dev, uint8_t *term_dir_dev, void *stream);

/* T consecutive steps in ONE launch with the state tile kept in shared memory between
 * steps (action-trace replay; BASELINE.json config 4 and the kernel benchmark).
 * actions_dev [T,N]; outputs are [T,N,...] with the same meaning as mgrl_step; any output
 * except reward/term/trunc may be NULL. */
int mgrl_step_many(mgrl_env *env, int T, const uint8_t *actions_dev, uint8_t *image_dev,
                   uint8_t *dir_dev, uint8_t *mission_dev, float *reward_dev, uint8_t *term_dev,
                   uint8_t *trunc_dev, uint8_t *ep_len_dev, void *stream);

/* canonical state dump / restore, bytes must equal num_envs*MGRL_STATE_BYTES.
 * set_state also (re)sets the seed used by later auto-resets. */
int mgrl_get_state(mgrl_env *env, void *dst_dev, size_t bytes, void *stream);
int mgrl_set_state(mgrl_env *env, const void *src_dev, size_t bytes, uint64_t seed, void *stream);
/* same, host buffers; synchronous */
int mgrl_get_state_host(mgrl_env *env, void *dst_host, size_t bytes, void *stream);
int mgrl_set_state_host(mgrl_env *env, const void *src_host, size_t bytes, uint64_t seed, void *stream);
/* raw device pointer to the [N,140] state array (zero-copy inspection) */
int mgrl_state_ptr(mgrl_env *env, void **state_dev);

/* current observation of every environment without stepping (gen_obs) */
int mgrl_observe(mgrl_env *env, uint8_t *image_dev, uint8_t *dir_dev, uint8_t *mission_dev, void *stream);
/* replaces minigrid FullyObsWrapper (experts_test.py:29): image [N,size,size,3], agent = (10,0,dir) */
int mgrl_full_obs(mgrl_env *env, uint8_t *image_dev, void *stream);
/* the same into a host buffer image_host [N,S,S,3] (device-to-host copy inside, synchronous): what
 * FullyObsWrapper.observation returns per environment at experts_test.py:27-30,46 */
int mgrl_full_obs_host(mgrl_env *env, uint8_t *image_host, void *stream);
/* max over environments of the per-env error byte (synchronises the stream) */
int mgrl_error_flags(mgrl_env *env, int *flags_out, void *stream);

/* replaces VecFrameStack(4,'first').step_wait on all three keys (ppo.py:126):
 * shift by one frame, zero the history of finished envs, append the newest frame.
 *   stack_image_dev [N,4,147] (== (N,12,7,7) for CHW frames), stack_dir_dev [N,16] one-hot
 *   (Discrete2BoxWrapper, environment.py:144-149), stack_mission_dev [N,128] int64 tokens
 *   (TokenizeVocabWrapper, environment.py:91-112) from token_table_dev [74,32] int64.
 *   done_dev NULL = reset (clear all history). */
int mgrl_stack_push(int num_envs, const uint8_t *image_dev, const uint8_t *dir_dev,
                    const uint8_t *mission_dev, const uint8_t *done_dev, const int64_t *token_table_dev,
                    uint8_t *stack_image_dev, uint8_t *stack_dir_dev, int64_t *stack_mission_dev,
                    void *stream);

/* replaces SB3 RolloutBuffer.compute_returns_and_advantage (float32, same operation order):
 * rewards/values/advantages/returns [T,N] f32, episode_starts [T,N] u8, last_values [N],
 * last_dones [N] u8. */
int mgrl_gae(const float *rewards_dev, const float *values_dev, const uint8_t *episode_starts_dev,
             const float *last_values_dev, const uint8_t *last_dones_dev, double gamma,
             double gae_lambda, int T, int N, float *advantages_dev, float *returns_dev, void *stream);

/* ---- policy forward for the rollout (K3) ---------------------------------------------- */
#define MGRL_POLICY_WEIGHTS 84940 /* floats in the packed weight buffer, layout in csrc/mgrl_policy.cu */
#define MGRL_POLICY_DETERMINISTIC 1 /* flags: action = argmax(logits) instead of a sample (evaluate_policy, ppo.py:161) */
#define MGRL_POLICY_TENSOR 2 /* flags: tensor-core kernel (split-TF32 mma, fp32-class results); the weight buffer then
                              * holds MGRL_POLICY_WEIGHTS + MGRL_POLICY_FRAGMENTS floats, see mgrl_policy_pack_fragments */
#define MGRL_POLICY_FRAGMENTS 93696 /* floats of the per-lane mma fragment section behind the fp32 weights */

/* replaces CustomPPOPolicy.forward (policies.py:227-244, CustomExtractor policies.py:21-120) together with
 * the VecFrameStack(4,'first') / VecTransposeImage it reads through (ppo.py:124-126), for the rollout:
 * gathers the 4-frame stack of every environment from an un-stacked HWC148 frame buffer (records of time
 * index b-3..b; frames older than the episode are zero), runs extractor + MLPs + heads in fp32 and samples
 * the action by inverse CDF on one Philox uniform keyed by (seed, env_id_base + i, step), or takes the argmax
 * when flags has MGRL_POLICY_DETERMINISTIC.
 *   frames_dev [>= b+1, N, 148] u8, dirs_dev [>= b+1, N] u8, mission_dev [N] (time b);
 *   prev_age_dev / prev_done_dev [N] or NULL (NULL = first observation after a reset): age = frames of
 *   history available, 0..3, written to age_out_dev [N]; start_out_dev [N] or NULL = episode_start flag;
 *   action_dev [N] u8 / logp_dev [N] f32 (or NULL: values only), value_dev [N] f32, logits_dev [N,7] or NULL.
 * The mission GRU is the look-up table at the end of the weight buffer ([74*4][128], row = mission*4 + age). */
int mgrl_policy_forward(const float *weights_dev, const uint8_t *frames_dev, const uint8_t *dirs_dev,
                        const uint8_t *mission_dev, const uint8_t *prev_age_dev, const uint8_t *prev_done_dev,
                        uint8_t *age_out_dev, uint8_t *start_out_dev, uint8_t *action_dev, float *logp_dev,
                        float *value_dev, float *logits_dev, int num_envs, int time_index, uint64_t seed,
                        uint64_t env_id_base, uint32_t step, int flags, void *stream);
/* Fills the fragment section (floats [MGRL_POLICY_WEIGHTS, MGRL_POLICY_WEIGHTS + MGRL_POLICY_FRAGMENTS) of
 * weights_dev, 16-byte aligned) from the fp32 weights in front of it: every convolution / linear layer of
 * CustomExtractor + MlpExtractor + heads (policies.py:21-120,227-244) as tf32 hi/lo B fragments of mma.m16n8k8.
 * Call after every weight change, before mgrl_policy_forward with MGRL_POLICY_TENSOR. */
int mgrl_policy_pack_fragments(float *weights_dev, void *stream);
const char *mgrl_policy_last_error(void);

/* PPO update, first stage of the image extractor (Conv2d(12,16,2) + ReLU + MaxPool2d(2): policies.py:59 over
 * single.yaml:44-47, inside SB3's PPO.train) for the minibatch samples (t_dev[b], i_dev[b]) of a rollout, read
 * straight from the un-stacked frame buffer (frames_dev [time, N, 148] u8; the newest frame of sample b is record
 * t+3, frames older than the episode - age_dev[b] = 0..3 frames of history - are zero):
 *   forward : w1_dev [16,12,2,2] f32 (torch layout), b1_dev [16] -> pooled_dev [B,9,16] f32 (cell = qh*3+qw, after
 *             bias and ReLU) and arg_dev [B,9,16] u8 (position of the maximum | 4 if positive);
 *   backward: dpooled_dev [B,9,16] -> dw1_dev [16,48] f32, db1_dev [16] (overwritten). */
int mgrl_conv1_pool_forward(const uint8_t *frames_dev, int num_envs, const int32_t *t_dev, const int32_t *i_dev,
                            const uint8_t *age_dev, int batch, const float *w1_dev, const float *b1_dev,
                            float *pooled_dev, uint8_t *arg_dev, void *stream);
int mgrl_conv1_pool_backward(const uint8_t *frames_dev, int num_envs, const int32_t *t_dev, const int32_t *i_dev,
                             const uint8_t *age_dev, int batch, const uint8_t *arg_dev, const float *dpooled_dev,
                             float *dw1_dev, float *db1_dev, void *stream);

/* PPO update, input patches of the second convolution (Conv2d(16,32,2), policies.py:59 over single.yaml:48) and their
 * adjoint: patches_dev [B,4,64] f32 with patches[b][o][kk*16+ci] = pooled[b][q(o,kk)][ci] (o = oh*2+ow, kk = kh*2+kw,
 * q = (oh+kh)*3 + ow+kw) from pooled_dev [B,9,16]; backward sums dpatches_dev [B,4,64] into dpooled_dev [B,9,16]
 * (overwritten).  16-byte aligned buffers. */
int mgrl_patch2x2_forward(const float *pooled_dev, int batch, float *patches_dev, void *stream);
int mgrl_patch2x2_backward(const float *dpatches_dev, int batch, float *dpooled_dev, void *stream);

/* PPO update, gradient of the mission look-up table (the table replaces the GRU of CustomExtractor's mission branch,
 * policies.py:59 over single.yaml:52-56, for the (mission, age) pairs that exist): out_dev [n_rows,128] f32 =
 * sum over b of d_dev [B,128] into row rows_dev[b] (int64, 0 <= row < n_rows <= 400); out_dev is overwritten. */
int mgrl_lut_grad(const float *d_dev, const int64_t *rows_dev, int batch, int n_rows, float *out_dev, void *stream);

/* PPO update, bias gradients of the convolution / linear layers (policies.py:59, SB3 MlpExtractor): out_dev [cols] f32 =
 * column sums of the row-major g_dev [rows, cols] f32 (cols <= 128); out_dev is overwritten. */
int mgrl_colsum(const float *g_dev, long long rows, int cols, float *out_dev, void *stream);

/* ---- PPO optimizer step, hand-written end to end (K5) --------------------------------- */
/* Replaces SB3 `PPO.train` as the reference runs it (ppo.py:94-113,159 on CustomPPOPolicy, policies.py:21-120,227-257;
 * hyper-parameters hydra_configs/algorithm/ppo.yaml:28-38): for one minibatch of rollout samples, the forward of
 * extractor + MLPs + heads, the clipped-surrogate / clipped-value / entropy loss, every parameter gradient, the global-norm
 * clip and the Adam step.  All kernels are in csrc/mgrl_update.cu (+ the first extractor stage of mgrl_policy_tc.cu); no
 * library kernel runs inside these calls.
 *   Parameters, gradients and Adam moments are caller-owned flat float buffers of MGRL_PPO_PARAMS entries in the order
 *   of SB3's state dict as listed in minigrid-rl_b200/policy.py SHAPES (direction Linear, the three convolutions,
 *   Embedding, GRU, policy_net, value_net, action_net, value_net).  sequences_dev [num_sequences,128] u8 are the
 *   stacked mission token sequences (row = mission*4 + age). */
#define MGRL_PPO_PARAMS 110216
typedef struct mgrl_ppo mgrl_ppo;
typedef struct {
    const uint8_t *frames;   /* [T+4, N, 148] u8 un-stacked frame records (newest frame of sample (t, i) is record t+3) */
    const uint8_t *dirs;     /* [T+4, N] */
    const uint8_t *mission;  /* [T+4, N] mission ids */
    const uint8_t *age;      /* [T+1, N] frames of history available (0..3) */
    const uint8_t *actions;  /* [T, N] */
    const float *values;     /* [T, N] old values */
    const float *logp;       /* [T, N] old log-probabilities */
    const float *adv;        /* [T, N] advantages (mgrl_gae) */
    const float *ret;        /* [T, N] returns */
    int num_envs;            /* N */
} mgrl_rollout_view;
typedef struct {
    float clip_range, clip_range_vf /* < 0: no value clipping */, ent_coef, vf_coef;
    int normalize_advantage; /* (adv - mean) / (std + 1e-8) with the minibatch moments of adv_sums_dev */
    int strict_fp32;         /* 1: three-term split TF32 (fp32-class results); 0: one TF32 pass like ppo.py:29-32 */
    int use_tcgen05;         /* with strict_fp32 = 0: the 208 x 128 MLP GEMMs (forward and dX) as tcgen05.mma kind::tf32 with the
                              * accumulator in TMEM (csrc/mgrl_linear_tc5.cu) instead of mma.sync */
} mgrl_ppo_hyper;
int mgrl_ppo_create(int device, int max_batch, int num_sequences, mgrl_ppo **out);
int mgrl_ppo_destroy(mgrl_ppo *ctx);
int mgrl_ppo_bind(mgrl_ppo *ctx, float *params_dev, float *grads_dev, float *adam_m_dev, float *adam_v_dev,
                  const uint8_t *sequences_dev);
/* GRU(Embedding(tokens)) final hidden state of every sequence -> lut_out_dev [num_sequences,128] (policies.py:61-67,87-93) */
int mgrl_ppo_mission_table(mgrl_ppo *ctx, float *lut_out_dev, void *stream);
/* (sum, sum of squares, count) in double of adv_dev[idx_dev[..]] for every consecutive minibatch of `batch` entries of
 * idx_dev [total] -> sums_dev [ceil(total/batch)][3]: what the per-minibatch advantage normalisation needs; summed over
 * ranks by the caller (one all-reduce per epoch). */
int mgrl_ppo_moments(const float *adv_dev, const int32_t *idx_dev, int batch, long long total, double *sums_dev,
                     void *stream);
/* forward + loss + backward of the minibatch idx_dev [batch] (flat indices t*N + i): the flat gradient buffer is overwritten;
 * loss_out_dev [4] (optional) = sums over the minibatch of policy loss, value loss and -entropy terms (divide by batch);
 * logits_out_dev [batch,7] / values_out_dev [batch] optional. */
int mgrl_ppo_gradients(mgrl_ppo *ctx, const mgrl_rollout_view *rollout, const int32_t *idx_dev, int batch,
                       const double *adv_sums_dev, const mgrl_ppo_hyper *hyper, float *loss_out_dev,
                       float *logits_out_dev, float *values_out_dev, void *stream);
/* clip_grad_norm_(max_grad_norm) + Adam step number `step` (1-based) on the bound buffers; gradients are multiplied by
 * grad_scale first (1 / world size after an all-reduce).  norm_out_dev (optional) receives the gradient norm. */
int mgrl_ppo_apply(mgrl_ppo *ctx, float lr, float max_grad_norm, float grad_scale, float beta1, float beta2, float eps,
                   int step, float *norm_out_dev, void *stream);
/* device pointer of an internal activation buffer of the last mgrl_ppo_gradients call (tests): "pooled", "h2", "f", "a1",
 * "a2", "dz2", "dz1", "df", "dh2", "dpooled", "lut", "dlut", "stats" */
int mgrl_ppo_debug_buffer(mgrl_ppo *ctx, const char *name, void **out);
/* copies `count` floats of that buffer to dst_dev (device to device, stream ordered) */
int mgrl_ppo_debug_copy(mgrl_ppo *ctx, const char *name, float *dst_dev, long long count, void *stream);

/* ---- host-buffer drop-in path (what B200VecEnv.reset/step with numpy arrays calls) --- */

/* VecEnv.reset(): stacked observation dict into host buffers; synchronous.
 *   image_host [N,4,147] u8, direction_host [N,16] u8, mission_host [N,128] i64 */
int mgrl_vec_reset_host(mgrl_env *env, uint64_t seed, uint8_t *image_host, uint8_t *direction_host,
                        int64_t *mission_host, void *stream);
/* VecEnv.step(actions): actions_host [N] u8 -> stacked obs, rewards [N] f32, term/trunc [N],
 * ep_len [N], term_image_host [N,147] + term_dir_host [N] (last frame of the terminal
 * observation, rows of finished envs only; both may be NULL).  Copies H2D/D2H inside;
 * synchronous. */
int mgrl_vec_step_host(mgrl_env *env, const uint8_t *actions_host, uint8_t *image_host,
                       uint8_t *direction_host, int64_t *mission_host, float *reward_host,
                       uint8_t *term_host, uint8_t *trunc_host, uint8_t *ep_len_host,
                       uint8_t *term_image_host, uint8_t *term_dir_host, void *stream);
/* VecEnv.step(actions) with the SB3 observation dict kept IN PLACE in host memory, the way VecFrameStack keeps
 * `stacked_obs` (ppo.py:124-126): image_host / direction_host / mission_host must be the arrays that the previous
 * mgrl_vec_reset_host or mgrl_vec_step_stacked_host call of this handle filled; the call shifts every environment's
 * stack by one frame and appends the new one (a finished environment restarts from zeros).  Only one 64-byte record per
 * environment crosses PCIe (one code byte per view cell + the step's scalars); the library's host threads expand it
 * (format conversion only).  For environments that finished, row e of term_image_host [N,4,147] / term_direction_host
 * [N,16] / term_mission_host [N,128] receives the stacked terminal observation (SB3's info['terminal_observation']);
 * other rows are left alone; the three may be NULL.  Synchronous. */
int mgrl_vec_step_stacked_host(mgrl_env *env, const uint8_t *actions_host, uint8_t *image_host,
                               uint8_t *direction_host, int64_t *mission_host, float *reward_host,
                               uint8_t *term_host, uint8_t *trunc_host, uint8_t *ep_len_host,
                               uint8_t *term_image_host, uint8_t *term_direction_host,
                               int64_t *term_mission_host, void *stream);
/* The same two calls without the SB3 wrapper stack: the un-stacked observation (what PlaygroundEnv.step +
 * Discrete2BoxWrapper's input look like before VecTransposeImage / VecFrameStack / TokenizeVocabWrapper) goes to
 * host buffers: image_host [N,pitch] in the handle's obs_layout, dir_host [N], mission_host [N] (mission ids),
 * reward / term / trunc / ep_len [N], term_image_host [N,pitch] + term_dir_host [N] (rows of finished envs; may be
 * NULL).  This is the output set of the CPU oracle's mg_vec_step, i.e. the like-for-like end-to-end path. */
int mgrl_vec_reset_frames_host(mgrl_env *env, uint64_t seed, uint8_t *image_host, uint8_t *dir_host,
                               uint8_t *mission_host, void *stream);
int mgrl_vec_step_frames_host(mgrl_env *env, const uint8_t *actions_host, uint8_t *image_host, uint8_t *dir_host,
                              uint8_t *mission_host, float *reward_host, uint8_t *term_host, uint8_t *trunc_host,
                              uint8_t *ep_len_host, uint8_t *term_image_host, uint8_t *term_dir_host, void *stream);
/* hardware probe of the tcgen05 operand layout of mgrl_conv1_tc5.cu (rows at a 16-byte pitch, one plane per K chunk, taps as
 * row-shifted descriptors): out_dev [4][128][16] f32 = the products for the shifts 0, 1, 7, 8 (profiles/tc5_shift_probe.py) */
int mgrl_debug_tc5_shift_probe(float *out_dev, void *stream);

/* mission-id -> token table used by the host path and by mgrl_stack_push callers: [74,32] i64 */
int mgrl_set_token_table(mgrl_env *env, const int64_t *table_host);

#ifdef __cplusplus
}
#endif
#endif /* MGRL_H */
