/*
 * oracle/mg_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT)
 *
 * Plain-C restatement of the reference hot path
 *   /root/reference/src/custom_env.py  (PlaygroundEnv: step :269-330, _gen_grid :122-267,
 *                                       generators :371-513, :595-2034, next2door :2036-2046)
 * on top of the upstream MiniGrid semantics restated in SURVEY.md Appendix A
 * (MiniGridEnv.step / gen_obs / place_obj / place_agent, Grid, WorldObj).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call this code, and only as the checker / CPU baseline.
 *
 * PARITY PIN: the reference ships no golden vectors and `minigrid` is not installable
 * here, so the [UPSTREAM] part is "parity unpinned".  The PlaygroundEnv part IS pinned:
 * tests/golden/ holds traces recorded by oracle/gen_golden.py from the UNMODIFIED
 * reference classes running on oracle/upstream_shim, driven by this file's Philox
 * stream, and tests/test_oracle_golden.py replays them through this oracle bit-exactly.
 *
 * The struct layouts and the kind-byte encoding below are shared verbatim with the CUDA
 * product (include/mgrl.h documents the same bytes); the oracle has its own copy so that
 * it never includes product headers.
 */
#ifndef MG_ORACLE_H
#define MG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_MAX_SIZE 11
#define MG_GRID_CELLS 121
#define MG_VIEW 7
#define MG_OBS_BYTES 147 /* 7*7*3, image[vx][vy][c] */
#define MG_NONE 0xFF

/* ---- kind byte: one byte per grid cell / carried object ------------------------------
 *  0 empty (1,0,0)   1 wall (2,5,0)   2 goal (8,1,0)   3 lava (9,0,0)
 *  8+c   key  colour c            (5,c,0)
 *  16+c  ball colour c            (6,c,0)
 *  24+8*s+c door colour c state s (4,c,s)  s: 0 open, 1 closed, 2 locked
 *  64+8*m+c box colour c          (7,c,0)  m: 0 empty box, 1..6 holds Key(colour m-1)
 *  colours: red 0 green 1 blue 2 purple 3 yellow 4 grey 5  (upstream COLOR_TO_IDX)
 * As `carrying`, 0 means "nothing" (renders as empty, like upstream gen_obs_grid). */
enum { MG_K_EMPTY = 0, MG_K_WALL = 1, MG_K_GOAL = 2, MG_K_LAVA = 3,
       MG_K_KEY = 8, MG_K_BALL = 16, MG_K_DOOR = 24, MG_K_BOX = 64 };

/* actions (custom_env.py:43-51) */
enum { MG_A_LEFT = 0, MG_A_RIGHT, MG_A_FORWARD, MG_A_PICKUP, MG_A_DROP, MG_A_TOGGLE, MG_A_DONE };

/* cfg.env.problem (custom_env.py:134-152) */
enum { MG_P_MULTI = 0, MG_P_GTO = 1, MG_P_GTG = 2, MG_P_OPN = 3, MG_P_PKP = 4, MG_P_DRP = 5, MG_P_MOV = 6, MG_P_FULL = 7 };

/* mission ids: group*24 + type4*6 + colour; type4: key 0 ball 1 box 2 door 3
 * group 0 'go to', 1 'toggle', 2 'pick up'; 72 'go to goal'; 73 'drop' */
#define MG_MISSION_GOAL 72
#define MG_MISSION_DROP 73
/* 'move left|right|up|down' (problems mov / full, custom_env.py:216-256) take ids 24..27: 'toggle <colour> key' is never
 * generated ('toggle' only picks boxes and doors, :197-201), so the table stays at 74 rows.  A 'move' episode has no target
 * position; its target_range (:221-254: per row / column the first empty cell seen from the named side, at generation
 * time) is kept in the four bytes target_x | target_y << 8 | target_action << 16 | pad << 24 as decimal digits:
 * digit k = the coordinate (1..size-2) for row y = k+1 (left/right) or column x = k+1 (up/down), 0 = no cell. */
#define MG_MISSION_MOVE0 24
#define MG_N_MISSIONS 74

/* per-environment state: 140 bytes (35 words) */
typedef struct {
    uint8_t grid[MG_GRID_CELLS]; /* kind bytes, grid[y*size + x] */
    uint8_t agent_x, agent_y, agent_dir;
    uint8_t carrying;            /* kind byte, 0 = nothing */
    uint8_t step_count;
    uint8_t target_x, target_y;  /* MG_NONE = no target position */
    uint8_t target_action;       /* 0 = None */
    uint8_t mission_id;
    uint8_t mission_done;        /* latch (custom_env.py:79); survives reset */
    uint8_t latch_step;          /* step_count when self.reward was stored */
    uint32_t episode;            /* RNG counter word: episodes generated so far for this env */
    uint16_t reset_draws;        /* RNG draws consumed by the last layout generation */
    uint8_t error;               /* bit0: invalid action seen, bit1: rejection cap hit */
    uint8_t pad;
} mg_state;

typedef struct {
    int32_t size;             /* cfg.env.size, 5..11 */
    int32_t num_objects;      /* cfg.env.num_objects */
    int32_t problem;          /* MG_P_* */
    int32_t mission;          /* cfg.env.mission: 0,1,2,5 or -1 for null (uniform over {0,1,2,5}) */
    int32_t all_doors_open;   /* cfg.env.all_doors_open */
    int32_t see_through_walls;
    int32_t max_steps;        /* size*size (custom_env.py:114) */
    int32_t reserved;
} mg_config;

/* ---- RNG: Philox4x32-10, key=(seed lo,hi), counter=(block, episode, env lo, env hi);
 * draw d of an episode = word d&3 of block d>>2; below(n) = mulhi32(word, n) ------------ */
void mg_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint32_t mg_draw_below(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t draw, uint32_t n);

void mg_reward_lut(int max_steps, float *lut /* [max_steps+1] */);
void mg_kind_encode(uint8_t kind, uint8_t out[3]);

/* layout generation for s->episode (custom_env.py:122-267); bumps s->episode afterwards.
 * keeps mission_done/latch_step (Q1).  returns number of draws, <0 on unsupported config */
int mg_generate(const mg_config *cfg, uint64_t seed, uint64_t env_id, mg_state *s);
/* fresh env: zero state, episode 0, then generate (== constructing PlaygroundEnv + reset) */
int mg_reset_env(const mg_config *cfg, uint64_t seed, uint64_t env_id, mg_state *s);

/* PlaygroundEnv.step without the observation; *carry_obs = what gen_obs saw as carrying (Q2) */
void mg_step_env(const mg_config *cfg, const float *lut, mg_state *s, int action,
                 float *reward, uint8_t *terminated, uint8_t *truncated, uint8_t *carry_obs);
/* MiniGridEnv.gen_obs: image[vx][vy][c] */
void mg_gen_obs(const mg_config *cfg, const mg_state *s, uint8_t carrying, uint8_t *image);
/* FullyObsWrapper: image[x][y][c] (size*size*3), agent cell = (10,0,dir) */
void mg_full_obs(const mg_config *cfg, const mg_state *s, uint8_t *image);

/* vector step with SB3-style auto-reset (DummyVecEnv.step_wait semantics), nthreads>=1 */
void mg_vec_reset(const mg_config *cfg, uint64_t seed, uint64_t env_id_base, int n, mg_state *states,
                  uint8_t *obs, uint8_t *dir, uint8_t *mission, int nthreads);
void mg_vec_step(const mg_config *cfg, uint64_t seed, uint64_t env_id_base, int n, mg_state *states,
                 const uint8_t *actions, uint8_t *obs, uint8_t *dir, uint8_t *mission, float *reward,
                 uint8_t *term, uint8_t *trunc, uint8_t *ep_len, uint8_t *term_obs, int nthreads);

/* SB3 RolloutBuffer.compute_returns_and_advantage, float32, exact operation order */
void mg_gae(const float *rewards, const float *values, const uint8_t *episode_starts,
            const float *last_values, const uint8_t *last_dones, double gamma, double gae_lambda,
            int T, int N, float *advantages, float *returns);

#ifdef __cplusplus
}
#endif
#endif
