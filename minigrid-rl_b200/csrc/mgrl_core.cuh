// mgrl_core.cuh — per-environment device functions of the batched MiniGrid simulator.
//
// One environment = one 140-byte packed state (EnvState).  Everything here is written as
// __host__ __device__ inline code over a state reference so that (a) the CUDA kernels in
// mgrl_kernels.cu run it one-lane-per-environment on shared-memory-resident tiles and
// (b) tests/ can compile the very same functions for the host and compare them with the
// CPU oracle without a GPU (tests/support/host_emul.cpp; test infrastructure, not a fallback).
//
// Reference behaviour implemented (paths relative to /root/reference/src/):
//   env_step      : [UPSTREAM] MiniGridEnv.step + PlaygroundEnv.step      custom_env.py:269-330
//   encode_view   : [UPSTREAM] MiniGridEnv.gen_obs / Grid.slice/rotate_left/process_vis/encode
//   generate      : PlaygroundEnv._gen_grid and the room generators       custom_env.py:122-267, 371-555, 595-2034
// This is a re-design, not a transcription: cells are one "kind" byte, the three room
// generators are one table-driven routine, object pools are bit masks, the 7x7 view is a
// closed-form affine map with clamping instead of slice+rotate.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MGRL_HD __host__ __device__ __forceinline__
#else
#define MGRL_HD inline
#endif

namespace mgrl {

constexpr int kMaxSize = 11;
constexpr int kGridCells = 121;
constexpr int kView = 7;
constexpr int kObsBytes = 147;
constexpr int kNone = 0xFF;
constexpr int kMaxTries = 1000;  // bound of the reference's `while True` rejection loops

// kind byte (see include/mgrl.h for the table)
constexpr int K_EMPTY = 0, K_WALL = 1, K_GOAL = 2, K_LAVA = 3, K_KEY = 8, K_BALL = 16, K_DOOR = 24, K_BOX = 64;
constexpr int A_LEFT = 0, A_RIGHT = 1, A_FORWARD = 2, A_PICKUP = 3, A_DROP = 4, A_TOGGLE = 5, A_DONE = 6;
constexpr int P_MULTI = 0, P_GTO = 1, P_GTG = 2, P_OPN = 3, P_PKP = 4, P_DRP = 5, P_MOV = 6, P_FULL = 7;
constexpr int MISSION_GOAL = 72, MISSION_DROP = 73;
// 'move left|right|up|down' (problems mov / full, custom_env.py:216-256) take the ids of 'toggle <colour> key', which is
// never generated ('toggle' only picks boxes and doors, :197-201): the mission table keeps its 74 rows.  Such an episode has
// no target position; its target_range (per row / column the first empty cell seen from the named side, at generation time)
// lives in target_x | target_y << 8 | target_action << 16 | pad << 24 as decimal digits: digit k = the coordinate
// (1..size-2) of the cell in row y = k+1 (left / right) or column x = k+1 (up / down), 0 = no cell.
constexpr int MISSION_MOVE0 = 24;
constexpr int ERR_BAD_ACTION = 1, ERR_TRIES = 2, ERR_SYNC = 4;  // ERR_SYNC: a prepared layout never arrived (kernel bug guard)
constexpr int OBS_HWC = 0;  // image[vx][vy][c]  (MiniGrid native), 147-byte records
constexpr int OBS_CHW = 1;  // image[c][vx][vy]  (after SB3 VecTransposeImage), 147-byte records

struct EnvState {  // 140 bytes = 35 words (odd word stride: conflict-free lane-per-env smem access)
    uint8_t grid[kGridCells];
    uint8_t agent_x, agent_y, agent_dir;
    uint8_t carrying;
    uint8_t step_count;
    uint8_t target_x, target_y;
    uint8_t target_action;
    uint8_t mission_id;
    uint8_t mission_done;
    uint8_t latch_step;
    uint32_t episode;
    uint16_t reset_draws;
    uint8_t error;
    uint8_t pad;
};
static_assert(sizeof(EnvState) == 140, "EnvState must be 140 bytes");

struct EnvCfg {
    int32_t size, num_objects, problem, mission, all_doors_open, see_through_walls, max_steps, num_obstacles;
};

// ------------------------------------------------------------------------------- helpers
MGRL_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
MGRL_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// position of the n-th (0-based) set bit of m
MGRL_HD int nth_set_bit(uint32_t m, int n) {
    int pos = 0, t;
    t = popc32(m & 0xFFFFu); if (n >= t) { n -= t; pos = 16; m >>= 16; }
    t = popc32(m & 0xFFu);   if (n >= t) { n -= t; pos += 8; m >>= 8; }
    t = popc32(m & 0xFu);    if (n >= t) { n -= t; pos += 4; m >>= 4; }
    t = popc32(m & 0x3u);    if (n >= t) { n -= t; pos += 2; m >>= 2; }
    t = (int)(m & 1u);       if (n >= t) { pos += 1; }
    return pos;
}

MGRL_HD bool k_is_key(int k) { return (k >> 3) == 1; }
MGRL_HD bool k_is_ball(int k) { return (k >> 3) == 2; }
MGRL_HD bool k_is_door(int k) { return (unsigned)(k - K_DOOR) < 24u; }
MGRL_HD bool k_is_box(int k) { return k >= K_BOX; }
MGRL_HD bool k_pickable(int k) { return (unsigned)(k - K_KEY) < 16u || k >= K_BOX; }
MGRL_HD bool k_opaque(int k) { return k == K_WALL || (unsigned)(k - (K_DOOR + 8)) < 16u; }  // !see_behind

// packed (type | colour<<8 | state<<16) observation encoding of a kind byte
MGRL_HD uint32_t kind_encode(int k) {
    if (k < 8) {
        const uint32_t t = (0x9821u >> ((k & 3) * 4)) & 0xFu;
        const uint32_t c = (0x0150u >> ((k & 3) * 4)) & 0xFu;
        return t | (c << 8);
    }
    const uint32_t c = (uint32_t)(k & 7) << 8;
    if (k < K_BALL) return 5u | c;
    if (k < K_DOOR) return 6u | c;
    if (k < K_DOOR + 24) return 4u | c | ((uint32_t)((k - K_DOOR) >> 3) << 16);
    return 7u | c;
}

// ---------------------------------------------------------------------------------- RNG
// Philox4x32-10, key = seed, counter = (block, episode, env_lo, env_hi).  Draw d of an
// episode is word d&3 of block d>>2; below(n) = mulhi32(word, n).  The generator keeps a ring of
// kRing draws per layout (lane-interleaved shared memory): kUpFront blocks before the straight-line
// prologue, then one more block per placement iteration whenever the ring has room, so the Philox
// rounds overlap the dependent load-compare chain of the placement loop; a draw the ring does not hold
// (only if a lane consumed five draws per iteration for many iterations in a row) is recomputed.
MGRL_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t ka, uint32_t kb, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;   // one wide multiply each
        const uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0;
        const uint32_t h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
        c0 = h1 ^ c1 ^ ka; c1 = l1; c2 = h0 ^ c3 ^ kb; c3 = l0;
        ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// A block as a call: the generator tops its ring up from five places (prologue, door loops, placement loop, target
// selection); inlined, every one of them is another ~100 instructions of rounds in the rollout kernel's instruction stream,
// and that kernel is bound by instruction issue and fetch.  Four words come back in registers.
struct PhiloxBlock { uint32_t w0, w1, w2, w3; };
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
PhiloxBlock philox_block(uint32_t block, uint32_t episode, uint32_t e0, uint32_t e1, uint32_t k0, uint32_t k1) {
    uint32_t w[4];
    philox4x32_10(block, episode, e0, e1, k0, k1, w);
    return PhiloxBlock{w[0], w[1], w[2], w[3]};
}
// one word of the stream, recomputed (draws beyond the precomputed buffer)
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
uint32_t philox_word(uint32_t idx, uint32_t episode, uint32_t e0, uint32_t e1, uint32_t k0, uint32_t k1) {
    uint32_t w[4];
    philox4x32_10(idx >> 2, episode, e0, e1, k0, k1, w);
    const uint32_t j = idx & 3u;
    return j == 0 ? w[0] : j == 1 ? w[1] : j == 2 ? w[2] : w[3];
}

// ---- 'move <direction>' missions (problems mov / full): rare, kept out of line so that the step and the generator of the
// other problems do not carry their code in the instruction stream
#if defined(__CUDACC__)
#define MGRL_COLD __host__ __device__ __noinline__
#else
#define MGRL_COLD inline
#endif
// custom_env.py:314-317: is (ax, ay) in the episode's target range (decimal digits v, MISSION_MOVE0 + d)
MGRL_COLD bool move_hit(uint32_t v, int d, int ax, int ay) {
    const bool rows = d < 2;
    const int kk = (rows ? ay : ax) - 1;
    uint32_t p10 = 1u;
    for (int i = 0; i < kk; ++i) p10 *= 10u;
    return (int)(v / p10 % 10u) == (rows ? ax : ay);
}
// custom_env.py:221-254: per row (left / right) or column (up / down) the first empty cell seen from the named side, at
// generation time (the agent's own cell is empty in the grid), as decimal digits
MGRL_COLD uint32_t move_range(const uint8_t* grid, int S, int d) {
    uint32_t v = 0u, p10 = 1u;
    for (int k = 1; k <= S - 2; ++k, p10 *= 10u) {
        const int step = (d & 1) ? -1 : 1;
        int c = (d & 1) ? S - 2 : 1;
        while (c > 0 && c < S - 1 && grid[d < 2 ? k * S + c : c * S + k] != K_EMPTY) c += step;
        if (c >= S - 1) c = 0;
        v += (uint32_t)c * p10;
    }
    return v;
}

// --------------------------------------------------------------------------------- step
struct StepOut {
    float reward;
    uint8_t terminated, truncated, carry_obs;
    uint8_t step_count;   // steps of the episode including this one
    uint8_t error;        // ERR_BAD_ACTION when the action was not a MiniGrid action (KEEP_IF_DONE only: not stored)
    uint32_t w32;         // word 32 after the step (target_action, mission_id, mission_done, latch_step)
};

// [UPSTREAM] MiniGridEnv.step, then PlaygroundEnv.step's mission bookkeeping.
// reward_lut[k] = float32(1 - 0.9*k/max_steps) computed in float64 on the host.
// KEEP_IF_DONE: when the step ends the episode nothing is written back to the state (the caller is about to
// overwrite it with the layout of the next episode, possibly by an asynchronous copy that is already in flight);
// the surviving fields (mission latch, error bit) are returned in the StepOut.
// what a step reads from the state: the scalar fields travel as three words: 30 = grid[120], agent x, y, dir;
// 31 = carrying, step_count, target x, y; 32 = target_action, mission_id, mission_done, latch_step; and the cell in front
struct StepIn {
    uint32_t w30, w31, w32;
    int fidx, k;
};
MGRL_HD StepIn env_step_load(const EnvState& s, int S) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&s);
    StepIn in;
    in.w30 = w[30]; in.w31 = w[31]; in.w32 = w[32];
    const int dir = (int)(in.w30 >> 24);
    const int ax = (int)((in.w30 >> 8) & 0xFFu), ay = (int)((in.w30 >> 16) & 0xFFu);
    in.fidx = (ay + (dir == 1) - (dir == 3)) * S + ax + (dir == 0) - (dir == 2);
    in.k = s.grid[in.fidx];
    return in;
}

// EXTRA: the problems mov / full ('move <direction>' missions) are compiled in; the rollout kernel of the other problems is
// instantiated without them (it is bound by instruction issue: their three rare branches cost it 3 %).
template <bool KEEP_IF_DONE, bool EXTRA = true>
MGRL_HD StepOut env_step_apply(EnvState& s, const StepIn& in, int action, int max_steps, const float* reward_lut) {
    uint32_t* w = reinterpret_cast<uint32_t*>(&s);
    const uint32_t w30 = in.w30, w31 = in.w31, w32 = in.w32;
    const int dir = (int)(w30 >> 24);
    int ax = (int)((w30 >> 8) & 0xFFu), ay = (int)((w30 >> 16) & 0xFFu);
    int carrying = (int)(w31 & 0xFFu);
    const int step = (int)((w31 >> 8) & 0xFFu) + 1;
    const int tx = (int)((w31 >> 16) & 0xFFu), ty = (int)(w31 >> 24);
    const int ta = (int)(w32 & 0xFFu), mission_id = (int)((w32 >> 8) & 0xFFu);
    int mdone = (int)((w32 >> 16) & 0xFFu), latch = (int)(w32 >> 24);
    const int dx = (dir == 0) - (dir == 2), dy = (dir == 1) - (dir == 3);
    const int fidx = in.fidx;
    const int k = in.k;
    int nk = k;                // front cell after the action
    int ndir = dir;
    bool term = false;
    float r = 0.0f;
    StepOut o;
    o.error = 0;

    if (action == A_LEFT) ndir = (dir + 3) & 3;
    else if (action == A_RIGHT) ndir = (dir + 1) & 3;
    else if (action == A_FORWARD) {
        if (k == K_EMPTY || k == K_GOAL || k == K_LAVA || (unsigned)(k - K_DOOR) < 8u) { ax += dx; ay += dy; }
        if (k == K_GOAL) { term = true; r = reward_lut[step]; }
        if (k == K_LAVA) term = true;
    } else if (action == A_PICKUP) {
        if (k_pickable(k) && carrying == 0) { carrying = k; nk = K_EMPTY; }
    } else if (action == A_DROP) {
        if (k == K_EMPTY && carrying != 0) { nk = carrying; carrying = 0; }
    } else if (action == A_TOGGLE) {
        const int st = (k - K_DOOR) >> 3, c = k & 7;
        if (k_is_door(k)) {
            if (st != 2) nk = K_DOOR + 8 * (st ^ 1) + c;
            else if (k_is_key(carrying) && (carrying & 7) == c) nk = K_DOOR + c;   // locked: needs a Key of its colour
        } else if (k_is_box(k)) {  // the box is replaced by its contents
            const int m = (k - K_BOX) >> 3;
            nk = m ? K_KEY + m - 1 : K_EMPTY;
        }
    } else if (action != A_DONE) {
        if (KEEP_IF_DONE) o.error = ERR_BAD_ACTION;
        else s.error |= ERR_BAD_ACTION;  // upstream raises ValueError
    }
    o.truncated = step >= max_steps;
    o.step_count = (uint8_t)step;
    o.carry_obs = (uint8_t)carrying;  // the observation is rendered here (custom_env.py:270)

    if (term) {  // custom_env.py:272-277
        if (mission_id != MISSION_GOAL) { mdone = 0; latch = 0; r = 0.0f; }
    } else {
        const int fx = ax + (ndir == 0) - (ndir == 2), fy = ay + (ndir == 1) - (ndir == 3);
        // :279-283 colour match only; a toggle neither turns nor moves, so the front cell is the one just rewritten
        if (action == A_TOGGLE && k_is_door(nk) && carrying != 0 && (nk & 7) == (carrying & 7)) carrying = 0;
        if (!mdone) {  // :288-317
            bool hit;
            if (EXTRA && (unsigned)(mission_id - MISSION_MOVE0) < 4u) {   // :314-317 agent_pos in target_range
                const uint32_t v = (w31 >> 16) | ((w32 & 0xFFu) << 16) | (reinterpret_cast<const uint32_t*>(&s)[34] & 0xFF000000u);
                hit = move_hit(v, mission_id - MISSION_MOVE0, ax, ay);
            } else if (tx != kNone) hit = ta ? (fx == tx && fy == ty && action == ta) : (ax == tx && ay == ty);
            else hit = ta != 0 && action == ta;
            if (hit) { mdone = 1; latch = step; }
        }
        if (action == A_DONE) {  // :319-328
            r = mdone ? reward_lut[latch] : 0.0f;
            mdone = 0; latch = 0;
            term = true;
        }
    }
    o.w32 = (w32 & 0xFFFFu) | ((uint32_t)mdone << 16) | ((uint32_t)latch << 24);
    o.reward = r;
    o.terminated = term;
    if (!(KEEP_IF_DONE && (term || o.truncated))) {
        if (nk != k) s.grid[fidx] = (uint8_t)nk;
        w[30] = (w30 & 0xFFu) | ((uint32_t)ax << 8) | ((uint32_t)ay << 16) | ((uint32_t)ndir << 24);
        w[31] = (uint32_t)carrying | ((uint32_t)step << 8) | (w31 & 0xFFFF0000u);
        w[32] = o.w32;
        if (KEEP_IF_DONE && o.error) s.error |= o.error;
    }
    return o;
}

MGRL_HD StepOut env_step(EnvState& s, int action, int S, int max_steps, const float* reward_lut) {
    return env_step_apply<false>(s, env_step_load(s, S), action, max_steps, reward_lut);
}

// -------------------------------------------------------------------------- observation
// View cell (vx,vy) shows world cell agent + (vx-3)*right + (6-vy)*dir.  The border of the
// grid is always wall and out-of-grid cells render as wall, so coordinates are clamped
// instead of bounds-checked.  The map is separable: the forward coordinate depends only on
// vy and the sideways coordinate only on vx, so 7+7 clamped offsets replace 49x2 clamps:
//     grid index of view cell (vx,vy) = ro[vx] + fo[vy].
constexpr int kObsPitch148 = 148;  // word-aligned record: 49 x (type,colour,state) + 1 pad byte
constexpr int OBS_HWC148 = 2;      // image[vx][vy][c], record pitch 148 B (fast path: packed word stores)

template <int LAYOUT>
MGRL_HD int obs_index(int vx, int vy, int c) {
    return LAYOUT == OBS_CHW ? c * (kView * kView) + vx * kView + vy : (vx * kView + vy) * 3 + c;
}

MGRL_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct ViewMap {
    int fo[kView], ro[kView];
};

MGRL_HD void view_map(const EnvState& s, int S, ViewMap& m) {
    const int dir = s.agent_dir;
    const bool even = (dir & 1) == 0;                 // forward axis is x for east/west
    const int sf = dir < 2 ? 1 : -1;                  // forward step: +1 east/south, -1 west/north
    const int sr = (dir == 0 || dir == 3) ? 1 : -1;   // sideways (view-x) step
    const int af = even ? s.agent_x : s.agent_y, ar = even ? s.agent_y : s.agent_x;
    const int mf = even ? 1 : S, mr = even ? S : 1;
#pragma unroll
    for (int i = 0; i < kView; ++i) {
        m.fo[i] = clampi(af + (6 - i) * sf, 0, S - 1) * mf;
        m.ro[i] = clampi(ar + (i - 3) * sr, 0, S - 1) * mr;
    }
}

// lut[k] = kind_encode(k) for k < 128 (shared memory on the device)
MGRL_HD void fill_kind_lut(uint32_t* lut, int tid, int nthreads) {
    for (int k = tid; k < 128; k += nthreads) lut[k] = kind_encode(k);
}

MGRL_HD uint32_t pack3(uint32_t lo, uint32_t hi, int shift_bytes) {  // bytes of (hi:lo) >> 8*shift, hi's top byte is 0
#if defined(__CUDA_ARCH__)
    return __byte_perm(lo, hi, shift_bytes == 0 ? 0x4210 : shift_bytes == 1 ? 0x5421 : 0x6542);
#else
    return shift_bytes == 0 ? (lo | (hi << 24)) : shift_bytes == 1 ? ((lo >> 8) | (hi << 16)) : ((lo >> 16) | (hi << 8));
#endif
}

// fast path: see-through view, HWC, 37 aligned words (148 B) per environment.  Row by row: the
// sideways offset of a view row is computed once, the seven forward offsets stay in registers.
MGRL_HD void encode_view_packed(const EnvState& s, int carrying, int S, const uint32_t* lut, uint32_t* out) {
    const int dir = s.agent_dir;
    const bool even = (dir & 1) == 0;                 // forward axis is x for east/west
    const int sf = dir < 2 ? 1 : -1;                  // forward step: +1 east/south, -1 west/north
    const int sr = (dir == 0 || dir == 3) ? 1 : -1;   // sideways (view-x) step
    const int af = even ? s.agent_x : s.agent_y, ar = even ? s.agent_y : s.agent_x;
    const int mf = even ? 1 : S, mr = even ? S : 1;
    int fo[kView];
#pragma unroll
    for (int i = 0; i < kView; ++i) fo[i] = clampi(af + (6 - i) * sf, 0, S - 1) * mf;
    uint32_t e[4];
#pragma unroll
    for (int vx = 0; vx < kView; ++vx) {
        const uint8_t* row = s.grid + clampi(ar + (vx - 3) * sr, 0, S - 1) * mr;
#pragma unroll
        for (int vy = 0; vy < kView; ++vy) {
            const int c = vx * kView + vy;
            const int k = (c == 3 * kView + 6) ? carrying : (int)row[fo[vy]];
            e[c & 3] = lut[k];
            if ((c & 3) == 3) {
                const int w = (c >> 2) * 3;
                out[w] = pack3(e[0], e[1], 0);
                out[w + 1] = pack3(e[1], e[2], 1);
                out[w + 2] = pack3(e[2], e[3], 2);
            }
        }
    }
    out[36] = e[0];  // cell 48 + pad byte
}

// byte-granular path (any layout, 147-byte records; also used for terminal observations)
template <int LAYOUT>
MGRL_HD void encode_view_see_through(const EnvState& s, int carrying, int S, const uint32_t* lut, uint8_t* out) {
    ViewMap m;
    view_map(s, S, m);
#pragma unroll
    for (int vx = 0; vx < kView; ++vx) {
#pragma unroll
        for (int vy = 0; vy < kView; ++vy) {
            const int k = (vx == 3 && vy == 6) ? carrying : (int)s.grid[m.ro[vx] + m.fo[vy]];
            const uint32_t e = lut[k];
            out[obs_index<LAYOUT>(vx, vy, 0)] = (uint8_t)e;
            out[obs_index<LAYOUT>(vx, vy, 1)] = (uint8_t)(e >> 8);
            out[obs_index<LAYOUT>(vx, vy, 2)] = (uint8_t)(e >> 16);
        }
    }
}

// see_through_walls == false: [UPSTREAM] Grid.process_vis, mask kept as a 49-bit set
template <int LAYOUT>
MGRL_HD void encode_view_occluded(const EnvState& s, int carrying, int S, const uint32_t* lut, uint8_t* out) {
    ViewMap m;
    view_map(s, S, m);
    uint64_t opaque = 0;  // bit vx*7+vy
    for (int vx = 0; vx < kView; ++vx)
        for (int vy = 0; vy < kView; ++vy)
            if (k_opaque(s.grid[m.ro[vx] + m.fo[vy]])) opaque |= 1ull << (vx * kView + vy);
    uint64_t mask = 1ull << (3 * kView + 6);
    for (int j = kView - 1; j >= 0; --j) {
        for (int i = 0; i < kView - 1; ++i) {
            const uint64_t b = 1ull << (i * kView + j);
            if (!(mask & b) || (opaque & b)) continue;
            mask |= 1ull << ((i + 1) * kView + j);
            if (j > 0) mask |= (1ull << ((i + 1) * kView + j - 1)) | (1ull << (i * kView + j - 1));
        }
        for (int i = kView - 1; i >= 1; --i) {
            const uint64_t b = 1ull << (i * kView + j);
            if (!(mask & b) || (opaque & b)) continue;
            mask |= 1ull << ((i - 1) * kView + j);
            if (j > 0) mask |= (1ull << ((i - 1) * kView + j - 1)) | (1ull << (i * kView + j - 1));
        }
    }
    for (int vx = 0; vx < kView; ++vx)
        for (int vy = 0; vy < kView; ++vy) {
            const int k = (vx == 3 && vy == 6) ? carrying : (int)s.grid[m.ro[vx] + m.fo[vy]];
            const uint32_t e = ((mask >> (vx * kView + vy)) & 1ull) ? lut[k] : 0u;
            const int pitch_layout = LAYOUT == OBS_HWC148 ? OBS_HWC : LAYOUT;
            out[obs_index<pitch_layout>(vx, vy, 0)] = (uint8_t)e;
            out[obs_index<pitch_layout>(vx, vy, 1)] = (uint8_t)(e >> 8);
            out[obs_index<pitch_layout>(vx, vy, 2)] = (uint8_t)(e >> 16);
        }
    if (LAYOUT == OBS_HWC148) out[147] = 0;
}

// record pitch of a layout
MGRL_HD constexpr int obs_pitch(int layout) { return layout == OBS_HWC148 ? kObsPitch148 : kObsBytes; }

// out must be 4-byte aligned for OBS_HWC148
template <int LAYOUT>
MGRL_HD void encode_view(const EnvState& s, int carrying, int S, bool see_through, const uint32_t* lut, uint8_t* out) {
    if (!see_through) encode_view_occluded<LAYOUT>(s, carrying, S, lut, out);
    else if (LAYOUT == OBS_HWC148) encode_view_packed(s, carrying, S, lut, reinterpret_cast<uint32_t*>(out));
    else encode_view_see_through<LAYOUT>(s, carrying, S, lut, out);
}

// [UPSTREAM] FullyObsWrapper: image[x][y][c], agent cell (10, 0, dir)
MGRL_HD void encode_full(const EnvState& s, int S, uint8_t* out) {
    for (int x = 0; x < S; ++x)
        for (int y = 0; y < S; ++y) {
            uint32_t e = kind_encode(s.grid[y * S + x]);
            if (x == s.agent_x && y == s.agent_y) e = 10u | ((uint32_t)s.agent_dir << 16);
            uint8_t* o = out + (x * S + y) * 3;
            o[0] = (uint8_t)e; o[1] = (uint8_t)(e >> 8); o[2] = (uint8_t)(e >> 16);
        }
}


// ----------------------------------------------------------------------------- generator
// PlaygroundEnv._gen_grid (custom_env.py:122-267) with the map generators (:371-555 single
// room, :595-2034 two/three/four rooms).
//
// The reference is a long chain of rejection loops ("draw a position, retry until it is
// admissible") whose number and kind depend on earlier draws.  It is run here one lane per
// environment on DENSE warps (every lane starts a layout at the same time), shaped so the lanes
// stay converged for as long as the data allows:
//   1. all Philox blocks of the episode are computed up front (no per-draw RNG branch);
//   2. the multi-room prologue (mission, room count, door colours/locks, door cells) is
//      straight-line code over the maximum of four doors, predicated per lane;
//   3. every placement kind (goal, agent, key / key-in-box, distractor, single-room object,
//      obstacle) is one shared draw-test-commit body driven by a 32-bit task word (rectangle to
//      draw from + admissibility flags); a lane's sequence of tasks is a row of a table indexed by
//      (rooms, agent room, goal room, locked doors), so advancing to the next task is a word
//      fetch instead of a nested room walk;
//   4. cells next to a door carry a flag bit while the layout is built, so "empty and not
//      next to a door" (next2door, :2036-2046) is one compare of the byte already loaded.
// The draw order is the reference's; the CPU oracle consumes the same stream.
#ifndef MGRL_GEN_UPFRONT
#define MGRL_GEN_UPFRONT 2
#endif
#ifndef MGRL_GEN_AHEAD
#define MGRL_GEN_AHEAD 6
#endif
constexpr int T_KEY = 0, T_BALL = 1, T_BOX = 2, T_DOOR = 3, T_GOAL = 4;
constexpr int kRing = 16;               // draws held per generation: a ring over the episode's Philox stream
constexpr int kUpFront = MGRL_GEN_UPFRONT;   // blocks computed before the prologue
constexpr int kAhead = MGRL_GEN_AHEAD;       // top-up threshold: a block is produced when no more than kAhead draws are ahead
constexpr int kObjWords = 14;           // the placed objects in insertion order, 16 bits each (<= 28 objects):
                                        //   type | colour << 3 | cell << 6
constexpr int kGenWords = kRing + kObjWords;   // words of generation scratch per layout besides the 35 state words
constexpr int kGridWords = 31;          // words 0..30 of EnvState cover grid[121] + agent x/y/dir
constexpr uint32_t kDoorFlag = 0x80u;   // "next to a door" mark on a grid byte (kinds are < 128)
// generation scratch words a configuration needs: the ring + its object records (multi: <= 4 doors, the goal, <= 4 keys and
// <= num_objects distractors; single room: the objects and the goal; full: 24 + goal) - the rollout kernel sizes its
// generator warps' shared memory with it (fewer words per warp = room for one more generator warp)
MGRL_HD int gen_words_for(const EnvCfg& cfg) {
    const int objs = cfg.problem == P_MULTI ? 9 + cfg.num_objects : cfg.problem == P_FULL ? 25 : cfg.num_objects + 1;
    const int words = (objs + 1) / 2;
    return kRing + (words < kObjWords ? words : kObjWords);
}

// Keys placed in `room` when the agent starts in `agent_room`: up to
// two door indices, 7 = none.  One byte per agent room: low nibble = first key, high = second.
MGRL_HD constexpr uint32_t key_byte(int first, int second) { return (uint32_t)(first | (second << 4)); }
MGRL_HD constexpr uint32_t key_row(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {
    return a0 | (a1 << 8) | (a2 << 16) | (a3 << 24);
}
MGRL_HD int key_door(int nrooms, int room, int agent_room, int j) {
    if (nrooms == 2) return (room == agent_room && j == 0) ? 0 : 7;  // the single door, agent's room
    if (nrooms == 3) {  // doors h=0 vu=1 vl=2; only the agent's room gets keys
        if (room != agent_room) return 7;
        const int first = room == 0 ? 1 : 2, second = room == 2 ? 1 : 0;
        return j == 0 ? first : second;
    }
    // 4 rooms UL LL UR LR, doors hl=0 hr=1 vu=2 vl=3; columns = agent in UL, LL, UR, LR
    constexpr uint32_t N = key_byte(7, 7);
    constexpr uint32_t UL = key_row(key_byte(2, 0), key_byte(2, 7), key_byte(0, 7), N);
    constexpr uint32_t LL = key_row(key_byte(3, 7), key_byte(3, 0), N, key_byte(0, 7));
    constexpr uint32_t UR = key_row(key_byte(1, 7), N, key_byte(2, 1), key_byte(2, 7));
    constexpr uint32_t LR = key_row(N, key_byte(1, 7), key_byte(3, 7), key_byte(3, 1));
    const uint32_t t = room == 0 ? UL : room == 1 ? LL : room == 2 ? UR : LR;
    return (int)((t >> (8 * agent_room + 4 * j)) & 0xFu);
}

MGRL_HD int room_of(int nrooms, int mid, int x, int y) {
    const bool left = x < mid, upper = y < mid;
    if (nrooms == 2) return left ? 0 : 1;
    if (nrooms == 3) return left ? (upper ? 0 : 1) : 2;
    return left ? (upper ? 0 : 1) : (upper ? 2 : 3);
}

// COLOR_NAMES sorted alphabetically (blue green grey purple red yellow) <-> COLOR_TO_IDX
MGRL_HD int sorted_colour(int i) { return (int)((0x403512u >> (4 * i)) & 0xFu); }
MGRL_HD int sorted_index(int colour) { return (int)((0x253014u >> (4 * colour)) & 0xFu); }
MGRL_HD int obj_kind(int type, int colour) {
    return type == T_KEY ? K_KEY + colour : type == T_BALL ? K_BALL + colour
         : type == T_BOX ? K_BOX + colour : type == T_DOOR ? K_DOOR + 8 + colour : K_GOAL;
}

// A placement task is one 32-bit word; a lane walks a zero-terminated list of them:
//   bits 0-3 x0, 4-7 width, 8-11 y0, 12-15 height of the rectangle the position is drawn from
//   bits 16-18 stage, 19-20 door (key tasks), 21 second key of the room, 22-23 room,
//   bit 24 reject x == mid / y == mid (lava on a multi map), 25 lava may be overwritten,
//   bit 26 only "not next to a door" is asked of the cell (key tasks), 27 agent (next2door ignored),
//   bit 28 the agent's own cell is rejected
// The list before the agent exists is configuration-constant (GenIO.prefix: objects, goal, agent); what
// follows (keys of locked doors, distractors per room, obstacles) is a row of a table indexed by
// (rooms, agent room, goal room, locked doors) built once per configuration on the host.
enum GenStage : int { G_OBJ = 1, G_GOAL, G_AGENT, G_KEY, G_DIST, G_OBST, G_FIXED };   // G_FIXED: object (bits 19-23) without a draw
constexpr uint32_t TF_MID = 1u << 24, TF_LAVA = 1u << 25, TF_KEYMODE = 1u << 26, TF_AGENTMODE = 1u << 27, TF_AGENT_CELL = 1u << 28;
constexpr int kTaskWords = 32;            // words per table row / prefix list, zero terminated
// rows exist for the valid combinations only: 2 rooms: 2 x 2 x 2 lock states of its one door; 3 rooms: 3 x 3 x 8;
// 4 rooms: 4 x 4 x 16  ->  8 + 72 + 256 rows
constexpr int kTaskEntries = 8 + 72 + 256;

MGRL_HD int task_index(int nrooms, int agent_room, int goal_room, uint32_t locked) {
    const int ndoors = nrooms == 2 ? 1 : nrooms;
    const int base = nrooms == 2 ? 0 : nrooms == 3 ? 8 : 80;
    return base + (((agent_room * nrooms + goal_room) << ndoors) | (int)locked);
}
inline uint32_t task_word(int stage, int x0, int x1, int y0, int y1, uint32_t flags, int room = 0, int door = 0, int second = 0) {
    return (uint32_t)x0 | ((uint32_t)(x1 - x0 + 1) << 4) | ((uint32_t)y0 << 8) | ((uint32_t)(y1 - y0 + 1) << 12) |
           ((uint32_t)stage << 16) | ((uint32_t)door << 19) | ((uint32_t)second << 21) | ((uint32_t)room << 22) | flags;
}
inline uint32_t room_task(int stage, int S, int nrooms, int r, uint32_t flags, int door, int second) {
    const int m = S / 2;
    const bool left = nrooms == 2 ? r == 0 : r < 2;
    const bool full_height = nrooms == 2 || (nrooms == 3 && r == 2);
    const bool upper = (r & 1) == 0;
    const int x0 = left ? 1 : m + 1, x1 = left ? m - 1 : S - 2;
    const int y0 = (full_height || upper) ? 1 : m + 1, y1 = (full_height || !upper) ? S - 2 : m - 1;
    return task_word(stage, x0, x1, y0, y1, flags, r, door, second);
}
inline void append_obstacles(const EnvCfg& cfg, uint32_t* out, int& n) {
    const int S = cfg.size;
    const bool multi = cfg.problem == P_MULTI;
    for (int i = 0; i < cfg.num_obstacles && n < kTaskWords - 1; ++i)
        out[n++] = multi ? task_word(G_OBST, 1, S - 2, 1, S - 2, TF_MID | TF_LAVA | TF_AGENT_CELL)     // :155-172
                         : task_word(G_OBST, 0, S - 1, 0, S - 1, TF_AGENT_CELL);                         // place_obj
}

// The per-room walk of _generate_{2,3,4}_rooms (:652-855, :931-1297, :1392-2034) for one
// (rooms, agent room, goal room, locked doors) combination: keys of the room's locked doors first
// (each uses up one of the room's distractor slots), one slot less in the goal's room, then the
// room's distractors.  Reference quirk (:1119, :1660): the lower-left loop reads the upper-left counter.
inline void build_task_row(const EnvCfg& cfg, int nrooms, int agent_room, int goal_room, uint32_t locked, uint32_t* out) {
    const int S = cfg.size, nl = cfg.num_objects / 2, nr = cfg.num_objects - nl;
    int cnt[4] = {0, 0, 0, 0};
    if (nrooms == 2) { cnt[0] = nl; cnt[1] = nr; }
    else if (nrooms == 3) { cnt[0] = nl / 2; cnt[1] = nl - nl / 2; cnt[2] = nr; }
    else { cnt[0] = nl / 2; cnt[1] = nl - nl / 2; cnt[2] = nr / 2; cnt[3] = nr - nr / 2; }
    int n = 0;
    for (int i = 0; i < kTaskWords; ++i) out[i] = 0u;
    if (agent_room >= nrooms || goal_room >= nrooms) return;
    for (int r = 0; r < nrooms; ++r) {
        for (int j = 0; j < 2; ++j) {
            const int d = key_door(nrooms, r, agent_room, j);
            if (d != 7 && ((locked >> d) & 1u)) {
                if (n < kTaskWords - 1)
                    out[n++] = room_task(G_KEY, S, nrooms, r, TF_KEYMODE | (r == agent_room ? TF_AGENT_CELL : 0u), d, j);
                --cnt[r];
            }
        }
        if (goal_room == r) --cnt[r];
        const int loops = (nrooms >= 3 && r == 1) ? cnt[0] : cnt[r];
        for (int q = 0; q < loops; ++q)
            if (n < kTaskWords - 1) out[n++] = room_task(G_DIST, S, nrooms, r, TF_AGENT_CELL, 0, 0);
    }
    append_obstacles(cfg, out, n);
}
// table[kTaskEntries][kTaskWords]; a single-room problem uses row 0 (its obstacles)
inline void build_task_table(const EnvCfg& cfg, uint32_t* table) {
    for (int i = 0; i < kTaskEntries * kTaskWords; ++i) table[i] = 0u;
    if (cfg.problem != P_MULTI) {
        int n = 0;
        append_obstacles(cfg, table, n);
        return;
    }
    for (int nrooms = 2; nrooms <= 4; ++nrooms)
        for (int a = 0; a < nrooms; ++a)
            for (int g = 0; g < nrooms; ++g)
                for (uint32_t l = 0; l < (1u << (nrooms == 2 ? 1 : nrooms)); ++l)
                    build_task_row(cfg, nrooms, a, g, l, table + (size_t)task_index(nrooms, a, g, l) * kTaskWords);
}
// Rows are mostly short (a dozen tasks at num_objects = 4): repack the table to the longest row so that it stays
// L1-resident next to the kernels' shared memory (36 KB instead of 96 KB).  Returns the row pitch in words.
inline int pack_task_table(const uint32_t* table, uint32_t* packed) {
    int longest = 0;
    for (int r = 0; r < kTaskEntries; ++r)
        for (int i = 0; i < kTaskWords; ++i)
            if (table[r * kTaskWords + i] != 0u && i + 1 > longest) longest = i + 1;
    int pitch = (longest + 1 + 3) & ~3;
    if (pitch > kTaskWords) pitch = kTaskWords;
    for (int r = 0; r < kTaskEntries; ++r)
        for (int i = 0; i < pitch; ++i) packed[r * pitch + i] = i < kTaskWords ? table[r * kTaskWords + i] : 0u;
    return pitch;
}
// tasks up to and including the agent: single-room objects (:371-555), goal, agent
inline void build_task_prefix(const EnvCfg& cfg, uint32_t* out /* [kTaskWords] */) {
    const int S = cfg.size;
    const bool multi = cfg.problem == P_MULTI;
    const bool has_goal = multi || cfg.problem == P_GTG || cfg.problem == P_DRP || cfg.problem == P_FULL;
    int n = 0;
    for (int i = 0; i < kTaskWords; ++i) out[i] = 0u;
    if (cfg.problem == P_FULL)   // _generate_full_map :332-369: every (type, colour) pair, types outer, COLOR_NAMES inner
        for (int i = 0; i < 24; ++i) out[n++] = task_word(G_FIXED, 0, S - 1, 0, S - 1, TF_AGENT_CELL) | ((uint32_t)i << 19);
    else if (!multi)
        for (int i = 0; i < cfg.num_objects && n < kTaskWords - 3; ++i) out[n++] = task_word(G_OBJ, 0, S - 1, 0, S - 1, TF_AGENT_CELL);
    if (has_goal) out[n++] = task_word(G_GOAL, 0, S - 1, 0, S - 1, TF_AGENT_CELL);
    out[n++] = task_word(G_AGENT, 0, S - 1, 0, S - 1, TF_AGENTMODE);
}
// words 0..30 of a fresh S x S grid: empty interior, wall border (Grid.wall_rect, custom_env.py:132)
inline void build_empty_grid(int S, uint32_t* words /* [kGridWords] */) {
    uint8_t g[kGridWords * 4];
    for (int i = 0; i < kGridWords * 4; ++i) g[i] = K_EMPTY;
    for (int i = 0; i < S; ++i) { g[i] = K_WALL; g[(S - 1) * S + i] = K_WALL; g[i * S] = K_WALL; g[i * S + S - 1] = K_WALL; }
    for (int i = 0; i < kGridWords; ++i)
        words[i] = (uint32_t)g[4 * i] | ((uint32_t)g[4 * i + 1] << 8) | ((uint32_t)g[4 * i + 2] << 16) | ((uint32_t)g[4 * i + 3] << 24);
}

// what a generation reads besides the configuration
struct GenIO {
    uint32_t* draws;           // this lane's generation scratch, word i at draws[i * stride]: kRing draws, then kObjWords
    int stride;                // 32 on the device (lane-interleaved shared memory), 1 on the host
    const uint32_t* tasks;     // [kTaskEntries][row_words] (build_task_table, then pack_task_table)
    int row_words = kTaskWords;  // row pitch of `tasks`: the longest row + terminator, rounded up to 4 words
    const uint32_t* prefix;    // [kTaskWords] (build_task_prefix)
    const uint32_t* empty;     // [kGridWords] fresh grid
    bool keep_marks = false;   // true: leave the next-to-a-door marks in the grid words; the caller strips them
                               //   (word & kMarkMask) while it copies the layout out
};
constexpr uint32_t kMarkMask = 0x7F7F7F7Fu;

// Builds the layout of `episode` into s: grid, agent, target, mission, carrying = 0, step_count = 0,
// episode = episode + 1, reset_draws; ORs ERR_TRIES into s.error.  mission_done / latch_step are
// not touched (they survive a reset in the reference).
// MODE: what is compiled in.  GEN_ALL: every problem; GEN_BASE: without the problems mov / full (see env_step_apply);
// GEN_MULTI_PLAIN: only the multi-room problem without obstacles and with the default door states (all_doors_open = false) -
// the configuration of every BASELINE benchmark - so that the single-room stages, the obstacle flags and the open-door
// draws are not in the rollout kernel's instruction stream.  The caller picks the instance from the configuration.
constexpr int GEN_ALL = 0, GEN_BASE = 1, GEN_MULTI_PLAIN = 2;
template <int MODE = GEN_ALL>
MGRL_HD void generate(EnvState& s, const EnvCfg& cfg, uint64_t seed, uint64_t env_id, uint32_t episode, const GenIO& io) {
    constexpr bool EXTRA = MODE == GEN_ALL, PLAIN = MODE == GEN_MULTI_PLAIN;
    const int S = cfg.size, m = S / 2;
    const bool multi = PLAIN || cfg.problem == P_MULTI;
    const bool doors_open = !PLAIN && cfg.all_doors_open != 0;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t e0 = (uint32_t)env_id, e1 = (uint32_t)(env_id >> 32);
    const int ds = io.stride;

    // ---- 1. the first draws of the episode (the ring is topped up inside the placement loop)
    int filled = 0;  // Philox blocks produced so far: draws [0, 4 * filled) exist, the last kRing of them are in the ring
    auto produce = [&]() {
        const PhiloxBlock b = philox_block((uint32_t)filled, episode, e0, e1, k0, k1);
        const int at = (4 * filled) & (kRing - 1);      // (a block never wraps: kRing is a multiple of 4)
        io.draws[at * ds] = b.w0; io.draws[(at + 1) * ds] = b.w1; io.draws[(at + 2) * ds] = b.w2; io.draws[(at + 3) * ds] = b.w3;
        ++filled;
    };
#pragma unroll 1
    for (int b = 0; b < kUpFront; ++b) produce();
    int nd = 0;  // draws consumed
    // top the ring up: a block whenever no more than kAhead draws are ahead (it never overwrites a draw that is still ahead:
    // 4 * filled + 4 - nd <= kRing).  Draws left in the ring when the layout is finished are wasted Philox rounds - a third
    // of a layout's instructions are Philox - so the ring is kept just deep enough for what one step of the prologue (<= 3
    // draws) or one placement iteration (<= 5, normally 2-4) consumes; a draw the ring does not hold is recomputed by draw().
    auto topup = [&]() { if (4 * filled - nd <= kAhead) produce(); };
    auto draw = [&](int i) -> uint32_t {  // word nd + i
        const int idx = nd + i;
        if (idx < 4 * filled) return io.draws[(idx & (kRing - 1)) * ds];
        return philox_word((uint32_t)idx, episode, e0, e1, k0, k1);
    };
    // placed objects, insertion order, 16 bits each behind the ring: type | colour << 3 | cell << 6
    int nobjs = 0;
    auto record = [&](uint32_t tc, int cell) {
        reinterpret_cast<uint16_t*>(io.draws + (kRing + (nobjs >> 1)) * ds)[nobjs & 1] = (uint16_t)(tc | ((uint32_t)cell << 6));
        ++nobjs;
    };

    // ---- fresh grid (Grid.wall_rect) and per-episode fields ([UPSTREAM] MiniGridEnv.reset, :125-127)
    uint32_t* gw = reinterpret_cast<uint32_t*>(&s);
#pragma unroll 4
    for (int i = 0; i < kGridWords; ++i) gw[i] = io.empty[i];
    s.carrying = 0; s.step_count = 0;
    s.target_x = s.target_y = kNone; s.target_action = 0;

    int cmd = multi ? cfg.mission
                    : (cfg.problem == P_GTO ? 0 : cfg.problem == P_GTG ? 5 : cfg.problem == P_OPN ? 1
                                                : cfg.problem == P_PKP ? 2 : cfg.problem == P_DRP ? 3 : 4);   // (full: drawn below)
    s.pad = 0;   // target_range = []
    int nrooms = 2;
    uint32_t doors = 0;          // per door: colour(3) | locked<<3 | key_in_box<<4
    uint32_t locked_mask = 0;
    uint32_t pool, pool_types;   // remaining (type, colour) pairs; 3 bits of type per 6-colour slot
    if (multi) { pool_types = T_KEY | (T_BALL << 3) | (T_BOX << 6); pool = (1u << 18) - 1u; }
    else if (cfg.problem == P_GTG) { pool_types = T_BOX | (T_DOOR << 3) | (T_KEY << 6) | (T_BALL << 9); pool = (1u << 24) - 1u; }
    else if (cfg.problem == P_OPN) { pool_types = T_BOX | (T_DOOR << 3); pool = (1u << 12) - 1u; }
    else if (cfg.problem == P_PKP) { pool_types = T_KEY | (T_BOX << 3) | (T_BALL << 6); pool = (1u << 18) - 1u; }
    else { pool_types = T_KEY | (T_BALL << 3) | (T_BOX << 6) | (T_DOOR << 9); pool = (1u << 24) - 1u; }

    // ---- 2. multi-room prologue (_generate_multi_map :601-611, doors :635-650, :880-929, :1324-1390)
    if (multi) {
        // (at most 2 + 4*3 + 4*2 = 22 draws here, at most 3 between two top-ups: always in the ring)
        auto pdraw = [&](int i) -> uint32_t { return io.draws[((nd + i) & (kRing - 1)) * ds]; };
        if (cmd < 0) { cmd = (int)((0x5210u >> (4 * mulhi32(pdraw(0), 4))) & 0xFu); ++nd; }  // choice([0,1,2,5])
        nrooms = 2 + (int)mulhi32(pdraw(0), 3); ++nd;                                         // randint(2,4)
        const int ndoors = nrooms == 2 ? 1 : nrooms;
        const int hi = nrooms == 3 ? m : S - 1;
        for (int i = 1; i < S - 1; ++i) {
            s.grid[i * S + m] = K_WALL;                                  // wall x = mid
            if (nrooms >= 3 && i < hi) s.grid[m * S + i] = K_WALL;      // wall y = mid
        }
        uint32_t colours = 0x3Fu;    // remaining door colours (sorted-name order)
#pragma unroll 1
        for (int d = 0; d < 4; ++d) {
            topup();
            if (d < ndoors) {
                const int i = (int)mulhi32(pdraw(0), (uint32_t)popc32(colours));
                const int bit = nth_set_bit(colours, i);
                colours &= ~(1u << bit);
                const int colour = sorted_colour(bit);
                int used = 1;
                const int locked = doors_open ? 0 : (mulhi32(pdraw(used++), 2) == 0);  // choice([True, False])
                const int kib = mulhi32(pdraw(used++), 2) == 0;
                nd += used;
                if (locked) {  // obj_choice.remove(('key', c)) [, ('box', c)]: pool slots are key, ball, box
                    pool &= ~(1u << (0 * 6 + sorted_index(colour)));
                    if (kib) pool &= ~(1u << (2 * 6 + sorted_index(colour)));
                }
                doors |= (uint32_t)(colour | (locked << 3) | (kib << 4)) << (8 * d);
                locked_mask |= (uint32_t)locked << d;
            }
        }
#pragma unroll 1
        for (int d = 0; d < 4; ++d) {
            topup();
            if (d < ndoors) {
                bool horizontal; int lo, hi2;
                if (nrooms == 2) { horizontal = false; lo = 1; hi2 = S - 2; }
                else if (nrooms == 3) { horizontal = d == 0; lo = d == 2 ? m + 1 : 1; hi2 = d == 2 ? S - 2 : m - 1; }
                else { horizontal = d < 2; lo = (d & 1) ? m + 1 : 1; hi2 = (d & 1) ? S - 2 : m - 1; }
                const int p = lo + (int)mulhi32(pdraw(0), (uint32_t)(hi2 - lo + 1)); ++nd;
                int is_open = 0;
                if (doors_open) { is_open = mulhi32(pdraw(0), 2) == 0; ++nd; }
                const int props = (int)((doors >> (8 * d)) & 0xFFu);
                const int colour = props & 7, state = is_open ? 0 : (((props >> 3) & 1) ? 2 : 1);
                const int x = horizontal ? p : m, y = horizontal ? m : p;
                const int c = y * S + x;
                s.grid[c] = (uint8_t)(K_DOOR + 8 * state + colour);
                s.grid[c - 1] |= kDoorFlag; s.grid[c + 1] |= kDoorFlag;
                s.grid[c - S] |= kDoorFlag; s.grid[c + S] |= kDoorFlag;
                record((uint32_t)(T_DOOR | (colour << 3)), c);
            }
        }
    }

    // ---- 3. placements: one try of the lane's current task per iteration
    // A task starts by fixing what it places: a pool entry (choice + remove: one draw), a door's key
    // or key-in-box, the goal, lava or a drawn obstacle kind.
    int kind = 0;           // kind byte the task writes
    uint32_t objw = 0;      // type | colour << 3 of the object record
    auto start_task = [&](uint32_t task, uint32_t word) {
        const int stage = (int)((task >> 16) & 7u);
        if ((!PLAIN && stage == G_OBJ) || stage == G_DIST) {
            const int i = (int)mulhi32(word, (uint32_t)popc32(pool)); ++nd;
            const int bit = nth_set_bit(pool, i);
            pool &= ~(1u << bit);
            const int slot = (bit >= 6) + (bit >= 12) + (bit >= 18);       // bit / 6 for bit < 24
            const int type = (int)((pool_types >> (3 * slot)) & 7u), colour = sorted_colour(bit - 6 * slot);
            kind = obj_kind(type, colour);
            objw = (uint32_t)(type | (colour << 3));
        } else if (stage == G_KEY) {
            const int props = (int)((doors >> (8 * ((task >> 19) & 3u))) & 0xFFu);
            const int colour = props & 7;
            const bool kib = (props >> 4) & 1;
            kind = kib ? K_BOX + 8 * (colour + 1) + colour : K_KEY + colour;
            objw = (uint32_t)((kib ? T_BOX : T_KEY) | (colour << 3));
        } else if (EXTRA && stage == G_FIXED) {
            const int i = (int)((task >> 19) & 31u), type = i / 6, colour = sorted_colour(i - 6 * type);
            kind = obj_kind(type, colour);
            objw = (uint32_t)(type | (colour << 3));
        } else if (!PLAIN && stage == G_OBST) {
            kind = K_LAVA;
            if (!multi) { kind = mulhi32(word, 2) == 0 ? K_LAVA : K_WALL; ++nd; }   // choice([Lava(), Wall()])
        } else {
            kind = K_GOAL; objw = T_GOAL;
        }
    };
    const uint32_t* list = io.prefix;
    int li = 0, tries = 0;
    uint32_t task = list[0];
    uint32_t agent_xy = 0xFFFFu, goal_xy = 0xFFFFu, key_xy = 0xFFFFu;   // x | y << 8
    if (!multi) start_task(task, draw(0));
    else { kind = K_GOAL; objw = T_GOAL; }

    while (task != 0u) {
        // the task that follows in the current list, fetched now so that an accepted placement does not wait for it
        const uint32_t task_after = list[li + 1];
        // Two tries of the current task per iteration.  A try is two draws (x, y); the four draws in hand are try A and,
        // should A be rejected, try B.  Both admissibility tests are independent (two grid bytes in flight, no state changes
        // on a rejection), so a rejected try costs no extra round trip through the draw buffer and the grid.  When A is
        // accepted, the third and fourth draw are what they were before: agent direction / next task's pool entry.
        topup();   // an iteration consumes at most five draws
        uint32_t d0, d1, d2, d3;
        if (nd + 4 <= 4 * filled) {
            d0 = io.draws[(nd & (kRing - 1)) * ds]; d1 = io.draws[((nd + 1) & (kRing - 1)) * ds];
            d2 = io.draws[((nd + 2) & (kRing - 1)) * ds]; d3 = io.draws[((nd + 3) & (kRing - 1)) * ds];
        } else {
            d0 = draw(0); d1 = draw(1); d2 = draw(2); d3 = draw(3);
        }
        const uint32_t rx0 = task & 15u, rw = (task >> 4) & 15u, ry0 = (task >> 8) & 15u, rh = (task >> 12) & 15u;
        const uint32_t mask = (task & TF_KEYMODE) ? kDoorFlag : (task & TF_AGENTMODE) ? 0x7Fu : 0xFFu;
        const int xa = (int)rx0 + (int)mulhi32(d0, rw), ya = (int)ry0 + (int)mulhi32(d1, rh);
        const int xb = (int)rx0 + (int)mulhi32(d2, rw), yb = (int)ry0 + (int)mulhi32(d3, rh);
        const uint32_t herea = s.grid[ya * S + xa], hereb = s.grid[yb * S + xb];   // kind | kDoorFlag (flag only on a multi map)
        const uint32_t xya = (uint32_t)xa | ((uint32_t)ya << 8), xyb = (uint32_t)xb | ((uint32_t)yb << 8);
        // admissibility as one predicate expression per try (no short-circuit branches)
        const bool lava_ok = !PLAIN && (task & TF_LAVA) != 0u, own = (task & TF_AGENT_CELL) != 0u, mid = !PLAIN && (task & TF_MID) != 0u;
        const bool bada = (((herea & mask) != 0u) & !(lava_ok & (herea == (uint32_t)K_LAVA))) | (own & (xya == agent_xy)) |
                          (xya == goal_xy) | (xya == key_xy) | (mid & ((xa == m) | (ya == m)));
        const bool badb = (((hereb & mask) != 0u) & !(lava_ok & (hereb == (uint32_t)K_LAVA))) | (own & (xyb == agent_xy)) |
                          (xyb == goal_xy) | (xyb == key_xy) | (mid & ((xb == m) | (yb == m)));
        // try B counts only if A was rejected and the rejection bound is not reached by A alone
        const bool second = bada & (tries + 1 < kMaxTries);
        int x = xa, y = ya;
        uint32_t xy = xya, here = herea;
        bool bad = bada;
        nd += 2;
        tries += bada;
        if (second) {
            x = xb; y = yb; xy = xyb; here = hereb; bad = badb;
            nd += 2;
            tries += badb;
        }
        if (tries >= kMaxTries) { s.error |= ERR_TRIES; bad = false; }
        if (!bad) {
            tries = 0;
            const int cell = y * S + x;
            const int stage = (int)((task >> 16) & 7u);
            // the draws that follow the accepted position
            uint32_t n0, n1;
            if (second) { n0 = draw(0); n1 = draw(1); } else { n0 = d2; n1 = d3; }
            uint32_t next_word = n0;
            uint32_t next_task = task_after;
            if (stage == G_AGENT) {
                agent_xy = xy;
                s.agent_dir = (uint8_t)mulhi32(n0, 4); ++nd;
                next_word = n1;
                if (EXTRA && cfg.problem == P_FULL) {   // np_random.choice(msn_commands) right after place_agent (:365)
                    cmd = (int)mulhi32(n1, 6); ++nd;
                    next_word = draw(0);
                }
                int row = 0;
                if (multi) {
                    const int gx = (int)(goal_xy & 0xFFu), gy = (int)(goal_xy >> 8);
                    row = task_index(nrooms, room_of(nrooms, m, x, y), room_of(nrooms, m, gx, gy), locked_mask);
                }
                list = io.tasks + (size_t)row * io.row_words;
                li = -1;
                next_task = list[0];
            } else {
                s.grid[cell] = (uint8_t)((uint32_t)kind | (here & kDoorFlag));
                if (PLAIN || stage != G_OBST) record(objw, cell);
                if (stage == G_GOAL) goal_xy = xy;
                if (stage == G_KEY && !(task & (1u << 21))) key_xy = xy;
            }
            ++li;
            task = next_task;
            start_task(task, next_word);
        }
    }
    s.agent_x = (uint8_t)(agent_xy & 0xFFu); s.agent_y = (uint8_t)(agent_xy >> 8);
    if (multi && !io.keep_marks) {  // drop the next-to-a-door marks
#pragma unroll 4
        for (int i = 0; i < kGridWords; ++i) gw[i] &= kMarkMask;
    }

    // ---- 4. target selection (:174-267)
    if (cmd <= 2) {
        const int n = nobjs;
        for (;;) {
            if (4 * filled - nd < 1) produce();   // (a draw the ring does not hold costs a whole block for one word)
            const int oi = (int)mulhi32(draw(0), (uint32_t)n); ++nd;
            const int o = reinterpret_cast<const uint16_t*>(io.draws + (kRing + (oi >> 1)) * ds)[oi & 1];
            const int t = o & 7;
            const bool ok = cmd == 0 ? t != T_GOAL : cmd == 1 ? (t == T_BOX || t == T_DOOR)
                                                              : (t == T_BOX || t == T_KEY || t == T_BALL);
            if (!ok && ++tries >= kMaxTries) s.error |= ERR_TRIES;
            if (ok || tries >= kMaxTries) {
                const int cell = o >> 6, ty = cell / S;
                s.mission_id = (uint8_t)(cmd * 24 + t * 6 + ((o >> 3) & 7));
                s.target_x = (uint8_t)(cell - ty * S); s.target_y = (uint8_t)ty;
                s.target_action = (uint8_t)(cmd == 0 ? A_DONE : cmd == 1 ? A_TOGGLE : A_PICKUP);
                break;
            }
        }
    } else if (!PLAIN && cmd == 3) {                                                      // :212-214
        s.mission_id = MISSION_DROP; s.target_action = A_DROP;
    } else if (EXTRA && cmd == 4) {                                                       // :216-256 'move <direction>'
        const int d = (int)mulhi32(draw(0), 4); ++nd;                                     // left right up down
        const uint32_t v = move_range(s.grid, S, d);
        s.mission_id = (uint8_t)(MISSION_MOVE0 + d);
        s.target_x = (uint8_t)v; s.target_y = (uint8_t)(v >> 8); s.target_action = (uint8_t)(v >> 16); s.pad = (uint8_t)(v >> 24);
    } else {                                                                              // :258-267
        s.mission_id = MISSION_GOAL; s.target_x = (uint8_t)(goal_xy & 0xFFu); s.target_y = (uint8_t)(goal_xy >> 8);
    }
    s.reset_draws = (uint16_t)nd;
    s.episode = episode + 1u;
}

// A finished environment takes over a prepared layout: everything generate() wrote, while the
// mission latch survives the reset (Q1) and the error byte is sticky.
MGRL_HD void adopt_layout(uint32_t* cur, const uint32_t* lay) {
    // words 0..30 grid + agent x/y/dir; 31 = carrying, step_count, target_x, target_y;
    // 32 = target_action, mission_id, mission_done, latch_step; 33 = episode; 34 = reset_draws, error, pad
#pragma unroll
    for (int i = 0; i < 32; ++i) cur[i] = lay[i];
    cur[32] = (lay[32] & 0x0000FFFFu) | (cur[32] & 0xFFFF0000u);
    cur[33] = lay[33];
    cur[34] = (lay[34] & 0xFF00FFFFu) | ((cur[34] | lay[34]) & 0x00FF0000u);   // (pad = top byte of a move mission's target range)
}

}  // namespace mgrl
