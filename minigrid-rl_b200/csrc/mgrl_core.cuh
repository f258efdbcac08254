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
constexpr int P_MULTI = 0, P_GTO = 1, P_GTG = 2, P_OPN = 3, P_PKP = 4, P_DRP = 5;
constexpr int MISSION_GOAL = 72, MISSION_DROP = 73;
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
#if defined(__CUDA_ARCH__)
    return (int)__fns(m, 0, n + 1);
#else
    for (int i = 0; i < n; ++i) m &= m - 1;
    return __builtin_ctz(m);
#endif
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
// episode is word d&3 of block d>>2; below(n) = mulhi32(word, n).  The generator computes the
// first kDrawBuf draws of an episode up front (16 blocks, straight-line, all lanes converged)
// into a lane-interleaved buffer and indexes it; draws beyond that (p < 0.5 %) are recomputed.
MGRL_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t ka, uint32_t kb, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ ka; c1 = l1; c2 = h0 ^ c3 ^ kb; c3 = l0;
        ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// --------------------------------------------------------------------------------- step
struct StepOut {
    float reward;
    uint8_t terminated, truncated, carry_obs;
};

// [UPSTREAM] MiniGridEnv.step, then PlaygroundEnv.step's mission bookkeeping.
// reward_lut[k] = float32(1 - 0.9*k/max_steps) computed in float64 on the host.
MGRL_HD StepOut env_step(EnvState& s, int action, int S, int max_steps, const float* reward_lut) {
    StepOut o;
    float r = 0.0f;
    bool term = false;
    const int dir = s.agent_dir;
    const int dx = (dir == 0) - (dir == 2), dy = (dir == 1) - (dir == 3);
    int ax = s.agent_x, ay = s.agent_y;
    const int step = s.step_count + 1;
    s.step_count = (uint8_t)step;
    const int fidx = (ay + dy) * S + ax + dx;
    const int k = s.grid[fidx];
    int carrying = s.carrying;
    int ndir = dir;

    if (action == A_LEFT) ndir = (dir + 3) & 3;
    else if (action == A_RIGHT) ndir = (dir + 1) & 3;
    else if (action == A_FORWARD) {
        if (k == K_EMPTY || k == K_GOAL || k == K_LAVA || (unsigned)(k - K_DOOR) < 8u) { ax += dx; ay += dy; }
        if (k == K_GOAL) { term = true; r = reward_lut[step]; }
        if (k == K_LAVA) term = true;
    } else if (action == A_PICKUP) {
        if (k_pickable(k) && carrying == 0) { carrying = k; s.grid[fidx] = K_EMPTY; }
    } else if (action == A_DROP) {
        if (k == K_EMPTY && carrying != 0) { s.grid[fidx] = (uint8_t)carrying; carrying = 0; }
    } else if (action == A_TOGGLE) {
        if (k_is_door(k)) {
            const int st = (k - K_DOOR) >> 3, c = k & 7;
            if (st == 2) {  // locked: opens only for a Key of its colour
                if (k_is_key(carrying) && (carrying & 7) == c) s.grid[fidx] = (uint8_t)(K_DOOR + c);
            } else {
                s.grid[fidx] = (uint8_t)(K_DOOR + 8 * (st ^ 1) + c);
            }
        } else if (k_is_box(k)) {  // the box is replaced by its contents
            const int m = (k - K_BOX) >> 3;
            s.grid[fidx] = (uint8_t)(m ? K_KEY + m - 1 : K_EMPTY);
        }
    } else if (action != A_DONE) {
        s.error |= ERR_BAD_ACTION;  // upstream raises ValueError
    }
    s.agent_x = (uint8_t)ax; s.agent_y = (uint8_t)ay; s.agent_dir = (uint8_t)ndir;
    o.truncated = step >= max_steps;
    o.carry_obs = (uint8_t)carrying;  // the observation is rendered here (custom_env.py:270)

    if (term) {  // custom_env.py:272-277
        if (s.mission_id != MISSION_GOAL) { s.mission_done = 0; s.latch_step = 0; r = 0.0f; }
    } else {
        const int ndx = (ndir == 0) - (ndir == 2), ndy = (ndir == 1) - (ndir == 3);
        const int fx = ax + ndx, fy = ay + ndy;
        if (action == A_TOGGLE) {  // :279-283 colour match only
            const int f = s.grid[fy * S + fx];
            if (k_is_door(f) && carrying != 0 && (f & 7) == (carrying & 7)) carrying = 0;
        }
        if (!s.mission_done) {  // :288-317
            const int ta = s.target_action;
            bool latch = false;
            if (s.target_x != kNone) {
                if (ta) latch = (fx == s.target_x && fy == s.target_y && action == ta);
                else latch = (ax == s.target_x && ay == s.target_y);
            } else {
                latch = (ta != 0 && action == ta);
            }
            if (latch) { s.mission_done = 1; s.latch_step = (uint8_t)step; }
        }
        if (action == A_DONE) {  // :319-328
            r = s.mission_done ? reward_lut[s.latch_step] : 0.0f;
            s.mission_done = 0; s.latch_step = 0;
            term = true;
        }
    }
    s.carrying = (uint8_t)carrying;
    o.reward = r;
    o.terminated = term;
    return o;
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

// fast path: see-through view, HWC, 37 aligned words (148 B) per environment
MGRL_HD void encode_view_packed(const EnvState& s, int carrying, int S, const uint32_t* lut, uint32_t* out) {
    ViewMap m;
    view_map(s, S, m);
    uint32_t e[4];
#pragma unroll
    for (int c = 0; c < kView * kView; ++c) {
        const int vx = c / kView, vy = c % kView;
        const int k = (c == 3 * kView + 6) ? carrying : (int)s.grid[m.ro[vx] + m.fo[vy]];
        e[c & 3] = lut[k];
        if ((c & 3) == 3) {
            const int w = (c >> 2) * 3;
            out[w] = pack3(e[0], e[1], 0);
            out[w + 1] = pack3(e[1], e[2], 1);
            out[w + 2] = pack3(e[2], e[3], 2);
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
//      obstacle) is one shared draw-test-commit body; a lane's sequence of key/distractor tasks
//      is a byte string looked up in a table indexed by (rooms, agent room, goal room, locked
//      doors), so advancing to the next task is a byte fetch instead of a nested room walk;
//   4. cells next to a door carry a flag bit while the layout is built, so "empty and not
//      next to a door" (next2door, :2036-2046) is one compare of the byte already loaded.
// The draw order is the reference's; the CPU oracle consumes the same stream.
constexpr int T_KEY = 0, T_BALL = 1, T_BOX = 2, T_DOOR = 3, T_GOAL = 4;
constexpr int kDrawBuf = 64;            // draws precomputed per generation = 16 Philox blocks
constexpr int kGridWords = 31;          // words 0..30 of EnvState cover grid[121] + agent x/y/dir
constexpr int kTaskBytes = 32;          // task string: [0] = n, [1..n] = tasks
constexpr int kTaskEntries = 3 * 4 * 4 * 16;
constexpr uint32_t kDoorFlag = 0x80u;   // "next to a door" mark on a grid byte (kinds are < 128)

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

// task byte: kind (bits 0-1: 1 = key of a locked door, 2 = distractor) | room << 2 | door << 4 | second << 6
constexpr int TASK_KEY = 1, TASK_DIST = 2;
MGRL_HD int task_index(int nrooms, int agent_room, int goal_room, uint32_t locked) {
    return (((nrooms - 2) * 4 + agent_room) * 4 + goal_room) * 16 + (int)locked;
}

// The per-room walk of _generate_{2,3,4}_rooms (:652-855, :931-1297, :1392-2034) for one
// (rooms, agent room, goal room, locked doors) combination: keys of the room's locked doors first
// (each uses up one of the room's distractor slots), one slot less in the goal's room, then the
// room's distractors.  Reference quirk (:1119, :1660): the lower-left loop reads the upper-left counter.
inline void build_task_string(int nrooms, int agent_room, int goal_room, uint32_t locked, int num_objects, uint8_t* out) {
    const int nl = num_objects / 2, nr = num_objects - nl;
    int cnt[4] = {0, 0, 0, 0};
    if (nrooms == 2) { cnt[0] = nl; cnt[1] = nr; }
    else if (nrooms == 3) { cnt[0] = nl / 2; cnt[1] = nl - nl / 2; cnt[2] = nr; }
    else { cnt[0] = nl / 2; cnt[1] = nl - nl / 2; cnt[2] = nr / 2; cnt[3] = nr - nr / 2; }
    int n = 0;
    for (int i = 0; i < kTaskBytes; ++i) out[i] = 0;
    if (agent_room >= nrooms || goal_room >= nrooms) return;
    for (int r = 0; r < nrooms; ++r) {
        for (int j = 0; j < 2; ++j) {
            const int d = key_door(nrooms, r, agent_room, j);
            if (d != 7 && ((locked >> d) & 1u)) {
                if (n < kTaskBytes - 1) out[1 + n++] = (uint8_t)(TASK_KEY | (r << 2) | (d << 4) | (j << 6));
                --cnt[r];
            }
        }
        if (goal_room == r) --cnt[r];
        const int loops = (nrooms >= 3 && r == 1) ? cnt[0] : cnt[r];
        for (int q = 0; q < loops; ++q)
            if (n < kTaskBytes - 1) out[1 + n++] = (uint8_t)(TASK_DIST | (r << 2));
    }
    out[0] = (uint8_t)n;
}
inline void build_task_table(int num_objects, uint8_t* table /* [kTaskEntries * kTaskBytes] */) {
    for (int nrooms = 2; nrooms <= 4; ++nrooms)
        for (int a = 0; a < 4; ++a)
            for (int g = 0; g < 4; ++g)
                for (uint32_t l = 0; l < 16; ++l)
                    build_task_string(nrooms, a, g, l, num_objects, table + (size_t)task_index(nrooms, a, g, l) * kTaskBytes);
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
    uint32_t* draws;           // this lane's draw buffer: word i at draws[i * stride]
    int stride;                // 32 on the device (lane-interleaved shared memory), 1 on the host
    const uint8_t* tasks;      // [kTaskEntries][kTaskBytes] task strings of cfg.num_objects
    const uint32_t* empty;     // [kGridWords] fresh grid
};

enum GenStage : int { G_OBJ = 0, G_GOAL, G_AGENT, G_KEY, G_DIST, G_OBST, G_PLACED };

// Builds the layout of `episode` into s: grid, agent, target, mission, carrying = 0, step_count = 0,
// episode = episode + 1, reset_draws; ORs ERR_TRIES into s.error.  mission_done / latch_step are
// not touched (they survive a reset in the reference).
MGRL_HD void generate(EnvState& s, const EnvCfg& cfg, uint64_t seed, uint64_t env_id, uint32_t episode, const GenIO& io) {
    const int S = cfg.size, m = S / 2;
    const bool multi = cfg.problem == P_MULTI;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t e0 = (uint32_t)env_id, e1 = (uint32_t)(env_id >> 32);
    const int ds = io.stride;

    // ---- 1. all draws of the episode
    for (int b = 0; b < kDrawBuf / 4; ++b) {
        uint32_t w[4];
        philox4x32_10((uint32_t)b, episode, e0, e1, k0, k1, w);
#pragma unroll
        for (int j = 0; j < 4; ++j) io.draws[(4 * b + j) * ds] = w[j];
    }
    int nd = 0;  // draws consumed
    auto draw = [&](int i) -> uint32_t {  // word nd + i
        const int idx = nd + i;
        if (idx < kDrawBuf) return io.draws[idx * ds];
        uint32_t w[4];
        philox4x32_10((uint32_t)(idx >> 2), episode, e0, e1, k0, k1, w);
        const int j = idx & 3;
        return j == 0 ? w[0] : j == 1 ? w[1] : j == 2 ? w[2] : w[3];
    };
    // placed objects, insertion order: type | colour<<3 | x<<6 | y<<10.  Entry i lives in draw slot i,
    // which is dead by then (every object consumes at least one draw before it is recorded).
    int nobjs = 0;
    auto push_obj = [&](int type, int colour, int x, int y) {
        io.draws[nobjs * ds] = (uint32_t)(type | (colour << 3) | (x << 6) | (y << 10));
        ++nobjs;
    };

    // ---- fresh grid (Grid.wall_rect) and per-episode fields ([UPSTREAM] MiniGridEnv.reset, :125-127)
    uint32_t* gw = reinterpret_cast<uint32_t*>(&s);
#pragma unroll
    for (int i = 0; i < kGridWords; ++i) gw[i] = io.empty[i];
    s.carrying = 0; s.step_count = 0;
    s.target_x = s.target_y = kNone; s.target_action = 0;

    int cmd = multi ? cfg.mission
                    : (cfg.problem == P_GTO ? 0 : cfg.problem == P_GTG ? 5 : cfg.problem == P_OPN ? 1
                                                : cfg.problem == P_PKP ? 2 : 3);
    int nrooms = 0, ndoors = 0;
    uint32_t doors = 0;          // per door: colour(3) | locked<<3 | key_in_box<<4
    uint32_t locked_mask = 0;
    uint32_t colours = 0x3Fu;    // remaining door colours (sorted-name order)
    uint32_t pool, pool_types;   // remaining (type, colour) pairs; 3 bits of type per 6-colour slot
    if (multi) { pool_types = T_KEY | (T_BALL << 3) | (T_BOX << 6); pool = (1u << 18) - 1u; }
    else if (cfg.problem == P_GTG) { pool_types = T_BOX | (T_DOOR << 3) | (T_KEY << 6) | (T_BALL << 9); pool = (1u << 24) - 1u; }
    else if (cfg.problem == P_OPN) { pool_types = T_BOX | (T_DOOR << 3); pool = (1u << 12) - 1u; }
    else if (cfg.problem == P_PKP) { pool_types = T_KEY | (T_BOX << 3) | (T_BALL << 6); pool = (1u << 18) - 1u; }
    else { pool_types = T_KEY | (T_BALL << 3) | (T_BOX << 6) | (T_DOOR << 9); pool = (1u << 24) - 1u; }

    // ---- 2. multi-room prologue (_generate_multi_map :601-611, doors :635-650, :880-929, :1324-1390)
    if (multi) {
        if (cmd < 0) { cmd = (int)((0x5210u >> (4 * mulhi32(draw(0), 4))) & 0xFu); ++nd; }  // choice([0,1,2,5])
        nrooms = 2 + (int)mulhi32(draw(0), 3); ++nd;                                         // randint(2,4)
        ndoors = nrooms == 2 ? 1 : nrooms;
        const int hi = nrooms == 3 ? m : S - 1;
        for (int i = 1; i < S - 1; ++i) {
            s.grid[i * S + m] = K_WALL;                                  // wall x = mid
            if (nrooms >= 3 && i < hi) s.grid[m * S + i] = K_WALL;      // wall y = mid
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (d < ndoors) {
                const int i = (int)mulhi32(draw(0), (uint32_t)popc32(colours));
                const int bit = nth_set_bit(colours, i);
                colours &= ~(1u << bit);
                const int colour = sorted_colour(bit);
                int used = 1;
                const int locked = cfg.all_doors_open ? 0 : (mulhi32(draw(used++), 2) == 0);  // choice([True, False])
                const int kib = mulhi32(draw(used++), 2) == 0;
                nd += used;
                if (locked) {  // obj_choice.remove(('key', c)) [, ('box', c)]: pool slots are key, ball, box
                    pool &= ~(1u << (0 * 6 + sorted_index(colour)));
                    if (kib) pool &= ~(1u << (2 * 6 + sorted_index(colour)));
                }
                doors |= (uint32_t)(colour | (locked << 3) | (kib << 4)) << (8 * d);
                locked_mask |= (uint32_t)locked << d;
            }
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (d < ndoors) {
                bool horizontal; int lo, hi2;
                if (nrooms == 2) { horizontal = false; lo = 1; hi2 = S - 2; }
                else if (nrooms == 3) { horizontal = d == 0; lo = d == 2 ? m + 1 : 1; hi2 = d == 2 ? S - 2 : m - 1; }
                else { horizontal = d < 2; lo = (d & 1) ? m + 1 : 1; hi2 = (d & 1) ? S - 2 : m - 1; }
                const int p = lo + (int)mulhi32(draw(0), (uint32_t)(hi2 - lo + 1)); ++nd;
                int is_open = 0;
                if (cfg.all_doors_open) { is_open = mulhi32(draw(0), 2) == 0; ++nd; }
                const int props = (int)((doors >> (8 * d)) & 0xFFu);
                const int colour = props & 7, state = is_open ? 0 : (((props >> 3) & 1) ? 2 : 1);
                const int x = horizontal ? p : m, y = horizontal ? m : p;
                const int c = y * S + x;
                s.grid[c] = (uint8_t)(K_DOOR + 8 * state + colour);
                s.grid[c - 1] |= kDoorFlag; s.grid[c + 1] |= kDoorFlag;
                s.grid[c - S] |= kDoorFlag; s.grid[c + S] |= kDoorFlag;
                push_obj(T_DOOR, colour, x, y);
            }
        }
    }

    // ---- 3. placements: one try of the lane's current task per iteration
    const bool has_goal = multi || cfg.problem == P_GTG || cfg.problem == P_DRP;
    int stage = multi ? G_GOAL : (cfg.num_objects > 0 ? G_OBJ : (has_goal ? G_GOAL : G_AGENT));
    int agent_x = -1, agent_y = -1, goal_x = -1, goal_y = -1, agent_room = 0;
    int r = 0, d = 0, second = 0, kx = -1, ky = -1;
    int obj_i = 0, obst_i = 0, tries = 0, ti = 0, nt = 0;
    int pend_type = 0, pend_colour = 0;
    bool pre_done = false;        // the task's pool / kind draw has happened
    const uint8_t* tasks = io.tasks;

    while (stage != G_PLACED) {
        const bool whole_grid = stage == G_OBJ || stage == G_GOAL || stage == G_AGENT || (stage == G_OBST && !multi);
        // pre-draw once per task: pool entry (choice + remove) or obstacle kind
        if (!pre_done && (stage == G_OBJ || stage == G_DIST)) {
            const int i = (int)mulhi32(draw(0), (uint32_t)popc32(pool)); ++nd;
            const int bit = nth_set_bit(pool, i);
            pool &= ~(1u << bit);
            pend_type = (int)((pool_types >> (3 * (bit / 6))) & 7u);
            pend_colour = sorted_colour(bit % 6);
            pre_done = true;
        } else if (!pre_done && stage == G_OBST && !multi) {
            pend_type = mulhi32(draw(0), 2) == 0 ? K_LAVA : K_WALL; ++nd;  // choice([Lava(), Wall()])
            pre_done = true;
        }
        // room rectangle (inclusive) or the whole grid (place_obj draws over [0,S))
        int x0, x1, y0, y1;
        if (whole_grid) { x0 = 0; x1 = S - 1; y0 = 0; y1 = S - 1; }
        else if (stage == G_OBST) { x0 = 1; x1 = S - 2; y0 = 1; y1 = S - 2; }
        else {
            const bool left = nrooms == 2 ? r == 0 : r < 2;
            const bool full_height = nrooms == 2 || (nrooms == 3 && r == 2);
            const bool upper = (r & 1) == 0;
            x0 = left ? 1 : m + 1; x1 = left ? m - 1 : S - 2;
            y0 = (full_height || upper) ? 1 : m + 1;
            y1 = (full_height || !upper) ? S - 2 : m - 1;
        }
        const int x = x0 + (int)mulhi32(draw(0), (uint32_t)(x1 - x0 + 1));
        const int y = y0 + (int)mulhi32(draw(1), (uint32_t)(y1 - y0 + 1));
        nd += 2;
        const int cell = y * S + x;
        const int here = s.grid[cell];           // kind | kDoorFlag (flag only ever set on a multi map)
        const bool at_agent = x == agent_x && y == agent_y;
        bool ok;
        if (stage == G_AGENT) ok = (here & 0x7F) == K_EMPTY;                        // place_agent: no next2door test
        else if (whole_grid || stage == G_DIST) ok = here == K_EMPTY && !at_agent;  // place_obj (+ not next2door)
        else if (stage == G_KEY)
            ok = !(x == goal_x && y == goal_y) && !(r == agent_room && at_agent) && !(x == kx && y == ky) && !(here & kDoorFlag);
        else ok = x != m && y != m && (here == K_EMPTY || here == K_LAVA) && !at_agent;  // lava on a multi map
        if (!ok && ++tries >= kMaxTries) { s.error |= ERR_TRIES; ok = true; }
        if (ok) {
            tries = 0; pre_done = false;
            bool next_task = false;
            if (stage == G_OBJ) {
                s.grid[cell] = (uint8_t)obj_kind(pend_type, pend_colour);
                push_obj(pend_type, pend_colour, x, y);
                if (++obj_i == cfg.num_objects) stage = has_goal ? G_GOAL : G_AGENT;
            } else if (stage == G_GOAL) {
                s.grid[cell] = (uint8_t)(K_GOAL | (here & kDoorFlag));
                goal_x = x; goal_y = y;
                push_obj(T_GOAL, 0, x, y);
                stage = G_AGENT;
            } else if (stage == G_AGENT) {
                agent_x = x; agent_y = y;
                s.agent_x = (uint8_t)x; s.agent_y = (uint8_t)y;
                s.agent_dir = (uint8_t)mulhi32(draw(0), 4); ++nd;
                if (multi) {
                    agent_room = room_of(nrooms, m, agent_x, agent_y);
                    const int goal_room = room_of(nrooms, m, goal_x, goal_y);
                    tasks = io.tasks + (size_t)task_index(nrooms, agent_room, goal_room, locked_mask) * kTaskBytes;
                    nt = tasks[0]; ti = 0;
                    next_task = true;
                } else {
                    stage = G_OBST;
                }
            } else if (stage == G_KEY) {
                const int props = (int)((doors >> (8 * d)) & 0xFFu);
                const int colour = props & 7;
                const bool kib = (props >> 4) & 1;
                s.grid[cell] = (uint8_t)(kib ? K_BOX + 8 * (colour + 1) + colour : K_KEY + colour);
                push_obj(kib ? T_BOX : T_KEY, colour, x, y);
                if (!second) { kx = x; ky = y; }
                next_task = true;
            } else if (stage == G_DIST) {
                s.grid[cell] = (uint8_t)obj_kind(pend_type, pend_colour);
                push_obj(pend_type, pend_colour, x, y);
                next_task = true;
            } else {  // G_OBST
                s.grid[cell] = (uint8_t)(multi ? K_LAVA : pend_type);
                ++obst_i;
            }
            if (next_task) {  // next key / distractor of the room walk, then the obstacles
                if (ti < nt) {
                    const int t = tasks[1 + ti++];
                    stage = (t & 3) == TASK_KEY ? G_KEY : G_DIST;
                    r = (t >> 2) & 3; d = (t >> 4) & 3; second = (t >> 6) & 1;
                } else {
                    stage = G_OBST;
                }
            }
            if (stage == G_OBST && obst_i >= cfg.num_obstacles) stage = G_PLACED;
        }
    }
    if (multi) {  // drop the next-to-a-door marks
#pragma unroll
        for (int i = 0; i < kGridWords; ++i) gw[i] &= 0x7F7F7F7Fu;
    }

    // ---- 4. target selection (:174-267)
    if (cmd <= 2) {
        const int n = nobjs;
        for (;;) {
            const int o = (int)io.draws[(int)mulhi32(draw(0), (uint32_t)n) * ds]; ++nd;
            const int t = o & 7;
            const bool ok = cmd == 0 ? t != T_GOAL : cmd == 1 ? (t == T_BOX || t == T_DOOR)
                                                              : (t == T_BOX || t == T_KEY || t == T_BALL);
            if (!ok && ++tries >= kMaxTries) s.error |= ERR_TRIES;
            if (ok || tries >= kMaxTries) {
                s.mission_id = (uint8_t)(cmd * 24 + t * 6 + ((o >> 3) & 7));
                s.target_x = (uint8_t)((o >> 6) & 15); s.target_y = (uint8_t)((o >> 10) & 15);
                s.target_action = (uint8_t)(cmd == 0 ? A_DONE : cmd == 1 ? A_TOGGLE : A_PICKUP);
                break;
            }
        }
    } else if (cmd == 3) {                                                                // :212-214
        s.mission_id = MISSION_DROP; s.target_action = A_DROP;
    } else {                                                                              // :258-267
        s.mission_id = MISSION_GOAL; s.target_x = (uint8_t)goal_x; s.target_y = (uint8_t)goal_y;
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
    cur[34] = (lay[34] & 0x0000FFFFu) | ((cur[34] | lay[34]) & 0x00FF0000u);
}

}  // namespace mgrl
