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
constexpr int ERR_BAD_ACTION = 1, ERR_TRIES = 2;
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
// episode is word d&3 of block d>>2; below(n) = mulhi32(word, n).
//
// The generator consumes draws through a 4-word *window*: sync() makes blocks A = d>>2 and
// B = A+1 current (one Philox site per generator iteration, so lanes of a warp that sit at
// different draw counts still execute the rounds together), peek(i) returns word d+i for
// i < 4, and the caller then advances ndraw by what it used.
struct Rng {
    uint32_t k0, k1, e0, e1, episode;
    uint32_t ndraw;
    uint32_t blk;  // block id held in a[]; b[] holds blk+1
    uint32_t a[4], b[4];

    MGRL_HD static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t ka, uint32_t kb,
                               uint32_t* out) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
            const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
            c0 = h1 ^ c1 ^ ka; c1 = l1; c2 = h0 ^ c3 ^ kb; c3 = l0;
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
    MGRL_HD void init(uint64_t seed, uint64_t env_id, uint32_t ep) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        e0 = (uint32_t)env_id; e1 = (uint32_t)(env_id >> 32);
        episode = ep; ndraw = 0; blk = 0;
        philox(0, episode, e0, e1, k0, k1, a);
        philox(1, episode, e0, e1, k0, k1, b);
    }
    // re-establish a[] = block(ndraw>>2), b[] = the block after it
    MGRL_HD void sync() {
        if ((ndraw >> 2) != blk) {  // the window moved by exactly one block (<= 4 draws per iteration)
            blk = ndraw >> 2;
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = b[i];
            philox(blk + 1, episode, e0, e1, k0, k1, b);
        }
    }
    // word ndraw+i, i in 0..3
    MGRL_HD uint32_t peek(int i) const {
        const int w = (int)(ndraw & 3u) + i;  // 0..6
        return w == 0 ? a[0] : w == 1 ? a[1] : w == 2 ? a[2] : w == 3 ? a[3] : w == 4 ? b[0] : w == 5 ? b[1] : b[2];
    }
};

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
// room, :595-2034 two/three/four rooms) as ONE flattened task machine.
//
// The reference is a long sequence of rejection loops ("draw a position, retry until it is
// admissible") whose number and kind depend on earlier draws.  Run lane-per-environment that
// diverges badly, so the generator is written as a single loop whose body performs exactly
// one *try* of the lane's current task: every placement kind (goal, agent, key / key-in-box,
// distractor, single-room object, obstacle) shares the same draw-test-commit code, driven by
// per-lane flags.  Lanes that need more tries simply take more iterations while the others
// move on to their next task, and the Philox rounds run at one site per iteration.
// The draw order is the reference's; the CPU oracle consumes the same stream.
constexpr int T_KEY = 0, T_BALL = 1, T_BOX = 2, T_DOOR = 3, T_GOAL = 4;
constexpr int kMaxObjs = 40;

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

enum GenStage : int {
    G_CMD = 0,   // multi: mission command (if cfg.mission is null) and room count
    G_DOORPROP,  // door colour / locked / key_in_box
    G_DOORCELL,  // door position (/ is_open)
    G_OBJ,       // single-room object: pool draw + place_obj
    G_GOAL,      // goal (multi: not next to a door)
    G_AGENT,     // place_agent
    G_KEY,       // key or box-with-key of a locked door
    G_DIST,      // distractor object of a room
    G_OBST,      // obstacle
    G_TARGET,    // target selection
    G_DONE
};

MGRL_HD void generate(EnvState& s, const EnvCfg& cfg, uint64_t seed, uint64_t env_id) {
    const int S = cfg.size, m = S / 2;
    const bool multi = cfg.problem == P_MULTI;
    // ---- grid: empty interior, wall border (Grid.wall_rect, custom_env.py:132)
    for (int i = 0; i < kGridCells; ++i) s.grid[i] = K_EMPTY;
    for (int i = 0; i < S; ++i) {
        s.grid[i] = K_WALL; s.grid[(S - 1) * S + i] = K_WALL; s.grid[i * S] = K_WALL; s.grid[i * S + S - 1] = K_WALL;
    }
    s.carrying = 0; s.step_count = 0;                                  // [UPSTREAM] MiniGridEnv.reset
    s.target_x = s.target_y = kNone; s.target_action = 0;              // :125-127

    Rng rng;
    rng.init(seed, env_id, s.episode);

    // ---- per-lane task state
    int stage = multi ? G_CMD : G_OBJ;
    int cmd = multi ? cfg.mission
                    : (cfg.problem == P_GTO ? 0 : cfg.problem == P_GTG ? 5 : cfg.problem == P_OPN ? 1
                                                : cfg.problem == P_PKP ? 2 : 3);
    int nrooms = 0, ndoors = 0, d = 0;          // door loop
    uint32_t doors = 0;                          // per door: colour(3) | locked<<3 | key_in_box<<4
    uint32_t colours = 0x3Fu;                    // remaining door colours (sorted-name order)
    uint32_t pool = 0, pool_types = 0;           // remaining (type, colour) pairs; 3 bits of type per slot
    int agent_x = -1, agent_y = -1, goal_x = -1, goal_y = -1, agent_room = 0, goal_room = 0;
    int r = 0, ph = 0, q = 0, loops = 0, kx = -1, ky = -1;   // room walk
    uint32_t counts = 0;                         // 4 signed bytes: distractor budget per room
    int obj_i = 0, obst_i = 0, tries = 0;
    int pend_type = 0, pend_colour = 0;          // object being placed
    bool pre_done = false;                       // the task's pool / kind draw has happened
    int nobjs = 0;
    uint16_t objs[kMaxObjs];                     // type | colour<<3 | x<<6 | y<<10, insertion order

    // pool of (type, colour) pairs in the reference's comprehension order
    if (multi) { pool_types = T_KEY | (T_BALL << 3) | (T_BOX << 6); pool = (1u << 18) - 1u; }
    else if (cfg.problem == P_GTG) { pool_types = T_BOX | (T_DOOR << 3) | (T_KEY << 6) | (T_BALL << 9); pool = (1u << 24) - 1u; }
    else if (cfg.problem == P_OPN) { pool_types = T_BOX | (T_DOOR << 3); pool = (1u << 12) - 1u; }
    else if (cfg.problem == P_PKP) { pool_types = T_KEY | (T_BOX << 3) | (T_BALL << 6); pool = (1u << 18) - 1u; }
    else { pool_types = T_KEY | (T_BALL << 3) | (T_BOX << 6) | (T_DOOR << 9); pool = (1u << 24) - 1u; }
    if (!multi && cfg.num_objects == 0) stage = (cfg.problem == P_GTG || cfg.problem == P_DRP) ? G_GOAL : G_AGENT;

#define MGRL_CNT(rr) ((int)(int8_t)((counts >> (8 * (rr))) & 0xFFu))
#define MGRL_CNT_DEC(rr) counts = (counts & ~(0xFFu << (8 * (rr)))) | ((uint32_t)((MGRL_CNT(rr) - 1) & 0xFF) << (8 * (rr)))

    while (stage != G_DONE) {
        rng.sync();
        int used = 0;

        if (stage == G_CMD) {  // _generate_multi_map :601-611
            if (cmd < 0) cmd = (int)((0x5210u >> (4 * mulhi32(rng.peek(used++), 4))) & 0xFu);  // choice([0,1,2,5])
            nrooms = 2 + (int)mulhi32(rng.peek(used++), 3);                                     // randint(2,4)
            ndoors = nrooms == 2 ? 1 : nrooms;
            for (int i = 1; i < S - 1; ++i) s.grid[i * S + m] = K_WALL;                         // wall x = mid
            if (nrooms >= 3) {
                const int hi = nrooms == 3 ? m : S - 1;
                for (int i = 1; i < hi; ++i) s.grid[m * S + i] = K_WALL;                        // wall y = mid
            }
            const int n = cfg.num_objects, nl = n / 2, nr = n - nl;
            int c0, c1, c2 = 0, c3 = 0;
            if (nrooms == 2) { c0 = nl; c1 = nr; }
            else if (nrooms == 3) { c0 = nl / 2; c1 = nl - nl / 2; c2 = nr; }
            else { c0 = nl / 2; c1 = nl - nl / 2; c2 = nr / 2; c3 = nr - nr / 2; }
            counts = (uint32_t)(c0 & 0xFF) | ((uint32_t)(c1 & 0xFF) << 8) | ((uint32_t)(c2 & 0xFF) << 16) |
                     ((uint32_t)(c3 & 0xFF) << 24);
            stage = G_DOORPROP; d = 0;
        } else if (stage == G_DOORPROP) {  // :635-643, :880-908, :1324-1362
            const int i = (int)mulhi32(rng.peek(used++), (uint32_t)popc32(colours));
            const int bit = nth_set_bit(colours, i);
            colours &= ~(1u << bit);
            const int colour = sorted_colour(bit);
            const int locked = cfg.all_doors_open ? 0 : (mulhi32(rng.peek(used++), 2) == 0);  // choice([True, False])
            const int kib = mulhi32(rng.peek(used++), 2) == 0;
            if (locked) {  // obj_choice.remove(('key', c)) [, ('box', c)]: pool slots are key, ball, box
                pool &= ~(1u << (0 * 6 + sorted_index(colour)));
                if (kib) pool &= ~(1u << (2 * 6 + sorted_index(colour)));
            }
            doors |= (uint32_t)(colour | (locked << 3) | (kib << 4)) << (8 * d);
            if (++d == ndoors) { stage = G_DOORCELL; d = 0; }
        } else if (stage == G_DOORCELL) {  // :646-650, :911-929, :1365-1390
            bool horizontal; int lo, hi;
            if (nrooms == 2) { horizontal = false; lo = 1; hi = S - 2; }
            else if (nrooms == 3) { horizontal = d == 0; lo = d == 2 ? m + 1 : 1; hi = d == 2 ? S - 2 : m - 1; }
            else { horizontal = d < 2; lo = (d & 1) ? m + 1 : 1; hi = (d & 1) ? S - 2 : m - 1; }
            const int p = lo + (int)mulhi32(rng.peek(used++), (uint32_t)(hi - lo + 1));
            const int is_open = cfg.all_doors_open ? (mulhi32(rng.peek(used++), 2) == 0) : 0;
            const int props = (int)((doors >> (8 * d)) & 0xFFu);
            const int colour = props & 7, state = is_open ? 0 : (((props >> 3) & 1) ? 2 : 1);
            const int x = horizontal ? p : m, y = horizontal ? m : p;
            s.grid[y * S + x] = (uint8_t)(K_DOOR + 8 * state + colour);
            if (nobjs < kMaxObjs) objs[nobjs] = (uint16_t)(T_DOOR | (colour << 3) | (x << 6) | (y << 10));
            ++nobjs;
            if (++d == ndoors) stage = G_GOAL;
        } else if (stage == G_TARGET) {  // :174-210
            const int n = nobjs < kMaxObjs ? nobjs : kMaxObjs;
            const int o = objs[mulhi32(rng.peek(used++), (uint32_t)n)];
            const int t = o & 7;
            const bool ok = cmd == 0 ? t != T_GOAL : cmd == 1 ? (t == T_BOX || t == T_DOOR)
                                                              : (t == T_BOX || t == T_KEY || t == T_BALL);
            if (!ok && ++tries >= kMaxTries) s.error |= ERR_TRIES;
            if (ok || tries >= kMaxTries) {
                s.mission_id = (uint8_t)(cmd * 24 + t * 6 + ((o >> 3) & 7));
                s.target_x = (uint8_t)((o >> 6) & 15); s.target_y = (uint8_t)((o >> 10) & 15);
                s.target_action = (uint8_t)(cmd == 0 ? A_DONE : cmd == 1 ? A_TOGGLE : A_PICKUP);
                stage = G_DONE;
            }
        } else {
            // ---- one placement try, shared by G_OBJ / G_GOAL / G_AGENT / G_KEY / G_DIST / G_OBST
            const bool whole_grid = stage == G_OBJ || stage == G_GOAL || stage == G_AGENT || (stage == G_OBST && !multi);
            // pre-draw once per task: pool entry (choice + remove) or obstacle kind
            if (!pre_done && (stage == G_OBJ || stage == G_DIST)) {
                const int i = (int)mulhi32(rng.peek(used++), (uint32_t)popc32(pool));
                const int bit = nth_set_bit(pool, i);
                pool &= ~(1u << bit);
                pend_type = (int)((pool_types >> (3 * (bit / 6))) & 7u);
                pend_colour = sorted_colour(bit % 6);
                pre_done = true;
            } else if (!pre_done && stage == G_OBST && !multi) {
                pend_type = mulhi32(rng.peek(used++), 2) == 0 ? K_LAVA : K_WALL;  // choice([Lava(), Wall()])
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
            const int x = x0 + (int)mulhi32(rng.peek(used++), (uint32_t)(x1 - x0 + 1));
            const int y = y0 + (int)mulhi32(rng.peek(used++), (uint32_t)(y1 - y0 + 1));
            const int here = s.grid[y * S + x];
            // admissibility
            bool ok;
            const bool at_agent = x == agent_x && y == agent_y;
            bool n2d = false;
            if (multi && stage != G_AGENT && x >= 1 && x <= S - 2 && y >= 1 && y <= S - 2)  // next2door :2036-2046
                n2d = k_is_door(s.grid[y * S + x - 1]) || k_is_door(s.grid[y * S + x + 1]) ||
                      k_is_door(s.grid[(y - 1) * S + x]) || k_is_door(s.grid[(y + 1) * S + x]);
            if (whole_grid) ok = here == K_EMPTY && !at_agent && !(stage == G_GOAL && multi && n2d);  // place_obj
            else if (stage == G_KEY)
                ok = !(x == goal_x && y == goal_y) && !(r == agent_room && at_agent) && !(x == kx && y == ky) && !n2d;
            else if (stage == G_DIST) ok = here == K_EMPTY && !at_agent && !n2d;   // objs occupy exactly the non-empty cells
            else ok = x != m && y != m && (here == K_EMPTY || here == K_LAVA) && !at_agent && !n2d;  // lava, multi
            if (!ok && ++tries >= kMaxTries) { s.error |= ERR_TRIES; ok = true; }
            if (ok) {
                tries = 0; pre_done = false;
                if (stage == G_OBJ) {
                    s.grid[y * S + x] = (uint8_t)obj_kind(pend_type, pend_colour);
                    if (nobjs < kMaxObjs) objs[nobjs] = (uint16_t)(pend_type | (pend_colour << 3) | (x << 6) | (y << 10));
                    ++nobjs;
                    if (++obj_i == cfg.num_objects)
                        stage = (cfg.problem == P_GTG || cfg.problem == P_DRP) ? G_GOAL : G_AGENT;
                } else if (stage == G_GOAL) {
                    s.grid[y * S + x] = K_GOAL;
                    goal_x = x; goal_y = y;
                    if (nobjs < kMaxObjs) objs[nobjs] = (uint16_t)(T_GOAL | (x << 6) | (y << 10));
                    ++nobjs;
                    stage = G_AGENT;
                } else if (stage == G_AGENT) {
                    agent_x = x; agent_y = y;
                    s.agent_x = (uint8_t)x; s.agent_y = (uint8_t)y;
                    s.agent_dir = (uint8_t)mulhi32(rng.peek(used++), 4);
                    if (multi) {
                        agent_room = room_of(nrooms, m, agent_x, agent_y);
                        goal_room = room_of(nrooms, m, goal_x, goal_y);
                        stage = G_KEY; r = 0; ph = 0; kx = ky = -1;   // resolved to a real task below
                    } else {
                        stage = G_OBST;
                    }
                } else if (stage == G_KEY) {
                    const int props = (int)((doors >> (8 * d)) & 0xFFu);
                    const int colour = props & 7;
                    const bool kib = (props >> 4) & 1;
                    s.grid[y * S + x] = (uint8_t)(kib ? K_BOX + 8 * (colour + 1) + colour : K_KEY + colour);
                    if (nobjs < kMaxObjs) objs[nobjs] = (uint16_t)((kib ? T_BOX : T_KEY) | (colour << 3) | (x << 6) | (y << 10));
                    ++nobjs;
                    MGRL_CNT_DEC(r);
                    if (ph == 0) { kx = x; ky = y; }
                    ++ph;
                } else if (stage == G_DIST) {
                    s.grid[y * S + x] = (uint8_t)obj_kind(pend_type, pend_colour);
                    if (nobjs < kMaxObjs) objs[nobjs] = (uint16_t)(pend_type | (pend_colour << 3) | (x << 6) | (y << 10));
                    ++nobjs;
                    ++q;
                } else {  // G_OBST
                    s.grid[y * S + x] = (uint8_t)(multi ? K_LAVA : pend_type);
                    ++obst_i;
                }
                // ---- next task of the room walk (keys of the room, then its distractors)
                if (stage == G_KEY || stage == G_DIST) {
                    stage = G_OBST;
                    while (r < nrooms) {
                        if (ph < 2) {
                            d = key_door(nrooms, r, agent_room, ph);
                            if (d != 7 && ((doors >> (8 * d + 3)) & 1u)) { stage = G_KEY; break; }
                            ++ph;
                        } else if (ph == 2) {
                            if (goal_room == r) MGRL_CNT_DEC(r);
                            // reference quirk (custom_env.py:1119, 1660): the lower-left loop uses the upper-left counter
                            loops = (nrooms >= 3 && r == 1) ? MGRL_CNT(0) : MGRL_CNT(r);
                            q = 0; ph = 3;
                        } else {
                            if (q < loops) { stage = G_DIST; break; }
                            ++r; ph = 0; kx = ky = -1;
                        }
                    }
                }
                if (stage == G_OBST && obst_i >= cfg.num_obstacles) stage = cmd <= 2 ? G_TARGET : G_DONE;
            }
        }
        rng.ndraw += (uint32_t)used;
    }
#undef MGRL_CNT
#undef MGRL_CNT_DEC
    if (cmd == 3) { s.mission_id = MISSION_DROP; s.target_action = A_DROP; }           // :212-214
    else if (cmd == 5) {                                                                 // :258-267
        s.mission_id = MISSION_GOAL; s.target_x = (uint8_t)goal_x; s.target_y = (uint8_t)goal_y;
    }
    s.reset_draws = (uint16_t)rng.ndraw;
    s.episode += 1;
}

}  // namespace mgrl
