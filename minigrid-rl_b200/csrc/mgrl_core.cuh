// mgrl_core.cuh — per-environment device functions of the batched MiniGrid simulator.
//
// One environment = one 140-byte packed state (EnvState).  Everything here is written as
// __host__ __device__ inline code over a state reference so that (a) the CUDA kernels in
// mgrl_kernels.cu run it one-lane-per-environment on shared-memory-resident tiles and
// (b) tests/ can compile the very same functions for the host and compare them with the
// CPU oracle without a GPU (tests/host_emul.cpp; test infrastructure, not a fallback).
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
constexpr int OBS_HWC = 0;  // image[vx][vy][c]  (MiniGrid native)
constexpr int OBS_CHW = 1;  // image[c][vx][vy]  (after SB3 VecTransposeImage)

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
struct Rng {
    uint32_t k0, k1, e0, e1, episode;
    uint32_t ndraw;
    uint32_t buf[4];

    MGRL_HD void init(uint64_t seed, uint64_t env_id, uint32_t ep) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        e0 = (uint32_t)env_id; e1 = (uint32_t)(env_id >> 32);
        episode = ep; ndraw = 0;
    }
    MGRL_HD void refill() {
        uint32_t c0 = ndraw >> 2, c1 = episode, c2 = e0, c3 = e1, a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
            const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
            c0 = h1 ^ c1 ^ a; c1 = l1; c2 = h0 ^ c3 ^ b; c3 = l0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        buf[0] = c0; buf[1] = c1; buf[2] = c2; buf[3] = c3;
    }
    MGRL_HD uint32_t below(uint32_t n) {
        const uint32_t w = ndraw & 3u;
        if (w == 0) refill();
        ++ndraw;
        const uint32_t v = w == 0 ? buf[0] : w == 1 ? buf[1] : w == 2 ? buf[2] : buf[3];
        return mulhi32(v, n);
    }
    MGRL_HD int randint(int a, int b) { return a + (int)below((uint32_t)(b - a + 1)); }  // inclusive
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
// instead of bounds-checked.  `out` receives 147 bytes in LAYOUT order.
template <int LAYOUT>
MGRL_HD int obs_index(int vx, int vy, int c) {
    return LAYOUT == OBS_HWC ? (vx * kView + vy) * 3 + c : c * (kView * kView) + vx * kView + vy;
}

MGRL_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <int LAYOUT>
MGRL_HD void encode_view_see_through(const EnvState& s, int carrying, int S, uint8_t* out) {
    const int dir = s.agent_dir;
    const int dx = (dir == 0) - (dir == 2), dy = (dir == 1) - (dir == 3);
    const int rx = -dy, ry = dx;
    const int bx = s.agent_x - 3 * rx + 6 * dx, by = s.agent_y - 3 * ry + 6 * dy;  // world of view (0,0)
#pragma unroll
    for (int vx = 0; vx < kView; ++vx) {
#pragma unroll
        for (int vy = 0; vy < kView; ++vy) {
            const int wx = clampi(bx + vx * rx - vy * dx, 0, S - 1);
            const int wy = clampi(by + vx * ry - vy * dy, 0, S - 1);
            int k = s.grid[wy * S + wx];
            if (vx == 3 && vy == 6) k = carrying;
            const uint32_t e = kind_encode(k);
            out[obs_index<LAYOUT>(vx, vy, 0)] = (uint8_t)e;
            out[obs_index<LAYOUT>(vx, vy, 1)] = (uint8_t)(e >> 8);
            out[obs_index<LAYOUT>(vx, vy, 2)] = (uint8_t)(e >> 16);
        }
    }
}

// see_through_walls == false: [UPSTREAM] Grid.process_vis, mask kept as a 49-bit set
template <int LAYOUT>
MGRL_HD void encode_view_occluded(const EnvState& s, int carrying, int S, uint8_t* out) {
    const int dir = s.agent_dir;
    const int dx = (dir == 0) - (dir == 2), dy = (dir == 1) - (dir == 3);
    const int rx = -dy, ry = dx;
    const int bx = s.agent_x - 3 * rx + 6 * dx, by = s.agent_y - 3 * ry + 6 * dy;
    uint64_t opaque = 0;  // bit vx*7+vy
    for (int vx = 0; vx < kView; ++vx)
        for (int vy = 0; vy < kView; ++vy) {
            const int wx = clampi(bx + vx * rx - vy * dx, 0, S - 1);
            const int wy = clampi(by + vx * ry - vy * dy, 0, S - 1);
            if (k_opaque(s.grid[wy * S + wx])) opaque |= 1ull << (vx * kView + vy);
        }
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
            const int wx = clampi(bx + vx * rx - vy * dx, 0, S - 1);
            const int wy = clampi(by + vx * ry - vy * dy, 0, S - 1);
            int k = s.grid[wy * S + wx];
            if (vx == 3 && vy == 6) k = carrying;
            const uint32_t e = ((mask >> (vx * kView + vy)) & 1ull) ? kind_encode(k) : 0u;
            out[obs_index<LAYOUT>(vx, vy, 0)] = (uint8_t)e;
            out[obs_index<LAYOUT>(vx, vy, 1)] = (uint8_t)(e >> 8);
            out[obs_index<LAYOUT>(vx, vy, 2)] = (uint8_t)(e >> 16);
        }
}

template <int LAYOUT>
MGRL_HD void encode_view(const EnvState& s, int carrying, int S, bool see_through, uint8_t* out) {
    if (see_through) encode_view_see_through<LAYOUT>(s, carrying, S, out);
    else encode_view_occluded<LAYOUT>(s, carrying, S, out);
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
// Object list entry: type(3) | colour(3) <<3 | x <<6 | y <<10
constexpr int T_KEY = 0, T_BALL = 1, T_BOX = 2, T_DOOR = 3, T_GOAL = 4;
constexpr int kMaxObjs = 40;

struct Gen {
    EnvState& s;
    const EnvCfg& cfg;
    Rng rng;
    int S, mid;
    int agent_x, agent_y, goal_x, goal_y;
    int nobjs;
    uint32_t pool;        // bit (slot*6 + sorted colour index): remaining (type, colour) pairs
    uint32_t pool_types;  // 3 bits per slot: object type of that slot
    uint16_t objs[kMaxObjs];

    MGRL_HD Gen(EnvState& st, const EnvCfg& c) : s(st), cfg(c) {}

    // COLOR_NAMES sorted alphabetically (blue green grey purple red yellow) -> COLOR_TO_IDX
    MGRL_HD static int sorted_colour(int i) { return (int)((0x403512u >> (4 * i)) & 0xFu); }
    MGRL_HD static int sorted_index(int colour) { return (int)((0x253014u >> (4 * colour)) & 0xFu); }

    MGRL_HD uint8_t& cell(int x, int y) { return s.grid[y * S + x]; }
    MGRL_HD void add_obj(int type, int colour, int x, int y) {
        if (nobjs < kMaxObjs) objs[nobjs] = (uint16_t)(type | (colour << 3) | (x << 6) | (y << 10));
        ++nobjs;
    }
    MGRL_HD bool on_obj(int x, int y) const {
        const int key = (x << 6) | (y << 10);
        for (int i = 0; i < nobjs && i < kMaxObjs; ++i)
            if ((objs[i] & 0xFFC0) == key) return true;
        return false;
    }
    MGRL_HD void pool_fill(int t0, int t1, int t2, int t3, int ntypes) {
        pool_types = (uint32_t)(t0 | (t1 << 3) | (t2 << 6) | (t3 << 9));
        pool = (1u << (6 * ntypes)) - 1u;
    }
    MGRL_HD void pool_remove(int type, int colour) {
        for (int slot = 0; slot < 4; ++slot)
            if ((int)((pool_types >> (3 * slot)) & 7u) == type) pool &= ~(1u << (slot * 6 + sorted_index(colour)));
    }
    MGRL_HD void pool_take(int& type, int& colour) {  // choice(pool) then remove
        const int i = (int)rng.below((uint32_t)popc32(pool));
        const int bit = nth_set_bit(pool, i);
        pool &= ~(1u << bit);
        type = (int)((pool_types >> (3 * (bit / 6))) & 7u);
        colour = sorted_colour(bit % 6);
    }
    MGRL_HD static int obj_kind(int type, int colour) {
        return type == T_KEY ? K_KEY + colour : type == T_BALL ? K_BALL + colour
             : type == T_BOX ? K_BOX + colour : type == T_DOOR ? K_DOOR + 8 + colour : K_GOAL;
    }
    MGRL_HD bool next2door(int x, int y) {  // custom_env.py:2036-2046
        return k_is_door(cell(x - 1, y)) || k_is_door(cell(x + 1, y)) || k_is_door(cell(x, y - 1)) ||
               k_is_door(cell(x, y + 1));
    }
    // [UPSTREAM] place_obj over the whole grid
    MGRL_HD void place_obj(int kind, int& px, int& py) {
        int x = 0, y = 0, tries = 0;
        for (;;) {
            x = (int)rng.below((uint32_t)S);
            y = (int)rng.below((uint32_t)S);
            if (++tries >= kMaxTries) { s.error |= ERR_TRIES; break; }
            if (cell(x, y) != K_EMPTY) continue;
            if (x == agent_x && y == agent_y) continue;
            break;
        }
        cell(x, y) = (uint8_t)kind;
        px = x; py = y;
    }
    MGRL_HD void place_agent() {  // [UPSTREAM] place_agent
        int x, y;
        agent_x = agent_y = -1;
        place_obj(K_EMPTY, x, y);
        agent_x = x; agent_y = y;
        s.agent_x = (uint8_t)x; s.agent_y = (uint8_t)y;
        s.agent_dir = (uint8_t)rng.below(4);
    }
    MGRL_HD void place_goal_away_from_doors() {
        int x, y, tries = 0;
        for (;;) {
            place_obj(K_GOAL, x, y);
            if (next2door(x, y) && ++tries < kMaxTries) { cell(x, y) = K_EMPTY; continue; }
            break;
        }
        goal_x = x; goal_y = y;
        add_obj(T_GOAL, 0, x, y);
    }
};

struct Room { int x0, x1, y0, y1; };

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

// 2 rooms: L R; 3 rooms: UL LL R; 4 rooms: UL LL UR LR (interior rectangles, inclusive)
MGRL_HD Room room_rect(int nrooms, int r, int S, int m) {
    Room R;
    const bool left = nrooms == 2 ? r == 0 : r < 2;
    const bool full_height = nrooms == 2 || (nrooms == 3 && r == 2);
    const bool upper = (r & 1) == 0;
    R.x0 = left ? 1 : m + 1; R.x1 = left ? m - 1 : S - 2;
    R.y0 = (full_height || upper) ? 1 : m + 1;
    R.y1 = (full_height || !upper) ? S - 2 : m - 1;
    return R;
}

// _generate_{2,3,4}_rooms as one table-driven routine (custom_env.py:617-2034)
MGRL_HD void generate_rooms(Gen& g, int nrooms) {
    const int S = g.S, m = g.mid, n = g.cfg.num_objects;
    g.pool_fill(T_KEY, T_BALL, T_BOX, 0, 3);
    for (int i = 1; i < S - 1; ++i) g.cell(m, i) = K_WALL;
    if (nrooms == 3) for (int i = 1; i < m; ++i) g.cell(i, m) = K_WALL;
    if (nrooms == 4) for (int i = 1; i < S - 1; ++i) g.cell(i, m) = K_WALL;

    int counts[4];
    const int nl = n / 2, nr = n - nl;
    if (nrooms == 2) { counts[0] = nl; counts[1] = nr; counts[2] = counts[3] = 0; }
    else if (nrooms == 3) { counts[0] = nl / 2; counts[1] = nl - nl / 2; counts[2] = nr; counts[3] = 0; }
    else { counts[0] = nl / 2; counts[1] = nl - nl / 2; counts[2] = nr / 2; counts[3] = nr - nr / 2; }

    // door properties: colour (without replacement), locked, key_in_box  — packed 8 bits/door
    uint32_t doors = 0;  // per door: colour(3) | locked<<3 | key_in_box<<4
    uint32_t colours = 0x3Fu;
    const int ndoors = nrooms == 2 ? 1 : nrooms;
    for (int d = 0; d < ndoors; ++d) {
        const int i = (int)g.rng.below((uint32_t)popc32(colours));
        const int bit = nth_set_bit(colours, i);
        colours &= ~(1u << bit);
        const int colour = Gen::sorted_colour(bit);
        const int locked = g.cfg.all_doors_open ? 0 : (g.rng.below(2) == 0);
        const int kib = g.rng.below(2) == 0;
        if (locked) { g.pool_remove(T_KEY, colour); if (kib) g.pool_remove(T_BOX, colour); }
        doors |= (uint32_t)(colour | (locked << 3) | (kib << 4)) << (8 * d);
    }
    // door cells: (horizontal?, lo, hi) per door in the reference's order
    for (int d = 0; d < ndoors; ++d) {
        bool horizontal; int lo, hi;
        if (nrooms == 2) { horizontal = false; lo = 1; hi = S - 2; }
        else if (nrooms == 3) { horizontal = d == 0; lo = d == 2 ? m + 1 : 1; hi = d == 2 ? S - 2 : m - 1; }
        else { horizontal = d < 2; lo = (d & 1) ? m + 1 : 1; hi = (d & 1) ? S - 2 : m - 1; }
        const int p = g.rng.randint(lo, hi);
        const int is_open = g.cfg.all_doors_open ? (g.rng.below(2) == 0) : 0;
        const int props = (int)((doors >> (8 * d)) & 0xFFu);
        const int colour = props & 7, locked = (props >> 3) & 1;
        const int state = is_open ? 0 : (locked ? 2 : 1);
        const int x = horizontal ? p : m, y = horizontal ? m : p;
        g.cell(x, y) = (uint8_t)(K_DOOR + 8 * state + colour);
        g.add_obj(T_DOOR, colour, x, y);
    }
    g.place_goal_away_from_doors();
    g.place_agent();
    const int agent_room = room_of(nrooms, m, g.agent_x, g.agent_y);
    const int goal_room = room_of(nrooms, m, g.goal_x, g.goal_y);

    for (int r = 0; r < nrooms; ++r) {
        const Room R = room_rect(nrooms, r, S, m);
        int kx = -1, ky = -1;
        for (int j = 0; j < 2; ++j) {
            const int d = key_door(nrooms, r, agent_room, j);
            if (d == 7) continue;
            const int props = (int)((doors >> (8 * d)) & 0xFFu);
            const int colour = props & 7;
            if (!((props >> 3) & 1)) continue;  // key only for a locked door
            int x = 0, y = 0, tries = 0;
            for (;;) {
                x = g.rng.randint(R.x0, R.x1);
                y = g.rng.randint(R.y0, R.y1);
                if (++tries >= kMaxTries) { g.s.error |= ERR_TRIES; break; }
                if (x == g.goal_x && y == g.goal_y) continue;
                if (r == agent_room && x == g.agent_x && y == g.agent_y) continue;
                if (x == kx && y == ky) continue;
                if (g.next2door(x, y)) continue;
                break;
            }
            if ((props >> 4) & 1) {  // Box(colour, Key(colour))
                g.cell(x, y) = (uint8_t)(K_BOX + 8 * (colour + 1) + colour);
                g.add_obj(T_BOX, colour, x, y);
            } else {
                g.cell(x, y) = (uint8_t)(K_KEY + colour);
                g.add_obj(T_KEY, colour, x, y);
            }
            counts[r]--;
            if (j == 0) { kx = x; ky = y; }
        }
        if (goal_room == r) counts[r]--;
        // reference quirk (custom_env.py:1119, 1660): the lower-left loop uses the upper-left counter
        const int loops = (nrooms >= 3 && r == 1) ? counts[0] : counts[r];
        for (int q = 0; q < loops; ++q) {
            int type, colour;
            g.pool_take(type, colour);
            int x = 0, y = 0, tries = 0;
            for (;;) {
                x = g.rng.randint(R.x0, R.x1);
                y = g.rng.randint(R.y0, R.y1);
                if (++tries >= kMaxTries) { g.s.error |= ERR_TRIES; break; }
                if (g.on_obj(x, y)) continue;
                if (x == g.agent_x && y == g.agent_y) continue;
                if (g.next2door(x, y)) continue;
                break;
            }
            g.cell(x, y) = (uint8_t)Gen::obj_kind(type, colour);
            g.add_obj(type, colour, x, y);
        }
    }
}

// single-room generators (custom_env.py:371-555)
MGRL_HD void generate_single(Gen& g, int problem) {
    if (problem == P_GTG) g.pool_fill(T_BOX, T_DOOR, T_KEY, T_BALL, 4);
    else if (problem == P_OPN) g.pool_fill(T_BOX, T_DOOR, 0, 0, 2);
    else if (problem == P_PKP) g.pool_fill(T_KEY, T_BOX, T_BALL, 0, 3);
    else g.pool_fill(T_KEY, T_BALL, T_BOX, T_DOOR, 4);
    for (int i = 0; i < g.cfg.num_objects; ++i) {
        int type, colour, x, y;
        g.pool_take(type, colour);
        g.place_obj(Gen::obj_kind(type, colour), x, y);
        g.add_obj(type, colour, x, y);
    }
    if (problem == P_GTG || problem == P_DRP) {
        int x, y;
        g.place_obj(K_GOAL, x, y);
        g.goal_x = x; g.goal_y = y;
        g.add_obj(T_GOAL, 0, x, y);
    }
    g.place_agent();
}

// obstacles (custom_env.py:155-172)
MGRL_HD void place_obstacles(Gen& g) {
    const int S = g.S;
    for (int i = 0; i < g.cfg.num_obstacles; ++i) {
        if (g.cfg.problem == P_MULTI) {
            int x = 0, y = 0, tries = 0;
            for (;;) {
                x = g.rng.randint(1, S - 2);
                y = g.rng.randint(1, S - 2);
                if (++tries >= kMaxTries) { g.s.error |= ERR_TRIES; break; }
                if (x == g.mid || y == g.mid) continue;
                if (g.on_obj(x, y)) continue;
                if (x == g.agent_x && y == g.agent_y) continue;
                if (g.next2door(x, y)) continue;
                break;
            }
            g.cell(x, y) = K_LAVA;
        } else {
            int x, y;
            const int kind = g.rng.below(2) == 0 ? K_LAVA : K_WALL;
            g.place_obj(kind, x, y);
        }
    }
}

// PlaygroundEnv._gen_grid (custom_env.py:122-267) for episode s.episode of env `env_id`.
// Leaves mission_done / latch_step untouched (they survive resets in the reference).
MGRL_HD void generate(EnvState& s, const EnvCfg& cfg, uint64_t seed, uint64_t env_id) {
    Gen g(s, cfg);
    g.S = cfg.size; g.mid = cfg.size / 2;
    g.agent_x = g.agent_y = g.goal_x = g.goal_y = -1;
    g.nobjs = 0; g.pool = 0; g.pool_types = 0;
    g.rng.init(seed, env_id, s.episode);
    const int S = g.S;
    for (int i = 0; i < kGridCells; ++i) s.grid[i] = K_EMPTY;
    for (int i = 0; i < S; ++i) {
        g.cell(i, 0) = K_WALL; g.cell(i, S - 1) = K_WALL; g.cell(0, i) = K_WALL; g.cell(S - 1, i) = K_WALL;
    }
    s.carrying = 0; s.step_count = 0;
    s.target_x = s.target_y = kNone; s.target_action = 0;

    int cmd;
    if (cfg.problem == P_MULTI) {
        cmd = cfg.mission >= 0 ? cfg.mission : (int)((0x5210u >> (4 * g.rng.below(4))) & 0xFu);  // choice([0,1,2,5])
        generate_rooms(g, g.rng.randint(2, 4));
    } else {
        generate_single(g, cfg.problem);
        cmd = cfg.problem == P_GTO ? 0 : cfg.problem == P_GTG ? 5 : cfg.problem == P_OPN ? 1
            : cfg.problem == P_PKP ? 2 : 3;
    }
    if (cfg.num_obstacles > 0) place_obstacles(g);

    // target selection (custom_env.py:174-267)
    if (cmd <= 2) {
        const int n = g.nobjs < kMaxObjs ? g.nobjs : kMaxObjs;
        int o = 0, tries = 0;
        for (;;) {
            o = g.objs[g.rng.below((uint32_t)n)];
            const int t = o & 7;
            const bool ok = cmd == 0 ? t != T_GOAL : cmd == 1 ? (t == T_BOX || t == T_DOOR)
                                                              : (t == T_BOX || t == T_KEY || t == T_BALL);
            if (ok) break;
            if (++tries >= kMaxTries) { s.error |= ERR_TRIES; break; }
        }
        s.mission_id = (uint8_t)(cmd * 24 + (o & 7) * 6 + ((o >> 3) & 7));
        s.target_x = (uint8_t)((o >> 6) & 15); s.target_y = (uint8_t)((o >> 10) & 15);
        s.target_action = (uint8_t)(cmd == 0 ? A_DONE : cmd == 1 ? A_TOGGLE : A_PICKUP);
    } else if (cmd == 3) {
        s.mission_id = MISSION_DROP; s.target_action = A_DROP;
    } else {
        s.mission_id = MISSION_GOAL;
        s.target_x = (uint8_t)g.goal_x; s.target_y = (uint8_t)g.goal_y;
    }
    s.reset_draws = (uint16_t)g.rng.ndraw;
    s.episode += 1;
}

}  // namespace mgrl
