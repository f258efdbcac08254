// mgrl_kernels.cu — sm_100a kernels and the C ABI (include/mgrl.h) of the batched MiniGrid
// simulator.  See DESIGN.md for the data layout and the roofline of each kernel.
//
// Kernel map (SURVEY.md §2.1 ids):
//   step_kernel<LAYOUT,SEE>      K1+K2  T >= 1 steps per launch: step -> mission bookkeeping -> reward/done ->
//                                       auto-reset by adopting a prepared layout -> 7x7x3 encode; the same
//                                       warps build new layouts in dense batches of 32 between steps
//   generate_kernel              K2     dense rebuild of the layouts that one-step launches used up (every kDepth steps)
//   reset_kernel                 K2     fresh environments + kDepth prepared layouts each (also after set_state)
//   gae_kernel                   K4     SB3 GAE reverse scan
//   stack_push_kernel                   VecFrameStack + Discrete2Box + TokenizeVocab gathers
//
// Thread mapping: one CTA = one tile of TILE consecutive environments, one lane per environment for
// the dynamics; the tile's packed state (TILE x 140 B) lives in shared memory for the whole launch and
// each warp stages its 32 observation records there so that every global access is a coalesced
// 16-byte vector transfer.
//
// Auto-reset without divergence: the layout of (env, episode) depends only on the RNG key, never on
// the actions, so it is built AHEAD of time.  Every environment owns kDepth layout slots in global
// memory (L2-resident).  A finished environment copies the slot of its next episode and queues a
// request to refill it; whenever 32 requests are waiting, the warp that is furthest ahead takes them
// and runs the generator with all 32 lanes busy.  Warps never meet at a block barrier inside the step
// loop, and the queue survives across launches, so one-step launches batch the same way.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "mgrl.h"
#include "mgrl_core.cuh"
#include "mgrl_wire.cuh"

using namespace mgrl;

namespace {

constexpr int STATE_WORDS = 35;           // sizeof(EnvState) / 4
constexpr int kDepth = 4;                 // layouts prepared ahead of time per environment (a power of two)
static_assert((kDepth & (kDepth - 1)) == 0, "kDepth must be a power of two");
constexpr int kSlotWords = 36;             // a prepared layout: the 35 state words + 1 pad word (16-byte aligned records)
constexpr int kQueueCap = 512;            // >= TILE * kDepth requests can be outstanding per tile
constexpr uint32_t kNoEntry = 0xFFFFu;
constexpr unsigned FULL = 0xffffffffu;
constexpr int kSpinLimit = 1 << 20;

static_assert(sizeof(EnvState) == MGRL_STATE_BYTES, "ABI state size");

struct EnvParams {
    EnvCfg cfg;
    uint64_t seed;
    uint64_t env_id_base;
    int n;
    int T;                    // steps per launch
    EnvState* states;         // [n] current state of every environment
    uint32_t* slots;          // [kDepth][n][kSlotWords] prepared layouts: slot (E % kDepth) holds episode E once its
                              //   episode word (the tag) reads E + 1
    uint8_t* tags;            // [kDepth][n] low byte of every slot's tag, compact copy for the coalesced prologue load
    uint16_t* qsave;          // [tiles][kQueueCap] generation requests left over by the previous launch
    uint32_t* qcount;         // [tiles]
    uint32_t* glist;          // [kDepth * n] deferred requests of one-step launches: env * kDepth + slot
    uint32_t* gcount;         // [1]
    int defer;                // 1: this launch only records its requests in glist (generate_kernel builds them later)
    const uint32_t* tasks;    // [kTaskEntries][task_row_words] (build_task_table + pack_task_table)
    int task_row_words;
    const uint32_t* prefix;   // [kTaskWords] (build_task_prefix)
    const uint32_t* empty;    // [kGridWords] (build_empty_grid)
    const float* reward_lut;  // [max_steps+1] device
    const uint8_t* actions;
    uint8_t* image;
    uint8_t* dir;
    uint8_t* mission;
    float* reward;
    uint8_t* term;
    uint8_t* trunc;
    uint8_t* ep_len;
    uint8_t* term_image;
    uint8_t* term_dir;
    // rollout_kernel: layout requests that were still queued when the step warps of a tile finished travel to the next
    // launch (carry = 1) instead of being built while the step warps idle; a launch with T = 0 and carry = 0 drains them
    uint16_t* rq_save;        // [tiles][qcap]
    uint32_t* rq_count;       // [tiles]
    int carry;
};

__device__ __forceinline__ uint32_t* slot_ptr(uint32_t* slots, int j, int n, int env) {
    return slots + ((size_t)j * n + env) * kSlotWords;
}

// ---- cooperative copies (coalesced; 16-byte vectors when size and address allow) ------------
template <int NTHREADS>
__device__ __noinline__ void coop_copy(void* dst, const void* src, int bytes, int tid) {
    if ((bytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const uint4* g = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll 1
        for (int i = tid; i < (bytes >> 4); i += NTHREADS) d[i] = g[i];
    } else if ((bytes & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 3) == 0) {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
#pragma unroll 1
        for (int i = tid; i < (bytes >> 2); i += NTHREADS) d[i] = g[i];
    } else {
        const uint8_t* g = reinterpret_cast<const uint8_t*>(src);
        uint8_t* d = reinterpret_cast<uint8_t*>(dst);
        const int n16 = (reinterpret_cast<uintptr_t>(dst) & 15) == 0 ? (bytes >> 4) : 0;  // smem src is 16-B aligned
#pragma unroll 1
        for (int i = tid; i < n16; i += NTHREADS) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(g)[i];
#pragma unroll 1
        for (int i = (n16 << 4) + tid; i < bytes; i += NTHREADS) d[i] = g[i];
    }
}

// one warp's 32 staged observation records -> global (the common case: full warp, 148-byte records)
template <int PITCH>
__device__ __forceinline__ void warp_copy_out(uint8_t* dst, const uint8_t* src, int nvw, int lane) {
    if (PITCH == kObsPitch148 && nvw == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const uint4* g = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int i = 0; i < (32 * kObsPitch148 / 16 + 31) / 32; ++i) {
            const int j = i * 32 + lane;
            if (j < 32 * kObsPitch148 / 16) __stcs(d + j, g[j]);   // streaming: the record is not read again here
        }
    } else {
        coop_copy<32>(dst, src, nvw * PITCH, lane);
    }
}

// ---- shared-memory-only observation encoder (the step kernel's fast path) --------------------
// Same result as encode_view_packed (mgrl_core.cuh), written against 32-bit shared addresses so that
// every access is an LDS/STS with one address instruction: per view row the 7 kind bytes are
// gathered first, then their 7 LUT words, then packed into aligned words of the staged record.
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }

__device__ __forceinline__ void encode_packed_smem(uint32_t state_sa, int agent_x, int agent_y, int dir, uint32_t carrying,
                                                   int S, uint32_t lut_sa, uint32_t out_sa) {
    const bool even = (dir & 1) == 0;
    const int sf = dir < 2 ? 1 : -1;
    const int sr = (dir == 0 || dir == 3) ? 1 : -1;
    const int af = even ? agent_x : agent_y, ar = even ? agent_y : agent_x;
    const int mf = even ? 1 : S, mr = even ? S : 1;
    int cf[kView];
#pragma unroll
    for (int i = 0; i < kView; ++i) cf[i] = clampi(af + (6 - i) * sf, 0, S - 1);
    uint32_t e[4];
    // software pipeline over the view rows: the kind bytes of row vx+1 are in flight while row vx goes
    // through the LUT and is packed, so each row exposes one shared-memory latency instead of two
    uint32_t k[kView], kn[kView];
    {
        const uint32_t row = state_sa + (uint32_t)(clampi(ar - 3 * sr, 0, S - 1) * mr);
#pragma unroll
        for (int vy = 0; vy < kView; ++vy) k[vy] = lds_u8(row + (uint32_t)(cf[vy] * mf));
    }
#pragma unroll
    for (int vx = 0; vx < kView; ++vx) {
        if (vx + 1 < kView) {
            const uint32_t row = state_sa + (uint32_t)(clampi(ar + (vx + 1 - 3) * sr, 0, S - 1) * mr);
#pragma unroll
            for (int vy = 0; vy < kView; ++vy)
                kn[vy] = (vx + 1 == 3 && vy == 6) ? carrying : lds_u8(row + (uint32_t)(cf[vy] * mf));
        }
        uint32_t w[kView];
#pragma unroll
        for (int vy = 0; vy < kView; ++vy) w[vy] = lds_u32(lut_sa + k[vy] * 4u);
#pragma unroll
        for (int vy = 0; vy < kView; ++vy) {
            const int c = vx * kView + vy;
            e[c & 3] = w[vy];
            if ((c & 3) == 3) {
                const uint32_t o = out_sa + (uint32_t)((c >> 2) * 12);
                sts_u32(o, pack3(e[0], e[1], 0));
                sts_u32(o + 4, pack3(e[1], e[2], 1));
                sts_u32(o + 8, pack3(e[2], e[3], 2));
            }
        }
#pragma unroll
        for (int vy = 0; vy < kView; ++vy) k[vy] = kn[vy];
    }
    sts_u32(out_sa + 144, e[0]);  // cell 48 + pad byte
}

// Shared memory of one tile.  Warps are independent between the prologue and the epilogue: each
// owns 32 environments (state + a staging area for their observation records) and the tile shares
// a queue of layout requests and NB draw buffers for the warps that serve it.
template <int TILE, int NB>
struct TileSmem {
    alignas(16) uint32_t state[TILE * STATE_WORDS];
    alignas(16) uint8_t obs[TILE / 32][32 * kObsPitch148];   // per warp: observation staging / layout scratch
    alignas(16) uint32_t draws[NB][kGenWords * 32];           // lane-interleaved draw ring + object list of a serving warp
    uint32_t kind_lut[128];
    float lut[kGridCells + 1];
    uint32_t empty[kGridWords];
    uint32_t prefix[kTaskWords];
    uint16_t queue[kQueueCap];   // ring of requests: local env | slot << 8; kNoEntry = not written yet
    uint32_t q_head, q_tail;
    int lock[NB];
    int warp_t[TILE / 32];       // steps finished by each warp
    uint8_t ready[kDepth][TILE]; // low byte of each slot's tag (episode + 1 of the layout it holds)
};
static_assert(32 * kObsPitch148 >= 32 * STATE_WORDS * 4, "layout scratch must fit the staging area");

template <int TILE, int NB>
__device__ __forceinline__ void tile_prologue(TileSmem<TILE, NB>& sm, const EnvParams& p, int tid) {
    fill_kind_lut(sm.kind_lut, tid, TILE);
    for (int i = tid; i <= p.cfg.max_steps; i += TILE) sm.lut[i] = p.reward_lut[i];
    for (int i = tid; i < kGridWords; i += TILE) sm.empty[i] = p.empty[i];
    for (int i = tid; i < kTaskWords; i += TILE) sm.prefix[i] = p.prefix[i];
}

// queue a layout request for every lane of `mask` (warp-aggregated; call converged)
template <int TILE, int NB>
__device__ __forceinline__ void push_requests(TileSmem<TILE, NB>& sm, unsigned mask, uint32_t entry, int lane) {
    if (mask == 0u) return;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&sm.q_tail, (uint32_t)__popc(mask));
    base = __shfl_sync(FULL, base, leader);
    if ((mask >> lane) & 1u) {
        const uint32_t pos = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
        *reinterpret_cast<volatile uint16_t*>(&sm.queue[pos & (kQueueCap - 1)]) = (uint16_t)entry;
    }
}

// the step kernel's single out-of-line copy of the generator
// (leaves the next-to-a-door marks in the grid words: strip them with kMarkMask on the way out)
template <int MODE = GEN_ALL>
__device__ __noinline__ void generate_layout(uint32_t* sc, const EnvCfg& cfg, uint64_t seed, uint64_t env_id, uint32_t episode,
                                             uint32_t* draws, const uint32_t* tasks, int row_words, const uint32_t* prefix,
                                             const uint32_t* empty) {
    sc[32] = 0u; sc[34] = 0u;
    GenIO io;
    io.draws = draws; io.stride = 32; io.tasks = tasks; io.row_words = row_words; io.prefix = prefix; io.empty = empty;
    io.keep_marks = true;
    generate<MODE>(*reinterpret_cast<EnvState*>(sc), cfg, seed, env_id, episode, io);
}

// Take up to 32 requests off the tile's queue and build their layouts, one lane each (dense
// generation).  A full batch is required unless `partial` (the caller is waiting for a layout).
// Returns false when there was nothing to take or no draw buffer was free.  Call converged.
template <int TILE, int NB>
__device__ __noinline__ bool serve_queue(TileSmem<TILE, NB>& sm, const EnvParams& p, int tile0, int warp, int lane,
                                         bool partial) {
    int b = -1, h = 0, n = 0;
    if (lane == 0) {
        const uint32_t head = *reinterpret_cast<volatile uint32_t*>(&sm.q_head);
        const uint32_t avail = *reinterpret_cast<volatile uint32_t*>(&sm.q_tail) - head;
        n = avail >= 32u ? 32 : (partial ? (int)avail : 0);
        if (n > 0) {
            for (int i = 0; i < NB && b < 0; ++i)
                if (atomicCAS(&sm.lock[i], 0, 1) == 0) b = i;
            if (b >= 0) {
                if (atomicCAS(&sm.q_head, head, head + (uint32_t)n) != head) { atomicExch(&sm.lock[b], 0); b = -1; }
                else h = (int)head;
            }
        }
    }
    b = __shfl_sync(FULL, b, 0);
    if (b < 0) return false;
    h = __shfl_sync(FULL, h, 0);
    n = __shfl_sync(FULL, n, 0);
    if (lane < n) {
        volatile uint16_t* q = reinterpret_cast<volatile uint16_t*>(&sm.queue[(h + lane) & (kQueueCap - 1)]);
        uint32_t ent;
        while ((ent = *q) == kNoEntry) {}
        *q = (uint16_t)kNoEntry;
        const int e = (int)(ent & 0xFFu), j = (int)(ent >> 8);
        uint32_t* slot = slot_ptr(p.slots, j, p.n, tile0 + e);
        // The slot's old layout was adopted, so it now gets the one episode in [E, E + kDepth) that maps to slot j, E being
        // the environment's next episode.  (E may advance while we look: every value it can take gives the same answer,
        // because the environment cannot adopt the layout we are about to build.)
        const uint32_t E = *reinterpret_cast<const volatile uint32_t*>(&sm.state[e * STATE_WORDS + 33]);
        const uint32_t episode = E + (((uint32_t)j - E) & (uint32_t)(kDepth - 1));
        uint32_t* sc = reinterpret_cast<uint32_t*>(sm.obs[warp]) + lane * STATE_WORDS;
        generate_layout(sc, p.cfg, p.seed, p.env_id_base + (uint64_t)(tile0 + e), episode, sm.draws[b] + lane, p.tasks,
                        p.task_row_words, sm.prefix, sm.empty);
#pragma unroll
        for (int i = 0; i < STATE_WORDS; ++i) __stcg(slot + i, i < kGridWords ? sc[i] & kMarkMask : sc[i]);
        // hand-off inside the CTA: layout (global) -> fence -> tag byte (shared); the adopting lane
        // reads the tag byte, fences, then reads the layout
        p.tags[(size_t)j * p.n + tile0 + e] = (uint8_t)sc[33];
        __threadfence_block();
        *reinterpret_cast<volatile uint8_t*>(&sm.ready[j][e]) = (uint8_t)sc[33];
    }
    __syncwarp();
    if (lane == 0) { __threadfence_block(); atomicExch(&sm.lock[b], 0); }
    return true;
}

// T simulator steps of one tile.  No block-wide barrier inside the loop.
template <int LAYOUT, bool SEE, int TILE, int NB>
__global__ void __launch_bounds__(TILE, TILE == 128 ? 4 : 7) step_kernel(const EnvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem<TILE, NB>& sm = *reinterpret_cast<TileSmem<TILE, NB>*>(smem_raw);
    constexpr int PITCH = obs_pitch(LAYOUT);
    constexpr int NW = TILE / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile0 = blockIdx.x * TILE;
    const int nv = min(TILE, p.n - tile0);
    const int S = p.cfg.size;

    if (nv == TILE) {   // full tile: all 16-byte loads of a thread in flight together (the generic copy is rolled)
        const uint4* g = reinterpret_cast<const uint4*>(p.states + tile0);
        uint4* d = reinterpret_cast<uint4*>(sm.state);
        constexpr int NV4 = TILE * (int)sizeof(EnvState) / 16;
        uint4 v[(NV4 + TILE - 1) / TILE];
#pragma unroll
        for (int i = 0; i < (NV4 + TILE - 1) / TILE; ++i)
            if (i * TILE + tid < NV4) v[i] = g[i * TILE + tid];
#pragma unroll
        for (int i = 0; i < (NV4 + TILE - 1) / TILE; ++i)
            if (i * TILE + tid < NV4) d[i * TILE + tid] = v[i];
    } else {
        coop_copy<TILE>(sm.state, p.states + tile0, nv * (int)sizeof(EnvState), tid);
    }
    tile_prologue<TILE, NB>(sm, p, tid);
    {   // requests the previous launch left behind
        uint32_t nq = p.qcount[blockIdx.x];
        const uint16_t* qs = p.qsave + (size_t)blockIdx.x * kQueueCap;
        if (p.defer && nq) {   // a one-step launch does not build layouts: hand the leftovers to generate_kernel
            __shared__ uint32_t gbase;
            if (tid == 0) gbase = atomicAdd(p.gcount, nq);
            __syncthreads();
            for (uint32_t i = tid; i < nq; i += TILE)
                p.glist[gbase + i] = (uint32_t)(tile0 + (qs[i] & 0xFFu)) * kDepth + (uint32_t)(qs[i] >> 8);
            nq = 0u;
        }
        for (int i = tid; i < kQueueCap; i += TILE) sm.queue[i] = (uint32_t)i < nq ? qs[i] : (uint16_t)kNoEntry;
        if (tid == 0) { sm.q_head = 0u; sm.q_tail = nq; }
        if (tid < NB) sm.lock[tid] = 0;
        if (tid < NW) sm.warp_t[tid] = 0;
#pragma unroll
        for (int j = 0; j < kDepth; ++j)
            sm.ready[j][tid] = tid < nv ? p.tags[(size_t)j * p.n + tile0 + tid] : (uint8_t)0;
    }
    __syncthreads();

    const bool active = tid < nv;
    uint32_t* cur = sm.state + tid * STATE_WORDS;
    EnvState& s = *reinterpret_cast<EnvState*>(cur);
    uint8_t* stage = sm.obs[warp];
    const int nvw = max(0, min(32, nv - warp * 32));
    volatile int* warp_t = sm.warp_t;

    __builtin_assume(__isShared(stage));
    __builtin_assume(__isShared(cur));
    int action = active ? p.actions[tile0 + tid] : 0;
    for (int t = 0; t < p.T; ++t) {
        const size_t gi = (size_t)t * (size_t)p.n + (size_t)(tile0 + tid);
        bool done = false;
        int carry = 0;
        const int a = action;
        if (active && t + 1 < p.T) action = p.actions[gi + (size_t)p.n];   // next step's action, in flight during this one
        // The `done` action always ends the episode (custom_env.py:319-328), so a lane that is about to take it knows
        // now that it will adopt its next layout in this step: if that layout is ready, fetch it asynchronously
        // (cp.async, 9 x 16 B, L2 -> the lane's row of the staging area) while the step itself is computed.
        bool early = false;
        if (active && a == A_DONE) {
            const uint32_t E0 = s.episode;
            const int j0 = (int)(E0 % kDepth);
            if (*reinterpret_cast<const volatile uint8_t*>(&sm.ready[j0][tid]) == (uint8_t)(E0 + 1u)) {
                __threadfence_block();
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + lane * (kSlotWords * 4));
                const uint32_t* src = slot_ptr(p.slots, j0, p.n, tile0 + tid);
#pragma unroll
                for (int i = 0; i < kSlotWords / 4; ++i)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 4 * i) : "memory");
                early = true;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (active) {
            const StepOut o = env_step(s, a, S, p.cfg.max_steps, sm.lut);
            p.reward[gi] = o.reward;
            p.term[gi] = o.terminated;
            p.trunc[gi] = o.truncated;
            done = o.terminated | o.truncated;
            if (p.ep_len) p.ep_len[gi] = done ? s.step_count : (uint8_t)0;
            carry = o.carry_obs;
            if (done) {   // info['terminal_observation']
                if (p.term_image) encode_view<LAYOUT>(s, carry, S, SEE, sm.kind_lut, p.term_image + gi * PITCH);
                if (p.term_dir) p.term_dir[gi] = s.agent_dir;
            }
        }
        if (__any_sync(FULL, done)) {
            // finished environments adopt the layout prepared for their next episode and ask for another
            bool pending = done;
            const uint32_t E = s.episode;
            const int j = (int)(E % kDepth);
            const uint32_t* slot = slot_ptr(p.slots, j, p.n, tile0 + tid);
            const uint32_t entry = (uint32_t)tid | ((uint32_t)j << 8);
            const volatile uint8_t* tag = &sm.ready[j][tid];
            int spins = 0;
            for (;;) {
                bool ready = false;
                if (pending && early) {          // the layout is already on its way into this lane's staging row
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    adopt_layout(cur, reinterpret_cast<const uint32_t*>(stage + lane * (kSlotWords * 4)));
                    ready = true; pending = false; carry = 0;
                } else if (pending && *tag == (uint8_t)(E + 1u)) {
                    __threadfence_block();
                    uint32_t w[STATE_WORDS];
#pragma unroll
                    for (int i = 0; i < STATE_WORDS; ++i) w[i] = __ldcg(slot + i);
                    adopt_layout(cur, w);
                    ready = true; pending = false; carry = 0;
                }
                const unsigned rm = __ballot_sync(FULL, ready);
                if (!p.defer) {
                    push_requests<TILE, NB>(sm, rm, entry, lane);
                } else if (rm) {   // one-step launch: record the request for generate_kernel (warp-aggregated)
                    const int leader = __ffs(rm) - 1;
                    uint32_t base = 0;
                    if (lane == leader) base = atomicAdd(p.gcount, (uint32_t)__popc(rm));
                    base = __shfl_sync(FULL, base, leader);
                    if (ready) p.glist[base + (uint32_t)__popc(rm & ((1u << lane) - 1u))] = (uint32_t)(tile0 + tid) * kDepth + (uint32_t)j;
                }
                if (!__any_sync(FULL, pending)) break;
                if (p.defer) {
                    // One-step launches do not serve requests.  generate_kernel refills every slot at least every kDepth
                    // steps, so this is only reached right after a switch from multi-step launches whose leftover requests
                    // covered all slots of an environment: build the layout here and leave the slot's pending request to
                    // produce the episode after it (the tag says "adopted").
                    int b = -1;
                    if (lane == 0)
                        while (b < 0)
                            for (int i = 0; i < NB && b < 0; ++i)
                                if (atomicCAS(&sm.lock[i], 0, 1) == 0) b = i;
                    b = __shfl_sync(FULL, b, 0);
                    if (pending) {
                        uint32_t* sc = reinterpret_cast<uint32_t*>(sm.obs[warp]) + lane * STATE_WORDS;
                        generate_layout(sc, p.cfg, p.seed, p.env_id_base + (uint64_t)(tile0 + tid), E, sm.draws[b] + lane, p.tasks,
                                        p.task_row_words, sm.prefix, sm.empty);
#pragma unroll 4
                        for (int i = 0; i < kGridWords; ++i) sc[i] &= kMarkMask;
                        adopt_layout(cur, sc);
                        __stcg(const_cast<uint32_t*>(slot) + 33, E + 1u);
                        p.tags[(size_t)j * p.n + tile0 + tid] = (uint8_t)(E + 1u);
                        *const_cast<volatile uint8_t*>(tag) = (uint8_t)(E + 1u);
                        pending = false; carry = 0;
                    }
                    __syncwarp();
                    if (lane == 0) atomicExch(&sm.lock[b], 0);
                    break;
                }
                // a layout is not there yet: serve the queue ourselves (any batch size), else back off
                if (!serve_queue<TILE, NB>(sm, p, tile0, warp, lane, true)) __nanosleep(200);
                if (++spins > kSpinLimit) {
                    if (pending) s.error |= ERR_SYNC;
                    break;
                }
            }
        }
        if (active) {
            if (p.image) {
                if (LAYOUT == OBS_HWC148 && SEE) {
                    asm volatile("" ::: "memory");   // the state bytes written above are read through asm loads
                    encode_packed_smem((uint32_t)__cvta_generic_to_shared(cur), s.agent_x, s.agent_y, s.agent_dir,
                                       (uint32_t)carry, S, (uint32_t)__cvta_generic_to_shared(sm.kind_lut),
                                       (uint32_t)__cvta_generic_to_shared(stage + lane * PITCH));
                    asm volatile("" ::: "memory");
                } else {
                    encode_view<LAYOUT>(s, carry, S, SEE, sm.kind_lut, stage + lane * PITCH);
                }
            }
            if (p.dir) p.dir[gi] = s.agent_dir;
            if (p.mission) p.mission[gi] = s.mission_id;
        }
        __syncwarp();
        if (p.image && nvw > 0)
            warp_copy_out<PITCH>(p.image + ((size_t)t * (size_t)p.n + (size_t)(tile0 + warp * 32)) * PITCH, stage, nvw, lane);
        __syncwarp();
        // the warp that is furthest ahead builds the next batch of layouts
        if (lane == 0) warp_t[warp] = t + 1;
        if (!p.defer) {
            // cheap look before the call: is a full batch waiting, and is this warp (one of) the furthest ahead?
            const uint32_t waiting = *reinterpret_cast<volatile uint32_t*>(&sm.q_tail) - *reinterpret_cast<volatile uint32_t*>(&sm.q_head);
            if (waiting >= 32u) {
                bool lead = true;
#pragma unroll
                for (int w2 = 0; w2 < NW; ++w2) lead = lead && warp_t[w2] <= t + 1;
                if (lead) serve_queue<TILE, NB>(sm, p, tile0, warp, lane, false);
            }
        }
    }
    // keep serving full batches until every warp of the tile has finished its steps
    for (int spins = 0; spins < kSpinLimit && !p.defer; ++spins) {
        bool all = true;
#pragma unroll
        for (int w2 = 0; w2 < NW; ++w2) all = all && warp_t[w2] >= p.T;
        if (all) break;
        if (!serve_queue<TILE, NB>(sm, p, tile0, warp, lane, false)) __nanosleep(100);
    }
    __syncthreads();
    coop_copy<TILE>(p.states + tile0, sm.state, nv * (int)sizeof(EnvState), tid);
    {   // requests still queued travel to the next launch
        const uint32_t head = sm.q_head, nq = sm.q_tail - head;
        uint16_t* qs = p.qsave + (size_t)blockIdx.x * kQueueCap;
        for (uint32_t i = tid; i < nq; i += TILE) qs[i] = sm.queue[(head + i) & (kQueueCap - 1)];
        if (tid == 0) p.qcount[blockIdx.x] = nq;
    }
}

// ======================================================================================= rollout_kernel
// T > 1 steps per launch (mgrl_step_many): one big tile per SM, warp-specialised.
//   * `sw` STEP warps, one lane per environment: step -> (adopt the prepared layout of the next episode) -> encode the
//     observation into the warp's staging area -> ONE bulk copy (cp.async.bulk, TMA engine) of the 32 records to HBM;
//   * `gw` GENERATOR warps that never step: they pop the tile's request queue in batches of 32 and build layouts
//     densely (one lane per layout) into their own scratch rows, then write the 32 slot records with coalesced 16-byte
//     stores and publish them through the tile's tag bytes.
// The layout generator is ~44 % of the instructions of a uniform-random rollout; on its own warps it adds independent
// warps per scheduler (22 instead of 14 per SM at 65 536 environments) instead of stalling the step warps.  Adoption is
// nine 16-byte cp.async per finished lane straight from the slot record (L2) into the lane's state row: env_step_apply
// <KEEP_IF_DONE> writes nothing when the episode ends, and for the `done` action (which always ends it,
// custom_env.py:319-328) the copy is issued before the step is computed.  Every slot is full when a launch starts and
// when it ends (the generator warps drain the queue), so launches leave no requests behind.
constexpr int kRowWords = 36;        // state row pitch in shared memory: 144 B = 16-byte aligned cp.async destination
constexpr int kScratchWords = 35;    // generator scratch row pitch (odd: lane-per-row accesses are conflict free)
constexpr int kMaxStepWarps = 14, kMaxGenWarps = 10;
constexpr uint32_t kEnvBits = 9;     // request = env-in-tile (9 bits, <= 448) | slot << 9

struct RolloutSmem {                 // byte offsets into the dynamic shared memory of one tile
    int state, stage, scratch, gen, queue, ready, kind_lut, lut, empty, prefix, ctrl, total, qcap;
};
__host__ __device__ inline RolloutSmem rollout_smem(int sw, int gw, int gen_words) {
    RolloutSmem L;
    int o = 0;
    L.state = o; o += sw * 32 * kRowWords * 4;
    L.stage = o; o += sw * 32 * kObsPitch148;
    L.scratch = o; o += gw * 32 * kScratchWords * 4;
    L.gen = o; o += gw * 32 * gen_words * 4;
    int qcap = 64;
    while (qcap < sw * 32 * kDepth) qcap <<= 1;
    L.qcap = qcap;
    L.queue = o; o += qcap * 2;
    L.ready = o; o += kDepth * sw * 32;
    L.kind_lut = o; o += 128 * 4;
    L.lut = o; o += 128 * 4;         // kGridCells + 1 = 122 floats
    L.empty = o; o += 32 * 4;
    L.prefix = o; o += kTaskWords * 4;
    L.ctrl = o; o += 64;
    L.total = o;
    return L;
}
struct RolloutCtrl {
    uint32_t q_head, q_tail;
    int finished;        // step warps that have done their T steps
    int starve;          // a step lane is waiting for a layout: generator warps take partial batches
};

__device__ __forceinline__ void fetch_layout(uint32_t row_sa, const uint32_t* slot) {
#pragma unroll
    for (int i = 0; i < kSlotWords / 4; ++i)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(row_sa + 16u * i), "l"(slot + 4 * i) : "memory");
}

// MODE: GEN_ALL / GEN_BASE / GEN_MULTI_PLAIN (mgrl_core.cuh, generate): which problems are compiled in
template <int LAYOUT, bool SEE, int MODE>
__global__ void __launch_bounds__((kMaxStepWarps + kMaxGenWarps) * 32, 1) rollout_kernel(const EnvParams p, int sw, int gw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int PITCH = obs_pitch(LAYOUT);
    const int gen_words = gen_words_for(p.cfg);
    const RolloutSmem L = rollout_smem(sw, gw, gen_words);
    uint32_t* state = reinterpret_cast<uint32_t*>(smem_raw + L.state);
    uint16_t* queue = reinterpret_cast<uint16_t*>(smem_raw + L.queue);
    uint8_t* ready = smem_raw + L.ready;                                     // [kDepth][tile_envs]
    uint32_t* kind_lut = reinterpret_cast<uint32_t*>(smem_raw + L.kind_lut);
    float* lut = reinterpret_cast<float*>(smem_raw + L.lut);
    uint32_t* empty = reinterpret_cast<uint32_t*>(smem_raw + L.empty);
    uint32_t* prefix = reinterpret_cast<uint32_t*>(smem_raw + L.prefix);
    RolloutCtrl& ctrl = *reinterpret_cast<RolloutCtrl*>(smem_raw + L.ctrl);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthreads = blockDim.x;
    const int tile_envs = sw * 32;
    const int tile0 = blockIdx.x * tile_envs;
    const int nv = min(tile_envs, p.n - tile0);
    const int S = p.cfg.size;
    const uint32_t qmask = (uint32_t)L.qcap - 1u;

    {   // ---- prologue: states [env][35 words] -> rows of kRowWords, constants, tags
        const uint32_t* g = reinterpret_cast<const uint32_t*>(p.states + tile0);
        for (int i = tid; i < nv * STATE_WORDS; i += nthreads) {
            const int r = i / STATE_WORDS;
            state[r * kRowWords + (i - r * STATE_WORDS)] = g[i];
        }
        for (int r = tid; r < tile_envs; r += nthreads) state[r * kRowWords + STATE_WORDS] = 0u;
        fill_kind_lut(kind_lut, tid, nthreads);
        for (int i = tid; i <= p.cfg.max_steps; i += nthreads) lut[i] = p.reward_lut[i];
        for (int i = tid; i < kGridWords; i += nthreads) empty[i] = p.empty[i];
        for (int i = tid; i < kTaskWords; i += nthreads) prefix[i] = p.prefix[i];
        const uint32_t nq0 = p.rq_count[blockIdx.x];   // requests the previous launch left queued
        for (int i = tid; i < L.qcap; i += nthreads)
            queue[i] = (uint32_t)i < nq0 ? p.rq_save[(size_t)blockIdx.x * L.qcap + i] : (uint16_t)kNoEntry;
        for (int i = tid; i < kDepth * tile_envs; i += nthreads) {
            const int j = i / tile_envs, e = i - j * tile_envs;
            ready[i] = e < nv ? p.tags[(size_t)j * p.n + tile0 + e] : (uint8_t)0;
        }
        if (tid == 0) { ctrl.q_head = 0u; ctrl.q_tail = nq0; ctrl.finished = 0; ctrl.starve = 0; }
    }
    __syncthreads();

    if (warp < sw) {
        // ================================================================ step warps
        const int e = warp * 32 + lane;
        const bool active = e < nv;
        uint32_t* cur = state + e * kRowWords;
        EnvState& s = *reinterpret_cast<EnvState*>(cur);
        uint8_t* stage = smem_raw + L.stage + warp * (32 * kObsPitch148);
        const int nvw = max(0, min(32, nv - warp * 32));
        const uint32_t row_sa = (uint32_t)__cvta_generic_to_shared(cur);
        const uint32_t stage_sa = (uint32_t)__cvta_generic_to_shared(stage);
        __builtin_assume(__isShared(stage));
        __builtin_assume(__isShared(cur));
        uint64_t policy = 0;   // the records are not read again by this kernel: keep them from displacing the layouts in L2
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        bool store_in_flight = false;   // (lane 0) a bulk copy may still be reading the staging area
        int action = (active && p.T > 0) ? p.actions[tile0 + e] : 0;
        for (int t = 0; t < p.T; ++t) {
            const size_t gi = (size_t)t * (size_t)p.n + (size_t)(tile0 + e);
            const int a = action;
            if (active && t + 1 < p.T) action = p.actions[gi + (size_t)p.n];   // next step's action, in flight during this one
            bool done = false, fetched = false;
            int carry = 0;
            StepOut o;
            o.w32 = 0u; o.error = 0;
            uint32_t E = 0, old34 = 0;
            int j = 0;
            if (active) {
                const StepIn in = env_step_load(s, S);
                E = cur[33]; old34 = cur[34];
                j = (int)(E & (uint32_t)(kDepth - 1));
                const uint32_t* slot = slot_ptr(p.slots, j, p.n, tile0 + e);
                // `done` always ends the episode: everything the step needs is in registers now, so the next layout can
                // start to overwrite the row while the step is computed
                if (a == A_DONE && *reinterpret_cast<const volatile uint8_t*>(&ready[j * tile_envs + e]) == (uint8_t)(E + 1u)) {
                    __threadfence_block();
                    fetch_layout(row_sa, slot);
                    fetched = true;
                }
                o = env_step_apply<true, MODE == GEN_ALL>(s, in, a, p.cfg.max_steps, lut);
                p.reward[gi] = o.reward;
                p.term[gi] = o.terminated;
                p.trunc[gi] = o.truncated;
                done = o.terminated | o.truncated;
                if (p.ep_len) p.ep_len[gi] = done ? o.step_count : (uint8_t)0;
                carry = o.carry_obs;
            }
            if (__any_sync(FULL, done)) {
                bool pending = done && !fetched;
                const volatile uint8_t* tag = &ready[j * tile_envs + e];
                for (int spins = 0; __any_sync(FULL, pending); ++spins) {
                    if (pending && *tag == (uint8_t)(E + 1u)) {
                        __threadfence_block();
                        fetch_layout(row_sa, slot_ptr(p.slots, j, p.n, tile0 + e));
                        fetched = true; pending = false;
                    }
                    if (pending) {   // the generator warps are behind: let them take what is queued, whatever the batch size
                        *reinterpret_cast<volatile int*>(&ctrl.starve) = 1;
                        __nanosleep(400);
                        if (spins > kSpinLimit) { s.error |= ERR_SYNC; pending = false; }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                const bool adopted = done && fetched;
                if (adopted) {
                    // what survives the reset: the mission latch (App. B Q1; zero after `done`) and the sticky error byte
                    if (o.w32 & 0xFFFF0000u) cur[32] = (cur[32] & 0x0000FFFFu) | (o.w32 & 0xFFFF0000u);
                    const uint32_t err = (old34 & 0x00FF0000u) | ((uint32_t)o.error << 16);
                    if (err) cur[34] |= err;
                    carry = 0;
                }
                const unsigned rm = __ballot_sync(FULL, adopted);
                if (rm) {   // ask for the layout that will follow the one just adopted (warp-aggregated push)
                    const int leader = __ffs(rm) - 1;
                    uint32_t base = 0;
                    if (lane == leader) base = atomicAdd(&ctrl.q_tail, (uint32_t)__popc(rm));
                    base = __shfl_sync(FULL, base, leader);
                    if (adopted) {
                        const uint32_t pos = base + (uint32_t)__popc(rm & ((1u << lane) - 1u));
                        *reinterpret_cast<volatile uint16_t*>(&queue[pos & qmask]) = (uint16_t)((uint32_t)e | ((uint32_t)j << kEnvBits));
                    }
                }
            }
            // the previous step's bulk copy must have read the staging area before it is rewritten
            if (store_in_flight) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); store_in_flight = false; }
            __syncwarp();
            if (active) {
                if (p.image) {
                    if (LAYOUT == OBS_HWC148 && SEE) {
                        asm volatile("" ::: "memory");   // the state bytes written above are read through asm loads
                        encode_packed_smem(row_sa, s.agent_x, s.agent_y, s.agent_dir, (uint32_t)carry, S,
                                           (uint32_t)__cvta_generic_to_shared(kind_lut), stage_sa + (uint32_t)(lane * PITCH));
                        asm volatile("" ::: "memory");
                    } else {
                        encode_view<LAYOUT>(s, carry, S, SEE, kind_lut, stage + lane * PITCH);
                    }
                }
                if (p.dir) p.dir[gi] = s.agent_dir;
                if (p.mission) p.mission[gi] = s.mission_id;
            }
            if (p.image && nvw > 0) {
                uint8_t* dst = p.image + ((size_t)t * (size_t)p.n + (size_t)(tile0 + warp * 32)) * PITCH;
                if (nvw == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    // generic-proxy writes of the staging area -> async proxy: fence by every writer, then one lane issues
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                                     ::"l"(dst), "r"(stage_sa), "r"(32 * PITCH), "l"(policy) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        store_in_flight = true;
                    }
                } else {
                    __syncwarp();
                    warp_copy_out<PITCH>(dst, stage, nvw, lane);
                    __syncwarp();
                }
            }
        }
        if (store_in_flight) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd(&ctrl.finished, 1); }
    } else {
        // ================================================================ generator warps
        const int g = warp - sw;
        uint32_t* scratch = reinterpret_cast<uint32_t*>(smem_raw + L.scratch) + g * (32 * kScratchWords);
        uint32_t* gen = reinterpret_cast<uint32_t*>(smem_raw + L.gen) + g * (32 * gen_words);
        int idle = 0;   // (lane 0) polls since this warp last had work
        for (;;) {
            int n_take = 0, h = 0;
            if (lane == 0) {
                const uint32_t head = *reinterpret_cast<volatile uint32_t*>(&ctrl.q_head);
                const uint32_t avail = *reinterpret_cast<volatile uint32_t*>(&ctrl.q_tail) - head;
                const bool fin = *reinterpret_cast<volatile int*>(&ctrl.finished) >= sw;
                // a partial batch wastes lanes: only when the rollout is over, or when a step lane waits for a layout and no
                // full batch has come together for a while (the whole tile may be waiting)
                const bool starve = *reinterpret_cast<volatile int*>(&ctrl.starve) != 0 && idle >= 8;
                n_take = avail >= 32u ? 32 : ((fin || starve) ? (int)avail : 0);
                // the step warps are done: what is still queued is left to the next launch, whose generator warps would
                // otherwise have nothing to do until its first episodes end (and this launch's step warps nothing now)
                const bool leave = fin && p.carry != 0;
                if (leave) n_take = 0;
                idle = n_take > 0 ? 0 : idle + 1;
                if (n_take > 0) {
                    if (atomicCAS(&ctrl.q_head, head, head + (uint32_t)n_take) != head) n_take = -1;   // another warp took them
                    else { h = (int)head; if (n_take < 32) *reinterpret_cast<volatile int*>(&ctrl.starve) = 0; }
                } else if (fin) {
                    n_take = -2;   // every step warp is done and the queue is empty (or left to the next launch)
                }
            }
            n_take = __shfl_sync(FULL, n_take, 0);
            if (n_take == -2) break;
            if (n_take <= 0) { if (n_take == 0) __nanosleep(100); continue; }
            h = __shfl_sync(FULL, h, 0);
            int e = 0, j = 0;
            uint32_t tagw = 0;
            uint32_t* sc = scratch + lane * kScratchWords;
            if (lane < n_take) {
                volatile uint16_t* q = reinterpret_cast<volatile uint16_t*>(&queue[((uint32_t)h + (uint32_t)lane) & qmask]);
                uint32_t ent;
                while ((ent = *q) == kNoEntry) {}
                *q = (uint16_t)kNoEntry;
                e = (int)(ent & ((1u << kEnvBits) - 1u)); j = (int)(ent >> kEnvBits);
                // The slot's old layout was adopted, so it gets the one episode in [E, E + kDepth) that maps to slot j, E being
                // the environment's next episode (E may advance while we look: every value it can take gives the same answer,
                // because the environment cannot adopt the layout we are about to build).
                const uint32_t E = *reinterpret_cast<const volatile uint32_t*>(&state[e * kRowWords + 33]);
                const uint32_t episode = E + (((uint32_t)j - E) & (uint32_t)(kDepth - 1));
                generate_layout<MODE>(sc, p.cfg, p.seed, p.env_id_base + (uint64_t)(tile0 + e), episode, gen + lane, p.tasks,
                                p.task_row_words, prefix, empty);
                tagw = sc[33];
            }
            __syncwarp();
            // write the n_take slot records (144 B each) with coalesced 16-byte stores: chunk c = record c / 9, part c % 9
            const int chunks = n_take * (kSlotWords / 4);
            for (int c0 = 0; c0 < chunks; c0 += 32) {
                const int c = c0 + lane;
                const bool valid = c < chunks;
                const int r = valid ? c / (kSlotWords / 4) : 0;
                const int q4 = c - r * (kSlotWords / 4);
                const int er = __shfl_sync(FULL, e, r), jr = __shfl_sync(FULL, j, r);
                if (valid) {
                    // the next-to-a-door marks are stripped on the way out (grid words 0..30); word 35 is the record's pad
                    const uint32_t* src = scratch + r * kScratchWords + 4 * q4;
                    uint4 v;
                    v.x = src[0]; v.y = src[1]; v.z = src[2]; v.w = q4 == 8 ? 0u : src[3];
                    if (q4 < 8) { v.x &= kMarkMask; v.y &= kMarkMask; v.z &= kMarkMask; if (q4 < 7) v.w &= kMarkMask; }
                    __stcg(reinterpret_cast<uint4*>(slot_ptr(p.slots, jr, p.n, tile0 + er)) + q4, v);
                }
            }
            __threadfence_block();
            __syncwarp();
            if (lane < n_take) {
                // hand-off inside the CTA: layout (global, L2) -> fence -> tag byte (shared); the adopting lane reads the
                // tag byte, fences, then reads the layout
                p.tags[(size_t)j * p.n + tile0 + e] = (uint8_t)tagw;
                __threadfence_block();
                *reinterpret_cast<volatile uint8_t*>(&ready[j * tile_envs + e]) = (uint8_t)tagw;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    {   // requests still queued travel to the next launch (none when carry = 0: the generator warps drained the queue)
        const uint32_t head = ctrl.q_head, nq = ctrl.q_tail - head;
        for (uint32_t i = tid; i < nq; i += nthreads) p.rq_save[(size_t)blockIdx.x * L.qcap + i] = queue[(head + i) & qmask];
        if (tid == 0) p.rq_count[blockIdx.x] = nq;
    }
    {   // ---- epilogue: rows -> states
        uint32_t* g = reinterpret_cast<uint32_t*>(p.states + tile0);
        for (int i = tid; i < nv * STATE_WORDS; i += nthreads) {
            const int r = i / STATE_WORDS;
            g[i] = state[r * kRowWords + (i - r * STATE_WORDS)];
        }
    }
}

// Dense layout generation for the requests recorded by one-step launches: every warp takes batches of 32 entries of
// the global list (grid-stride), one lane per layout, all lanes busy.
template <int NWARPS>
struct GenSmem {
    alignas(16) uint32_t scratch[NWARPS][32 * STATE_WORDS];
    alignas(16) uint32_t draws[NWARPS][kGenWords * 32];
    uint32_t empty[kGridWords];
    uint32_t prefix[kTaskWords];
};

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) generate_kernel(const EnvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GenSmem<NWARPS>& sm = *reinterpret_cast<GenSmem<NWARPS>*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kGridWords; i += NWARPS * 32) sm.empty[i] = p.empty[i];
    for (int i = tid; i < kTaskWords; i += NWARPS * 32) sm.prefix[i] = p.prefix[i];
    __syncthreads();
    const uint32_t count = *p.gcount;
    const uint32_t nwarps = gridDim.x * NWARPS;
    for (uint32_t b = blockIdx.x * NWARPS + warp; b * 32u < count; b += nwarps) {
        const uint32_t r = b * 32u + (uint32_t)lane;
        if (r < count) {
            const uint32_t ent = p.glist[r];
            const uint32_t env = ent / kDepth, j = ent % kDepth;
            uint32_t* slot = slot_ptr(p.slots, (int)j, p.n, (int)env);
            const uint32_t episode = slot[33] - 1u + kDepth;   // the slot's old layout was adopted
            uint32_t* sc = sm.scratch[warp] + lane * STATE_WORDS;
            sc[32] = 0u; sc[34] = 0u;
            GenIO io;
            io.draws = sm.draws[warp] + lane; io.stride = 32; io.tasks = p.tasks; io.row_words = p.task_row_words;
            io.prefix = sm.prefix; io.empty = sm.empty; io.keep_marks = true;
            generate(*reinterpret_cast<EnvState*>(sc), p.cfg, p.seed, p.env_id_base + (uint64_t)env, episode, io);
#pragma unroll
            for (int i = 0; i < STATE_WORDS; ++i) slot[i] = i < kGridWords ? sc[i] & kMarkMask : sc[i];
            p.tags[(size_t)j * p.n + env] = (uint8_t)sc[33];
        }
        __syncwarp();
    }
}

// reset (PRIME = false): fresh environments, episode 0 built in place, episodes 1..kDepth into the slots,
// first observation.  PRIME = true (after mgrl_set_state): only the slots, for episodes E..E+kDepth-1.
template <int LAYOUT, bool SEE, bool PRIME, int TILE>
struct ResetSmem {
    alignas(16) uint32_t state[TILE * STATE_WORDS];
    alignas(16) uint8_t obs[TILE / 32][32 * kObsPitch148];
    alignas(16) uint32_t draws[TILE / 32][kGenWords * 32];
    uint32_t kind_lut[128];
    uint32_t empty[kGridWords];
    uint32_t prefix[kTaskWords];
};

template <int LAYOUT, bool SEE, bool PRIME, int TILE>
__global__ void __launch_bounds__(TILE) reset_kernel(const EnvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = ResetSmem<LAYOUT, SEE, PRIME, TILE>;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    constexpr int PITCH = obs_pitch(LAYOUT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile0 = blockIdx.x * TILE;
    const int nv = min(TILE, p.n - tile0);
    if (PRIME) coop_copy<TILE>(sm.state, p.states + tile0, nv * (int)sizeof(EnvState), tid);
    fill_kind_lut(sm.kind_lut, tid, TILE);
    for (int i = tid; i < kGridWords; i += TILE) sm.empty[i] = p.empty[i];
    for (int i = tid; i < kTaskWords; i += TILE) sm.prefix[i] = p.prefix[i];
    if (tid == 0) p.qcount[blockIdx.x] = 0u;
    __syncthreads();
    const bool active = tid < nv;
    uint32_t* cur = sm.state + tid * STATE_WORDS;
    EnvState& s = *reinterpret_cast<EnvState*>(cur);
    GenIO io;
    io.draws = sm.draws[warp] + lane; io.stride = 32; io.tasks = p.tasks; io.row_words = p.task_row_words;
    io.prefix = sm.prefix; io.empty = sm.empty;
    const uint64_t env_id = p.env_id_base + (uint64_t)(tile0 + tid);
    if (active) {
        if (!PRIME) {
#pragma unroll
            for (int i = 0; i < STATE_WORDS; ++i) cur[i] = 0u;
            generate(s, p.cfg, p.seed, env_id, 0u, io);
        }
        uint32_t* sc = reinterpret_cast<uint32_t*>(sm.obs[warp]) + lane * STATE_WORDS;
        for (int k = 0; k < kDepth; ++k) {
            const uint32_t episode = s.episode + (uint32_t)k;
            sc[32] = 0u; sc[34] = 0u;
            generate(*reinterpret_cast<EnvState*>(sc), p.cfg, p.seed, env_id, episode, io);   // (marks stripped inside)
            uint32_t* slot = slot_ptr(p.slots, (int)(episode % kDepth), p.n, tile0 + tid);
#pragma unroll
            for (int i = 0; i < STATE_WORDS; ++i) slot[i] = sc[i];
            p.tags[(size_t)(episode % kDepth) * p.n + tile0 + tid] = (uint8_t)sc[33];
        }
    }
    if (PRIME) return;
    __syncwarp();
    uint8_t* stage = sm.obs[warp];
    if (active) {
        if (p.image) encode_view<LAYOUT>(s, 0, p.cfg.size, SEE, sm.kind_lut, stage + lane * PITCH);
        if (p.dir) p.dir[tile0 + tid] = s.agent_dir;
        if (p.mission) p.mission[tile0 + tid] = s.mission_id;
    }
    __syncthreads();
    const int nvw = max(0, min(32, nv - warp * 32));
    if (p.image && nvw > 0) coop_copy<32>(p.image + (size_t)(tile0 + warp * 32) * PITCH, stage, nvw * PITCH, lane);
    coop_copy<TILE>(p.states + tile0, sm.state, nv * (int)sizeof(EnvState), tid);
}

// gen_obs of the current states without stepping
template <int LAYOUT, bool SEE, int TILE>
__global__ void __launch_bounds__(TILE) observe_kernel(const EnvParams p) {
    __shared__ uint32_t kind_lut[128];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * TILE + tid;
    constexpr int PITCH = obs_pitch(LAYOUT);
    fill_kind_lut(kind_lut, tid, TILE);
    __syncthreads();
    if (i >= p.n) return;
    EnvState s = p.states[i];
    uint32_t rec[kObsPitch148 / 4];
    uint8_t* out = reinterpret_cast<uint8_t*>(rec);
    if (p.image) {
        encode_view<LAYOUT>(s, s.carrying, p.cfg.size, SEE, kind_lut, out);
        uint8_t* dst = p.image + (size_t)i * PITCH;
        for (int b = 0; b < PITCH; ++b) dst[b] = out[b];
    }
    if (p.dir) p.dir[i] = s.agent_dir;
    if (p.mission) p.mission[i] = s.mission_id;
}

__global__ void full_obs_kernel(const EnvState* __restrict__ states, int n, int S, uint8_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) encode_full(states[i], S, out + (size_t)i * S * S * 3);
}

__global__ void error_flags_kernel(const EnvState* __restrict__ states, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int e = i < n ? states[i].error : 0;
    e = __reduce_or_sync(0xffffffffu, e);
    if ((threadIdx.x & 31) == 0 && e) atomicOr(out, e);
}

// VecFrameStack(4,'first') over image / direction one-hot / mission tokens.
// One thread per (env, element): 147 image bytes, 4 direction bytes, 32 mission tokens.
constexpr int STACK_ELEMS = kObsBytes + 4 + MGRL_MISSION_TOKENS;
__global__ void stack_push_kernel(int n, const uint8_t* __restrict__ image, const uint8_t* __restrict__ dir,
                                  const uint8_t* __restrict__ mission, const uint8_t* __restrict__ done,
                                  const int64_t* __restrict__ table, uint8_t* __restrict__ s_img,
                                  uint8_t* __restrict__ s_dir, int64_t* __restrict__ s_mis) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * STACK_ELEMS) return;
    const int e = (int)(idx / STACK_ELEMS), j = (int)(idx % STACK_ELEMS);
    const bool clear = done == nullptr || done[e] != 0;
    if (j < kObsBytes) {
        if (!s_img) return;
        uint8_t* b = s_img + (size_t)e * (MGRL_FRAMES * kObsBytes) + j;
        const uint8_t f1 = b[kObsBytes], f2 = b[2 * kObsBytes], f3 = b[3 * kObsBytes];
        b[0] = clear ? 0 : f1; b[kObsBytes] = clear ? 0 : f2; b[2 * kObsBytes] = clear ? 0 : f3;
        b[3 * kObsBytes] = image[(size_t)e * kObsBytes + j];
    } else if (j < kObsBytes + 4) {
        if (!s_dir) return;
        const int k = j - kObsBytes;
        uint8_t* b = s_dir + (size_t)e * 16 + k;
        const uint8_t f1 = b[4], f2 = b[8], f3 = b[12];
        b[0] = clear ? 0 : f1; b[4] = clear ? 0 : f2; b[8] = clear ? 0 : f3;
        b[12] = dir[e] == k;
    } else {
        if (!s_mis) return;
        const int k = j - kObsBytes - 4;
        int64_t* b = s_mis + (size_t)e * (MGRL_FRAMES * MGRL_MISSION_TOKENS) + k;
        const int64_t f1 = b[32], f2 = b[64], f3 = b[96];
        b[0] = clear ? 0 : f1; b[32] = clear ? 0 : f2; b[64] = clear ? 0 : f3;
        b[96] = table[(int)mission[e] * MGRL_MISSION_TOKENS + k];
    }
}

// SB3 GAE: one lane per environment walks the time axis backwards.  The loads of a step do not depend on the recurrence:
// they are issued kGaeGroup steps at a time into register arrays (3 x 16 loads in flight per lane; left to the unroller
// the compiler kept the loop at 32 registers and a couple of loads in flight: 0.41 of the HBM roofline), then the group's
// recurrence runs on registers.  __f*_rn intrinsics pin the float32 operation order (no FMA contraction) -> bit-exact, which
// is why the time axis is not a parallel scan (a scan re-associates the sums).  Blocks of kGaeBlock environments: 65 536
// environments are 1024 blocks = 6.9 per SM.
constexpr int kGaeBlock = 64;
constexpr int kGaeGroup = 16;
__global__ void __launch_bounds__(kGaeBlock) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const uint8_t* __restrict__ starts,
                                                  const float* __restrict__ last_values,
                                                  const uint8_t* __restrict__ last_dones, float g, float gl, int T,
                                                  int N, float* __restrict__ adv, float* __restrict__ ret) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float A = 0.0f;
    float nv = last_values[n];
    float nnt = 1.0f - (float)last_dones[n];
    for (int t0 = T - 1; t0 >= 0; t0 -= kGaeGroup) {
        float vt[kGaeGroup], rt[kGaeGroup];
        uint8_t st[kGaeGroup];
#pragma unroll
        for (int u = 0; u < kGaeGroup; ++u) {
            const int t = t0 - u;
            const size_t i = (size_t)(t < 0 ? 0 : t) * N + n;
            vt[u] = __ldcs(values + i); rt[u] = __ldcs(rewards + i); st[u] = __ldcs(starts + i);
        }
#pragma unroll
        for (int u = 0; u < kGaeGroup; ++u) {
            const int t = t0 - u;
            if (t >= 0) {
                const size_t i = (size_t)t * N + n;
                const float delta = __fsub_rn(__fadd_rn(rt[u], __fmul_rn(__fmul_rn(g, nv), nnt)), vt[u]);
                A = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnt), A));
                __stcs(adv + i, A);
                __stcs(ret + i, __fadd_rn(A, vt[u]));
                nv = vt[u];
                nnt = 1.0f - (float)st[u];
            }
        }
    }
}

// ------------------------------------------------------------------------------------ host side
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof g_err, fmt, detail);
    return code;
}

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            snprintf(g_err, sizeof g_err, "%s: %s", #expr, cudaGetErrorString(_e)); \
            return MGRL_ERR_CUDA;                                                   \
        }                                                                           \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

}  // namespace

// the library's one thread-local error text, shared with the policy translation units
char* mgrl_error_buffer() { return g_err; }

struct mgrl_env {
    mgrl_config cfg;
    EnvCfg ecfg;
    int device;
    int tile;    // environments per CTA (64 / 128) of the one-step kernel
    int n_sms;   // multiprocessors of the device (tile shape of the rollout kernel)
    uint16_t* rq_save;   // rollout_kernel's carried layout requests [tiles][qcap] and their counts [tiles]
    uint32_t* rq_count;
    int rq_tiles;
    mutable bool rollout_pending;   // the last rollout launch left requests queued: drained before anything but another rollout
    int rollout; // 1: mgrl_step_many runs rollout_kernel; 0 (MGRL_ROLLOUT=0): the one-step kernel with T steps per launch
    uint64_t seed;
    EnvState* states;
    uint32_t* slots;     // [kDepth][N][kSlotWords]
    uint8_t* tags;       // [kDepth][N]
    uint16_t* qsave;     // [tiles(64)][kQueueCap]
    uint32_t* qcount;    // [tiles(64)]
    uint32_t* glist;     // [kDepth * N] deferred requests of one-step launches
    uint32_t* gcount;    // [1]
    int deferred_steps;  // one-step launches since generate_kernel last ran
    uint32_t* tasks;
    int task_row_words;
    uint32_t* prefix;
    uint32_t* empty;
    float* lut;
    int* err_flags;
    // host-path (VecEnv drop-in) buffers, allocated on first use
    bool host_ready;
    uint8_t *h_actions, *h_image, *h_dir, *h_mission, *h_term, *h_trunc, *h_eplen, *h_termimg, *h_termdir, *h_stack_img, *h_stack_dir;
    float* h_reward;
    int64_t *h_stack_mis, *h_table;
    uint8_t* h_full;     // full-grid observation staging of mgrl_full_obs_host
    bool table_set;
    int64_t table_host[MGRL_N_MISSIONS * MGRL_MISSION_TOKENS];   // host copy for the in-place stack of mgrl_vec_step_stacked_host
    bool stack_on_host;          // the newest observation stack lives in the caller's arrays (mgrl_vec_step_stacked_host), not in h_stack_*
    const void* stack_ptrs[3];   // the arrays the last reset / stacked step filled
    mgrl_wire::Path* wire;   // compact wire format + expansion pool of mgrl_vec_step_frames_host (mgrl_wire.cu), created on first use
};

namespace {

EnvParams make_params(const mgrl_env* e) {
    EnvParams p;
    memset(&p, 0, sizeof p);
    p.cfg = e->ecfg;
    p.seed = e->seed;
    p.env_id_base = e->cfg.env_id_base;
    p.n = e->cfg.num_envs;
    p.T = 1;
    p.states = e->states;
    p.slots = e->slots;
    p.tags = e->tags;
    p.qsave = e->qsave;
    p.qcount = e->qcount;
    p.glist = e->glist;
    p.gcount = e->gcount;
    p.tasks = e->tasks;
    p.task_row_words = e->task_row_words;
    p.prefix = e->prefix;
    p.empty = e->empty;
    p.reward_lut = e->lut;
    p.rq_save = e->rq_save;
    p.rq_count = e->rq_count;
    return p;
}

constexpr int MODE_STEP = 0, MODE_RESET = 1, MODE_PRIME = 2, MODE_OBSERVE = 3;

template <typename K>
int launch_kernel(K kernel, size_t smem, int grid, int block, cudaStream_t s, const EnvParams& p) {
    // > 48 KB of dynamic shared memory needs the opt-in attribute (idempotent, cheap)
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, block, smem, s>>>(p);
    CUDA_TRY(cudaGetLastError());
    return MGRL_OK;
}

template <int LAYOUT, bool SEE, int TILE, int NB>
int launch_tile(int mode, const EnvParams& p, cudaStream_t s) {
    const int grid = (p.n + TILE - 1) / TILE;
    if (mode == MODE_STEP)
        return launch_kernel(step_kernel<LAYOUT, SEE, TILE, NB>, sizeof(TileSmem<TILE, NB>), grid, TILE, s, p);
    if (mode == MODE_RESET)
        return launch_kernel(reset_kernel<LAYOUT, SEE, false, TILE>, sizeof(ResetSmem<LAYOUT, SEE, false, TILE>), grid, TILE, s, p);
    if (mode == MODE_PRIME)
        return launch_kernel(reset_kernel<LAYOUT, SEE, true, TILE>, sizeof(ResetSmem<LAYOUT, SEE, true, TILE>), grid, TILE, s, p);
    return launch_kernel(observe_kernel<LAYOUT, SEE, TILE>, 0, grid, TILE, s, p);
}

template <int LAYOUT, bool SEE>
int launch_layout(const mgrl_env* e, int mode, const EnvParams& p, cudaStream_t s) {
    if (e->tile == 64) return launch_tile<LAYOUT, SEE, 64, 1>(mode, p, s);
    return launch_tile<LAYOUT, SEE, 128, 2>(mode, p, s);
}

int drain_rollout(const mgrl_env* e, cudaStream_t s);

int launch_env(const mgrl_env* e, int mode, const EnvParams& p, cudaStream_t s) {
    if (e->rollout_pending) {
        if (mode == MODE_RESET || mode == MODE_PRIME) {   // every slot is rebuilt: the carried requests are void
            CUDA_TRY(cudaMemsetAsync(e->rq_count, 0, (size_t)e->rq_tiles * sizeof(uint32_t), s));
            e->rollout_pending = false;
        } else {
            const int rc = drain_rollout(e, s);
            if (rc) return rc;
        }
    }
    const bool see = e->ecfg.see_through_walls != 0;
    switch (e->cfg.obs_layout) {
    case MGRL_OBS_CHW:
        return see ? launch_layout<OBS_CHW, true>(e, mode, p, s) : launch_layout<OBS_CHW, false>(e, mode, p, s);
    case MGRL_OBS_HWC148:
        return see ? launch_layout<OBS_HWC148, true>(e, mode, p, s) : launch_layout<OBS_HWC148, false>(e, mode, p, s);
    default:
        return see ? launch_layout<OBS_HWC, true>(e, mode, p, s) : launch_layout<OBS_HWC, false>(e, mode, p, s);
    }
}

// tile shape of the rollout kernel: one tile per SM when the environments fit one wave (14 step warps = 448 environments at
// most), 5 generator warps per 7 step warps (what shared memory allows next to 14 step warps); MGRL_SW / MGRL_GW override
constexpr int kRolloutSmemMax = 227 * 1024;   // opt-in dynamic shared memory per CTA
void rollout_shape(const mgrl_env* e, int* sw_out, int* gw_out) {
    const int per_sm = (e->cfg.num_envs + e->n_sms - 1) / e->n_sms;
    int sw = (per_sm + 31) / 32;
    sw = sw < 1 ? 1 : (sw > kMaxStepWarps ? kMaxStepWarps : sw);
    // the generator warps are the limiter (one layout per seven env-steps of a uniform-random rollout costs more than the
    // seven steps): as many as the shared memory and the register file (24 warps of 80 registers; an eleventh generator warp
    // fits the memory but caps the kernel at 72 registers, which costs what the warp brings) take
    int gw = sw + 2;
    if (const char* v = getenv("MGRL_SW")) sw = atoi(v);
    if (const char* v = getenv("MGRL_GW")) gw = atoi(v);
    sw = sw < 1 ? 1 : (sw > kMaxStepWarps ? kMaxStepWarps : sw);
    gw = gw < 1 ? 1 : (gw > kMaxGenWarps ? kMaxGenWarps : gw);
    while (gw > 1 && rollout_smem(sw, gw, gen_words_for(e->ecfg)).total > kRolloutSmemMax) --gw;
    *sw_out = sw; *gw_out = gw;
}

template <int LAYOUT, bool SEE>
int launch_rollout_t(const mgrl_env* e, const EnvParams& p, cudaStream_t s) {
    int sw, gw;
    rollout_shape(e, &sw, &gw);
    const RolloutSmem L = rollout_smem(sw, gw, gen_words_for(e->ecfg));
    const int grid = (p.n + sw * 32 - 1) / (sw * 32);
    // the instance with the least code that covers the configuration (the kernel is bound by instruction issue)
    const int mode = e->ecfg.problem >= P_MOV ? GEN_ALL
                   : (e->ecfg.problem == P_MULTI && e->ecfg.num_obstacles == 0 && !e->ecfg.all_doors_open) ? GEN_MULTI_PLAIN : GEN_BASE;
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t ce = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
        if (ce != cudaSuccess) return ce;
        kernel<<<grid, (sw + gw) * 32, L.total, s>>>(p, sw, gw);
        return cudaSuccess;
    };
    if (mode == GEN_ALL) CUDA_TRY(launch(rollout_kernel<LAYOUT, SEE, GEN_ALL>));
    else if (mode == GEN_MULTI_PLAIN) CUDA_TRY(launch(rollout_kernel<LAYOUT, SEE, GEN_MULTI_PLAIN>));
    else CUDA_TRY(launch(rollout_kernel<LAYOUT, SEE, GEN_BASE>));
    CUDA_TRY(cudaGetLastError());
    return MGRL_OK;
}

int launch_rollout(const mgrl_env* e, const EnvParams& p, cudaStream_t s) {
    const bool see = e->ecfg.see_through_walls != 0;
    switch (e->cfg.obs_layout) {
    case MGRL_OBS_CHW:
        return see ? launch_rollout_t<OBS_CHW, true>(e, p, s) : launch_rollout_t<OBS_CHW, false>(e, p, s);
    case MGRL_OBS_HWC148:
        return see ? launch_rollout_t<OBS_HWC148, true>(e, p, s) : launch_rollout_t<OBS_HWC148, false>(e, p, s);
    default:
        return see ? launch_rollout_t<OBS_HWC, true>(e, p, s) : launch_rollout_t<OBS_HWC, false>(e, p, s);
    }
}

// the layouts a rollout launch left to its successor, when the successor is not a rollout: a launch without steps
int drain_rollout(const mgrl_env* e, cudaStream_t s) {
    if (!e->rollout_pending) return MGRL_OK;
    EnvParams p = make_params(e);
    p.T = 0; p.carry = 0;
    const int rc = launch_rollout(e, p, s);
    if (rc == MGRL_OK) e->rollout_pending = false;
    return rc;
}

// build the layouts requested by the one-step launches since the last flush (dense generate_kernel)
int flush_deferred(mgrl_env* e, cudaStream_t s) {
    if (e->deferred_steps == 0) return MGRL_OK;
    constexpr int NWARPS = 4;
    const EnvParams p = make_params(e);
    // sized for a quarter of the environments finishing on every step; the kernel strides over anything beyond that
    long long batches = ((long long)e->cfg.num_envs * e->deferred_steps / 4 + 31) / 32;
    int grid = (int)((batches + NWARPS - 1) / NWARPS);
    grid = grid < 1 ? 1 : (grid > 148 * 4 ? 148 * 4 : grid);
    const size_t smem = sizeof(GenSmem<NWARPS>);
    CUDA_TRY(cudaFuncSetAttribute(generate_kernel<NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generate_kernel<NWARPS><<<grid, NWARPS * 32, smem, s>>>(p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemsetAsync(e->gcount, 0, sizeof(uint32_t), s));
    e->deferred_steps = 0;
    return MGRL_OK;
}

int ensure_host_buffers(mgrl_env* e) {
    if (e->host_ready) return MGRL_OK;
    const size_t n = (size_t)e->cfg.num_envs;
    CUDA_TRY(cudaMalloc(&e->h_actions, n));
    CUDA_TRY(cudaMalloc(&e->h_image, n * kObsPitch148));
    // the small per-step outputs live in one slab (reward | dir | mission | term | trunc | ep_len | term_dir) so that a
    // caller whose host buffers are laid out the same way gets them with ONE device-to-host copy
    CUDA_TRY(cudaMalloc(&e->h_reward, n * 10));
    e->h_dir = reinterpret_cast<uint8_t*>(e->h_reward) + 4 * n;
    e->h_mission = e->h_dir + n; e->h_term = e->h_mission + n; e->h_trunc = e->h_term + n; e->h_eplen = e->h_trunc + n;
    e->h_termdir = e->h_eplen + n;
    CUDA_TRY(cudaMalloc(&e->h_termimg, n * kObsPitch148));
    CUDA_TRY(cudaMemset(e->h_termdir, 0, n));
    CUDA_TRY(cudaMalloc(&e->h_stack_img, n * MGRL_FRAMES * kObsBytes));
    CUDA_TRY(cudaMalloc(&e->h_stack_dir, n * 16));
    CUDA_TRY(cudaMalloc(&e->h_stack_mis, n * MGRL_FRAMES * MGRL_MISSION_TOKENS * sizeof(int64_t)));
    CUDA_TRY(cudaMemset(e->h_termimg, 0, n * kObsPitch148));
    e->host_ready = true;
    return MGRL_OK;
}

int launch_stack(int n, const uint8_t* image, const uint8_t* dir, const uint8_t* mission, const uint8_t* done,
                 const int64_t* table, uint8_t* s_img, uint8_t* s_dir, int64_t* s_mis, cudaStream_t s) {
    const size_t total = (size_t)n * STACK_ELEMS;
    const int threads = 256;
    const unsigned grid = (unsigned)((total + threads - 1) / threads);
    stack_push_kernel<<<grid, threads, 0, s>>>(n, image, dir, mission, done, table, s_img, s_dir, s_mis);
    CUDA_TRY(cudaGetLastError());
    return MGRL_OK;
}

}  // namespace

extern "C" {

int mgrl_abi_version(void) { return MGRL_ABI_VERSION; }
const char* mgrl_last_error(void) { return g_err; }

int mgrl_create(const mgrl_config* cfg, int device, mgrl_env** out) {
    if (!cfg || !out) return fail(MGRL_ERR_INVALID, "mgrl_create: null argument%s");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MGRL_ERR_NO_DEVICE, "mgrl_create: no CUDA device (this library has no CPU path)%s");
    }
    if (device < 0 || device >= ndev) return fail(MGRL_ERR_INVALID, "mgrl_create: bad device index%s");
    if (cfg->size < 5 || cfg->size > kMaxSize) return fail(MGRL_ERR_INVALID, "mgrl_create: size must be in 5..11%s");
    if (cfg->problem < MGRL_MULTI || cfg->problem > MGRL_FULL)
        return fail(MGRL_ERR_INVALID, "mgrl_create: unknown problem%s");
    if (!(cfg->mission == -1 || cfg->mission == 0 || cfg->mission == 1 || cfg->mission == 2 || cfg->mission == 5))
        return fail(MGRL_ERR_INVALID, "mgrl_create: mission must be 0, 1, 2, 5 or -1 (null)%s");
    if (cfg->num_envs <= 0) return fail(MGRL_ERR_INVALID, "mgrl_create: num_envs must be positive%s");
    if (cfg->num_objects < 0 || cfg->num_objects > 18)
        return fail(MGRL_ERR_INVALID, "mgrl_create: num_objects must be in 0..18%s");
    if (cfg->num_obstacles < 0 || cfg->num_obstacles > 8)
        return fail(MGRL_ERR_INVALID, "mgrl_create: num_obstacles must be in 0..8 (task rows hold 31 placements)%s");
    if (cfg->obs_layout != MGRL_OBS_HWC && cfg->obs_layout != MGRL_OBS_CHW && cfg->obs_layout != MGRL_OBS_HWC148)
        return fail(MGRL_ERR_INVALID, "mgrl_create: bad obs_layout%s");
    DeviceGuard guard(device);
    mgrl_env* e = new (std::nothrow) mgrl_env();
    if (!e) return fail(MGRL_ERR_INVALID, "mgrl_create: out of host memory%s");
    memset(e, 0, sizeof *e);
    e->cfg = *cfg;
    e->device = device;
    // 64 environments (2 warps, one draw buffer) per CTA: 7 CTAs per SM, so 65 536 environments spread evenly over the
    // 148 SMs (6.9 tiles each); 128 per CTA (4 warps, two draw buffers) is the alternative shape (MGRL_TILE=128)
    e->tile = 64;
    if (const char* t = getenv("MGRL_TILE")) e->tile = atoi(t) == 128 ? 128 : 64;
    e->n_sms = 148;
    cudaDeviceGetAttribute(&e->n_sms, cudaDevAttrMultiProcessorCount, device);
    e->rollout = 1;
    if (const char* t = getenv("MGRL_ROLLOUT")) e->rollout = atoi(t) != 0;
    e->ecfg.size = cfg->size;
    e->ecfg.num_objects = cfg->num_objects;
    e->ecfg.problem = cfg->problem;
    e->ecfg.mission = cfg->mission;
    e->ecfg.all_doors_open = cfg->all_doors_open;
    e->ecfg.see_through_walls = cfg->see_through_walls;
    e->ecfg.max_steps = cfg->max_steps > 0 ? cfg->max_steps : cfg->size * cfg->size;
    e->ecfg.num_obstacles = cfg->num_obstacles;
    if (e->ecfg.max_steps > kGridCells) {
        delete e;
        return fail(MGRL_ERR_INVALID, "mgrl_create: max_steps above 121 is not supported%s");
    }
    // reward LUT: float32(1 - 0.9*(k/max_steps)) evaluated in float64 like MiniGridEnv._reward
    float lut[kGridCells + 1];
    for (int k = 0; k <= kGridCells; ++k) {
        volatile double q = (double)k / (double)e->ecfg.max_steps;
        volatile double m = 0.9 * q;
        volatile double r = 1.0 - m;
        lut[k] = (float)r;
    }
    // generator inputs: task strings of this num_objects and the fresh grid of this size
    static thread_local uint32_t full_table[kTaskEntries * kTaskWords], tasks[kTaskEntries * kTaskWords];
    uint32_t empty[kGridWords], prefix[kTaskWords];
    build_task_table(e->ecfg, full_table);
    e->task_row_words = pack_task_table(full_table, tasks);
    build_task_prefix(e->ecfg, prefix);
    build_empty_grid(cfg->size, empty);
    const size_t n_tiles = ((size_t)cfg->num_envs + 63) / 64;
    cudaError_t err = cudaMalloc(&e->states, (size_t)cfg->num_envs * sizeof(EnvState));
    if (err == cudaSuccess) err = cudaMemset(e->states, 0, (size_t)cfg->num_envs * sizeof(EnvState));
    if (err == cudaSuccess) err = cudaMalloc(&e->slots, (size_t)kDepth * cfg->num_envs * kSlotWords * sizeof(uint32_t));
    if (err == cudaSuccess) err = cudaMemset(e->slots, 0, (size_t)kDepth * cfg->num_envs * kSlotWords * sizeof(uint32_t));
    if (err == cudaSuccess) err = cudaMalloc(&e->tags, (size_t)kDepth * cfg->num_envs);
    if (err == cudaSuccess) err = cudaMemset(e->tags, 0, (size_t)kDepth * cfg->num_envs);
    if (err == cudaSuccess) err = cudaMalloc(&e->qsave, n_tiles * kQueueCap * sizeof(uint16_t));
    if (err == cudaSuccess) err = cudaMalloc(&e->qcount, n_tiles * sizeof(uint32_t));
    if (err == cudaSuccess) err = cudaMemset(e->qcount, 0, n_tiles * sizeof(uint32_t));
    {   // the rollout kernel's carried requests: one queue image per tile, whatever its shape (MGRL_SW may change it):
        // tiles(sw) * qcap(sw) <= (n / (32 sw) + 1) * 2 * 32 sw * kDepth
        const size_t entries = 2 * (size_t)kDepth * (size_t)cfg->num_envs + 2 * (size_t)kMaxStepWarps * 32 * kDepth;
        e->rq_tiles = (cfg->num_envs + 31) / 32;
        if (err == cudaSuccess) err = cudaMalloc(&e->rq_save, entries * sizeof(uint16_t));
        if (err == cudaSuccess) err = cudaMalloc(&e->rq_count, (size_t)e->rq_tiles * sizeof(uint32_t));
        if (err == cudaSuccess) err = cudaMemset(e->rq_count, 0, (size_t)e->rq_tiles * sizeof(uint32_t));
    }
    if (err == cudaSuccess) err = cudaMalloc(&e->glist, (size_t)kDepth * cfg->num_envs * sizeof(uint32_t));
    if (err == cudaSuccess) err = cudaMalloc(&e->gcount, sizeof(uint32_t));
    if (err == cudaSuccess) err = cudaMemset(e->gcount, 0, sizeof(uint32_t));
    if (err == cudaSuccess) err = cudaMalloc(&e->tasks, sizeof tasks);
    if (err == cudaSuccess) err = cudaMemcpy(e->tasks, tasks, sizeof tasks, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMalloc(&e->prefix, sizeof prefix);
    if (err == cudaSuccess) err = cudaMemcpy(e->prefix, prefix, sizeof prefix, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMalloc(&e->empty, sizeof empty);
    if (err == cudaSuccess) err = cudaMemcpy(e->empty, empty, sizeof empty, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMalloc(&e->lut, sizeof lut);
    if (err == cudaSuccess) err = cudaMemcpy(e->lut, lut, sizeof lut, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMalloc(&e->err_flags, sizeof(int));
    if (err != cudaSuccess) {
        snprintf(g_err, sizeof g_err, "mgrl_create: %s", cudaGetErrorString(err));
        mgrl_destroy(e);
        return MGRL_ERR_CUDA;
    }
    *out = e;
    return MGRL_OK;
}

int mgrl_destroy(mgrl_env* e) {
    if (!e) return MGRL_OK;
    DeviceGuard guard(e->device);
    void* bufs[] = {e->rq_save, e->rq_count, e->states, e->slots, e->tags, e->qsave, e->qcount, e->glist, e->gcount, e->tasks, e->prefix, e->empty, e->lut, e->err_flags, e->h_actions, e->h_image,
                    e->h_termimg, e->h_reward, e->h_stack_img, e->h_stack_dir,
                    e->h_stack_mis, e->h_table, e->h_full};
    for (void* b : bufs)
        if (b) cudaFree(b);
    mgrl_wire::destroy(e->wire);
    delete e;
    return MGRL_OK;
}

int mgrl_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(MGRL_ERR_INVALID, "mgrl_host_alloc: null argument%s");
    CUDA_TRY(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return MGRL_OK;
}
int mgrl_host_free(void* ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return MGRL_OK;
}

int mgrl_reset(mgrl_env* e, uint64_t seed, uint8_t* image, uint8_t* dir, uint8_t* mission, void* stream) {
    if (!e) return fail(MGRL_ERR_INVALID, "mgrl_reset: null handle%s");
    DeviceGuard guard(e->device);
    e->seed = seed;
    e->deferred_steps = 0;
    CUDA_TRY(cudaMemsetAsync(e->gcount, 0, sizeof(uint32_t), (cudaStream_t)stream));
    EnvParams p = make_params(e);
    p.image = image; p.dir = dir; p.mission = mission;
    return launch_env(e, MODE_RESET, p, (cudaStream_t)stream);
}

int mgrl_step(mgrl_env* e, const uint8_t* actions, uint8_t* image, uint8_t* dir, uint8_t* mission, float* reward,
              uint8_t* term, uint8_t* trunc, uint8_t* ep_len, uint8_t* term_image, uint8_t* term_dir, void* stream) {
    if (!e) return fail(MGRL_ERR_INVALID, "mgrl_step: null handle%s");
    if (!actions || !reward || !term || !trunc)
        return fail(MGRL_ERR_INVALID, "mgrl_step: actions, reward, term and trunc are required%s");
    DeviceGuard guard(e->device);
    EnvParams p = make_params(e);
    p.actions = actions; p.image = image; p.dir = dir; p.mission = mission; p.reward = reward;
    p.term = term; p.trunc = trunc; p.ep_len = ep_len; p.term_image = term_image; p.term_dir = term_dir;
    // a one-step launch only records which layouts it used up; generate_kernel rebuilds them, densely, every
    // kDepth steps (an environment cannot use more than kDepth layouts in between)
    p.defer = 1;
    int rc = launch_env(e, MODE_STEP, p, (cudaStream_t)stream);
    if (rc) return rc;
    if (++e->deferred_steps >= kDepth) rc = flush_deferred(e, (cudaStream_t)stream);
    return rc;
}

int mgrl_step_many(mgrl_env* e, int T, const uint8_t* actions, uint8_t* image, uint8_t* dir, uint8_t* mission,
                   float* reward, uint8_t* term, uint8_t* trunc, uint8_t* ep_len, void* stream) {
    if (!e) return fail(MGRL_ERR_INVALID, "mgrl_step_many: null handle%s");
    if (T <= 0 || !actions || !reward || !term || !trunc)
        return fail(MGRL_ERR_INVALID, "mgrl_step_many: T>0, actions, reward, term and trunc are required%s");
    DeviceGuard guard(e->device);
    int rc = flush_deferred(e, (cudaStream_t)stream);   // layouts owed to earlier one-step launches first
    if (rc) return rc;
    EnvParams p = make_params(e);
    p.T = T;
    p.actions = actions; p.image = image; p.dir = dir; p.mission = mission; p.reward = reward;
    p.term = term; p.trunc = trunc; p.ep_len = ep_len;
    if (e->rollout) {
        // MGRL_CARRY=1: leave the layout requests that are still queued when the step warps finish to the next launch (drained
        // by a launch without steps before anything that is not a rollout).  Off by default: measured 13.05 against 13.01 G
        // env-steps/s - the generator warps are the limiter either way, the work only moves from one launch's tail to the next.
        static const bool carry = [] { const char* v = getenv("MGRL_CARRY"); return v && v[0] == '1'; }();
        p.carry = carry ? 1 : 0;
        rc = launch_rollout(e, p, (cudaStream_t)stream);
        if (rc == MGRL_OK && carry) e->rollout_pending = true;
        return rc;
    }
    return launch_env(e, MODE_STEP, p, (cudaStream_t)stream);
}

int mgrl_observe(mgrl_env* e, uint8_t* image, uint8_t* dir, uint8_t* mission, void* stream) {
    if (!e) return fail(MGRL_ERR_INVALID, "mgrl_observe: null handle%s");
    DeviceGuard guard(e->device);
    EnvParams p = make_params(e);
    p.image = image; p.dir = dir; p.mission = mission;
    return launch_env(e, MODE_OBSERVE, p, (cudaStream_t)stream);
}

int mgrl_get_state(mgrl_env* e, void* dst, size_t bytes, void* stream) {
    if (!e || !dst) return fail(MGRL_ERR_INVALID, "mgrl_get_state: null argument%s");
    if (bytes != (size_t)e->cfg.num_envs * sizeof(EnvState)) return fail(MGRL_ERR_INVALID, "mgrl_get_state: bad size%s");
    DeviceGuard guard(e->device);
    CUDA_TRY(cudaMemcpyAsync(dst, e->states, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return MGRL_OK;
}

int mgrl_set_state(mgrl_env* e, const void* src, size_t bytes, uint64_t seed, void* stream) {
    if (!e || !src) return fail(MGRL_ERR_INVALID, "mgrl_set_state: null argument%s");
    if (bytes != (size_t)e->cfg.num_envs * sizeof(EnvState)) return fail(MGRL_ERR_INVALID, "mgrl_set_state: bad size%s");
    DeviceGuard guard(e->device);
    e->seed = seed;
    e->deferred_steps = 0;
    CUDA_TRY(cudaMemsetAsync(e->gcount, 0, sizeof(uint32_t), (cudaStream_t)stream));
    CUDA_TRY(cudaMemcpyAsync(e->states, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return launch_env(e, MODE_PRIME, make_params(e), (cudaStream_t)stream);   // layouts of the next episodes
}

int mgrl_get_state_host(mgrl_env* e, void* dst, size_t bytes, void* stream) {
    if (!e || !dst) return fail(MGRL_ERR_INVALID, "mgrl_get_state_host: null argument%s");
    if (bytes != (size_t)e->cfg.num_envs * sizeof(EnvState))
        return fail(MGRL_ERR_INVALID, "mgrl_get_state_host: bad size%s");
    DeviceGuard guard(e->device);
    CUDA_TRY(cudaMemcpyAsync(dst, e->states, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return MGRL_OK;
}

int mgrl_set_state_host(mgrl_env* e, const void* src, size_t bytes, uint64_t seed, void* stream) {
    if (!e || !src) return fail(MGRL_ERR_INVALID, "mgrl_set_state_host: null argument%s");
    if (bytes != (size_t)e->cfg.num_envs * sizeof(EnvState))
        return fail(MGRL_ERR_INVALID, "mgrl_set_state_host: bad size%s");
    DeviceGuard guard(e->device);
    e->seed = seed;
    e->deferred_steps = 0;
    CUDA_TRY(cudaMemsetAsync(e->gcount, 0, sizeof(uint32_t), (cudaStream_t)stream));
    CUDA_TRY(cudaMemcpyAsync(e->states, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    const int rc = launch_env(e, MODE_PRIME, make_params(e), (cudaStream_t)stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return MGRL_OK;
}

int mgrl_state_ptr(mgrl_env* e, void** state_dev) {
    if (!e || !state_dev) return fail(MGRL_ERR_INVALID, "mgrl_state_ptr: null argument%s");
    *state_dev = e->states;
    return MGRL_OK;
}

int mgrl_full_obs(mgrl_env* e, uint8_t* image, void* stream) {
    if (!e || !image) return fail(MGRL_ERR_INVALID, "mgrl_full_obs: null argument%s");
    DeviceGuard guard(e->device);
    const int n = e->cfg.num_envs;
    full_obs_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e->states, n, e->ecfg.size, image);
    CUDA_TRY(cudaGetLastError());
    return MGRL_OK;
}

int mgrl_full_obs_host(mgrl_env* e, uint8_t* image_host, void* stream) {
    if (!e || !image_host) return fail(MGRL_ERR_INVALID, "mgrl_full_obs_host: null argument%s");
    DeviceGuard guard(e->device);
    const size_t bytes = (size_t)e->cfg.num_envs * e->ecfg.size * e->ecfg.size * 3;
    if (!e->h_full) CUDA_TRY(cudaMalloc(&e->h_full, bytes));
    const int rc = mgrl_full_obs(e, e->h_full, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(image_host, e->h_full, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return MGRL_OK;
}

int mgrl_error_flags(mgrl_env* e, int* flags_out, void* stream) {
    if (!e || !flags_out) return fail(MGRL_ERR_INVALID, "mgrl_error_flags: null argument%s");
    DeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int n = e->cfg.num_envs;
    CUDA_TRY(cudaMemsetAsync(e->err_flags, 0, sizeof(int), s));
    error_flags_kernel<<<(n + 255) / 256, 256, 0, s>>>(e->states, n, e->err_flags);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(flags_out, e->err_flags, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return MGRL_OK;
}

int mgrl_stack_push(int num_envs, const uint8_t* image, const uint8_t* dir, const uint8_t* mission,
                    const uint8_t* done, const int64_t* table, uint8_t* s_img, uint8_t* s_dir, int64_t* s_mis,
                    void* stream) {
    if (num_envs <= 0) return fail(MGRL_ERR_INVALID, "mgrl_stack_push: num_envs must be positive%s");
    if ((s_img && !image) || (s_dir && !dir) || (s_mis && (!mission || !table)))
        return fail(MGRL_ERR_INVALID, "mgrl_stack_push: a stack was given without its source%s");
    return launch_stack(num_envs, image, dir, mission, done, table, s_img, s_dir, s_mis, (cudaStream_t)stream);
}

int mgrl_gae(const float* rewards, const float* values, const uint8_t* starts, const float* last_values,
             const uint8_t* last_dones, double gamma, double gae_lambda, int T, int N, float* adv, float* ret,
             void* stream) {
    if (!rewards || !values || !starts || !last_values || !last_dones || !adv || !ret || T <= 0 || N <= 0)
        return fail(MGRL_ERR_INVALID, "mgrl_gae: null argument or empty shape%s");
    // SB3 multiplies float32 arrays by Python floats: gamma -> f32; gamma*lambda in f64 -> f32
    const float g = (float)gamma;
    const float gl = (float)(gamma * gae_lambda);
    gae_kernel<<<(N + kGaeBlock - 1) / kGaeBlock, kGaeBlock, 0, (cudaStream_t)stream>>>(rewards, values, starts, last_values, last_dones, g,
                                                                  gl, T, N, adv, ret);
    CUDA_TRY(cudaGetLastError());
    return MGRL_OK;
}

int mgrl_set_token_table(mgrl_env* e, const int64_t* table_host) {
    if (!e || !table_host) return fail(MGRL_ERR_INVALID, "mgrl_set_token_table: null argument%s");
    DeviceGuard guard(e->device);
    const size_t bytes = (size_t)MGRL_N_MISSIONS * MGRL_MISSION_TOKENS * sizeof(int64_t);
    if (!e->h_table) CUDA_TRY(cudaMalloc(&e->h_table, bytes));
    CUDA_TRY(cudaMemcpy(e->h_table, table_host, bytes, cudaMemcpyHostToDevice));
    memcpy(e->table_host, table_host, bytes);
    e->table_set = true;
    return MGRL_OK;
}

static int copy_stacked_out(mgrl_env* e, uint8_t* image_host, uint8_t* direction_host, int64_t* mission_host,
                            cudaStream_t s) {
    const size_t n = (size_t)e->cfg.num_envs;
    if (image_host)
        CUDA_TRY(cudaMemcpyAsync(image_host, e->h_stack_img, n * MGRL_FRAMES * kObsBytes, cudaMemcpyDeviceToHost, s));
    if (direction_host) CUDA_TRY(cudaMemcpyAsync(direction_host, e->h_stack_dir, n * 16, cudaMemcpyDeviceToHost, s));
    if (mission_host)
        CUDA_TRY(cudaMemcpyAsync(mission_host, e->h_stack_mis, n * MGRL_FRAMES * MGRL_MISSION_TOKENS * sizeof(int64_t),
                                 cudaMemcpyDeviceToHost, s));
    return MGRL_OK;
}

int mgrl_vec_reset_host(mgrl_env* e, uint64_t seed, uint8_t* image_host, uint8_t* direction_host,
                        int64_t* mission_host, void* stream) {
    if (!e) return fail(MGRL_ERR_INVALID, "mgrl_vec_reset_host: null handle%s");
    if (!e->table_set) return fail(MGRL_ERR_INVALID, "mgrl_vec_reset_host: call mgrl_set_token_table first%s");
    if (e->cfg.obs_layout == MGRL_OBS_HWC148)
        return fail(MGRL_ERR_INVALID, "mgrl_vec_reset_host: the host path needs a 147-byte layout (CHW or HWC)%s");
    DeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_host_buffers(e);
    if (rc) return rc;
    rc = mgrl_reset(e, seed, e->h_image, e->h_dir, e->h_mission, stream);
    if (rc) return rc;
    rc = launch_stack(e->cfg.num_envs, e->h_image, e->h_dir, e->h_mission, nullptr, e->h_table, e->h_stack_img,
                      e->h_stack_dir, e->h_stack_mis, s);
    if (rc) return rc;
    rc = copy_stacked_out(e, image_host, direction_host, mission_host, s);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(s));
    e->stack_on_host = false;
    e->stack_ptrs[0] = image_host; e->stack_ptrs[1] = direction_host; e->stack_ptrs[2] = mission_host;
    if (e->wire) mgrl_wire::reset_stacked(e->wire);
    return MGRL_OK;
}

int mgrl_vec_reset_frames_host(mgrl_env* e, uint64_t seed, uint8_t* image_host, uint8_t* dir_host, uint8_t* mission_host,
                               void* stream) {
    if (!e || !image_host || !dir_host || !mission_host) return fail(MGRL_ERR_INVALID, "mgrl_vec_reset_frames_host: null argument%s");
    DeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_host_buffers(e);
    if (rc) return rc;
    rc = mgrl_reset(e, seed, e->h_image, e->h_dir, e->h_mission, stream);
    if (rc) return rc;
    const size_t n = (size_t)e->cfg.num_envs, pitch = e->cfg.obs_layout == MGRL_OBS_HWC148 ? kObsPitch148 : kObsBytes;
    CUDA_TRY(cudaMemcpyAsync(image_host, e->h_image, n * pitch, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(dir_host, e->h_dir, n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(mission_host, e->h_mission, n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return MGRL_OK;
}

int mgrl_vec_step_frames_host(mgrl_env* e, const uint8_t* actions_host, uint8_t* image_host, uint8_t* dir_host,
                              uint8_t* mission_host, float* reward_host, uint8_t* term_host, uint8_t* trunc_host,
                              uint8_t* ep_len_host, uint8_t* term_image_host, uint8_t* term_dir_host, void* stream) {
    if (!e || !actions_host || !image_host || !dir_host || !mission_host || !reward_host || !term_host || !trunc_host)
        return fail(MGRL_ERR_INVALID, "mgrl_vec_step_frames_host: null argument%s");
    if (!e->host_ready) return fail(MGRL_ERR_INVALID, "mgrl_vec_step_frames_host: call mgrl_vec_reset_frames_host first%s");
    DeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)e->cfg.num_envs, pitch = e->cfg.obs_layout == MGRL_OBS_HWC148 ? kObsPitch148 : kObsBytes;
    CUDA_TRY(cudaMemcpyAsync(e->h_actions, actions_host, n, cudaMemcpyHostToDevice, s));
    int rc = mgrl_step(e, e->h_actions, e->h_image, e->h_dir, e->h_mission, e->h_reward, e->h_term, e->h_trunc,
                       e->h_eplen, term_image_host ? e->h_termimg : nullptr, term_dir_host ? e->h_termdir : nullptr, stream);
    if (rc) return rc;
    // Default: the observation crosses PCIe as 64-byte records (one code byte per view cell + the step's scalars) and is
    // expanded into the caller's arrays by the handle's host threads while later chunks are still in flight (mgrl_wire.cu);
    // MGRL_WIRE=0 copies the full 148-byte records instead.
    static const bool wire_on = [] { const char* v = getenv("MGRL_WIRE"); return !(v && v[0] == '0'); }();
    if (wire_on) {
        if (!e->wire) {
            e->wire = mgrl_wire::create(e->cfg.num_envs);
            if (!e->wire) return fail(MGRL_ERR_CUDA, "mgrl_vec_step_frames_host: wire staging allocation failed%s");
        }
        mgrl_wire::Outputs o = {};
        o.layout = e->cfg.obs_layout; o.image_dev = e->h_image; o.image_host = image_host; o.reward_dev = e->h_reward; o.reward_host = reward_host;
        o.dir_dev = e->h_dir; o.mission_dev = e->h_mission; o.term_dev = e->h_term; o.trunc_dev = e->h_trunc; o.eplen_dev = e->h_eplen;
        o.tdir_dev = term_dir_host ? e->h_termdir : nullptr;
        o.dir_host = dir_host; o.mission_host = mission_host; o.term_host = term_host; o.trunc_host = trunc_host; o.eplen_host = ep_len_host;
        o.tdir_host = term_dir_host;
        mgrl_wire::Outputs x = {};
        x.layout = e->cfg.obs_layout; x.image_dev = e->h_termimg; x.image_host = term_image_host;
        CUDA_TRY(mgrl_wire::step(e->wire, o, term_image_host ? &x : nullptr, s));
        return MGRL_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(image_host, e->h_image, n * pitch, cudaMemcpyDeviceToHost, s));
    uint8_t* hb = reinterpret_cast<uint8_t*>(reward_host);
    const bool slab = dir_host == hb + 4 * n && mission_host == dir_host + n && term_host == mission_host + n &&
                      trunc_host == term_host + n && ep_len_host == trunc_host + n &&
                      (!term_dir_host || term_dir_host == ep_len_host + n);
    if (slab) {   // host buffers mirror the device slab: one copy for everything but the images
        CUDA_TRY(cudaMemcpyAsync(reward_host, e->h_reward, n * (term_dir_host ? 10 : 9), cudaMemcpyDeviceToHost, s));
    } else {
        CUDA_TRY(cudaMemcpyAsync(dir_host, e->h_dir, n, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(mission_host, e->h_mission, n, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(reward_host, e->h_reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(term_host, e->h_term, n, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(trunc_host, e->h_trunc, n, cudaMemcpyDeviceToHost, s));
        if (ep_len_host) CUDA_TRY(cudaMemcpyAsync(ep_len_host, e->h_eplen, n, cudaMemcpyDeviceToHost, s));
        if (term_dir_host) CUDA_TRY(cudaMemcpyAsync(term_dir_host, e->h_termdir, n, cudaMemcpyDeviceToHost, s));
    }
    if (term_image_host) CUDA_TRY(cudaMemcpyAsync(term_image_host, e->h_termimg, n * pitch, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return MGRL_OK;
}

int mgrl_vec_step_host(mgrl_env* e, const uint8_t* actions_host, uint8_t* image_host, uint8_t* direction_host,
                       int64_t* mission_host, float* reward_host, uint8_t* term_host, uint8_t* trunc_host,
                       uint8_t* ep_len_host, uint8_t* term_image_host, uint8_t* term_dir_host, void* stream) {
    if (!e || !actions_host || !reward_host || !term_host || !trunc_host)
        return fail(MGRL_ERR_INVALID, "mgrl_vec_step_host: null argument%s");
    if (!e->host_ready) return fail(MGRL_ERR_INVALID, "mgrl_vec_step_host: call mgrl_vec_reset_host first%s");
    if (!e->table_set) return fail(MGRL_ERR_INVALID, "mgrl_vec_step_host: call mgrl_set_token_table first%s");
    if (e->cfg.obs_layout == MGRL_OBS_HWC148)
        return fail(MGRL_ERR_INVALID, "mgrl_vec_step_host: the host path needs a 147-byte layout (CHW or HWC)%s");
    if (e->stack_on_host)
        return fail(MGRL_ERR_INVALID, "mgrl_vec_step_host: the stack of this handle lives in host arrays since mgrl_vec_step_stacked_host; "
                                      "call mgrl_vec_reset_host first%s");
    DeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)e->cfg.num_envs;
    CUDA_TRY(cudaMemcpyAsync(e->h_actions, actions_host, n, cudaMemcpyHostToDevice, s));
    int rc = mgrl_step(e, e->h_actions, e->h_image, e->h_dir, e->h_mission, e->h_reward, e->h_term, e->h_trunc,
                       e->h_eplen, term_image_host ? e->h_termimg : nullptr, term_dir_host ? e->h_termdir : nullptr, stream);
    if (rc) return rc;
    // done flag for the frame stack = term | trunc; h_eplen is non-zero exactly on done steps
    rc = launch_stack(e->cfg.num_envs, e->h_image, e->h_dir, e->h_mission, e->h_eplen, e->h_table, e->h_stack_img,
                      e->h_stack_dir, e->h_stack_mis, s);
    if (rc) return rc;
    rc = copy_stacked_out(e, image_host, direction_host, mission_host, s);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(reward_host, e->h_reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(term_host, e->h_term, n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(trunc_host, e->h_trunc, n, cudaMemcpyDeviceToHost, s));
    if (ep_len_host) CUDA_TRY(cudaMemcpyAsync(ep_len_host, e->h_eplen, n, cudaMemcpyDeviceToHost, s));
    if (term_image_host)
        CUDA_TRY(cudaMemcpyAsync(term_image_host, e->h_termimg, n * kObsBytes, cudaMemcpyDeviceToHost, s));
    if (term_dir_host) CUDA_TRY(cudaMemcpyAsync(term_dir_host, e->h_termdir, n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    e->stack_ptrs[0] = image_host; e->stack_ptrs[1] = direction_host; e->stack_ptrs[2] = mission_host;
    return MGRL_OK;
}

int mgrl_vec_step_stacked_host(mgrl_env* e, const uint8_t* actions_host, uint8_t* image_host, uint8_t* direction_host,
                               int64_t* mission_host, float* reward_host, uint8_t* term_host, uint8_t* trunc_host,
                               uint8_t* ep_len_host, uint8_t* term_image_host, uint8_t* term_direction_host,
                               int64_t* term_mission_host, void* stream) {
    const char* what = "mgrl_vec_step_stacked_host";
    if (!e || !actions_host || !image_host || !direction_host || !mission_host || !reward_host || !term_host || !trunc_host)
        return fail(MGRL_ERR_INVALID, "%s: null argument", what);
    if (!e->host_ready || !e->table_set) return fail(MGRL_ERR_INVALID, "%s: call mgrl_set_token_table and mgrl_vec_reset_host first", what);
    if (e->cfg.obs_layout == MGRL_OBS_HWC148) return fail(MGRL_ERR_INVALID, "%s: the host path needs a 147-byte layout (CHW or HWC)", what);
    if (e->stack_ptrs[0] != image_host || e->stack_ptrs[1] != direction_host || e->stack_ptrs[2] != mission_host)
        return fail(MGRL_ERR_INVALID, "%s: the observation stack is updated in place: pass the arrays the previous reset / step filled", what);
    DeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)e->cfg.num_envs;
    if (!e->wire) {
        e->wire = mgrl_wire::create(e->cfg.num_envs);
        if (!e->wire) return fail(MGRL_ERR_CUDA, "%s: wire staging allocation failed", what);
    }
    CUDA_TRY(cudaMemcpyAsync(e->h_actions, actions_host, n, cudaMemcpyHostToDevice, s));
    const bool want_term = term_image_host != nullptr;
    int rc = mgrl_step(e, e->h_actions, e->h_image, e->h_dir, e->h_mission, e->h_reward, e->h_term, e->h_trunc, e->h_eplen,
                       want_term ? e->h_termimg : nullptr, e->h_termdir, stream);
    if (rc) return rc;
    mgrl_wire::Outputs o = {};
    o.layout = e->cfg.obs_layout; o.image_dev = e->h_image; o.reward_dev = e->h_reward; o.reward_host = reward_host;
    o.dir_dev = e->h_dir; o.mission_dev = e->h_mission; o.term_dev = e->h_term; o.trunc_dev = e->h_trunc; o.eplen_dev = e->h_eplen;
    o.tdir_dev = e->h_termdir;
    o.term_host = term_host; o.trunc_host = trunc_host; o.eplen_host = ep_len_host;
    mgrl_wire::Outputs x = {};
    x.layout = e->cfg.obs_layout; x.image_dev = e->h_termimg;
    mgrl_wire::Stacked st = {};
    st.image = image_host; st.direction = direction_host; st.mission = mission_host; st.term_image = term_image_host;
    st.term_direction = term_direction_host; st.term_mission = term_mission_host; st.table = e->table_host;
    e->stack_on_host = true;
    CUDA_TRY(mgrl_wire::step_stacked(e->wire, o, want_term ? &x : nullptr, st, s));
    return MGRL_OK;
}

}  // extern "C"
