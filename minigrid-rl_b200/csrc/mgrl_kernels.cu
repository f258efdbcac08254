// mgrl_kernels.cu — sm_100a kernels and the C ABI (include/mgrl.h) of the batched MiniGrid
// simulator.  See DESIGN.md for the data layout and the roofline of each kernel.
//
// Kernel map (SURVEY.md §2.1 ids):
//   env_kernel<LAYOUT,SEE,MODE>  K1+K2  step -> mission bookkeeping -> reward/done -> compacted
//                                       auto-reset (layout generation) -> 7x7x3 encode
//   env_many_kernel              K1+K2  T steps per launch, state tile resident in shared memory
//   gae_kernel                   K4     SB3 GAE reverse scan
//   stack_push_kernel                   VecFrameStack + Discrete2Box + TokenizeVocab gathers
//
// Thread mapping: one CTA = one tile of TILE consecutive environments, one lane per
// environment for the dynamics; the tile's packed state (TILE x 140 B) and its observation
// bytes (TILE x 147 B) are staged in shared memory so that every global access is a
// coalesced 16-byte vector transfer.  Finished environments are compacted per CTA so the
// (divergent, RNG-heavy) layout generator runs on dense lanes.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "mgrl.h"
#include "mgrl_core.cuh"

using namespace mgrl;

namespace {

constexpr int STATE_WORDS = 35;           // sizeof(EnvState) / 4
constexpr int MODE_STEP = 0, MODE_RESET = 1, MODE_OBSERVE = 2;

static_assert(sizeof(EnvState) == MGRL_STATE_BYTES, "ABI state size");

struct EnvParams {
    EnvCfg cfg;
    uint64_t seed;
    uint64_t env_id_base;
    int n;
    int T;       // env_many_kernel only
    int spread;  // 1: deal finished environments round-robin over the CTA's warps (shortest critical
                 //    path, for small batches); 0: pack them into the first warps (fewest issue slots)
    EnvState* states;
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
};

// ---- cooperative tile copies (coalesced; 16-byte vectors when size and address allow) -------
template <int TILE>
__device__ __forceinline__ void tile_copy(void* dst, const void* src, int bytes, int tid) {
    if ((bytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const uint4* g = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
        for (int i = tid; i < (bytes >> 4); i += TILE) d[i] = g[i];
    } else if ((bytes & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 3) == 0) {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (int i = tid; i < (bytes >> 2); i += TILE) d[i] = g[i];
    } else {
        const uint8_t* g = reinterpret_cast<const uint8_t*>(src);
        uint8_t* d = reinterpret_cast<uint8_t*>(dst);
        const int n16 = (reinterpret_cast<uintptr_t>(dst) & 15) == 0 ? (bytes >> 4) : 0;  // smem src is 16-B aligned
        for (int i = tid; i < n16; i += TILE) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(g)[i];
        for (int i = (n16 << 4) + tid; i < bytes; i += TILE) d[i] = g[i];
    }
}

template <int TILE>
struct TileSmem {
    alignas(16) uint32_t state[TILE * STATE_WORDS];
    alignas(16) uint8_t obs[TILE * kObsPitch148];
    uint32_t kind_lut[128];
    float lut[kGridCells + 1];
    uint16_t done_list[TILE];
    uint8_t carry[TILE];
    int n_done;
};

// one simulator step of the tile held in `sm` (state already resident); writes the per-step
// outputs of global env index tile0+tid; leaves the new observation records in sm.obs
template <int LAYOUT, bool SEE, int MODE, int TILE>
__device__ __forceinline__ void tile_step(TileSmem<TILE>& sm, const EnvParams& p, int tile0, int nv, int tid,
                                          size_t out_off /* element offset of this step's [N] outputs */) {
    constexpr int PITCH = obs_pitch(LAYOUT);
    EnvState* st = reinterpret_cast<EnvState*>(sm.state);
    const bool active = tid < nv;
    const int S = p.cfg.size;

    if (MODE == MODE_STEP) {
        if (active) {
            EnvState& s = st[tid];
            const int a = p.actions[out_off + tile0 + tid];
            const StepOut o = env_step(s, a, S, p.cfg.max_steps, sm.lut);
            const size_t gi = out_off + tile0 + tid;
            p.reward[gi] = o.reward;
            p.term[gi] = o.terminated;
            p.trunc[gi] = o.truncated;
            const bool done = o.terminated | o.truncated;
            if (p.ep_len) p.ep_len[gi] = done ? s.step_count : (uint8_t)0;
            sm.carry[tid] = o.carry_obs;
            if (done) sm.done_list[atomicAdd(&sm.n_done, 1)] = (uint16_t)tid;
        }
        __syncthreads();
        // pass over the finished environments only: terminal observation, then a new layout
        const int nd = sm.n_done;
        constexpr int NW = TILE / 32;
        const int slot = p.spread ? (tid & 31) * NW + (tid >> 5) : tid;  // which finished env this lane takes
        if (slot < nd) {
            const int e = sm.done_list[slot];
            EnvState& s = st[e];
            if (p.term_image)
                encode_view<LAYOUT>(s, sm.carry[e], S, SEE, sm.kind_lut,
                                    p.term_image + (out_off + tile0 + e) * PITCH);
            if (p.term_dir) p.term_dir[out_off + tile0 + e] = s.agent_dir;
            generate(s, p.cfg, p.seed, p.env_id_base + (uint64_t)(tile0 + e));
            sm.carry[e] = 0;
        }
        __syncthreads();
        if (tid == 0) sm.n_done = 0;
    } else if (MODE == MODE_RESET) {
        if (active) {
            uint32_t* w = sm.state + tid * STATE_WORDS;
#pragma unroll
            for (int i = 0; i < STATE_WORDS; ++i) w[i] = 0u;
            generate(st[tid], p.cfg, p.seed, p.env_id_base + (uint64_t)(tile0 + tid));
            sm.carry[tid] = 0;
        }
    } else {
        if (active) sm.carry[tid] = st[tid].carrying;
    }

    if (active) {
        const EnvState& s = st[tid];
        const size_t gi = out_off + tile0 + tid;
        if (p.image) encode_view<LAYOUT>(s, sm.carry[tid], S, SEE, sm.kind_lut, sm.obs + tid * PITCH);
        if (p.dir) p.dir[gi] = s.agent_dir;
        if (p.mission) p.mission[gi] = s.mission_id;
    }
    __syncthreads();
    if (p.image) tile_copy<TILE>(p.image + (out_off + tile0) * PITCH, sm.obs, nv * PITCH, tid);
}

template <int TILE>
__device__ __forceinline__ void tile_prologue(TileSmem<TILE>& sm, const EnvParams& p, int tid) {
    fill_kind_lut(sm.kind_lut, tid, TILE);
    for (int i = tid; i <= p.cfg.max_steps; i += TILE) sm.lut[i] = p.reward_lut[i];
    if (tid == 0) sm.n_done = 0;
}

template <int LAYOUT, bool SEE, int MODE, int TILE>
__global__ void __launch_bounds__(TILE) env_kernel(const EnvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem<TILE>& sm = *reinterpret_cast<TileSmem<TILE>*>(smem_raw);
    const int tid = threadIdx.x;
    const int tile0 = blockIdx.x * TILE;
    const int nv = min(TILE, p.n - tile0);
    if (MODE != MODE_RESET) tile_copy<TILE>(sm.state, p.states + tile0, nv * (int)sizeof(EnvState), tid);
    tile_prologue<TILE>(sm, p, tid);
    __syncthreads();
    tile_step<LAYOUT, SEE, MODE, TILE>(sm, p, tile0, nv, tid, 0);
    if (MODE != MODE_OBSERVE) tile_copy<TILE>(p.states + tile0, sm.state, nv * (int)sizeof(EnvState), tid);
}

// T steps per launch; the state tile never leaves shared memory between steps
template <int LAYOUT, bool SEE, int TILE>
__global__ void __launch_bounds__(TILE) env_many_kernel(const EnvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem<TILE>& sm = *reinterpret_cast<TileSmem<TILE>*>(smem_raw);
    const int tid = threadIdx.x;
    const int tile0 = blockIdx.x * TILE;
    const int nv = min(TILE, p.n - tile0);
    tile_copy<TILE>(sm.state, p.states + tile0, nv * (int)sizeof(EnvState), tid);
    tile_prologue<TILE>(sm, p, tid);
    __syncthreads();
    for (int t = 0; t < p.T; ++t) {
        tile_step<LAYOUT, SEE, MODE_STEP, TILE>(sm, p, tile0, nv, tid, (size_t)t * (size_t)p.n);
        __syncthreads();  // sm.obs is rewritten by the next step
    }
    tile_copy<TILE>(p.states + tile0, sm.state, nv * (int)sizeof(EnvState), tid);
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

// SB3 GAE: one lane per environment walks the time axis backwards; the loads of a step do
// not depend on the recurrence, so the unrolled loop keeps 8 steps of loads in flight.
// __f*_rn intrinsics pin the float32 operation order (no FMA contraction) -> bit-exact.
__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const uint8_t* __restrict__ starts,
                                                  const float* __restrict__ last_values,
                                                  const uint8_t* __restrict__ last_dones, float g, float gl, int T,
                                                  int N, float* __restrict__ adv, float* __restrict__ ret) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float A = 0.0f;
    float nv = last_values[n];
    float nnt = 1.0f - (float)last_dones[n];
#pragma unroll 8
    for (int t = T - 1; t >= 0; --t) {
        const size_t i = (size_t)t * N + n;
        const float vt = values[i], rt = rewards[i];
        const float st = (float)starts[i];
        const float delta = __fsub_rn(__fadd_rn(rt, __fmul_rn(__fmul_rn(g, nv), nnt)), vt);
        A = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnt), A));
        adv[i] = A;
        ret[i] = __fadd_rn(A, vt);
        nv = vt;
        nnt = 1.0f - st;
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

struct mgrl_env {
    mgrl_config cfg;
    EnvCfg ecfg;
    int device;
    int tile;    // environments per CTA (64 / 128 / 256)
    int spread;  // reset scheduling, see EnvParams
    uint64_t seed;
    EnvState* states;
    float* lut;
    int* err_flags;
    // host-path (VecEnv drop-in) buffers, allocated on first use
    bool host_ready;
    uint8_t *h_actions, *h_image, *h_dir, *h_mission, *h_term, *h_trunc, *h_eplen, *h_termimg, *h_termdir, *h_stack_img, *h_stack_dir;
    float* h_reward;
    int64_t *h_stack_mis, *h_table;
    bool table_set;
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
    p.spread = e->spread;
    p.states = e->states;
    p.reward_lut = e->lut;
    return p;
}

template <typename K>
int launch_kernel(K kernel, size_t smem, int grid, int block, cudaStream_t s, const EnvParams& p) {
    // > 48 KB of dynamic shared memory needs the opt-in attribute (idempotent, cheap)
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, block, smem, s>>>(p);
    CUDA_TRY(cudaGetLastError());
    return MGRL_OK;
}

template <int LAYOUT, bool SEE, int TILE>
int launch_tile(int mode, bool many, const EnvParams& p, cudaStream_t s) {
    const int grid = (p.n + TILE - 1) / TILE;
    const size_t smem = sizeof(TileSmem<TILE>);
    if (many) return launch_kernel(env_many_kernel<LAYOUT, SEE, TILE>, smem, grid, TILE, s, p);
    if (mode == MODE_STEP) return launch_kernel(env_kernel<LAYOUT, SEE, MODE_STEP, TILE>, smem, grid, TILE, s, p);
    if (mode == MODE_RESET) return launch_kernel(env_kernel<LAYOUT, SEE, MODE_RESET, TILE>, smem, grid, TILE, s, p);
    return launch_kernel(env_kernel<LAYOUT, SEE, MODE_OBSERVE, TILE>, smem, grid, TILE, s, p);
}

template <int LAYOUT, bool SEE>
int launch_layout(const mgrl_env* e, int mode, bool many, const EnvParams& p, cudaStream_t s) {
    if (e->tile == 256) return launch_tile<LAYOUT, SEE, 256>(mode, many, p, s);
    if (e->tile == 64) return launch_tile<LAYOUT, SEE, 64>(mode, many, p, s);
    return launch_tile<LAYOUT, SEE, 128>(mode, many, p, s);
}

int launch_env(const mgrl_env* e, int mode, bool many, const EnvParams& p, cudaStream_t s) {
    const bool see = e->ecfg.see_through_walls != 0;
    switch (e->cfg.obs_layout) {
    case MGRL_OBS_CHW:
        return see ? launch_layout<OBS_CHW, true>(e, mode, many, p, s) : launch_layout<OBS_CHW, false>(e, mode, many, p, s);
    case MGRL_OBS_HWC148:
        return see ? launch_layout<OBS_HWC148, true>(e, mode, many, p, s)
                   : launch_layout<OBS_HWC148, false>(e, mode, many, p, s);
    default:
        return see ? launch_layout<OBS_HWC, true>(e, mode, many, p, s) : launch_layout<OBS_HWC, false>(e, mode, many, p, s);
    }
}

int ensure_host_buffers(mgrl_env* e) {
    if (e->host_ready) return MGRL_OK;
    const size_t n = (size_t)e->cfg.num_envs;
    CUDA_TRY(cudaMalloc(&e->h_actions, n));
    CUDA_TRY(cudaMalloc(&e->h_image, n * kObsBytes));
    CUDA_TRY(cudaMalloc(&e->h_dir, n));
    CUDA_TRY(cudaMalloc(&e->h_mission, n));
    CUDA_TRY(cudaMalloc(&e->h_term, n));
    CUDA_TRY(cudaMalloc(&e->h_trunc, n));
    CUDA_TRY(cudaMalloc(&e->h_eplen, n));
    CUDA_TRY(cudaMalloc(&e->h_termimg, n * kObsBytes));
    CUDA_TRY(cudaMalloc(&e->h_termdir, n));
    CUDA_TRY(cudaMemset(e->h_termdir, 0, n));
    CUDA_TRY(cudaMalloc(&e->h_reward, n * sizeof(float)));
    CUDA_TRY(cudaMalloc(&e->h_stack_img, n * MGRL_FRAMES * kObsBytes));
    CUDA_TRY(cudaMalloc(&e->h_stack_dir, n * 16));
    CUDA_TRY(cudaMalloc(&e->h_stack_mis, n * MGRL_FRAMES * MGRL_MISSION_TOKENS * sizeof(int64_t)));
    CUDA_TRY(cudaMemset(e->h_termimg, 0, n * kObsBytes));
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
    if (cfg->problem < MGRL_MULTI || cfg->problem > MGRL_DRP)
        return fail(MGRL_ERR_INVALID, "mgrl_create: unsupported problem (full/mov are out of scope)%s");
    if (!(cfg->mission == -1 || cfg->mission == 0 || cfg->mission == 1 || cfg->mission == 2 || cfg->mission == 5))
        return fail(MGRL_ERR_INVALID, "mgrl_create: mission must be 0, 1, 2, 5 or -1 (null)%s");
    if (cfg->num_envs <= 0) return fail(MGRL_ERR_INVALID, "mgrl_create: num_envs must be positive%s");
    if (cfg->num_objects < 0 || cfg->num_objects > 18)
        return fail(MGRL_ERR_INVALID, "mgrl_create: num_objects must be in 0..18%s");
    if (cfg->obs_layout != MGRL_OBS_HWC && cfg->obs_layout != MGRL_OBS_CHW && cfg->obs_layout != MGRL_OBS_HWC148)
        return fail(MGRL_ERR_INVALID, "mgrl_create: bad obs_layout%s");
    DeviceGuard guard(device);
    mgrl_env* e = new (std::nothrow) mgrl_env();
    if (!e) return fail(MGRL_ERR_INVALID, "mgrl_create: out of host memory%s");
    memset(e, 0, sizeof *e);
    e->cfg = *cfg;
    e->device = device;
    // launch shape: small batches are latency-bound (deal resets over all warps), large ones issue-bound
    e->tile = 128;
    e->spread = cfg->num_envs <= 262144 ? 1 : 0;
    if (const char* t = getenv("MGRL_TILE")) e->tile = atoi(t) == 256 ? 256 : atoi(t) == 64 ? 64 : 128;
    if (const char* t = getenv("MGRL_SPREAD")) e->spread = atoi(t) != 0;
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
    cudaError_t err = cudaMalloc(&e->states, (size_t)cfg->num_envs * sizeof(EnvState));
    if (err == cudaSuccess) err = cudaMemset(e->states, 0, (size_t)cfg->num_envs * sizeof(EnvState));
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
    void* bufs[] = {e->states, e->lut, e->err_flags, e->h_actions, e->h_image, e->h_dir, e->h_mission, e->h_term,
                    e->h_trunc, e->h_eplen, e->h_termimg, e->h_termdir, e->h_reward, e->h_stack_img, e->h_stack_dir,
                    e->h_stack_mis, e->h_table};
    for (void* b : bufs)
        if (b) cudaFree(b);
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
    EnvParams p = make_params(e);
    p.image = image; p.dir = dir; p.mission = mission;
    return launch_env(e, MODE_RESET, false, p, (cudaStream_t)stream);
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
    return launch_env(e, MODE_STEP, false, p, (cudaStream_t)stream);
}

int mgrl_step_many(mgrl_env* e, int T, const uint8_t* actions, uint8_t* image, uint8_t* dir, uint8_t* mission,
                   float* reward, uint8_t* term, uint8_t* trunc, uint8_t* ep_len, void* stream) {
    if (!e) return fail(MGRL_ERR_INVALID, "mgrl_step_many: null handle%s");
    if (T <= 0 || !actions || !reward || !term || !trunc)
        return fail(MGRL_ERR_INVALID, "mgrl_step_many: T>0, actions, reward, term and trunc are required%s");
    DeviceGuard guard(e->device);
    EnvParams p = make_params(e);
    p.T = T;
    p.actions = actions; p.image = image; p.dir = dir; p.mission = mission; p.reward = reward;
    p.term = term; p.trunc = trunc; p.ep_len = ep_len;
    return launch_env(e, MODE_STEP, true, p, (cudaStream_t)stream);
}

int mgrl_observe(mgrl_env* e, uint8_t* image, uint8_t* dir, uint8_t* mission, void* stream) {
    if (!e) return fail(MGRL_ERR_INVALID, "mgrl_observe: null handle%s");
    DeviceGuard guard(e->device);
    EnvParams p = make_params(e);
    p.image = image; p.dir = dir; p.mission = mission;
    return launch_env(e, MODE_OBSERVE, false, p, (cudaStream_t)stream);
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
    CUDA_TRY(cudaMemcpyAsync(e->states, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return MGRL_OK;
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
    CUDA_TRY(cudaMemcpyAsync(e->states, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
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
    gae_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rewards, values, starts, last_values, last_dones, g,
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
    return MGRL_OK;
}

int mgrl_vec_step_host(mgrl_env* e, const uint8_t* actions_host, uint8_t* image_host, uint8_t* direction_host,
                       int64_t* mission_host, float* reward_host, uint8_t* term_host, uint8_t* trunc_host,
                       uint8_t* ep_len_host, uint8_t* term_image_host, uint8_t* term_dir_host, void* stream) {
    if (!e || !actions_host || !reward_host || !term_host || !trunc_host)
        return fail(MGRL_ERR_INVALID, "mgrl_vec_step_host: null argument%s");
    if (!e->host_ready) return fail(MGRL_ERR_INVALID, "mgrl_vec_step_host: call mgrl_vec_reset_host first%s");
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
    return MGRL_OK;
}

}  // extern "C"
