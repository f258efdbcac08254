// mgrl_policy.cu — K3: fused observation -> policy forward -> categorical sample, fp32 on the CUDA cores.
//
// Replaces, for the rollout, CustomPPOPolicy.forward (/root/reference/src/policies.py:227-244 over
// CustomExtractor, policies.py:21-120, built from hydra_configs/single.yaml:38-57) and what it takes from
// Stable-Baselines3: VecFrameStack(4,'first') + VecTransposeImage (ppo.py:124-126) are a gather from the
// un-stacked [time, env, 148] frame buffer with zero fill across episode boundaries (no stacked copy in HBM),
// preprocess_obs (image / 255), MlpExtractor pi/vf 208->64->64 Tanh, action_net, value_net, Categorical.
// The GRU over the stacked mission tokens is a look-up: the mission is constant within an episode, so its
// 128-float feature depends only on (mission id, frames in the stack) -> table [74*4][128] recomputed by the
// host whenever the weights change.
//
// Mapping: 64 observations per CTA, 256 threads: a lane owns one observation (its activations live in a private
// 293-word shared-memory block, odd pitch: conflict free) and FOUR warps share each group of 32 observations,
// splitting every layer by output (pooled cells for conv1, output positions for conv2, 16 of the 64 channels for
// conv3 and the hidden layers).  All weight reads are therefore warp-uniform 16-byte read-only loads (one
// transaction broadcast to the warp); with 24 warps per SM their L2 latency is covered by the other warps.
// fp32 FMA throughout (single-pass TF32 would miss the 1e-5 parity bar).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "mgrl.h"

namespace {

constexpr int OB = 64;             // observations per CTA
constexpr int PITCH = 293;         // words of shared memory per observation
constexpr int FRAME_WORDS = 37;    // 148-byte frame record
constexpr int POOL_OFF = 148;      // pooled conv1 output: 9 cells x 16 channels
constexpr int H_OFF = 208;         // hidden layer scratch (64 floats) behind the 208 features

// float offsets into the packed weight buffer (mirrored in minigrid-rl_b200/policy.py: WEIGHT_LAYOUT)
constexpr int W1 = 0, B1 = W1 + 48 * 16, W2 = B1 + 16, B2 = W2 + 64 * 32, W3 = B2 + 32, B3 = W3 + 128 * 64;
constexpr int WD = B3 + 64, BD = WD + 16 * 16;
constexpr int PI1 = BD + 16, PI1B = PI1 + 208 * 64, PI2 = PI1B + 64, PI2B = PI2 + 64 * 64;
constexpr int VF1 = PI2B + 64, VF1B = VF1 + 208 * 64, VF2 = VF1B + 64, VF2B = VF2 + 64 * 64;
constexpr int WA = VF2B + 64, BA = WA + 64 * 8, WV = BA + 8, BV = WV + 64, LUT = BV + 4;
constexpr int N_WEIGHTS = LUT + MGRL_N_MISSIONS * 4 * 128;
static_assert(N_WEIGHTS == MGRL_POLICY_WEIGHTS, "weight layout");

struct PolicyArgs {
    const float* w;
    const uint8_t* frames;   // record of (time b, env i) at frames + (b * n + i) * 148; the kernel reads b-3..b
    const uint8_t* dirs;     // (b * n + i)
    const uint8_t* mission;  // [n] mission id at time b
    const uint8_t* prev_age; // [n] or null (= first observation after a reset)
    const uint8_t* prev_done;// [n] done flag of the step that produced this observation, or null
    uint8_t* age_out;        // [n] frames of history available for this observation, 0..3
    uint8_t* start_out;      // [n] or null: episode_start flag
    uint8_t* action;         // [n] or null
    float* logp;             // [n] or null
    float* value;            // [n]
    float* logits;           // [n,7] or null
    int n, b;
    uint64_t seed, env_id_base;
    uint32_t step;           // sampling counter (global step index)
    int deterministic;       // 1: action = argmax(logits) (evaluate_policy / test(), ppo.py:161,174-292)
};

__device__ __forceinline__ void philox_u01(uint64_t seed, uint64_t env, uint32_t step, float& u) {
    uint32_t c0 = step, c1 = 0x504F4C49u /* "POLI" */, c2 = (uint32_t)env, c3 = (uint32_t)(env >> 32);
    uint32_t ka = (uint32_t)seed, kb = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ ka; c1 = l1; c2 = h0 ^ c3 ^ kb; c3 = l0;
        ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    u = (float)(c0 >> 8) * (1.0f / 16777216.0f);
}

// acc[0..NOUT) += x * w[0..NOUT), w 16-byte aligned and the same address in every lane
template <int NOUT>
__device__ __forceinline__ void fma_row(float* acc, float x, const float* __restrict__ w) {
    const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll
    for (int i = 0; i < NOUT / 4; ++i) {
        const float4 v = __ldg(w4 + i);
        acc[4 * i + 0] = fmaf(x, v.x, acc[4 * i + 0]);
        acc[4 * i + 1] = fmaf(x, v.y, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(x, v.z, acc[4 * i + 2]);
        acc[4 * i + 3] = fmaf(x, v.w, acc[4 * i + 3]);
    }
}

constexpr int PARTS = 4;           // warps sharing one group of 32 observations
constexpr int NT = OB * PARTS;     // threads per CTA

// 16 of the 64 outputs of two hidden layers at once (policy and value net read the same input):
// acc_a += in_a[k] * wa[k][c0..c0+16), acc_b += in_b[k] * wb[k][c0..c0+16)
template <int K>
__device__ __forceinline__ void dense16x2(const float* __restrict__ wa, const float* __restrict__ ba, const float* in_a,
                                          const float* __restrict__ wb, const float* __restrict__ bb, const float* in_b,
                                          int c0, float* acc_a, float* acc_b) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc_a[i] = __ldg(ba + c0 + i); acc_b[i] = __ldg(bb + c0 + i); }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        fma_row<16>(acc_a, in_a[k], wa + k * 64 + c0);
        fma_row<16>(acc_b, in_b[k], wb + k * 64 + c0);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc_a[i] = tanhf(acc_a[i]); acc_b[i] = tanhf(acc_b[i]); }
}

__global__ void __launch_bounds__(NT, 3) policy_forward_kernel(const PolicyArgs p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint8_t s_age[OB];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int part = warp >> 1;                 // 0..3: which slice of every layer this warp computes
    const int ob = (warp & 1) * 32 + lane;      // observation of this lane within the CTA
    const int i0 = blockIdx.x * OB;
    const int i = i0 + ob;
    const int nv = min(OB, p.n - i0);
    const bool valid = ob < nv;
    const float* __restrict__ w = p.w;

    // ---- frames of history available (VecFrameStack zero-fills what precedes the episode)
    if (tid < OB) {
        int a = 0;
        if (tid < nv) {
            const int gi = i0 + tid;
            const bool start = p.prev_done == nullptr || p.prev_done[gi] != 0;
            a = start ? 0 : min((int)(p.prev_age ? p.prev_age[gi] : 0) + 1, 3);
            p.age_out[gi] = (uint8_t)a;
            if (p.start_out) p.start_out[gi] = start;
        }
        s_age[tid] = (uint8_t)a;
    }
    __syncthreads();
    const int age = s_age[ob];

    // ---- gather the 4-frame stack: coalesced reads of [nv x 148 B] per frame, scattered to per-observation blocks
    uint32_t* blocks = reinterpret_cast<uint32_t*>(smem);
    for (int f = 0; f < 4; ++f) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.frames + ((size_t)(p.b - 3 + f) * p.n + i0) * 148);
        for (int e = tid; e < nv * FRAME_WORDS; e += NT) {
            const int o = e / FRAME_WORDS, j = e - o * FRAME_WORDS;
            blocks[o * PITCH + f * FRAME_WORDS + j] = (3 - f) <= (int)s_age[o] ? src[e] : 0u;
        }
    }
    __syncthreads();

    float* blk = smem + ob * PITCH;
    const uint8_t* px = reinterpret_cast<const uint8_t*>(blk);

    // ---- image: Conv2d(12,16,2) + ReLU + MaxPool2d(2)  (7x7 -> 6x6 -> 3x3), input / 255; pooled cells q = part, part+4, 8
    if (valid) {
        for (int q = part; q < 9; q += PARTS) {
            const int qh = q / 3, qw = q - qh * 3;
            float pooled[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) pooled[c] = -3.0e38f;
            for (int s = 0; s < 4; ++s) {
                const int ph = 2 * qh + (s >> 1), pw = 2 * qw + (s & 1);
                float acc[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[c] = 0.0f;
                for (int f = 0; f < 4; ++f) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int cell = (ph + (kk >> 1)) * 7 + pw + (kk & 1);
                        const uint8_t* b = px + f * 148 + cell * 3;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            fma_row<16>(acc, (float)b[c] * (1.0f / 255.0f), w + W1 + (((f * 3 + c) * 4 + kk) * 16));
                    }
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) pooled[c] = fmaxf(pooled[c], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) blk[POOL_OFF + q * 16 + c] = fmaxf(pooled[c] + __ldg(w + B1 + c), 0.0f);
        }
    }
    __syncthreads();
    // ---- Conv2d(16,32,2) + ReLU  (3x3 -> 2x2): output position o = part, at words [o*32 + c2] over the consumed frames
    if (valid) {
        const int o = part;
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = __ldg(w + B2 + c);
        for (int kk = 0; kk < 4; ++kk) {
            const int q = ((o >> 1) + (kk >> 1)) * 3 + (o & 1) + (kk & 1);
#pragma unroll 4
            for (int c1 = 0; c1 < 16; ++c1) fma_row<32>(acc, blk[POOL_OFF + q * 16 + c1], w + W2 + (kk * 16 + c1) * 32);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) blk[o * 32 + c] = fmaxf(acc[c], 0.0f);
    }
    __syncthreads();
    // ---- Conv2d(32,64,2) + ReLU + Flatten  (2x2 -> 1x1): channels [16 part, 16 part + 16)
    const int c0 = part * 16;
    float fa[16], fb[16];
    if (valid) {
#pragma unroll
        for (int c = 0; c < 16; ++c) fa[c] = __ldg(w + B3 + c0 + c);
#pragma unroll 4
        for (int j = 0; j < 128; ++j) fma_row<16>(fa, blk[j], w + W3 + j * 64 + c0);
    }
    __syncthreads();          // conv2's output is consumed: the block becomes the 208 features
    if (valid) {
        // features = [direction 0:16 | image 16:80 | mission 80:208]  (policies.py:83-102)
#pragma unroll
        for (int c = 0; c < 16; ++c) blk[16 + c0 + c] = fmaxf(fa[c], 0.0f);
        if (part == 0) {      // direction: Linear(16,16) on the stacked one-hot (frames older than the episode are zero)
            float fd[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) fd[c] = __ldg(w + BD + c);
            for (int f = 0; f < 4; ++f) {
                if ((3 - f) <= age) {
                    const int d = p.dirs[(size_t)(p.b - 3 + f) * p.n + i] & 3;
                    fma_row<16>(fd, 1.0f, w + WD + (f * 4 + d) * 16);
                }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) blk[c] = fd[c];
        }
        const float4* row = reinterpret_cast<const float4*>(w + LUT + ((int)p.mission[i] * 4 + age) * 128) + part * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = __ldg(row + c);
            float* d = blk + 80 + part * 32 + 4 * c;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    }
    __syncthreads();
    // ---- first hidden layer of the policy and the value net (208 -> 64 each, Tanh)
    if (valid) dense16x2<208>(w + PI1, w + PI1B, blk, w + VF1, w + VF1B, blk, c0, fa, fb);
    __syncthreads();          // every part has read the features
    if (valid) {
#pragma unroll
        for (int c = 0; c < 16; ++c) { blk[H_OFF + c0 + c] = fa[c]; blk[c0 + c] = fb[c]; }
    }
    __syncthreads();
    // ---- second hidden layer (64 -> 64, Tanh): policy reads [208,272), value reads [0,64)
    if (valid) dense16x2<64>(w + PI2, w + PI2B, blk + H_OFF, w + VF2, w + VF2B, blk, c0, fa, fb);
    __syncthreads();
    if (valid) {
#pragma unroll
        for (int c = 0; c < 16; ++c) { blk[64 + c0 + c] = fa[c]; blk[128 + c0 + c] = fb[c]; }
    }
    __syncthreads();
    if (!valid) return;
    if (part == 1) {          // value_net
        float v = __ldg(w + BV);
#pragma unroll 8
        for (int k = 0; k < 64; ++k) v = fmaf(blk[128 + k], __ldg(w + WV + k), v);
        p.value[i] = v;
    }
    if (part != 0) return;
    // ---- action_net + Categorical(logits): log-softmax, inverse-CDF sample on one Philox uniform per (env, step)
    float lg[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) lg[a] = __ldg(w + BA + a);
#pragma unroll 4
    for (int k = 0; k < 64; ++k) fma_row<8>(lg, blk[64 + k], w + WA + k * 8);
    if (p.logits) {
#pragma unroll
        for (int a = 0; a < 7; ++a) p.logits[(size_t)i * 7 + a] = lg[a];
    }
    if (p.action) {
        float m = lg[0];
#pragma unroll
        for (int a = 1; a < 7; ++a) m = fmaxf(m, lg[a]);
        float sum = 0.0f;
#pragma unroll
        for (int a = 0; a < 7; ++a) sum += expf(lg[a] - m);
        const float lse = m + logf(sum);
        float u;
        philox_u01(p.seed, p.env_id_base + (uint64_t)i, p.step, u);
        int act = 6;
        float c = 0.0f, chosen = lg[6];
        bool found = false;
#pragma unroll
        for (int a = 0; a < 7; ++a) {
            c += expf(lg[a] - lse);
            const bool take = p.deterministic ? (lg[a] == m) : (u < c);   // argmax: first maximum, like torch.argmax
            if (!found && take) { act = a; chosen = lg[a]; found = true; }
        }
        p.action[i] = (uint8_t)act;
        if (p.logp) p.logp[i] = chosen - lse;
    }
}

thread_local char g_perr[256] = "";

}  // namespace

extern "C" {

const char* mgrl_policy_last_error(void) { return g_perr; }

int mgrl_policy_forward(const float* weights_dev, const uint8_t* frames_dev, const uint8_t* dirs_dev,
                        const uint8_t* mission_dev, const uint8_t* prev_age_dev, const uint8_t* prev_done_dev,
                        uint8_t* age_out_dev, uint8_t* start_out_dev, uint8_t* action_dev, float* logp_dev,
                        float* value_dev, float* logits_dev, int num_envs, int time_index, uint64_t seed,
                        uint64_t env_id_base, uint32_t step, int flags, void* stream) {
    if (!weights_dev || !frames_dev || !dirs_dev || !mission_dev || !age_out_dev || !value_dev || num_envs <= 0 ||
        time_index < 3) {
        snprintf(g_perr, sizeof g_perr, "mgrl_policy_forward: null argument, empty batch or time_index < 3");
        return MGRL_ERR_INVALID;
    }
    PolicyArgs a;
    a.w = weights_dev; a.frames = frames_dev; a.dirs = dirs_dev; a.mission = mission_dev;
    a.prev_age = prev_age_dev; a.prev_done = prev_done_dev; a.age_out = age_out_dev; a.start_out = start_out_dev;
    a.action = action_dev; a.logp = logp_dev; a.value = value_dev; a.logits = logits_dev;
    a.n = num_envs; a.b = time_index; a.seed = seed; a.env_id_base = env_id_base; a.step = step;
    a.deterministic = (flags & MGRL_POLICY_DETERMINISTIC) ? 1 : 0;
    const size_t smem = (size_t)OB * PITCH * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(policy_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        policy_forward_kernel<<<(num_envs + OB - 1) / OB, NT, smem, (cudaStream_t)stream>>>(a);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        snprintf(g_perr, sizeof g_perr, "mgrl_policy_forward: %s", cudaGetErrorString(e));
        return MGRL_ERR_CUDA;
    }
    return MGRL_OK;
}

}  // extern "C"
