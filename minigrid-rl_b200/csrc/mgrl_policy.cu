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
// Mapping: one thread = one observation, 64 observations per CTA.  All weight reads are warp-uniform 16-byte
// read-only loads (one L1 transaction broadcast to the warp); each thread keeps up to 64 accumulators in
// registers and its activations in a private 293-word shared-memory block (odd pitch: conflict free).
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

// 64-wide hidden layer: out = tanh(W^T in + b), `in` and `out` in this thread's shared block
template <int K>
__device__ __forceinline__ void dense64_tanh(const float* __restrict__ wt, const float* __restrict__ bias, const float* in,
                                             float* out) {
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = __ldg(bias + i);
#pragma unroll 2
    for (int k = 0; k < K; ++k) fma_row<64>(acc, in[k], wt + k * 64);
#pragma unroll
    for (int i = 0; i < 64; ++i) out[i] = tanhf(acc[i]);
}

__global__ void __launch_bounds__(OB) policy_forward_kernel(const PolicyArgs p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint8_t s_age[OB];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * OB;
    const int i = i0 + tid;
    const int nv = min(OB, p.n - i0);
    const float* __restrict__ w = p.w;

    // ---- frames of history available (VecFrameStack zero-fills what precedes the episode)
    int age = 0;
    if (i < p.n) {
        const bool start = p.prev_done == nullptr || p.prev_done[i] != 0;
        age = start ? 0 : min((int)(p.prev_age ? p.prev_age[i] : 0) + 1, 3);
        p.age_out[i] = (uint8_t)age;
        if (p.start_out) p.start_out[i] = start;
    }
    s_age[tid] = (uint8_t)age;
    __syncthreads();

    // ---- gather the 4-frame stack: coalesced reads of [nv x 148 B] per frame, scattered to per-thread blocks
    uint32_t* blocks = reinterpret_cast<uint32_t*>(smem);
    for (int f = 0; f < 4; ++f) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.frames + ((size_t)(p.b - 3 + f) * p.n + i0) * 148);
        for (int e = tid; e < nv * FRAME_WORDS; e += OB) {
            const int o = e / FRAME_WORDS, j = e - o * FRAME_WORDS;
            blocks[o * PITCH + f * FRAME_WORDS + j] = (3 - f) <= (int)s_age[o] ? src[e] : 0u;
        }
    }
    __syncthreads();
    if (i >= p.n) return;

    float* blk = smem + tid * PITCH;
    const uint8_t* px = reinterpret_cast<const uint8_t*>(blk);

    // ---- image: Conv2d(12,16,2) + ReLU + MaxPool2d(2)  (7x7 -> 6x6 -> 3x3), input / 255
    for (int q = 0; q < 9; ++q) {
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
    // ---- Conv2d(16,32,2) + ReLU  (3x3 -> 2x2); output (o, c2) at words [o*32 + c2] over the consumed frames
    for (int o = 0; o < 4; ++o) {
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
    // ---- Conv2d(32,64,2) + ReLU + Flatten  (2x2 -> 1x1)
    float feat_img[64];
    {
#pragma unroll
        for (int c = 0; c < 64; ++c) feat_img[c] = __ldg(w + B3 + c);
#pragma unroll 2
        for (int j = 0; j < 128; ++j) fma_row<64>(feat_img, blk[j], w + W3 + j * 64);
    }
    // ---- direction: Linear(16,16) on the stacked one-hot (frames older than the episode are all-zero)
    float feat_dir[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) feat_dir[c] = __ldg(w + BD + c);
    for (int f = 0; f < 4; ++f) {
        if ((3 - f) <= age) {
            const int d = p.dirs[(size_t)(p.b - 3 + f) * p.n + i] & 3;
            fma_row<16>(feat_dir, 1.0f, w + WD + (f * 4 + d) * 16);
        }
    }
    // features = [direction 0:16 | image 16:80 | mission 80:208]  (policies.py:83-102)
#pragma unroll
    for (int c = 0; c < 16; ++c) blk[c] = feat_dir[c];
#pragma unroll
    for (int c = 0; c < 64; ++c) blk[16 + c] = fmaxf(feat_img[c], 0.0f);
    {
        const float4* row = reinterpret_cast<const float4*>(w + LUT + ((int)p.mission[i] * 4 + age) * 128);
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
            const float4 v = __ldg(row + c);
            blk[80 + 4 * c] = v.x; blk[81 + 4 * c] = v.y; blk[82 + 4 * c] = v.z; blk[83 + 4 * c] = v.w;
        }
    }
    // ---- policy head
    dense64_tanh<208>(w + PI1, w + PI1B, blk, blk + H_OFF);
    dense64_tanh<64>(w + PI2, w + PI2B, blk + H_OFF, blk + H_OFF);
    float lg[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) lg[a] = __ldg(w + BA + a);
#pragma unroll 4
    for (int k = 0; k < 64; ++k) fma_row<8>(lg, blk[H_OFF + k], w + WA + k * 8);
    // ---- value head
    dense64_tanh<208>(w + VF1, w + VF1B, blk, blk + H_OFF);
    dense64_tanh<64>(w + VF2, w + VF2B, blk + H_OFF, blk + H_OFF);
    float v = __ldg(w + BV);
#pragma unroll 8
    for (int k = 0; k < 64; ++k) v = fmaf(blk[H_OFF + k], __ldg(w + WV + k), v);
    p.value[i] = v;
    if (p.logits) {
#pragma unroll
        for (int a = 0; a < 7; ++a) p.logits[(size_t)i * 7 + a] = lg[a];
    }
    // ---- Categorical(logits): log-softmax, inverse-CDF sample on one Philox uniform per (env, step)
    if (p.action) {
        float m = lg[0];
#pragma unroll
        for (int a = 1; a < 7; ++a) m = fmaxf(m, lg[a]);
        float e[7], sum = 0.0f;
#pragma unroll
        for (int a = 0; a < 7; ++a) { e[a] = expf(lg[a] - m); sum += e[a]; }
        const float lse = m + logf(sum);
        float u;
        philox_u01(p.seed, p.env_id_base + (uint64_t)i, p.step, u);
        int act = 6;
        float c = 0.0f, chosen = lg[6];
        bool found = false;
#pragma unroll
        for (int a = 0; a < 7; ++a) {
            c += expf(lg[a] - lse);
            if (!found && u < c) { act = a; chosen = lg[a]; found = true; }
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
                        uint64_t env_id_base, uint32_t step, void* stream) {
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
    const size_t smem = (size_t)OB * PITCH * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(policy_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        policy_forward_kernel<<<(num_envs + OB - 1) / OB, OB, smem, (cudaStream_t)stream>>>(a);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        snprintf(g_perr, sizeof g_perr, "mgrl_policy_forward: %s", cudaGetErrorString(e));
        return MGRL_ERR_CUDA;
    }
    return MGRL_OK;
}

}  // extern "C"
