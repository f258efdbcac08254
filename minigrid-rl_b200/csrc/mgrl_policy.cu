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
// Mapping: 32 observations per CTA, 128 threads: a lane owns one observation (its activations live in a private
// 293-word shared-memory block, odd pitch: conflict free) and the CTA's FOUR warps share the 32 observations,
// splitting every layer by output (pooled cells for conv1, output positions for conv2, 16 of the 64 channels for
// conv3 and the hidden layers).  Weights are staged layer by layer (16 KB chunks) through shared memory by the whole
// CTA, so every weight read in the inner loops is a warp-uniform 16-byte shared-memory broadcast instead of an L2
// round trip; 4 CTAs per SM overlap one CTA's staging with the others' arithmetic.
// fp32 FMA throughout (single-pass TF32 would miss the 1e-5 parity bar).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "mgrl.h"
#include "mgrl_policy_layout.cuh"
#include "mgrl_conv1_tc5.cuh"

using namespace mgrl_policy;

// one error buffer for the whole library (mgrl_kernels.cu): mgrl_last_error() and mgrl_policy_last_error() read the same text
char* mgrl_error_buffer();

namespace {

constexpr int OB = 32;             // observations per CTA
constexpr int PITCH = 293;         // words of shared memory per observation
constexpr int FRAME_WORDS = 37;    // 148-byte frame record
constexpr int POOL_OFF = 148;      // pooled conv1 output: 9 cells x 16 channels
constexpr int H_OFF = 208;         // hidden layer scratch (64 floats) behind the 208 features

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

constexpr int PARTS = 4;           // warps sharing the CTA's 32 observations
constexpr int NT = OB * PARTS;     // threads per CTA
constexpr int WBUF = 4096;         // floats of weights staged in shared memory at a time (16 KB)

// acc[0..NOUT) += x * w[0..NOUT) with w in shared memory (the same address in every lane: one broadcast per 16 bytes)
template <int NOUT>
__device__ __forceinline__ void fma_row_s(float* acc, float x, const float* w) {
    const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll
    for (int i = 0; i < NOUT / 4; ++i) {
        const float4 v = w4[i];
        acc[4 * i + 0] = fmaf(x, v.x, acc[4 * i + 0]);
        acc[4 * i + 1] = fmaf(x, v.y, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(x, v.z, acc[4 * i + 2]);
        acc[4 * i + 3] = fmaf(x, v.w, acc[4 * i + 3]);
    }
}

// the CTA copies `nfloats` (a multiple of 4, 16-byte aligned source) of the packed weights into the staging buffer
__device__ __forceinline__ void stage_weights(float* wbuf, const float* __restrict__ src, int nfloats, int tid) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(wbuf);
    for (int e = tid; e < nfloats / 4; e += NT) d4[e] = __ldg(s4 + e);
}

// two 64-wide layers at once (policy and value net), K inputs each, weights streamed through the staging buffer in
// chunks of 32 rows per net; this thread owns outputs [c0, c0 + 16) of both
template <int K>
__device__ __forceinline__ void dense16x2(float* wbuf, const float* __restrict__ wa, const float* __restrict__ ba,
                                          const float* in_a, const float* __restrict__ wb, const float* __restrict__ bb,
                                          const float* in_b, int c0, float* acc_a, float* acc_b, int tid, bool valid) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc_a[i] = __ldg(ba + c0 + i); acc_b[i] = __ldg(bb + c0 + i); }
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int rows = min(32, K - k0);
        __syncthreads();                                   // the previous chunk (or layer) is done with the buffer
        stage_weights(wbuf, wa + k0 * 64, rows * 64, tid);
        stage_weights(wbuf + 2048, wb + k0 * 64, rows * 64, tid);
        __syncthreads();
        if (valid) {
#pragma unroll 4
            for (int k = 0; k < rows; ++k) {
                fma_row_s<16>(acc_a, in_a[k0 + k], wbuf + k * 64 + c0);
                fma_row_s<16>(acc_b, in_b[k0 + k], wbuf + 2048 + k * 64 + c0);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc_a[i] = tanhf(acc_a[i]); acc_b[i] = tanhf(acc_b[i]); }
}

__global__ void __launch_bounds__(NT, 4) policy_forward_kernel(const PolicyArgs p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint8_t s_age[OB];
    float* wbuf = smem + OB * PITCH;             // 16-byte aligned: OB * PITCH * 4 = 37 504
    const int tid = threadIdx.x;
    const int lane = tid & 31, part = tid >> 5;  // part 0..3: which slice of every layer this warp computes
    const int ob = lane;                         // observation of this lane within the CTA
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
    stage_weights(wbuf, w + W1, 768 + 16, tid);  // conv1 weights + bias
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
                            fma_row_s<16>(acc, (float)b[c] * (1.0f / 255.0f), wbuf + (((f * 3 + c) * 4 + kk) * 16));
                    }
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) pooled[c] = fmaxf(pooled[c], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) blk[POOL_OFF + q * 16 + c] = fmaxf(pooled[c] + wbuf[768 + c], 0.0f);
        }
    }
    __syncthreads();
    stage_weights(wbuf, w + W2, 2048 + 32, tid);   // conv2 weights + bias
    __syncthreads();
    // ---- Conv2d(16,32,2) + ReLU  (3x3 -> 2x2): output position o = part, at words [o*32 + c2] over the consumed frames
    if (valid) {
        const int o = part;
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = wbuf[2048 + c];
        for (int kk = 0; kk < 4; ++kk) {
            const int q = ((o >> 1) + (kk >> 1)) * 3 + (o & 1) + (kk & 1);
#pragma unroll 4
            for (int c1 = 0; c1 < 16; ++c1) fma_row_s<32>(acc, blk[POOL_OFF + q * 16 + c1], wbuf + (kk * 16 + c1) * 32);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) blk[o * 32 + c] = fmaxf(acc[c], 0.0f);
    }
    // ---- Conv2d(32,64,2) + ReLU + Flatten  (2x2 -> 1x1): channels [16 part, 16 part + 16), weights in two chunks of 64 rows
    const int c0 = part * 16;
    float fa[16], fb[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) fa[c] = __ldg(w + B3 + c0 + c);
    for (int j0 = 0; j0 < 128; j0 += 64) {
        __syncthreads();
        stage_weights(wbuf, w + W3 + j0 * 64, 64 * 64, tid);
        __syncthreads();
        if (valid) {
#pragma unroll 4
            for (int j = 0; j < 64; ++j) fma_row_s<16>(fa, blk[j0 + j], wbuf + j * 64 + c0);
        }
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
    // ---- first hidden layer of the policy and the value net (208 -> 64 each, Tanh)  (the chunk loop starts with a barrier)
    dense16x2<208>(wbuf, w + PI1, w + PI1B, blk, w + VF1, w + VF1B, blk, c0, fa, fb, tid, valid);
    __syncthreads();          // every part has read the features
    if (valid) {
#pragma unroll
        for (int c = 0; c < 16; ++c) { blk[H_OFF + c0 + c] = fa[c]; blk[c0 + c] = fb[c]; }
    }
    // ---- second hidden layer (64 -> 64, Tanh): policy reads [208,272), value reads [0,64)
    dense16x2<64>(wbuf, w + PI2, w + PI2B, blk + H_OFF, w + VF2, w + VF2B, blk, c0, fa, fb, tid, valid);
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

constexpr size_t kErrBytes = 512;

// MGRL_CONV1_SIMT=1 selects the CUDA-core kernels of the update's first stage (kept for A/B measurements)
bool conv1_tensor_cores() {
    static const bool simt = [] { const char* v = getenv("MGRL_CONV1_SIMT"); return v && atoi(v) != 0; }();
    return !simt;
}

}  // namespace

extern "C" {

const char* mgrl_policy_last_error(void) { return mgrl_error_buffer(); }

int mgrl_policy_forward(const float* weights_dev, const uint8_t* frames_dev, const uint8_t* dirs_dev,
                        const uint8_t* mission_dev, const uint8_t* prev_age_dev, const uint8_t* prev_done_dev,
                        uint8_t* age_out_dev, uint8_t* start_out_dev, uint8_t* action_dev, float* logp_dev,
                        float* value_dev, float* logits_dev, int num_envs, int time_index, uint64_t seed,
                        uint64_t env_id_base, uint32_t step, int flags, void* stream) {
    if (!weights_dev || !frames_dev || !dirs_dev || !mission_dev || !age_out_dev || !value_dev || num_envs <= 0 ||
        time_index < 3) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_policy_forward: null argument, empty batch or time_index < 3");
        return MGRL_ERR_INVALID;
    }
    PolicyArgs a;
    a.w = weights_dev; a.frames = frames_dev; a.dirs = dirs_dev; a.mission = mission_dev;
    a.prev_age = prev_age_dev; a.prev_done = prev_done_dev; a.age_out = age_out_dev; a.start_out = start_out_dev;
    a.action = action_dev; a.logp = logp_dev; a.value = value_dev; a.logits = logits_dev;
    a.n = num_envs; a.b = time_index; a.seed = seed; a.env_id_base = env_id_base; a.step = step;
    a.deterministic = (flags & MGRL_POLICY_DETERMINISTIC) ? 1 : 0;
    cudaError_t e;
    if (flags & MGRL_POLICY_TENSOR) {       // tensor-core kernel: the buffer carries the fragment section
        e = launch_policy_forward_tc(a, (cudaStream_t)stream);
    } else {
        const size_t smem = (size_t)(OB * PITCH + WBUF) * sizeof(float);
        e = cudaFuncSetAttribute(policy_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            policy_forward_kernel<<<(num_envs + OB - 1) / OB, NT, smem, (cudaStream_t)stream>>>(a);
            e = cudaGetLastError();
        }
    }
    if (e != cudaSuccess) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_policy_forward: %s", cudaGetErrorString(e));
        return MGRL_ERR_CUDA;
    }
    return MGRL_OK;
}

int mgrl_policy_pack_fragments(float* weights_dev, void* stream) {
    if (!weights_dev) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_policy_pack_fragments: null argument");
        return MGRL_ERR_INVALID;
    }
    const cudaError_t e = launch_pack_fragments(weights_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_policy_pack_fragments: %s", cudaGetErrorString(e));
        return MGRL_ERR_CUDA;
    }
    return MGRL_OK;
}

}  // extern "C"

// ================================================================================================================
// PPO update, first stage of the image extractor by hand: Conv2d(12,16,2) + ReLU + MaxPool2d(2) forward, and the
// gradient of its weight / bias (the input is an observation: it needs no gradient).  In the library formulation
// this stage alone is ~45 % of a minibatch update (1.8 GB of float patches written and re-read); here the frames are
// read as bytes straight from the rollout buffer (the (t, env) samples of the minibatch are gathered on the fly,
// frames older than the episode zero-filled), and only the pooled 3x3x16 activations leave the kernel.
// Replaces, inside PPO.train ([UPSTREAM] SB3, driven from /root/reference/src/ppo.py:159), the first three modules of
// CustomExtractor's image branch (policies.py:59, hydra_configs/single.yaml:44-47) and their autograd.
namespace {

constexpr int C1_OB = 64;        // samples per CTA (forward)
constexpr int C1_PITCH = 149;    // words per staged sample: 4 frames x 37 words, odd pitch

// stage the 4-frame stacks of samples [s0, s0 + count) into shared memory, PITCH words per sample
__device__ __forceinline__ void stage_samples(const Conv1Args& p, int s0, int count, uint32_t* blocks, int pitch, int tid,
                                              int nthreads) {
    for (int e = tid; e < count * 4 * FRAME_WORDS; e += nthreads) {
        const int o = e / (4 * FRAME_WORDS), r = e - o * (4 * FRAME_WORDS);
        const int f = r / FRAME_WORDS, j = r - f * FRAME_WORDS;
        const int s = s0 + o;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.frames + ((size_t)(p.t[s] + f) * p.n + p.i[s]) * 148);
        blocks[o * pitch + f * FRAME_WORDS + j] = (3 - f) <= (int)p.age[s] ? __ldg(src + j) : 0u;
    }
}

__global__ void __launch_bounds__(C1_OB * 4) conv1_pool_fwd_kernel(const Conv1Args p) {
    extern __shared__ __align__(16) uint32_t c1_smem[];
    float* w1t = reinterpret_cast<float*>(c1_smem + C1_OB * C1_PITCH);   // [48][16]
    float* div255 = w1t + 768;                                           // byte -> float(byte) / 255 (preprocess_obs, exact)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int part = warp >> 1, ob = (warp & 1) * 32 + lane;
    const int s0 = blockIdx.x * C1_OB;
    const int count = min(C1_OB, p.B - s0);
    div255[tid] = (float)tid / 255.0f;
    for (int e = tid; e < 768; e += C1_OB * 4) {   // [co][j] -> [j][co]
        const int co = e / 48, j = e - co * 48;
        w1t[j * 16 + co] = p.w1[e];
    }
    stage_samples(p, s0, count, c1_smem, C1_PITCH, tid, C1_OB * 4);
    __syncthreads();
    if (ob >= count) return;
    const uint8_t* px = reinterpret_cast<const uint8_t*>(c1_smem + ob * C1_PITCH);
    const size_t s = (size_t)(s0 + ob);
    for (int q = part; q < 9; q += 4) {
        const int qh = q / 3, qw = q - qh * 3;
        float best[16];
        int pos[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { best[c] = -3.0e38f; pos[c] = 0; }
        for (int sp = 0; sp < 4; ++sp) {
            const int ph = 2 * qh + (sp >> 1), pw = 2 * qw + (sp & 1);
            float acc[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = 0.0f;
            for (int f = 0; f < 4; ++f) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint8_t* b = px + f * 148 + ((ph + (kk >> 1)) * 7 + pw + (kk & 1)) * 3;
#pragma unroll
                    for (int c3 = 0; c3 < 3; ++c3) {
                        const float x = div255[b[c3]];
                        const float4* w4 = reinterpret_cast<const float4*>(w1t + ((f * 3 + c3) * 4 + kk) * 16);
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const float4 w = w4[v];
                            acc[4 * v + 0] = fmaf(x, w.x, acc[4 * v + 0]); acc[4 * v + 1] = fmaf(x, w.y, acc[4 * v + 1]);
                            acc[4 * v + 2] = fmaf(x, w.z, acc[4 * v + 2]); acc[4 * v + 3] = fmaf(x, w.w, acc[4 * v + 3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (acc[c] > best[c]) { best[c] = acc[c]; pos[c] = sp; }   // first maximum wins, like max_pool2d
        }
        float* out = p.pooled + (s * 9 + q) * 16;
        uint8_t* ao = p.arg + (s * 9 + q) * 16;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float v = best[c] + __ldg(p.b1 + c);
            out[c] = fmaxf(v, 0.0f);
            ao[c] = (uint8_t)(pos[c] | (v > 0.0f ? 4 : 0));
        }
    }
}

// weight / bias gradient: thread = (input index j of 48, channel quad cq of 4), accumulating over every sample of the
// CTA's chunks in registers, one atomic per weight per CTA at the end
constexpr int C1B_CHUNK = 32;
__global__ void __launch_bounds__(192) conv1_pool_bwd_kernel(const Conv1Args p) {
    extern __shared__ __align__(16) uint32_t c1_smem[];
    uint32_t* blocks = c1_smem;                                                        // [32][149] frames
    float* dz = reinterpret_cast<float*>(c1_smem + C1B_CHUNK * C1_PITCH);              // [32][9][16] masked gradient
    uint32_t* arg = reinterpret_cast<uint32_t*>(dz + C1B_CHUNK * 144);                 // [32][9][4] words of 4 positions
    float* div255 = reinterpret_cast<float*>(arg + C1B_CHUNK * 36);                    // byte -> float(byte) / 255
    const int tid = threadIdx.x;
    for (int e = tid; e < 256; e += 192) div255[e] = (float)e / 255.0f;
    const int j = tid % 48, cq = tid / 48;
    const int f = j / 12, c3 = (j / 4) % 3, kk = j & 3;                                // j = (f*3 + c3)*4 + kk
    const int joff = f * 148 + ((kk >> 1) * 7 + (kk & 1)) * 3 + c3;                    // byte offset of input j at position (0,0)
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float bacc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int nchunks = (p.B + C1B_CHUNK - 1) / C1B_CHUNK;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int s0 = ch * C1B_CHUNK;
        const int count = min(C1B_CHUNK, p.B - s0);
        __syncthreads();
        stage_samples(p, s0, count, blocks, C1_PITCH, tid, 192);
        for (int e = tid; e < count * 144; e += 192) {
            const uint8_t a = p.arg[(size_t)s0 * 144 + e];
            dz[e] = (a & 4) ? p.dpooled[(size_t)s0 * 144 + e] : 0.0f;
            reinterpret_cast<uint8_t*>(arg)[e] = a & 3;
        }
        __syncthreads();
        for (int o = 0; o < count; ++o) {
            const uint8_t* px = reinterpret_cast<const uint8_t*>(blocks + o * C1_PITCH) + joff;
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int qh = q / 3, qw = q - qh * 3;
                const uint8_t* b = px + ((2 * qh) * 7 + 2 * qw) * 3;
                // the four candidate positions of this pooled cell: (0,0) (0,1) (1,0) (1,1)
                const uint32_t x4 = (uint32_t)b[0] | ((uint32_t)b[3] << 8) | ((uint32_t)b[21] << 16) | ((uint32_t)b[24] << 24);
                const uint32_t pos4 = arg[(o * 9 + q) * 4 + cq];
                const float4 g = *reinterpret_cast<const float4*>(dz + (o * 9 + q) * 16 + cq * 4);
                const float gs[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t x = (x4 >> (8 * ((pos4 >> (8 * c)) & 3u))) & 0xFFu;
                    acc[c] = fmaf(div255[x], gs[c], acc[c]);
                    bacc[c] += gs[c];
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        atomicAdd(p.dw1 + (cq * 4 + c) * 48 + j, acc[c]);
        if (j == 0) atomicAdd(p.db1 + cq * 4 + c, bacc[c]);
    }
}

}  // namespace

// PPO update, input patches of the second convolution (Conv2d(16,32,2) on the pooled 3x3x16 map, policies.py:59 over
// single.yaml:48): patches[b][o][kk * 16 + ci] = pooled[b][q(o, kk)][ci], o = oh * 2 + ow, kk = kh * 2 + kw,
// q = (oh + kh) * 3 + ow + kw, and its adjoint.  The library formulation (two unfolds + reshape) spends 0.74 ms of a
// 262 144-sample minibatch in unfold_backward alone; these are two float4 copy kernels.
namespace {

// PPO update, gradient of the mission look-up table: out[row[b]][:] += d[b][:] over the minibatch (the adjoint of the
// gather lut[mission * 4 + age] that stands in for the GRU).  A few dozen distinct rows receive 10^5..10^6
// contributions: the library's embedding backward sorts the indices first (0.43 ms per 262 144 samples); here every CTA
// sums its share of the samples into a shared-memory copy of the whole table (296 x 128 floats = 148 KB, conflict-free
// column-interleaved rows, shared-memory float atomics) and adds the touched entries to global memory once.
__global__ void __launch_bounds__(512) lut_grad_kernel(const float4* __restrict__ d, const long long* __restrict__ row, int batch,
                                                       int n_rows, float* __restrict__ out) {
    extern __shared__ float tab[];                 // [n_rows][4][32]: logical column lane * 4 + j at j * 32 + lane
    const int tid = threadIdx.x, lane = tid & 31;
    for (int e = tid; e < n_rows * 128; e += blockDim.x) tab[e] = 0.f;
    __syncthreads();
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int b = (blockIdx.x * blockDim.x + tid) >> 5; b < batch; b += warps) {
        const int r = (int)row[b];
        const float4 v = __ldg(d + (size_t)b * 32 + lane);
        float* t = tab + r * 128 + lane;
        atomicAdd(t, v.x); atomicAdd(t + 32, v.y); atomicAdd(t + 64, v.z); atomicAdd(t + 96, v.w);
    }
    __syncthreads();
    for (int e = tid; e < n_rows * 128; e += blockDim.x) {
        const float v = tab[e];
        if (v != 0.f) {
            const int r = e >> 7, p = e & 127;
            atomicAdd(out + r * 128 + (p & 31) * 4 + (p >> 5), v);
        }
    }
}

// PPO update, bias gradients: out[c] = sum over rows of g[r][c] for a tall row-major matrix (rows = minibatch samples or
// sample x position, cols <= 128).  One pass at streaming speed: a thread owns column (tid % cols) of every
// (blockDim / cols)-th row of its CTA's slab, partial sums meet in shared memory, one atomic per column per CTA.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ g, long long rows, int cols, float* __restrict__ out) {
    __shared__ float part[256];
    const int tid = threadIdx.x;
    const int groups = 256 / cols, grp = tid / cols, c = tid - grp * cols;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    if (grp < groups) {
        const long long stride = (long long)gridDim.x * groups;
        long long r = (long long)blockIdx.x * groups + grp;
        for (; r + 3 * stride < rows; r += 4 * stride) {     // four independent loads in flight
            acc0 += __ldg(g + r * cols + c);
            acc1 += __ldg(g + (r + stride) * cols + c);
            acc2 += __ldg(g + (r + 2 * stride) * cols + c);
            acc3 += __ldg(g + (r + 3 * stride) * cols + c);
        }
        for (; r < rows; r += stride) acc0 += __ldg(g + r * cols + c);
    }
    part[tid] = (acc0 + acc1) + (acc2 + acc3);
    __syncthreads();
    if (tid < cols) {
        float s = 0.f;
        for (int k = 0; k < groups; ++k) s += part[k * cols + tid];
        atomicAdd(out + tid, s);
    }
}

__global__ void patch2x2_fwd_kernel(const float4* __restrict__ pooled, float4* __restrict__ patches, int batch) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // (b, o, kk, ci / 4)
    if (e >= (size_t)batch * 64) return;
    const int c4 = (int)(e & 3), kk = (int)((e >> 2) & 3), o = (int)((e >> 4) & 3);
    const size_t b = e >> 6;
    const int q = ((o >> 1) + (kk >> 1)) * 3 + (o & 1) + (kk & 1);
    patches[e] = __ldg(pooled + (b * 9 + q) * 4 + c4);
}

__global__ void patch2x2_bwd_kernel(const float4* __restrict__ dpatches, float4* __restrict__ dpooled, int batch) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // (b, q, ci / 4)
    if (e >= (size_t)batch * 36) return;
    const size_t b = e / 36;
    const int r = (int)(e - b * 36), q = r >> 2, c4 = r & 3;
    const int qh = q / 3, qw = q - qh * 3;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int oh = qh - (kk >> 1), ow = qw - (kk & 1);
        if (oh < 0 || oh > 1 || ow < 0 || ow > 1) continue;
        const float4 v = __ldg(dpatches + (b * 4 + oh * 2 + ow) * 16 + kk * 4 + c4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    dpooled[e] = acc;
}

}  // namespace

extern "C" {

int mgrl_conv1_pool_forward(const uint8_t* frames_dev, int num_envs, const int32_t* t_dev, const int32_t* i_dev,
                            const uint8_t* age_dev, int batch, const float* w1_dev, const float* b1_dev, float* pooled_dev,
                            uint8_t* arg_dev, void* stream) {
    if (!frames_dev || !t_dev || !i_dev || !age_dev || !w1_dev || !b1_dev || !pooled_dev || !arg_dev || batch <= 0 || num_envs <= 0) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_conv1_pool_forward: null argument or empty batch");
        return MGRL_ERR_INVALID;
    }
    Conv1Args a = {};
    a.frames = frames_dev; a.t = t_dev; a.i = i_dev; a.age = age_dev; a.w1 = w1_dev; a.b1 = b1_dev; a.pooled = pooled_dev;
    a.arg = arg_dev; a.n = num_envs; a.B = batch;
    cudaError_t e;
    // MGRL_CONV1_TC5=1 / 2: the tcgen05 kernel (mgrl_conv1_tc5.cu) with the two-term split / one TF32 pass
    const char* tc5_env = getenv("MGRL_CONV1_TC5");   // (read per call: the parity test switches it)
    const int tc5 = tc5_env ? atoi(tc5_env) : 0;
    if (tc5) {
        a.onepass = tc5 == 2;
        e = mgrl_tc5::launch_conv1_pool_fwd(a, (cudaStream_t)stream);
    } else if (conv1_tensor_cores()) {
        e = launch_conv1_pool_fwd_tc(a, (cudaStream_t)stream);
    } else {
        const size_t smem = (size_t)(C1_OB * C1_PITCH + 768 + 256) * 4;
        e = cudaFuncSetAttribute(conv1_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            conv1_pool_fwd_kernel<<<(batch + C1_OB - 1) / C1_OB, C1_OB * 4, smem, (cudaStream_t)stream>>>(a);
            e = cudaGetLastError();
        }
    }
    if (e != cudaSuccess) { snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_conv1_pool_forward: %s", cudaGetErrorString(e)); return MGRL_ERR_CUDA; }
    return MGRL_OK;
}

int mgrl_conv1_pool_backward(const uint8_t* frames_dev, int num_envs, const int32_t* t_dev, const int32_t* i_dev,
                             const uint8_t* age_dev, int batch, const uint8_t* arg_dev, const float* dpooled_dev, float* dw1_dev,
                             float* db1_dev, void* stream) {
    if (!frames_dev || !t_dev || !i_dev || !age_dev || !arg_dev || !dpooled_dev || !dw1_dev || !db1_dev || batch <= 0 || num_envs <= 0) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_conv1_pool_backward: null argument or empty batch");
        return MGRL_ERR_INVALID;
    }
    Conv1Args a = {};
    a.frames = frames_dev; a.t = t_dev; a.i = i_dev; a.age = age_dev; a.arg = const_cast<uint8_t*>(arg_dev); a.dpooled = dpooled_dev;
    a.dw1 = dw1_dev; a.db1 = db1_dev; a.n = num_envs; a.B = batch;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(dw1_dev, 0, 768 * sizeof(float), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(db1_dev, 0, 16 * sizeof(float), s);
    if (e == cudaSuccess && conv1_tensor_cores()) {
        e = launch_conv1_pool_bwd_tc(a, s);
    } else if (e == cudaSuccess) {
        const size_t smem = (size_t)(C1B_CHUNK * C1_PITCH) * 4 + (size_t)C1B_CHUNK * 144 * 4 + (size_t)C1B_CHUNK * 144 + 256 * 4;
        e = cudaFuncSetAttribute(conv1_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            const int nchunks = (batch + C1B_CHUNK - 1) / C1B_CHUNK;
            const int grid = nchunks < 148 * 4 ? nchunks : 148 * 4;
            conv1_pool_bwd_kernel<<<grid, 192, smem, s>>>(a);
            e = cudaGetLastError();
        }
    }
    if (e != cudaSuccess) { snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_conv1_pool_backward: %s", cudaGetErrorString(e)); return MGRL_ERR_CUDA; }
    return MGRL_OK;
}

int mgrl_patch2x2_forward(const float* pooled_dev, int batch, float* patches_dev, void* stream) {
    if (!pooled_dev || !patches_dev || batch <= 0) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_patch2x2_forward: null argument or empty batch");
        return MGRL_ERR_INVALID;
    }
    const size_t total = (size_t)batch * 64;
    patch2x2_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(pooled_dev), reinterpret_cast<float4*>(patches_dev), batch);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_patch2x2_forward: %s", cudaGetErrorString(e)); return MGRL_ERR_CUDA; }
    return MGRL_OK;
}

int mgrl_patch2x2_backward(const float* dpatches_dev, int batch, float* dpooled_dev, void* stream) {
    if (!dpatches_dev || !dpooled_dev || batch <= 0) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_patch2x2_backward: null argument or empty batch");
        return MGRL_ERR_INVALID;
    }
    const size_t total = (size_t)batch * 36;
    patch2x2_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(dpatches_dev), reinterpret_cast<float4*>(dpooled_dev), batch);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_patch2x2_backward: %s", cudaGetErrorString(e)); return MGRL_ERR_CUDA; }
    return MGRL_OK;
}

int mgrl_lut_grad(const float* d_dev, const int64_t* rows_dev, int batch, int n_rows, float* out_dev, void* stream) {
    if (!d_dev || !rows_dev || !out_dev || batch <= 0 || n_rows <= 0 || n_rows > 400) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_lut_grad: null argument, empty batch or more than 400 table rows");
        return MGRL_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)n_rows * 128 * sizeof(float);
    static thread_local size_t opted = 0;          // the opt-in is not a stream operation: once, outside any capture
    cudaError_t e = cudaSuccess;
    if (smem > opted) {
        e = cudaFuncSetAttribute(lut_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) opted = smem;
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(out_dev, 0, smem, s);
    if (e == cudaSuccess) {
        const int grid = batch < 148 * 16 ? (batch + 15) / 16 : 148;
        lut_grad_kernel<<<grid, 512, smem, s>>>(reinterpret_cast<const float4*>(d_dev), reinterpret_cast<const long long*>(rows_dev),
                                               batch, n_rows, out_dev);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_lut_grad: %s", cudaGetErrorString(e)); return MGRL_ERR_CUDA; }
    return MGRL_OK;
}

int mgrl_colsum(const float* g_dev, long long rows, int cols, float* out_dev, void* stream) {
    if (!g_dev || !out_dev || rows <= 0 || cols <= 0 || cols > 128) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_colsum: null argument, empty matrix or more than 128 columns");
        return MGRL_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out_dev, 0, (size_t)cols * sizeof(float), s);
    if (e == cudaSuccess) {
        const long long per = 256 / cols * 64;            // about 64 rows per thread
        long long grid = (rows + per - 1) / per;
        grid = grid < 1 ? 1 : (grid > 148 * 8 ? 148 * 8 : grid);
        colsum_kernel<<<(unsigned)grid, 256, 0, s>>>(g_dev, rows, cols, out_dev);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_colsum: %s", cudaGetErrorString(e)); return MGRL_ERR_CUDA; }
    return MGRL_OK;
}

}  // extern "C"
