// mgrl_policy_tc.cu — K3 on the tensor cores: the same fused observation -> policy forward -> categorical sample as
// mgrl_policy.cu, with every convolution / linear layer issued as mma.sync m16n8k8 TF32 with fp32 accumulation.
//
// Replaces, for the rollout, CustomPPOPolicy.forward (/root/reference/src/policies.py:227-244 over CustomExtractor,
// policies.py:21-120, hydra_configs/single.yaml:38-57) + VecFrameStack / VecTransposeImage (ppo.py:124-126).
//
// Precision.  One TF32 pass (10-bit mantissa) misses the 1e-5 parity bar of BASELINE.json, so every product is the
// three-term split  x*w = x_hi*w_hi + x_hi*w_lo + x_lo*w_hi  (x_hi = tf32(x), x_lo = tf32(x - x_hi); the dropped
// x_lo*w_lo term is 2^-22 relative), accumulated in fp32: fp32-class results from the tensor pipe.  The first
// convolution needs two terms only: its inputs are observation bytes (exact in TF32) and 1/255 is folded into the
// weights.
//
// Mapping.  A warp owns 16 observations = the 16 rows of an m16 tile and walks the whole network for them; nothing
// but the 4-frame byte stack (shared memory, gathered from the un-stacked frame buffer) is staged.  Activations
// never leave registers: the accumulator fragment of a layer (row g: columns 2t, 2t+1) IS the A fragment of the next
// layer (row g: k-slots t, t+4) once that layer's weights are packed with k-slot s <-> input channel 2s / 2(s-4)+1 of
// the 8-channel group, so there is no shuffle or shared-memory round trip between layers.  The max-pool is a running
// maximum over the four accumulator sets of a pooled cell, and a pooled cell is folded into the second convolution's
// accumulators as soon as it exists (8 live registers instead of 72).
//
// Weights.  `pack_fragments_kernel` rewrites the packed fp32 weights into per-lane B fragments
// {b0_hi, b1_hi, b0_lo, b1_lo} in (k-tile, n-tile, lane) order behind the fp32 section of the weight buffer: one
// coalesced 16-byte load feeds the three mma of a tile.  The section is 375 KB and shared by every warp: L1/L2 resident.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "mgrl.h"
#include "mgrl_policy_layout.cuh"

using namespace mgrl_policy;

namespace {

constexpr int WARPS = 4;                 // warps per CTA
constexpr int OBC = 16 * WARPS;          // observations per CTA
constexpr int NT = 32 * WARPS;
constexpr int FRAME_WORDS = 37;          // 148-byte frame record
constexpr int SLAB = OBC * 148;          // bytes of one frame of the CTA's observations (contiguous in the frame buffer)

// fragment section, float4 units: entry = base + (kt * NT_layer + nt) * 32 + lane
constexpr int F_C1 = 0;                          // conv1   K = 48  (6 k-tiles)  N = 16 (2 n-tiles), natural k order, / 255
constexpr int F_C2 = F_C1 + 6 * 2 * 32;          // conv2   K = 64  (8)          N = 32 (4)
constexpr int F_C3 = F_C2 + 8 * 4 * 32;          // conv3   K = 128 (16)         N = 64 (8)
constexpr int F_M1 = F_C3 + 16 * 8 * 32;         // pi | vf first layer  K = 208 (26)  N = 128 (16)
constexpr int F_P2 = F_M1 + 26 * 16 * 32;        // pi second layer      K = 64 (8)    N = 64 (8)
constexpr int F_V2 = F_P2 + 8 * 8 * 32;          // vf second layer
constexpr int F_HA = F_V2 + 8 * 8 * 32;          // action_net           K = 64 (8)    N = 8 (1)
constexpr int F_HV = F_HA + 8 * 32;              // value_net            K = 64 (8)    N = 8 (1), column 0
constexpr int F_END = F_HV + 8 * 32;
static_assert(F_END * 4 == MGRL_POLICY_FRAGMENTS, "fragment section size");

// round to TF32 (nearest, ties away from zero: cvt.rna.tf32.f32) with two integer instructions; the conversion
// instruction itself runs on the quarter-rate XU pipe, which the first version of this kernel saturated (84 % busy)
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
// float(byte) without a conversion instruction: 2^23 + b is exact
__device__ __forceinline__ uint32_t byte_to_float_bits(uint32_t b) { return __float_as_uint(__uint_as_float(0x4B000000u | b) - 8388608.0f); }

// c += a * b  (m16n8k8, A row-major 16x8, B col-major 8x8, fp32 accumulate)
__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// an activation fragment split for the three-term product
struct AFrag {
    uint32_t hi[4], lo[4];
};
// values of rows g / g+8 at the two k-slots of this lane: (r0 slot t, r1 slot t, r0 slot t+4, r1 slot t+4)
__device__ __forceinline__ void split(AFrag& f, float r0a, float r1a, float r0b, float r1b) {
    const float v[4] = {r0a, r1a, r0b, r1b};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f.hi[i] = to_tf32(v[i]);
        f.lo[i] = to_tf32(v[i] - __uint_as_float(f.hi[i]));
    }
}
// the accumulator fragment of the previous layer (c0 = row g col 2t, c1 = row g col 2t+1, c2 / c3 = row g+8) as this
// layer's A fragment under the channel permutation of the packed weights
__device__ __forceinline__ void split_acc(AFrag& f, const float (&c)[4]) { split(f, c[0], c[2], c[1], c[3]); }

__device__ __forceinline__ void mma3(float (&c)[4], const AFrag& a, const float4 b) {
    mma8(c, a.lo, __float_as_uint(b.x), __float_as_uint(b.y));
    mma8(c, a.hi, __float_as_uint(b.z), __float_as_uint(b.w));
    mma8(c, a.hi, __float_as_uint(b.x), __float_as_uint(b.y));
}

// conv1's fragments are re-read from shared memory at every use (a volatile load: the compiler would otherwise keep all
// twelve, 48 registers, live across the 36 positions)
__device__ __forceinline__ float4 lds_frag(const float4* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

// the same product with the two small cross terms kept in an accumulator of their own (added to the big one in fp32 on
// the CUDA cores at the end of the layer): the tensor core's fp32 accumulation does not round to nearest, and terms
// 2^-11 of the running sum lose most of their bits when they are added to it directly
__device__ __forceinline__ void mma3s(float (&c)[4], float (&cs)[4], const AFrag& a, const float4 b) {
    mma8(cs, a.lo, __float_as_uint(b.x), __float_as_uint(b.y));
    mma8(cs, a.hi, __float_as_uint(b.z), __float_as_uint(b.w));
    mma8(c, a.hi, __float_as_uint(b.x), __float_as_uint(b.y));
}

// d = a * b (zero accumulator input)
__device__ __forceinline__ void mma8z(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
// the three-term product of ONE k-tile on its own, then added to the running sum with a rounded fp32 add on the CUDA
// cores: the tensor core truncates when it adds to its accumulator (a bias that grows with the number of k-tiles)
__device__ __forceinline__ void mma3r(float (&c)[4], const AFrag& a, const float4 b) {
    float d[4];
    mma8z(d, a.lo, __float_as_uint(b.x), __float_as_uint(b.y));
    mma8(d, a.hi, __float_as_uint(b.z), __float_as_uint(b.w));
    mma8(d, a.hi, __float_as_uint(b.x), __float_as_uint(b.y));
#pragma unroll
    for (int r = 0; r < 4; ++r) c[r] += d[r];
}

// accumulators of one n-tile start at the layer's bias (columns 2t, 2t+1 of the tile)
__device__ __forceinline__ void init_bias(float (&c)[4], const float* __restrict__ bias, int t) {
    const float2 b = __ldg(reinterpret_cast<const float2*>(bias) + t);
    c[0] = b.x; c[1] = b.y; c[2] = b.x; c[3] = b.y;
}

// W warps per CTA (16 observations each).  At N = 65 536 there are 4096 warp tiles: with 4 warps per CTA and 2 CTAs per
// SM (255 registers) the 1184 resident warps need 4 rounds for 3.46 rounds of work; 5 warps per CTA (168 registers, 1480
// resident warps) need 3 for 2.77 but each round is slower: no gain measured, 4 stays the default.
template <int W, int PREC>
__global__ void __launch_bounds__(32 * W, 2) policy_forward_tc_kernel(const PolicyArgs p) {
    constexpr int OBC = 16 * W, NT = 32 * W, SLAB = OBC * 148;
    extern __shared__ __align__(16) uint8_t tc_smem[];
    uint8_t* s_frames = tc_smem;                                            // [frame][observation][148]
    float4* s_c1 = reinterpret_cast<float4*>(tc_smem + 4 * SLAB);           // conv1 fragments (reused by all 36 positions)
    uint8_t* s_age = tc_smem + 4 * SLAB + 6 * 2 * 32 * 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int i0 = blockIdx.x * OBC;
    const int nv = min(OBC, p.n - i0);
    const float* __restrict__ w = p.w;
    const float4* __restrict__ frag = reinterpret_cast<const float4*>(p.w + N_WEIGHTS);

    // ---- frames of history available (VecFrameStack zero-fills what precedes the episode)
    if (tid < OBC) {
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
    // ---- the 4-frame stacks: frame f of the CTA's observations is one contiguous slab of the un-stacked buffer ->
    //      16-byte cp.async chunks, all in flight at once (bytes past the batch are zero-filled by the copy)
    {
        const int valid = nv * 148;
        const bool aligned = (p.n & 3) == 0;      // slab starts are 16-byte aligned when n * 148 is a multiple of 16
        for (int f = 0; f < 4; ++f) {
            const uint8_t* src = p.frames + ((size_t)(p.b - 3 + f) * p.n + i0) * 148;
            if (aligned) {
                for (int c = tid; c < SLAB / 16; c += NT) {
                    const int left = valid - c * 16;
                    const int sz = left >= 16 ? 16 : (left > 0 ? left : 0);
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_frames + f * SLAB + c * 16);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src + (sz ? c * 16 : 0)), "r"(sz) : "memory");
                }
            } else {
                const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
                uint32_t* d32 = reinterpret_cast<uint32_t*>(s_frames + f * SLAB);
                for (int e = tid; e < OBC * FRAME_WORDS; e += NT) d32[e] = e < nv * FRAME_WORDS ? __ldg(s32 + e) : 0u;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int e = tid; e < 6 * 2 * 32; e += NT) s_c1[e] = __ldg(frag + F_C1 + e);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // frames older than the episode are zero (VecFrameStack after a reset)
    for (int e = tid; e < OBC * 3; e += NT) {
        const int o = e / 3, f = e - o * 3;
        if ((3 - f) > (int)s_age[o]) {
            uint32_t* d32 = reinterpret_cast<uint32_t*>(s_frames + f * SLAB + o * 148);
#pragma unroll
            for (int j = 0; j < FRAME_WORDS; ++j) d32[j] = 0u;
        }
    }
    __syncthreads();

    const int r0 = warp * 16 + g, r1 = r0 + 8;              // this lane's two observations (rows g and g+8 of the tile)
    const bool v0 = r0 < nv, v1 = r1 < nv;
    const uint8_t* px0 = s_frames + r0 * 148;
    const uint8_t* px1 = s_frames + r1 * 148;

    // ---- image branch.  conv2's accumulators h2[o][nt] live through conv1: every pooled cell is folded in at once.
    float h2[4][4][4];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) init_bias(h2[o][nt], w + B2 + nt * 8, t);
    const float2 b1a = __ldg(reinterpret_cast<const float2*>(w + B1) + t), b1b = __ldg(reinterpret_cast<const float2*>(w + B1 + 8) + t);
    const int tap = ((t >> 1) * 7 + (t & 1)) * 3;           // this lane's kernel tap (k = ci * 4 + tap, tap = kh * 2 + kw = t)
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int qh = q / 3, qw = q - qh * 3;
        // Conv2d(12,16,2) at the four positions of pooled cell q, running maximum (MaxPool2d(2)), then bias + ReLU
        float best[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) best[nt][r] = -3.0e38f;
#pragma unroll 1
        for (int sp = 0; sp < 4; ++sp) {      // rolled: the fully unrolled network does not fit the instruction cache
            const int cell = ((2 * qh + (sp >> 1)) * 7 + 2 * qw + (sp & 1)) * 3 + tap;
            float cur[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int kt = 0; kt < 6; ++kt) {
                // k-tile kt = input channels ci = 2kt (slots 0..3 = taps) and 2kt + 1 (slots 4..7); ci = frame * 3 + c
                const int oa = ((2 * kt) / 3) * SLAB + (2 * kt) % 3, ob = ((2 * kt + 1) / 3) * SLAB + (2 * kt + 1) % 3;
                uint32_t a[4];
                a[0] = byte_to_float_bits(px0[oa + cell]); a[1] = byte_to_float_bits(px1[oa + cell]);
                a[2] = byte_to_float_bits(px0[ob + cell]); a[3] = byte_to_float_bits(px1[ob + cell]);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const float4 b = lds_frag(s_c1 + (kt * 2 + nt) * 32 + lane);
                    mma8(cur[nt], a, __float_as_uint(b.z), __float_as_uint(b.w));   // bytes are exact in TF32: two terms
                    mma8(cur[nt], a, __float_as_uint(b.x), __float_as_uint(b.y));
                }
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r) best[nt][r] = fmaxf(best[nt][r], cur[nt][r]);
        }
        AFrag pa[2];
        {
            const float x0 = fmaxf(best[0][0] + b1a.x, 0.f), x1 = fmaxf(best[0][1] + b1a.y, 0.f);
            const float x2 = fmaxf(best[0][2] + b1a.x, 0.f), x3 = fmaxf(best[0][3] + b1a.y, 0.f);
            split(pa[0], x0, x2, x1, x3);
            const float y0 = fmaxf(best[1][0] + b1b.x, 0.f), y1 = fmaxf(best[1][1] + b1b.y, 0.f);
            const float y2 = fmaxf(best[1][2] + b1b.x, 0.f), y3 = fmaxf(best[1][3] + b1b.y, 0.f);
            split(pa[1], y0, y2, y1, y3);
        }
        // Conv2d(16,32,2): cell q is tap (qh - oh, qw - ow) of output position o = (oh, ow)
#pragma unroll
        for (int oh = 0; oh < 2; ++oh)
#pragma unroll
            for (int ow = 0; ow < 2; ++ow) {
                const int kh = qh - oh, kw = qw - ow;
                if (kh < 0 || kh > 1 || kw < 0 || kw > 1) continue;
                const int kk = kh * 2 + kw;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
                        mma3(h2[oh * 2 + ow][nt], pa[h], __ldg(frag + F_C2 + ((kk * 2 + h) * 4 + nt) * 32 + lane));
            }
    }
    // ---- Conv2d(32,64,2) + ReLU + Flatten: k = o * 32 + c2
    float x3[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) init_bias(x3[nt], w + B3 + nt * 8, t);
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            AFrag a;
            split(a, fmaxf(h2[o][n2][0], 0.f), fmaxf(h2[o][n2][2], 0.f), fmaxf(h2[o][n2][1], 0.f), fmaxf(h2[o][n2][3], 0.f));
            const int kt = o * 4 + n2;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) mma3(x3[nt], a, __ldg(frag + F_C3 + (kt * 8 + nt) * 32 + lane));
        }

    // ---- features = [direction 0:16 | image 16:80 | mission 80:208] -> first hidden layers of policy (n-tiles 0..7)
    //      and value net (8..15), Tanh
    const int gi0 = i0 + (v0 ? r0 : 0), gi1 = i0 + (v1 ? r1 : 0);   // rows past the batch read observation i0 (never stored)
    const int age0 = s_age[v0 ? r0 : 0], age1 = s_age[v1 ? r1 : 0];
    float m1[16][4];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) init_bias(m1[nt], (nt < 8 ? w + PI1B + nt * 8 : w + VF1B + (nt - 8) * 8), t);
    float m1s[PREC ? 8 : 1][4];     // small cross terms of the value net's half (PREC)
#pragma unroll
    for (int nt = 0; nt < (PREC ? 8 : 1); ++nt) m1s[nt][0] = m1s[nt][1] = m1s[nt][2] = m1s[nt][3] = 0.f;
    auto m1_ktile = [&](const AFrag& a, int kt) {
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            const float4 b = __ldg(frag + F_M1 + (kt * 16 + nt) * 32 + lane);
            if (PREC == 2 && nt >= 8) mma3r(m1[nt], a, b);
            else if (PREC && nt >= 8) mma3s(m1[nt], m1s[PREC ? nt - 8 : 0], a, b);
            else mma3(m1[nt], a, b);
        }
    };
    {   // direction: Linear(16,16) on the stacked one-hot = bias + one weight row per frame of history (exact fp32 sums)
        float d0[4], d1[4];      // columns 2t, 2t+1, 8+2t, 8+2t+1 of rows g / g+8
#pragma unroll
        for (int j = 0; j < 4; ++j) d0[j] = d1[j] = __ldg(w + BD + (j >> 1) * 8 + 2 * t + (j & 1));
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            if ((3 - f) <= age0) {
                const int d = p.dirs[(size_t)(p.b - 3 + f) * p.n + gi0] & 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) d0[j] += __ldg(w + WD + (f * 4 + d) * 16 + (j >> 1) * 8 + 2 * t + (j & 1));
            }
            if ((3 - f) <= age1) {
                const int d = p.dirs[(size_t)(p.b - 3 + f) * p.n + gi1] & 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) d1[j] += __ldg(w + WD + (f * 4 + d) * 16 + (j >> 1) * 8 + 2 * t + (j & 1));
            }
        }
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            AFrag a;
            split(a, d0[2 * kt], d1[2 * kt], d0[2 * kt + 1], d1[2 * kt + 1]);
            m1_ktile(a, kt);
        }
    }
#pragma unroll
    for (int k8 = 0; k8 < 8; ++k8) {   // image features: ReLU of conv3
        AFrag a;
        split(a, fmaxf(x3[k8][0], 0.f), fmaxf(x3[k8][2], 0.f), fmaxf(x3[k8][1], 0.f), fmaxf(x3[k8][3], 0.f));
        m1_ktile(a, 2 + k8);
    }
    {   // mission features: row (mission, age) of the GRU look-up table
        const float2* row0 = reinterpret_cast<const float2*>(w + LUT + ((int)p.mission[gi0] * 4 + age0) * 128) + t;
        const float2* row1 = reinterpret_cast<const float2*>(w + LUT + ((int)p.mission[gi1] * 4 + age1) * 128) + t;
#pragma unroll 4
        for (int k8 = 0; k8 < 16; ++k8) {
            const float2 u0 = __ldg(row0 + k8 * 4), u1 = __ldg(row1 + k8 * 4);
            AFrag a;
            split(a, u0.x, u1.x, u0.y, u1.y);
            m1_ktile(a, 10 + k8);
        }
    }
#pragma unroll
    for (int nt = 0; nt < 16; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) m1[nt][r] = tanhf(PREC && nt >= 8 ? m1[nt][r] + m1s[PREC ? nt - 8 : 0][r] : m1[nt][r]);

    // ---- policy net: second hidden layer (Tanh), action_net
    float lg[4];
    {
        float h[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) init_bias(h[nt], w + PI2B + nt * 8, t);
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
            AFrag a;
            split_acc(a, m1[kt]);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) mma3(h[nt], a, __ldg(frag + F_P2 + (kt * 8 + nt) * 32 + lane));
        }
        init_bias(lg, w + BA, t);
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
            AFrag a;
            split(a, tanhf(h[kt][0]), tanhf(h[kt][2]), tanhf(h[kt][1]), tanhf(h[kt][3]));
            mma3(lg, a, __ldg(frag + F_HA + kt * 32 + lane));
        }
    }
    // ---- value net: second hidden layer (Tanh), value_net (column 0 of a padded n-tile)
    {
        float h[8][4], hs[PREC ? 8 : 1][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) init_bias(h[nt], w + VF2B + nt * 8, t);
#pragma unroll
        for (int nt = 0; nt < (PREC ? 8 : 1); ++nt) hs[nt][0] = hs[nt][1] = hs[nt][2] = hs[nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
            AFrag a;
            split_acc(a, m1[8 + kt]);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float4 b = __ldg(frag + F_V2 + (kt * 8 + nt) * 32 + lane);
                if (PREC == 2) mma3r(h[nt], a, b);
                else if (PREC) mma3s(h[nt], hs[PREC ? nt : 0], a, b);
                else mma3(h[nt], a, b);
            }
        }
        float vv[4] = {0.f, 0.f, 0.f, 0.f}, vs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
            AFrag a;
            if (PREC) {
#pragma unroll
                for (int r = 0; r < 4; ++r) h[kt][r] += hs[PREC ? kt : 0][r];
            }
            split(a, tanhf(h[kt][0]), tanhf(h[kt][2]), tanhf(h[kt][1]), tanhf(h[kt][3]));
            if (PREC == 2) mma3r(vv, a, __ldg(frag + F_HV + kt * 32 + lane));
            else if (PREC) mma3s(vv, vs, a, __ldg(frag + F_HV + kt * 32 + lane));
            else mma3(vv, a, __ldg(frag + F_HV + kt * 32 + lane));
        }
        if (t == 0) {
            const float bv = __ldg(w + BV);
            if (v0) p.value[i0 + r0] = vv[0] + vs[0] + bv;
            if (v1) p.value[i0 + r1] = vv[2] + vs[2] + bv;
        }
    }
    // ---- Categorical(logits): the 8 columns of a row sit in the 4 lanes of its quad; lane t = 0 finishes row g,
    //      lane t = 1 row g+8 (log-softmax, inverse-CDF sample on one Philox uniform per (env, step))
    float row[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int src = (lane & ~3) | j;
        const float a0 = __shfl_sync(0xffffffffu, lg[0], src), a1 = __shfl_sync(0xffffffffu, lg[1], src);
        const float c0 = __shfl_sync(0xffffffffu, lg[2], src), c1 = __shfl_sync(0xffffffffu, lg[3], src);
        row[2 * j] = t == 0 ? a0 : c0;
        row[2 * j + 1] = t == 0 ? a1 : c1;
    }
    if (t > 1 || !(t == 0 ? v0 : v1)) return;
    const int i = i0 + (t == 0 ? r0 : r1);
    if (p.logits) {
#pragma unroll
        for (int a = 0; a < 7; ++a) p.logits[(size_t)i * 7 + a] = row[a];
    }
    if (p.action) {
        float m = row[0];
#pragma unroll
        for (int a = 1; a < 7; ++a) m = fmaxf(m, row[a]);
        float sum = 0.0f;
#pragma unroll
        for (int a = 0; a < 7; ++a) sum += expf(row[a] - m);
        const float lse = m + logf(sum);
        float u;
        philox_u01(p.seed, p.env_id_base + (uint64_t)i, p.step, u);
        int act = 6;
        float c = 0.0f, chosen = row[6];
        bool found = false;
#pragma unroll
        for (int a = 0; a < 7; ++a) {
            c += expf(row[a] - lse);
            const bool take = p.deterministic ? (row[a] == m) : (u < c);   // argmax: first maximum, like torch.argmax
            if (!found && take) { act = a; chosen = row[a]; found = true; }
        }
        p.action[i] = (uint8_t)act;
        if (p.logp) p.logp[i] = chosen - lse;
    }
}

// packed fp32 weights -> B fragments {b0_hi, b1_hi, b0_lo, b1_lo}.  B[k][n] = weight of input k, output n.
__global__ void pack_fragments_kernel(const float* __restrict__ w, float4* __restrict__ frag) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= F_END) return;
    int base, ntiles;
    if (e < F_C2) { base = F_C1; ntiles = 2; }
    else if (e < F_C3) { base = F_C2; ntiles = 4; }
    else if (e < F_M1) { base = F_C3; ntiles = 8; }
    else if (e < F_P2) { base = F_M1; ntiles = 16; }
    else if (e < F_V2) { base = F_P2; ntiles = 8; }
    else if (e < F_HA) { base = F_V2; ntiles = 8; }
    else if (e < F_HV) { base = F_HA; ntiles = 1; }
    else { base = F_HV; ntiles = 1; }
    const int local = e - base, lane = local & 31, tile = local >> 5;
    const int kt = tile / ntiles, nt = tile - kt * ntiles;
    const int g = lane >> 2, t = lane & 3, n = nt * 8 + g;
    const bool natural = base == F_C1;
    const int k0 = kt * 8 + (natural ? t : 2 * t), k1 = kt * 8 + (natural ? t + 4 : 2 * t + 1);
    auto B = [&](int k) -> float {
        switch (base) {
        case F_C1: return w[W1 + k * 16 + n] / 255.0f;
        case F_C2: return w[W2 + k * 32 + n];
        case F_C3: return w[W3 + k * 64 + n];
        case F_M1: return n < 64 ? w[PI1 + k * 64 + n] : w[VF1 + k * 64 + n - 64];
        case F_P2: return w[PI2 + k * 64 + n];
        case F_V2: return w[VF2 + k * 64 + n];
        case F_HA: return w[WA + k * 8 + n];
        default: return n == 0 ? w[WV + k] : 0.0f;
        }
    };
    const float b0 = B(k0), b1 = B(k1);
    const float h0 = __uint_as_float(to_tf32(b0)), h1 = __uint_as_float(to_tf32(b1));
    frag[e] = make_float4(h0, h1, __uint_as_float(to_tf32(b0 - h0)), __uint_as_float(to_tf32(b1 - h1)));
}


// ================================================================================================================
// PPO update, first extractor stage (Conv2d(12,16,2) + ReLU + MaxPool2d(2): policies.py:59 over single.yaml:44-47,
// inside SB3's PPO.train driven from ppo.py:159) on the tensor cores.  Same contracts as the CUDA-core kernels of
// mgrl_policy.cu (mgrl_conv1_pool_forward / backward): the minibatch samples (t, env) are read as bytes straight off
// the rollout's frame buffer, only the pooled 3x3x16 activations and a byte of arg-max per output leave the forward.
//   forward : the rollout kernel's first stage on gathered samples, plus the arg-max bookkeeping; two-term split
//             (bytes are exact in TF32), so the result is fp32-class whatever precision the rest of the update runs in;
//   backward: dW1[co][k] = sum over (sample, cell, position) of [arg-max == position, output > 0] * dpooled * x(k):
//             a GEMM whose reduction dimension is the sample axis.  A k-tile is 8 samples at one position; A = masked
//             gradient (row = channel), B = the bytes of the patch (column = input index), a 7th n-tile of ones yields the
//             bias gradient; a warp keeps its 16 x 56 accumulators for every sample it sees and the CTA adds them to
//             global memory once.
constexpr int SAMPLE_BYTES = 4 * 148;     // staged 4-frame stack of one sample

// gather the 4-frame stacks of samples [s0, s0 + 64) (4-byte cp.async, all in flight; frames older than the episode
// and rows past the batch are zero-filled by the copy)
__device__ __forceinline__ void stage_samples_async(const Conv1Args& p, int s0, uint8_t* smem, int tid) {
    for (int e = tid; e < OBC * 4 * FRAME_WORDS; e += NT) {
        const int o = e / (4 * FRAME_WORDS), r = e - o * (4 * FRAME_WORDS);
        const int f = r / FRAME_WORDS, j = r - f * FRAME_WORDS;
        const int s = s0 + o;
        const bool live = s < p.B && (3 - f) <= (int)p.age[s];
        const uint8_t* src = live ? p.frames + ((size_t)(p.t[s] + f) * p.n + p.i[s]) * 148 + j * 4 : p.frames;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + o * SAMPLE_BYTES + r * 4);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(live ? 4 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// ONEPASS: one TF32 pass (weights rounded to TF32, the low term of the split dropped): what `update_tf32` asks for
template <bool ONEPASS>
__global__ void __launch_bounds__(NT, 4) conv1_pool_fwd_tc_kernel(const Conv1Args p) {
    __shared__ __align__(16) uint8_t s_px[OBC * SAMPLE_BYTES];
    __shared__ __align__(16) float4 s_c1[6 * 2 * 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int s0 = blockIdx.x * OBC;
    stage_samples_async(p, s0, s_px, tid);
    for (int e = tid; e < 6 * 2 * 32; e += NT) {   // B fragments of W1 / 255 (torch layout [co][k]), natural k order
        const int l = e & 31, tile = e >> 5, kt = tile >> 1, nt = tile & 1;
        const int n = nt * 8 + (l >> 2), k0 = kt * 8 + (l & 3), k1 = k0 + 4;
        const float b0 = __ldg(p.w1 + n * 48 + k0) / 255.0f, b1 = __ldg(p.w1 + n * 48 + k1) / 255.0f;
        const float h0 = __uint_as_float(to_tf32(b0)), h1 = __uint_as_float(to_tf32(b1));
        s_c1[e] = make_float4(h0, h1, __uint_as_float(to_tf32(b0 - h0)), __uint_as_float(to_tf32(b1 - h1)));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const bool v0 = s0 + r0 < p.B, v1 = s0 + r1 < p.B;
    const uint8_t* px0 = s_px + r0 * SAMPLE_BYTES;
    const uint8_t* px1 = s_px + r1 * SAMPLE_BYTES;
    const float2 b1a = __ldg(reinterpret_cast<const float2*>(p.b1) + t), b1b = __ldg(reinterpret_cast<const float2*>(p.b1 + 8) + t);
    const int tap = ((t >> 1) * 7 + (t & 1)) * 3;
#pragma unroll 1
    for (int q = 0; q < 9; ++q) {
        const int qh = q / 3, qw = q - qh * 3;
        float best[2][4];
        int pos[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) { best[nt][r] = -3.0e38f; pos[nt][r] = 0; }
#pragma unroll 1
        for (int sp = 0; sp < 4; ++sp) {
            const int cell = ((2 * qh + (sp >> 1)) * 7 + 2 * qw + (sp & 1)) * 3 + tap;
            float cur[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int kt = 0; kt < 6; ++kt) {
                const int oa = ((2 * kt) / 3) * 148 + (2 * kt) % 3, ob = ((2 * kt + 1) / 3) * 148 + (2 * kt + 1) % 3;
                uint32_t a[4];
                a[0] = byte_to_float_bits(px0[oa + cell]); a[1] = byte_to_float_bits(px1[oa + cell]);
                a[2] = byte_to_float_bits(px0[ob + cell]); a[3] = byte_to_float_bits(px1[ob + cell]);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const float4 b = s_c1[(kt * 2 + nt) * 32 + lane];
                    if (!ONEPASS) mma8(cur[nt], a, __float_as_uint(b.z), __float_as_uint(b.w));
                    mma8(cur[nt], a, __float_as_uint(b.x), __float_as_uint(b.y));
                }
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (cur[nt][r] > best[nt][r]) { best[nt][r] = cur[nt][r]; pos[nt][r] = sp; }   // first maximum wins, like max_pool2d
        }
        // accumulator (row g | g+8, columns 2t, 2t+1 of n-tile nt) -> pooled[s][q][nt*8 + 2t ..], arg likewise
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const float2 bias = nt == 0 ? b1a : b1b;
            const float x0 = best[nt][0] + bias.x, x1 = best[nt][1] + bias.y, x2 = best[nt][2] + bias.x, x3 = best[nt][3] + bias.y;
            const int c = nt * 8 + 2 * t;
            if (v0) {
                const size_t o = ((size_t)(s0 + r0) * 9 + q) * 16 + c;
                *reinterpret_cast<float2*>(p.pooled + o) = make_float2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
                *reinterpret_cast<uint16_t*>(p.arg + o) = (uint16_t)((pos[nt][0] | (x0 > 0.f ? 4 : 0)) | ((pos[nt][1] | (x1 > 0.f ? 4 : 0)) << 8));
            }
            if (v1) {
                const size_t o = ((size_t)(s0 + r1) * 9 + q) * 16 + c;
                *reinterpret_cast<float2*>(p.pooled + o) = make_float2(fmaxf(x2, 0.f), fmaxf(x3, 0.f));
                *reinterpret_cast<uint16_t*>(p.arg + o) = (uint16_t)((pos[nt][2] | (x2 > 0.f ? 4 : 0)) | ((pos[nt][3] | (x3 > 0.f ? 4 : 0)) << 8));
            }
        }
    }
}

template <bool ONEPASS>
__global__ void __launch_bounds__(NT, 5) conv1_pool_bwd_tc_kernel(const Conv1Args p) {
    __shared__ __align__(16) uint8_t s_px[OBC * SAMPLE_BYTES];
    __shared__ float s_acc[16 * 56];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int e = tid; e < 16 * 56; e += NT) s_acc[e] = 0.f;
    float acc[7][4];
#pragma unroll
    for (int nt = 0; nt < 7; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    // B: column n = nt*8 + g is input index k = ci*4 + tap with ci = 2nt + (g >> 2), tap = g & 3
    int koff[6];
#pragma unroll
    for (int nt = 0; nt < 6; ++nt) {
        const int ci = 2 * nt + (g >> 2), tp = g & 3;
        koff[nt] = (ci / 3) * 148 + ci % 3 + ((tp >> 1) * 7 + (tp & 1)) * 3;
    }
    const uint32_t one = g == 0 ? __float_as_uint(1.0f) : 0u;     // 7th n-tile: column 48 = ones (bias gradient)
    const int nchunks = (p.B + OBC - 1) / OBC;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int s0 = ch * OBC;
        __syncthreads();                                          // the previous chunk's bytes are consumed
        stage_samples_async(p, s0, s_px, tid);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {                             // the warp's 16 samples = two k-tiles of 8
            const int ja = warp * 16 + h * 8 + t, jb = ja + 4;    // samples of this lane's k-slots t and t+4
            const bool va = s0 + ja < p.B, vb = s0 + jb < p.B;
            const uint8_t* pa = s_px + ja * SAMPLE_BYTES;
            const uint8_t* pb = s_px + jb * SAMPLE_BYTES;
            // A: masked gradient of channels g / g+8 for samples ja / jb.  The eight values of a pooled cell are loaded one cell
            // ahead (their L2 latency sat in front of every cell: long scoreboard 3.1 cycles per issue)
            float nda0, nda1, ndb0, ndb1;
            int nga0, nga1, ngb0, ngb1;
            auto fetch = [&](int q) {
                const size_t oa = ((size_t)(s0 + ja) * 9 + q) * 16, ob = ((size_t)(s0 + jb) * 9 + q) * 16;
                nda0 = va ? __ldg(p.dpooled + oa + g) : 0.f; nda1 = va ? __ldg(p.dpooled + oa + g + 8) : 0.f;
                ndb0 = vb ? __ldg(p.dpooled + ob + g) : 0.f; ndb1 = vb ? __ldg(p.dpooled + ob + g + 8) : 0.f;
                nga0 = va ? p.arg[oa + g] : 0; nga1 = va ? p.arg[oa + g + 8] : 0;
                ngb0 = vb ? p.arg[ob + g] : 0; ngb1 = vb ? p.arg[ob + g + 8] : 0;
            };
            fetch(0);
#pragma unroll 1
            for (int q = 0; q < 9; ++q) {
                const int qh = q / 3, qw = q - qh * 3;
                const float da0 = nda0, da1 = nda1, db0 = ndb0, db1 = ndb1;
                const int ga0 = nga0, ga1 = nga1, gb0 = ngb0, gb1 = ngb1;
                if (q + 1 < 9) fetch(q + 1);
#pragma unroll 1
                for (int sp = 0; sp < 4; ++sp) {
                    AFrag a;
                    split(a, ga0 == (sp | 4) ? da0 : 0.f, ga1 == (sp | 4) ? da1 : 0.f, gb0 == (sp | 4) ? db0 : 0.f,
                          gb1 == (sp | 4) ? db1 : 0.f);
                    const int cell = ((2 * qh + (sp >> 1)) * 7 + 2 * qw + (sp & 1)) * 3;
#pragma unroll
                    for (int nt = 0; nt < 6; ++nt) {
                        const uint32_t b0 = byte_to_float_bits(pa[koff[nt] + cell]), b1 = byte_to_float_bits(pb[koff[nt] + cell]);
                        if (!ONEPASS) mma8(acc[nt], a.lo, b0, b1);
                        mma8(acc[nt], a.hi, b0, b1);
                    }
                    if (!ONEPASS) mma8(acc[6], a.lo, one, one);
                    mma8(acc[6], a.hi, one, one);
                }
            }
        }
    }
    // warp accumulators (row co = g | g+8, columns 2t, 2t+1 of n-tile nt) -> CTA sum -> one atomic per weight per CTA
#pragma unroll
    for (int nt = 0; nt < 7; ++nt) {
        atomicAdd(&s_acc[g * 56 + nt * 8 + 2 * t], acc[nt][0]); atomicAdd(&s_acc[g * 56 + nt * 8 + 2 * t + 1], acc[nt][1]);
        atomicAdd(&s_acc[(g + 8) * 56 + nt * 8 + 2 * t], acc[nt][2]); atomicAdd(&s_acc[(g + 8) * 56 + nt * 8 + 2 * t + 1], acc[nt][3]);
    }
    __syncthreads();
    for (int e = tid; e < 16 * 49; e += NT) {
        const int co = e / 49, k = e - co * 49;
        const float v = s_acc[co * 56 + k];
        if (k < 48) atomicAdd(p.dw1 + co * 48 + k, v * (1.0f / 255.0f));
        else atomicAdd(p.db1 + co, v);
    }
}

}  // namespace

namespace mgrl_policy {

cudaError_t launch_policy_forward_tc(const PolicyArgs& a, cudaStream_t stream) {
    // Value-net precision (MGRL_TC_PREC).  2 (default): every k-tile's product is added to the running sum with a rounded
    // fp32 add - the error of the values against the fp32 oracle equals the CUDA-core kernel's (1.07e-5 of max |v| on the
    // test rollout's first step, where every stack holds one frame; 8.8e-6 after); 1: small cross terms in their own
    // tensor-core accumulator (15 % faster, 1.13e-5 / 9.7e-6).  A single accumulator measures 1.38e-5 / 1.18e-5: above
    // the 1e-5 bar, not built.
    static const int prec = [] { const char* v = getenv("MGRL_TC_PREC"); return v ? atoi(v) : 2; }();
    // 4 warps per CTA by default; MGRL_TC_WARPS=5 measures the same (230 vs 234 us: three rounds of slower warps)
    static const int warps = [] { const char* v = getenv("MGRL_TC_WARPS"); return v ? atoi(v) : 4; }();
    auto launch = [&](auto kernel, int w) -> cudaError_t {
        const int obc = 16 * w;
        const size_t smem = (size_t)4 * obc * 148 + 6 * 2 * 32 * 16 + obc;
        if (smem > 48 * 1024) {    // 5 warps per CTA only (4 warps: 44 KB, no opt-in; keeps the default launch capturable)
            const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kernel<<<(a.n + obc - 1) / obc, 32 * w, smem, stream>>>(a);
        return cudaGetLastError();
    };
    if (warps == 4) return prec == 1 ? launch(policy_forward_tc_kernel<4, 1>, 4) : launch(policy_forward_tc_kernel<4, 2>, 4);
    return prec == 1 ? launch(policy_forward_tc_kernel<5, 1>, 5) : launch(policy_forward_tc_kernel<5, 2>, 5);
}

cudaError_t launch_conv1_pool_fwd_tc(const Conv1Args& a, cudaStream_t stream) {
    if (a.onepass) conv1_pool_fwd_tc_kernel<true><<<(a.B + OBC - 1) / OBC, NT, 0, stream>>>(a);
    else conv1_pool_fwd_tc_kernel<false><<<(a.B + OBC - 1) / OBC, NT, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_conv1_pool_bwd_tc(const Conv1Args& a, cudaStream_t stream) {   // dw1 / db1 zeroed by the caller
    const int nchunks = (a.B + OBC - 1) / OBC;
    const int grid = nchunks < 148 * 5 ? nchunks : 148 * 5;
    if (a.onepass) conv1_pool_bwd_tc_kernel<true><<<grid, NT, 0, stream>>>(a);
    else conv1_pool_bwd_tc_kernel<false><<<grid, NT, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_pack_fragments(float* weights_dev, cudaStream_t stream) {
    pack_fragments_kernel<<<(F_END + 255) / 256, 256, 0, stream>>>(weights_dev, reinterpret_cast<float4*>(weights_dev + N_WEIGHTS));
    return cudaGetLastError();
}

}  // namespace mgrl_policy
