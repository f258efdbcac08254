// mgrl_update.cu — K5: the PPO optimizer step (SB3 `PPO.train` on CustomPPOPolicy, /root/reference/src/ppo.py:94-113,159;
// policies.py:21-120,227-257; hydra_configs/algorithm/ppo.yaml:28-38) as hand-written kernels only: no cuBLAS / cuDNN /
// torch kernel runs between `mgrl_ppo_gradients` and `mgrl_ppo_apply`.
//
//   mission table   gru_gi_kernel, gru_fwd_kernel           GRU(Embedding(tokens)) of the 74 x 4 distinct stacked missions: W_hh lives in
//                                                           REGISTERS (one gate row per thread), two sequences per CTA, 128 dependent steps
//   samples         prep_kernel                             (t, env) -> indices, age, mission row, directions, old values / log-probs ...
//   extractor       conv1_pool_fwd_tc (mgrl_policy_tc.cu), rows_gemm<conv2, patch loader>, rows_gemm<conv3>, assemble_kernel (direction
//                   Linear on a one-hot = column sums, mission table rows)
//   MLPs            rows_gemm<pi|vf first layer>, rows_gemm<pi|vf second layer>
//   heads + loss    loss_kernel: action / value heads, log-softmax, clipped surrogate, clipped value loss, entropy bonus, their
//                   gradients, the head weight gradients and d(loss)/d(second-layer pre-activation)
//   backward        rows_gemm with transposed fragment packs (dX = dZ W, previous activation's derivative in the epilogue),
//                   wgrad_kernel (dW = dZ^T X, db = column sums; reduction over the samples, accumulators in registers for the whole
//                   launch), dir_grad_kernel, lut_grad_strided_kernel, conv1_pool_bwd_tc, gru_bwd_kernel + wgrad for W_hh,
//                   gru_embed_grad_kernel
//   optimizer       gradnorm_kernel + adam_kernel (global-norm clip + Adam on the flat 110 216-float buffer)
//
// Every contraction is mma.sync.m16n8k8 TF32 with fp32 accumulation; `strict` selects the three-term split (fp32-class
// results, the parity mode) instead of one TF32 pass (what the reference runs with, ppo.py:29-32).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <new>

#include "mgrl.h"
#include "mgrl_policy_layout.cuh"
#include "mgrl_conv1_tc5.cuh"
#include "mgrl_linear_tc5.cuh"

char* mgrl_error_buffer();

namespace {

constexpr int kErrBytes = 512;
using mgrl_policy::Conv1Args;

// ------------------------------------------------------------------------------------------------ flat parameter layout
// order and shapes = minigrid-rl_b200/policy.py SHAPES (SB3 state-dict names)
constexpr int P_WD = 0, P_BD = 256, P_WC1 = 272, P_BC1 = 1040, P_WC2 = 1056, P_BC2 = 3104, P_WC3 = 3136, P_BC3 = 11328;
constexpr int P_EMB = 11392, P_WIH = 12416, P_WHH = 24704, P_BIH = 73856, P_BHH = 74240;
constexpr int P_PI1 = 74624, P_PI1B = 87936, P_PI2 = 88000, P_PI2B = 92096;
constexpr int P_VF1 = 92160, P_VF1B = 105472, P_VF2 = 105536, P_VF2B = 109632;
constexpr int P_WA = 109696, P_BA = 110144, P_WV = 110151, P_BV = 110215, P_TOTAL = 110216;
static_assert(P_TOTAL == MGRL_PPO_PARAMS, "flat parameter count");

constexpr int NSEQ_MAX = 400;
constexpr int SEQ_LEN = 128;
constexpr int HID = 128;
constexpr int G3 = 384;

// ------------------------------------------------------------------------------------------------ mma helpers
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
struct AFrag {
    uint32_t hi[4], lo[4];
};
template <bool STRICT>
__device__ __forceinline__ void make_afrag(AFrag& f, const float (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f.hi[i] = to_tf32(v[i]);
        if (STRICT) f.lo[i] = to_tf32(v[i] - __uint_as_float(f.hi[i]));
    }
}
__device__ __forceinline__ void cp_async16(uint32_t dst_sa, const void* src, bool valid) {
    const int n = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_sa), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------ weight packs
// B fragments of mma.m16n8k8 in (group, k-tile, n-tile, lane) order: b0 = B[kt*8 + t][nt*8 + g], b1 = B[kt*8 + t + 4][nt*8 + g]
// (g = lane / 4, t = lane % 4).  One TF32 pass: float2 {b0, b1} rounded to TF32; strict: float4 {b0_hi, b1_hi, b0_lo, b1_lo}.
enum Pack : int { PK_C2F = 0, PK_C3F, PK_L1F, PK_L2F, PK_L2B, PK_L1B, PK_C3B, PK_C2B, PK_COUNT };
struct PackDesc {
    int groups, KT, NT, base;   // base: entry offset of the pack in the fragment buffer
};
__host__ __device__ inline PackDesc pack_desc(int pk) {
    // entries per group = KT * NT * 32
    const int g[PK_COUNT] = {1, 1, 1, 2, 2, 2, 1, 1};
    const int kt[PK_COUNT] = {8, 16, 26, 8, 8, 16, 8, 4};
    const int nt[PK_COUNT] = {4, 8, 16, 8, 8, 13, 16, 8};
    PackDesc d;
    int base = 0;
    for (int i = 0; i < pk; ++i) base += g[i] * kt[i] * nt[i] * 32;
    d.groups = g[pk]; d.KT = kt[pk]; d.NT = nt[pk]; d.base = base;
    return d;
}
constexpr int kPackEntries = 1024 + 4096 + 13312 + 4096 + 4096 + 13312 + 4096 + 1024;   // 45056
// packed biases behind the fragments: conv2 [32], conv3 [64], first MLP layer pi|vf [128], second [128]
constexpr int PB_C2 = 0, PB_C3 = 32, PB_L1 = 96, PB_L2 = 224, PB_TOTAL = 352;

// B(k, n) of pack `pk`, group `grp`
__device__ __forceinline__ float pack_source(const float* __restrict__ P, int pk, int grp, int k, int n) {
    switch (pk) {
        case PK_C2F: { const int kk = k >> 4, ci = k & 15; return P[P_WC2 + n * 64 + ci * 4 + kk]; }            // k = kk*16+ci, n = co
        case PK_C3F: { const int kk = k >> 5, ci = k & 31; return P[P_WC3 + n * 128 + ci * 4 + kk]; }           // k = kk*32+ci, n = co
        case PK_L1F: return n < 64 ? P[P_PI1 + n * 208 + k] : P[P_VF1 + (n - 64) * 208 + k];                   // pi | vf side by side
        case PK_L2F: return P[(grp ? P_VF2 : P_PI2) + n * 64 + k];
        case PK_L2B: return P[(grp ? P_VF2 : P_PI2) + k * 64 + n];                                               // k = out j, n = in i
        case PK_L1B: { const int i = grp * 104 + n; return k < 64 ? P[P_PI1 + k * 208 + i] : P[P_VF1 + (k - 64) * 208 + i]; }
        case PK_C3B: { const int kk = n >> 5, ci = n & 31; return P[P_WC3 + k * 128 + ci * 4 + kk]; }           // k = co, n = kk*32+ci
        default:     { const int kk = n >> 4, ci = n & 15; return P[P_WC2 + k * 64 + ci * 4 + kk]; }                       // PK_C2B, k = co
    }
}

__global__ void pack_update_kernel(const float* __restrict__ P, float* __restrict__ frag, float* __restrict__ pbias, int strict) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < PB_TOTAL) {
        float v;
        if (e < PB_C3) v = P[P_BC2 + e];
        else if (e < PB_L1) v = P[P_BC3 + e - PB_C3];
        else if (e < PB_L2) { const int j = e - PB_L1; v = j < 64 ? P[P_PI1B + j] : P[P_VF1B + j - 64]; }
        else { const int j = e - PB_L2; v = j < 64 ? P[P_PI2B + j] : P[P_VF2B + j - 64]; }
        pbias[e] = v;
    }
    if (e >= kPackEntries) return;
    int pk = 0;
    PackDesc d = pack_desc(0);
    for (int i = 1; i < PK_COUNT; ++i) {
        const PackDesc di = pack_desc(i);
        if (e >= di.base) { pk = i; d = di; }
    }
    int r = e - d.base;
    const int lane = r & 31; r >>= 5;
    const int nt = r % d.NT; r /= d.NT;
    const int kt = r % d.KT; const int grp = r / d.KT;
    const int g = lane >> 2, t = lane & 3;
    const float b0 = pack_source(P, pk, grp, kt * 8 + t, nt * 8 + g);
    const float b1 = pack_source(P, pk, grp, kt * 8 + t + 4, nt * 8 + g);
    const float h0 = __uint_as_float(to_tf32(b0)), h1 = __uint_as_float(to_tf32(b1));
    if (strict) {
        reinterpret_cast<float4*>(frag)[e] = make_float4(h0, h1, __uint_as_float(to_tf32(b0 - h0)), __uint_as_float(to_tf32(b1 - h1)));
    } else {
        reinterpret_cast<float2*>(frag)[e] = make_float2(h0, h1);
    }
}

// ------------------------------------------------------------------------------------------------ rows_gemm
// out[r, :] = epilogue( A[r, :] . B )  for a tall A (rows = samples or sample x position): one CTA = 128 rows, warp w = rows
// 16w..16w+15 against all 8*NT columns of its column group (blockIdx.y).  The A tile goes to shared memory with cp.async
// (row pitch K + 4 words = 4 x odd: fragment loads are conflict free), B fragments come pre-packed from L1/L2.
enum Loader : int { LD_PLAIN = 0, LD_PATCH };
enum Epi : int { EP_BIAS_RELU = 0, EP_BIAS_TANH, EP_GRAD_TANH, EP_GRAD_RELU, EP_GRAD_MIX, EP_PATCH_ADJ };
struct GemmArgs {
    const float* a; int lda; int a_col_step;      // A rows (LD_PATCH: pooled [B,9,16]); column offset per group
    const void* frag; int frag_step;              // packed B fragments; entries per group
    const float* bias; int bias_step;             // packed bias (forward epilogues)
    const float* y; int ldy; int y_col_step;      // activation whose derivative the gradient epilogues apply
    float* out; int ldo; int out_col_step;
    long long rows;
};
constexpr int GM_ROWS = 128;

// the four patch rows of a sample: row = b*4 + o (o = oh*2 + ow), column kk*16 + ci (kk = kh*2 + kw) <- pooled[b][(oh+kh)*3 + ow+kw][ci]
__device__ __forceinline__ const float* patch_src(const float* pooled, long long row, int kk) {
    const int o = (int)(row & 3);
    const int q = ((o >> 1) + (kk >> 1)) * 3 + (o & 1) + (kk & 1);
    return pooled + ((row >> 2) * 9 + q) * 16;
}

template <int K, int LOADER, int C4_LO = 0, int C4_HI = K / 4>
__device__ __forceinline__ void load_rows_tile(float* tile, const float* a, int lda, long long row0, long long rows, int nrows, int tid,
                                               int nthreads) {
    constexpr int LD = K + 4, C4 = C4_HI - C4_LO;
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(tile);
    for (int e = tid; e < nrows * C4; e += nthreads) {
        const int r = e / C4, c4 = C4_LO + e - r * C4;
        const long long row = row0 + r;
        const bool valid = row < rows;
        const float* src;
        if (LOADER == LD_PATCH) src = patch_src(a, valid ? row : 0, c4 >> 2) + (c4 & 3) * 4;
        else src = a + (valid ? row : 0) * (long long)lda + c4 * 4;
        cp_async16(sa + (uint32_t)(r * LD + c4 * 4) * 4u, src, valid);
    }
}

template <int K, int NT, int LOADER, int EPI, bool STRICT>
__global__ void __launch_bounds__(256, 2) rows_gemm_kernel(const GemmArgs p) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int LD = K + 4, KT = K / 8;
    float* tile = smem_f;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int grp = blockIdx.y;
    const long long row0 = (long long)blockIdx.x * GM_ROWS;
    // the tile arrives in two column halves: the k-loop starts on the first while the second is still in flight
    constexpr int KT_A = (KT + 1) / 2;
    load_rows_tile<K, LOADER, 0, KT_A * 2>(tile, p.a + grp * p.a_col_step, p.lda, row0, p.rows, GM_ROWS, tid, 256);
    cp_async_commit();
    load_rows_tile<K, LOADER, KT_A * 2, K / 4>(tile, p.a + grp * p.a_col_step, p.lda, row0, p.rows, GM_ROWS, tid, 256);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    const float* ta = tile + (warp * 16 + g) * LD + t;
    const size_t fbase = (size_t)grp * p.frag_step + lane;
#pragma unroll 2
    for (int kt = 0; kt < KT; ++kt) {
        if (kt == KT_A) { cp_async_wait<0>(); __syncthreads(); }
        const float v[4] = {ta[kt * 8], ta[kt * 8 + 8 * LD], ta[kt * 8 + 4], ta[kt * 8 + 8 * LD + 4]};
        AFrag af;
        make_afrag<STRICT>(af, v);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const size_t e = fbase + (size_t)(kt * NT + nt) * 32;
            if (STRICT) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.frag) + e);
                mma8(acc[nt], af.lo, __float_as_uint(b.x), __float_as_uint(b.y));
                mma8(acc[nt], af.hi, __float_as_uint(b.z), __float_as_uint(b.w));
                mma8(acc[nt], af.hi, __float_as_uint(b.x), __float_as_uint(b.y));
            } else {
                const float2 b = __ldg(reinterpret_cast<const float2*>(p.frag) + e);
                mma8(acc[nt], af.hi, __float_as_uint(b.x), __float_as_uint(b.y));
            }
        }
    }

    const long long ra = row0 + warp * 16 + g, rb = ra + 8;
    if (EPI == EP_PATCH_ADJ) {
        // dpatches [128 rows = 32 samples x 4 positions][64] -> dpooled [32][9][16]: the adjoint of the patch gather, summed in a
        // fixed order (no atomics)
        __syncthreads();                       // everyone is done with the A tile
        constexpr int PL = 68;
        float* dp = smem_f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int c = nt * 8 + 2 * t;
            dp[(warp * 16 + g) * PL + c] = acc[nt][0]; dp[(warp * 16 + g) * PL + c + 1] = acc[nt][1];
            dp[(warp * 16 + g + 8) * PL + c] = acc[nt][2]; dp[(warp * 16 + g + 8) * PL + c + 1] = acc[nt][3];
        }
        __syncthreads();
        const long long b0 = row0 >> 2, nb = (p.rows >> 2);
        for (int e = tid; e < 32 * 144; e += 256) {
            const int bl = e / 144, r = e - bl * 144, q = r >> 4, ci = r & 15, qh = q / 3, qw = q - qh * 3;
            if (b0 + bl >= nb) continue;
            float s = 0.f;
#pragma unroll
            for (int kh = 0; kh < 2; ++kh)
#pragma unroll
                for (int kw = 0; kw < 2; ++kw) {
                    const int oh = qh - kh, ow = qw - kw;
                    if (oh >= 0 && oh < 2 && ow >= 0 && ow < 2) s += dp[(bl * 4 + oh * 2 + ow) * PL + (kh * 2 + kw) * 16 + ci];
                }
            p.out[(b0 + bl) * 144 + r] = s;
        }
        return;
    }
    float* out = p.out + grp * p.out_col_step;
    const float* y = p.y ? p.y + grp * p.y_col_step : nullptr;
    const float* bias = p.bias ? p.bias + grp * p.bias_step : nullptr;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int c = nt * 8 + 2 * t;
        float v0 = acc[nt][0], v1 = acc[nt][1], v2 = acc[nt][2], v3 = acc[nt][3];
        if (EPI == EP_BIAS_RELU || EPI == EP_BIAS_TANH) {
            const float2 b = __ldg(reinterpret_cast<const float2*>(bias + c));
            v0 += b.x; v1 += b.y; v2 += b.x; v3 += b.y;
            if (EPI == EP_BIAS_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
            else { v0 = tanhf(v0); v1 = tanhf(v1); v2 = tanhf(v2); v3 = tanhf(v3); }
        } else {
            float2 ya = make_float2(0.f, 0.f), yb = ya;
            if (ra < p.rows) ya = *reinterpret_cast<const float2*>(y + ra * p.ldy + c);
            if (rb < p.rows) yb = *reinterpret_cast<const float2*>(y + rb * p.ldy + c);
            if (EPI == EP_GRAD_TANH) {
                v0 *= 1.f - ya.x * ya.x; v1 *= 1.f - ya.y * ya.y; v2 *= 1.f - yb.x * yb.x; v3 *= 1.f - yb.y * yb.y;
            } else {
                // EP_GRAD_MIX: columns 16..79 of the 208 features are the ReLU output of the third convolution, the others
                // (direction Linear, mission table row) have no activation
                const int cg = grp * p.out_col_step + c;
                const bool relu = EPI == EP_GRAD_RELU || (cg >= 16 && cg < 80);
                if (relu) {
                    v0 = ya.x > 0.f ? v0 : 0.f; v1 = ya.y > 0.f ? v1 : 0.f; v2 = yb.x > 0.f ? v2 : 0.f; v3 = yb.y > 0.f ? v3 : 0.f;
                }
            }
        }
        if (ra < p.rows) *reinterpret_cast<float2*>(out + ra * p.ldo + c) = make_float2(v0, v1);
        if (rb < p.rows) *reinterpret_cast<float2*>(out + rb * p.ldo + c) = make_float2(v2, v3);
    }
}

// ------------------------------------------------------------------------------------------------ wgrad
// dW[n][k] += sum over rows of dZ[r][n] * X[r][k], db[n] += sum of dZ[r][n].  The reduction runs over the samples: a CTA owns a
// contiguous range of 64-row stages (double buffered with cp.async) and keeps its share of the N x K outputs in registers for all
// of them; one atomicAdd per output and CTA at the end.  Inside a k-step the 8 rows are taken in the order 0,2,4,6,1,3,5,7 (the same
// for both operands), which makes the transposed fragment reads conflict free under the 4 x odd row pitch.
struct WgradArgs {
    const float* dz; int ldz; int dz_col_step;
    const float* x; int ldx; int x_col_step;
    float* dw; int dw_step;       // [N][K] row-major (CONV: torch [co][ci][kh][kw] with k = kk*C + ci)
    float* db; int db_step;
    float* dw_hi; float* db_hi; int split;   // split > 0: output rows >= split go to dw_hi / db_hi (row - split)
    long long rows;
    int stages_per_cta;
};
constexpr int WG_ROWS = 64;

template <int N, int K, int LOADER, bool CONV, bool STRICT>
__global__ void __launch_bounds__(256, 1) wgrad_kernel(const WgradArgs p) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int LZ = N + 4, LX = K + 4, MT = N / 16, WGN = 8 / MT, NTW = (K / 8) / WGN;
    static_assert(MT * WGN == 8 && NTW * WGN * 8 == K, "warp tiling");
    constexpr int STAGE = WG_ROWS * (LZ + LX);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int mt = warp % MT, wg = warp / MT;
    const int grp = blockIdx.y;
    const float* dz = p.dz + grp * p.dz_col_step;
    const float* x = p.x + grp * p.x_col_step;
    const long long nstages = (p.rows + WG_ROWS - 1) / WG_ROWS;
    const long long s0 = (long long)blockIdx.x * p.stages_per_cta;
    const long long s1 = s0 + p.stages_per_cta < nstages ? s0 + p.stages_per_cta : nstages;
    if (s0 >= s1) return;

    float acc[NTW][4];
#pragma unroll
    for (int i = 0; i < NTW; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    float bs0 = 0.f, bs1 = 0.f;

    auto issue = [&](long long s, int buf) {
        float* zt = smem_f + buf * STAGE;
        float* xt = zt + WG_ROWS * LZ;
        load_rows_tile<N, LD_PLAIN>(zt, dz, p.ldz, s * WG_ROWS, p.rows, WG_ROWS, tid, 256);
        load_rows_tile<K, LOADER>(xt, x, p.ldx, s * WG_ROWS, p.rows, WG_ROWS, tid, 256);
        cp_async_commit();
    };
    issue(s0, 0);
    for (long long s = s0; s < s1; ++s) {
        const int buf = (int)((s - s0) & 1);
        if (s + 1 < s1) { issue(s + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float* zt = smem_f + buf * STAGE;
        const float* xt = zt + WG_ROWS * LZ;
#pragma unroll 2
        for (int ks = 0; ks < WG_ROWS / 8; ++ks) {
            const float* zr = zt + (ks * 8 + 2 * t) * LZ + mt * 16 + g;     // rows 2t (k = t) and 2t + 1 (k = t + 4)
            const float v[4] = {zr[0], zr[8], zr[LZ], zr[LZ + 8]};
            AFrag af;
            make_afrag<STRICT>(af, v);
            if (wg == 0) { bs0 += v[0] + v[2]; bs1 += v[1] + v[3]; }
            const float* xr = xt + (ks * 8 + 2 * t) * LX + wg * (NTW * 8) + g;
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) {
                const float x0 = xr[nt * 8], x1 = xr[nt * 8 + LX];
                if (STRICT) {
                    const uint32_t h0 = to_tf32(x0), h1 = to_tf32(x1);
                    const uint32_t l0 = to_tf32(x0 - __uint_as_float(h0)), l1 = to_tf32(x1 - __uint_as_float(h1));
                    mma8(acc[nt], af.lo, h0, h1);
                    mma8(acc[nt], af.hi, l0, l1);
                    mma8(acc[nt], af.hi, h0, h1);
                } else {
                    mma8(acc[nt], af.hi, __float_as_uint(x0), __float_as_uint(x1));   // X is truncated to TF32 by the tensor core
                }
            }
        }
        __syncthreads();
    }
    float* dw = p.dw + grp * p.dw_step;
    constexpr int C = K / 4;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = mt * 16 + g + (i >> 1) * 8;
            const int n = wg * (NTW * 8) + nt * 8 + 2 * t + (i & 1);
            float* dst = dw;
            if (p.split > 0 && m >= p.split) { dst = p.dw_hi; m -= p.split; }
            const int off = CONV ? m * K + (n % C) * 4 + n / C : m * K + n;
            atomicAdd(dst + off, acc[nt][i]);
        }
    }
    if (wg == 0 && p.db) {
        bs0 += __shfl_xor_sync(0xffffffffu, bs0, 1); bs0 += __shfl_xor_sync(0xffffffffu, bs0, 2);
        bs1 += __shfl_xor_sync(0xffffffffu, bs1, 1); bs1 += __shfl_xor_sync(0xffffffffu, bs1, 2);
        if (t == 0) {
            float* db = p.db + grp * p.db_step;
            int m = mt * 16 + g;
            if (p.split > 0 && m >= p.split) { db = p.db_hi; m -= p.split; }     // (split is a multiple of 16)
            atomicAdd(db + m, bs0);
            atomicAdd(db + m + 8, bs1);
        }
    }
}

// ------------------------------------------------------------------------------------------------ samples
struct RolloutView {
    const uint8_t *frames, *dirs, *mission, *age, *actions;
    const float *values, *logp, *adv, *ret;
    int n;
};
struct SampleBufs {
    int32_t *t, *i;
    uint8_t *age, *dcode, *act;
    long long* mrow;
    float *oldv, *oldlp, *adv, *ret;
};

// sample b = flat index idx[b] = t * N + env of the [T, N] rollout arrays (policy.py evaluate_samples)
__global__ void prep_kernel(const RolloutView v, const int32_t* __restrict__ idx, int B, SampleBufs s) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int f = idx[b];
    const int t = f / v.n, i = f - t * v.n;
    s.t[b] = t; s.i[b] = i;
    const int a = v.age[f];
    s.age[b] = (uint8_t)a;
    s.mrow[b] = (long long)v.mission[(size_t)(t + 3) * v.n + i] * 4 + a;
    uint32_t dc = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) dc |= (uint32_t)(v.dirs[(size_t)(t + k) * v.n + i] & 3) << (2 * k);
    s.dcode[b] = (uint8_t)dc;
    s.act[b] = v.actions[f];
    s.oldv[b] = v.values[f]; s.oldlp[b] = v.logp[f]; s.adv[b] = v.adv[f]; s.ret[b] = v.ret[f];
}

// advantage moments of n_mb consecutive minibatches of `batch` samples: sums[mb] = (sum, sum of squares, count) in double
__global__ void __launch_bounds__(256) adv_moments_kernel(const float* __restrict__ adv, const int32_t* __restrict__ idx, int batch, long long total,
                                                          double* __restrict__ sums) {
    const int mb = blockIdx.y;
    const long long lo = (long long)mb * batch;
    const long long hi = lo + batch < total ? lo + batch : total;
    double s1 = 0.0, s2 = 0.0;
    for (long long e = lo + blockIdx.x * 256 + threadIdx.x; e < hi; e += (long long)gridDim.x * 256) {
        const double a = (double)adv[idx[e]];
        s1 += a; s2 += a * a;
    }
    __shared__ double r1[256], r2[256];
    r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd(sums + 3 * mb, r1[0]); atomicAdd(sums + 3 * mb + 1, r2[0]);
        if (blockIdx.x == 0) atomicAdd(sums + 3 * mb + 2, (double)(hi - lo));
    }
}

// features f [B,208]: columns 0..15 = direction Linear on the stacked one-hot (a sum of <= 4 weight columns + bias), columns
// 80..207 = row mission*4 + age of the mission table; columns 16..79 are written by the third convolution
__global__ void __launch_bounds__(256) assemble_kernel(const float* __restrict__ P, const float* __restrict__ lut, const uint8_t* __restrict__ dcode,
                                                       const uint8_t* __restrict__ age, const long long* __restrict__ mrow, int B, float* __restrict__ f) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int dc = dcode[warp], a = age[warp];
    float* fr = f + (size_t)warp * 208;
    if (lane < 16) {
        float v = P[P_BD + lane];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (3 - k <= a) v += P[P_WD + lane * 16 + 4 * k + ((dc >> (2 * k)) & 3)];
        fr[lane] = v;
    }
    const float4 r = __ldg(reinterpret_cast<const float4*>(lut + (size_t)mrow[warp] * 128) + lane);
    *reinterpret_cast<float4*>(fr + 80 + lane * 4) = r;
}

// gradient of the direction Linear: dW[o][4k + dir_k] += df[b][o] for the frames k that exist, db[o] += df[b][o].  Thread =
// (output o, row group): 17 register accumulators, one shared-memory reduction per CTA at the end.
__global__ void __launch_bounds__(256) dir_grad_kernel(const float* __restrict__ df, const uint8_t* __restrict__ dcode, const uint8_t* __restrict__ age,
                                                       int B, float* __restrict__ G) {
    __shared__ float red[16][17 * 16 + 1];
    const int o = threadIdx.x & 15, grp = threadIdx.x >> 4;
    float acc[17];
#pragma unroll
    for (int e = 0; e < 17; ++e) acc[e] = 0.f;
    for (int b = blockIdx.x * 16 + grp; b < B; b += gridDim.x * 16) {
        const float v = df[(size_t)b * 208 + o];
        const int dc = dcode[b], a = age[b];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int d = (dc >> (2 * k)) & 3;
            const float vk = 3 - k <= a ? v : 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[4 * k + q] += d == q ? vk : 0.f;
        }
        acc[16] += v;
    }
#pragma unroll
    for (int e = 0; e < 17; ++e) red[grp][e * 16 + o] = acc[e];
    __syncthreads();
    for (int e = threadIdx.x; e < 272; e += 256) {
        float v = 0.f;
        for (int g2 = 0; g2 < 16; ++g2) v += red[g2][e];
        const int in = e >> 4, oo = e & 15;            // e = in * 16 + o
        if (v != 0.f) atomicAdd(G + (in < 16 ? P_WD + oo * 16 + in : P_BD + oo), v);
    }
}

// gradient of the mission table rows: out[row[b]][:] += d[b][0..127] (d rows `ld` floats apart); see lut_grad_kernel in
// mgrl_policy.cu for the shared-memory accumulation scheme
__global__ void __launch_bounds__(512) lut_grad_strided_kernel(const float* __restrict__ d, int ld, const long long* __restrict__ row, int batch,
                                                               int n_rows, float* __restrict__ out) {
    extern __shared__ float tab[];                 // [n_rows][4][32]: logical column lane * 4 + j at j * 32 + lane
    const int tid = threadIdx.x, lane = tid & 31;
    for (int e = tid; e < n_rows * 128; e += blockDim.x) tab[e] = 0.f;
    __syncthreads();
    const int warps = (gridDim.x * blockDim.x) >> 5;
    int b = (blockIdx.x * blockDim.x + tid) >> 5;
    for (; b + 3 * warps < batch; b += 4 * warps) {
        float4 v[4];
        int r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            r[u] = (int)row[b + u * warps];
            v[u] = __ldg(reinterpret_cast<const float4*>(d + (size_t)(b + u * warps) * ld) + lane);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float* tp = tab + r[u] * 128 + lane;
            atomicAdd(tp, v[u].x); atomicAdd(tp + 32, v[u].y); atomicAdd(tp + 64, v[u].z); atomicAdd(tp + 96, v[u].w);
        }
    }
    for (; b < batch; b += warps) {
        const int r = (int)row[b];
        const float4 v = __ldg(reinterpret_cast<const float4*>(d + (size_t)b * ld) + lane);
        float* tp = tab + r * 128 + lane;
        atomicAdd(tp, v.x); atomicAdd(tp + 32, v.y); atomicAdd(tp + 64, v.z); atomicAdd(tp + 96, v.w);
    }
    __syncthreads();
    for (int e = tid; e < n_rows * 128; e += blockDim.x) {
        const float v = tab[e];
        if (v != 0.f) {
            const int r = e >> 7, q = e & 127;
            atomicAdd(out + r * 128 + (q & 31) * 4 + (q >> 5), v);
        }
    }
}

// ------------------------------------------------------------------------------------------------ heads + loss
struct LossArgs {
    const float* P;          // flat parameters (action_net, value_net)
    const float* a2;         // [B,128] second-layer activations: pi 0..63 | vf 64..127
    const uint8_t* act;
    const float *oldv, *oldlp, *adv, *ret;
    const double* sums;      // (sum, sum of squares, count) of the minibatch advantages over all ranks, or null
    float* dz2;              // [B,128] d loss / d second-layer pre-activation
    float* G;                // flat gradients (action_net / value_net entries are accumulated here)
    float* stats;            // [4]: sum of policy loss, value loss, -entropy terms, (unused) — divided by B by the caller
    float* logits_out;       // [B,7] or null
    float* values_out;       // [B] or null
    int B;
    float clip, clip_vf, ent_coef, vf_coef;   // clip_vf < 0: no value clipping
    int normalize;
};
constexpr int LS_ROWS = 128, LS_LD = 129;

__global__ void __launch_bounds__(LS_ROWS) loss_kernel(const LossArgs p) {
    extern __shared__ __align__(16) float smem_f[];
    float* tile = smem_f;                         // [128][129]
    float* dl = tile + LS_ROWS * LS_LD;           // [128][8]: dlogits 0..6, dvalue 7
    float* wa = dl + LS_ROWS * 8;                 // [7][64] action_net.weight, then value_net.weight [64], biases [8]
    __shared__ float red[3][LS_ROWS / 32];
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * LS_ROWS;
    for (int e = tid; e < 512; e += LS_ROWS) {            // wa[c][8] = action_net.weight[0..6][c], value_net.weight[c]
        const int c = e >> 3, j = e & 7;
        wa[e] = j < 7 ? p.P[P_WA + j * 64 + c] : p.P[P_WV + c];
    }
    if (tid < 7) wa[512 + tid] = p.P[P_BA + tid];
    if (tid == 7) wa[512 + 7] = p.P[P_BV];
    for (int e = tid; e < LS_ROWS * 32; e += LS_ROWS) {
        const int r = e >> 5, c4 = e & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < p.B) v = *reinterpret_cast<const float4*>(p.a2 + (row0 + r) * 128 + c4 * 4);
        float* d = tile + r * LS_LD + c4 * 4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();

    const long long b = row0 + tid;
    const bool valid = b < p.B;
    const float* h = tile + tid * LS_LD;
    float lg[7], val = wa[512 + 7];
#pragma unroll
    for (int j = 0; j < 7; ++j) lg[j] = wa[512 + j];
    for (int c = 0; c < 64; ++c) {
        const float hp = h[c], hv = h[64 + c];
        const float4 w0 = *reinterpret_cast<const float4*>(wa + c * 8), w1 = *reinterpret_cast<const float4*>(wa + c * 8 + 4);
        lg[0] = fmaf(w0.x, hp, lg[0]); lg[1] = fmaf(w0.y, hp, lg[1]); lg[2] = fmaf(w0.z, hp, lg[2]); lg[3] = fmaf(w0.w, hp, lg[3]);
        lg[4] = fmaf(w1.x, hp, lg[4]); lg[5] = fmaf(w1.y, hp, lg[5]); lg[6] = fmaf(w1.z, hp, lg[6]);
        val = fmaf(w1.w, hv, val);
    }
    float s_pl = 0.f, s_vl = 0.f, s_el = 0.f;
    float dlg[7], dv = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) dlg[j] = 0.f;
    if (valid) {
        if (p.logits_out) { for (int j = 0; j < 7; ++j) p.logits_out[b * 7 + j] = lg[j]; }
        if (p.values_out) p.values_out[b] = val;
        float mx = lg[0];
#pragma unroll
        for (int j = 1; j < 7; ++j) mx = fmaxf(mx, lg[j]);
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) se += expf(lg[j] - mx);
        const float lse = mx + logf(se);
        float lp[7], pr[7], ent = 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) { lp[j] = lg[j] - lse; pr[j] = expf(lp[j]); ent -= pr[j] * lp[j]; }
        const int a = p.act[b];
        float lpa = lp[0];
#pragma unroll
        for (int j = 1; j < 7; ++j) lpa = a == j ? lp[j] : lpa;
        float adv = p.adv[b];
        if (p.normalize && p.sums[2] > 1.0) {
            // mean and unbiased std of the minibatch (all ranks): torch `adv.mean()`, `adv.std()`
            const double n = p.sums[2], mean = p.sums[0] / n;
            double var = (p.sums[1] - n * mean * mean) / (n - 1.0);
            var = var > 0.0 ? var : 0.0;
            adv = (adv - (float)mean) / ((float)sqrt(var) + 1e-8f);
        }
        const float invB = 1.f / (float)p.B;
        const float ratio = expf(lpa - p.oldlp[b]);
        const float lo = 1.f - p.clip, hi = 1.f + p.clip;
        const float s1 = adv * ratio, s2 = adv * fminf(fmaxf(ratio, lo), hi);
        s_pl = -fminf(s1, s2);
        // torch.min(a, b) backward: the smaller one gets the gradient, a tie splits it; clamp passes it inside [lo, hi]
        const float g1 = s1 < s2 ? 1.f : (s1 == s2 ? 0.5f : 0.f), g2 = 1.f - g1;
        const float inr = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
        const float dlpa = -(g1 + g2 * inr) * adv * ratio * invB;
        // entropy bonus: ent_coef * mean(sum_j p_j lp_j)
        float dlp[7], sum_dlp = 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            dlp[j] = p.ent_coef * invB * pr[j] * (lp[j] + 1.f) + (a == j ? dlpa : 0.f);
            sum_dlp += dlp[j];
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) dlg[j] = dlp[j] - pr[j] * sum_dlp;
        s_el = -ent;
        // value loss: vf_coef * mean((ret - vp)^2), vp = old + clamp(v - old, -c, c)
        const float ov = p.oldv[b];
        float vp = val, pass = 1.f;
        if (p.clip_vf >= 0.f) {
            const float d = val - ov;
            vp = ov + fminf(fmaxf(d, -p.clip_vf), p.clip_vf);
            pass = (d >= -p.clip_vf && d <= p.clip_vf) ? 1.f : 0.f;
        }
        const float e = p.ret[b] - vp;
        s_vl = e * e;
        dv = -2.f * e * p.vf_coef * invB * pass;
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) dl[tid * 8 + j] = dlg[j];
    dl[tid * 8 + 7] = dv;
    // loss terms: warp shuffle -> one atomic per CTA
    for (int o = 16; o > 0; o >>= 1) {
        s_pl += __shfl_xor_sync(0xffffffffu, s_pl, o); s_vl += __shfl_xor_sync(0xffffffffu, s_vl, o); s_el += __shfl_xor_sync(0xffffffffu, s_el, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = s_pl; red[1][tid >> 5] = s_vl; red[2][tid >> 5] = s_el; }
    __syncthreads();
    if (tid < 3) {
        float s = 0.f;
        for (int w = 0; w < LS_ROWS / 32; ++w) s += red[tid][w];
        atomicAdd(p.stats + tid, s);
    }
    // head weight gradients: thread (j, c) sums dl[r][j] * h[r][c] over the tile's rows.  576 = 7*64 + 64 (value_net) products
    // + 8 bias sums per row; 128 threads take 4-5 outputs each
    for (int o = tid; o < 448 + 64 + 8; o += LS_ROWS) {
        float s = 0.f;
        if (o < 448) {
            const int j = o >> 6, c = o & 63;
            for (int r = 0; r < LS_ROWS; ++r) s = fmaf(dl[r * 8 + j], tile[r * LS_LD + c], s);
            atomicAdd(p.G + P_WA + o, s);
        } else if (o < 512) {
            const int c = o - 448;
            for (int r = 0; r < LS_ROWS; ++r) s = fmaf(dl[r * 8 + 7], tile[r * LS_LD + 64 + c], s);
            atomicAdd(p.G + P_WV + c, s);
        } else {
            const int j = o - 512;
            for (int r = 0; r < LS_ROWS; ++r) s += dl[r * 8 + j];
            atomicAdd(p.G + (j < 7 ? P_BA + j : P_BV), s);
        }
    }
    __syncthreads();
    // d loss / d pre-activation of the second layer, in place over the tile, then one coalesced store
    {
        float* hrow = tile + tid * LS_LD;
        for (int c = 0; c < 64; ++c) {
            const float4 w0 = *reinterpret_cast<const float4*>(wa + c * 8), w1 = *reinterpret_cast<const float4*>(wa + c * 8 + 4);
            float dp = dlg[0] * w0.x;
            dp = fmaf(dlg[1], w0.y, dp); dp = fmaf(dlg[2], w0.z, dp); dp = fmaf(dlg[3], w0.w, dp);
            dp = fmaf(dlg[4], w1.x, dp); dp = fmaf(dlg[5], w1.y, dp); dp = fmaf(dlg[6], w1.z, dp);
            const float hp = hrow[c], hv = hrow[64 + c];
            hrow[c] = dp * (1.f - hp * hp);
            hrow[64 + c] = dv * w1.w * (1.f - hv * hv);
        }
    }
    __syncthreads();
    for (int e = tid; e < LS_ROWS * 32; e += LS_ROWS) {
        const int r = e >> 5, c4 = e & 31;
        if (row0 + r < p.B) {
            const float* s = tile + r * LS_LD + c4 * 4;
            *reinterpret_cast<float4*>(p.dz2 + (row0 + r) * 128 + c4 * 4) = make_float4(s[0], s[1], s[2], s[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ mission GRU
// GI[v][j] = b_ih[j] + W_ih[j][:] . Emb[v][:]  (the input half of the gates depends on the token only: 32 tokens)
__global__ void gru_gi_kernel(const float* __restrict__ P, float* __restrict__ gi) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 32 * G3) return;
    const int v = e / G3, j = e - v * G3;
    float s = P[P_BIH + j];
    for (int c = 0; c < 32; ++c) s = fmaf(P[P_WIH + j * 32 + c], P[P_EMB + v * 32 + c], s);
    gi[e] = s;
}

// gate non-linearities on the exp2 / reciprocal units (a few ulp each): the 128-step recurrence is a latency chain, and the IEEE
// divide + tanhf were half of a step
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) {
    const float ax = fabsf(x);
    const float e = __expf(-2.f * ax);                       // in (0, 1]: no overflow
    const float r = __fdividef(1.f - e, 1.f + e);
    return copysignf(r, x);
}

// Two sequences per CTA, 384 threads, W_hh in REGISTERS for all 128 steps: thread (rq = tid / 4, ks = tid % 4) keeps rows
// 4rq..4rq+3 x columns 32ks..32ks+31 (128 floats).  A step reads only the thread's quarter of the two hidden states from shared
// memory (shared-memory wavefronts, not FMAs, bound the all-columns-per-thread layout: 3100 cycles per step against ~900 here);
// the four partial dot products meet with two quad shuffles.
// store: per (t, seq) the values the backward pass needs: r, z, n, gh_n, h_prev [5][128]; lut[seq] = final hidden state.
struct GruArgs {
    const float* P;
    const float* gi;          // [32][384]
    const uint8_t* tokens;    // [nseq][128]
    int nseq;
    float* store;             // [128][nseq][5][128] (may be null: table only)
    float* hprev_rows;        // [128 * nseq][128]   (h_{t-1} as the X operand of the W_hh weight gradient; may be null)
    float* lut;               // [nseq][128]
    // backward
    const float* dlut;        // [nseq][128]
    float* dgh_rows;          // [128 * nseq][384]
    float* dgi_tab;           // [32][384] accumulated with atomics (caller zeroes)
};
constexpr int HS = 36;        // padded quarter of a hidden state: the four quarters of a quad's 16-byte reads fall on distinct banks

__global__ void __launch_bounds__(384, 1) gru_fwd_kernel(const GruArgs p) {
    extern __shared__ __align__(16) float gi_s[];      // [32][384] input halves of the gates per token (48 KB, dynamic)
    __shared__ __align__(16) float h[2][4 * HS];
    __shared__ float gh[2][G3];
    __shared__ uint8_t tok[2][SEQ_LEN];
    const int tid = threadIdx.x, ks = tid & 3, rq = tid >> 2;
    float w[4][32];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) w[r][kk] = p.P[P_WHH + (4 * rq + r) * HID + ks * 32 + kk];
    const float bhh = p.P[P_BHH + 4 * rq + ks];        // bias of the row this lane writes
    const int seq0 = blockIdx.x * 2;
    for (int e = tid; e < 2 * SEQ_LEN; e += 384) {
        const int s = e / SEQ_LEN, q = e - s * SEQ_LEN;
        tok[s][q] = seq0 + s < p.nseq ? p.tokens[(size_t)(seq0 + s) * SEQ_LEN + q] : (uint8_t)0;
    }
    if (tid < 2 * 4 * HS) (&h[0][0])[tid] = 0.f;
    for (int e = tid; e < 32 * G3; e += 384) gi_s[e] = p.gi[e];
    __syncthreads();
    for (int t = 0; t < SEQ_LEN; ++t) {
        float acc[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            const float4 x0 = *reinterpret_cast<const float4*>(&h[0][ks * HS + k4 * 4]);
            const float4 x1 = *reinterpret_cast<const float4*>(&h[1][ks * HS + k4 * 4]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[r][0] = fmaf(w[r][k4 * 4], x0.x, acc[r][0]); acc[r][0] = fmaf(w[r][k4 * 4 + 1], x0.y, acc[r][0]);
                acc[r][0] = fmaf(w[r][k4 * 4 + 2], x0.z, acc[r][0]); acc[r][0] = fmaf(w[r][k4 * 4 + 3], x0.w, acc[r][0]);
                acc[r][1] = fmaf(w[r][k4 * 4], x1.x, acc[r][1]); acc[r][1] = fmaf(w[r][k4 * 4 + 1], x1.y, acc[r][1]);
                acc[r][1] = fmaf(w[r][k4 * 4 + 2], x1.z, acc[r][1]); acc[r][1] = fmaf(w[r][k4 * 4 + 3], x1.w, acc[r][1]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                acc[r][s] += __shfl_xor_sync(0xffffffffu, acc[r][s], 1);
                acc[r][s] += __shfl_xor_sync(0xffffffffu, acc[r][s], 2);
            }
        {   // lane ks writes row 4rq + ks
            const float v0 = ks == 0 ? acc[0][0] : ks == 1 ? acc[1][0] : ks == 2 ? acc[2][0] : acc[3][0];
            const float v1 = ks == 0 ? acc[0][1] : ks == 1 ? acc[1][1] : ks == 2 ? acc[2][1] : acc[3][1];
            gh[0][4 * rq + ks] = v0 + bhh; gh[1][4 * rq + ks] = v1 + bhh;
        }
        __syncthreads();
        if (tid < 2 * HID) {
            const int s = tid >> 7, c = tid & (HID - 1);
            const int seq = seq0 + s;
            const float* gi = gi_s + (int)tok[s][t] * G3;
            const float r = sigmoidf_(gi[c] + gh[s][c]);
            const float z = sigmoidf_(gi[HID + c] + gh[s][HID + c]);
            const float ghn = gh[s][2 * HID + c];
            const float n = tanhf_(gi[2 * HID + c] + r * ghn);
            float* hc = &h[s][(c >> 5) * HS + (c & 31)];
            const float hp = *hc;
            const float hn = (1.f - z) * n + z * hp;
            if (seq < p.nseq) {
                if (p.store) {
                    float* st = p.store + ((size_t)t * p.nseq + seq) * 5 * HID + c;
                    st[0] = r; st[HID] = z; st[2 * HID] = n; st[3 * HID] = ghn; st[4 * HID] = hp;
                }
                if (p.hprev_rows) p.hprev_rows[((size_t)t * p.nseq + seq) * HID + c] = hp;
                if (t == SEQ_LEN - 1) p.lut[(size_t)seq * HID + c] = hn;
            }
            *hc = hn;              // (s, c) is read and written by this thread only in this phase
        }
        __syncthreads();
    }
}

// Backward through time.  Thread (js = warp, kq = lane) keeps W_hh[32js .. 32js+31][4kq .. 4kq+3] in registers:
// dh_prev[k] += sum_j W[j][k] dgh[j] is a sum of twelve partial sums (one per warp), which meet in shared memory.
__global__ void __launch_bounds__(384, 1) gru_bwd_kernel(const GruArgs p) {
    extern __shared__ __align__(16) float dgi_acc[];   // [32][384] (48 KB, dynamic)
    __shared__ __align__(16) float dgh[2][G3];
    __shared__ __align__(16) float part[12][2][HID];
    __shared__ float dh[2][HID];
    __shared__ uint8_t tok[2][SEQ_LEN];
    const int tid = threadIdx.x, kq = tid & 31, js = tid >> 5;
    float w[4][32];                                    // w[kk][jj] = W_hh[32js + jj][4kq + kk]
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
        const float4 v = *reinterpret_cast<const float4*>(p.P + P_WHH + (js * 32 + jj) * HID + 4 * kq);
        w[0][jj] = v.x; w[1][jj] = v.y; w[2][jj] = v.z; w[3][jj] = v.w;
    }
    const int seq0 = blockIdx.x * 2;
    for (int e = tid; e < 32 * G3; e += 384) dgi_acc[e] = 0.f;
    for (int e = tid; e < 2 * SEQ_LEN; e += 384) {
        const int s = e / SEQ_LEN, qq = e - s * SEQ_LEN;
        tok[s][qq] = seq0 + s < p.nseq ? p.tokens[(size_t)(seq0 + s) * SEQ_LEN + qq] : (uint8_t)0;
    }
    for (int e = tid; e < 12 * 2 * HID; e += 384) (&part[0][0][0])[e] = 0.f;
    if (tid < 2 * HID) {
        const int s = tid >> 7, c = tid & (HID - 1);
        dh[s][c] = seq0 + s < p.nseq ? p.dlut[(size_t)(seq0 + s) * HID + c] : 0.f;
    }
    __syncthreads();
    float pr = 0.f, pz = 0.f, pn = 0.f, pg = 0.f, ph = 0.f;      // r, z, n, gh_n, h_prev of the step about to be processed
    auto fetch = [&](int t) {
        if (tid < 2 * HID && seq0 + (tid >> 7) < p.nseq && t >= 0) {
            const float* st = p.store + ((size_t)t * p.nseq + seq0 + (tid >> 7)) * 5 * HID + (tid & (HID - 1));
            pr = __ldg(st); pz = __ldg(st + HID); pn = __ldg(st + 2 * HID); pg = __ldg(st + 3 * HID); ph = __ldg(st + 4 * HID);
        }
    };
    fetch(SEQ_LEN - 1);
    for (int t = SEQ_LEN - 1; t >= 0; --t) {
        if (tid < 2 * HID) {
            const int s = tid >> 7, c = tid & (HID - 1);
            const int seq = seq0 + s;
            float d = dh[s][c];
#pragma unroll
            for (int q = 0; q < 12; ++q) d += part[q][s][c];
            float drp = 0.f, dzp = 0.f, dnp = 0.f, dghn = 0.f, dprev = 0.f;
            if (seq < p.nseq) {
                const float r = pr, z = pz, n = pn, ghn = pg, hp = ph;
                const float dn = d * (1.f - z);
                dnp = dn * (1.f - n * n);
                dzp = d * (hp - n) * z * (1.f - z);
                dghn = dnp * r;
                drp = dnp * ghn * r * (1.f - r);
                dprev = d * z;
                float* o = p.dgh_rows + ((size_t)t * p.nseq + seq) * G3 + c;
                o[0] = drp; o[HID] = dzp; o[2 * HID] = dghn;
                float* ga = dgi_acc + (int)tok[s][t] * G3 + c;
                atomicAdd(ga, drp); atomicAdd(ga + HID, dzp); atomicAdd(ga + 2 * HID, dnp);
            }
            dgh[s][c] = drp; dgh[s][HID + c] = dzp; dgh[s][2 * HID + c] = dghn;
            dh[s][c] = dprev;
        }
        fetch(t - 1);
        __syncthreads();
        float acc[4][2];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) acc[kk][0] = acc[kk][1] = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 x0 = *reinterpret_cast<const float4*>(&dgh[0][js * 32 + j4 * 4]);
            const float4 x1 = *reinterpret_cast<const float4*>(&dgh[1][js * 32 + j4 * 4]);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                acc[kk][0] = fmaf(w[kk][j4 * 4], x0.x, acc[kk][0]); acc[kk][0] = fmaf(w[kk][j4 * 4 + 1], x0.y, acc[kk][0]);
                acc[kk][0] = fmaf(w[kk][j4 * 4 + 2], x0.z, acc[kk][0]); acc[kk][0] = fmaf(w[kk][j4 * 4 + 3], x0.w, acc[kk][0]);
                acc[kk][1] = fmaf(w[kk][j4 * 4], x1.x, acc[kk][1]); acc[kk][1] = fmaf(w[kk][j4 * 4 + 1], x1.y, acc[kk][1]);
                acc[kk][1] = fmaf(w[kk][j4 * 4 + 2], x1.z, acc[kk][1]); acc[kk][1] = fmaf(w[kk][j4 * 4 + 3], x1.w, acc[kk][1]);
            }
        }
        *reinterpret_cast<float4*>(&part[js][0][4 * kq]) = make_float4(acc[0][0], acc[1][0], acc[2][0], acc[3][0]);
        *reinterpret_cast<float4*>(&part[js][1][4 * kq]) = make_float4(acc[0][1], acc[1][1], acc[2][1], acc[3][1]);
        __syncthreads();
    }
    for (int e = tid; e < 32 * G3; e += 384) {
        const float v = dgi_acc[e];
        if (v != 0.f) atomicAdd(p.dgi_tab + e, v);
    }
}

// from dGI [32][384]: db_ih[j] = sum_v dGI[v][j]; dW_ih[j][c] = sum_v dGI[v][j] Emb[v][c]; dEmb[v][c] = sum_j dGI[v][j] W_ih[j][c]
__global__ void gru_embed_grad_kernel(const float* __restrict__ P, const float* __restrict__ dgi, float* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < G3 * 32) {                       // dW_ih
        const int j = e >> 5, c = e & 31;
        float s = 0.f;
        for (int v = 0; v < 32; ++v) s = fmaf(dgi[v * G3 + j], P[P_EMB + v * 32 + c], s);
        G[P_WIH + e] = s;
    } else if (e < G3 * 32 + 1024) {         // dEmb
        const int r = e - G3 * 32, v = r >> 5, c = r & 31;
        float s = 0.f;
        for (int j = 0; j < G3; ++j) s = fmaf(dgi[v * G3 + j], P[P_WIH + j * 32 + c], s);
        G[P_EMB + r] = s;
    } else if (e < G3 * 32 + 1024 + G3) {    // db_ih
        const int j = e - G3 * 32 - 1024;
        float s = 0.f;
        for (int v = 0; v < 32; ++v) s += dgi[v * G3 + j];
        G[P_BIH + j] = s;
    }
}

// ------------------------------------------------------------------------------------------------ optimizer
constexpr int NORM_CTAS = 32;
__global__ void __launch_bounds__(256) gradnorm_kernel(const float* __restrict__ G, int n, float scale, double* __restrict__ partial) {
    double s = 0.0;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < n; e += gridDim.x * 256) {
        const double g = (double)(G[e] * scale);
        s += g * g;
    }
    __shared__ double r[256];
    r[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) r[threadIdx.x] += r[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = r[0];
}
// torch.nn.utils.clip_grad_norm_ (coef = max_norm / (norm + 1e-6), clamped to 1) followed by torch.optim.Adam (no weight decay)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ P, const float* __restrict__ G, float* __restrict__ M, float* __restrict__ V, int n,
                                                   float scale, const double* __restrict__ partial, float max_norm, float lr, float beta1, float beta2,
                                                   float eps, float bc1, float bc2_sqrt, float* __restrict__ norm_out) {
    double tot = 0.0;
    for (int i = 0; i < NORM_CTAS; ++i) tot += partial[i];
    const float norm = (float)sqrt(tot);
    float coef = max_norm / (norm + 1e-6f);
    coef = coef < 1.f ? coef : 1.f;
    if (max_norm <= 0.f) coef = 1.f;
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
    const float step = lr / bc1;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < n; e += gridDim.x * 256) {
        const float g = G[e] * scale * coef;
        const float m = beta1 * M[e] + (1.f - beta1) * g;
        const float v = beta2 * V[e] + (1.f - beta2) * g * g;
        M[e] = m; V[e] = v;
        P[e] -= step * m / (sqrtf(v) / bc2_sqrt + eps);
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct Ctx {
    int device = 0, max_batch = 0, nseq = 0;
    // caller-owned
    float *P = nullptr, *G = nullptr, *M = nullptr, *V = nullptr;
    const uint8_t* tokens = nullptr;
    // owned
    void* frag = nullptr; float* pbias = nullptr;
    float* canon = nullptr;          // two core-matrix images of the first MLP layer's weights (tcgen05 path)
    int32_t *t = nullptr, *i = nullptr;
    uint8_t *age = nullptr, *dcode = nullptr, *act = nullptr, *arg = nullptr;
    long long* mrow = nullptr;
    float *oldv = nullptr, *oldlp = nullptr, *adv = nullptr, *ret = nullptr;
    float *pooled = nullptr, *h2 = nullptr, *f = nullptr, *a1 = nullptr, *a2 = nullptr, *dz2 = nullptr, *dz1 = nullptr, *df = nullptr, *dh2 = nullptr,
          *dpooled = nullptr;
    float *gi = nullptr, *store = nullptr, *hprev_rows = nullptr, *dgh_rows = nullptr, *lut = nullptr, *dlut = nullptr, *dgi_tab = nullptr;
    float* stats = nullptr;
    double* partial = nullptr;
    int sms = 148;
    bool opted = false;
    // Side streams of one optimizer step: the mission GRU (forward before `assemble`, backward after `df`) and the weight-
    // gradient reductions do not sit on the chain conv1 -> ... -> loss -> ... -> conv1 backward; they are forked off the
    // caller's stream with events and joined before the function returns (also under stream capture: the fork/join
    // becomes parallel branches of the captured graph).  MGRL_UPDATE_STREAMS=0: everything on the caller's stream.
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t ev[10] = {};
    bool forked = true;
};

template <typename T>
cudaError_t dev_alloc(T** p, size_t count) { return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)); }

#define UP_TRY(expr)                                                                                          \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) {                                                                              \
            snprintf(mgrl_error_buffer(), kErrBytes, "%s: %s: %s", what, #expr, cudaGetErrorString(_e));      \
            return MGRL_ERR_CUDA;                                                                             \
        }                                                                                                     \
    } while (0)

template <int K, int NT, int LOADER, int EPI>
cudaError_t launch_rows_gemm(const GemmArgs& a, int groups, bool strict, cudaStream_t s) {
    const size_t smem = (size_t)GM_ROWS * ((EPI == EP_PATCH_ADJ && K + 4 < 68) ? 68 : (K + 4)) * sizeof(float);
    const dim3 grid((unsigned)((a.rows + GM_ROWS - 1) / GM_ROWS), (unsigned)groups);
    cudaError_t e;
    if (strict) {
        e = cudaFuncSetAttribute(rows_gemm_kernel<K, NT, LOADER, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rows_gemm_kernel<K, NT, LOADER, EPI, true><<<grid, 256, smem, s>>>(a);
    } else {
        e = cudaFuncSetAttribute(rows_gemm_kernel<K, NT, LOADER, EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rows_gemm_kernel<K, NT, LOADER, EPI, false><<<grid, 256, smem, s>>>(a);
    }
    return cudaGetLastError();
}

template <int N, int K, int LOADER, bool CONV>
cudaError_t launch_wgrad(WgradArgs a, int groups, bool strict, int sms, cudaStream_t s) {
    const size_t smem = (size_t)2 * WG_ROWS * (N + 4 + K + 4) * sizeof(float);
    const long long nstages = (a.rows + WG_ROWS - 1) / WG_ROWS;
    long long ctas = sms / groups > 0 ? sms / groups : 1;
    if (ctas > nstages) ctas = nstages;
    a.stages_per_cta = (int)((nstages + ctas - 1) / ctas);
    ctas = (nstages + a.stages_per_cta - 1) / a.stages_per_cta;
    const dim3 grid((unsigned)ctas, (unsigned)groups);
    cudaError_t e;
    if (strict) {
        e = cudaFuncSetAttribute(wgrad_kernel<N, K, LOADER, CONV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        wgrad_kernel<N, K, LOADER, CONV, true><<<grid, 256, smem, s>>>(a);
    } else {
        e = cudaFuncSetAttribute(wgrad_kernel<N, K, LOADER, CONV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        wgrad_kernel<N, K, LOADER, CONV, false><<<grid, 256, smem, s>>>(a);
    }
    return cudaGetLastError();
}

const void* frag_ptr(const Ctx* c, int pk, bool strict) {
    const size_t off = (size_t)pack_desc(pk).base * (strict ? 16 : 8);
    return reinterpret_cast<const char*>(c->frag) + off;
}
int frag_step(int pk) { const PackDesc d = pack_desc(pk); return d.KT * d.NT * 32; }

GruArgs gru_args(const Ctx* c) {
    GruArgs g = {};
    g.P = c->P; g.gi = c->gi; g.tokens = c->tokens; g.nseq = c->nseq; g.store = c->store; g.hprev_rows = c->hprev_rows; g.lut = c->lut;
    g.dlut = c->dlut; g.dgh_rows = c->dgh_rows; g.dgi_tab = c->dgi_tab;
    return g;
}

int run_gru_forward(Ctx* c, bool keep, cudaStream_t s, const char* what) {
    gru_gi_kernel<<<(32 * G3 + 255) / 256, 256, 0, s>>>(c->P, c->gi);
    GruArgs g = gru_args(c);
    if (!keep) { g.store = nullptr; g.hprev_rows = nullptr; }
    gru_fwd_kernel<<<(c->nseq + 1) / 2, 384, 32 * G3 * sizeof(float), s>>>(g);
    UP_TRY(cudaGetLastError());
    return MGRL_OK;
}

}  // namespace

extern "C" {

struct mgrl_ppo {
    Ctx c;
};

int mgrl_ppo_create(int device, int max_batch, int num_sequences, mgrl_ppo** out) {
    const char* what = "mgrl_ppo_create";
    if (!out || max_batch <= 0 || num_sequences <= 0 || num_sequences > NSEQ_MAX) {
        snprintf(mgrl_error_buffer(), kErrBytes, "%s: bad argument (max_batch > 0, 0 < num_sequences <= %d)", what, NSEQ_MAX);
        return MGRL_ERR_INVALID;
    }
    UP_TRY(cudaSetDevice(device));
    mgrl_ppo* h = new (std::nothrow) mgrl_ppo();
    if (!h) return MGRL_ERR_INVALID;
    Ctx& c = h->c;
    c.device = device; c.max_batch = max_batch; c.nseq = num_sequences;
    cudaDeviceProp prop;
    UP_TRY(cudaGetDeviceProperties(&prop, device));
    c.sms = prop.multiProcessorCount;
    const size_t B = (size_t)max_batch, R = (size_t)SEQ_LEN * num_sequences;
    UP_TRY(cudaMalloc(&c.frag, (size_t)kPackEntries * 16));
    UP_TRY(dev_alloc(&c.pbias, PB_TOTAL));
    UP_TRY(dev_alloc(&c.canon, 2 * mgrl_tc5::CANON_FLOATS));
    UP_TRY(dev_alloc(&c.t, B)); UP_TRY(dev_alloc(&c.i, B));
    UP_TRY(dev_alloc(&c.age, B)); UP_TRY(dev_alloc(&c.dcode, B)); UP_TRY(dev_alloc(&c.act, B)); UP_TRY(dev_alloc(&c.arg, B * 144));
    UP_TRY(dev_alloc(&c.mrow, B));
    UP_TRY(dev_alloc(&c.oldv, B)); UP_TRY(dev_alloc(&c.oldlp, B)); UP_TRY(dev_alloc(&c.adv, B)); UP_TRY(dev_alloc(&c.ret, B));
    UP_TRY(dev_alloc(&c.pooled, B * 144)); UP_TRY(dev_alloc(&c.h2, B * 128)); UP_TRY(dev_alloc(&c.f, B * 208));
    UP_TRY(dev_alloc(&c.a1, B * 128)); UP_TRY(dev_alloc(&c.a2, B * 128)); UP_TRY(dev_alloc(&c.dz2, B * 128));
    UP_TRY(dev_alloc(&c.dz1, B * 128)); UP_TRY(dev_alloc(&c.df, B * 208)); UP_TRY(dev_alloc(&c.dh2, B * 128));
    UP_TRY(dev_alloc(&c.dpooled, B * 144));
    UP_TRY(dev_alloc(&c.gi, 32 * G3)); UP_TRY(dev_alloc(&c.store, R * 5 * HID)); UP_TRY(dev_alloc(&c.hprev_rows, R * HID));
    UP_TRY(dev_alloc(&c.dgh_rows, R * G3)); UP_TRY(dev_alloc(&c.lut, (size_t)num_sequences * HID));
    UP_TRY(dev_alloc(&c.dlut, (size_t)num_sequences * HID)); UP_TRY(dev_alloc(&c.dgi_tab, 32 * G3));
    UP_TRY(dev_alloc(&c.stats, 8)); UP_TRY(dev_alloc(&c.partial, NORM_CTAS));
    if (const char* v = getenv("MGRL_UPDATE_STREAMS")) c.forked = v[0] != '0';
    for (int k = 0; k < 2; ++k) UP_TRY(cudaStreamCreateWithFlags(&c.side[k], cudaStreamNonBlocking));
    for (int k = 0; k < 10; ++k) UP_TRY(cudaEventCreateWithFlags(&c.ev[k], cudaEventDisableTiming));
    UP_TRY(cudaFuncSetAttribute(lut_grad_strided_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, num_sequences * 128 * 4));
    UP_TRY(cudaFuncSetAttribute(loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (LS_ROWS * LS_LD + LS_ROWS * 8 + 520) * 4));
    UP_TRY(cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * G3 * 4));
    UP_TRY(cudaFuncSetAttribute(gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * G3 * 4));
    *out = h;
    return MGRL_OK;
}

int mgrl_ppo_destroy(mgrl_ppo* h) {
    if (!h) return MGRL_OK;
    Ctx& c = h->c;
    cudaSetDevice(c.device);
    void* ptrs[] = {c.frag, c.pbias, c.canon, c.t, c.i, c.age, c.dcode, c.act, c.arg, c.mrow, c.oldv, c.oldlp, c.adv, c.ret, c.pooled, c.h2, c.f, c.a1,
                    c.a2, c.dz2, c.dz1, c.df, c.dh2, c.dpooled, c.gi, c.store, c.hprev_rows, c.dgh_rows, c.lut, c.dlut, c.dgi_tab, c.stats, c.partial};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int k = 0; k < 2; ++k) if (c.side[k]) cudaStreamDestroy(c.side[k]);
    for (int k = 0; k < 10; ++k) if (c.ev[k]) cudaEventDestroy(c.ev[k]);
    delete h;
    return MGRL_OK;
}

int mgrl_ppo_bind(mgrl_ppo* h, float* params_dev, float* grads_dev, float* adam_m_dev, float* adam_v_dev, const uint8_t* sequences_dev) {
    if (!h || !params_dev || !grads_dev || !adam_m_dev || !adam_v_dev || !sequences_dev) {
        snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_ppo_bind: null argument");
        return MGRL_ERR_INVALID;
    }
    Ctx& c = h->c;
    c.P = params_dev; c.G = grads_dev; c.M = adam_m_dev; c.V = adam_v_dev; c.tokens = sequences_dev;
    return MGRL_OK;
}

int mgrl_ppo_mission_table(mgrl_ppo* h, float* lut_out_dev, void* stream) {
    const char* what = "mgrl_ppo_mission_table";
    if (!h || !h->c.P || !lut_out_dev) { snprintf(mgrl_error_buffer(), kErrBytes, "%s: null argument or unbound context", what); return MGRL_ERR_INVALID; }
    Ctx& c = h->c;
    cudaStream_t s = (cudaStream_t)stream;
    const int rc = run_gru_forward(&c, false, s, what);
    if (rc != MGRL_OK) return rc;
    UP_TRY(cudaMemcpyAsync(lut_out_dev, c.lut, (size_t)c.nseq * HID * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return MGRL_OK;
}

int mgrl_ppo_moments(const float* adv_dev, const int32_t* idx_dev, int batch, long long total, double* sums_dev, void* stream) {
    const char* what = "mgrl_ppo_moments";
    if (!adv_dev || !idx_dev || !sums_dev || batch <= 0 || total <= 0) { snprintf(mgrl_error_buffer(), kErrBytes, "%s: bad argument", what); return MGRL_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const int n_mb = (int)((total + batch - 1) / batch);
    UP_TRY(cudaMemsetAsync(sums_dev, 0, (size_t)n_mb * 3 * sizeof(double), s));
    int gx = (batch + 256 * 8 - 1) / (256 * 8);
    gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
    adv_moments_kernel<<<dim3((unsigned)gx, (unsigned)n_mb), 256, 0, s>>>(adv_dev, idx_dev, batch, total, sums_dev);
    UP_TRY(cudaGetLastError());
    return MGRL_OK;
}

int mgrl_ppo_gradients(mgrl_ppo* h, const mgrl_rollout_view* rv, const int32_t* idx_dev, int batch, const double* adv_sums_dev,
                       const mgrl_ppo_hyper* hp, float* loss_out_dev, float* logits_out_dev, float* values_out_dev, void* stream) {
    const char* what = "mgrl_ppo_gradients";
    if (!h || !rv || !idx_dev || !hp || batch <= 0 || !h->c.P) { snprintf(mgrl_error_buffer(), kErrBytes, "%s: null argument or unbound context", what); return MGRL_ERR_INVALID; }
    Ctx& c = h->c;
    if (batch > c.max_batch) { snprintf(mgrl_error_buffer(), kErrBytes, "%s: batch exceeds the context's max_batch", what); return MGRL_ERR_INVALID; }
    if (hp->normalize_advantage && !adv_sums_dev) { snprintf(mgrl_error_buffer(), kErrBytes, "%s: normalize_advantage needs adv_sums_dev", what); return MGRL_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const bool strict = hp->strict_fp32 != 0;
    const int B = batch;
    UP_TRY(cudaMemsetAsync(c.G, 0, (size_t)P_TOTAL * sizeof(float), s));
    UP_TRY(cudaMemsetAsync(c.stats, 0, 8 * sizeof(float), s));
    UP_TRY(cudaMemsetAsync(c.dgi_tab, 0, 32 * G3 * sizeof(float), s));
    UP_TRY(cudaMemsetAsync(c.dlut, 0, (size_t)c.nseq * HID * sizeof(float), s));

    // ---- forward
    // sg: the mission GRU's stream, sw: the weight-gradient stream (both = s when not forked)
    const bool fork = c.forked;
    cudaStream_t sg = fork ? c.side[0] : s, sw = fork ? c.side[1] : s;
    int nev = 0;
    // `to` continues after everything issued on `from` so far
    auto after = [&](cudaStream_t from, cudaStream_t to) -> cudaError_t {
        if (!fork || from == to) return cudaSuccess;
        cudaEvent_t e = c.ev[nev++];
        cudaError_t ce = cudaEventRecord(e, from);
        return ce != cudaSuccess ? ce : cudaStreamWaitEvent(to, e, 0);
    };
    pack_update_kernel<<<(kPackEntries + 255) / 256, 256, 0, s>>>(c.P, (float*)c.frag, c.pbias, strict ? 1 : 0);
    UP_TRY(after(s, sg));
    { const int rc = run_gru_forward(&c, true, sg, what); if (rc != MGRL_OK) return rc; }
    RolloutView v = {rv->frames, rv->dirs, rv->mission, rv->age, rv->actions, rv->values, rv->logp, rv->adv, rv->ret, rv->num_envs};
    SampleBufs sb = {c.t, c.i, c.age, c.dcode, c.act, c.mrow, c.oldv, c.oldlp, c.adv, c.ret};
    prep_kernel<<<(B + 255) / 256, 256, 0, s>>>(v, idx_dev, B, sb);
    UP_TRY(cudaGetLastError());
    Conv1Args ca = {};
    ca.frames = rv->frames; ca.t = c.t; ca.i = c.i; ca.age = c.age; ca.w1 = c.P + P_WC1; ca.b1 = c.P + P_BC1; ca.pooled = c.pooled; ca.arg = c.arg;
    ca.n = rv->num_envs; ca.B = B; ca.onepass = strict ? 0 : 1;
    // one TF32 pass + tcgen05 enabled: the im2col-free tcgen05 kernel (mgrl_conv1_tc5.cu); MGRL_CONV1_TC5=0 keeps mma.sync
    static const bool conv1_tc5 = [] { const char* v = getenv("MGRL_CONV1_TC5"); return !(v && v[0] == '0'); }();
    if (!strict && hp->use_tcgen05 != 0 && conv1_tc5) UP_TRY(mgrl_tc5::launch_conv1_pool_fwd(ca, s));
    else UP_TRY(mgrl_policy::launch_conv1_pool_fwd_tc(ca, s));
    GemmArgs g = {};
    // conv2: patches [4B,64] -> h2 [4B,32] = [B,128] (oh, ow, c2)
    g = GemmArgs{}; g.a = c.pooled; g.frag = frag_ptr(&c, PK_C2F, strict); g.bias = c.pbias + PB_C2; g.out = c.h2; g.ldo = 32; g.rows = 4LL * B;
    UP_TRY((launch_rows_gemm<64, 4, LD_PATCH, EP_BIAS_RELU>(g, 1, strict, s)));
    // conv3: h2 [B,128] -> f[:, 16:80]
    g = GemmArgs{}; g.a = c.h2; g.lda = 128; g.frag = frag_ptr(&c, PK_C3F, strict); g.bias = c.pbias + PB_C3; g.out = c.f + 16; g.ldo = 208; g.rows = B;
    UP_TRY((launch_rows_gemm<128, 8, LD_PLAIN, EP_BIAS_RELU>(g, 1, strict, s)));
    UP_TRY(after(sg, s));   // the mission table
    assemble_kernel<<<(unsigned)(((size_t)B * 32 + 255) / 256), 256, 0, s>>>(c.P, c.lut, c.dcode, c.age, c.mrow, B, c.f);
    UP_TRY(cudaGetLastError());
    // first MLP layer, pi | vf: f [B,208] -> a1 [B,128]
    const bool tc5 = !strict && hp->use_tcgen05 != 0;
    if (tc5) {
        UP_TRY(mgrl_tc5::pack_canonical(c.P, c.canon, mgrl_tc5::W_L1F, s));
        UP_TRY(mgrl_tc5::pack_canonical(c.P, c.canon + mgrl_tc5::CANON_FLOATS, mgrl_tc5::W_L1B, s));
        mgrl_tc5::Args ta = {};
        ta.a = c.f; ta.lda = 208; ta.w_canon = c.canon; ta.bias = c.pbias + PB_L1; ta.out = c.a1; ta.ldo = 128; ta.rows = B;
        UP_TRY(mgrl_tc5::launch_l1_forward(ta, s));
    } else {
        g = GemmArgs{}; g.a = c.f; g.lda = 208; g.frag = frag_ptr(&c, PK_L1F, strict); g.bias = c.pbias + PB_L1;
        g.out = c.a1; g.ldo = 128; g.rows = B;
        UP_TRY((launch_rows_gemm<208, 16, LD_PLAIN, EP_BIAS_TANH>(g, 1, strict, s)));
    }
    // second MLP layer: a1[:, g*64 ..] -> a2[:, g*64 ..]
    g = GemmArgs{}; g.a = c.a1; g.lda = 128; g.a_col_step = 64; g.frag = frag_ptr(&c, PK_L2F, strict); g.frag_step = frag_step(PK_L2F);
    g.bias = c.pbias + PB_L2; g.bias_step = 64; g.out = c.a2; g.ldo = 128; g.out_col_step = 64; g.rows = B;
    UP_TRY((launch_rows_gemm<64, 8, LD_PLAIN, EP_BIAS_TANH>(g, 2, strict, s)));

    // ---- heads, loss, head gradients
    LossArgs la = {};
    la.P = c.P; la.a2 = c.a2; la.act = c.act; la.oldv = c.oldv; la.oldlp = c.oldlp; la.adv = c.adv; la.ret = c.ret; la.sums = adv_sums_dev;
    la.dz2 = c.dz2; la.G = c.G; la.stats = c.stats; la.logits_out = logits_out_dev; la.values_out = values_out_dev; la.B = B;
    la.clip = hp->clip_range; la.clip_vf = hp->clip_range_vf; la.ent_coef = hp->ent_coef; la.vf_coef = hp->vf_coef;
    la.normalize = hp->normalize_advantage;
    loss_kernel<<<(B + LS_ROWS - 1) / LS_ROWS, LS_ROWS, (LS_ROWS * LS_LD + LS_ROWS * 8 + 520) * 4, s>>>(la);
    UP_TRY(cudaGetLastError());

    // ---- backward
    WgradArgs w = {};
    // second layer: dW = dz2^T a1 per group, dz1 = (dz2 W2) * (1 - a1^2)
    w = WgradArgs{}; w.dz = c.dz2; w.ldz = 128; w.dz_col_step = 64; w.x = c.a1; w.ldx = 128; w.x_col_step = 64; w.dw = c.G + P_PI2; w.dw_step = P_VF2 - P_PI2;
    w.db = c.G + P_PI2B; w.db_step = P_VF2B - P_PI2B; w.rows = B;
    UP_TRY(after(s, sw));   // dz2
    UP_TRY((launch_wgrad<64, 64, LD_PLAIN, false>(w, 2, strict, c.sms, sw)));
    g = GemmArgs{}; g.a = c.dz2; g.lda = 128; g.a_col_step = 64; g.frag = frag_ptr(&c, PK_L2B, strict); g.frag_step = frag_step(PK_L2B);
    g.y = c.a1; g.ldy = 128; g.y_col_step = 64; g.out = c.dz1; g.ldo = 128; g.out_col_step = 64; g.rows = B;
    UP_TRY((launch_rows_gemm<64, 8, LD_PLAIN, EP_GRAD_TANH>(g, 2, strict, s)));
    // first layer: dW = dz1[:, g*64..]^T f, df = dz1 [W_pi; W_vf] with ReLU' on the conv columns
    w = WgradArgs{}; w.dz = c.dz1; w.ldz = 128; w.x = c.f; w.ldx = 208; w.dw = c.G + P_PI1; w.db = c.G + P_PI1B;
    w.dw_hi = c.G + P_VF1; w.db_hi = c.G + P_VF1B; w.split = 64; w.rows = B;
    UP_TRY(after(s, sw));   // dz1
    UP_TRY((launch_wgrad<128, 208, LD_PLAIN, false>(w, 1, strict, c.sms, sw)));
    if (tc5) {
        mgrl_tc5::Args ta = {};
        ta.a = c.dz1; ta.lda = 128; ta.w_canon = c.canon + mgrl_tc5::CANON_FLOATS; ta.y = c.f; ta.ldy = 208; ta.out = c.df; ta.ldo = 208; ta.rows = B;
        UP_TRY(mgrl_tc5::launch_l1_backward(ta, s));
    } else {
        g = GemmArgs{}; g.a = c.dz1; g.lda = 128; g.frag = frag_ptr(&c, PK_L1B, strict); g.frag_step = frag_step(PK_L1B);
        g.y = c.f; g.ldy = 208; g.y_col_step = 104; g.out = c.df; g.ldo = 208; g.out_col_step = 104; g.rows = B;
        UP_TRY((launch_rows_gemm<128, 13, LD_PLAIN, EP_GRAD_MIX>(g, 2, strict, s)));
    }
    // direction Linear and mission table rows, then the mission GRU's backward pass, on the GRU's stream
    UP_TRY(after(s, sg));   // df
    UP_TRY(after(s, sw));
    {
        int grid = (B + 16 * 64 - 1) / (16 * 64);
        grid = grid < 1 ? 1 : (grid > c.sms * 4 ? c.sms * 4 : grid);
        dir_grad_kernel<<<grid, 256, 0, sg>>>(c.df, c.dcode, c.age, B, c.G);
        const int lg = B < c.sms * 16 ? (B + 15) / 16 : c.sms;
        lut_grad_strided_kernel<<<lg, 512, (size_t)c.nseq * 128 * 4, sg>>>(c.df + 80, 208, c.mrow, B, c.nseq, c.dlut);
        UP_TRY(cudaGetLastError());
    }
    // mission GRU
    {
        GruArgs ga = gru_args(&c);
        gru_bwd_kernel<<<(c.nseq + 1) / 2, 384, 32 * G3 * sizeof(float), sg>>>(ga);
        UP_TRY(cudaGetLastError());
        w = WgradArgs{}; w.dz = c.dgh_rows; w.ldz = G3; w.dz_col_step = 128; w.x = c.hprev_rows; w.ldx = 128; w.dw = c.G + P_WHH; w.dw_step = 128 * 128;
        w.db = c.G + P_BHH; w.db_step = 128; w.rows = (long long)SEQ_LEN * c.nseq;
        UP_TRY((launch_wgrad<128, 128, LD_PLAIN, false>(w, 3, strict, c.sms, sg)));
        gru_embed_grad_kernel<<<(G3 * 32 + 1024 + G3 + 255) / 256, 256, 0, sg>>>(c.P, c.dgi_tab, c.G);
        UP_TRY(cudaGetLastError());
    }
    // conv3: dz3 = df[:, 16:80]; dW3 = dz3^T h2; dh2 = (dz3 W3) * (h2 > 0)
    w = WgradArgs{}; w.dz = c.df + 16; w.ldz = 208; w.x = c.h2; w.ldx = 128; w.dw = c.G + P_WC3; w.db = c.G + P_BC3; w.rows = B;
    UP_TRY((launch_wgrad<64, 128, LD_PLAIN, true>(w, 1, strict, c.sms, sw)));
    g = GemmArgs{}; g.a = c.df + 16; g.lda = 208; g.frag = frag_ptr(&c, PK_C3B, strict);
    g.y = c.h2; g.ldy = 128; g.out = c.dh2; g.ldo = 128; g.rows = B;
    UP_TRY((launch_rows_gemm<64, 16, LD_PLAIN, EP_GRAD_RELU>(g, 1, strict, s)));
    // conv2: dz2c = dh2 as [4B,32]; dW2 = dz2c^T patches; dpooled = adjoint gather of (dz2c W2)
    w = WgradArgs{}; w.dz = c.dh2; w.ldz = 32; w.x = c.pooled; w.dw = c.G + P_WC2; w.db = c.G + P_BC2; w.rows = 4LL * B;
    UP_TRY(after(s, sw));   // dh2
    UP_TRY((launch_wgrad<32, 64, LD_PATCH, true>(w, 1, strict, c.sms, sw)));
    g = GemmArgs{}; g.a = c.dh2; g.lda = 32; g.frag = frag_ptr(&c, PK_C2B, strict); g.out = c.dpooled; g.rows = 4LL * B;
    UP_TRY((launch_rows_gemm<32, 8, LD_PLAIN, EP_PATCH_ADJ>(g, 1, strict, s)));
    // conv1 (+ pool) weight gradient
    ca.dpooled = c.dpooled; ca.dw1 = c.G + P_WC1; ca.db1 = c.G + P_BC1;
    UP_TRY(mgrl_policy::launch_conv1_pool_bwd_tc(ca, s));
    UP_TRY(after(sg, s));   // join
    UP_TRY(after(sw, s));
    if (loss_out_dev) UP_TRY(cudaMemcpyAsync(loss_out_dev, c.stats, 4 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return MGRL_OK;
}

int mgrl_ppo_apply(mgrl_ppo* h, float lr, float max_grad_norm, float grad_scale, float beta1, float beta2, float eps, int step, float* norm_out_dev,
                   void* stream) {
    const char* what = "mgrl_ppo_apply";
    if (!h || !h->c.P || step < 1) { snprintf(mgrl_error_buffer(), kErrBytes, "%s: unbound context or step < 1", what); return MGRL_ERR_INVALID; }
    Ctx& c = h->c;
    cudaStream_t s = (cudaStream_t)stream;
    gradnorm_kernel<<<NORM_CTAS, 256, 0, s>>>(c.G, P_TOTAL, grad_scale, c.partial);
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    const float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    adam_kernel<<<64, 256, 0, s>>>(c.P, c.G, c.M, c.V, P_TOTAL, grad_scale, c.partial, max_grad_norm, lr, beta1, beta2, eps, bc1, bc2s, norm_out_dev);
    UP_TRY(cudaGetLastError());
    return MGRL_OK;
}

int mgrl_ppo_debug_buffer(mgrl_ppo* h, const char* name, void** out) {
    if (!h || !name || !out) return MGRL_ERR_INVALID;
    Ctx& c = h->c;
    struct { const char* n; void* p; } tab[] = {{"pooled", c.pooled}, {"h2", c.h2}, {"f", c.f}, {"a1", c.a1}, {"a2", c.a2}, {"dz2", c.dz2}, {"dz1", c.dz1},
                                                {"df", c.df}, {"dh2", c.dh2}, {"dpooled", c.dpooled}, {"lut", c.lut}, {"dlut", c.dlut}, {"stats", c.stats}};
    for (auto& e : tab) {
        const char *a = e.n, *b = name;
        while (*a && *a == *b) { ++a; ++b; }
        if (!*a && !*b) { *out = e.p; return MGRL_OK; }
    }
    snprintf(mgrl_error_buffer(), kErrBytes, "mgrl_ppo_debug_buffer: unknown buffer");
    return MGRL_ERR_INVALID;
}

int mgrl_ppo_debug_copy(mgrl_ppo* h, const char* name, float* dst_dev, long long count, void* stream) {
    void* src = nullptr;
    const int rc = mgrl_ppo_debug_buffer(h, name, &src);
    if (rc != MGRL_OK) return rc;
    const char* what = "mgrl_ppo_debug_copy";
    if (!dst_dev || count <= 0 || !src) { snprintf(mgrl_error_buffer(), kErrBytes, "%s: bad argument", what); return MGRL_ERR_INVALID; }
    UP_TRY(cudaMemcpyAsync(dst_dev, src, (size_t)count * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return MGRL_OK;
}

}  // extern "C"
