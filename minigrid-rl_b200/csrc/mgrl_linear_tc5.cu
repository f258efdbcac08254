// mgrl_linear_tc5.cu — the wide row GEMMs of the PPO update (SB3 MlpExtractor inside PPO.train, /root/reference/src/ppo.py:159;
// policies.py:227-257) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 issued by ONE thread per CTA, both operands in
// shared memory (canonical no-swizzle K-major core-matrix layout), the 128 x N fp32 accumulator in TMEM, read back with
// tcgen05.ld for the epilogue.  Used by mgrl_ppo_gradients when the update runs with one TF32 pass (update_tf32, the
// reference's own setting ppo.py:29-32); the three-term split mode keeps the mma.sync kernels of mgrl_update.cu.
//
//   out[r, 0:N] = epilogue( sum_k A[r, k] * W[n, k] )        A: [rows, K] fp32 row-major (lda), W: [N, K] K-major
//
// A tile = 128 rows (TMEM lane = row).  Shared memory holds all of W and a ring of A half-tiles as 8-row x 16-byte core
// matrices (core matrices adjacent along K: LBO = 128 B; 8-row groups (K/4) * 128 B apart: SBO).  W is pre-arranged in global
// memory in exactly that image by pack_canonical_kernel, so its load is a linear 16-byte cp.async stream.
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "mgrl_linear_tc5.cuh"

namespace mgrl_tc5 {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst_sa, const void* src, bool valid) {
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_sa), "l"(src), "r"(n) : "memory");
}

// shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor): start address, leading (K) and stride
// (M/N) byte offsets in 16-byte units, version 1 (Blackwell), layout type 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// W [N][K] (value(n, k) given by the functor) -> canonical core-matrix image, rounded to TF32:
//   float offset = (n/8) * (K/4) * 32 + (k/4) * 32 + (n%8) * 4 + (k%4)
__global__ void pack_canonical_kernel(const float* __restrict__ P, float* __restrict__ out, int which) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    int N, K;
    if (which == W_L1F) { N = 128; K = 208; } else { N = 208; K = 128; }
    if (e >= N * K) return;
    const int n = e / K, k = e - n * K;
    float v;
    if (which == W_L1F) v = n < 64 ? P[OFF_PI1 + n * 208 + k] : P[OFF_VF1 + (n - 64) * 208 + k];      // W[n = out][k = in]
    else v = k < 64 ? P[OFF_PI1 + k * 208 + n] : P[OFF_VF1 + (k - 64) * 208 + n];                      // W^T: [n = in][k = out]
    const uint32_t r = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;
    out[(n >> 3) * (K / 4) * 32 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)] = __uint_as_float(r);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// tcgen05.commit: the barrier gets one arrival when every tcgen05.mma issued by this thread so far has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = __expf(-2.f * fabsf(x));
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

// Persistent, warp-specialised: one CTA per SM walks the 128-row tiles with stride gridDim.x.
//   warp 4   producer: W once, then the A tiles as K-halves into a ring of shared-memory stages (cp.async; a stage is
//            published with fence.proxy.async + mbarrier arrive once its copies have landed)
//   warp 5   one thread issues tcgen05.mma (K/16 instructions per stage) into one of TWO TMEM accumulators and commits: the
//            stage's "empty" barrier when its MMAs are done, the accumulator's "full" barrier after the tile's last one
//   warps 0-3 epilogue: TMEM lane quadrant w -> registers (tcgen05.ld, 32 columns at a time) -> bias / activation / gradient mask
//            -> global; the accumulator goes back to the MMA warp as soon as it has been read
template <int K, int N>
struct Tc5Cfg {
    static constexpr int KH0 = ((K / 2 + 15) / 16) * 16, KH1 = K - KH0;   // columns of the two stages of a tile (multiples of 16)
    static constexpr uint32_t W_BYTES = (uint32_t)N * K * 4u, STAGE_BYTES = 128u * KH0 * 4u;
    static constexpr int STAGES_FIT = (int)((227u * 1024u - 1024u - W_BYTES) / STAGE_BYTES);
    static constexpr int STAGES = STAGES_FIT > 4 ? 4 : STAGES_FIT;
    static constexpr uint32_t ACC_COLS = N <= 128 ? 128 : 256, TMEM_COLS = 2 * ACC_COLS;
    static constexpr uint32_t SMEM = W_BYTES + STAGES * STAGE_BYTES + 256;
    static_assert(STAGES >= 2 && K % 16 == 0 && KH1 > 0 && N % 16 == 0 && N <= 256, "tile shape");
};

template <int K, int N, int EPI>
__global__ void __launch_bounds__(192, 1) linear_tc5_kernel(const Args p) {
    using Cfg = Tc5Cfg<K, N>;
    constexpr int KH0 = Cfg::KH0, KH1 = Cfg::KH1, S = Cfg::STAGES;
    constexpr uint32_t LBO = 128u, SBO_W = (K / 4) * 128u;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sW = smem;
    uint8_t* sA = smem + Cfg::W_BYTES;
    const uint32_t bars = smem_u32(smem + Cfg::W_BYTES + S * Cfg::STAGE_BYTES);     // 8-byte barriers
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (S + s); };
    auto acc_full = [&](int a) { return bars + 8u * (2 * S + a); };
    auto acc_empty = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::W_BYTES + S * Cfg::STAGE_BYTES + 8 * (2 * S + 4));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (p.rows + 127) / 128;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 32); mbar_init(empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 4) {
        // ------------------------------------------------------------------ producer
        {   // all of W, in its global core-matrix image: part of the first stage's copy group
            const uint32_t sb = smem_u32(sW);
            const float4* w = reinterpret_cast<const float4*>(p.w_canon);
            for (int e = lane; e < (int)(Cfg::W_BYTES / 16); e += 32) cp_async16(sb + (uint32_t)e * 16u, w + e, true);
        }
        const int r8 = lane & 7, cq = lane >> 3;
        int it = 0, pending = -1;                         // pending: stage whose copies are in flight
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int h = 0; h < 2; ++h, ++it) {
                const int s = it % S;
                mbar_wait(empty(s), (uint32_t)(((it / S) & 1) ^ 1));
                const uint32_t sa = smem_u32(sA) + (uint32_t)s * Cfg::STAGE_BYTES;
                // a warp instruction = 8 rows x 64 bytes (quarter warp = one 16-byte chunk column of a core matrix); the single
                // producer warp must not spend instructions on addresses: one row pointer per 8-row group, constant steps inside
                auto fill = [&](auto cols_tag) {
                    constexpr int cols = decltype(cols_tag)::value, c4n = cols / 16;
                    constexpr uint32_t sbo = (uint32_t)(cols / 4) * 128u;
#pragma unroll 2
                    for (int g8 = 0; g8 < 16; ++g8) {
                        const long long grow = tile * 128 + g8 * 8 + r8;
                        const bool valid = grow < p.rows;
                        const float* src = p.a + (valid ? grow : 0) * (long long)p.lda + h * KH0 + cq * 4;
                        const uint32_t dst = sa + (uint32_t)g8 * sbo + (uint32_t)cq * LBO + (uint32_t)r8 * 16u;
#pragma unroll
                        for (int c4 = 0; c4 < c4n; ++c4) cp_async16(dst + (uint32_t)c4 * 4u * LBO, src + c4 * 16, valid);
                    }
                };
                if (h == 0) fill(std::integral_constant<int, KH0>{}); else fill(std::integral_constant<int, KH1>{});
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (pending >= 0) {                       // the stage before this one has landed: publish it
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(full(pending));
                }
                pending = s;
            }
        }
        if (pending >= 0) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full(pending));
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, N);
            const uint32_t w0 = smem_u32(sW);
            int it = 0, j = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++j) {
                const int a = j & 1;
                mbar_wait(acc_empty(a), (uint32_t)(((j >> 1) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + (uint32_t)a * Cfg::ACC_COLS;
                for (int h = 0; h < 2; ++h, ++it) {
                    const int s = it % S;
                    mbar_wait(full(s), (uint32_t)((it / S) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(sA) + (uint32_t)s * Cfg::STAGE_BYTES;
                    const uint32_t b0 = w0 + (uint32_t)h * (KH0 / 4) * LBO;
                    const int cols = h ? KH1 : KH0;
                    const uint32_t sbo = (uint32_t)(cols / 4) * 128u;
#pragma unroll 1
                    for (int kt = 0; kt < cols / 8; ++kt) {   // K 8 per instruction = two core-matrix columns (256 B further on)
                        const uint64_t da = make_desc(a0 + (uint32_t)kt * 2u * LBO, LBO, sbo);
                        const uint64_t db = make_desc(b0 + (uint32_t)kt * 2u * LBO, LBO, SBO_W);
                        mma_tf32(d, da, db, idesc, (h | kt) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty(s));                    // the stage is free once these MMAs have read it
                }
                umma_commit(acc_full(a));                     // the tile's accumulator is complete
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: warp w = TMEM lanes 32w .. 32w+31 = rows
        int j = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++j) {
            const int a = j & 1;
            mbar_wait(acc_full(a), (uint32_t)((j >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long row = tile * 128 + warp * 32 + lane;
            const uint32_t t0 = tmem + (uint32_t)a * Cfg::ACC_COLS + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < N; c0 += 32) {
                float v[32];
                tmem_ld32(t0 + (uint32_t)c0, v);
                if (c0 + 32 >= N) {                           // last read of this accumulator: hand it back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(acc_empty(a));
                }
                if (row < p.rows) {
                    float* o = p.out + row * (long long)p.ldo + c0;
                    const float* y = (EPI == EPI_GRAD_MIX) ? p.y + row * (long long)p.ldy + c0 : nullptr;
                    // 32 bytes per store instruction (st.global.v8.f32, sm_100): a thread's row chunk is a whole sector; with
                    // 16-byte stores every sector reached L2 as two half writes from different instructions
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        float x[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) x[q] = v[i + q];
                        if (EPI == EPI_BIAS_TANH) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i + 4));
                            x[0] = fast_tanh(x[0] + b0.x); x[1] = fast_tanh(x[1] + b0.y); x[2] = fast_tanh(x[2] + b0.z); x[3] = fast_tanh(x[3] + b0.w);
                            x[4] = fast_tanh(x[4] + b1.x); x[5] = fast_tanh(x[5] + b1.y); x[6] = fast_tanh(x[6] + b1.z); x[7] = fast_tanh(x[7] + b1.w);
                        } else if (EPI == EPI_GRAD_MIX) {
                            // columns 16..79 of the 208 features are ReLU outputs of the third convolution: pass where the output was > 0
                            const int c = c0 + i;
                            if (c >= 16 && c < 80) {
                                const float4 y0 = *reinterpret_cast<const float4*>(y + i), y1 = *reinterpret_cast<const float4*>(y + i + 4);
                                x[0] = y0.x > 0.f ? x[0] : 0.f; x[1] = y0.y > 0.f ? x[1] : 0.f; x[2] = y0.z > 0.f ? x[2] : 0.f; x[3] = y0.w > 0.f ? x[3] : 0.f;
                                x[4] = y1.x > 0.f ? x[4] : 0.f; x[5] = y1.y > 0.f ? x[5] : 0.f; x[6] = y1.z > 0.f ? x[6] : 0.f; x[7] = y1.w > 0.f ? x[7] : 0.f;
                            }
                        }
                        if (c0 + i < N)
                            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + i), "f"(x[0]), "f"(x[1]), "f"(x[2]),
                                         "f"(x[3]), "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7]) : "memory");
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

template <int K, int N, int EPI>
cudaError_t launch_t(const Args& a, cudaStream_t s) {
    using Cfg = Tc5Cfg<K, N>;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    cudaError_t e = cudaFuncSetAttribute(linear_tc5_kernel<K, N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    const long long ntiles = (a.rows + 127) / 128;
    const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
    linear_tc5_kernel<K, N, EPI><<<grid, 192, Cfg::SMEM, s>>>(a);
    return cudaGetLastError();
}

}  // namespace

cudaError_t pack_canonical(const float* params, float* out, int which, cudaStream_t s) {
    pack_canonical_kernel<<<(128 * 208 + 255) / 256, 256, 0, s>>>(params, out, which);
    return cudaGetLastError();
}

cudaError_t launch_l1_forward(const Args& a, cudaStream_t s) { return launch_t<208, 128, EPI_BIAS_TANH>(a, s); }
cudaError_t launch_l1_backward(const Args& a, cudaStream_t s) { return launch_t<128, 208, EPI_GRAD_MIX>(a, s); }

}  // namespace mgrl_tc5
