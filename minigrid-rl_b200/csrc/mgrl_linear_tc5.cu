// mgrl_linear_tc5.cu — the wide row GEMMs of the PPO update (SB3 MlpExtractor inside PPO.train, /root/reference/src/ppo.py:159;
// policies.py:227-257) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 issued by ONE thread per CTA, both operands in
// shared memory (canonical no-swizzle K-major core-matrix layout), the 128 x N fp32 accumulator in TMEM, read back with
// tcgen05.ld for the epilogue.  Used by mgrl_ppo_gradients when the update runs with one TF32 pass (update_tf32, the
// reference's own setting ppo.py:29-32); the three-term split mode keeps the mma.sync kernels of mgrl_update.cu.
//
//   out[r, 0:N] = epilogue( sum_k A[r, k] * W[n, k] )        A: [rows, K] fp32 row-major (lda), W: [N, K] K-major
//
// One CTA = one 128-row tile (TMEM lane = row).  Shared memory: the A tile and all of W as 8-row x 16-byte core matrices
// (core matrices adjacent along K: LBO = 128 B; 8-row groups K/4 * 128 B apart: SBO).  W is pre-arranged in global memory in
// exactly that image by pack_canonical_kernel, so its load is a linear 16-byte cp.async stream.
#include <cuda_runtime.h>

#include <cstdint>

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

template <int K, int N, int EPI>
__global__ void __launch_bounds__(128, 1) linear_tc5_kernel(const Args p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr uint32_t A_BYTES = 128u * K * 4u, B_BYTES = (uint32_t)N * K * 4u;
    constexpr uint32_t LBO = 128u, SBO = (K / 4) * 128u;
    constexpr uint32_t TMEM_COLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + A_BYTES + B_BYTES + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long row0 = (long long)blockIdx.x * 128;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // A tile: a warp copies 8 rows x 4 chunks at a time (64 contiguous bytes per row from global; one 128-byte core-matrix
    // column per quarter warp into shared memory)
    {
        const uint32_t sa = smem_u32(sA);
        constexpr int CH = K / 4;                         // 16-byte chunks per row
        constexpr int UNITS = 16 * (CH / 4);              // (8-row group, 4-chunk group) units per tile
        const int r8 = lane & 7, cq = lane >> 3;
        for (int u = warp; u < UNITS; u += 4) {
            const int g8 = u / (CH / 4), c4 = u - g8 * (CH / 4);
            const int row = g8 * 8 + r8, chunk = c4 * 4 + cq;
            const long long grow = row0 + row;
            const bool valid = grow < p.rows;
            const float* src = p.a + (valid ? grow : 0) * (long long)p.lda + chunk * 4;
            cp_async16(sa + (uint32_t)g8 * SBO + (uint32_t)chunk * LBO + (uint32_t)r8 * 16u, src, valid);
        }
        const uint32_t sb = smem_u32(sB);
        const float4* w = reinterpret_cast<const float4*>(p.w_canon);
        for (int e = tid; e < (int)(B_BYTES / 16); e += 128) cp_async16(sb + (uint32_t)e * 16u, w + e, true);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // generic-proxy writes (cp.async) -> visible to the tensor core's async proxy; TMEM address -> visible to everyone
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (tid == 0) {
        constexpr uint32_t idesc = make_idesc(128, N);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll 1
        for (int kt = 0; kt < K / 8; ++kt) {              // one instruction = K 8 = two core-matrix columns (256 B further on)
            const uint64_t da = make_desc(a0 + (uint32_t)kt * 2u * LBO, LBO, SBO);
            const uint64_t db = make_desc(b0 + (uint32_t)kt * 2u * LBO, LBO, SBO);
            mma_tf32(tmem, da, db, idesc, kt > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    mbar_wait(smem_u32(mbar), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 = rows; 32 columns per tcgen05.ld
    const long long row = row0 + warp * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (row < p.rows) {
            float* o = p.out + row * (long long)p.ldo + c0;
            const float* y = (EPI == EPI_GRAD_MIX) ? p.y + row * (long long)p.ldy + c0 : nullptr;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                float4 x = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                if (EPI == EPI_BIAS_TANH) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i));
                    x.x = tanhf(x.x + b.x); x.y = tanhf(x.y + b.y); x.z = tanhf(x.z + b.z); x.w = tanhf(x.w + b.w);
                } else if (EPI == EPI_GRAD_MIX) {
                    // columns 16..79 of the 208 features are ReLU outputs of the third convolution: pass where the output was > 0
                    const int c = c0 + i;
                    if (c >= 16 && c < 80) {
                        const float4 yy = *reinterpret_cast<const float4*>(y + i);
                        x.x = yy.x > 0.f ? x.x : 0.f; x.y = yy.y > 0.f ? x.y : 0.f; x.z = yy.z > 0.f ? x.z : 0.f; x.w = yy.w > 0.f ? x.w : 0.f;
                    }
                }
                if (c0 + i < N) *reinterpret_cast<float4*>(o + i) = x;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

template <int K, int N, int EPI>
cudaError_t launch_t(const Args& a, cudaStream_t s) {
    const size_t smem = (size_t)128 * K * 4 + (size_t)N * K * 4 + 64;
    cudaError_t e = cudaFuncSetAttribute(linear_tc5_kernel<K, N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    linear_tc5_kernel<K, N, EPI><<<(unsigned)((a.rows + 127) / 128), 128, smem, s>>>(a);
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
