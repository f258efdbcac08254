// mgrl_policy_layout.cuh — what the two policy-forward kernels (mgrl_policy.cu: fp32 on the CUDA cores,
// mgrl_policy_tc.cu: split-TF32 on the tensor cores) share: the packed weight layout, the launch arguments and the
// sampling RNG.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "mgrl.h"

namespace mgrl_policy {

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


struct Conv1Args {
    const uint8_t* frames;   // [time, n, 148]
    const int32_t* t;        // [B] time index of the sample (its newest frame is record t + 3)
    const int32_t* i;        // [B] environment
    const uint8_t* age;      // [B] frames of history available (0..3)
    const float* w1;         // [16][12][2][2] (torch layout)
    const float* b1;         // [16]
    float* pooled;           // [B][9][16] post bias + ReLU, pooled cell q = qh * 3 + qw
    uint8_t* arg;            // [B][9][16] position of the maximum (0..3) | 4 if the output is positive
    const float* dpooled;    // backward: [B][9][16]
    float* dw1;              // backward: [16][48] accumulated with atomics (caller zeroes)
    float* db1;              // backward: [16]
    int n, B;
    int onepass;             // tensor-core kernels: 1 = one TF32 pass instead of the two-term split (update_tf32)
};

// mgrl_policy_tc.cu: the tensor-core forward and the fragment packing of its weights (section of MGRL_POLICY_FRAGMENTS
// floats behind the N_WEIGHTS fp32 weights of the same buffer)
cudaError_t launch_policy_forward_tc(const PolicyArgs& a, cudaStream_t stream);
cudaError_t launch_pack_fragments(float* weights_dev, cudaStream_t stream);
// mgrl_policy_tc.cu: first extractor stage of the PPO update on the tensor cores
cudaError_t launch_conv1_pool_fwd_tc(const Conv1Args& a, cudaStream_t stream);
cudaError_t launch_conv1_pool_bwd_tc(const Conv1Args& a, cudaStream_t stream);

}  // namespace mgrl_policy
