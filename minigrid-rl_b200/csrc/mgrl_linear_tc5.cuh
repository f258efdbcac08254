// mgrl_linear_tc5.cuh — tcgen05 / TMEM row GEMMs of the PPO update (mgrl_linear_tc5.cu)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mgrl_tc5 {

enum : int { EPI_BIAS_TANH = 0, EPI_GRAD_MIX = 1 };
enum : int { W_L1F = 0, W_L1B = 1 };                    // which weight image pack_canonical builds
constexpr int OFF_PI1 = 74624, OFF_VF1 = 92160;         // flat parameter offsets of mlp_extractor.{policy,value}_net.0.weight
constexpr int CANON_FLOATS = 128 * 208;                 // floats of one packed image (both are 128 x 208 / 208 x 128)

struct Args {
    const float* a; int lda;        // [rows, K] fp32
    const float* w_canon;           // pack_canonical image of W
    const float* bias;              // [N] (EPI_BIAS_TANH)
    const float* y; int ldy;        // activations whose ReLU derivative EPI_GRAD_MIX applies
    float* out; int ldo;
    long long rows;
};

// W -> core-matrix image (TF32-rounded): W_L1F = [pi | vf] first MLP layer as [N = 128 out][K = 208 in];
// W_L1B = its transpose [N = 208 in][K = 128 out] (the dX GEMM)
cudaError_t pack_canonical(const float* params_dev, float* out_dev, int which, cudaStream_t s);
// a1 = tanh(f W^T + b): K = 208, N = 128
cudaError_t launch_l1_forward(const Args& a, cudaStream_t s);
// df = dz1 [W_pi; W_vf] with ReLU' on the convolution columns: K = 128, N = 208
cudaError_t launch_l1_backward(const Args& a, cudaStream_t s);

}  // namespace mgrl_tc5
