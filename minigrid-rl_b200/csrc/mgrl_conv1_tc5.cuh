// mgrl_conv1_tc5.cuh — first extractor stage of the PPO update on tcgen05 / TMEM (mgrl_conv1_tc5.cu)
#pragma once
#include <cuda_runtime.h>

#include "mgrl_policy_layout.cuh"

namespace mgrl_tc5 {

// same contract as mgrl_policy::launch_conv1_pool_fwd_tc: pooled [B][9][16] and arg bytes from the rollout's frame buffer
cudaError_t launch_conv1_pool_fwd(const mgrl_policy::Conv1Args& a, cudaStream_t s);

}  // namespace mgrl_tc5
