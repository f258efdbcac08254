// mgrl_wire.cuh — the compact device-to-host wire format of the host-buffer drop-in path (mgrl_wire.cu)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mgrl_wire {

struct Path;   // device + pinned staging, copy events, expansion thread pool of one environment handle

// one set of per-step outputs: encoded observation on the device (layout = MGRL_OBS_*) and where it goes on the host; the
// small outputs are optional (null = not part of this set)
struct Outputs {
    int layout;
    const uint8_t* image_dev; uint8_t* image_host;
    const float* reward_dev; float* reward_host;
    const uint8_t *dir_dev, *mission_dev, *term_dev, *trunc_dev, *eplen_dev, *tdir_dev;
    uint8_t *dir_host, *mission_host, *term_host, *trunc_host, *eplen_host, *tdir_host;
};

// The SB3 observation dict of the wrapper stack (VecTransposeImage + VecFrameStack(4,'first') + Discrete2BoxWrapper +
// TokenizeVocabWrapper, /root/reference/src/ppo.py:118-126) kept IN PLACE in the caller's host arrays, the way
// VecFrameStack keeps `stacked_obs`: every step shifts an environment's three newest frames down and appends the new one
// (a finished environment restarts from zeros), and the stacked terminal observation of a finished environment goes to
// row e of the term_* arrays (optional).
struct Stacked {
    uint8_t* image;            // [n][4][147]
    uint8_t* direction;        // [n][16]  one-hot x 4 frames
    int64_t* mission;          // [n][128] tokens x 4 frames
    uint8_t* term_image;       // [n][4][147] rows of finished environments only; may be null
    uint8_t* term_direction;   // [n][16]
    int64_t* term_mission;     // [n][128]
    const int64_t* table;      // [MGRL_N_MISSIONS][32] mission id -> tokens (host)
};

Path* create(int num_envs);     // nullptr when an allocation fails
void destroy(Path* p);
int host_threads(const Path* p);
// pack -> chunked device-to-host copies -> expansion into the host arrays; returns when every host array is complete.
// `extra` (optional) is a second, image-only set (the terminal observations).
cudaError_t step(Path* p, const Outputs& main, const Outputs* extra, cudaStream_t stream);
// the same transfer, consumed into the in-place observation stack; main.image_host / extra->image_host are ignored
cudaError_t step_stacked(Path* p, const Outputs& main, const Outputs* extra, const Stacked& st, cudaStream_t stream);
// after a reset filled the stack arrays: every environment holds one frame
void reset_stacked(Path* p);

}  // namespace mgrl_wire
