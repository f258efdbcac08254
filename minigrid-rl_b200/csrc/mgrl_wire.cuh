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

Path* create(int num_envs);     // nullptr when an allocation fails
void destroy(Path* p);
int host_threads(const Path* p);
// pack -> chunked device-to-host copies -> expansion into the host arrays; returns when every host array is complete.
// `extra` (optional) is a second, image-only set (the terminal observations).
cudaError_t step(Path* p, const Outputs& main, const Outputs* extra, cudaStream_t stream);

}  // namespace mgrl_wire
