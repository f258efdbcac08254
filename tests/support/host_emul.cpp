// tests/support/host_emul.cpp — TEST INFRASTRUCTURE.
// Compiles the product's per-environment device functions (minigrid-rl_b200/csrc/mgrl_core.cuh)
// for the HOST so that the kernel logic can be compared with the CPU oracle in the CPU-only
// test tier (-m "not gpu").  Never linked into the product; the product has no CPU path.
#include <cstddef>
#include <cstring>
#include "mgrl_core.cuh"

using namespace mgrl;

static const uint32_t* host_lut() {
    static uint32_t lut[128];
    static bool init = false;
    if (!init) { fill_kind_lut(lut, 0, 1); init = true; }
    return lut;
}

// generation inputs the kernels keep in device memory: task strings and the fresh grid of the configuration
struct HostGen {
    uint32_t full[kTaskEntries * kTaskWords];
    uint32_t tasks[kTaskEntries * kTaskWords];
    uint32_t prefix[kTaskWords];
    uint32_t empty[kGridWords];
    uint32_t draws[kGenWords];
    GenIO io;
    explicit HostGen(const EnvCfg& cfg) {
        build_task_table(cfg, full);
        io.row_words = pack_task_table(full, tasks);
        build_task_prefix(cfg, prefix);
        build_empty_grid(cfg.size, empty);
        io.draws = draws; io.stride = 1; io.tasks = tasks; io.prefix = prefix; io.empty = empty;
    }
};

extern "C" {

// in-place generation of episode s->episode (what the reset kernel does)
void emul_generate(const EnvCfg* cfg, uint64_t seed, uint64_t env_id, EnvState* s) {
    HostGen g(*cfg);
    generate(*s, *cfg, seed, env_id, s->episode, g.io);
}

void emul_step(const EnvCfg* cfg, const float* lut, EnvState* s, int action, float* reward, uint8_t* term,
               uint8_t* trunc, uint8_t* carry_obs) {
    StepOut o = env_step(*s, action, cfg->size, cfg->max_steps, lut);
    *reward = o.reward; *term = o.terminated; *trunc = o.truncated; *carry_obs = o.carry_obs;
}

void emul_obs(const EnvCfg* cfg, const EnvState* s, int carrying, int layout, uint8_t* out) {
    const bool see = cfg->see_through_walls != 0;
    if (layout == OBS_HWC) encode_view<OBS_HWC>(*s, carrying, cfg->size, see, host_lut(), out);
    else if (layout == OBS_CHW) encode_view<OBS_CHW>(*s, carrying, cfg->size, see, host_lut(), out);
    else encode_view<OBS_HWC148>(*s, carrying, cfg->size, see, host_lut(), out);
}

void emul_full_obs(const EnvCfg* cfg, const EnvState* s, uint8_t* out) { encode_full(*s, cfg->size, out); }

uint32_t emul_kind_encode(int k) { return kind_encode(k); }

// DummyVecEnv-style vector step, same contract as the CUDA step kernel
void emul_vec_step(const EnvCfg* cfg, uint64_t seed, uint64_t base, int n, const float* lut, EnvState* st,
                   const uint8_t* act, uint8_t* obs, uint8_t* dir, uint8_t* mis, float* rew, uint8_t* term,
                   uint8_t* trunc, uint8_t* ep_len, uint8_t* term_obs) {
    HostGen g(*cfg);
    for (int i = 0; i < n; ++i) {
        EnvState& s = st[i];
        StepOut o = env_step(s, act[i], cfg->size, cfg->max_steps, lut);
        rew[i] = o.reward; term[i] = o.terminated; trunc[i] = o.truncated;
        const bool done = o.terminated | o.truncated;
        ep_len[i] = done ? s.step_count : 0;
        int carry = o.carry_obs;
        if (done) {
            encode_view<OBS_HWC>(s, carry, cfg->size, cfg->see_through_walls != 0, host_lut(), term_obs + (size_t)i * kObsBytes);
            // the step kernel's path: a layout prepared ahead of time is adopted by the finished env
            EnvState lay;
            memset(&lay, 0, sizeof lay);
            generate(lay, *cfg, seed, base + (uint64_t)i, s.episode, g.io);
            adopt_layout(reinterpret_cast<uint32_t*>(&s), reinterpret_cast<const uint32_t*>(&lay));
            carry = 0;
        }
        encode_view<OBS_HWC>(s, carry, cfg->size, cfg->see_through_walls != 0, host_lut(), obs + (size_t)i * kObsBytes);
        dir[i] = s.agent_dir; mis[i] = s.mission_id;
    }
}
}
