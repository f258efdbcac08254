// mgrl_wire.cu — the PCIe wire format of the host-buffer drop-in path (mgrl_vec_step_frames_host / mgrl_vec_reset_frames_host).
//
// The host path of VecEnv.step (/root/reference/src/ppo.py:159,242 through SB3's VecEnv) is bound by the device-to-host copy
// of the observation: 148 B of (type, colour, state) triples per environment and step.  A view cell has fewer than 256
// distinct triples, so the device packs every cell into ONE code byte (49 B + the step's scalars = one 64-byte record per
// environment), the records cross PCIe in ONE copy, and a small pool of host threads expands them into the caller's image /
// reward / flag arrays while the copy is still in flight: a record is exactly one 64-byte cache line of pinned memory and
// carries the step's tag byte, so a host thread knows that a record has landed by looking at the record itself (no events,
// no chunked copies: splitting the copy in eight cost 60 us per step).  The expansion is a table look-up of the format
// only; no environment logic runs on the CPU.
//
//   code = 128 | state << 3 | colour   for doors (type 4, the only objects with a state)
//        = type << 3 | colour          otherwise (type <= 10, colour <= 5)
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "mgrl.h"
#include "mgrl_wire.cuh"

extern "C" int mgrl_wire_have_ssse3(void);                                              // mgrl_wire_host.cpp
extern "C" void mgrl_wire_expand_hwc_ssse3(const uint8_t* rec, uint8_t* out, int pad148);
extern "C" void mgrl_wire_expand_chw_ssse3(const uint8_t* rec, uint8_t* out);
extern "C" int mgrl_wire_have_avx512vbmi(void);
extern "C" int mgrl_wire_expand_block_hwc_avx512(const uint8_t* recs, int count, uint8_t* out, int pitch, uint8_t tag, int tag_offset,
                                                 const volatile int* abort_flag, unsigned long long* poll_ns);
extern "C" int mgrl_wire_expand_groups_hwc148_avx512(const uint8_t* recs, int count, uint8_t* out, uint8_t tag, int tag_offset,
                                                     const volatile int* abort_flag, unsigned long long* poll_ns, uint8_t* dir,
                                                     uint8_t* mission, uint8_t* term, uint8_t* trunc, uint8_t* eplen, uint8_t* tdir,
                                                     float* reward);
extern "C" int mgrl_wire_expand_block_hwc_ssse3(const uint8_t* recs, int count, uint8_t* out, int pitch, uint8_t tag, int tag_offset,
                                                const volatile int* abort_flag, unsigned long long* poll_ns);

namespace mgrl_wire {

namespace {

constexpr int kRec = 64;           // bytes per wire record
constexpr int kCells = 49;
// record layout: [0..48] cell codes, 49 dir, 50 mission, 51 terminated, 52 truncated, 53 episode length, 54 terminal dir,
// 55 step tag (1..255, changes every step: "this cache line belongs to the current step"), [56..59] reward, [60..63] pad
constexpr int O_DIR = 49, O_MIS = 50, O_TERM = 51, O_TRUNC = 52, O_EPLEN = 53, O_TDIR = 54, O_TAG = 55, O_REW = 56;
constexpr int kBlock = 512;        // records per work item

__device__ __forceinline__ uint32_t cell_code(uint32_t t, uint32_t c, uint32_t s) {
    return t == 4u ? (128u | (s << 3) | c) : ((t << 3) | c);
}

struct PackArgs {
    const uint8_t* image;    // [n][pitch] encoded observation
    const float* reward;     // small outputs, any may be null (image-only job)
    const uint8_t *dir, *mission, *term, *trunc, *eplen, *tdir;
    uint8_t* wire;           // [n][64]
    int n, layout;
    uint32_t tag;
    int first;               // records are built for environments [first, n)
};

// 16 threads per environment: 12 x four cells, cell 48 + (dir, mission, terminated), (truncated, length, terminal dir), reward, pad
__global__ void __launch_bounds__(256) pack_codes_kernel(const PackArgs p) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int env = p.first + (gid >> 4), j = gid & 15;
    if (env >= p.n) return;
    uint32_t* out = reinterpret_cast<uint32_t*>(p.wire) + (size_t)env * 16 + j;
    auto triple = [&](int cell, uint32_t& t, uint32_t& c, uint32_t& s) {
        if (p.layout == MGRL_OBS_CHW) {
            const uint8_t* b = p.image + (size_t)env * 147 + cell;
            t = b[0]; c = b[49]; s = b[98];
        } else {
            const uint8_t* b = p.image + (size_t)env * (p.layout == MGRL_OBS_HWC148 ? 148 : 147) + 3 * cell;
            t = b[0]; c = b[1]; s = b[2];
        }
    };
    if (j < 12) {
        uint32_t w = 0;
        if (p.layout == MGRL_OBS_HWC148) {   // aligned words: cells 4j..4j+3 are bytes 12j..12j+11
            const uint32_t* src = reinterpret_cast<const uint32_t*>(p.image + (size_t)env * 148) + 3 * j;
            const uint32_t a = src[0], b = src[1], c = src[2];
            w = cell_code(a & 0xFF, (a >> 8) & 0xFF, (a >> 16) & 0xFF) |
                cell_code(a >> 24, b & 0xFF, (b >> 8) & 0xFF) << 8 |
                cell_code((b >> 16) & 0xFF, b >> 24, c & 0xFF) << 16 |
                cell_code((c >> 8) & 0xFF, (c >> 16) & 0xFF, c >> 24) << 24;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t t, c, s;
                triple(4 * j + q, t, c, s);
                w |= cell_code(t, c, s) << (8 * q);
            }
        }
        *out = w;
    } else if (j == 12) {
        uint32_t t, c, s;
        triple(48, t, c, s);
        uint32_t w = cell_code(t, c, s);
        if (p.dir) w |= (uint32_t)p.dir[env] << 8;
        if (p.mission) w |= (uint32_t)p.mission[env] << 16;
        if (p.term) w |= (uint32_t)p.term[env] << 24;
        *out = w;
    } else if (j == 13) {
        uint32_t w = 0;
        if (p.trunc) w |= p.trunc[env];
        if (p.eplen) w |= (uint32_t)p.eplen[env] << 8;
        if (p.tdir) w |= (uint32_t)p.tdir[env] << 16;
        *out = w | (p.tag << 24);
    } else if (j == 14) {
        *out = p.reward ? __float_as_uint(p.reward[env]) : 0u;
    } else {
        *out = 0u;
    }
}

// the scalars of environments whose image goes to the host by a direct copy: one 16-byte record each
//   0 dir, 1 mission, 2 terminated, 3 truncated, 4 episode length, 5 terminal dir, 7 step tag, [8..11] reward
constexpr int kSmall = 16, S_TAG = 7, S_REW = 8;
__global__ void __launch_bounds__(256) pack_small_kernel(const PackArgs p, uint4* __restrict__ out, int count) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= count) return;
    uint4 v;
    v.x = (p.dir ? p.dir[env] : 0u) | (p.mission ? (uint32_t)p.mission[env] << 8 : 0u) | (p.term ? (uint32_t)p.term[env] << 16 : 0u) |
          (p.trunc ? (uint32_t)p.trunc[env] << 24 : 0u);
    v.y = (p.eplen ? p.eplen[env] : 0u) | (p.tdir ? (uint32_t)p.tdir[env] << 8 : 0u) | (p.tag << 24);
    v.z = p.reward ? __float_as_uint(p.reward[env]) : 0u;
    v.w = 0u;
    out[env] = v;
}

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

}  // namespace

struct Job {
    const uint8_t* wire = nullptr;     // pinned host records of this job
    uint8_t* image = nullptr;          // caller's image array [n][pitch]
    int layout = 0;
    uint8_t *dir = nullptr, *mission = nullptr, *term = nullptr, *trunc = nullptr, *eplen = nullptr, *tdir = nullptr;
    float* reward = nullptr;
    uint8_t tag = 0;
    const uint8_t* small = nullptr;    // 16-byte scalar records of the environments whose image is copied directly
    // in-place observation stack (step_stacked): the terminal frames' records ride in the second staging buffer
    bool stacked = false;
    Stacked st = {};
    const uint8_t* wire_term = nullptr;
    uint8_t tag_term = 0;
};

struct Path {
    int n = 0;
    uint8_t* wire_dev[2] = {nullptr, nullptr};
    uint8_t* wire_host[2] = {nullptr, nullptr};
    uint32_t lut[256];
    // pool
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv;
    std::atomic<uint64_t> next{0};     // epoch << 32 | next item: a new epoch publishes a step's jobs; items are claimed by CAS so
                                       // that a thread still leaving the previous step can never claim (or skip) one of this step
    std::atomic<int> done{0};
    uint32_t epoch = 0;
    std::atomic<bool> stop{false};
    volatile int abort_flag = 0;       // set when the step's copy failed: give up on records that will never land
    Job jobs[2];
    int njobs = 0, nitems = 0, blocks = 0;
    bool ssse3 = false, avx512 = false, stream_stores = true, groups16 = true;
    uint8_t tag[2] = {0, 0};           // per staging buffer: bumped whenever that buffer is used (1..255)
    std::vector<uint8_t> age;          // step_stacked: frames of the current episode in an environment's stack (1..4)
    // Hybrid transfer of step(): the images of the first `nd_blocks` blocks cross PCIe as they are (148 B per environment, no
    // host work), the rest as 64-byte records expanded by the host threads.  With many cores per GPU the copy is the limit and
    // everything goes as records; with few (8 ranks sharing 16 cores) the expansion is, and part of the batch is better off
    // on the copy engine.  The split follows the threads' idle share, step by step.
    uint8_t* small_dev = nullptr;
    uint8_t* small_host = nullptr;
    int nd_blocks = 0;
    bool direct_auto = true;
    std::atomic<unsigned long long> poll_ns{0}, busy_ns{0};
    int wire_items = 0;

    bool small_live = false;           // small_host holds records of the last step (their tags may match again 255 steps later)

    // Blocks that change sides keep records of an older step on the side they join; a tag byte repeats every 255 steps, so
    // those records are marked "not landed" (tag 0 is never used) before the split changes (no copy is in flight then).
    void invalidate(int b0, int b1) {
        const size_t lo = (size_t)b0 * kBlock, hi = (size_t)b1 * kBlock < (size_t)n ? (size_t)b1 * kBlock : (size_t)n;
        for (size_t r = lo; r < hi; ++r) { wire_host[0][r * kRec + O_TAG] = 0; small_host[r * kSmall + S_TAG] = 0; }
    }
    void set_split(int want) {
        if (want == nd_blocks) return;
        invalidate(want < nd_blocks ? want : nd_blocks, want < nd_blocks ? nd_blocks : want);
        nd_blocks = want;
    }

    // Split controller: a one-dimensional search on the measured step time, four steps per sample.  A probe moves the split
    // by `probe_step` blocks; it is kept when the step got at least 3 % faster (and the next probe is twice as wide), otherwise
    // the split goes back, the step halves, the direction flips and the next probe waits (twice as long after every failure).  The threads' idle share
    // (time spent waiting for records / time inside work items) picks the first direction: threads that never wait mean
    // the expansion is the limit (more direct copies may help), threads that mostly wait mean the copy engine is.
    struct Tune {
        int probe_step = 0, dir = +1, window = 0, base_nd = 0, cooldown = 0, backoff = 1;
        double acc_us = 0, acc_idle = 0, base_us = -1;
        bool probing = false;
    } tune_state;
    void tune(double step_us, double idle) {
        Tune& t = tune_state;
        constexpr int W = 4;
        t.acc_us += step_us; t.acc_idle += idle;
        if (++t.window < W) return;
        const double mean = t.acc_us / W, idle_m = t.acc_idle / W;
        t.acc_us = t.acc_idle = 0; t.window = 0;
        const int cap = blocks - (blocks + 7) / 8;                        // at least an eighth stays on the record path
        auto clamp = [&](int v) { return v < 0 ? 0 : (v > cap ? cap : v); };
        if (t.probe_step == 0) t.probe_step = blocks >= 8 ? blocks / 4 : 1;
        if (t.probing) {
            if (mean < t.base_us * 0.97) {                                // better: keep it and go on in the same direction
                t.base_us = mean; t.base_nd = nd_blocks; t.backoff = 1;
                if (t.probe_step * 2 <= blocks / 4) t.probe_step *= 2;  // accelerate while it keeps paying
                const int next = clamp(nd_blocks + t.dir * t.probe_step);
                if (next != nd_blocks) set_split(next); else t.probing = false;
            } else {
                set_split(t.base_nd);
                t.probing = false; t.dir = -t.dir;
                t.probe_step = t.probe_step > 1 ? t.probe_step / 2 : 1;
                t.cooldown = t.backoff; t.backoff = t.backoff < 32 ? t.backoff * 2 : 32;
            }
            return;
        }
        t.base_us = mean; t.base_nd = nd_blocks;                          // (re)measured at the current split
        if (t.cooldown > 0) { --t.cooldown; return; }
        int dir = t.dir;
        if (idle_m > 0.30) dir = -1; else if (idle_m < 0.10 && nd_blocks == 0) dir = +1;
        const int next = clamp(nd_blocks + dir * t.probe_step);
        if (next != nd_blocks) { t.dir = dir; t.probing = true; set_split(next); }
    }

    struct Clock {                     // per thread and step
        unsigned long long poll = 0, skip = 0;
        bool first = true;             // the wait for the step's first record (kernel + pack + copy latency) is not idle time: skip
    };
    static unsigned long long now_ns() {
        return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }

    bool wait_record(const uint8_t* rec, uint8_t want, Clock* ck = nullptr, int tag_offset = O_TAG) const {
        // the record is one cache line written by the copy engine: its tag says whether it is this step's
        const volatile uint8_t* vt = rec + tag_offset;
        if (*vt != want) {
            const unsigned long long t0 = ck ? now_ns() : 0ull;
            while (*vt != want) {
                if (abort_flag) return false;
                cpu_relax();
            }
            if (ck) { if (ck->first) ck->skip += now_ns() - t0; else ck->poll += now_ns() - t0; }
        }
        if (ck) ck->first = false;
        std::atomic_thread_fence(std::memory_order_acquire);
        return true;
    }

    // scalars of the directly copied environments [lo, hi)
    void extract_small(const Job& jb, int lo, int hi, Clock* ck) const {
        int r = lo;
        // four records (one cache line) at a time: a 4-byte store per byte field, one 16-byte store of the rewards
        if (jb.dir && jb.mission && jb.term && jb.trunc && jb.eplen && jb.reward) {
            for (; r + 4 <= hi; r += 4) {
                const uint8_t* rec = jb.small + (size_t)r * kSmall;
                for (int q = 0; q < 4; ++q)
                    if (!wait_record(rec + q * kSmall, jb.tag, ck, S_TAG)) return;
                uint64_t w[4];
                uint32_t rew[4];
                for (int q = 0; q < 4; ++q) { memcpy(&w[q], rec + q * kSmall, 8); memcpy(&rew[q], rec + q * kSmall + S_REW, 4); }
                auto field = [&](int byte) {
                    const int sh = 8 * byte;
                    return (uint32_t)((w[0] >> sh) & 0xFF) | (uint32_t)((w[1] >> sh) & 0xFF) << 8 | (uint32_t)((w[2] >> sh) & 0xFF) << 16 |
                           (uint32_t)((w[3] >> sh) & 0xFF) << 24;
                };
                uint32_t v;
                v = field(0); memcpy(jb.dir + r, &v, 4);
                v = field(1); memcpy(jb.mission + r, &v, 4);
                v = field(2); memcpy(jb.term + r, &v, 4);
                v = field(3); memcpy(jb.trunc + r, &v, 4);
                v = field(4); memcpy(jb.eplen + r, &v, 4);
                if (jb.tdir) { v = field(5); memcpy(jb.tdir + r, &v, 4); }
                memcpy(jb.reward + r, rew, 16);
            }
        }
        for (; r < hi; ++r) {
            const uint8_t* rec = jb.small + (size_t)r * kSmall;
            if (!wait_record(rec, jb.tag, ck, S_TAG)) return;
            if (jb.dir) jb.dir[r] = rec[0];
            if (jb.mission) jb.mission[r] = rec[1];
            if (jb.term) jb.term[r] = rec[2];
            if (jb.trunc) jb.trunc[r] = rec[3];
            if (jb.eplen) jb.eplen[r] = rec[4];
            if (jb.tdir) jb.tdir[r] = rec[5];
            if (jb.reward) memcpy(jb.reward + r, rec + S_REW, 4);
        }
    }

    void expand_frame(const uint8_t* rec, uint8_t* o, int layout) const {     // one 147-byte frame
        if (layout == MGRL_OBS_CHW) {
            if (ssse3) { mgrl_wire_expand_chw_ssse3(rec, o); return; }
            for (int i = 0; i < kCells; ++i) {
                const uint32_t e = lut[rec[i]];
                o[i] = (uint8_t)e; o[49 + i] = (uint8_t)(e >> 8); o[98 + i] = (uint8_t)(e >> 16);
            }
        } else if (ssse3) {
            mgrl_wire_expand_hwc_ssse3(rec, o, 0);
        } else {
            for (int i = 0; i < kCells; ++i) {
                const uint32_t e = lut[rec[i]];
                o[3 * i] = (uint8_t)e; o[3 * i + 1] = (uint8_t)(e >> 8); o[3 * i + 2] = (uint8_t)(e >> 16);
            }
        }
    }

    // VecFrameStack(4,'first') + VecTransposeImage + Discrete2BoxWrapper + TokenizeVocabWrapper on the caller's persistent
    // arrays.  The mission of an episode never changes, so once an environment's four frames belong to one episode its
    // token rows are left alone (they are 1 KB of the 1.6 KB observation).
    void expand_stacked(const Job& jb, int lo, int hi) {
        constexpr int F = 147, TOK = 32;
        const Stacked& st = jb.st;
        for (int r = lo; r < hi; ++r) {
            const uint8_t* rec = jb.wire + (size_t)r * kRec;
            if (!wait_record(rec, jb.tag)) return;
            uint8_t* img = st.image + (size_t)r * 4 * F;
            uint8_t* sd = st.direction + (size_t)r * 16;
            int64_t* sm = st.mission + (size_t)r * 4 * TOK;
            const int64_t* tok = st.table + (size_t)rec[O_MIS] * TOK;
            if (rec[O_TERM] | rec[O_TRUNC]) {
                if (st.term_image) {
                    const uint8_t* trec = jb.wire_term + (size_t)r * kRec;
                    if (!wait_record(trec, jb.tag_term)) return;
                    uint8_t* ti = st.term_image + (size_t)r * 4 * F;
                    memcpy(ti, img + F, 3 * F);
                    expand_frame(trec, ti + 3 * F, jb.layout);
                }
                if (st.term_direction) {
                    uint8_t* td = st.term_direction + (size_t)r * 16;
                    memcpy(td, sd + 4, 12);
                    const uint32_t hot = 1u << (8 * (rec[O_TDIR] & 3));
                    memcpy(td + 12, &hot, 4);
                }
                if (st.term_mission) {
                    int64_t* tm = st.term_mission + (size_t)r * 4 * TOK;
                    memcpy(tm, sm + TOK, 3 * TOK * sizeof(int64_t));
                    memcpy(tm + 3 * TOK, sm + 3 * TOK, TOK * sizeof(int64_t));
                }
                memset(img, 0, 3 * F);
                memset(sd, 0, 12);
                memset(sm, 0, 3 * TOK * sizeof(int64_t));
                memcpy(sm + 3 * TOK, tok, TOK * sizeof(int64_t));
                age[r] = 1;
            } else {
                memmove(img, img + F, 3 * F);
                memmove(sd, sd + 4, 12);
                if (age[r] < 4) {
                    memmove(sm, sm + TOK, 3 * TOK * sizeof(int64_t));
                    memcpy(sm + 3 * TOK, tok, TOK * sizeof(int64_t));
                    age[r] += 1;
                }
            }
            expand_frame(rec, img + 3 * F, jb.layout);
            const uint32_t hot = 1u << (8 * (rec[O_DIR] & 3));
            memcpy(sd + 12, &hot, 4);
            if (jb.term) jb.term[r] = rec[O_TERM];
            if (jb.trunc) jb.trunc[r] = rec[O_TRUNC];
            if (jb.eplen) jb.eplen[r] = rec[O_EPLEN];
            if (jb.tdir) jb.tdir[r] = rec[O_TDIR];
            if (jb.reward) memcpy(jb.reward + r, rec + O_REW, 4);
        }
    }

    void expand(const Job& jb, int lo, int hi, Clock* ck = nullptr) {
        if (jb.stacked) { expand_stacked(jb, lo, hi); return; }
        const int pitch = jb.layout == MGRL_OBS_HWC148 ? 148 : 147;
        bool image_done = false;
        if (ssse3 && stream_stores && jb.layout != MGRL_OBS_CHW && ((reinterpret_cast<uintptr_t>(jb.image) + (size_t)lo * pitch) & 15) == 0) {
            // the whole block's images with aligned non-temporal stores (this also waits for every record of the block)
            unsigned long long polled = 0;
            if (ck && ck->first) { if (!wait_record(jb.wire + (size_t)lo * kRec, jb.tag, ck)) return; }
            // AVX-512 VBMI (one record = one register, a full byte permute per output vector): 64-byte aligned blocks only
            const bool wide = avx512 && ((reinterpret_cast<uintptr_t>(jb.image) + (size_t)lo * pitch) & 63) == 0;
            if (wide && pitch == 148 && groups16) {
                // groups of 16 records (37 aligned units): images realigned in registers, scalars 16 at a time
                const int want = (hi - lo) & ~15;
                const int got = mgrl_wire_expand_groups_hwc148_avx512(
                    jb.wire + (size_t)lo * kRec, hi - lo, jb.image + (size_t)lo * pitch, jb.tag, O_TAG, &abort_flag, ck ? &polled : nullptr,
                    jb.dir ? jb.dir + lo : nullptr, jb.mission ? jb.mission + lo : nullptr, jb.term ? jb.term + lo : nullptr,
                    jb.trunc ? jb.trunc + lo : nullptr, jb.eplen ? jb.eplen + lo : nullptr, jb.tdir ? jb.tdir + lo : nullptr,
                    jb.reward ? jb.reward + lo : nullptr);
                if (ck) ck->poll += polled;
                if (got < want) return;
                lo += got;                                      // a ragged tail (< 16 records) goes record by record below
            } else {
                const int got = (wide ? mgrl_wire_expand_block_hwc_avx512 : mgrl_wire_expand_block_hwc_ssse3)(
                    jb.wire + (size_t)lo * kRec, hi - lo, jb.image + (size_t)lo * pitch, pitch, jb.tag, O_TAG, &abort_flag, ck ? &polled : nullptr);
                if (ck) ck->poll += polled;
                if (got < hi - lo) return;
                image_done = true;
            }
        }
        for (int r = lo; r < hi; ++r) {
            const uint8_t* rec = jb.wire + (size_t)r * kRec;
            if (!wait_record(rec, jb.tag, image_done ? nullptr : ck)) return;
            uint8_t* o = jb.image + (size_t)r * pitch;
            if (image_done) {
            } else if (jb.layout == MGRL_OBS_CHW) {
                for (int i = 0; i < kCells; ++i) {
                    const uint32_t e = lut[rec[i]];
                    o[i] = (uint8_t)e; o[49 + i] = (uint8_t)(e >> 8); o[98 + i] = (uint8_t)(e >> 16);
                }
            } else if (ssse3) {
                mgrl_wire_expand_hwc_ssse3(rec, o, pitch == 148);
            } else {
                for (int i = 0; i < kCells - 1; ++i) {          // overlapping 4-byte stores, three bytes apart
                    const uint32_t e = lut[rec[i]];
                    memcpy(o + 3 * i, &e, 4);
                }
                const uint32_t e = lut[rec[kCells - 1]];
                if (pitch == 148) memcpy(o + 144, &e, 4);       // cell 48 + the record's pad byte (0)
                else { o[144] = (uint8_t)e; o[145] = (uint8_t)(e >> 8); o[146] = (uint8_t)(e >> 16); }
            }
            if (jb.dir) jb.dir[r] = rec[O_DIR];
            if (jb.mission) jb.mission[r] = rec[O_MIS];
            if (jb.term) jb.term[r] = rec[O_TERM];
            if (jb.trunc) jb.trunc[r] = rec[O_TRUNC];
            if (jb.eplen) jb.eplen[r] = rec[O_EPLEN];
            if (jb.tdir) jb.tdir[r] = rec[O_TDIR];
            if (jb.reward) memcpy(jb.reward + r, rec + O_REW, 4);
        }
    }

    // item i = (job, block of kBlock records), in copy order
    void run_items(uint32_t ep) {
        Clock ck;
        for (;;) {
            uint64_t v = next.load(std::memory_order_acquire);
            int i;
            for (;;) {
                if ((uint32_t)(v >> 32) != ep) return;
                i = (int)(uint32_t)v;
                if (i >= nitems) return;
                if (next.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel, std::memory_order_acquire)) break;
            }
            if (jobs[0].small) {
                // hybrid step: the record blocks [nd_blocks, blocks) in copy order, then the scalar blocks of the direct part
                const bool wire = i < wire_items;
                const int b = wire ? nd_blocks + i : i - wire_items;
                const int lo = b * kBlock, hi = lo + kBlock < n ? lo + kBlock : n;
                const unsigned long long t0 = now_ns();
                ck.poll = ck.skip = 0;
                if (!wire) ck.first = false;    // these records arrive last: waiting for them is waiting for the copy engine
                if (wire) expand(jobs[0], lo, hi, &ck); else extract_small(jobs[0], lo, hi, &ck);
                const unsigned long long spent = now_ns() - t0;       // = busy + polling (+ the skipped first wait)
                poll_ns.fetch_add(ck.poll, std::memory_order_relaxed);
                busy_ns.fetch_add(spent - ck.poll - ck.skip, std::memory_order_relaxed);
            } else {
                const int job = i / blocks, b = i - job * blocks;
                const int lo = b * kBlock, hi = lo + kBlock < n ? lo + kBlock : n;
                expand(jobs[job], lo, hi);
            }
            done.fetch_add(1, std::memory_order_release);
        }
    }

    uint32_t current_epoch() const { return (uint32_t)(next.load(std::memory_order_acquire) >> 32); }

    void worker() {
        uint32_t seen = 0;
        while (!stop.load(std::memory_order_acquire)) {
            // spin briefly for the next step (vector steps arrive back to back), then sleep
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            for (int spins = 0;; ++spins) {
                if (current_epoch() != seen) { got = true; break; }
                if (stop.load(std::memory_order_acquire)) return;
                cpu_relax();
                if ((spins & 1023) == 1023 &&
                    std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) break;
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(m);
                cv.wait_for(lk, std::chrono::milliseconds(50),
                            [&] { return current_epoch() != seen || stop.load(std::memory_order_acquire); });
                if (current_epoch() == seen) continue;
            }
            seen = current_epoch();
            run_items(seen);
        }
    }
};

Path* create(int n) {
    Path* p = new Path();
    p->n = n;
    p->blocks = (n + kBlock - 1) / kBlock;
    p->ssse3 = mgrl_wire_have_ssse3() != 0 && !getenv("MGRL_WIRE_SCALAR");
    p->stream_stores = !getenv("MGRL_WIRE_NO_STREAM");
    p->avx512 = p->ssse3 && mgrl_wire_have_avx512vbmi() != 0 && !getenv("MGRL_WIRE_NO_AVX512");
    p->groups16 = !getenv("MGRL_WIRE_NO_GROUPS");
    for (int c = 0; c < 256; ++c) {
        const uint32_t t = c >= 128 ? 4u : (uint32_t)(c >> 3), col = (uint32_t)(c & 7), s = c >= 128 ? (uint32_t)((c >> 3) & 3) : 0u;
        p->lut[c] = t | (col << 8) | (s << 16);
    }
    bool ok = true;
    for (int k = 0; k < 2 && ok; ++k) {
        ok = cudaMalloc(&p->wire_dev[k], (size_t)n * kRec) == cudaSuccess &&
             cudaHostAlloc(reinterpret_cast<void**>(&p->wire_host[k]), (size_t)n * kRec, cudaHostAllocDefault) == cudaSuccess;
        if (ok) memset(p->wire_host[k], 0, (size_t)n * kRec);       // tag 0 = nothing has landed
    }
    if (ok) {   // MGRL_WIRE_DIRECT: "auto" (default), or the fixed share of the batch whose images are copied directly (0 = none)
        const char* v = getenv("MGRL_WIRE_DIRECT");
        if (v && v[0] && strcmp(v, "auto") != 0) {
            p->direct_auto = false;
            double f = atof(v);
            f = f < 0 ? 0 : (f > 1 ? 1 : f);
            p->nd_blocks = (int)(f * p->blocks + 0.5);
        }
        ok = cudaMalloc(&p->small_dev, (size_t)n * kSmall) == cudaSuccess &&
             cudaHostAlloc(reinterpret_cast<void**>(&p->small_host), (size_t)n * kSmall, cudaHostAllocDefault) == cudaSuccess;
        if (ok) memset(p->small_host, 0, (size_t)n * kSmall);
    }
    if (!ok) { destroy(p); return nullptr; }
    int nthreads = (int)std::thread::hardware_concurrency();
    if (const char* v = getenv("LOCAL_WORLD_SIZE")) { const int w = atoi(v); if (w > 1) nthreads /= w; }
    if (const char* v = getenv("MGRL_HOST_THREADS")) nthreads = atoi(v);
    nthreads = nthreads < 1 ? 1 : (nthreads > 32 ? 32 : nthreads);
    if (nthreads > p->blocks) nthreads = p->blocks;
    for (int t = 0; t + 1 < nthreads; ++t) p->threads.emplace_back([p] { p->worker(); });   // the calling thread works too
    if (p->direct_auto) {
        // starting point of the search: few threads per GPU (ranks sharing the host's cores) cannot keep up with the copy
        // engine, many can (measured on 16 cores: 2 threads 109 M env-steps/s as records, 190 M copied directly; 16 threads
        // 392 M against 269 M)
        if (nthreads <= 2) p->nd_blocks = p->blocks * 3 / 4;
        else if (nthreads <= 4) p->nd_blocks = p->blocks / 2;
        else if (nthreads >= 12) { p->tune_state.probe_step = p->blocks >= 16 ? p->blocks / 16 : 1; p->tune_state.backoff = 16; p->tune_state.cooldown = 4; }
        const int cap = p->blocks - (p->blocks + 7) / 8;
        if (p->nd_blocks > cap) p->nd_blocks = cap;
    }
    return p;
}

void destroy(Path* p) {
    if (!p) return;
    p->stop.store(true, std::memory_order_release);
    { std::lock_guard<std::mutex> lk(p->m); }
    p->cv.notify_all();
    for (auto& t : p->threads) t.join();
    for (int k = 0; k < 2; ++k) {
        if (p->wire_dev[k]) cudaFree(p->wire_dev[k]);
        if (p->wire_host[k]) cudaFreeHost(p->wire_host[k]);
    }
    if (p->small_dev) cudaFree(p->small_dev);
    if (p->small_host) cudaFreeHost(p->small_host);
    delete p;
}

int host_threads(const Path* p) { return p ? (int)p->threads.size() + 1 : 0; }

void reset_stacked(Path* p) { p->age.assign((size_t)p->n, 1); }

static cudaError_t run_step(Path* p, const Outputs& main, const Outputs* extra, const Stacked* st, cudaStream_t s);

cudaError_t step(Path* p, const Outputs& main, const Outputs* extra, cudaStream_t s) { return run_step(p, main, extra, nullptr, s); }

cudaError_t step_stacked(Path* p, const Outputs& main, const Outputs* extra, const Stacked& st, cudaStream_t s) {
    if (p->age.size() != (size_t)p->n) reset_stacked(p);
    return run_step(p, main, extra, &st, s);
}

static cudaError_t run_step(Path* p, const Outputs& main, const Outputs* extra, const Stacked* st, cudaStream_t s) {
    static const bool debug = getenv("MGRL_WIRE_DEBUG") != nullptr;
    static double acc[2] = {0, 0}, acc_idle = 0;
    static int nacc = 0;
    const auto t_begin = std::chrono::steady_clock::now();
    auto us_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    };
    const Outputs* outs[2] = {&main, extra};
    const int njobs = extra ? 2 : 1;
    const int n = p->n;
    // hybrid: images of environments [0, first) by direct copy (un-stacked outputs without terminal images only)
    const bool hybrid = !st && !extra && p->small_dev && main.image_host;
    const int nd = hybrid ? p->nd_blocks : 0;
    const int first = nd * kBlock < n ? nd * kBlock : n;
    for (int k = 0; k < njobs; ++k) {
        const Outputs& o = *outs[k];
        p->tag[k] = (uint8_t)(p->tag[k] == 255 ? 1 : p->tag[k] + 1);
        PackArgs a = {};
        a.image = o.image_dev; a.reward = o.reward_dev; a.dir = o.dir_dev; a.mission = o.mission_dev; a.term = o.term_dev;
        a.trunc = o.trunc_dev; a.eplen = o.eplen_dev; a.tdir = o.tdir_dev; a.wire = p->wire_dev[k]; a.n = n; a.layout = o.layout;
        a.tag = p->tag[k]; a.first = first;
        cudaError_t e = cudaSuccess;
        const size_t pitch = o.layout == MGRL_OBS_HWC148 ? 148 : 147;
        // both pack kernels first, then the copies back to back (a kernel between two copies costs a bubble on the copy engine)
        if (first < n) {
            pack_codes_kernel<<<(unsigned)(((size_t)(n - first) * 16 + 255) / 256), 256, 0, s>>>(a);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && first > 0) {
            pack_small_kernel<<<(first + 255) / 256, 256, 0, s>>>(a, reinterpret_cast<uint4*>(p->small_dev), first);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && first < n)
            e = cudaMemcpyAsync(p->wire_host[k] + (size_t)first * kRec, p->wire_dev[k] + (size_t)first * kRec, (size_t)(n - first) * kRec,
                                cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && first > 0) {
            // the direct part: the images as they are, then the 16-byte scalar records (they land last: a host thread that
            // waits for them is waiting for the copy engine, which is what the split controller reads)
            e = cudaMemcpyAsync(o.image_host, o.image_dev, (size_t)first * pitch, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(p->small_host, p->small_dev, (size_t)first * kSmall, cudaMemcpyDeviceToHost, s);
        }
        if (e != cudaSuccess) return e;
        Job& jb = p->jobs[k];
        jb.wire = p->wire_host[k]; jb.image = o.image_host; jb.layout = o.layout; jb.dir = o.dir_host; jb.mission = o.mission_host;
        jb.term = o.term_host; jb.trunc = o.trunc_host; jb.eplen = o.eplen_host; jb.tdir = o.tdir_host; jb.reward = o.reward_host;
        jb.tag = p->tag[k];
        jb.stacked = false;
        jb.small = hybrid ? p->small_host : nullptr;
    }
    if (st) {   // one job: the main records, with the terminal frames' records (second buffer) read for finished environments
        Job& jb = p->jobs[0];
        jb.stacked = true; jb.st = *st;
        jb.wire_term = extra ? p->wire_host[1] : nullptr; jb.tag_term = p->tag[1];
        if (!extra) { jb.st.term_image = nullptr; }
    }
    // publish the step to the pool: the threads chase the copy, record by record
    p->njobs = st ? 1 : njobs;
    p->nitems = p->njobs * p->blocks;
    if (!hybrid && p->small_live) {   // (small_host is not a destination of this step's copies)
        for (size_t r = 0; r < (size_t)n; ++r) p->small_host[r * kSmall + S_TAG] = 0;
        p->small_live = false;
    }
    if (hybrid) p->small_live = first > 0;
    if (hybrid) {   // record blocks [nd, blocks), then the scalar blocks [0, nd)
        p->wire_items = p->blocks - nd;
        p->poll_ns.store(0, std::memory_order_relaxed);
        p->busy_ns.store(0, std::memory_order_relaxed);
    }
    p->done.store(0, std::memory_order_relaxed);
    p->abort_flag = 0;
    p->epoch += 1;
    p->next.store((uint64_t)p->epoch << 32, std::memory_order_release);
    if (!p->threads.empty()) p->cv.notify_all();
    const double t_issued = us_since(t_begin);
    p->run_items(p->epoch);
    cudaError_t err = cudaSuccess;
    for (int spins = 0; p->done.load(std::memory_order_acquire) < p->nitems; ++spins) {
        cpu_relax();
        if ((spins & 0xFFFF) == 0xFFFF && cudaStreamQuery(s) != cudaErrorNotReady) {
            // the copy is over (or failed): records that still show an old tag will never arrive
            err = cudaStreamSynchronize(s);
            if (err != cudaSuccess) p->abort_flag = 1;   // expand() gives up on missing records
        }
    }
    if (err != cudaSuccess) return err;
    err = cudaStreamSynchronize(s);
    if (hybrid && p->direct_auto && err == cudaSuccess) {
        const double poll = (double)p->poll_ns.load(std::memory_order_relaxed), busy = (double)p->busy_ns.load(std::memory_order_relaxed);
        const double idle = poll + busy > 0 ? poll / (poll + busy) : 0.0;
        p->tune(us_since(t_begin), idle);
        if (debug) acc_idle += idle;
    }
    if (debug) {
        acc[0] += t_issued; acc[1] += us_since(t_begin);
        if (++nacc == 64) {
            fprintf(stderr, "[mgrl_wire] %s us per step: issued %.1f, expanded %.1f (threads %d, direct blocks %d of %d, idle %.2f)\n",
                    st ? "stacked" : (extra ? "frames+terminal" : "frames"), acc[0] / 64, acc[1] / 64, (int)p->threads.size() + 1,
                    hybrid ? p->nd_blocks : 0, p->blocks, acc_idle / 64);
            acc_idle = 0;
            acc[0] = acc[1] = 0; nacc = 0;
        }
    }
    return err;
}

}  // namespace mgrl_wire
