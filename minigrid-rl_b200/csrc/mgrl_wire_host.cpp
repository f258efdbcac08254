// mgrl_wire_host.cpp — host half of the PCIe wire format (mgrl_wire.cu): one 64-byte record -> the 49 (type, colour, state)
// triples of an HWC observation record, with SSSE3 byte shuffles (16 cells per iteration).  Compiled with -mssse3; the caller
// picks this routine only when the CPU reports SSSE3 (mgrl_wire_have_ssse3), else its scalar table loop.
//   code = 128 | state << 3 | colour  for doors (type 4),  type << 3 | colour  otherwise
#include <tmmintrin.h>

#include <chrono>
#include <cstdint>
#include <cstring>

extern "C" int mgrl_wire_have_ssse3(void) {
    __builtin_cpu_init();
    return __builtin_cpu_supports("ssse3") ? 1 : 0;
}

namespace {
struct Masks {
    __m128i t[3], c[3], s[3];
};
inline Masks make_masks() {
    Masks m;
    alignas(16) uint8_t b[3][3][16];
    for (int v = 0; v < 3; ++v)
        for (int j = 0; j < 16; ++j) {
            const int byte = 16 * v + j, q = byte / 3, r = byte % 3;     // output byte = component r of cell q (of 16)
            for (int comp = 0; comp < 3; ++comp) b[comp][v][j] = (uint8_t)(r == comp ? q : 0x80);
        }
    for (int v = 0; v < 3; ++v) {
        m.t[v] = _mm_load_si128(reinterpret_cast<const __m128i*>(b[0][v]));
        m.c[v] = _mm_load_si128(reinterpret_cast<const __m128i*>(b[1][v]));
        m.s[v] = _mm_load_si128(reinterpret_cast<const __m128i*>(b[2][v]));
    }
    return m;
}
}  // namespace

// rec: 49 code bytes (at least 64 readable); out: 147 bytes (pad148 != 0: 148, the last byte written 0)
extern "C" void mgrl_wire_expand_hwc_ssse3(const uint8_t* rec, uint8_t* out, int pad148) {
    static const Masks m = make_masks();
    const __m128i seven = _mm_set1_epi8(7), three = _mm_set1_epi8(3), four = _mm_set1_epi8(4), lo5 = _mm_set1_epi8(0x1F);
    for (int g = 0; g < 3; ++g) {
        const __m128i code = _mm_loadu_si128(reinterpret_cast<const __m128i*>(rec + 16 * g));
        const __m128i door = _mm_cmpgt_epi8(_mm_setzero_si128(), code);                       // code >= 128
        const __m128i hi = _mm_and_si128(_mm_srli_epi16(code, 3), lo5);                        // code >> 3 per byte
        const __m128i T = _mm_or_si128(_mm_andnot_si128(door, hi), _mm_and_si128(door, four));
        const __m128i C = _mm_and_si128(code, seven);
        const __m128i S = _mm_and_si128(_mm_and_si128(hi, three), door);
        uint8_t* o = out + 48 * g;
        for (int v = 0; v < 3; ++v) {
            const __m128i x = _mm_or_si128(_mm_or_si128(_mm_shuffle_epi8(T, m.t[v]), _mm_shuffle_epi8(C, m.c[v])),
                                           _mm_shuffle_epi8(S, m.s[v]));
            _mm_storeu_si128(reinterpret_cast<__m128i*>(o + 16 * v), x);
        }
    }
    const uint32_t c = rec[48];
    const uint32_t t = c >= 128 ? 4u : (c >> 3), s = c >= 128 ? ((c >> 3) & 3u) : 0u;
    out[144] = (uint8_t)t; out[145] = (uint8_t)(c & 7u); out[146] = (uint8_t)s;
    if (pad148) out[147] = 0;
}

// the same record as three planes (CHW: 49 types, 49 colours, 49 states), 147 bytes
extern "C" void mgrl_wire_expand_chw_ssse3(const uint8_t* rec, uint8_t* out) {
    const __m128i seven = _mm_set1_epi8(7), three = _mm_set1_epi8(3), four = _mm_set1_epi8(4), lo5 = _mm_set1_epi8(0x1F);
    for (int g = 0; g < 3; ++g) {
        const __m128i code = _mm_loadu_si128(reinterpret_cast<const __m128i*>(rec + 16 * g));
        const __m128i door = _mm_cmpgt_epi8(_mm_setzero_si128(), code);
        const __m128i hi = _mm_and_si128(_mm_srli_epi16(code, 3), lo5);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(out + 16 * g), _mm_or_si128(_mm_andnot_si128(door, hi), _mm_and_si128(door, four)));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(out + 49 + 16 * g), _mm_and_si128(code, seven));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(out + 98 + 16 * g), _mm_and_si128(_mm_and_si128(hi, three), door));
    }
    const uint32_t c = rec[48];
    out[48] = (uint8_t)(c >= 128 ? 4u : (c >> 3)); out[97] = (uint8_t)(c & 7u); out[146] = (uint8_t)(c >= 128 ? ((c >> 3) & 3u) : 0u);
}

// A block of `count` consecutive records (rec pitch 64) -> `count` consecutive HWC observation records of `pitch` bytes at
// `out`, written with non-temporal 16-byte stores: the images are write-only here, and ordinary stores would first READ every
// destination line into the cache (the expansion is bound by host memory traffic, not by the shuffles).  `out` must be 16-byte
// aligned; bytes are carried across records so that every store is an aligned unit; the tail (< 16 bytes) uses memcpy.
// `ready(i)` is called before record i is read and returns 0 to stop (the caller's "has this record landed" poll).
// `poll_ns` (optional) accumulates the time spent waiting for records that had not landed yet.
extern "C" int mgrl_wire_expand_block_hwc_ssse3(const uint8_t* recs, int count, uint8_t* out, int pitch, uint8_t tag, int tag_offset,
                                                const volatile int* abort_flag, unsigned long long* poll_ns) {
    alignas(16) uint8_t tmp[16 + 160];
    int carry = 0;                       // bytes already in tmp (< 16)
    uint8_t* dst = out;                  // next aligned unit
    for (int r = 0; r < count; ++r) {
        const uint8_t* rec = recs + (size_t)r * 64;
        const volatile uint8_t* vt = rec + tag_offset;
        if (*vt != tag) {
            const auto t0 = std::chrono::steady_clock::now();
            while (*vt != tag) {
                if (*abort_flag) return r;
                _mm_pause();
            }
            if (poll_ns) *poll_ns += (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        mgrl_wire_expand_hwc_ssse3(rec, tmp + carry, pitch == 148);
        const int have = carry + pitch, units = have >> 4;
        for (int u = 0; u < units; ++u)
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + u, _mm_load_si128(reinterpret_cast<const __m128i*>(tmp) + u));
        dst += units * 16;
        carry = have & 15;
        if (carry) memcpy(tmp, tmp + units * 16, 16);       // the partial unit moves to the front
    }
    if (carry) memcpy(dst, tmp, carry);
    _mm_sfence();
    return count;
}

// ---- AVX-512 VBMI: one 64-byte record is one register; a full byte permute (vpermb) per output vector and component replaces
// the nine 16-byte shuffles per 16 cells above.  Picked at run time when the CPU has AVX-512 BW + VBMI.
#include <immintrin.h>

extern "C" int mgrl_wire_have_avx512vbmi(void) {
    __builtin_cpu_init();
    return (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi")) ? 1 : 0;
}

namespace {
struct Vbmi {
    alignas(64) uint8_t idx[3][64];    // output byte 64 v + b <- cell (64 v + b) / 3 (63 = a zero lane past cell 48)
    uint64_t mc[3], ms[3];             // bytes of output vector v that hold a colour / a state
};
inline Vbmi make_vbmi() {
    Vbmi t;
    for (int v = 0; v < 3; ++v) {
        t.mc[v] = t.ms[v] = 0;
        for (int b = 0; b < 64; ++b) {
            const int j = 64 * v + b;
            t.idx[v][b] = (uint8_t)(j < 147 ? j / 3 : 63);
            if (j < 147 && j % 3 == 1) t.mc[v] |= 1ull << b;
            if (j < 147 && j % 3 == 2) t.ms[v] |= 1ull << b;
        }
    }
    return t;
}
}  // namespace

// same contract as mgrl_wire_expand_block_hwc_ssse3; `out` must be 64-byte aligned (non-temporal 64-byte stores)
extern "C" __attribute__((target("avx512f,avx512bw,avx512vbmi"))) int mgrl_wire_expand_block_hwc_avx512(
    const uint8_t* recs, int count, uint8_t* out, int pitch, uint8_t tag, int tag_offset, const volatile int* abort_flag,
    unsigned long long* poll_ns) {
    static const Vbmi tab = make_vbmi();
    const __m512i i0 = _mm512_load_si512(tab.idx[0]), i1 = _mm512_load_si512(tab.idx[1]), i2 = _mm512_load_si512(tab.idx[2]);
    const __m512i seven = _mm512_set1_epi8(7), three = _mm512_set1_epi8(3), four = _mm512_set1_epi8(4), lo5 = _mm512_set1_epi8(0x1F);
    const __mmask64 cells = (1ull << 49) - 1;                       // bytes 49..63 of a record are not cells
    alignas(64) uint8_t tmp[64 + 192];
    int carry = 0;                       // bytes already in tmp (< 64)
    uint8_t* dst = out;                  // next aligned unit
    for (int r = 0; r < count; ++r) {
        const uint8_t* rec = recs + (size_t)r * 64;
        const volatile uint8_t* vt = rec + tag_offset;
        if (*vt != tag) {
            const auto t0 = std::chrono::steady_clock::now();
            while (*vt != tag) {
                if (*abort_flag) return r;
                _mm_pause();
            }
            if (poll_ns) *poll_ns += (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        const __m512i code = _mm512_maskz_loadu_epi8(cells, rec);
        const __mmask64 door = _mm512_movepi8_mask(code);                                  // code >= 128
        const __m512i hi = _mm512_and_si512(_mm512_srli_epi16(code, 3), lo5);              // code >> 3 per byte
        const __m512i T = _mm512_mask_blend_epi8(door, hi, four);
        const __m512i C = _mm512_and_si512(code, seven);
        const __m512i S = _mm512_maskz_and_epi32(0xFFFF, _mm512_maskz_mov_epi8(door, hi), three);
        __m512i x0 = _mm512_permutexvar_epi8(i0, T), x1 = _mm512_permutexvar_epi8(i1, T), x2 = _mm512_permutexvar_epi8(i2, T);
        x0 = _mm512_mask_permutexvar_epi8(x0, tab.mc[0], i0, C); x0 = _mm512_mask_permutexvar_epi8(x0, tab.ms[0], i0, S);
        x1 = _mm512_mask_permutexvar_epi8(x1, tab.mc[1], i1, C); x1 = _mm512_mask_permutexvar_epi8(x1, tab.ms[1], i1, S);
        x2 = _mm512_mask_permutexvar_epi8(x2, tab.mc[2], i2, C); x2 = _mm512_mask_permutexvar_epi8(x2, tab.ms[2], i2, S);
        _mm512_storeu_si512(tmp + carry, x0); _mm512_storeu_si512(tmp + carry + 64, x1); _mm512_storeu_si512(tmp + carry + 128, x2);
        const int have = carry + pitch, units = have >> 6;
        for (int u = 0; u < units; ++u)
            _mm512_stream_si512(reinterpret_cast<__m512i*>(dst) + u, _mm512_load_si512(reinterpret_cast<const __m512i*>(tmp) + u));
        dst += units * 64;
        carry = have & 63;
        if (carry) _mm512_store_si512(tmp, _mm512_load_si512(tmp + units * 64));       // the partial unit moves to the front
    }
    if (carry) memcpy(dst, tmp, carry);
    _mm_sfence();
    return count;
}

// ---- AVX-512 VBMI, groups of 16 records of the 148-byte pitch: 16 x 148 B = 37 x 64 B, so a group starts and ends on an
// aligned unit and record r of a group starts (5 r mod 16) dwords into a unit - a compile-time constant.  The output
// vectors are realigned in REGISTERS (valignd against the previous vector) instead of through a staging buffer: the staging
// buffer's unaligned stores followed by aligned loads defeated store-to-load forwarding on every record (13.6 ns per
// record; the shuffles themselves are ~4 ns).  The step's scalars (bytes 48..59 of a record) are gathered as three dwords
// per record, 16 records at a time, and stored as 16-byte (64-byte for the rewards) units.
namespace {
#define MGRL_AVX512 __attribute__((target("avx512f,avx512bw,avx512vbmi"), always_inline)) inline
template <int S>
MGRL_AVX512 __m512i join_units(__m512i pending, __m512i x) {      // the top S dwords of `pending`, then the low 16 - S of x
    if constexpr (S == 0) return x;
    else return _mm512_alignr_epi32(x, pending, 16 - S);
}
struct Expand16 {
    __m512i i0, i1, i2, seven, three, four, lo5;
    __mmask64 mc0, mc1, mc2, ms0, ms1, ms2;
};
template <int R>
MGRL_AVX512 void expand_one(const Expand16& k, const uint8_t* rec, __m512i& pending, __m512i*& dst) {
    constexpr int S = (5 * R) % 16;                                // dwords of this group's output already pending
    const __mmask64 cells = (1ull << 49) - 1;
    const __m512i code = _mm512_maskz_loadu_epi8(cells, rec);
    const __mmask64 door = _mm512_movepi8_mask(code);
    const __m512i hi = _mm512_and_si512(_mm512_srli_epi16(code, 3), k.lo5);
    const __m512i T = _mm512_mask_blend_epi8(door, hi, k.four);
    const __m512i C = _mm512_and_si512(code, k.seven);
    const __m512i Sx = _mm512_and_si512(_mm512_maskz_mov_epi8(door, hi), k.three);
    __m512i x0 = _mm512_permutexvar_epi8(k.i0, T), x1 = _mm512_permutexvar_epi8(k.i1, T), x2 = _mm512_permutexvar_epi8(k.i2, T);
    x0 = _mm512_mask_permutexvar_epi8(x0, k.mc0, k.i0, C); x0 = _mm512_mask_permutexvar_epi8(x0, k.ms0, k.i0, Sx);
    x1 = _mm512_mask_permutexvar_epi8(x1, k.mc1, k.i1, C); x1 = _mm512_mask_permutexvar_epi8(x1, k.ms1, k.i1, Sx);
    x2 = _mm512_mask_permutexvar_epi8(x2, k.mc2, k.i2, C); x2 = _mm512_mask_permutexvar_epi8(x2, k.ms2, k.i2, Sx);
    _mm512_stream_si512(dst++, join_units<S>(pending, x0));
    _mm512_stream_si512(dst++, join_units<S>(x0, x1));
    if constexpr (S + 5 >= 16) {                                   // the record's last 5 dwords complete a unit
        _mm512_stream_si512(dst++, join_units<S>(x1, x2));
        pending = _mm512_alignr_epi32(x2, x2, 5);                  // x2's low 5 dwords on top; the top S + 5 - 16 are pending
    } else {
        pending = _mm512_alignr_epi32(x2, x1, 5);                  // x1's top S dwords, then x2's low 5: S + 5 pending
    }
}
}  // namespace

// `count` records -> floor(count / 16) * 16 observation records of 148 bytes at `out` (64-byte aligned) plus their scalars
// (any of the scalar arrays may be null); returns the number of records done (a multiple of 16, or less when aborted: then
// also a multiple of 16).  The caller finishes a ragged tail with the one-record routines.
extern "C" __attribute__((target("avx512f,avx512bw,avx512vbmi"))) int mgrl_wire_expand_groups_hwc148_avx512(
    const uint8_t* recs, int count, uint8_t* out, uint8_t tag, int tag_offset, const volatile int* abort_flag,
    unsigned long long* poll_ns, uint8_t* dir, uint8_t* mission, uint8_t* term, uint8_t* trunc, uint8_t* eplen, uint8_t* tdir, float* reward) {
    static const Vbmi tab = make_vbmi();
    Expand16 k;
    k.i0 = _mm512_load_si512(tab.idx[0]); k.i1 = _mm512_load_si512(tab.idx[1]); k.i2 = _mm512_load_si512(tab.idx[2]);
    k.seven = _mm512_set1_epi8(7); k.three = _mm512_set1_epi8(3); k.four = _mm512_set1_epi8(4); k.lo5 = _mm512_set1_epi8(0x1F);
    k.mc0 = tab.mc[0]; k.mc1 = tab.mc[1]; k.mc2 = tab.mc[2]; k.ms0 = tab.ms[0]; k.ms1 = tab.ms[1]; k.ms2 = tab.ms[2];
    const __m512i rec_off = _mm512_setr_epi32(0, 64, 128, 192, 256, 320, 384, 448, 512, 576, 640, 704, 768, 832, 896, 960);
    __m512i* dst = reinterpret_cast<__m512i*>(out);
    int done = 0;
    for (; done + 16 <= count; done += 16) {
        const uint8_t* g = recs + (size_t)done * 64;
        for (int r = 15; r >= 0; --r) {                            // the copy lands in address order: the last record first
            const volatile uint8_t* vt = g + r * 64 + tag_offset;
            if (*vt != tag) {
                const auto t0 = std::chrono::steady_clock::now();
                while (*vt != tag) {
                    if (*abort_flag) { _mm_sfence(); return done; }
                    _mm_pause();
                }
                if (poll_ns) *poll_ns += (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
            }
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        __m512i pending = _mm512_setzero_si512();
        expand_one<0>(k, g, pending, dst); expand_one<1>(k, g + 64, pending, dst); expand_one<2>(k, g + 128, pending, dst);
        expand_one<3>(k, g + 192, pending, dst); expand_one<4>(k, g + 256, pending, dst); expand_one<5>(k, g + 320, pending, dst);
        expand_one<6>(k, g + 384, pending, dst); expand_one<7>(k, g + 448, pending, dst); expand_one<8>(k, g + 512, pending, dst);
        expand_one<9>(k, g + 576, pending, dst); expand_one<10>(k, g + 640, pending, dst); expand_one<11>(k, g + 704, pending, dst);
        expand_one<12>(k, g + 768, pending, dst); expand_one<13>(k, g + 832, pending, dst); expand_one<14>(k, g + 896, pending, dst);
        expand_one<15>(k, g + 960, pending, dst);
        // scalars: dword 12 = cell 48 | dir | mission | terminated, dword 13 = truncated | length | terminal dir | tag, dword 14 = reward
        const __m512i d12 = _mm512_i32gather_epi32(rec_off, g + 48, 1), d13 = _mm512_i32gather_epi32(rec_off, g + 52, 1);
        if (dir) _mm_storeu_si128(reinterpret_cast<__m128i*>(dir + done), _mm512_cvtepi32_epi8(_mm512_srli_epi32(d12, 8)));
        if (mission) _mm_storeu_si128(reinterpret_cast<__m128i*>(mission + done), _mm512_cvtepi32_epi8(_mm512_srli_epi32(d12, 16)));
        if (term) _mm_storeu_si128(reinterpret_cast<__m128i*>(term + done), _mm512_cvtepi32_epi8(_mm512_srli_epi32(d12, 24)));
        if (trunc) _mm_storeu_si128(reinterpret_cast<__m128i*>(trunc + done), _mm512_cvtepi32_epi8(d13));
        if (eplen) _mm_storeu_si128(reinterpret_cast<__m128i*>(eplen + done), _mm512_cvtepi32_epi8(_mm512_srli_epi32(d13, 8)));
        if (tdir) _mm_storeu_si128(reinterpret_cast<__m128i*>(tdir + done), _mm512_cvtepi32_epi8(_mm512_srli_epi32(d13, 16)));
        if (reward) _mm512_storeu_si512(reward + done, _mm512_i32gather_epi32(rec_off, g + 56, 1));
    }
    _mm_sfence();
    return done;
}
