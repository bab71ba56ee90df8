// fjsp_wire.cpp — HOST decode of wire rows (include/fjsp_b200.h) into the float32 / int8 tensors fjsp_step writes.
// Format conversion only.  Two bodies per K: portable C++ and AVX2 (chosen once at run time with
// __builtin_cpu_supports); both produce the same bits (integer -> float32 conversions are exact, one IEEE division).
#if defined(__x86_64__) || defined(_M_X64)
#define FJSP_WIRE_X86 1
#include <immintrin.h>
#endif
#include <string.h>

#include "fjsp_wire.h"

namespace fjsp {

// eight mask bits -> eight bytes (0/1)
static inline uint64_t spread8(uint32_t x) {
    const uint64_t y = ((uint64_t)(x & 0xffu) * 0x0101010101010101ULL) & 0x8040201008040201ULL;
    return ((y + 0x7f7f7f7f7f7f7f7fULL) >> 7) & 0x0101010101010101ULL;
}

struct PosTab {
    float row[8], col[8];
    explicit PosTab(const Params& P) {
        for (int i = 0; i < 8; i++) row[i] = (float)P.pos_row[i < FJSP_NUM_LOCATIONS ? i : 0], col[i] = (float)P.pos_col[i < FJSP_NUM_LOCATIONS ? i : 0];
    }
};

// the ten fields per cell that are not "the byte as a float": AGV row / column, four progress values, four int8 queue lengths
template <int K>
static inline void obs_fixups(const Params& P, const PosTab& T, const uint8_t* b, float* o) {
    for (int c = 0; c < K; c++) {
        const int base = 7 + 31 * c;
        o[base + 4] = T.row[b[base + 4] & 7], o[base + 5] = T.col[b[base + 5] & 7];
        for (int i = 0; i < 4; i++) {
            o[base + 20 + 3 * i] = P.progress_tab[b[base + 20 + 3 * i]];
            o[base + 21 + 3 * i] = (float)(int)(int8_t)b[base + 21 + 3 * i];
        }
    }
}

template <int K>
static void decode_generic(const Params& P, const u32* wire, int64_t lo, int64_t hi, float* obs, int8_t* masks, float* rewards,
                           uint8_t* flags) {
    constexpr int OBS = Lay<K>::OBS, MASK = Lay<K>::MASK, ACT = Lay<K>::ACT, A = Lay<K>::AGENTS, WORDS = Wire<K>::WORDS;
    const PosTab T(P);
    const float denom = (float)(10 * A);
    for (int64_t e = lo; e < hi; e++) {
        const u32* row = wire + e * WORDS;
        if (obs) {
            const uint8_t* b = reinterpret_cast<const uint8_t*>(row);
            float* o = obs + e * OBS;
            for (int i = 0; i < OBS; i++) o[i] = (float)b[i];
            obs_fixups<K>(P, T, b, o);
        }
        if (masks) {
            for (int w = 0; w < Wire<K>::MW; w++) {
                const u32 bits = row[Wire<K>::OFF_MASK + w];
                const uint64_t v[4] = {spread8(bits), spread8(bits >> 8), spread8(bits >> 16), spread8(bits >> 24)};
                memcpy(masks + e * MASK + 32 * w, v, 32);
            }
        }
        if (rewards) {
            const int g = wire_g(row[Wire<K>::OFF_G]);
            const int16_t* l = reinterpret_cast<const int16_t*>(row + Wire<K>::OFF_LOCAL);
            float* r = rewards + e * ACT;
            for (int i = 0; i < ACT; i++) r[i] = i < A ? (float)(g + A * (int)l[i]) / denom : 0.0f;
        }
        if (flags) {
            const u32 f = wire_flags(row[Wire<K>::OFF_G]);
            memcpy(flags + e * FJSP_FLAG_DIM, &f, 4);
        }
    }
}

#ifdef FJSP_WIRE_X86
// One env, AVX2.  NT = false: straight into the caller's rows.  NT = true (used on blocks of 8 envs): obs into a small
// aligned stack buffer, masks / rewards as register values handed back to the caller for streaming stores.
template <int K>
struct Avx2Consts {
    __m256 denom;
    __m256i vA, obs_tail, rew_tail, sel, bit, one;
};
template <int K>
__attribute__((target("avx2"))) static inline Avx2Consts<K> avx2_consts() {
    constexpr int OBS = Lay<K>::OBS, A = Lay<K>::AGENTS;
    alignas(32) static const int32_t lane_lt[16] = {-1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0};
    Avx2Consts<K> c;
    c.denom = _mm256_set1_ps((float)(10 * A));
    c.vA = _mm256_set1_epi32(A);
    c.obs_tail = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(lane_lt + 8 - (OBS & 7)));  // lanes 0..(OBS%8 - 1)
    c.rew_tail = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(lane_lt + 8 - (A & 7)));    // A % 8 live columns
    c.sel = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3);
    c.bit = _mm256_set1_epi64x((long long)0x8040201008040201ULL);
    c.one = _mm256_set1_epi8(1);
    return c;
}
template <int K>
__attribute__((target("avx2"))) static inline void obs_avx2(const Params& P, const PosTab& T, const Avx2Consts<K>& c, const u32* row,
                                                            float* o) {
    constexpr int OBS = Lay<K>::OBS;
    const uint8_t* b = reinterpret_cast<const uint8_t*>(row);
#pragma GCC unroll 8
    for (int i = 0; i < OBS / 8; i++) {
        const __m256i v = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(b + 8 * i)));
        _mm256_storeu_ps(o + 8 * i, _mm256_cvtepi32_ps(v));
    }
    if (OBS & 7) {  // the row's obs bytes are padded to a multiple of 4 and followed by the mask words: 8 readable bytes
        const __m256i v = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(b + (OBS & ~7))));
        _mm256_maskstore_ps(o + (OBS & ~7), c.obs_tail, _mm256_cvtepi32_ps(v));
    }
    obs_fixups<K>(P, T, b, o);
}
template <int K>
__attribute__((target("avx2"))) static inline __m256i mask_avx2(const Avx2Consts<K>& c, u32 bits) {
    const __m256i v = _mm256_shuffle_epi8(_mm256_set1_epi32((int)bits), c.sel);
    return _mm256_and_si256(_mm256_cmpeq_epi8(_mm256_and_si256(v, c.bit), c.bit), c.one);
}
template <int K>
__attribute__((target("avx2"))) static inline __m256 reward_avx2(const Avx2Consts<K>& c, const u32* row, int i) {
    const __m256i g = _mm256_set1_epi32(wire_g(row[Wire<K>::OFF_G]));
    const int16_t* l = reinterpret_cast<const int16_t*>(row + Wire<K>::OFF_LOCAL);
    const __m256i li = _mm256_cvtepi16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(l + 8 * i)));
    __m256 q = _mm256_div_ps(_mm256_cvtepi32_ps(_mm256_add_epi32(g, _mm256_mullo_epi32(li, c.vA))), c.denom);
    if (8 * i + 8 > Lay<K>::AGENTS) q = _mm256_and_ps(q, _mm256_castsi256_ps(c.rew_tail));  // padding columns are 0.0f
    return q;
}

static bool nt_stores() {
    static const bool v = !getenv("FJSP_DECODE_NO_NT");
    return v;
}

template <int K>
__attribute__((target("avx2"))) static void decode_avx2(const Params& P, const u32* wire, int64_t lo, int64_t hi, float* obs,
                                                        int8_t* masks, float* rewards, uint8_t* flags) {
    constexpr int OBS = Lay<K>::OBS, MASK = Lay<K>::MASK, ACT = Lay<K>::ACT, WORDS = Wire<K>::WORDS, MW = Wire<K>::MW;
    const PosTab T(P);
    const Avx2Consts<K> c = avx2_consts<K>();
    int64_t e = lo;
    // Blocks of 8 envs whose output rows start on 32-byte boundaries: written with streaming (non-temporal) stores, so
    // the 220 B per env of results do not cost a read-for-ownership of every destination line on top of the write.
    // 8 rows of any of the four tensors are a whole number of 32-byte vectors.
    const bool aligned = (!obs || (reinterpret_cast<uintptr_t>(obs) & 31) == 0) && (!masks || (reinterpret_cast<uintptr_t>(masks) & 31) == 0) &&
                         (!rewards || (reinterpret_cast<uintptr_t>(rewards) & 31) == 0) && (!flags || (reinterpret_cast<uintptr_t>(flags) & 31) == 0);
    if (aligned && nt_stores()) {
        for (; e < hi && (e & 7); e++) {  // head: up to the first multiple of 8
            const u32* row = wire + e * WORDS;
            if (obs) obs_avx2<K>(P, T, c, row, obs + e * OBS);
            if (masks)
                for (int w = 0; w < MW; w++) _mm256_storeu_si256(reinterpret_cast<__m256i*>(masks + e * MASK + 32 * w), mask_avx2<K>(c, row[Wire<K>::OFF_MASK + w]));
            if (rewards)
                for (int i = 0; i < ACT / 8; i++) _mm256_storeu_ps(rewards + e * ACT + 8 * i, reward_avx2<K>(c, row, i));
            if (flags) {
            const u32 f = wire_flags(row[Wire<K>::OFF_G]);
            memcpy(flags + e * FJSP_FLAG_DIM, &f, 4);
        }
        }
        alignas(32) float tmp[8 * OBS];
        for (; e + 8 <= hi; e += 8) {
            const u32* row0 = wire + e * WORDS;
            if (obs) {
                for (int j = 0; j < 8; j++) obs_avx2<K>(P, T, c, row0 + j * WORDS, tmp + j * OBS);
                float* dst = obs + e * OBS;
#pragma GCC unroll 8
                for (int i = 0; i < OBS; i++) _mm256_stream_ps(dst + 8 * i, _mm256_load_ps(tmp + 8 * i));
            }
            if (masks)
                for (int j = 0; j < 8; j++)
                    for (int w = 0; w < MW; w++)
                        _mm256_stream_si256(reinterpret_cast<__m256i*>(masks + (e + j) * MASK + 32 * w), mask_avx2<K>(c, row0[j * WORDS + Wire<K>::OFF_MASK + w]));
            if (rewards)
                for (int j = 0; j < 8; j++)
                    for (int i = 0; i < ACT / 8; i++) _mm256_stream_ps(rewards + (e + j) * ACT + 8 * i, reward_avx2<K>(c, row0 + j * WORDS, i));
            if (flags) {
                const __m256i f = _mm256_setr_epi32((int)wire_flags(row0[0 * WORDS + Wire<K>::OFF_G]), (int)wire_flags(row0[1 * WORDS + Wire<K>::OFF_G]),
                                                    (int)wire_flags(row0[2 * WORDS + Wire<K>::OFF_G]), (int)wire_flags(row0[3 * WORDS + Wire<K>::OFF_G]),
                                                    (int)wire_flags(row0[4 * WORDS + Wire<K>::OFF_G]), (int)wire_flags(row0[5 * WORDS + Wire<K>::OFF_G]),
                                                    (int)wire_flags(row0[6 * WORDS + Wire<K>::OFF_G]), (int)wire_flags(row0[7 * WORDS + Wire<K>::OFF_G]));
                _mm256_stream_si256(reinterpret_cast<__m256i*>(flags + e * FJSP_FLAG_DIM), f);
            }
        }
        _mm_sfence();
    }
    for (; e < hi; e++) {  // tail, or everything when the caller's buffers are not 32-byte aligned
        const u32* row = wire + e * WORDS;
        if (obs) obs_avx2<K>(P, T, c, row, obs + e * OBS);
        if (masks)
            for (int w = 0; w < MW; w++) _mm256_storeu_si256(reinterpret_cast<__m256i*>(masks + e * MASK + 32 * w), mask_avx2<K>(c, row[Wire<K>::OFF_MASK + w]));
        if (rewards)
            for (int i = 0; i < ACT / 8; i++) _mm256_storeu_ps(rewards + e * ACT + 8 * i, reward_avx2<K>(c, row, i));
        if (flags) {
            const u32 f = wire_flags(row[Wire<K>::OFF_G]);
            memcpy(flags + e * FJSP_FLAG_DIM, &f, 4);
        }
    }
}

#endif  // FJSP_WIRE_X86

typedef void (*DecodeFn)(const Params&, const u32*, int64_t, int64_t, float*, int8_t*, float*, uint8_t*);

static bool have_avx2() {
#ifdef FJSP_WIRE_X86
    static const bool v = __builtin_cpu_supports("avx2");
    return v && !getenv("FJSP_DECODE_GENERIC");
#else
    return false;
#endif
}
const char* wire_decode_isa() { return have_avx2() ? "avx2" : "generic"; }

void wire_decode(int cells, const Params& P, const u32* wire, int64_t lo, int64_t hi, float* obs, int8_t* masks, float* rewards,
                 uint8_t* flags) {
    static const DecodeFn gen[4] = {decode_generic<1>, decode_generic<2>, decode_generic<3>, decode_generic<4>};
    const int k = cells < 1 ? 0 : cells > 4 ? 3 : cells - 1;
#ifdef FJSP_WIRE_X86
    static const DecodeFn avx[4] = {decode_avx2<1>, decode_avx2<2>, decode_avx2<3>, decode_avx2<4>};
    if (have_avx2()) {
        avx[k](P, wire, lo, hi, obs, masks, rewards, flags);
        return;
    }
#endif
    gen[k](P, wire, lo, hi, obs, masks, rewards, flags);
}

}  // namespace fjsp
