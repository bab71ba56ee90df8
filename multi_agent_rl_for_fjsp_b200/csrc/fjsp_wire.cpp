// fjsp_wire.cpp — HOST decode of wire rows (include/fjsp_b200.h) into the float32 / int8 tensors fjsp_step writes.
// Format conversion only.  Two bodies per K: portable C++ and AVX2 (chosen once at run time with
// __builtin_cpu_supports); both produce the same bits (integer -> float32 conversions are exact, one IEEE division).
#include <immintrin.h>
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
            const int g = (int)row[Wire<K>::OFF_G];
            const int16_t* l = reinterpret_cast<const int16_t*>(row + Wire<K>::OFF_LOCAL);
            float* r = rewards + e * ACT;
            for (int i = 0; i < ACT; i++) r[i] = i < A ? (float)(g + A * (int)l[i]) / denom : 0.0f;
        }
        if (flags) memcpy(flags + e * FJSP_FLAG_DIM, row + Wire<K>::OFF_FLAGS, 4);
    }
}

template <int K>
__attribute__((target("avx2"))) static void decode_avx2(const Params& P, const u32* wire, int64_t lo, int64_t hi, float* obs,
                                                        int8_t* masks, float* rewards, uint8_t* flags) {
    constexpr int OBS = Lay<K>::OBS, MASK = Lay<K>::MASK, ACT = Lay<K>::ACT, A = Lay<K>::AGENTS, WORDS = Wire<K>::WORDS;
    const PosTab T(P);
    const __m256 denom = _mm256_set1_ps((float)(10 * A));
    const __m256i vA = _mm256_set1_epi32(A);
    // lanes 0..(n-1) of a group of 8
    alignas(32) static const int32_t lane_lt[16] = {-1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0};
    const __m256i obs_tail = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(lane_lt + 8 - (OBS & 7)));
    const __m256i rew_tail = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(lane_lt + 8 - (A & 7)));  // A % 8 live columns
    const __m256i sel = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3);
    const __m256i bit = _mm256_set1_epi64x((long long)0x8040201008040201ULL);
    const __m256i one = _mm256_set1_epi8(1);
    for (int64_t e = lo; e < hi; e++) {
        const u32* row = wire + e * WORDS;
        if (obs) {
            const uint8_t* b = reinterpret_cast<const uint8_t*>(row);
            float* o = obs + e * OBS;
#pragma GCC unroll 8
            for (int i = 0; i < OBS / 8; i++) {
                const __m256i v = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(b + 8 * i)));
                _mm256_storeu_ps(o + 8 * i, _mm256_cvtepi32_ps(v));
            }
            if (OBS & 7) {  // the row's obs bytes are padded to a multiple of 4 and followed by the mask words: 8 readable bytes
                const __m256i v = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(b + (OBS & ~7))));
                _mm256_maskstore_ps(o + (OBS & ~7), obs_tail, _mm256_cvtepi32_ps(v));
            }
            obs_fixups<K>(P, T, b, o);
        }
        if (masks) {
            for (int w = 0; w < Wire<K>::MW; w++) {
                __m256i v = _mm256_shuffle_epi8(_mm256_set1_epi32((int)row[Wire<K>::OFF_MASK + w]), sel);
                v = _mm256_and_si256(_mm256_cmpeq_epi8(_mm256_and_si256(v, bit), bit), one);
                _mm256_storeu_si256(reinterpret_cast<__m256i*>(masks + e * MASK + 32 * w), v);
            }
        }
        if (rewards) {
            const __m256i g = _mm256_set1_epi32((int)row[Wire<K>::OFF_G]);
            const int16_t* l = reinterpret_cast<const int16_t*>(row + Wire<K>::OFF_LOCAL);
            float* r = rewards + e * ACT;
#pragma GCC unroll 4
            for (int i = 0; i < ACT / 8; i++) {
                const __m256i li = _mm256_cvtepi16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(l + 8 * i)));
                __m256 q = _mm256_div_ps(_mm256_cvtepi32_ps(_mm256_add_epi32(g, _mm256_mullo_epi32(li, vA))), denom);
                if (8 * i + 8 > A) q = _mm256_and_ps(q, _mm256_castsi256_ps(rew_tail));  // padding columns are 0.0f
                _mm256_storeu_ps(r + 8 * i, q);
            }
        }
        if (flags) memcpy(flags + e * FJSP_FLAG_DIM, row + Wire<K>::OFF_FLAGS, 4);
    }
}

typedef void (*DecodeFn)(const Params&, const u32*, int64_t, int64_t, float*, int8_t*, float*, uint8_t*);

static bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v && !getenv("FJSP_DECODE_GENERIC");
}
const char* wire_decode_isa() { return have_avx2() ? "avx2" : "generic"; }

void wire_decode(int cells, const Params& P, const u32* wire, int64_t lo, int64_t hi, float* obs, int8_t* masks, float* rewards,
                 uint8_t* flags) {
    static const DecodeFn gen[4] = {decode_generic<1>, decode_generic<2>, decode_generic<3>, decode_generic<4>};
    static const DecodeFn avx[4] = {decode_avx2<1>, decode_avx2<2>, decode_avx2<3>, decode_avx2<4>};
    const int k = cells < 1 ? 0 : cells > 4 ? 3 : cells - 1;
    (have_avx2() ? avx : gen)[k](P, wire, lo, hi, obs, masks, rewards, flags);
}

}  // namespace fjsp
