// fjsp_wire.cpp — HOST decode of wire rows (include/fjsp_b200.h) into the float32 / int8 tensors fjsp_step writes.
// Format conversion only: every output value is one bit-field of one word of the row (shift + mask), plus three tiny
// tables (station -> grid position, progress index -> float32(100 / L), reward code -> local reward) and the single
// IEEE division per reward that the kernels perform.  Two bodies: portable C++ (any K) and AVX2 for K = 1 (the
// reference shop, the host-buffer path bench.py times), chosen once at run time; both produce the same bits.
#if defined(__x86_64__) || defined(_M_X64)
#define FJSP_WIRE_X86 1
#include <immintrin.h>
#endif
#include <string.h>

#include "fjsp_wire.h"

#include <chrono>

namespace fjsp {

static const int32_t kRewardLut[32] = FJSP_WIRE_REWARD_LUT;

struct PosTab {
    float row[8], col[8];
    explicit PosTab(const Params& P) {
        for (int i = 0; i < 8; i++) row[i] = (float)P.pos_row[i < FJSP_NUM_LOCATIONS ? i : 0], col[i] = (float)P.pos_col[i < FJSP_NUM_LOCATIONS ? i : 0];
    }
};

static inline u32 bf(u32 w, int shift, u32 mask) { return (w >> shift) & mask; }

// the 32 (or 32 * n) mask bytes of a row as bits: bit 0 of every agent is the constant 1
template <int K>
static inline void mask_bits(const u32* row, u32* out /* Lay<K>::MASK / 32 words */) {
    uint64_t acc[(Lay<K>::MASK / 32 + 1) / 2 + 1] = {0};
    auto put = [&](int pos, uint64_t v) {
        acc[pos >> 6] |= v << (pos & 63);
        if ((pos & 63) > 38) acc[(pos >> 6) + 1] |= v >> (64 - (pos & 63));  // a cell's 26 bits may straddle a 64-bit word
    };
    put(0, 1u | (bf(row[0], 25, 3) << 1));
    for (int c = 0; c < K; c++) {
        const u32* w = row + 2 + 5 * c;
        uint64_t b = 1u | ((uint64_t)bf(w[0], 22, 0x7f) << 1);            // agv: 8
        b |= (uint64_t)(1u | (bf(w[1], 24, 3) << 1)) << 8;                 // small machine: 3
        b |= (uint64_t)(1u | (bf(w[1], 26, 3) << 1)) << 11;                // big machine: 3
        b |= (uint64_t)(1u | (bf(w[2], 17, 3) << 1)) << 14;                // packaging_blue_1
        b |= (uint64_t)(1u | (bf(w[3], 17, 3) << 1)) << 17;                // packaging_blue_2
        b |= (uint64_t)(1u | (bf(w[4], 17, 3) << 1)) << 20;                // packaging_red
        b |= (uint64_t)(1u | (bf(w[2], 30, 3) << 1)) << 23;                // packaging_green
        put(3 + 26 * c, b);
    }
    for (int i = 0; i < Lay<K>::MASK / 32; i++) out[i] = (u32)(acc[i >> 1] >> ((i & 1) * 32));
}

// eight mask bits -> eight bytes (0/1)
static inline uint64_t spread8(u32 x) {
    const uint64_t y = ((uint64_t)(x & 0xffu) * 0x0101010101010101ULL) & 0x8040201008040201ULL;
    return ((y + 0x7f7f7f7f7f7f7f7fULL) >> 7) & 0x0101010101010101ULL;
}

template <int K>
static inline void obs_row(const Params& P, const PosTab& T, const u32* row, float* o) {
    const u32 s0 = row[0];
    o[0] = (float)bf(s0, 0, 3), o[1] = (float)bf(s0, 2, 7), o[2] = (float)bf(s0, 5, 3), o[3] = (float)bf(s0, 7, 3);
    o[4] = (float)bf(s0, 9, 3), o[5] = (float)bf(s0, 11, 15), o[6] = (float)bf(s0, 15, 15);
    const float ready = (float)(row[1] >> 19);
    for (int c = 0; c < K; c++) {
        const u32* w = row + 2 + 5 * c;
        float* q = o + 7 + 31 * c;
        const u32 loc = bf(w[0], 0, 7);
        q[0] = (float)bf(w[1], 16, 1), q[1] = (float)bf(w[0], 10, 63), q[2] = (float)bf(w[0], 3, 1), q[3] = ready;
        q[4] = T.row[loc], q[5] = T.col[loc];
        q[6] = (float)bf(w[1], 8, 1), q[7] = (float)bf(w[0], 16, 63), q[8] = (float)bf(w[1], 0, 255), q[9] = q[2];
        q[10] = (float)bf(w[0], 4, 1), q[11] = (float)bf(w[0], 5, 7), q[12] = (float)bf(w[0], 8, 3);
        q[13] = q[6], q[14] = (float)bf(w[1], 9, 1), q[15] = (float)bf(w[1], 10, 63);
        q[16] = q[0], q[17] = (float)bf(w[1], 17, 1), q[18] = (float)bf(w[1], 18, 63);
        for (int i = 0; i < 3; i++) {
            q[19 + 3 * i] = (float)bf(w[2 + i], 0, 1);
            q[20 + 3 * i] = P.progress_tab[bf(w[2 + i], 1, 255)];
            q[21 + 3 * i] = (float)(int)(int8_t)bf(w[2 + i], 9, 255);
        }
        q[28] = (float)bf(w[2], 29, 1), q[29] = P.progress_tab[bf(w[2], 21, 255)], q[30] = (float)(int)(int8_t)bf(w[3], 21, 255);
    }
}

template <int K>
static inline void reward_row(const Params& P, const u32* row, float* r) {
    constexpr int A = Lay<K>::AGENTS, ACT = Lay<K>::ACT;
    const int g = 10 * (100 * (int)bf(row[1], 10, 511) + 10 * (int)bf(row[1], 0, 1023)) - P.step_size;
    const float denom = (float)(10 * A);
    r[0] = (float)(g + A * kRewardLut[bf(row[0], 27, 3)]) / denom;
    for (int c = 0; c < K; c++) {
        const u32* w = row + 2 + 5 * c;
        float* q = r + 1 + 7 * c;
        q[0] = (float)(g + A * kRewardLut[8 + bf(w[0], 29, 7)]) / denom;
        q[1] = (float)(g + A * kRewardLut[16 + bf(w[1], 28, 3)]) / denom;
        q[2] = (float)(g + A * kRewardLut[16 + bf(w[1], 30, 3)]) / denom;
        q[3] = (float)(g + A * kRewardLut[24 + bf(w[2], 19, 3)]) / denom;
        q[4] = (float)(g + A * kRewardLut[24 + bf(w[3], 19, 3)]) / denom;
        q[5] = (float)(g + A * kRewardLut[24 + bf(w[4], 19, 3)]) / denom;
        q[6] = (float)(g + A * kRewardLut[24 + bf(w[3], 29, 3)]) / denom;
    }
    for (int i = A; i < ACT; i++) r[i] = 0.0f;
}

template <int K>
static void decode_generic(const Params& P, const u32* wire, int64_t lo, int64_t hi, float* obs, int8_t* masks, float* rewards,
                           uint8_t* flags) {
    constexpr int OBS = Lay<K>::OBS, MASK = Lay<K>::MASK, ACT = Lay<K>::ACT, WORDS = Wire<K>::WORDS;
    const PosTab T(P);
    for (int64_t e = lo; e < hi; e++) {
        const u32* row = wire + e * WORDS;
        if (obs) obs_row<K>(P, T, row, obs + e * OBS);
        if (masks) {
            u32 bits[MASK / 32];
            mask_bits<K>(row, bits);
            for (int w = 0; w < MASK / 32; w++) {
                const uint64_t v[4] = {spread8(bits[w]), spread8(bits[w] >> 8), spread8(bits[w] >> 16), spread8(bits[w] >> 24)};
                memcpy(masks + e * MASK + 32 * w, v, 32);
            }
        }
        if (rewards) reward_row<K>(P, row, rewards + e * ACT);
        if (flags) {
            const u32 f = wire_flags(row[0]);
            memcpy(flags + e * FJSP_FLAG_DIM, &f, 4);
        }
    }
}

#ifdef FJSP_WIRE_X86
// ---- AVX2, K = 1: the row is ONE 32-byte vector.  Each group of 8 consecutive output floats is
//      permutevar8x32 (pick every lane's source word) -> srlv (its shift) -> and (its mask) -> cvt; the int8 queue lengths
//      are sign-extended with xor / sub, the two position fields come from 8-entry tables with permutevar8x32_ps.
struct Field {
    int word, shift;
    u32 mask;
    int sign8;
};
// source bit-field of each of the 38 observation floats (2 padding lanes), K = 1: words S0 S1 C0 C1 C2 C3 C4 = 0..6
static const Field kObsField[40] = {
    {0, 0, 3, 0}, {0, 2, 7, 0}, {0, 5, 3, 0}, {0, 7, 3, 0}, {0, 9, 3, 0}, {0, 11, 15, 0}, {0, 15, 15, 0},       // pickup station
    {3, 16, 1, 0}, {2, 10, 63, 0}, {2, 3, 1, 0}, {1, 19, 8191, 0}, {2, 0, 7, 0}, {2, 0, 7, 0}, {3, 8, 1, 0},    // agv 0..6
    {2, 16, 63, 0}, {3, 0, 255, 0}, {2, 3, 1, 0}, {2, 4, 1, 0}, {2, 5, 7, 0}, {2, 8, 3, 0},                    // agv 7..12
    {3, 8, 1, 0}, {3, 9, 1, 0}, {3, 10, 63, 0}, {3, 16, 1, 0}, {3, 17, 1, 0}, {3, 18, 63, 0},                  // machines
    {4, 0, 1, 0}, {4, 1, 255, 0}, {4, 9, 255, 1}, {5, 0, 1, 0}, {5, 1, 255, 0}, {5, 9, 255, 1},                // blue_1, blue_2
    {6, 0, 1, 0}, {6, 1, 255, 0}, {6, 9, 255, 1}, {4, 29, 1, 0}, {4, 21, 255, 0}, {5, 21, 255, 1},            // red, green
    {7, 0, 0, 0}, {7, 0, 0, 0}};
// source of the 8 reward codes and the kind offset into kRewardLut
static const Field kRewField[8] = {{0, 27, 3, 0}, {2, 29, 7, 8}, {3, 28, 3, 16}, {3, 30, 3, 16}, {4, 19, 3, 24}, {5, 19, 3, 24},
                                   {6, 19, 3, 24}, {5, 29, 3, 24}};

struct Avx2K1 {
    __m256i idx[5], sh[5], msk[5], sgn[5];
    __m256i ridx, rsh, rmsk, roff;
    __m256 rowtab, coltab, denom;
    __m256i sel, bit, one;
};
__attribute__((target("avx2"))) static Avx2K1 make_avx2_k1(const PosTab& T) {
    Avx2K1 c;
    alignas(32) int32_t a[8], b[8], m[8], s[8];
    for (int g = 0; g < 5; g++) {
        for (int l = 0; l < 8; l++) {
            const Field& f = kObsField[8 * g + l];
            a[l] = f.word, b[l] = f.shift, m[l] = (int32_t)f.mask, s[l] = f.sign8 ? 0x80 : 0;
        }
        c.idx[g] = _mm256_load_si256((const __m256i*)a), c.sh[g] = _mm256_load_si256((const __m256i*)b);
        c.msk[g] = _mm256_load_si256((const __m256i*)m), c.sgn[g] = _mm256_load_si256((const __m256i*)s);
    }
    for (int l = 0; l < 8; l++) a[l] = kRewField[l].word, b[l] = kRewField[l].shift, m[l] = (int32_t)kRewField[l].mask, s[l] = kRewField[l].sign8;
    c.ridx = _mm256_load_si256((const __m256i*)a), c.rsh = _mm256_load_si256((const __m256i*)b);
    c.rmsk = _mm256_load_si256((const __m256i*)m), c.roff = _mm256_load_si256((const __m256i*)s);
    c.rowtab = _mm256_loadu_ps(T.row), c.coltab = _mm256_loadu_ps(T.col);
    c.denom = _mm256_set1_ps(80.0f);
    c.sel = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3);
    c.bit = _mm256_set1_epi64x((long long)0x8040201008040201ULL);
    c.one = _mm256_set1_epi8(1);
    return c;
}
// 40 floats at o (38 used; the caller provides room for the 2 padding lanes)
__attribute__((target("avx2"))) static inline void obs_avx2_k1(const Params& P, const Avx2K1& c, __m256i R, const u32* row, float* o) {
#pragma GCC unroll 5
    for (int g = 0; g < 5; g++) {
        __m256i x = _mm256_and_si256(_mm256_srlv_epi32(_mm256_permutevar8x32_epi32(R, c.idx[g]), c.sh[g]), c.msk[g]);
        x = _mm256_sub_epi32(_mm256_xor_si256(x, c.sgn[g]), c.sgn[g]);
        __m256 f = _mm256_cvtepi32_ps(x);
        if (g == 1) {  // floats 11, 12 = the AGV's grid row / column (lanes 3, 4)
            f = _mm256_blend_ps(f, _mm256_permutevar8x32_ps(c.rowtab, x), 1 << 3);
            f = _mm256_blend_ps(f, _mm256_permutevar8x32_ps(c.coltab, x), 1 << 4);
        }
        _mm256_storeu_ps(o + 8 * g, f);
    }
    o[27] = P.progress_tab[bf(row[4], 1, 255)], o[30] = P.progress_tab[bf(row[5], 1, 255)];
    o[33] = P.progress_tab[bf(row[6], 1, 255)], o[36] = P.progress_tab[bf(row[4], 21, 255)];
}
__attribute__((target("avx2"))) static inline __m256i mask_avx2_k1(const Avx2K1& c, const u32* row) {
    u32 bits;
    mask_bits<1>(row, &bits);
    const __m256i v = _mm256_shuffle_epi8(_mm256_set1_epi32((int)bits), c.sel);
    return _mm256_and_si256(_mm256_cmpeq_epi8(_mm256_and_si256(v, c.bit), c.bit), c.one);
}
__attribute__((target("avx2"))) static inline __m256 reward_avx2_k1(const Params& P, const Avx2K1& c, __m256i R, const u32* row) {
    const int g = 10 * (100 * (int)bf(row[1], 10, 511) + 10 * (int)bf(row[1], 0, 1023)) - P.step_size;
    const __m256i code = _mm256_and_si256(_mm256_srlv_epi32(_mm256_permutevar8x32_epi32(R, c.ridx), c.rsh), c.rmsk);
    const __m256i l10 = _mm256_i32gather_epi32(kRewardLut, _mm256_add_epi32(code, c.roff), 4);
    const __m256i num = _mm256_add_epi32(_mm256_set1_epi32(g), _mm256_slli_epi32(l10, 3));  // g + 8 * local
    return _mm256_div_ps(_mm256_cvtepi32_ps(num), c.denom);
}

static bool nt_stores() {
    static const bool v = !getenv("FJSP_DECODE_NO_NT");
    return v;
}

// one env with plain stores
__attribute__((target("avx2"))) static inline void one_avx2_k1(const Params& P, const Avx2K1& c, const u32* wire, int64_t e, float* tmp, float* obs,
                                                               int8_t* masks, float* rewards, uint8_t* flags) {
    const u32* row = wire + e * 8;
    const __m256i R = _mm256_loadu_si256((const __m256i*)row);
    if (obs) {
        obs_avx2_k1(P, c, R, row, tmp);
        memcpy(obs + e * 38, tmp, 38 * 4);
    }
    if (masks) _mm256_storeu_si256((__m256i*)(masks + e * 32), mask_avx2_k1(c, row));
    if (rewards) _mm256_storeu_ps(rewards + e * 8, reward_avx2_k1(P, c, R, row));
    if (flags) {
        const u32 f = wire_flags(row[0]);
        memcpy(flags + e * FJSP_FLAG_DIM, &f, 4);
    }
}

__attribute__((target("avx2"))) static void decode_avx2_k1(const Params& P, const u32* wire, int64_t lo, int64_t hi, float* obs,
                                                           int8_t* masks, float* rewards, uint8_t* flags) {
    constexpr int OBS = 38, MASK = 32, ACT = 8, WORDS = 8;
    const PosTab T(P);
    const Avx2K1 c = make_avx2_k1(T);
    alignas(32) float tmp[8 * OBS + 8];
    int64_t e = lo;
    // Blocks of 8 envs whose output rows start on 32-byte boundaries: written with streaming (non-temporal) stores, so
    // the 220 B per env of results do not cost a read-for-ownership of every destination line on top of the write.
    const bool aligned = (!obs || (reinterpret_cast<uintptr_t>(obs) & 31) == 0) && (!masks || (reinterpret_cast<uintptr_t>(masks) & 31) == 0) &&
                         (!rewards || (reinterpret_cast<uintptr_t>(rewards) & 31) == 0) && (!flags || (reinterpret_cast<uintptr_t>(flags) & 31) == 0);
    if (aligned && nt_stores()) {
        for (; e < hi && (e & 7); e++) one_avx2_k1(P, c, wire, e, tmp, obs, masks, rewards, flags);
        for (; e + 8 <= hi; e += 8) {
            const u32* row0 = wire + e * WORDS;
            __m256i R[8];
            for (int j = 0; j < 8; j++) R[j] = _mm256_loadu_si256((const __m256i*)(row0 + j * WORDS));
            if (obs) {
                for (int j = 0; j < 8; j++) obs_avx2_k1(P, c, R[j], row0 + j * WORDS, tmp + j * OBS);  // each row's 2 padding lanes are overwritten by the next
                float* dst = obs + e * OBS;
#pragma GCC unroll 8
                for (int i = 0; i < OBS; i++) _mm256_stream_ps(dst + 8 * i, _mm256_load_ps(tmp + 8 * i));
            }
            if (masks)
                for (int j = 0; j < 8; j++) _mm256_stream_si256((__m256i*)(masks + (e + j) * MASK), mask_avx2_k1(c, row0 + j * WORDS));
            if (rewards)
                for (int j = 0; j < 8; j++) _mm256_stream_ps(rewards + (e + j) * ACT, reward_avx2_k1(P, c, R[j], row0 + j * WORDS));
            if (flags) {
                const __m256i f = _mm256_setr_epi32((int)wire_flags(row0[0 * WORDS]), (int)wire_flags(row0[1 * WORDS]), (int)wire_flags(row0[2 * WORDS]),
                                                    (int)wire_flags(row0[3 * WORDS]), (int)wire_flags(row0[4 * WORDS]), (int)wire_flags(row0[5 * WORDS]),
                                                    (int)wire_flags(row0[6 * WORDS]), (int)wire_flags(row0[7 * WORDS]));
                _mm256_stream_si256((__m256i*)(flags + e * FJSP_FLAG_DIM), f);
            }
        }
        _mm_sfence();
    }
    for (; e < hi; e++) one_avx2_k1(P, c, wire, e, tmp, obs, masks, rewards, flags);  // tail, or everything when the caller's buffers are not 32-byte aligned
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
const char* wire_decode_isa() { return have_avx2() ? "avx2 (K = 1), generic (K > 1)" : "generic"; }

void wire_decode(int cells, const Params& P, const u32* wire, int64_t lo, int64_t hi, float* obs, int8_t* masks, float* rewards,
                 uint8_t* flags) {
    static const DecodeFn gen[4] = {decode_generic<1>, decode_generic<2>, decode_generic<3>, decode_generic<4>};
    const int k = cells < 1 ? 0 : cells > 4 ? 3 : cells - 1;
#ifdef FJSP_WIRE_X86
    if (k == 0 && have_avx2()) {
        decode_avx2_k1(P, wire, lo, hi, obs, masks, rewards, flags);
        return;
    }
#endif
    gen[k](P, wire, lo, hi, obs, masks, rewards, flags);
}

// The decode's ceiling on a box: streaming (non-temporal) stores of `bytes` from `threads` threads, the pattern the decode
// writes the tensors with.  Returns seconds for one pass.
#ifdef FJSP_WIRE_X86
__attribute__((target("avx2"))) static void stream_fill(unsigned char* p, size_t n) {
    const __m256i v = _mm256_set1_epi32(0x3f800000);
    size_t i = 0;
    for (; i < n && ((reinterpret_cast<uintptr_t>(p + i)) & 31); i++) p[i] = 0;
    for (; i + 32 <= n; i += 32) _mm256_stream_si256(reinterpret_cast<__m256i*>(p + i), v);
    for (; i < n; i++) p[i] = 0;
    _mm_sfence();
}
#endif
double host_stream_write_seconds(void* buf, size_t bytes, int threads) {
    if (threads < 1) threads = 1;
    unsigned char* p = static_cast<unsigned char*>(buf);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        const size_t lo = bytes * t / threads, hi = bytes * (t + 1) / threads;
        th.emplace_back([=] {
#ifdef FJSP_WIRE_X86
            if (have_avx2()) {
                stream_fill(p + lo, hi - lo);
                return;
            }
#endif
            memset(p + lo, 0, hi - lo);
        });
    }
    for (auto& t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace fjsp
