// fjsp_umma.cuh — tcgen05 (5th-generation tensor core) GEMM for the batched A2C trainer's actor / critic networks
// (/root/reference/networks.py:22-61, a2c.py:168-252,647-731).  Hand-written for sm_100a:
//   * operands are staged in shared memory in the canonical no-swizzle K-major core-matrix layout (8 rows x 16 bytes
//     per core matrix), accumulators live in TMEM, one elected thread issues tcgen05.mma.kind::tf32, the epilogue reads
//     the accumulator back with tcgen05.ld;
//   * fp32-level accuracy from the TF32 pipe: every fp32 operand is split on the fly into hi = tf32(x) and
//     lo = tf32(x - hi); acc += hi*hi + lo*hi + hi*lo  ("3xTF32": ~2^-21 relative error per product, fp32 accumulate).
//     PASSES = 1 is plain TF32 (reported separately, never as the headline);
//   * producers are ordinary warps (global fp32 -> registers -> split -> st.shared), because the split needs the ALUs
//     anyway; that also lets one kernel read either operand in either orientation (forward, dX = dY W^T, dW = X^T dY)
//     without transposed copies in HBM;
//   * epilogues: bias + ReLU (forward), ReLU-mask + column sums (backward through a layer: dX and the bias gradient),
//     atomic accumulate (split-K weight gradients).
// One launch works on a TABLE of problems (blockIdx.y): the 8 actors and the critic are one grouped launch per layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fjsp_kernels.cuh"

namespace fjsp {
namespace umma {

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, both K-major, tf32 inputs, fp32 accumulate; issued by ONE thread for the CTA
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait that gives up loudly: a malformed descriptor or a lost arrive would otherwise hang the GPU box (the
// MMA pipeline's barriers are completed by hardware events); ~2 s at 2 GHz, far beyond any legitimate wait
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
// the same load without the wait: several can be in flight; tmem_ld_wait() then makes all of them readable.  (The empty
// volatile asm per register after the wait keeps the compiler from moving a use of the loaded values above it.)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_pin(uint32_t* r) {
#pragma unroll
    for (int i = 0; i < 16; i++) asm volatile("" : "+r"(r[i]));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// Shared-memory matrix descriptor (tcgen05 "matrix descriptor", no swizzle, K-major): start address, leading-dimension
// byte offset (distance between the two 16-byte K-halves = core matrices adjacent in K), stride byte offset (distance
// between 8-row groups), all in units of 16 bytes; bits 46..47 = 0b01 (sm_100 descriptor version).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor, kind::tf32: D = f32 (bits 4..5 = 1), A and B = tf32 (bits 7..9, 10..12 = 2), both K-major
// (bits 15, 16 = 0), N >> 3 at bits 17..22, M >> 4 at bits 24..28.
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------
// Grouped GEMM
// ---------------------------------------------------------------------------------------------------------------
// C[M x N] (+)= A[M x K] * B[N x K]^T for every problem of a table.  How A(m, k) and B(n, k) are read from global
// memory is a template parameter of the kernel (the same for all problems of a launch):
//   OP_KC  : k contiguous        X(r, k) = X[r * ld + k]        16-byte loads (ld % 4 == 0, pointer 16-byte aligned, K % 4 == 0)
//   OP_KCS : the same, scalar loads (any alignment, any K)
//   OP_MC  : r contiguous        X(r, k) = X[k * ld + r]
//   OP_PK  : (B only) a PACKED IMAGE made by fjsp_gemm_pack_kernel from a weight matrix: per 16-wide K chunk the bytes the
//            tensor core reads from shared memory — four K-quad planes of the hi terms, then four of the lo terms, a plane
//            = npad rows x 16 bytes (npad = N rounded up to 16) — so a stage's B tile is a straight copy, no conversion.
//            Weights are constant over a whole rollout + update; activations (A) are split on the fly as before.
enum { OP_KC = 0, OP_KCS = 1, OP_MC = 2, OP_PK = 3 };
enum { GEMM_RELU = 1, GEMM_ATOMIC = 2 };

struct GemmProb {          // 128 bytes; device array, one per problem
    const float* A;
    const float* B;
    float* C;
    const float* bias;     // [N] added to every row, or null
    const float* mask;     // same indexing as C: C = acc * (mask > 0), or null
    float* colsum;         // [N]: += column sums of what was stored (bias gradient), or null
    int32_t M, N, K;       // N <= 256
    int32_t lda, ldb;
    int32_t csm, csn;      // C(m, n) = C[m * csm + n * csn]
    int32_t flags;         // GEMM_RELU, GEMM_ATOMIC (atomicAdd into C: split-K partial sums)
    int32_t splitk;        // >= 1: the K range is cut into this many parts, one CTA each
    int32_t head_n;        // 0: rowdot_* is the one-column head below; 1..8: a fused head of that many columns (see below)
    int32_t head_ld;       // head_n >= 1: row stride of rowdot_out in floats
    int32_t reserved;
    // optional fused head (head_n = 0, N <= 128): rowdot_out[m] = sum_n stored(m, n) * rowdot_w[n] + rowdot_bias[0] — the critic's
    // last layer (128 -> 1, networks.py:41-61) in the epilogue of the layer before it.
    // head_n = R in 1..8 (any N <= 256): rowdot_out[m * head_ld + j] = sum_n stored(m, n) * rowdot_w[n * R + j] + rowdot_bias[j],
    // j < R — an actor's last layer (256 -> 3..8 logits, networks.py:22-38) as fp32 FMAs in the epilogue of its second layer:
    // as a GEMM launch of its own it costs a rollout step as much as the 256 x 256 layer (16 K chunks, latency-bound)
    const float* rowdot_w;
    float* rowdot_out;
    const float* rowdot_bias;
    int64_t reserved2;
};
static_assert(sizeof(GemmProb) == 128, "GemmProb layout is part of the ABI (include/fjsp_b200.h FjspGemmProb)");

constexpr int G_BM = 128;             // rows per CTA tile = TMEM lanes
constexpr int G_BN = 256;             // max columns per CTA tile = TMEM columns
constexpr int G_KC = 16;              // K per pipeline stage (two tf32 MMAs of K = 8)
constexpr int G_NQ = G_KC / 4;        // 16-byte K-quads per stage
constexpr int G_PAD = 32;             // bytes between K-quad planes: keeps the producers' st.shared.v4 conflict-free
constexpr int G_LBO_A = G_BM * 16 + G_PAD, G_LBO_B = G_BN * 16 + G_PAD;  // K-quad plane = all rows' 16-byte pieces
constexpr int G_A_BYTES = G_NQ * G_LBO_A, G_B_BYTES = G_NQ * G_LBO_B;    // one of (hi, lo)
constexpr int G_STAGE_BYTES = 2 * (G_A_BYTES + G_B_BYTES);
constexpr int G_STAGES = 2;
constexpr int G_SMEM_BYTES = G_STAGES * G_STAGE_BYTES;                   // 99,328: two CTAs per SM
constexpr int G_EPI_LD = 132;          // floats per row of the epilogue's staging area (128 columns + 4: conflict-free)
static_assert(G_BM * G_EPI_LD * 4 + G_BM * 8 * 4 <= G_SMEM_BYTES, "the epilogue stages 128 x 128 floats in the pipeline's shared memory");
constexpr int G_PRODUCERS = 256;      // warps 0..7 load; warps 0..3 are also the epilogue; warp 8 issues the MMAs
constexpr int G_THREADS = G_PRODUCERS + 32;

__device__ __forceinline__ uint32_t plane_off(int r, int q, int lbo) { return (uint32_t)(q * lbo + (r >> 3) * 128 + (r & 7) * 16); }

// fp32 -> (hi, lo) with hi = the nearest TF32 (round half away: one integer add + mask) and lo = x - hi (exact in fp32),
// of which the tensor core reads the leading 11 bits.
// (cvt.rna.tf32.f32 would do the same rounding on the XU pipe at 16 lanes per clock per SM: two of them per element made
// the conversion, not the MMAs, the limiter of the first version of this kernel — ncu: xu pipe saturated, tensor 18 %.)
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
template <bool ROUND_LO>
__device__ __forceinline__ void split_store(unsigned char* hi, unsigned char* lo, uint32_t off, float4 v, bool three) {
    const uint4 h = make_uint4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    *reinterpret_cast<uint4*>(hi + off) = h;
    if (three) {
        const float4 l = make_float4(v.x - __uint_as_float(h.x), v.y - __uint_as_float(h.y), v.z - __uint_as_float(h.z),
                                     v.w - __uint_as_float(h.w));
        if (ROUND_LO) {
            // lo rounded to TF32 as well: left to the hardware it is CUT to 11 bits, a bias that adds up linearly over the
            // K = batch-size sums of the weight gradients (there the two extra integer ops per element are spent)
            *reinterpret_cast<uint4*>(lo + off) = make_uint4(tf32_hi(l.x), tf32_hi(l.y), tf32_hi(l.z), tf32_hi(l.w));
        } else {
            *reinterpret_cast<float4*>(lo + off) = l;
        }
    }
}

// One operand tile of ROWS rows x G_KC k, fetched into registers (`fetch`) and later split + stored (`stash`).
template <int OP, int ROWS, bool ROUND_LO>
struct Loader {
    static constexpr int ITEMS = OP == OP_PK ? 2 * G_NQ * ROWS / G_PRODUCERS : OP == OP_MC ? (ROWS == 128 ? G_NQ / 2 : G_NQ) : ROWS * G_NQ / G_PRODUCERS;
    float4 v[ITEMS];
    // OP_PK: `ld` = npad (rows of a plane), `rmax` = number of 16-byte pieces of a chunk to move (8 * npad, or 4 * npad when
    // only the hi terms are used)
    __device__ __forceinline__ void fetch(const float* __restrict__ X, int ld, int r0, int rmax, int k0, int kmax, int tid) {
        const bool kfull = k0 + G_KC <= kmax;  // uniform: interior stages need no per-element K guards
        if (OP == OP_PK) {
            const float4* src = reinterpret_cast<const float4*>(X) + (int64_t)(k0 / G_KC) * (2 * G_NQ * ld);
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const int idx = tid + i * G_PRODUCERS;
                if (idx < rmax) v[i] = __ldg(src + idx);
            }
        } else if (OP == OP_MC) {
            // thread -> one row r (contiguous across the warp), a few K-quads; 4 coalesced scalar loads per quad
            const int r = ROWS == 128 ? (tid & 127) : tid;
            const int qb = ROWS == 128 ? (tid >> 7) : 0, qs = ROWS == 128 ? 2 : 1;
            const bool rok = r0 + r < rmax;
            const float* p0 = X + (int64_t)(k0 + 4 * qb) * ld + r0 + r;
            const int64_t qstep = (int64_t)4 * qs * ld;
            if (kfull && rok) {
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    const float* p = p0 + i * qstep;
                    v[i] = make_float4(__ldg(p), __ldg(p + ld), __ldg(p + 2 * (int64_t)ld), __ldg(p + 3 * (int64_t)ld));
                }
            } else {
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    const int k = k0 + 4 * (qb + qs * i);
                    const float* p = p0 + i * qstep;
                    v[i].x = (rok && k + 0 < kmax) ? __ldg(p) : 0.f;
                    v[i].y = (rok && k + 1 < kmax) ? __ldg(p + ld) : 0.f;
                    v[i].z = (rok && k + 2 < kmax) ? __ldg(p + 2 * (int64_t)ld) : 0.f;
                    v[i].w = (rok && k + 3 < kmax) ? __ldg(p + 3 * (int64_t)ld) : 0.f;
                }
            }
        } else {
            // thread -> (row, K-quad) with the quad index fastest: a warp reads 8 rows x 64 contiguous bytes
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const int idx = tid + i * G_PRODUCERS, r = idx / G_NQ, q = idx % G_NQ;
                const int k = k0 + 4 * q;
                const float* p = X + (int64_t)(r0 + r) * ld + k;
                const bool rok = r0 + r < rmax;
                if (OP == OP_KC) {
                    v[i] = (rok && k < kmax) ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else if (kfull && rok) {
                    v[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
                } else {
                    v[i].x = (rok && k + 0 < kmax) ? __ldg(p) : 0.f;
                    v[i].y = (rok && k + 1 < kmax) ? __ldg(p + 1) : 0.f;
                    v[i].z = (rok && k + 2 < kmax) ? __ldg(p + 2) : 0.f;
                    v[i].w = (rok && k + 3 < kmax) ? __ldg(p + 3) : 0.f;
                }
            }
        }
    }
    __device__ __forceinline__ void stash(unsigned char* hi, unsigned char* lo, int lbo, int tid, bool three) const {
        if (OP == OP_PK) {   // `lbo` = number of 16-byte pieces to move; hi planes then lo planes are contiguous from `hi`
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const int idx = tid + i * G_PRODUCERS;
                if (idx < lbo) reinterpret_cast<float4*>(hi)[idx] = v[i];
            }
        } else if (OP == OP_MC) {
            const int r = ROWS == 128 ? (tid & 127) : tid;
            const int qb = ROWS == 128 ? (tid >> 7) : 0, qs = ROWS == 128 ? 2 : 1;
#pragma unroll
            for (int i = 0; i < ITEMS; i++) split_store<ROUND_LO>(hi, lo, plane_off(r, qb + qs * i, lbo), v[i], three);
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const int idx = tid + i * G_PRODUCERS;
                split_store<ROUND_LO>(hi, lo, plane_off(idx / G_NQ, idx % G_NQ, lbo), v[i], three);
            }
        }
    }
};

// 16 per-thread values (one row, 16 columns) -> per-column sums over the warp's 32 rows: transposing butterfly, 16
// shuffles; on return lane l holds in v[0] the sum of column ((l >> 4) & 1) * 8 + ((l >> 3) & 1) * 4 + ((l >> 2) & 1) * 2 + ((l >> 1) & 1)
__device__ __forceinline__ float warp_colsum16(float* v, int lane) {
#pragma unroll
    for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
        const bool up = lane & bit;
#pragma unroll
        for (int i = 0; i < w; i++) {
            const float keep = up ? v[i + w] : v[i], send = up ? v[i] : v[i + w];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <int AOP, int BOP>
__global__ void __launch_bounds__(G_THREADS, 2) fjsp_gemm_kernel(const GemmProb* __restrict__ probs, int passes) {
    extern __shared__ __align__(128) unsigned char g_smem[];
    __shared__ uint64_t s_full[G_STAGES], s_empty[G_STAGES], s_acc;
    __shared__ uint32_t s_tmem;
    __shared__ float s_colsum[G_BN], s_bias[G_BN];
    __shared__ __align__(16) float s_head_w[G_BN * 8];   // head_n >= 1: the head's weights, [n][8] (columns >= head_n zero)

    // Launched as a programmatic dependent (fjsp_api.cu launch_gemm): this grid may become resident while its predecessor in
    // the stream drains; the problem table is immutable, everything else is read after griddepcontrol.wait below.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const GemmProb P = probs[blockIdx.y];
    const int mtiles = (P.M + G_BM - 1) / G_BM;
    const int splitk = P.splitk > 0 ? P.splitk : 1;
    if ((int)blockIdx.x >= mtiles * splitk) return;
    const int mt = blockIdx.x % mtiles, sp = blockIdx.x / mtiles;
    const int m0 = mt * G_BM;
    const int npad = (P.N + 15) & ~15;
    const int chunks_all = (P.K + G_KC - 1) / G_KC;
    const int per = (chunks_all + splitk - 1) / splitk;
    const int c_begin = sp * per, c_end = min(chunks_all, c_begin + per);
    const int nchunks = c_end - c_begin;
    if (nchunks <= 0) return;
    const bool three = passes >= 3;
    uint32_t ncols = 32;
    while ((int)ncols < npad) ncols <<= 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        // OP_PK: one more arrival per stage, the B loader's arrive.expect_tx (its bulk copy completes the transaction bytes)
        for (int s = 0; s < G_STAGES; s++) mbar_init(&s_full[s], G_PRODUCERS / 32 + (BOP == OP_PK ? 1 : 0)), mbar_init(&s_empty[s], 1);
        mbar_init(&s_acc, 1);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the predecessor's writes (activations, weights) are visible from here
    if (tid == 32 && AOP != OP_MC) {
        // The pipeline reads the A tile 64 bytes per row and stage: 128 row segments 1 KB apart, a DRAM-hostile pattern
        // (ncu, first version: long-scoreboard stalls on these loads dominated, DRAM at 19 % with the tensor pipe at 18 %).
        // The rows of a tile are (nearly) contiguous in memory, so ONE bulk L2 prefetch streams the whole tile — and the
        // ReLU-mask tile the epilogue will read — in DRAM order; the strided loads then hit L2.
        const int rows = min(G_BM, P.M - m0);
        const int kb = c_begin * G_KC, kn = min(P.K, c_end * G_KC) - kb;
        if (P.lda <= 2 * P.K) {
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(P.A + (int64_t)m0 * P.lda + kb);
            const uintptr_t a1 = a0 + ((size_t)(rows - 1) * P.lda + kn) * 4;
            bulk_prefetch_l2(reinterpret_cast<const void*>(a0 & ~(uintptr_t)15), (uint32_t)(((a1 + 15) & ~(uintptr_t)15) - (a0 & ~(uintptr_t)15)));
        }
        if (P.mask && P.csn == 1 && P.csm <= 2 * P.N) {
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(P.mask + (int64_t)m0 * P.csm);
            const uintptr_t a1 = a0 + ((size_t)(rows - 1) * P.csm + P.N) * 4;
            bulk_prefetch_l2(reinterpret_cast<const void*>(a0 & ~(uintptr_t)15), (uint32_t)(((a1 + 15) & ~(uintptr_t)15) - (a0 & ~(uintptr_t)15)));
        }
    }
    if (tid < G_BN) s_colsum[tid] = 0.f, s_bias[tid] = (P.bias && tid < P.N) ? __ldg(P.bias + tid) : 0.f;
    const int head_n = P.rowdot_w ? P.head_n : 0;
    if (head_n > 0) {
        for (int i = tid; i < G_BN * 8; i += G_THREADS) {
            const int n = i >> 3, j = i & 7;
            s_head_w[i] = (n < P.N && j < head_n) ? __ldg(P.rowdot_w + n * head_n + j) : 0.f;
        }
    }
    if (warp == G_PRODUCERS / 32) tmem_alloc(&s_tmem, ncols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem;

    if (warp < G_PRODUCERS / 32) {
        // ===== producers: global fp32 -> registers -> (hi, lo) tf32 -> shared memory, one stage ahead in registers =====
        // lo terms rounded (not cut) to TF32 everywhere: two integer ops per element buy unbiased errors, which matters
        // when gradients flow through three chained GEMMs and are then summed over the batch (critic W2: 2.6e-4 of the
        // largest gradient with cut lo terms, measured against float64)
        constexpr bool ROUND_LO = true;
        // TWO chunks ahead in registers (two loader sets, the loop unrolled by two): with one set the loads of chunk c + 1
        // were only issued after chunk c had been stored, so every chunk paid a full load latency in series
        // (ncu: 60 % of the cycles no warp eligible, long scoreboard) — now a chunk's loads have a whole chunk time to land.
        Loader<AOP, G_BM, ROUND_LO> la0, la1;
        Loader<BOP, G_BN, ROUND_LO> lb0, lb1;
        auto fetch = [&](Loader<AOP, G_BM, ROUND_LO>& la, Loader<BOP, G_BN, ROUND_LO>& lb, int c) {
            la.fetch(P.A, P.lda, m0, P.M, (c_begin + c) * G_KC, P.K, tid);
            if (BOP != OP_PK) lb.fetch(P.B, P.ldb, 0, P.N, (c_begin + c) * G_KC, P.K, tid);   // OP_PK: the B loader below
        };
        auto stash = [&](const Loader<AOP, G_BM, ROUND_LO>& la, const Loader<BOP, G_BN, ROUND_LO>& lb, int c) {
            const int s = c % G_STAGES;
            if (c >= G_STAGES) mbar_wait_or_trap(&s_empty[s], ((c / G_STAGES) - 1) & 1);
            unsigned char* st = g_smem + s * G_STAGE_BYTES;
            la.stash(st, st + G_A_BYTES, G_LBO_A, tid, three);
            if (BOP != OP_PK) lb.stash(st + 2 * G_A_BYTES, st + 2 * G_A_BYTES + G_B_BYTES, G_LBO_B, tid, three);
            fence_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_full[s]);
        };
        constexpr bool TWO_AHEAD = true;
        fetch(la0, lb0, 0);
        if (TWO_AHEAD) {
            if (nchunks > 1) fetch(la1, lb1, 1);
            for (int c = 0; c < nchunks; c += 2) {
                stash(la0, lb0, c);
                if (c + 2 < nchunks) fetch(la0, lb0, c + 2);
                if (c + 1 < nchunks) {
                    stash(la1, lb1, c + 1);
                    if (c + 3 < nchunks) fetch(la1, lb1, c + 3);
                }
            }
        } else {
            for (int c = 0; c < nchunks; c++) {
                stash(la0, lb0, c);
                if (c + 1 < nchunks) fetch(la0, lb0, c + 1);
            }
        }
    } else if (lane == 0) {
        // ===== MMA issuer: one thread =====
        const uint32_t idesc = instr_desc_tf32(G_BM, npad);
        const uint32_t sbase = smem_u32(g_smem);
        const uint32_t lbo_b = BOP == OP_PK ? (uint32_t)npad * 16u : (uint32_t)G_LBO_B;   // bytes between the K-quad planes of B
        for (int c = 0; c < nchunks; c++) {
            const int s = c % G_STAGES;
            mbar_wait_or_trap(&s_full[s], (c / G_STAGES) & 1);
            fence_after_sync();
            const uint32_t a_hi = sbase + s * G_STAGE_BYTES, a_lo = a_hi + G_A_BYTES;
            const uint32_t b_hi = a_hi + 2 * G_A_BYTES, b_lo = b_hi + G_NQ * lbo_b;
#pragma unroll
            for (int ks = 0; ks < G_KC / 8; ks++) {
                const uint64_t ah = smem_desc(a_hi + 2 * ks * G_LBO_A, G_LBO_A, 128), bh = smem_desc(b_hi + 2 * ks * lbo_b, lbo_b, 128);
                if (three) {  // small terms first
                    const uint64_t al = smem_desc(a_lo + 2 * ks * G_LBO_A, G_LBO_A, 128), bl = smem_desc(b_lo + 2 * ks * lbo_b, lbo_b, 128);
                    mma_tf32(tmem, al, bh, idesc, (c | ks) != 0);
                    mma_tf32(tmem, ah, bl, idesc, 1u);
                    mma_tf32(tmem, ah, bh, idesc, 1u);
                } else {
                    mma_tf32(tmem, ah, bh, idesc, (c | ks) != 0);
                }
            }
            mma_commit(&s_empty[s]);  // the stage is free again once these MMAs have read it
        }
        mma_commit(&s_acc);           // accumulator complete
    } else if (BOP == OP_PK && lane == 1) {
        // ===== B loader (packed weight images): one thread, one bulk copy per chunk =====
        // A chunk of the image is the stage's B tile byte for byte (hi planes, then lo planes), so the copy engine moves it:
        // no thread touches the weights.  (Through registers — 8 + 8 vector loads / stores per producer thread — the B tile
        // cost a chunk ~1 us whichever form the weights had: tools/gemm_chunk_cost.py.)
        const uint32_t bytes = (uint32_t)((three ? 2 : 1) * G_NQ * npad) * 16u;
        const size_t chunk_stride = (size_t)2 * G_NQ * npad * 16;   // bytes per chunk in the image (always hi + lo)
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.B) + (size_t)c_begin * chunk_stride;
        for (int c = 0; c < nchunks; c++) {
            const int s = c % G_STAGES;
            if (c >= G_STAGES) mbar_wait_or_trap(&s_empty[s], ((c / G_STAGES) - 1) & 1);
            mbar_expect_tx(&s_full[s], bytes);
            bulk_g2s(g_smem + s * G_STAGE_BYTES + 2 * G_A_BYTES, src + (size_t)c * chunk_stride, bytes, &s_full[s]);
        }
    }

    if (warp < G_PRODUCERS / 32) {
        // ===== epilogue =====
        // Phase 1, thread = one row of the tile (TMEM lane 32 * warp + lane): accumulator -> registers (tcgen05.ld) -> + bias,
        // ReLU -> shared memory (the pipeline stages are free now), 128 columns at a time, row stride 132 floats so the
        // 16-byte stores of 8 consecutive rows fall into 8 different bank groups.
        // Phase 2, warp = the same 32 rows, lane = 4 consecutive columns: one row per iteration leaves as ONE coalesced
        // 512-byte store (or 128 coalesced atomics); the ReLU mask is read the same way, column sums stay in registers.
        // All eight producer warps take part: warps q and q + 4 may both read TMEM lanes 32q..32q+31 (a warp's lane quarter is
        // warp % 4), so they share those 32 rows — in phase 1 each converts 64 of the 128 columns, in phase 2 each stores 16 of
        // the 32 rows — and meet at a 64-thread named barrier between the phases.  (With four warps the epilogue of a
        // 128 x 256 tile cost as much as ~6 K-chunks of the main loop; in the one-wave launches of a rollout step nothing
        // overlaps it.)
        const int q = warp & 3, half = warp >> 2;
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };
        mbar_wait_or_trap(&s_acc, 0);
        fence_after_sync();
        const bool atomic = P.flags & GEMM_ATOMIC, relu = P.flags & GEMM_RELU;
        const bool vec = !atomic && P.csn == 1 && (P.csm & 3) == 0 && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0) &&
                         (!P.mask || (reinterpret_cast<uintptr_t>(P.mask) & 15) == 0);
        float* stage = reinterpret_cast<float*>(g_smem) + (q * 32) * G_EPI_LD;
        float hd[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // head_n >= 1: this thread's row, its half of the columns
        for (int h0 = 0; h0 < npad; h0 += 128) {
            const int ncol = min(128, npad - h0);
            // this warp's 64 columns: two TMEM loads in flight per wait (four would need 64 registers: spills)
#pragma unroll
            for (int kk = 0; kk < 4; kk += 2) {
                uint32_t raw[2][16];
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int n0 = 64 * half + 16 * (kk + k);
                    if (n0 < ncol) tmem_ld16_issue(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h0 + n0), raw[k]);
                }
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int n0 = 64 * half + 16 * (kk + k);
                    if (n0 < ncol) {
                        tmem_ld_pin(raw[k]);
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; i++) v[i] = __uint_as_float(raw[k][i]);
                        if (P.bias) {
#pragma unroll
                            for (int i = 0; i < 16; i++) v[i] += s_bias[h0 + n0 + i];
                        }
                        if (relu) {
#pragma unroll
                            for (int i = 0; i < 16; i++) v[i] = fmaxf(v[i], 0.f);
                        }
                        float4* dst = reinterpret_cast<float4*>(stage + lane * G_EPI_LD + n0);
#pragma unroll
                        for (int i = 0; i < 4; i++) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        if (head_n > 0) {   // warp-uniform weight reads (broadcast), eight FMAs per stored value
                            const float4* hw = reinterpret_cast<const float4*>(s_head_w) + 2 * (h0 + n0);
#pragma unroll
                            for (int i = 0; i < 16; i++) {
                                const float4 wa = hw[2 * i], wb = hw[2 * i + 1];
                                hd[0] = fmaf(v[i], wa.x, hd[0]), hd[1] = fmaf(v[i], wa.y, hd[1]), hd[2] = fmaf(v[i], wa.z, hd[2]);
                                hd[3] = fmaf(v[i], wa.w, hd[3]), hd[4] = fmaf(v[i], wb.x, hd[4]), hd[5] = fmaf(v[i], wb.y, hd[5]);
                                hd[6] = fmaf(v[i], wb.z, hd[6]), hd[7] = fmaf(v[i], wb.w, hd[7]);
                            }
                        }
                    }
                }
            }
            pair_sync();
            const int n = h0 + 4 * lane;         // this lane's 4 columns
            const bool nany = 4 * lane < ncol && n < P.N;
            float cs[4] = {0.f, 0.f, 0.f, 0.f};
            float w4[4] = {0.f, 0.f, 0.f, 0.f};
            const bool rowdot1 = P.rowdot_w && head_n == 0;
            if (rowdot1 && nany) {
#pragma unroll
                for (int i = 0; i < 4; i++) w4[i] = n + i < P.N ? __ldg(P.rowdot_w + n + i) : 0.f;
            }
            const int rbeg = 16 * half, rows_here = max(0, min(16, P.M - (m0 + q * 32 + rbeg)));
            const bool fast = vec && !rowdot1 && n + 4 <= P.N;   // (per lane; the loop bounds are uniform)
            if (!rowdot1 && __all_sync(0xffffffffu, fast || !nany)) {
                // every active lane stores whole 16-byte pieces: two rows per iteration, the gate loads issued before the stores
                for (int rr = 0; rr < rows_here; rr += 2) {
                    float4 o[2], k4[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int r = rbeg + rr + u;
                        if (nany && rr + u < rows_here) {
                            o[u] = *reinterpret_cast<const float4*>(stage + r * G_EPI_LD + 4 * lane);
                            if (P.mask) k4[u] = __ldg(reinterpret_cast<const float4*>(P.mask + (int64_t)(m0 + q * 32 + r) * P.csm + n));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int r = rbeg + rr + u;
                        if (nany && rr + u < rows_here) {
                            if (P.mask) {
                                o[u].x = k4[u].x > 0.f ? o[u].x : 0.f, o[u].y = k4[u].y > 0.f ? o[u].y : 0.f;
                                o[u].z = k4[u].z > 0.f ? o[u].z : 0.f, o[u].w = k4[u].w > 0.f ? o[u].w : 0.f;
                            }
                            *reinterpret_cast<float4*>(P.C + (int64_t)(m0 + q * 32 + r) * P.csm + n) = o[u];
                            cs[0] += o[u].x, cs[1] += o[u].y, cs[2] += o[u].z, cs[3] += o[u].w;
                        }
                    }
                }
            } else
            for (int r = 16 * half; r < 16 * half + 16; r++) {
                const int m = m0 + q * 32 + r;
                if (m >= P.M) break;              // uniform over the warp
                if (rowdot1) {                    // fused head: every lane takes part in the row's reduction
                    float d = 0.f;
                    if (nany) {
                        const float4 o = *reinterpret_cast<const float4*>(stage + r * G_EPI_LD + 4 * lane);
                        d = o.x * w4[0] + o.y * w4[1] + o.z * w4[2] + o.w * w4[3];
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
                    if (lane == 0) P.rowdot_out[m] = d + (P.rowdot_bias ? __ldg(P.rowdot_bias) : 0.f);
                }
                if (!nany) continue;
                float4 o = *reinterpret_cast<const float4*>(stage + r * G_EPI_LD + 4 * lane);
                float* cp = P.C + (int64_t)m * P.csm + (int64_t)n * P.csn;
                if (vec && n + 4 <= P.N) {
                    if (P.mask) {
                        const float4 k4 = __ldg(reinterpret_cast<const float4*>(P.mask + (int64_t)m * P.csm + n));
                        o.x = k4.x > 0.f ? o.x : 0.f, o.y = k4.y > 0.f ? o.y : 0.f, o.z = k4.z > 0.f ? o.z : 0.f, o.w = k4.w > 0.f ? o.w : 0.f;
                    }
                    *reinterpret_cast<float4*>(cp) = o;
                    cs[0] += o.x, cs[1] += o.y, cs[2] += o.z, cs[3] += o.w;
                } else {
                    const float e[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (n + i < P.N) {
                            float x = e[i];
                            if (P.mask && !(__ldg(P.mask + (int64_t)m * P.csm + (int64_t)(n + i) * P.csn) > 0.f)) x = 0.f;
                            if (atomic) atomicAdd(cp + (int64_t)i * P.csn, x);
                            else cp[(int64_t)i * P.csn] = x;
                            cs[i] += x;
                        }
                    }
                }
            }
            if (P.colsum && nany) {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (n + i < P.N) atomicAdd(&s_colsum[n + i], cs[i]);
            }
            pair_sync();    // the staging rows are rewritten by the next column block
        }
        if (head_n > 0) {   // the pair's two column halves of a row meet behind the staging area; warp q adds the bias and stores
            float* part = reinterpret_cast<float*>(g_smem) + G_BM * G_EPI_LD + (q * 32 + lane) * 8;
            if (half == 1) {
                reinterpret_cast<float4*>(part)[0] = make_float4(hd[0], hd[1], hd[2], hd[3]);
                reinterpret_cast<float4*>(part)[1] = make_float4(hd[4], hd[5], hd[6], hd[7]);
            }
            pair_sync();
            const int m = m0 + q * 32 + lane;
            if (half == 0 && m < P.M) {
                const float4 pa = reinterpret_cast<const float4*>(part)[0], pb = reinterpret_cast<const float4*>(part)[1];
                const float o[8] = {hd[0] + pa.x, hd[1] + pa.y, hd[2] + pa.z, hd[3] + pa.w, hd[4] + pb.x, hd[5] + pb.y, hd[6] + pb.z, hd[7] + pb.w};
                float* out = P.rowdot_out + (int64_t)m * P.head_ld;
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (j < head_n) out[j] = o[j] + (P.rowdot_bias ? __ldg(P.rowdot_bias + j) : 0.f);
            }
        }
        fence_before_sync();
    }
    __syncthreads();
    if (P.colsum && tid < P.N) atomicAdd(P.colsum + tid, s_colsum[tid]);
    if (warp == G_PRODUCERS / 32) {
        __syncwarp();
        fence_after_sync();
        tmem_dealloc(tmem, ncols);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Packed B images (OP_PK): one job per weight matrix and orientation.  image[chunk][hi | lo][quad][npad rows][4 k] floats,
// zero beyond N and K; the (hi, lo) split is the one the loaders apply on the fly (tf32_hi of x and of x - hi).
// ---------------------------------------------------------------------------------------------------------------
struct PackJob {           // 32 bytes; device array
    const float* src;
    float* dst;
    int32_t op;            // orientation of src: OP_KC / OP_KCS: B(n, k) = src[n * ld + k]; OP_MC: src[k * ld + n]
    int32_t ld, N, K;
};
static_assert(sizeof(PackJob) == 32, "PackJob layout is part of the ABI (include/fjsp_b200.h FjspPackJob)");

__global__ void fjsp_gemm_pack_kernel(const PackJob* __restrict__ jobs) {
    const PackJob J = jobs[blockIdx.y];
    const int npad = (J.N + 15) & ~15, chunks = (J.K + G_KC - 1) / G_KC;
    const int total = chunks * G_NQ * npad;   // one thread per (chunk, quad, row): 4 k values -> one hi and one lo piece
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i % npad, q = (i / npad) % G_NQ, c = i / (npad * G_NQ);
        float x[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = c * G_KC + 4 * q + j;
            x[j] = (n < J.N && k < J.K) ? __ldg(J.src + (J.op == OP_MC ? (int64_t)k * J.ld + n : (int64_t)n * J.ld + k)) : 0.f;
        }
        const uint4 h = make_uint4(tf32_hi(x[0]), tf32_hi(x[1]), tf32_hi(x[2]), tf32_hi(x[3]));
        const uint4 l = make_uint4(tf32_hi(x[0] - __uint_as_float(h.x)), tf32_hi(x[1] - __uint_as_float(h.y)),
                                   tf32_hi(x[2] - __uint_as_float(h.z)), tf32_hi(x[3] - __uint_as_float(h.w)));
        uint4* chunk = reinterpret_cast<uint4*>(J.dst) + (int64_t)c * (2 * G_NQ * npad);
        chunk[q * npad + n] = h;
        chunk[(G_NQ + q) * npad + n] = l;
    }
}


}  // namespace umma
}  // namespace fjsp
