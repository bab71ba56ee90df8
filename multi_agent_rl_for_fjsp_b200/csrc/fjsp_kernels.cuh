// fjsp_kernels.cuh — sm_100a kernels of the batched FJSP environment (templated on K, the number of cells of the shop;
// K = 1 is the reference shop, K = 2..4 the scaled shop of DESIGN.md §10).
//
// HBM layout: the packed state is an ARRAY OF TILES.  One tile = 64 envs x W words (W = 128 for K = 1),
// stored word-major  tile[w][lane]  (u32), contiguous in HBM.  A CTA owns one tile per step:
//   1. one elected thread issues ONE bulk async copy (TMA engine, cp.async.bulk, SASS UBLKCP) of the tile's
//      dynamically indexed words (24 .. 64+64K: completion steps, orders, tray pools; 26 KB for K = 1) into shared memory
//      and arms an mbarrier with the byte count; all threads meanwhile fetch their hot words (coalesced 32-bit loads: a
//      warp reads 128 consecutive bytes per word) and the action bytes of their env;
//   2. every thread steps its own env: hot words in registers, the rest out of shared memory (thread t owns column t
//      of the sub-tile, so all accesses of a warp hit 32 distinct banks regardless of the word index each lane follows —
//      the FIFO pointer chasing is conflict-free by construction);
//   3. observations are staged row-major in shared memory; masks / rewards / flags leave as 128-bit stores;
//   4. hot words go back with coalesced 32-bit stores; after a proxy fence + barrier the elected thread issues two bulk
//      async stores (the sub-tile, the observation rows).
// Nothing but the step's own inputs and outputs crosses HBM: for K = 1, 512 B state in, 512 B state out,
// 8 + 152 + 32 + 32 + 4 bytes of I/O per env-step.
#pragma once

#include <cuda_runtime.h>

#include "fjsp_core.h"

namespace fjsp {

constexpr int TILE = FJSP_TILE_ENVS;  // 64 envs per tile

template <int K, bool LONG = false>
struct Geo {
    static constexpr int WORDS = Lay<K, LONG>::TOTAL;               // u32 per env (compact 128 / 380, long 256 / 520)
    static constexpr int TILE_WORDS = WORDS * TILE;
    static constexpr int TILE_BYTES = TILE_WORDS * 4;               // 32768 / 97280
    static constexpr int DYN0 = Lay<K, LONG>::DYN0;                 // first word that goes through shared memory
    static constexpr int DYN_WORDS = (Lay<K, LONG>::DYN_END - DYN0) * TILE;
    static constexpr int DYN_BYTES = DYN_WORDS * 4;                 // 26624 / 75776: through shared memory (one bulk copy)
    static constexpr int OBS_ROW_BYTES = Lay<K>::OBS * 4;           // 152 / 524
    static constexpr int OBS_TILE_BYTES = OBS_ROW_BYTES * TILE;     // 9728 / 33536
    static constexpr int STEP_SMEM_BYTES = DYN_BYTES + OBS_TILE_BYTES + 16;
    static constexpr int WIRE_ROW_BYTES = Wire<K>::WORDS * 4;        // 32 / 88
    static constexpr int WIRE_TILE_BYTES = WIRE_ROW_BYTES * TILE;    // 2048 / 5632
    static constexpr int STEP_WIRE_SMEM_BYTES = DYN_BYTES + WIRE_TILE_BYTES + 16;
    static constexpr int ROLLOUT_SMEM_BYTES = DYN_BYTES + 16;
    // cell-parallel step (K >= 2): TILE x K threads per CTA, + the exchange block
    static constexpr int X_BYTES = Xl<K, LONG>::WORDS * TILE * 4;                 // 6144 for K = 4
    static constexpr int CELLS_SMEM_BYTES = DYN_BYTES + OBS_TILE_BYTES + X_BYTES + 16;       // 115,472 for K = 4: 2 CTAs / SM
    static constexpr int CELLS_WIRE_SMEM_BYTES = DYN_BYTES + WIRE_TILE_BYTES + X_BYTES + 16;
    // resident CTAs per SM by shared memory: 228 KB per SM, 1 KB of it reserved per CTA (K = 4 compact: 2 x 116,496 fit).
    // It is also the kernel's register budget (__launch_bounds__): 65,536 / (64K threads x CTAs).
    static constexpr int CELLS_FIT = (228 * 1024) / (CELLS_SMEM_BYTES + 1024);
    static constexpr int CELLS_CTAS_PER_SM = CELLS_FIT >= 3 ? 3 : CELLS_FIT >= 2 ? 2 : 1;
    static constexpr int ROLLOUT_CELLS_SMEM_BYTES = DYN_BYTES + X_BYTES + 16;
};

// One env of a tile inside a kernel: the dynamically indexed words live in shared memory, column `lane` of the
// sub-tile; the hot words are read/written straight from/to the HBM tile with compile-time indices (a warp touches 128
// consecutive bytes per word: coalesced) and otherwise live in registers (struct Hot / HotCell).
// Long layout only (fields unused otherwise): where an order's attributes come from when the pickup station pops it —
// this env's explicit order table, or the Philox order stream — and the arrival stream.
#define FJSP_ORDER_SOURCE_MEMBERS                                                                          \
    uint64_t seed = 0, genv = 0;                                                                           \
    const FjspOrderRec* otab = nullptr;                                                                    \
    u32* rq = nullptr; /* this env's ready FIFO (HBM side buffer, READY_FIFO_WORDS entries) */             \
    __device__ __forceinline__ u32 rq_ld(int i) const { return rq[i & (READY_FIFO_WORDS - 1)]; }           \
    __device__ __forceinline__ void rq_st(int i, u32 v) { rq[i & (READY_FIFO_WORDS - 1)] = v; }            \
    __device__ __forceinline__ u32 fetch_order(int o, u32 episode) const {                                 \
        if (otab) {                                                                                        \
            bool bad = false;                                                                              \
            return order_from_rec(__ldg(otab + o), bad);                                                   \
        }                                                                                                  \
        return philox_order(seed, genv, episode, o);                                                       \
    }                                                                                                      \
    __device__ __forceinline__ u32 arrival_draw(u32 episode, u32 step) const {                             \
        u32 r[4];                                                                                          \
        philox4x32_10((u32)genv, episode, step, 5u, (u32)seed, (u32)(seed >> 32), r);                      \
        return r[0];                                                                                       \
    }
template <bool LONG_ = false>
struct TileColumnT {
    static constexpr bool LONG = LONG_;
    u32* dyn;  // &s_dyn[0][lane] - DYN0 * TILE, so dyn[w * TILE] is word w
    u32* hot;  // &g_tile[0][lane]
    FJSP_ORDER_SOURCE_MEMBERS
    __device__ __forceinline__ u32 ld(int w) const { return dyn[w * TILE]; }
    __device__ __forceinline__ void st(int w, u32 v) { dyn[w * TILE] = v; }
    __device__ __forceinline__ u32 ld_hot(int w) const { return hot[w * TILE]; }
    __device__ __forceinline__ void st_hot(int w, u32 v) { hot[w * TILE] = v; }
    __device__ __forceinline__ u32 or_word(int w, u32 v) {  // one thread owns the env: plain read-modify-write
        const u32 old = dyn[w * TILE];
        dyn[w * TILE] = old | v;
        return old;
    }
};
using TileColumn = TileColumnT<false>;
// The same column when the K cells of an env run on K threads (cell-parallel step): order words are shared.
template <bool LONG_ = false>
struct TileColumnSharedT : TileColumnT<LONG_> {
    __device__ __forceinline__ u32 or_word(int w, u32 v) { return atomicOr(&this->dyn[w * TILE], v); }
};
using TileColumnShared = TileColumnSharedT<false>;
// Per-env exchange area of the cell-parallel step (fjsp_core.h "X slots"): column `lane` of a [Xl<K>::WORDS][TILE] block.
struct XchgColumn {
    u32* x;  // &s_x[0][lane]
    __device__ __forceinline__ u32 ld(int i) const { return x[i * TILE]; }
    __device__ __forceinline__ void st(int i, u32 v) { x[i * TILE] = v; }
    __device__ __forceinline__ void atom_or(int i, u32 v) { atomicOr(&x[i * TILE], v); }
    __device__ __forceinline__ void atom_add(int i, u32 v) { atomicAdd(&x[i * TILE], v); }
    // u16 table: entry i16 lives in half (i16 & 1) of word i16 >> 1
    __device__ __forceinline__ void st16(int i16, u32 v) {
        reinterpret_cast<unsigned short*>(x + (i16 >> 1) * TILE)[i16 & 1] = (unsigned short)v;
    }
    __device__ __forceinline__ u32 ld16(int i16) const { return reinterpret_cast<const unsigned short*>(x + (i16 >> 1) * TILE)[i16 & 1]; }
};
// Whole column addressed directly in HBM (reset / export paths, not hot).
template <bool LONG_ = false>
struct GmemColumnT {
    static constexpr bool LONG = LONG_;
    u32* base;
    FJSP_ORDER_SOURCE_MEMBERS
    __device__ __forceinline__ u32 ld(int w) const { return base[w * TILE]; }
    __device__ __forceinline__ void st(int w, u32 v) { base[w * TILE] = v; }
    __device__ __forceinline__ u32 ld_hot(int w) const { return base[w * TILE]; }
    __device__ __forceinline__ void st_hot(int w, u32 v) { base[w * TILE] = v; }
    __device__ __forceinline__ u32 or_word(int w, u32 v) {
        const u32 old = base[w * TILE];
        base[w * TILE] = old | v;
        return old;
    }
};

// ---- PTX wrappers (mbarrier + bulk async copy; see /opt/skills/guides/blackwell_cuda_programming.md) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
// L2 prefetch of a whole tile that a later CTA of this SM will load (takes the HBM latency off that CTA's start-up)
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
// Programmatic dependent launch (small batches, where a step is launch-latency bound): a kernel launched with the
// programmatic-serialization attribute may become resident while its predecessor in the stream still runs; it must not
// touch global memory before pdl_wait() (= the predecessor has completed and its writes are visible).  Both are no-ops
// for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Warp-cooperative auto-reset.  FJSP_MAX_ORDERS == 32 == warp size: for every lane whose episode just ended (ballot),
// the 32 lanes draw that env's 32 Philox orders in parallel (one order per lane, counter = (global env, episode, lane))
// and store them into the ending lane's shared-memory column; the ending lane itself re-initialises its scalars.
// A per-lane reset would make the whole warp wait for 32 sequential Philox calls whenever ANY of its envs ends.
// Must be called by all 32 lanes of the warp.  Returns true for lanes that were reset.
template <int K, class S>
__device__ __forceinline__ bool warp_autoreset(S s, u32* s_dyn, int tid, bool do_reset, u32 cur_episode, int num_orders, uint64_t seed,
                                               uint64_t genv_lane0) {
    const unsigned need = __ballot_sync(0xffffffffu, do_reset);
    if (need == 0u) return false;
    const int lane = tid & 31, wbase = tid & ~31;
    u32 episode = 0u;
    if (do_reset) {
        episode = cur_episode + 1u;
        reset_env_base<K>(s, num_orders, episode);  // hot words go to the HBM tile; the caller reloads its registers
    }
    if (!S::LONG) {  // (the long layout's ring fills as the pickup station pops orders: nothing to draw at reset)
        unsigned rem = need;
        while (rem) {
            const int src = __ffs((int)rem) - 1;
            rem &= rem - 1u;
            const u32 ep = __shfl_sync(0xffffffffu, episode, src);
            const u32 ow = lane < num_orders ? philox_order(seed, genv_lane0 + (uint64_t)src, ep, lane) : 0u;
            s_dyn[(W_ORDER - W_CSTEP + lane) * TILE + wbase + src] = ow;
        }
    }
    __syncwarp();
    return do_reset;
}

struct StepArgs {
    u32* state;              // tiles
    const uint8_t* actions;  // [N][ACT]
    float* obs;              // [N][OBS]
    int8_t* masks;           // [N][MASK]
    float* rewards;          // [N][ACT]
    uint8_t* flags;          // [N][4]
    uint8_t* results;        // [N][ACT] or null
    int32_t* infos;          // [N][4] or null
    u32* wire;               // [N][Wire<K>::WORDS]: the WIRE instantiation writes this instead of obs/masks/rewards/flags
    int64_t num_envs, first_env;
    int64_t tile_begin;      // first tile of this launch (host-buffer path steps the batch in pipelined chunks)
    uint64_t seed;
    int32_t num_orders, autoreset;
    int32_t prefetch_tiles;  // cell-parallel kernel: L2-prefetch the tile this many CTAs ahead (0 = off)
    int32_t prefetch_tiles_env;  // the same for the thread-per-env kernel
    const FjspOrderRec* otab;    // long layout: explicit order tables [N][otab_stride], or null = Philox order stream
    int32_t otab_stride;
    u32* rq;                     // long layout: ready FIFOs [N][READY_FIFO_WORDS]
};

// ---------------------------------------------------------------------------------------------
// Reset: FJSPSimulation.reset for the masked envs.  Thread per env, state addressed in place.
// ---------------------------------------------------------------------------------------------
template <int K, bool LONG>
__global__ void __launch_bounds__(TILE) fjsp_reset_kernel(const __grid_constant__ Params P, u32* state, const uint8_t* env_mask,
                                                           const FjspOrderRec* orders, int orders_stride, int num_orders, uint64_t seed,
                                                           int64_t num_envs, int64_t first_env, float* obs, int8_t* masks) {
    const int64_t env = (int64_t)blockIdx.x * TILE + threadIdx.x;
    const bool pad = env >= num_envs;
    if (!pad && env_mask && env_mask[env] == 0) return;
    GmemColumnT<LONG> s{state + (int64_t)blockIdx.x * Geo<K, LONG>::TILE_WORDS + threadIdx.x};
    s.seed = seed, s.genv = (uint64_t)(first_env + env), s.otab = (pad || !orders) ? nullptr : orders + env * orders_stride;
    reset_env<K>(s, P, pad ? 0 : num_orders, s.otab, seed, (uint64_t)(first_env + env), 0u);
    if (pad) return;
    if (obs && masks) {
        float o[Lay<K>::OBS];
        u32 mw[Lay<K>::MASK / 4];
        observe_env<K>(s, P, o, mw);
#pragma unroll
        for (int i = 0; i < Lay<K>::OBS; i++) obs[env * Lay<K>::OBS + i] = o[i];
        uint4* m4 = reinterpret_cast<uint4*>(masks + env * Lay<K>::MASK);
#pragma unroll
        for (int i = 0; i < Lay<K>::MASK / 16; i++) m4[i] = make_uint4(mw[4 * i], mw[4 * i + 1], mw[4 * i + 2], mw[4 * i + 3]);
    }
}

// rows of `src` ([n][width]) whose env is selected by `mask` replace the same rows of `dst` (masked reset of explicit order tables)
__global__ void fjsp_copy_masked_rows_kernel(FjspOrderRec* dst, const FjspOrderRec* src, const uint8_t* mask, int64_t n, int width) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * width && mask[i / width]) dst[i] = src[i];
}

template <int K>
__device__ __forceinline__ void load_actions(const uint8_t* actions, int64_t env, bool valid, int* a) {
    const u32* src = reinterpret_cast<const u32*>(actions + env * Lay<K>::ACT);
#pragma unroll
    for (int i = 0; i < Lay<K>::ACT / 4; i++) {
        const u32 v = valid ? __ldg(src + i) : 0u;
#pragma unroll
        for (int j = 0; j < 4; j++) a[4 * i + j] = (int)((v >> (8 * j)) & 0xffu);
    }
}

// ---------------------------------------------------------------------------------------------
// Step: one CTA = one tile of 64 envs; one launch = one lockstep step of all envs.
// ---------------------------------------------------------------------------------------------
template <int K, bool WIRE, bool LONG = false>
__global__ void __launch_bounds__(TILE) fjsp_step_kernel(const __grid_constant__ Params P, const StepArgs A) {
    using G = Geo<K, LONG>;
    constexpr int OUT_ROW_BYTES = WIRE ? G::WIRE_ROW_BYTES : G::OBS_ROW_BYTES;   // staged row per env
    constexpr int OUT_TILE_BYTES = OUT_ROW_BYTES * TILE;
    constexpr int MODE = WIRE ? OBS_WIRE : OBS_FLOAT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32* s_dyn = reinterpret_cast<u32*>(smem_raw);
    u32* s_out = reinterpret_cast<u32*>(smem_raw + G::DYN_BYTES);  // float observations, or wire rows
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + G::DYN_BYTES + OUT_TILE_BYTES);

    const int tid = threadIdx.x;
    const int64_t tile = A.tile_begin + blockIdx.x;
    const int64_t env = tile * TILE + tid;
    const bool valid = env < A.num_envs;
    u32* g_tile = A.state + tile * G::TILE_WORDS;

    pdl_launch_dependents();  // the next step's CTAs may take their places now; they wait below before touching memory
    if (tid == 0) mbar_init(bar, 1);
    pdl_wait();
    if (tid == 0) {  // ONE bulk async copy (TMA engine) for the dynamically indexed words of the tile
        mbar_expect_tx(bar, G::DYN_BYTES);
        bulk_g2s(s_dyn, g_tile + G::DYN0 * TILE, G::DYN_BYTES, bar);
    }
    if (tid == 32 && A.prefetch_tiles_env > 0) {  // L2 prefetch of the tile a later CTA of this SM will load
        const int64_t nt = (int64_t)blockIdx.x + A.prefetch_tiles_env;
        if (nt < (int64_t)gridDim.x) bulk_prefetch_l2(g_tile + (int64_t)A.prefetch_tiles_env * G::TILE_WORDS, G::TILE_BYTES);
    }
    // meanwhile: the hot words of the pickup station and of cell 0 (coalesced 32-bit loads, straight into registers)
    // and the action bytes
    TileColumnT<LONG> s{s_dyn + tid - G::DYN0 * TILE, g_tile + tid};
    s.seed = A.seed, s.genv = (uint64_t)(A.first_env + env), s.otab = (A.otab && valid) ? A.otab + env * A.otab_stride : nullptr;
    s.rq = (LONG && valid) ? A.rq + env * READY_FIFO_WORDS : nullptr;
    Hot h;
    HotCell c0;
    load_hot(s, h);
    load_cell<K>(s, 0, c0);
    int a[Lay<K>::ACT];
    load_actions<K>(A.actions, env, valid, a);
    __syncthreads();  // the mbarrier is initialised for everyone
    mbar_wait(bar, 0);

    // padding lanes of a ragged last tile are inert: their words travel through unchanged
    StepOut<K> out;
    out.obs = reinterpret_cast<float*>(s_out) + tid * Lay<K>::OBS;
    out.flags = 0u;
    if (valid) step_env_hot<K, MODE>(s, P, h, c0, a, out);
    const bool ended = valid && A.autoreset && (out.flags & 0x00ffffffu);
    if (ended && K > 1) store_cell<K>(s, 0, c0);  // (kept simple: the reset below rewrites every hot word anyway)
    if (warp_autoreset<K>(s, s_dyn, tid, ended, h.episode, A.num_orders, A.seed, (uint64_t)(A.first_env + env - (tid & 31)))) {
        load_hot(s, h);
        load_cell<K>(s, 0, c0);
        observe_out<K, MODE>(s, P, h, c0, out);  // the observation returned with an ended episode is the new episode's first
        out.flags |= 1u << 24;
    }
    store_hot(s, h);
    store_cell<K>(s, 0, c0);
    if (valid) {
        if (WIRE) {
            u32 row[Wire<K>::WORDS];
            wire_row<K>(out, row);
            uint2* dst = reinterpret_cast<uint2*>(s_out + tid * Wire<K>::WORDS);
#pragma unroll
            for (int i = 0; i < Wire<K>::WORDS / 2; i++) dst[i] = make_uint2(row[2 * i], row[2 * i + 1]);
        } else {
            uint4* m4 = reinterpret_cast<uint4*>(A.masks + env * Lay<K>::MASK);
#pragma unroll
            for (int i = 0; i < Lay<K>::MASK / 16; i++)
                m4[i] = make_uint4(out.mask[4 * i], out.mask[4 * i + 1], out.mask[4 * i + 2], out.mask[4 * i + 3]);
            float4* r4 = reinterpret_cast<float4*>(A.rewards + env * Lay<K>::ACT);
#pragma unroll
            for (int i = 0; i < Lay<K>::ACT / 4; i++)
                r4[i] = make_float4(out.reward[4 * i], out.reward[4 * i + 1], out.reward[4 * i + 2], out.reward[4 * i + 3]);
            reinterpret_cast<u32*>(A.flags)[env] = out.flags;
        }
        if (A.results) {
            u32* rs = reinterpret_cast<u32*>(A.results + env * Lay<K>::ACT);
#pragma unroll
            for (int i = 0; i < Lay<K>::ACT / 4; i++) rs[i] = out.results[i];
        }
        if (A.infos) reinterpret_cast<int4*>(A.infos)[env] = make_int4(out.info[0], out.info[1], out.info[2], out.info[3]);
    }
    // generic-proxy writes to shared memory must be visible to the async proxy before the bulk stores
    fence_async_smem();
    __syncthreads();
    const int64_t remaining = A.num_envs - tile * TILE;
    const int nvalid = remaining >= TILE ? TILE : (int)remaining;
    u32* g_out = WIRE ? A.wire + tile * TILE * Wire<K>::WORDS : reinterpret_cast<u32*>(A.obs + tile * TILE * Lay<K>::OBS);
    const bool out_bulk = ((nvalid * OUT_ROW_BYTES) & 15) == 0 && ((reinterpret_cast<uintptr_t>(g_out) & 15) == 0);
    if (tid == 0) {
        bulk_s2g(g_tile + G::DYN0 * TILE, s_dyn, G::DYN_BYTES);
        if (out_bulk) bulk_s2g(g_out, s_out, (uint32_t)(nvalid * OUT_ROW_BYTES));
        bulk_commit();
    }
    if (!out_bulk) {  // ragged last tile or unaligned caller buffer: cooperative coalesced 32-bit stores
        for (int i = tid; i < nvalid * (OUT_ROW_BYTES / 4); i += TILE) g_out[i] = s_out[i];
    }
    if (tid == 0) bulk_wait_read0();  // shared memory must stay allocated until the TMA engine has read it
}

// ---------------------------------------------------------------------------------------------
// Cell-parallel step (scaled shop, K >= 2): one CTA = one tile of 64 envs x K cells = 64K threads; thread (c, e) runs
// cell c of env e (fjsp_core.h "cell-parallel step").  Warps are uniform in c, lanes of a warp are 32 consecutive envs,
// so every shared-memory access stays on the conflict-free column layout and the hot words of cell c are coalesced
// 32-bit loads.  The thread-per-env kernel above keeps 4 warps per SM resident at K = 4 (the tile's 76 KB of tray pools
// set the limit) and is latency-bound at 0.27 of the HBM roofline; this one keeps 16.
// ---------------------------------------------------------------------------------------------
template <int K, bool WIRE, bool LONG = false>
__global__ void __launch_bounds__(TILE* K, Geo<K, LONG>::CELLS_CTAS_PER_SM) fjsp_step_cells_kernel(const __grid_constant__ Params P, const StepArgs A) {
    using G = Geo<K, LONG>;
    constexpr int NT = TILE * K;
    constexpr int OUT_ROW_BYTES = WIRE ? G::WIRE_ROW_BYTES : G::OBS_ROW_BYTES;
    constexpr int OUT_TILE_BYTES = OUT_ROW_BYTES * TILE;
    constexpr int AG = Lay<K>::AGENTS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32* s_dyn = reinterpret_cast<u32*>(smem_raw);
    u32* s_out = reinterpret_cast<u32*>(smem_raw + G::DYN_BYTES);
    u32* s_x = reinterpret_cast<u32*>(smem_raw + G::DYN_BYTES + OUT_TILE_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + G::DYN_BYTES + OUT_TILE_BYTES + G::X_BYTES);

    const int tid = threadIdx.x;
    const int c = tid / TILE, e = tid % TILE;  // c is uniform over a warp
    const int64_t tile = A.tile_begin + blockIdx.x;
    const int64_t env = tile * TILE + e;
    const bool valid = env < A.num_envs;
    u32* g_tile = A.state + tile * G::TILE_WORDS;

    pdl_launch_dependents();
    if (tid == 0) mbar_init(bar, 1);
    pdl_wait();
    if (tid == 0) {  // the tile's bulk copy is under way before anything else happens in the CTA
        mbar_expect_tx(bar, G::DYN_BYTES);
        bulk_g2s(s_dyn, g_tile + G::DYN0 * TILE, G::DYN_BYTES, bar);
    }
    if (tid == 32 && A.prefetch_tiles > 0) {
        // the tile this SM will work on about one CTA generation from now goes to L2 meanwhile: its hot-word loads and
        // its bulk copy then start from L2 instead of HBM (with 2 CTAs per SM the start-up latency is not hidden otherwise)
        const int64_t nt = (int64_t)blockIdx.x + A.prefetch_tiles;
        if (nt < (int64_t)gridDim.x) bulk_prefetch_l2(g_tile + (int64_t)A.prefetch_tiles * G::TILE_WORDS, G::TILE_BYTES);
    }
    TileColumnSharedT<LONG> s;
    s.dyn = s_dyn + e - G::DYN0 * TILE, s.hot = g_tile + e;
    s.seed = A.seed, s.genv = (uint64_t)(A.first_env + env), s.otab = (A.otab && valid) ? A.otab + env * A.otab_stride : nullptr;
    s.rq = (LONG && valid) ? A.rq + env * READY_FIFO_WORDS : nullptr;
    XchgColumn x{s_x + e};
    CellLane L;
    L.c = c;
    load_hot(s, L.h);            // every lane of an env mirrors the pickup station's words
    load_cell<K>(s, c, L.hc);    // its own cell's 20 hot words (coalesced: a warp reads 128 consecutive bytes per word)
    int a0 = 0, a7[7];
    {
        const uint8_t* arow = A.actions + env * Lay<K>::ACT;
        if (valid) a0 = __ldg(arow);
#pragma unroll
        for (int i = 0; i < 7; i++) a7[i] = valid ? (int)__ldg(arow + 1 + 7 * c + i) : 0;
    }
    for (int i = tid; i < Xl<K, LONG>::WORDS * TILE; i += NT) s_x[i] = 0u;
    __syncthreads();  // mbarrier initialised and exchange area zeroed for everyone
    mbar_wait(bar, 0);

    if (valid) cells_begin<K>(s, x, P, L, a0, a7);
    __syncthreads();
    if (valid) cells_act_run<K>(s, x, P, L, a7);
    __syncthreads();
    int32_t info[4] = {0, 0, 0, 0};
    L.flags = 0u, L.g = 0;
    if (valid) cells_finish<K, TileColumnSharedT<LONG>>(x, P, L, info);
    const bool ended = valid && A.autoreset && (L.flags & 0x00ffffffu);
    if (__syncthreads_or(ended)) {
        // the two warps of cell 0 reset the ended envs (all hot words of all cells go to the HBM tile, pools and orders
        // to shared memory); every lane of a reset env then reloads its registers
        if (c == 0) warp_autoreset<K>(s, s_dyn, tid, ended, L.h.episode, A.num_orders, A.seed, (uint64_t)(A.first_env + env - (tid & 31)));
        __syncthreads();
        if (ended) {
            load_hot(s, L.h);
            load_cell<K>(s, c, L.hc);
            L.flags |= 1u << 24;
        }
    }
    if (valid) {
        if (WIRE) {  // the lane's five words of the row (+ the two shared ones on the lane of cell 0): plain stores
            u32 w5[5], s2[2];
            cells_observe_wire<K>(s, P, L, s2, w5);
            u32* row = s_out + e * Wire<K>::WORDS;
#pragma unroll
            for (int i = 0; i < 5; i++) row[2 + 5 * c + i] = w5[i];
            if (c == 0) {
                row[0] = s2[0], row[1] = s2[1];
                if (Wire<K>::WORDS > Wire<K>::USED) row[Wire<K>::WORDS - 1] = 0u;
            }
        } else {
            float* orow = reinterpret_cast<float*>(s_out) + e * Lay<K>::OBS;
            cells_observe<K>(s, x, P, L, FloatSink{orow, P}, FloatSink{orow + 7 + 31 * c, P});
        }
    }
    if (c == 0) store_hot(s, L.h);
    store_cell<K>(s, c, L.hc);
    __syncthreads();
    // ---- output rows, one 32-byte piece per lane: mask bytes [32c, 32c+32), action columns [8c, 8c+8)
    if (valid) {
        const u32 mbits = cells_mask_word<K>(x, c);
        u32 v16[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v16[i] = x.ld16(2 * Xl<K>::LOCAL + 8 * c + i);
        if (!WIRE) {
            uint4* m4 = reinterpret_cast<uint4*>(A.masks + env * Lay<K>::MASK + 32 * c);
            m4[0] = make_uint4(nibble_bytes(mbits), nibble_bytes(mbits >> 4), nibble_bytes(mbits >> 8), nibble_bytes(mbits >> 12));
            m4[1] = make_uint4(nibble_bytes(mbits >> 16), nibble_bytes(mbits >> 20), nibble_bytes(mbits >> 24), nibble_bytes(mbits >> 28));
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; i++) r[i] = (8 * c + i < AG) ? (float)(L.g + AG * x_local10(v16[i])) / (float)(10 * AG) : 0.0f;
            float4* r4 = reinterpret_cast<float4*>(A.rewards + env * Lay<K>::ACT + 8 * c);
            r4[0] = make_float4(r[0], r[1], r[2], r[3]);
            r4[1] = make_float4(r[4], r[5], r[6], r[7]);
            if (c == 0) reinterpret_cast<u32*>(A.flags)[env] = L.flags;
        }
        if (A.results) {
            u32 lo = 0u, hi = 0u;
#pragma unroll
            for (int i = 0; i < 4; i++) lo |= x_result(v16[i]) << (8 * i), hi |= x_result(v16[4 + i]) << (8 * i);
            *reinterpret_cast<uint2*>(A.results + env * Lay<K>::ACT + 8 * c) = make_uint2(lo, hi);
        }
        if (A.infos && c == 0) reinterpret_cast<int4*>(A.infos)[env] = make_int4(info[0], info[1], info[2], info[3]);
    }
    fence_async_smem();
    __syncthreads();
    const int64_t remaining = A.num_envs - tile * TILE;
    const int nvalid = remaining >= TILE ? TILE : (int)remaining;
    u32* g_out = WIRE ? A.wire + tile * TILE * Wire<K>::WORDS : reinterpret_cast<u32*>(A.obs + tile * TILE * Lay<K>::OBS);
    const bool out_bulk = ((nvalid * OUT_ROW_BYTES) & 15) == 0 && ((reinterpret_cast<uintptr_t>(g_out) & 15) == 0);
    if (tid == 0) {
        bulk_s2g(g_tile + G::DYN0 * TILE, s_dyn, G::DYN_BYTES);
        if (out_bulk) bulk_s2g(g_out, s_out, (uint32_t)(nvalid * OUT_ROW_BYTES));
        bulk_commit();
    }
    if (!out_bulk) {
        for (int i = tid; i < nvalid * (OUT_ROW_BYTES / 4); i += NT) g_out[i] = s_out[i];
    }
    if (tid == 0) bulk_wait_read0();
}

// ---------------------------------------------------------------------------------------------
// Uniform-random policy stand-in (BASELINE.json configs 1/2/4): a_i ~ U{0..n_i-1} from Philox.
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void fjsp_random_actions_kernel(uint8_t* actions, int64_t num_envs, int64_t first_env, uint64_t seed, uint64_t t) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();  // (the buffer may still be read by the step before)
    if (env >= num_envs) return;
    int a[Lay<K>::ACT];
#pragma unroll
    for (int i = 0; i < Lay<K>::ACT; i++) a[i] = 0;
    philox_actions_k<K>(seed, (uint64_t)(first_env + env), t, a);
    u32* dst = reinterpret_cast<u32*>(actions + env * Lay<K>::ACT);
#pragma unroll
    for (int i = 0; i < Lay<K>::ACT / 4; i++)
        dst[i] = (u32)a[4 * i] | ((u32)a[4 * i + 1] << 8) | ((u32)a[4 * i + 2] << 16) | ((u32)a[4 * i + 3] << 24);
}

// ---------------------------------------------------------------------------------------------
// Rollout: `steps` lockstep steps per launch with in-kernel Philox actions and auto-reset; the tile stays in
// shared memory between steps, so HBM sees the state once per `steps` steps.  Accumulates exact integer stats.
// ---------------------------------------------------------------------------------------------
template <int K, bool LONG = false>
__global__ void __launch_bounds__(TILE) fjsp_rollout_kernel(const __grid_constant__ Params P, u32* state, int64_t num_envs,
                                                             int64_t first_env, uint64_t seed, uint64_t t0, int steps,
                                                             int num_orders, unsigned long long* stats, u32* rq) {
    using G = Geo<K, LONG>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32* s_dyn = reinterpret_cast<u32*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + G::DYN_BYTES);
    __shared__ unsigned long long s_acc[6];
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t env = tile * TILE + tid;
    const bool valid = env < num_envs;
    u32* g_tile = state + tile * G::TILE_WORDS;
    if (tid == 0) mbar_init(bar, 1);
    if (tid < 6) s_acc[tid] = 0ull;
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, G::DYN_BYTES);
        bulk_g2s(s_dyn, g_tile + G::DYN0 * TILE, G::DYN_BYTES, bar);
    }
    TileColumnT<LONG> s{s_dyn + tid - G::DYN0 * TILE, g_tile + tid};
    s.seed = seed, s.genv = (uint64_t)(first_env + env);
    s.rq = (LONG && valid) ? rq + env * READY_FIFO_WORDS : nullptr;
    Hot h;  // hot words of the pickup station and of cell 0 live in registers for the whole launch
    HotCell c0;
    load_hot(s, h);
    load_cell<K>(s, 0, c0);
    mbar_wait(bar, 0);
    unsigned long long n_steps = 0, n_eps = 0, n_orders = 0, n_prod = 0, n_fault = 0;
    long long units = 0;
    const uint64_t genv = (uint64_t)(first_env + env);
    for (int k = 0; k < steps; k++) {  // uniform trip count: all lanes stay together for the cooperative reset
        bool ended = false;
        if (valid) {
            int a[Lay<K>::ACT];
            philox_actions_k<K>(seed, genv, t0 + (uint64_t)k, a);
            StepOut<K> out;
            out.obs = nullptr;
            const int before_o = h.completed_orders, before_p = h.total_packaged;
            step_env_hot<K, OBS_NONE>(s, P, h, c0, a, out);
            n_steps += 1;
            n_orders += (unsigned long long)(h.completed_orders - before_o);
            n_prod += (unsigned long long)(h.total_packaged - before_p);
            units += out.reward_units;
            if (out.flags & 0x00ffffffu) {
                ended = true;
                n_eps += 1;
                n_fault += (out.flags >> 16) & 0xffu ? 1 : 0;
            }
        }
        if (warp_autoreset<K>(s, s_dyn, tid, ended, h.episode, num_orders, seed, genv - (uint64_t)(tid & 31))) {
            load_hot(s, h);
            load_cell<K>(s, 0, c0);
        }
    }
    store_hot(s, h);
    store_cell<K>(s, 0, c0);
    // warp shuffle reduce, then one shared atomic per warp, one global atomic per CTA and counter
    unsigned long long v[6] = {n_steps, n_eps, n_orders, n_prod, n_fault, (unsigned long long)units};
#pragma unroll
    for (int j = 0; j < 6; j++) {
        unsigned long long x = v[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if ((tid & 31) == 0) atomicAdd(&s_acc[j], x);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(g_tile + G::DYN0 * TILE, s_dyn, G::DYN_BYTES);
        bulk_commit();
    }
    if (tid < 6) atomicAdd(&stats[tid], s_acc[tid]);
    if (tid == 0) bulk_wait_read0();
}

// ---------------------------------------------------------------------------------------------
// Rollout, cell-parallel (K >= 2): the K-steps-per-launch kernel with one thread per (env, cell).  The tile and the
// exchange area stay in shared memory for the whole launch, every lane keeps its cell's hot words (and its mirror of
// the pickup station's) in registers; per step: Philox actions for the lane's own agents, the three phases of the
// cell-parallel step (no observation), the cooperative reset by the two warps of cell 0, and a fresh exchange area.
// ---------------------------------------------------------------------------------------------
template <int K, bool LONG = false>
__global__ void __launch_bounds__(TILE* K, Geo<K, LONG>::CELLS_CTAS_PER_SM) fjsp_rollout_cells_kernel(const __grid_constant__ Params P, u32* state,
                                                                                                int64_t num_envs, int64_t first_env,
                                                                                                uint64_t seed, uint64_t t0, int steps,
                                                                                                int num_orders, unsigned long long* stats, u32* rq) {
    using G = Geo<K, LONG>;
    constexpr int NT = TILE * K;
    constexpr int AG = Lay<K>::AGENTS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32* s_dyn = reinterpret_cast<u32*>(smem_raw);
    u32* s_x = reinterpret_cast<u32*>(smem_raw + G::DYN_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + G::DYN_BYTES + G::X_BYTES);
    __shared__ unsigned long long s_acc[6];
    const int tid = threadIdx.x;
    const int c = tid / TILE, e = tid % TILE;
    const int64_t tile = blockIdx.x;
    const int64_t env = tile * TILE + e;
    const bool valid = env < num_envs;
    u32* g_tile = state + tile * G::TILE_WORDS;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, G::DYN_BYTES);
        bulk_g2s(s_dyn, g_tile + G::DYN0 * TILE, G::DYN_BYTES, bar);
    }
    if (tid < 6) s_acc[tid] = 0ull;
    TileColumnSharedT<LONG> s;
    s.dyn = s_dyn + e - G::DYN0 * TILE, s.hot = g_tile + e;
    s.seed = seed, s.genv = (uint64_t)(first_env + env);
    s.rq = (LONG && valid) ? rq + env * READY_FIFO_WORDS : nullptr;
    XchgColumn x{s_x + e};
    CellLane L;
    L.c = c;
    load_hot(s, L.h);
    load_cell<K>(s, c, L.hc);
    for (int i = tid; i < Xl<K, LONG>::WORDS * TILE; i += NT) s_x[i] = 0u;
    __syncthreads();
    mbar_wait(bar, 0);
    unsigned long long n_steps = 0, n_eps = 0, n_orders = 0, n_prod = 0, n_fault = 0;
    long long units = 0;
    const uint64_t genv = (uint64_t)(first_env + env);
    for (int k = 0; k < steps; k++) {
        const uint64_t t = t0 + (uint64_t)k;
        int a0 = 0, a7[7];
        {   // the lane's own action columns of the Philox stream (philox_actions_k): cell c draws with counter word 3 = 1 + 16c
            u32 r[4];
            philox4x32_10((u32)genv, (u32)t, (u32)(t >> 32), 1u + 16u * (u32)c, (u32)seed, (u32)(seed >> 32), r);
#pragma unroll
            for (int j = 1; j < 8; j++) {
                const u32 hw = (j & 1) ? (r[j >> 1] >> 16) : (r[j >> 1] & 0xffffu);
                a7[j - 1] = (int)((hw * (j == 1 ? 8u : 3u)) >> 16);
            }
            if (c != 0) philox4x32_10((u32)genv, (u32)t, (u32)(t >> 32), 1u, (u32)seed, (u32)(seed >> 32), r);
            a0 = (int)(((r[0] & 0xffffu) * 3u) >> 16);
        }
        if (valid) cells_begin<K>(s, x, P, L, a0, a7);
        __syncthreads();
        if (valid) cells_act_run<K>(s, x, P, L, a7);
        __syncthreads();
        bool ended = false;
        if (valid) {
            int32_t info[4];
            cells_finish<K, TileColumnSharedT<LONG>>(x, P, L, info);
#pragma unroll
            for (int i = 1; i < 8; i++) units += L.g + AG * L.local10[i];
            if (c == 0) {
                units += L.g + AG * L.local10[0];
                n_steps += 1;
                n_orders += (unsigned long long)(L.h.completed_orders - L.orders_in);
                n_prod += (unsigned long long)(L.h.total_packaged - L.packaged_in);
            }
            if (L.flags & 0x00ffffffu) {
                ended = true;
                if (c == 0) n_eps += 1, n_fault += ((L.flags >> 16) & 0xffu) ? 1 : 0;
            }
        }
        if (__syncthreads_or(ended)) {  // also: every lane has read the exchange area
            if (c == 0) warp_autoreset<K>(s, s_dyn, tid, ended, L.h.episode, num_orders, seed, genv - (uint64_t)(tid & 31));
            __syncthreads();
            if (ended) {
                load_hot(s, L.h);
                load_cell<K>(s, c, L.hc);
            }
        }
        for (int i = c; i < Xl<K, LONG>::WORDS; i += K) x.st(i, 0u);
        __syncthreads();
    }
    if (c == 0) store_hot(s, L.h);
    store_cell<K>(s, c, L.hc);
    unsigned long long v[6] = {n_steps, n_eps, n_orders, n_prod, n_fault, (unsigned long long)units};
#pragma unroll
    for (int j = 0; j < 6; j++) {
        unsigned long long y = v[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) y += __shfl_xor_sync(0xffffffffu, y, off);
        if ((tid & 31) == 0) atomicAdd(&s_acc[j], y);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(g_tile + G::DYN0 * TILE, s_dyn, G::DYN_BYTES);
        bulk_commit();
    }
    if (tid < 6) atomicAdd(&stats[tid], s_acc[tid]);
    if (tid == 0) bulk_wait_read0();
}

}  // namespace fjsp
