// fjsp_shared.cuh — kernels of the SHARED FLOOR (fjsp_shared.h; include/fjsp_b200.h "shared floor"): A = 2..4 AGVs on one
// set of stations.  Same data path as the reference shop's step kernel (fjsp_kernels.cuh): one CTA = one tile of 64 envs,
// the dynamically indexed words (order table, tray pool: 26 KB) travel by one bulk async copy each way, the hot words —
// 24 of the reference shop + one per further AGV — are coalesced 32-bit loads / stores, the observation rows are staged
// in shared memory and leave with a second bulk copy.  Tile = 132 words x 64 envs = 33,792 B.
// The observation rows (up to 19.7 KB per tile) are staged in the SAME shared memory as the dynamically indexed words: the
// step first takes the handful of words an observation reads (ObsSnapshot), the tile's words leave with their bulk store,
// and once the copy engine has read them the rows are written over them.  26.6 KB per CTA instead of 46.3: 8 CTAs per SM
// instead of 4 (the kernel is latency-bound at 8 warps per SM: 0.65 of the HBM peak measured with separate buffers).
#pragma once

#include "fjsp_kernels.cuh"
#include "fjsp_shared.h"

namespace fjsp {

template <int A>
struct ShGeo {
    static constexpr int TILE_WORDS = ShLay<A>::TOTAL * TILE;
    static constexpr int TILE_BYTES = TILE_WORDS * 4;
    static constexpr int DYN0 = ShLay<A>::DYN0;
    static constexpr int DYN_BYTES = (ShLay<A>::DYN_END - DYN0) * TILE * 4;   // 26,624
    static constexpr int OBS_ROW_BYTES = ShLay<A>::OBS * 4;
    static constexpr int OBS_TILE_BYTES = OBS_ROW_BYTES * TILE;
    static_assert(OBS_TILE_BYTES <= DYN_BYTES, "the observation rows are staged over the dynamically indexed words");
    static constexpr int STEP_SMEM_BYTES = DYN_BYTES + 16;
};

template <int A>
__global__ void __launch_bounds__(TILE) fjsp_shared_reset_kernel(const __grid_constant__ Params P, u32* state, const uint8_t* env_mask,
                                                                  const FjspOrderRec* orders, int num_orders, uint64_t seed,
                                                                  int64_t num_envs, int64_t first_env, float* obs, int8_t* masks) {
    using L = ShLay<A>;
    const int64_t env = (int64_t)blockIdx.x * TILE + threadIdx.x;
    const bool pad = env >= num_envs;
    if (!pad && env_mask && env_mask[env] == 0) return;
    GmemColumnT<false> s{state + (int64_t)blockIdx.x * ShGeo<A>::TILE_WORDS + threadIdx.x};
    shared_reset<A>(s, P, pad ? 0 : num_orders, (pad || !orders) ? nullptr : orders + env * FJSP_MAX_ORDERS, seed, (uint64_t)(first_env + env), 0u);
    if (pad || !obs || !masks) return;
    Hot h;
    HotCell c0;
    u32 ax[3] = {0u, 0u, 0u}, mw[L::MASK / 4];
    float o[L::OBS];
    load_hot(s, h), load_cell<1>(s, 0, c0), shared_load_agvs<A>(s, ax);
    shared_observe<A>(s, P, h, c0, ax, FloatSink{o, P}, mw);
#pragma unroll
    for (int i = 0; i < L::OBS; i++) obs[env * L::OBS + i] = o[i];
    uint4* m4 = reinterpret_cast<uint4*>(masks + env * L::MASK);
#pragma unroll
    for (int i = 0; i < L::MASK / 16; i++) m4[i] = make_uint4(mw[4 * i], mw[4 * i + 1], mw[4 * i + 2], mw[4 * i + 3]);
}

template <int A>
__global__ void __launch_bounds__(TILE) fjsp_shared_step_kernel(const __grid_constant__ Params P, const StepArgs Ar) {
    using L = ShLay<A>;
    using G = ShGeo<A>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32* s_dyn = reinterpret_cast<u32*>(smem_raw);
    u32* s_out = s_dyn;   // (after the tile's words have left)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + G::DYN_BYTES);

    const int tid = threadIdx.x;
    const int64_t tile = Ar.tile_begin + blockIdx.x;
    const int64_t env = tile * TILE + tid;
    const bool valid = env < Ar.num_envs;
    u32* g_tile = Ar.state + tile * G::TILE_WORDS;

    pdl_launch_dependents();
    if (tid == 0) mbar_init(bar, 1);
    pdl_wait();
    if (tid == 0) {
        mbar_expect_tx(bar, G::DYN_BYTES);
        bulk_g2s(s_dyn, g_tile + G::DYN0 * TILE, G::DYN_BYTES, bar);
    }
    if (tid == 32 && Ar.prefetch_tiles_env > 0) {
        const int64_t nt = (int64_t)blockIdx.x + Ar.prefetch_tiles_env;
        if (nt < (int64_t)gridDim.x) bulk_prefetch_l2(g_tile + (int64_t)Ar.prefetch_tiles_env * G::TILE_WORDS, G::TILE_BYTES);
    }
    TileColumnT<false> s{s_dyn + tid - G::DYN0 * TILE, g_tile + tid};
    Hot h;
    HotCell c0;
    u32 ax[3] = {0u, 0u, 0u};
    load_hot(s, h);
    load_cell<1>(s, 0, c0);
    shared_load_agvs<A>(s, ax);
    int a[L::ACT];
    {
        const u32* src = reinterpret_cast<const u32*>(Ar.actions + env * L::ACT);
#pragma unroll
        for (int i = 0; i < L::ACT / 4; i++) {
            const u32 v = valid ? __ldg(src + i) : 0u;
#pragma unroll
            for (int j = 0; j < 4; j++) a[4 * i + j] = (int)((v >> (8 * j)) & 0xffu);
        }
    }
    __syncthreads();
    mbar_wait(bar, 0);

    ShOut<A> out;
    out.obs = reinterpret_cast<float*>(s_out) + tid * L::OBS;
    out.flags = 0u;
    if (valid) shared_step<A, false>(s, P, h, c0, ax, a, out);
    const bool ended = valid && Ar.autoreset && (out.flags & 0x00ffffffu);
    if (warp_autoreset<1>(s, s_dyn, tid, ended, h.episode, Ar.num_orders, Ar.seed, (uint64_t)(Ar.first_env + env - (tid & 31)))) {
        shared_reset_agvs<A>(s);
        load_hot(s, h);
        load_cell<1>(s, 0, c0);
        shared_load_agvs<A>(s, ax);
        out.flags |= 1u << 24;   // the observation returned with an ended episode is the new episode's first
    }
    store_hot(s, h);
    store_cell<1>(s, 0, c0);
    shared_store_agvs<A>(s, ax);
    ObsSnapshot<A> snap;
    snap.take(s, h, c0, ax);
    // the tile's words leave; the rows of the observations take their place once the copy engine has read them
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(g_tile + G::DYN0 * TILE, s_dyn, G::DYN_BYTES);
        bulk_commit();
        bulk_wait_read0();
    }
    __syncthreads();
    if (valid) shared_observe<A>(snap, P, h, c0, ax, FloatSink{out.obs, P}, out.mask);
    if (valid) {
        uint4* m4 = reinterpret_cast<uint4*>(Ar.masks + env * L::MASK);
#pragma unroll
        for (int i = 0; i < L::MASK / 16; i++) m4[i] = make_uint4(out.mask[4 * i], out.mask[4 * i + 1], out.mask[4 * i + 2], out.mask[4 * i + 3]);
        float4* r4 = reinterpret_cast<float4*>(Ar.rewards + env * L::ACT);
#pragma unroll
        for (int i = 0; i < L::ACT / 4; i++) r4[i] = make_float4(out.reward[4 * i], out.reward[4 * i + 1], out.reward[4 * i + 2], out.reward[4 * i + 3]);
        reinterpret_cast<u32*>(Ar.flags)[env] = out.flags;
        if (Ar.results) {
            u32* rs = reinterpret_cast<u32*>(Ar.results + env * L::ACT);
#pragma unroll
            for (int i = 0; i < L::ACT / 4; i++) rs[i] = out.results[i];
        }
        if (Ar.infos) reinterpret_cast<int4*>(Ar.infos)[env] = make_int4(out.info[0], out.info[1], out.info[2], out.info[3]);
    }
    fence_async_smem();
    __syncthreads();
    const int64_t remaining = Ar.num_envs - tile * TILE;
    const int nvalid = remaining >= TILE ? TILE : (int)remaining;
    u32* g_out = reinterpret_cast<u32*>(Ar.obs + tile * TILE * L::OBS);
    const bool out_bulk = ((nvalid * G::OBS_ROW_BYTES) & 15) == 0 && ((reinterpret_cast<uintptr_t>(g_out) & 15) == 0);
    if (tid == 0 && out_bulk) {
        bulk_s2g(g_out, s_out, (uint32_t)(nvalid * G::OBS_ROW_BYTES));
        bulk_commit();
    }
    if (!out_bulk) {
        for (int i = tid; i < nvalid * (G::OBS_ROW_BYTES / 4); i += TILE) g_out[i] = s_out[i];
    }
    if (tid == 0) bulk_wait_read0();
}

template <int A>
__global__ void fjsp_shared_random_actions_kernel(uint8_t* actions, int64_t num_envs, int64_t first_env, uint64_t seed, uint64_t t) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (env >= num_envs) return;
    int a[ShLay<A>::ACT];
    philox_actions_shared<A>(seed, (uint64_t)(first_env + env), t, a);
    u32* dst = reinterpret_cast<u32*>(actions + env * ShLay<A>::ACT);
#pragma unroll
    for (int i = 0; i < ShLay<A>::ACT / 4; i++)
        dst[i] = (u32)a[4 * i] | ((u32)a[4 * i + 1] << 8) | ((u32)a[4 * i + 2] << 16) | ((u32)a[4 * i + 3] << 24);
}

}  // namespace fjsp
