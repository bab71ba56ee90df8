// fjsp_a2c.cuh — fused device ops for the batched A2C trainer (the caller of the hot path).
//   fjsp_policy_sample_kernel : per env row, for the 8 agents: softmax over the agent's logits, multiply by the action
//       mask, renormalise (uniform over valid actions if the masked mass is 0), draw a Categorical sample from a
//       Philox uniform, emit the action byte and log q[a]          (a2c.py:214-247 for one sample at a time)
//   fjsp_gae_kernel           : reverse scan over T of discounted returns and GAE per (env, agent) with the shared
//       critic value; an episode end zeroes the bootstrap         (transition_memory.py:83-105, a2c.py:325-332,357)
// Both replace O(100) elementwise torch launches per rollout step; the GEMMs stay library calls.
#pragma once

#include <cuda_runtime.h>

#include "fjsp_core.h"

namespace fjsp {

__device__ __forceinline__ void agent_segment(int agent, int& off, int& n) {
    // mask-block layout of the env: 3 | 8 | 6 x 3 (env.py MASK_OFFSETS)
    off = agent == 0 ? 0 : agent == 1 ? 3 : 11 + 3 * (agent - 2);
    n = agent == 1 ? 8 : 3;
}

// One thread per (row, agent): 8 consecutive lanes share a row.  (One thread per row kept 4096 rows on 32 SMs with eight
// softmaxes and two Philox calls in one dependent chain: 11 us per rollout step; the same arithmetic per agent, so the
// same actions bit for bit.)
__global__ void __launch_bounds__(128) fjsp_policy_sample_kernel(const float* __restrict__ logits, const int8_t* __restrict__ masks,
                                                                 uint8_t* __restrict__ actions, float* __restrict__ logp,
                                                                 int64_t rows, int64_t first_row, uint64_t seed,
                                                                 const unsigned long long* __restrict__ ctr, uint64_t t_off) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = gid >> 3;
    const int ag = (int)(gid & 7);
    if (row >= rows) return;
    int off, n;
    agent_segment(ag, off, n);
    float z[8];
    u32 mbit[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        z[j] = j < n ? __ldg(logits + row * 32 + off + j) : 0.f;
        mbit[j] = j < n ? (u32)(__ldg(masks + row * 32 + off + j) & 1) : 0u;
    }
    const uint64_t t = (ctr ? (uint64_t)ctr[0] : 0ull) + t_off;
    const uint64_t grow = (uint64_t)(first_row + row);
    u32 r[4];
    philox4x32_10((u32)grow, (u32)t, (u32)(t >> 32), ag < 4 ? 2u : 3u, (u32)seed, (u32)(seed >> 32), r);
    float zmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j < n) zmax = fmaxf(zmax, z[j]);
    float p[8], psum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j < n) p[j] = expf(z[j] - zmax), psum += p[j];
    float msum = 0.f, cnt = 0.f;
    float q[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j < n) {
            const float m = (float)mbit[j];
            p[j] = p[j] / psum;   // softmax (networks.py:33)
            q[j] = p[j] * m;      // probs * mask (a2c.py:221)
            msum += q[j], cnt += m;
            p[j] = m;             // keep the mask for the fallback
        }
    float total = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j < n) {
            q[j] = msum > 0.f ? q[j] / msum : p[j] / fmaxf(cnt, 1.f);  // renormalise / uniform over valid (a2c.py:224-231)
            total += q[j];
        }
    const u32 rbits = r[ag & 3];
    const float u = (float)(rbits >> 8) * (1.0f / 16777216.0f) * total;
    int a = 0, last_valid = 0;
    float cdf = 0.f;
    bool done = false;
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j < n) {
            cdf += q[j];
            if (q[j] > 0.f) {
                last_valid = j;
                if (!done && u < cdf) a = j, done = true;
            }
        }
    if (!done) a = last_valid;  // rounding at the end of the cdf never selects a zero-probability action
    float qa = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j < n && j == a) qa = q[j] / total;
    const float eps = 1.1920929e-07f;  // torch.distributions clamps probs to [eps, 1 - eps]
    actions[row * 8 + ag] = (uint8_t)a;
    if (logp) logp[row * 8 + ag] = logf(fminf(fmaxf(qa, eps), 1.0f - eps));
}

__global__ void fjsp_counter_add_kernel(unsigned long long* ctr, unsigned long long inc) { ctr[0] += inc; }

__global__ void __launch_bounds__(256) fjsp_gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                       const uint8_t* __restrict__ flags, float* __restrict__ returns,
                                                       float* __restrict__ adv, int T, int64_t N, float gamma, float lamb) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // env * 8 + agent
    if (idx >= N * 8) return;
    const int64_t env = idx >> 3;
    float ret = values[(int64_t)T * N + env], nxt = ret, gae = 0.f;
    for (int t = T - 1; t >= 0; t--) {
        const u32 f = reinterpret_cast<const u32*>(flags)[(int64_t)t * N + env];
        const float nd = (f & 0x00ffffffu) ? 0.f : 1.f;  // terminated | truncated | fault ends the episode
        const float r = rewards[(int64_t)t * N * 8 + idx];
        const float v = values[(int64_t)t * N + env];
        ret = r + gamma * ret * nd;
        const float td = r + gamma * nxt * nd - v;
        gae = td + gamma * lamb * gae * nd;
        returns[(int64_t)t * N * 8 + idx] = ret;
        adv[(int64_t)t * N * 8 + idx] = gae;
        nxt = v;
    }
}


// ---------------------------------------------------------------------------------------------
// Loss gradients of one A2C update, analytically (a2c.py:647-731): per row (one env step of the rollout) and agent
//   actor_loss_i = -mean(adv_n * log q[a]) - entropy_coef * mean(H(p)),   p = softmax(z), q = p*mask renormalised
//     d log q[a] / dz_j = [a == j] - q_j        (0 when Categorical's probability clamp [eps, 1-eps] is active, or when
//                                                 the masked mass is 0: the uniform fallback does not depend on z)
//     dH / dz_j         = p_j * (g_j - sum_k g_k p_k),   g_k = -(log(p_k + 1e-10) + p_k / (p_k + 1e-10))   (unmasked p)
//   critic_loss  = mean over (row, agent) of (v - return_i)^2   ->   dL/dv = 2/(8B) * sum_i (v - return_i)
// Writes dlogits[B][32] (same layout as the logits; pad columns 0) and dvalue[B]; accumulates into sums[64]:
//   [0..31] column sums of dlogits (the last layers' bias gradients), [32..39] sum of -adv_n*logq per agent,
//   [40..47] sum of H per agent, [48] sum of squared critic errors, [49] sum of dvalue.
// ---------------------------------------------------------------------------------------------
enum { LS_DB = 0, LS_ACTOR = 32, LS_ENT = 40, LS_CRITIC = 48, LS_DV = 49, LS_WORDS = 64 };

__global__ void __launch_bounds__(128) fjsp_a2c_loss_grad_kernel(const float* __restrict__ logits, const int8_t* __restrict__ masks,
                                                                 const uint8_t* __restrict__ actions, const float* __restrict__ adv,
                                                                 const float* __restrict__ returns, const float* __restrict__ values,
                                                                 const float* __restrict__ adv_mean, const float* __restrict__ adv_rstd,
                                                                 float entropy_coef, float inv_b, float* __restrict__ dlogits,
                                                                 float* __restrict__ dvalue, float* __restrict__ sums, int64_t rows) {
    __shared__ float s_sums[LS_WORDS];
    if (threadIdx.x < LS_WORDS) s_sums[threadIdx.x] = 0.f;
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = row < rows;
    float dz[32], acc[18];  // acc: 8 actor terms, 8 entropies, squared error, dvalue
#pragma unroll
    for (int i = 0; i < 32; i++) dz[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 18; i++) acc[i] = 0.f;
    if (live) {
        float z[32];
        u32 mw[8];
        const float4* z4 = reinterpret_cast<const float4*>(logits + row * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 v = __ldg(z4 + i);
            z[4 * i] = v.x, z[4 * i + 1] = v.y, z[4 * i + 2] = v.z, z[4 * i + 3] = v.w;
        }
        const uint4* m4 = reinterpret_cast<const uint4*>(masks + row * 32);
        const uint4 ma = __ldg(m4), mb = __ldg(m4 + 1);
        mw[0] = ma.x, mw[1] = ma.y, mw[2] = ma.z, mw[3] = ma.w, mw[4] = mb.x, mw[5] = mb.y, mw[6] = mb.z, mw[7] = mb.w;
        const uint2 ab = __ldg(reinterpret_cast<const uint2*>(actions) + row);
        const float4* a4 = reinterpret_cast<const float4*>(adv + row * 8);
        const float4* r4 = reinterpret_cast<const float4*>(returns + row * 8);
        const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1), r0 = __ldg(r4), r1 = __ldg(r4 + 1);
        const float advv[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float retv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        const float v = __ldg(values + row);
        float dv = 0.f, sq = 0.f;
#pragma unroll
        for (int ag = 0; ag < 8; ag++) {
            int off, n;
            agent_segment(ag, off, n);
            const int a = (int)(((ag < 4 ? ab.x : ab.y) >> ((ag & 3) * 8)) & 0xffu);
            float zmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < n) zmax = fmaxf(zmax, z[off + j]);
            float p[8], q[8], psum = 0.f, msum = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < n) p[j] = expf(z[off + j] - zmax), psum += p[j];
            float gp = 0.f, ent = 0.f, g[8];
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < n) {
                    const float m = (float)((mw[(off + j) >> 2] >> (((off + j) & 3) * 8)) & 1u);
                    p[j] = p[j] / psum;
                    q[j] = p[j] * m, msum += q[j];
                    const float lg = logf(p[j] + 1e-10f);
                    ent -= p[j] * lg;                               // a2c.py:704-707
                    g[j] = -(lg + p[j] / (p[j] + 1e-10f));
                    gp += g[j] * p[j];
                }
            const float adv_n = (advv[ag] - __ldg(adv_mean + ag)) * __ldg(adv_rstd + ag);  // a2c.py:727-729
            const float eps = 1.1920929e-07f;
            float qa = 0.f;
            bool through = false;
            if (msum > 0.f) {
                const float s = fmaxf(msum, 1e-38f);
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (j < n) {
                        q[j] = q[j] / s;
                        if (j == a) qa = q[j];
                    }
                through = qa >= eps && qa <= 1.0f - eps;       // clamp passes the gradient inside [eps, 1 - eps]
            } else {                                            // uniform over the valid actions (a2c.py:226-231)
                float cnt = 0.f;
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (j < n) cnt += (float)((mw[(off + j) >> 2] >> (((off + j) & 3) * 8)) & 1u);
                qa = (a < n && ((mw[(off + a) >> 2] >> (((off + a) & 3) * 8)) & 1u)) ? 1.0f / fmaxf(cnt, 1.f) : 0.f;
            }
            const float logq = logf(fminf(fmaxf(qa, eps), 1.0f - eps));
            acc[ag] = -adv_n * logq;
            acc[8 + ag] = ent;
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < n) {
                    float d = -entropy_coef * inv_b * p[j] * (g[j] - gp);
                    if (through) d -= adv_n * inv_b * ((j == a ? 1.f : 0.f) - q[j]);
                    dz[off + j] = d;
                }
            const float e = v - retv[ag];
            sq += e * e, dv += e;
        }
        dv *= 2.0f * inv_b * 0.125f;
        acc[16] = sq, acc[17] = dv;
        dvalue[row] = dv;
        float4* o4 = reinterpret_cast<float4*>(dlogits + row * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) o4[i] = make_float4(dz[4 * i], dz[4 * i + 1], dz[4 * i + 2], dz[4 * i + 3]);
    }
    // warp reduce (shuffles), one shared atomic per warp and value, one global atomic per block and value
#pragma unroll
    for (int i = 0; i < 29; i++) {
        float x = dz[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_sums[LS_DB + i], x);
    }
#pragma unroll
    for (int i = 0; i < 18; i++) {
        float x = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_sums[LS_ACTOR + i], x);
    }
    __syncthreads();
    if (threadIdx.x < LS_WORDS && s_sums[threadIdx.x] != 0.f) atomicAdd(sums + threadIdx.x, s_sums[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// Cell views of a scaled shop (K cells): the batched trainer sees every (env, cell) as one row of the reference's own
// 8-agent layout — the pickup station's 7 observation fields followed by the cell's 31, its 3 + 26 mask bytes, the
// pickup station's reward followed by the cell's seven — so the reference's actor/critic networks are shared by the
// cells.  View row = env * K + cell.  The pickup station acts through the row of cell 0; in the other rows its mask
// allows action 0 only (its log-probability is then constant: no gradient).
// ---------------------------------------------------------------------------------------------
__global__ void fjsp_cells_pack_actions_kernel(const uint8_t* __restrict__ v_actions, uint8_t* __restrict__ actions, int64_t num_envs,
                                               int K, int act_dim) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (env, action column)
    if (i >= num_envs * act_dim) return;
    const int64_t env = i / act_dim;
    const int col = (int)(i % act_dim);
    uint8_t a = 0;
    if (col == 0) a = v_actions[env * K * 8];
    else if (col < 1 + 7 * K) a = v_actions[(env * K + (col - 1) / 7) * 8 + 1 + (col - 1) % 7];
    actions[i] = a;
}

__global__ void fjsp_cells_unpack_views_kernel(const float* __restrict__ obs, const int8_t* __restrict__ masks,
                                               const float* __restrict__ rewards, const uint8_t* __restrict__ flags,
                                               float* __restrict__ v_obs, int8_t* __restrict__ v_masks, float* __restrict__ v_rewards,
                                               uint8_t* __restrict__ v_flags, int64_t num_envs, int K, int obs_dim, int mask_dim,
                                               int act_dim) {
    // one thread per (view row, 0..81): 38 observation floats, 32 mask bytes, 8 rewards, 4 flag bytes
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_envs * K * 82) return;
    const int64_t row = i / 82, env = row / K;
    const int j = (int)(i % 82), c = (int)(row % K);
    if (j < 38) {
        v_obs[row * 38 + j] = obs[env * obs_dim + (j < 7 ? j : 31 * c + j)];
    } else if (j < 70) {
        const int m = j - 38;
        int8_t v = 0;
        if (m < 3) v = c == 0 ? masks[env * mask_dim + m] : (int8_t)(m == 0);
        else if (m < 29) v = masks[env * mask_dim + 26 * c + m];
        v_masks[row * 32 + m] = v;
    } else if (j < 78) {
        const int a = j - 70;
        v_rewards[row * 8 + a] = rewards[env * act_dim + (a == 0 ? 0 : 7 * c + a)];
    } else {
        v_flags[row * 4 + (j - 78)] = flags[env * 4 + (j - 78)];
    }
}

// ---------------------------------------------------------------------------------------------
// Weight gradients with a NARROW side: G[i][j] += sum_b X[b][i] * Y[b][j] with nx <= 256 columns of X and ny <= 40 of Y —
// dW of the actors' heads (256 x 3..8), of the first layers (3..38 x 256, computed transposed) and of the critic's value
// head (128 x 1).  As tcgen05 GEMMs these cost as much as a 256 x 256 product (the wide operand is loaded, split and
// staged all the same; the main loop is latency-bound per K chunk) for 1..15 % of its flops: 18 of the 28 problems of the
// weight-gradient launch, 2/3 of its time.  Here: plain fp32 FMAs, X streamed once from HBM (coalesced: thread = column
// of X), the narrow rows of Y broadcast from shared memory, one atomic per (i, j) per slab of rows.
// ---------------------------------------------------------------------------------------------
struct WgradJob {          // 64 bytes; device array
    const float* X;        // [B][ldx], columns 0..nx-1
    const float* Y;        // [B][ldy], columns 0..ny-1
    float* G;              // G[i * gsi + j * gsj] += ...
    int32_t B, nx, ny, ldx, ldy, gsi, gsj;
    int32_t reserved[3];
};
static_assert(sizeof(WgradJob) == 64, "WgradJob layout is part of the ABI (include/fjsp_b200.h FjspWgradJob)");
constexpr int WG_SLAB = 512, WG_SUB = 32;   // rows per CTA, rows per shared-memory tile of Y

constexpr int WG_STAGES = 3;
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_SUB * 256 * 4;   // 98,304: ring of X sub-blocks (32 rows x <= 256 floats)

template <int NY>
__global__ void __launch_bounds__(256) fjsp_a2c_wgrad_small_kernel(const WgradJob* __restrict__ jobs) {
    __shared__ float sy[WG_SUB][NY];
    __shared__ uint64_t bar[WG_STAGES];
    extern __shared__ __align__(128) unsigned char wg_smem[];
    const WgradJob J = jobs[blockIdx.y];
    const int r0 = blockIdx.x * WG_SLAB;
    if (r0 >= J.B || J.ny > NY) return;
    const int r1 = min(J.B, r0 + WG_SLAB), i = threadIdx.x;
    const bool act = i < J.nx;
    float acc[NY];
#pragma unroll
    for (int j = 0; j < NY; j++) acc[j] = 0.f;
    // Rows of X that are contiguous in memory (ldx == nx, 16-byte aligned: the trainer's activation gradients) arrive by
    // cp.async.bulk into a ring of three 32-row sub-blocks, as in fjsp_a2c_head_backward_kernel; otherwise plain loads below.
    if (J.ldx == J.nx && (J.nx & 3) == 0 && (reinterpret_cast<uintptr_t>(J.X) & 15) == 0) {
        float* sh = reinterpret_cast<float*>(wg_smem);
        const int nsub = (r1 - r0 + WG_SUB - 1) / WG_SUB;
        auto issue = [&](int it) {
            const int rb = r0 + it * WG_SUB, nr = min(WG_SUB, r1 - rb);
            const uint32_t bytes = (uint32_t)(nr * J.nx * 4);
            uint64_t* b = &bar[it % WG_STAGES];
            mbar_expect_tx(b, bytes);
            bulk_g2s(sh + (it % WG_STAGES) * WG_SUB * J.nx, J.X + (int64_t)rb * J.nx, bytes, b);
        };
        if (i == 0) {
#pragma unroll
            for (int s = 0; s < WG_STAGES; s++) mbar_init(&bar[s], 1);
        }
        __syncthreads();
        if (i == 0) {
            for (int it = 0; it < WG_STAGES && it < nsub; it++) issue(it);
        }
        for (int it = 0; it < nsub; it++) {
            const int rb = r0 + it * WG_SUB, nr = min(WG_SUB, r1 - rb);
            __syncthreads();   // everyone is done with sub-block it - 1 (its rows of Y, its stage of the ring)
            if (i == 0 && it >= 1 && it - 1 + WG_STAGES < nsub) {
                fence_async_smem();
                issue(it - 1 + WG_STAGES);
            }
            for (int e = threadIdx.x; e < WG_SUB * NY; e += 256) {
                const int r = e / NY, j = e % NY;
                sy[r][j] = (r < nr && j < J.ny) ? __ldg(J.Y + (int64_t)(rb + r) * J.ldy + j) : 0.f;
            }
            __syncthreads();
            mbar_wait(&bar[it % WG_STAGES], (uint32_t)((it / WG_STAGES) & 1));
            if (act) {
                const float* xs = sh + (it % WG_STAGES) * WG_SUB * J.nx + i;
#pragma unroll 8
                for (int r = 0; r < WG_SUB; r++) {
                    const float x = r < nr ? xs[r * J.nx] : 0.f;
#pragma unroll
                    for (int j = 0; j < NY; j++) acc[j] = fmaf(x, sy[r][j], acc[j]);
                }
            }
        }
        if (act) {
#pragma unroll
            for (int j = 0; j < NY; j++)
                if (j < J.ny) atomicAdd(J.G + (int64_t)i * J.gsi + (int64_t)j * J.gsj, acc[j]);
        }
        return;
    }
    for (int rb = r0; rb < r1; rb += WG_SUB) {
        const int nr = min(WG_SUB, r1 - rb);
        // the wide operand's 32 loads are issued first: they are in flight while the block's rows of Y go through shared memory
        // (issued after the second barrier, every 32 rows paid the two latencies one after the other)
        float x[WG_SUB];
        if (act) {
#pragma unroll
            for (int r = 0; r < WG_SUB; r++) x[r] = r < nr ? __ldg(J.X + (int64_t)(rb + r) * J.ldx + i) : 0.f;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < WG_SUB * NY; e += 256) {
            const int r = e / NY, j = e % NY;
            sy[r][j] = (r < nr && j < J.ny) ? __ldg(J.Y + (int64_t)(rb + r) * J.ldy + j) : 0.f;
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int r = 0; r < WG_SUB; r++) {
#pragma unroll
                for (int j = 0; j < NY; j++) acc[j] = fmaf(x[r], sy[r][j], acc[j]);
            }
        }
    }
    if (act) {
#pragma unroll
        for (int j = 0; j < NY; j++)
            if (j < J.ny) atomicAdd(J.G + (int64_t)i * J.gsi + (int64_t)j * J.gsj, acc[j]);
    }
}

// ---------------------------------------------------------------------------------------------
// clip_grad_norm_ per network + Adam (a2c.py:668,686-690: nn.utils.clip_grad_norm_(net.parameters(), 0.5) then
// optim.Adam.step()), over a table of parameter SEGMENTS — one per (tensor, network): the six small actors are slices of
// stacked tensors.  torch's fused multi-tensor Adam spends 0.34 ms on these 654,366 parameters in 28 tensors and the
// clipping another dozen launches; here: one launch for the nine squared norms, one for clip + Adam, one that bumps the
// step counters and clears the norms.  Same arithmetic as torch.optim.Adam (capturable, no amsgrad / weight decay):
//   m = lerp(m, g, 1 - b1); v = b2 v + (1 - b2) g^2; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// on the optimizer's OWN state tensors (exp_avg, exp_avg_sq, step), so checkpoints and a later opt.step() see them.
// ---------------------------------------------------------------------------------------------
struct OptSeg {            // 64 bytes; device array
    float* param;
    float* grad;
    float* m;
    float* v;
    float* step;           // the parameter tensor's step counter (float, device); bumped by the segment with bump != 0
    int32_t n, net;        // elements; network index (norm / clip group), < 16
    float lr;
    int32_t bump;
    int64_t reserved;
};
static_assert(sizeof(OptSeg) == 64, "OptSeg layout is part of the ABI (include/fjsp_b200.h FjspOptSeg)");

__global__ void __launch_bounds__(256) fjsp_a2c_gradnorm_kernel(const OptSeg* __restrict__ segs, float* __restrict__ norms_sq) {
    const OptSeg S = segs[blockIdx.y];
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += gridDim.x * blockDim.x) {
        const float g = S.grad[i];
        acc = fmaf(g, g, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += part[w];
        if (t != 0.f) atomicAdd(norms_sq + S.net, t);
    }
}

__global__ void __launch_bounds__(256) fjsp_a2c_clip_adam_kernel(const OptSeg* __restrict__ segs, const float* __restrict__ norms_sq,
                                                                 float max_norm, double beta1, double beta2, double eps_d) {
    const OptSeg S = segs[blockIdx.y];
    const float coef = fminf(max_norm / (sqrtf(norms_sq[S.net]) + 1e-6f), 1.0f);   // clip_grad_norm_: clamped, always applied
    // the scalars as torch's fused (capturable) Adam forms them: the betas rounded to float first — 1 - 0.999f is 1.3e-5 below
    // 0.001, and measurably so in exp_avg_sq — then float per-element arithmetic
    const float b1 = (float)beta1, b2 = (float)beta2, eps = (float)eps_d;
    const double t = (double)S.step[0] + 1.0;
    const float bc1 = (float)(1.0 - pow((double)b1, t)), bc2 = (float)(1.0 - pow((double)b2, t));
    const float step_size = S.lr / bc1, bc2_sqrt = sqrtf(bc2);
    const float w1 = 1.0f - b1, w2 = 1.0f - b2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += gridDim.x * blockDim.x) {
        const float g = S.grad[i] * coef;
        S.grad[i] = g;
        float m = S.m[i], v = S.v[i];
        m = m + w1 * (g - m);                          // torch.lerp(m, g, 1 - beta1), weight < 0.5
        v = b2 * v + w2 * g * g;
        S.m[i] = m, S.v[i] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        S.param[i] -= step_size * m / denom;
    }
}

__global__ void fjsp_a2c_opt_finish_kernel(const OptSeg* __restrict__ segs, int nseg, float* __restrict__ norms_sq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg && segs[i].bump) segs[i].step[0] += 1.0f;
    if (i < 16) norms_sq[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// First layers of the actors in a rollout step (networks.py:22-38: Linear(3..13 -> 256) + ReLU): y = relu(x W + b) with
// K <= 40 as plain fp32 FMAs, one job per network.  As a tensor-core launch this layer is ONE K chunk wrapped in the
// GEMM kernel's fixed costs (TMEM allocation, 99 KB pipeline, staged epilogue): 17 us of a 70 us rollout step for 0.04 GFLOP.
// CTA = 32 rows x 256 columns; thread = 4 consecutive columns of 8 rows; x tile and W in shared memory; a warp stores
// 512 contiguous bytes per row.
// ---------------------------------------------------------------------------------------------
struct Layer1Job {         // 64 bytes; device array
    const float* X;        // [rows][ldx], columns 0..k-1 (an observation slice)
    const float* W;        // [k][n] row-major
    const float* bias;     // [n]
    float* Y;              // [rows][ldy]
    int32_t rows, k, n, ldx, ldy, relu;
    int32_t reserved[2];
};
static_assert(sizeof(Layer1Job) == 64, "Layer1Job layout is part of the ABI (include/fjsp_b200.h FjspLayer1Job)");
constexpr int L1_ROWS = 32, L1_KMAX = 40, L1_NMAX = 256;

__global__ void __launch_bounds__(256) fjsp_a2c_layer1_kernel(const Layer1Job* __restrict__ jobs, int max_k) {
    extern __shared__ __align__(16) float l1_smem[];   // W: [max_k][256], then the x tile: [32][41]
    float* sw = l1_smem;
    float (*sx)[L1_KMAX + 1] = reinterpret_cast<float (*)[L1_KMAX + 1]>(l1_smem + max_k * L1_NMAX);
    const Layer1Job J = jobs[blockIdx.y];
    const int r0 = blockIdx.x * L1_ROWS;
    if (r0 >= J.rows || J.k > max_k) return;
    const int tid = threadIdx.x, nrows = min(L1_ROWS, J.rows - r0);
    // independent loads, several in flight: with a plain strided loop each of the k iterations paid an L2 round trip in series
    if (tid < J.n) {
#pragma unroll 8
        for (int kk = 0; kk < J.k; kk++) sw[kk * L1_NMAX + tid] = __ldg(J.W + kk * J.n + tid);
    }
#pragma unroll 5
    for (int i = tid; i < nrows * J.k; i += 256) sx[i / J.k][i % J.k] = __ldg(J.X + (int64_t)(r0 + i / J.k) * J.ldx + (i % J.k));
    __syncthreads();
    const int c = 4 * (tid & 63), rq = tid >> 6;     // columns c..c+3; rows rq, rq + 4, ...
    if (c >= J.n) return;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (J.bias) {
        b.x = __ldg(J.bias + c);
        if (c + 1 < J.n) b.y = __ldg(J.bias + c + 1);
        if (c + 2 < J.n) b.z = __ldg(J.bias + c + 2);
        if (c + 3 < J.n) b.w = __ldg(J.bias + c + 3);
    }
    const bool vec = c + 4 <= J.n && (J.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(J.Y) & 15) == 0;
    for (int r = rq; r < nrows; r += 4) {
        float4 a = b;
        for (int k = 0; k < J.k; k++) {
            const float x = sx[r][k];
            const float4 w = *reinterpret_cast<const float4*>(sw + k * L1_NMAX + c);
            a.x = fmaf(x, w.x, a.x), a.y = fmaf(x, w.y, a.y), a.z = fmaf(x, w.z, a.z), a.w = fmaf(x, w.w, a.w);
        }
        if (J.relu) a.x = fmaxf(a.x, 0.f), a.y = fmaxf(a.y, 0.f), a.z = fmaxf(a.z, 0.f), a.w = fmaxf(a.w, 0.f);
        float* y = J.Y + (int64_t)(r0 + r) * J.ldy + c;
        if (vec) {
            *reinterpret_cast<float4*>(y) = a;
        } else {
            y[0] = a.x;
            if (c + 1 < J.n) y[1] = a.y;
            if (c + 2 < J.n) y[2] = a.z;
            if (c + 3 < J.n) y[3] = a.w;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Backward through an actor's logits layer (networks.py:22-38: Linear(256 -> 3..8)) — and the critic's value head
// (Linear(128 -> 1)) — in ONE pass over the layer's input activations H (a2c.py:647-731 backward):
//     dH[m, n]  = (sum_j dl[m, j] * W[n, j]) * (H[m, n] > 0)      gradient into the ReLU layer below
//     gb[n]    += sum_m dH[m, n]                                   that layer's bias gradient
//     gW[n, j] += sum_m H[m, n] * dl[m, j]                         the head's own weight gradient
// As two kernels (a K <= 8 tensor-core launch that is all epilogue + the narrow-gradient kernel) H was read twice;
// here a row of H is read once: thread = 4 columns of 64 rows of a 512-row slab, the head's weights and its gradient's
// partial sums in registers, the rows of dl broadcast from shared memory, one reduction per CTA.
// ---------------------------------------------------------------------------------------------
struct HeadBwdJob {        // 64 bytes; device array
    const float* dl;       // [rows][ld_dl], columns 0..na-1 (gradient of the head's outputs)
    const float* W;        // [n][na] row-major
    const float* H;        // [rows][n]: post-ReLU input of the head
    float* dH;             // [rows][n]
    float* gW;             // [n][na], accumulated into
    float* gb;             // [n], accumulated into
    int32_t rows, n, na, ld_dl;
};
static_assert(sizeof(HeadBwdJob) == 64, "HeadBwdJob layout is part of the ABI (include/fjsp_b200.h FjspHeadBwdJob)");
constexpr int HB_SLAB = 512, HB_SUB = 32, HB_STAGES = 3;
constexpr int HB_SMEM_BYTES = HB_STAGES * HB_SUB * 256 * 4;   // 98,304: the ring of H sub-blocks (32 rows x <= 256 floats each)

__global__ void __launch_bounds__(256, 2) fjsp_a2c_head_backward_kernel(const HeadBwdJob* __restrict__ jobs) {
    // The rows of H arrive by bulk copies (cp.async.bulk) into a ring of three 32-row sub-blocks — two CTAs keep 6 x 32 KB in
    // flight per SM; with plain loads the 16 warps an SM holds (the head's weights and its gradient's partial sums take 64
    // registers per thread) reached half of the HBM bandwidth (ncu: no warp eligible in 65 % of the cycles).
    extern __shared__ __align__(128) unsigned char hb_smem[];
    float* sh = reinterpret_cast<float*>(hb_smem);
    __shared__ uint64_t bar[HB_STAGES];
    __shared__ __align__(16) float sdl[HB_SUB][8];
    __shared__ float sred[256][9];   // per column: the head's 8 weight-gradient sums and the column sum
    const HeadBwdJob J = jobs[blockIdx.y];
    const int r0 = blockIdx.x * HB_SLAB;
    if (r0 >= J.rows) return;
    const int r1 = min(J.rows, r0 + HB_SLAB), tid = threadIdx.x;
    const int nsub = (r1 - r0 + HB_SUB - 1) / HB_SUB;
    const int c = 4 * (tid & 63), rq = tid >> 6;
    const bool active = c < J.n;      // (n % 4 == 0: checked by the host)
    auto issue = [&](int it) {        // thread 0: sub-block `it` of the slab into stage it % HB_STAGES
        const int rb = r0 + it * HB_SUB, nr = min(HB_SUB, r1 - rb);
        const uint32_t bytes = (uint32_t)(nr * J.n * 4);
        uint64_t* b = &bar[it % HB_STAGES];
        mbar_expect_tx(b, bytes);
        bulk_g2s(sh + (it % HB_STAGES) * HB_SUB * J.n, J.H + (int64_t)rb * J.n, bytes, b);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < HB_STAGES; s++) mbar_init(&bar[s], 1);
    }
    float w[4][8], acc[4][8], cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            w[i][j] = (active && j < J.na) ? __ldg(J.W + (int64_t)(c + i) * J.na + j) : 0.f;
            acc[i][j] = 0.f;
        }
    }
    for (int i = tid; i < 256 * 9; i += 256) (&sred[0][0])[i] = 0.f;
    __syncthreads();   // barriers initialised
    if (tid == 0) {
        for (int it = 0; it < HB_STAGES && it < nsub; it++) issue(it);
    }
    for (int it = 0; it < nsub; it++) {
        const int rb = r0 + it * HB_SUB;
        __syncthreads();   // everyone is done with sub-block it - 1: its rows of dl, and its stage of the ring
        if (tid == 0 && it >= 1 && it - 1 + HB_STAGES < nsub) {
            fence_async_smem();
            issue(it - 1 + HB_STAGES);
        }
        {
            const int r = tid >> 3, j = tid & 7;   // 32 rows x 8 columns = 256 threads
            sdl[r][j] = (rb + r < r1 && j < J.na) ? __ldg(J.dl + (int64_t)(rb + r) * J.ld_dl + j) : 0.f;
        }
        __syncthreads();
        mbar_wait(&bar[it % HB_STAGES], (uint32_t)((it / HB_STAGES) & 1));
        if (active) {
            const int nr = min(HB_SUB, r1 - rb);
            const float* hs = sh + (it % HB_STAGES) * HB_SUB * J.n;
#pragma unroll 2
            for (int r = rq; r < nr; r += 4) {
                const int64_t m = rb + r;
                const float4 h4 = *reinterpret_cast<const float4*>(hs + r * J.n + c);
                const float4 d0 = *reinterpret_cast<const float4*>(&sdl[r][0]), d1 = *reinterpret_cast<const float4*>(&sdl[r][4]);
                const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                const float h[4] = {h4.x, h4.y, h4.z, h4.w};
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float v = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; j++) v = fmaf(d[j], w[i][j], v), acc[i][j] = fmaf(h[i], d[j], acc[i][j]);
                    o[i] = h[i] > 0.f ? v : 0.f;
                    cs[i] += o[i];
                }
                *reinterpret_cast<float4*>(J.dH + m * J.n + c) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < J.na) atomicAdd(&sred[c + i][j], acc[i][j]);
            atomicAdd(&sred[c + i][8], cs[i]);
        }
    }
    __syncthreads();
    for (int e = tid; e < J.n * 9; e += 256) {
        const int col = e / 9, j = e % 9;
        const float v = sred[col][j];
        if (j == 8) {
            if (J.gb && v != 0.f) atomicAdd(J.gb + col, v);
        } else if (j < J.na && v != 0.f) {
            atomicAdd(J.gW + (int64_t)col * J.na + j, v);
        }
    }
}

}  // namespace fjsp
