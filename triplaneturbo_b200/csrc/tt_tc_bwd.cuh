// Tensor-core (tcgen05) kernels of the backward path.  One group of 128 threads per CTA (thread = sample point =
// TMEM lane), persistent over 128-point tiles.
//
// Every matrix product of the backward runs on the tensor cores:
//   * decoder recompute: 3xTF32 (SDF branch: the ReLU masks must equal the forward's) / 1xTF32 (colour branch),
//   * adjoint / tangent layers: 1xTF32 (gradient precision), A operand in TMEM,
//   * weight gradients dW = Σ_points a ⊗ x: both operands written by the threads into K-major shared-memory tiles
//     (K = the tile's 128 points), accumulated in TMEM across all tiles of the CTA and flushed once with atomicAdd.
//     The MMA shape is M = 128: only the first 64 (or 3) rows of A are meaningful, the remaining rows read whatever
//     follows in shared memory and land in accumulator rows that are never read.
// Plane gradients: cooperative scatter (consecutive lanes = consecutive 16-byte chunks of one texel) with
// red.global.add.v4.f32.
#pragma once
#include "tt_tc.cuh"

namespace tt {

template <int C, int NPL>
__device__ __forceinline__ void coop_scatter(float* __restrict__ gplanes, size_t ps, const int* tap_o, const float* tap_w,
                                             const uint32_t* pbase, int plane0, const float* stage, int tg) {
    constexpr int U = C / 4, NT = 4 * NPL, SP = C + 4;
#pragma unroll 1
    for (int j = 0; j < U; ++j) {
        const int item = tg + TC_GROUP * j;
        const int pt = item / U, ch = item - pt * U;
        const float4 v = *reinterpret_cast<const float4*>(stage + pt * SP + ch * 4);
        float* base = gplanes + (size_t)pbase[pt] * 6 * ps + (size_t)plane0 * ps + ch * 4;
#pragma unroll
        for (int k = 0; k < NPL; ++k)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int o = tap_o[pt * NT + k * 4 + t];
                const float ww = tap_w[pt * NT + k * 4 + t];
                if (ww != 0.f)
                    red_add4(base + k * ps + (size_t)o * C, make_float4(v.x * ww, v.y * ww, v.z * ww, v.w * ww));
            }
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#ifndef TT_EMUL
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
#endif
    return v;
}

// ================================================================================================ SDF branch
// The forward (k_geo_tc<C,true>) saves the two ReLU masks of every non-empty sample (16 B), so nothing of the SDF
// decoder is recomputed here: per tile  a1 = m1 ⊙ W2ᵀ(m2 ⊙ w3), de = W1ᵀ a1  (adjoint, unit seed), scatter de·ω,
// gather ẽ = Σ ω·texel, h̃1 = m1 ⊙ W1 ẽ  (tangent; h̃2 = m2 ⊙ W2 h̃1 is never formed), with dW1 += a1 ẽᵀ and Q += m2 h̃1ᵀ
// accumulated in TMEM (dW2 = w3 ⊙ Q row-wise and dw3 = rowsum(W2 ⊙ Q) at the final flush).
// G independent groups of 128 threads per CTA share the weight tiles; each group owns 256 TMEM columns:
//   A [0,64)  dW2 [64,128)  D [128,192)  dW1 [192,192+CP).
template <int C>
struct BwdGeoSmem {
    static constexpr int CP = (C + 15) / 16 * 16;
    static constexpr int W1H = 0, W2TH = W1H + 64 * C, W1TH = W2TH + 4096, W3 = W1TH + CP * 64;
    static constexpr int GROUP0 = W3 + 64;
    // per group; the operand tiles come first: the M=128 MMA over-reads the 64-row A tile into the B tile behind it.
    // STAGE (point-major [128][C+4] hand-off of the cooperative gather / scatter) aliases the B tile.
    static constexpr int AT = 0, BT = AT + wg_tile_floats(64), STAGE = BT;
    static constexpr int TAP_O = BT + wg_tile_floats(64), TAP_OM = TAP_O + 128 * 12, PBASE = TAP_OM + 128 * 12;
    static constexpr int GROUP_FLOATS = PBASE + 128;
    static constexpr int G = (GROUP0 + 2 * GROUP_FLOATS + 16) * 4 <= 227 * 1024 ? 2 : 1;
    static constexpr int TOTAL = GROUP0 + G * GROUP_FLOATS + 16;
    static constexpr uint32_t COL_G2 = 64, COL_G1 = 192;
    static_assert(128 * (C + 4) <= wg_tile_floats(64), "stage must fit in the B tile");
};

template <int C>
__global__ void __launch_bounds__(BwdGeoSmem<C>::G * TC_GROUP, 1)
k_bwd_geo_tc(const float* __restrict__ planes, const float* __restrict__ wp, tt_config cfg, TcSrc src, int64_t N,
             const float* __restrict__ gs_i, const float* __restrict__ u_i, const uint64_t* __restrict__ masks,
             float* __restrict__ gplanes, float* __restrict__ gw) {
    TT_SHARED(smem);
    using L = BwdGeoSmem<C>;
    constexpr int CP = L::CP, SP = C + 4, G = L::G, NT = G * TC_GROUP;
    const int tid = threadIdx.x, group = tid / TC_GROUP, tg = tid % TC_GROUP, warp = tid >> 5;
    const WOff wo = woff(C);
    const GOff go = goff(C);
    // single-pass (rounded tf32) operands: W1 for the tangent pass, W2ᵀ, W1ᵀ for the adjoint pass
    for (int i = tid; i < 64 * C; i += NT) {
        const int n = i / C, k = i % C;
        smem[L::W1H + btile_off(n, k, C)] = tf32_rn(__ldg(wp + wo.w1s + n * C + k));
    }
    for (int i = tid; i < 64 * 64; i += NT) {
        const int n = i / 64, k = i % 64;
        smem[L::W2TH + btile_off(n, k, 64)] = tf32_rn(__ldg(wp + wo.w2s + k * 64 + n));
    }
    for (int i = tid; i < CP * 64; i += NT) {
        const int n = i / 64, k = i % 64;
        smem[L::W1TH + btile_off(n, k, 64)] = n < C ? tf32_rn(__ldg(wp + wo.w1s + k * C + n)) : 0.f;
    }
    if (tid < 64) smem[L::W3 + tid] = __ldg(wp + wo.w3s + tid);
    float* gsm = smem + L::GROUP0 + group * L::GROUP_FLOATS;
    for (int i = tg; i < 2 * wg_tile_floats(64); i += TC_GROUP) gsm[L::AT + i] = 0.f;
    uint64_t* mbars = reinterpret_cast<uint64_t*>(smem + L::GROUP0 + G * L::GROUP_FLOATS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbars + G);
    if (tid == 0) for (int g = 0; g < G; ++g) mbar_init(mbars + g);
    if (warp == 0) tmem_alloc_warp(tmem_slot, G * TC_COLS_PER_GROUP);
    async_proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    Umma u;
    u.tmem = *tmem_slot + (uint32_t)group * TC_COLS_PER_GROUP;
    u.lane_base = (uint32_t)((warp & 3) * 32) << 16;
    u.mbar = smem_u32(mbars + group); u.phase = 0; u.group = group;
    const bool leader = tg == 0;
    const BTile bW1 = btile_make(smem + L::W1H, smem + L::W1H, 64, C);
    const BTile bW2T = btile_make(smem + L::W2TH, smem + L::W2TH, 64, 64);
    const BTile bW1T = btile_make(smem + L::W1TH, smem + L::W1TH, CP, 64);
    float* At = gsm + L::AT; float* Bt = gsm + L::BT;
    const uint32_t at_addr = smem_u32(At), bt_addr = smem_u32(Bt);
    int* tap_o = reinterpret_cast<int*>(gsm + L::TAP_O);
    float* tap_om = gsm + L::TAP_OM;
    uint32_t* pbase = reinterpret_cast<uint32_t*>(gsm + L::PBASE);
    float* stage = gsm + L::STAGE;
    const float* w3 = smem + L::W3;
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const int64_t n_live = src.count ? (int64_t)*src.count : N;
    const int64_t n_tiles = (n_live + TC_GROUP - 1) / TC_GROUP;
    bool any_tile = false;

    for (int64_t tile = (int64_t)blockIdx.x * G + group; tile < n_tiles; tile += (int64_t)gridDim.x * G) {
        const int64_t slot = tile * TC_GROUP + tg;
        const bool valid = slot < n_live;
        const int64_t id = valid ? (src.index ? (int64_t)src.index[slot] : slot) : 0;
        float gs = 0.f, uu[3] = {0.f, 0.f, 0.f};
        if (valid) { gs = gs_i[id]; uu[0] = u_i[id * 3]; uu[1] = u_i[id * 3 + 1]; uu[2] = u_i[id * 3 + 2]; }
        const bool active = valid && (gs != 0.f || uu[0] != 0.f || uu[1] != 0.f || uu[2] != 0.f);
        const uint64_t m1 = active ? masks[id * 4 + 2] : 0ull, m2 = active ? masks[id * 4 + 3] : 0ull;
        int prompt = 0;
        {
            float x[3] = {0.f, 0.f, 0.f}, p[3];
            if (active) tc_point(src, id, x, prompt);
#pragma unroll
            for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
            const float sc = 0.5f * (float)cfg.R / cfg.radius;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const Taps t = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
                const float ixd = uu[plane_ax(k)] * sc, iyd = uu[plane_ay(k)] * sc;
                const float om[4] = {gs * t.w[0] + (-t.wy0 * ixd - t.wx0 * iyd), gs * t.w[1] + (t.wy0 * ixd - t.wx1 * iyd),
                                     gs * t.w[2] + (-t.wy1 * ixd + t.wx0 * iyd), gs * t.w[3] + (t.wy1 * ixd + t.wx1 * iyd)};
                int4 o4; float4 w4;
                o4.x = (active && t.o[0] >= 0) ? t.o[0] : 0; w4.x = (active && t.o[0] >= 0) ? om[0] : 0.f;
                o4.y = (active && t.o[1] >= 0) ? t.o[1] : 0; w4.y = (active && t.o[1] >= 0) ? om[1] : 0.f;
                o4.z = (active && t.o[2] >= 0) ? t.o[2] : 0; w4.z = (active && t.o[2] >= 0) ? om[2] : 0.f;
                o4.w = (active && t.o[3] >= 0) ? t.o[3] : 0; w4.w = (active && t.o[3] >= 0) ? om[3] : 0.f;
                *reinterpret_cast<int4*>(tap_o + tg * 12 + k * 4) = o4;
                *reinterpret_cast<float4*>(tap_om + tg * 12 + k * 4) = w4;
            }
        }
        pbase[tg] = (uint32_t)prompt;
        float d[64];
        // ---- unit-seed adjoint: a1 = m1 ⊙ W2ᵀ(m2 ⊙ w3), de = W1ᵀ a1 ---------------------------------------------
        {
            float a[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) a[j] = ((m2 >> j) & 1ull) ? w3[j] : 0.f;
            umma_layer<64, 64, 1>(u, leader, a, bW2T, d);
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                a[j] = ((m1 >> j) & 1ull) ? d[j] : 0.f;
                At[wg_off(j, tg)] = tf32_rn(a[j]);                        // a1 -> A tile of dW1
            }
            float de[CP];
            umma_layer<64, CP, 1>(u, leader, a, bW1T, de);
#pragma unroll
            for (int c = 0; c < C; c += 4)
                *reinterpret_cast<float4*>(stage + tg * SP + c) = make_float4(de[c], de[c + 1], de[c + 2], de[c + 3]);
        }
        group_sync(group);
        if (gplanes) coop_scatter<C, 3>(gplanes, ps, tap_o, tap_om, pbase, 0, stage, tg);     // d L / d texel = de · ω
        group_sync(group);
        coop_gather<C, 3>(planes, ps, tap_o, tap_om, pbase, 0, stage, tg);                    // ẽ = Σ ω · texel
        group_sync(group);
        {
            float e[C];
#pragma unroll
            for (int c = 0; c < C; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(stage + tg * SP + c);
                e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
            }
            group_sync(group);                                                       // stage is the B tile
#pragma unroll
            for (int c = 0; c < C; ++c) Bt[wg_off(c, tg)] = tf32_rn(e[c]);
#pragma unroll
            for (int c = C; c < CP; ++c) Bt[wg_off(c, tg)] = 0.f;
            umma_put_A1<C>(u, e);
            async_proxy_fence();
            group_sync(group);
            if (leader) {
                if (gw) umma_mma_ss(u, at_addr, bt_addr, CP, L::COL_G1, any_tile);   // dW1 += a1 ẽᵀ
                umma_mma<1>(u, bW1, C, false);                                       // W1 ẽ
                umma_commit(u);
            }
            umma_wait(u);
            umma_get_D<64>(u, d);
        }
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            Bt[wg_off(j, tg)] = ((m1 >> j) & 1ull) ? tf32_rn(d[j]) : 0.f;            // h̃1
            At[wg_off(j, tg)] = ((m2 >> j) & 1ull) ? 1.f : 0.f;                     // m2 (exact in tf32)
        }
        async_proxy_fence();
        group_sync(group);
        // Q += m2 h̃1ᵀ.  Not committed here: the next commit of this thread (next tile's first layer, or the one after
        // the loop) also covers it, and nothing touches the operand tiles before that wait.
        if (leader && gw) umma_mma_ss(u, at_addr, bt_addr, 64, L::COL_G2, any_tile);
        any_tile = true;
    }
    if (any_tile) {           // (group-uniform) drain the last tile's MMAs
        if (leader) umma_commit(u);
        umma_wait(u);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (gw && any_tile) {
        if (tg < 64) {        // accumulator rows 0..63 = TMEM lanes 0..63
            float g1[CP];
            umma_get_D<CP>(u, g1, L::COL_G1);
#pragma unroll
            for (int c = 0; c < C; ++c) if (g1[c] != 0.f) atomicAdd(gw + go.g1s + tg * C + c, g1[c]);
            // Q[j][k] = Σ_p m2[p][j] h̃1[p][k]:  dW2[j][k] = w3[j] Q[j][k],  dw3[j] = Σ_p h̃2[p][j] = Σ_k W2[j][k] Q[j][k]
            float q[64];
            umma_get_D<64>(u, q, L::COL_G2);
            const float w3j = w3[tg];
            float s3 = 0.f;
#pragma unroll
            for (int k = 0; k < 64; ++k) {
                if (q[k] != 0.f) atomicAdd(gw + go.g2s + tg * 64 + k, w3j * q[k]);
                s3 = fmaf(__ldg(wp + wo.w2s + tg * 64 + k), q[k], s3);
            }
            if (s3 != 0.f) atomicAdd(gw + go.g3s + tg, s3);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_warp(*tmem_slot, G * TC_COLS_PER_GROUP);
}

// ================================================================================================ colour branch
template <int C>
struct BwdTexSmem {
    static constexpr int CP = (C + 15) / 16 * 16;
    static constexpr int AT = 0, BT = AT + wg_tile_floats(72);                  // rows 0-63: g_h; rows 64-71: gf (3 used)
    static constexpr int W1H = BT + wg_tile_floats(64), W2H = W1H + 3 * 64 * C, W2TH = W2H + 4096, W1TH = W2TH + 4096;
    static constexpr int W3 = W1TH + 3 * CP * 64;
    static constexpr int TAP_O = W3 + 192, TAP_W = TAP_O + 128 * 12, PBASE = TAP_W + 128 * 12, STAGE = PBASE + 128;
    static constexpr int TOTAL = STAGE + 128 * (C + 4) + 16;
    // TMEM columns: A [0,64), D1 [64,128), D0 [128,192), dW2 [192,256), dW3 [256,320), dW1 3 x CP from 320,
    // de_k: k=0 at D0, k=1 at D1, k=2 at 320 + 3 CP
    static constexpr uint32_t COL_D1 = 64, COL_GW2 = 192, COL_GW3 = 256, COL_GW1 = 320, COL_DE2 = 320 + 3 * CP;
    static constexpr bool TMEM_OK = 320 + 4 * CP <= 512;     // otherwise the host falls back to the SIMT kernels
};

template <int C>
__global__ void __launch_bounds__(TC_GROUP, 1) k_bwd_tex_tc(const float* __restrict__ planes, const float* __restrict__ wp,
                                                           tt_config cfg, TcSrc src, int64_t N,
                                                           const float* __restrict__ gf_i,
                                                           const uint64_t* __restrict__ masks,
                                                           float* __restrict__ gplanes, float* __restrict__ gw) {
    TT_SHARED(smem);
    using L = BwdTexSmem<C>;
    constexpr int CP = L::CP, SP = C + 4;
    const int tid = threadIdx.x, warp = tid >> 5;
    const WOff wo = woff(C);
    const GOff go = goff(C);
    for (int k = 0; k < 3; ++k) {
        for (int i = tid; i < 64 * C; i += TC_GROUP) {          // W1f plane tiles [64][C] (forward, single pass)
            const int n = i / C, kk = i % C;
            smem[L::W1H + k * 64 * C + btile_off(n, kk, C)] = tf32_rn(__ldg(wp + wo.w1f + n * 3 * C + k * C + kk));
        }
        for (int i = tid; i < CP * 64; i += TC_GROUP) {         // W1f_kᵀ tiles [CP][64]
            const int n = i / 64, kk = i % 64;
            smem[L::W1TH + k * CP * 64 + btile_off(n, kk, 64)] = n < C ? tf32_rn(__ldg(wp + wo.w1f + kk * 3 * C + k * C + n)) : 0.f;
        }
    }
    for (int i = tid; i < 4096; i += TC_GROUP) {
        const int n = i / 64, k = i % 64;
        smem[L::W2H + btile_off(n, k, 64)] = tf32_rn(__ldg(wp + wo.w2f + n * 64 + k));
        smem[L::W2TH + btile_off(n, k, 64)] = tf32_rn(__ldg(wp + wo.w2f + k * 64 + n));
    }
    for (int i = tid; i < wg_tile_floats(72) + wg_tile_floats(64); i += TC_GROUP) smem[L::AT + i] = 0.f;
    for (int i = tid; i < 192; i += TC_GROUP) smem[L::W3 + i] = __ldg(wp + wo.w3f + i);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + L::STAGE + 128 * (C + 4));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    if (tid == 0) mbar_init(mbar);
    if (warp == 0) tmem_alloc_warp(tmem_slot, 512);
    async_proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    Umma u;
    u.tmem = *tmem_slot; u.lane_base = (uint32_t)((warp & 3) * 32) << 16;
    u.mbar = smem_u32(mbar); u.phase = 0; u.group = 0;
    const bool leader = tid == 0;
    BTile bW1[3], bW1T[3];
    for (int k = 0; k < 3; ++k) {
        bW1[k] = btile_make(smem + L::W1H + k * 64 * C, smem + L::W1H + k * 64 * C, 64, C);
        bW1T[k] = btile_make(smem + L::W1TH + k * CP * 64, smem + L::W1TH + k * CP * 64, CP, 64);
    }
    const BTile bW2 = btile_make(smem + L::W2H, smem + L::W2H, 64, 64);
    const BTile bW2T = btile_make(smem + L::W2TH, smem + L::W2TH, 64, 64);
    float* At = smem + L::AT; float* Bt = smem + L::BT;
    const uint32_t at_addr = smem_u32(At), bt_addr = smem_u32(Bt);
    const uint32_t at_gf_addr = at_addr + 8 * WG_SBO;            // row group 8: rows 64..71 hold gf
    int* tap_o = reinterpret_cast<int*>(smem + L::TAP_O);
    float* tap_w = smem + L::TAP_W;
    uint32_t* pbase = reinterpret_cast<uint32_t*>(smem + L::PBASE);
    float* stage = smem + L::STAGE;
    const float* w3 = smem + L::W3;
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const int64_t n_live = src.count ? (int64_t)*src.count : N;
    const int64_t n_tiles = (n_live + TC_GROUP - 1) / TC_GROUP;
    bool any_tile = false;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t slot = tile * TC_GROUP + tid;
        const bool valid = slot < n_live;
        const int64_t id = valid ? (src.index ? (int64_t)src.index[slot] : slot) : 0;
        float gf[3] = {0.f, 0.f, 0.f};
        if (valid) { gf[0] = gf_i[id * 3]; gf[1] = gf_i[id * 3 + 1]; gf[2] = gf_i[id * 3 + 2]; }
        const bool active = valid && (gf[0] != 0.f || gf[1] != 0.f || gf[2] != 0.f);
        if (!__syncthreads_or(active)) continue;        // nothing to do in this tile
        int prompt = 0;
        {
            float x[3] = {0.f, 0.f, 0.f}, p[3];
            if (active) tc_point(src, id, x, prompt);
#pragma unroll
            for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const Taps t = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool in = active && t.o[q] >= 0;
                    tap_o[k * 512 + tid * 4 + q] = in ? t.o[q] : 0;
                    tap_w[k * 512 + tid * 4 + q] = in ? t.w[q] : 0.f;
                }
            }
        }
        pbase[tid] = (uint32_t)prompt;
        float d[64];
        // ---- recompute (single pass): h1 = relu(Σ_k W1f_k e_k), h2 = relu(W2f h1) ----------------------------------
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            group_sync(0);
            coop_gather<C, 1>(planes, ps, tap_o + k * 512, tap_w + k * 512, pbase, 3 + k, stage, tid);
            group_sync(0);
            float e[C];
#pragma unroll
            for (int c = 0; c < C; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(stage + tid * SP + c);
                e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
            }
            if (k > 0) umma_wait(u);
            umma_put_A1<C>(u, e);
            group_sync(0);
            if (leader) { umma_mma<1>(u, bW1[k], C, k > 0); umma_commit(u); }
        }
        umma_wait(u);
        umma_get_D<64>(u, d);
        // the ReLU masks are the forward's (3xTF32) masks: a single-pass recompute may flip units near zero
        const uint64_t m1 = active ? masks[id * 4] : 0ull, m2 = active ? masks[id * 4 + 1] : 0ull;
        {
            float h[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const bool on = (m1 >> j) & 1ull;
                h[j] = on ? d[j] : 0.f;
                Bt[wg_off(j, tid)] = tf32_rn(h[j]);                                   // h1 -> B tile of dW2
            }
            umma_layer<64, 64, 1>(u, leader, h, bW2, d);                              // pre-activations stay in D0
        }
        // ---- g_h2 = m2 ⊙ W3ᵀ gf ; dW2 += g_h2 h1ᵀ ; g_h1 = m1 ⊙ W2ᵀ g_h2 (output to D1 so that D0 keeps h2) ---------
        {
            float g2[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const float v = gf[0] * w3[j] + gf[1] * w3[64 + j] + gf[2] * w3[128 + j];
                g2[j] = ((m2 >> j) & 1ull) ? v : 0.f;
                At[wg_off(j, tid)] = tf32_rn(g2[j]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) At[wg_off(64 + c, tid)] = tf32_rn(gf[c]);
            umma_put_A1<64>(u, g2);
            async_proxy_fence();
            group_sync(0);
            if (leader) {
                if (gw) umma_mma_ss(u, at_addr, bt_addr, 64, L::COL_GW2, any_tile);
                umma_mma<1>(u, bW2T, 64, false);
                umma_commit(u);
            }
            umma_wait(u);
        }
        float gh1[64];                       // (`d` still holds the layer-2 pre-activations: h2 = relu(d))
        umma_get_D<64>(u, gh1);
#pragma unroll
        for (int j = 0; j < 64; ++j) gh1[j] = ((m1 >> j) & 1ull) ? gh1[j] : 0.f;
        // ---- dW3 += gf h2ᵀ (A rows 64..66), then the three data gradients de_k = W1f_kᵀ g_h1 ---------------------------
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            Bt[wg_off(j, tid)] = ((m2 >> j) & 1ull) ? tf32_rn(d[j]) : 0.f;             // h2
            At[wg_off(j, tid)] = tf32_rn(gh1[j]);                                      // g_h1 -> A tile of dW1
        }
        umma_put_A1<64>(u, gh1);
        async_proxy_fence();
        group_sync(0);
        if (leader) {
            if (gw) umma_mma_ss(u, at_gf_addr, bt_addr, 64, L::COL_GW3, any_tile);
            umma_commit(u);
        }
        umma_wait(u);
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            // de_k = W1f_kᵀ g_h1 (A = g_h1 still in TMEM), and dW1_k += g_h1 e_kᵀ with e_k re-gathered
            coop_gather<C, 1>(planes, ps, tap_o + k * 512, tap_w + k * 512, pbase, 3 + k, stage, tid);
            group_sync(0);
            {
                float e[C];
#pragma unroll
                for (int c = 0; c < C; c += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(stage + tid * SP + c);
                    e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
                }
#pragma unroll
                for (int c = 0; c < C; ++c) Bt[wg_off(c, tid)] = tf32_rn(e[c]);
            }
            async_proxy_fence();
            tc_fence_before();
            group_sync(0);
            if (leader) {
                if (gw) umma_mma_ss(u, at_addr, bt_addr, CP, L::COL_GW1 + k * CP, any_tile);
                umma_mma<1>(u, bW1T[k], 64, false);
                umma_commit(u);
            }
            umma_wait(u);
            float de[CP];
            umma_get_D<CP>(u, de);
#pragma unroll
            for (int c = 0; c < C; c += 4)
                *reinterpret_cast<float4*>(stage + tid * SP + c) = make_float4(de[c], de[c + 1], de[c + 2], de[c + 3]);
            group_sync(0);
            if (gplanes) coop_scatter<C, 1>(gplanes, ps, tap_o + k * 512, tap_w + k * 512, pbase, 3 + k, stage, tid);
            group_sync(0);
        }
        any_tile = true;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (gw && any_tile) {
        if (tid < 64) {
            float g[64];
            umma_get_D<64>(u, g, L::COL_GW2);
#pragma unroll
            for (int j = 0; j < 64; ++j) if (g[j] != 0.f) atomicAdd(gw + go.g2f + tid * 64 + j, g[j]);
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                float g1[CP];
                umma_get_D<CP>(u, g1, L::COL_GW1 + k * CP);
#pragma unroll
                for (int c = 0; c < C; ++c) if (g1[c] != 0.f) atomicAdd(gw + go.g1f + tid * 3 * C + k * C + c, g1[c]);
            }
        }
        if (tid < 32) {        // warp 0 reads lanes 0..31; rows 0..2 hold dW3
            float g[64];
            umma_get_D<64>(u, g, L::COL_GW3);
            if (tid < 3) {
#pragma unroll
                for (int j = 0; j < 64; ++j) if (g[j] != 0.f) atomicAdd(gw + go.g3f + tid * 64 + j, g[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_warp(*tmem_slot, 512);
}

}  // namespace tt
