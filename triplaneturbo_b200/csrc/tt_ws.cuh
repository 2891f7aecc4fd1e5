// Warp-specialised tensor-core kernels of the forward path (impl = 2, the default).
//
// The round-1 kernels (tt_tc.cuh) ran every phase of a tile on the same 128 threads: gather, layer round trips and
// the normal pass were serialised per group, and their times ADDED UP (DESIGN.md "Phase anatomy").  Here a CTA has
// three roles that overlap through mbarrier hand-offs:
//
//   gather warps (threads 0..127, "M group")   compute the sample positions and the bilinear tap tables of a 128-point
//                                              tile, gather the texels cooperatively (consecutive lanes = consecutive
//                                              16-byte chunks of a channel-last texel, 24 loads in flight per lane) and
//                                              write BLENDED rows into a shared-memory stage; they run ahead of the
//                                              consumers by up to two stage buffers per consumer group
//   2 consumer groups (2 x 128 threads)        thread = stage row = TMEM lane: row -> tf32 hi/lo split -> TMEM A operand,
//                                              one elected thread issues the tcgen05.mma of a layer (3xTF32), the
//                                              epilogue reads the accumulator back in 32-column halves (ReLU / masks /
//                                              split), second layer, head.  The two groups alternate chunks, so one
//                                              group's epilogue overlaps the other's MMAs.
//
// Analytic normal without a second gather: instead of the adjoint pass (de = W1ᵀ(m1 ⊙ W2ᵀ(m2 ⊙ w3)) followed by a second
// pass over the 12 taps), the gather warps also blend the three TANGENT vectors  V_a = ∂e/∂x_a  (same 12 texels, tap
// coefficients ∂w_t/∂x_a), and a chunk is 32 points x 4 rows (e, V_x, V_y, V_z).  The tangent rows run through the
// same two layers with the ReLU masks of their point's primal row (one warp shuffle: the 4 rows of a point are 4
// adjacent lanes), and  ∂sdf/∂x_a = Σ_j m2_j w3_j (W2 (m1 ⊙ W1 V_a))_j  comes out of the same epilogue as the SDF.
// That removes the second gather (41 of the 137 ms of the round-1 fine pass at config 2), both transposed weight tiles
// (-48 KB of shared memory, which the L1 cache gets back) and two of the four dependent layer round trips per point.
//
// TMA note: a cp.async.bulk.tensor producer (one 2x2xC box per point and plane, hardware zero fill = zeros padding) was
// measured and is NOT used: box issue costs ~67 cycles per box and warp (tools/tma_gather_probe.cu,
// profiles/r02_tma_gather_probe.txt: 2.1-4.1 TB/s against 9.8-19 TB/s for the cooperative LDG gather with 4-12 warps).
#pragma once
#include "tt_tc.cuh"

namespace tt {

constexpr int WS_M = 128;                               // gather threads
constexpr int WS_CG = 2;                                // consumer groups
constexpr int WS_THREADS = WS_M + WS_CG * TC_GROUP;     // 384
constexpr int WS_NBUF = 2;                              // stage buffers per consumer group

template <int C, bool NORMAL>
struct GeoWs {
    static constexpr int SP = C + 4;
    static constexpr int NK = NORMAL ? 4 : 1;           // stage rows per point: e (+ V_x, V_y, V_z)
    static constexpr int PTS = TC_GROUP / NK;           // points per chunk (= 128 stage rows)
    static constexpr int NCOEF = NORMAL ? 3 : 1;        // coefficient tables per tap: w (+ dw/dix, dw/diy)
    // float offsets
    static constexpr int W1H = 0, W1L = W1H + 64 * C, W2H = W1L + 64 * C, W2L = W2H + 4096, W3 = W2L + 4096;
    static constexpr int MTAB = W3 + 64;                // gather-group tables of one 128-point tile
    static constexpr int TAP_O = 0, TAP_W = TAP_O + 128 * 12, TAP_CX = TAP_W + 128 * 12;
    static constexpr int TAP_CY = TAP_CX + (NORMAL ? 128 * 12 : 0), PBASE = TAP_CY + (NORMAL ? 128 * 12 : 0);
    static constexpr int MMETA = PBASE + 128, MTAB_FLOATS = MMETA + 128 * 4;
    static constexpr int STAGE0 = MTAB + MTAB_FLOATS;
    static constexpr int ROWS = 0, META = ROWS + 128 * SP, STAGE_FLOATS = META + PTS * 4;
    static constexpr int NSTAGE = WS_CG * WS_NBUF;
    static constexpr int BARS = STAGE0 + NSTAGE * STAGE_FLOATS;        // uint64: full[NSTAGE], empty[NSTAGE], mma[WS_CG]
    static constexpr int TOTAL = BARS + 2 * (2 * NSTAGE + WS_CG) + 4;
    static_assert(BARS % 2 == 0, "mbarriers must be 8-byte aligned");
};

// ---- gather of one chunk -------------------------------------------------------------------------------------------------
// item = (point, 16-byte channel chunk); thread mt takes items mt, mt + 128, ...  JB items per batch so that 12 * JB
// independent 16-byte loads are in flight per lane.  p0 = first point of the chunk inside the tile's tables.
template <int C, bool NORMAL>
__device__ __forceinline__ void ws_gather_chunk(const float* __restrict__ planes, size_t ps, const int* tap_o,
                                                const float* tap_w, const float* tap_cx, const float* tap_cy,
                                                const uint32_t* pbase, int p0, float* rows, int mt) {
    using L = GeoWs<C, NORMAL>;
    constexpr int U = C / 4, SP = C + 4, JB = 2, ITEMS = L::PTS * U;
#pragma unroll 1
    for (int i0 = mt; i0 < ITEMS; i0 += JB * WS_M) {
        float4 v[JB][12];
        int pt[JB], ch[JB];
#pragma unroll
        for (int b = 0; b < JB; ++b) {
            const int item = i0 + b * WS_M < ITEMS ? i0 + b * WS_M : ITEMS - 1;      // tail: repeat the last item (not stored)
            pt[b] = item / U; ch[b] = item - pt[b] * U;
            const int P = p0 + pt[b];
            const float* base = planes + (size_t)pbase[P] * 6 * ps + ch[b] * 4;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int4 o4 = *reinterpret_cast<const int4*>(tap_o + P * 12 + k * 4);
                const float* pb = base + (size_t)k * ps;
                v[b][k * 4 + 0] = ldg4(pb + (size_t)o4.x * C); v[b][k * 4 + 1] = ldg4(pb + (size_t)o4.y * C);
                v[b][k * 4 + 2] = ldg4(pb + (size_t)o4.z * C); v[b][k * 4 + 3] = ldg4(pb + (size_t)o4.w * C);
            }
        }
#pragma unroll
        for (int b = 0; b < JB; ++b) {
            if (i0 + b * WS_M >= ITEMS) continue;
            const int P = p0 + pt[b];
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 V[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) V[a] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 w4 = *reinterpret_cast<const float4*>(tap_w + P * 12 + k * 4);
                const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float4 q = v[b][k * 4 + t];
                    s.x = fmaf(ww[t], q.x, s.x); s.y = fmaf(ww[t], q.y, s.y); s.z = fmaf(ww[t], q.z, s.z); s.w = fmaf(ww[t], q.w, s.w);
                }
                e.x += s.x; e.y += s.y; e.z += s.z; e.w += s.w;
                if (NORMAL) {
                    const float4 x4 = *reinterpret_cast<const float4*>(tap_cx + P * 12 + k * 4);
                    const float4 y4 = *reinterpret_cast<const float4*>(tap_cy + P * 12 + k * 4);
                    const float cx[4] = {x4.x, x4.y, x4.z, x4.w}, cy[4] = {y4.x, y4.y, y4.z, y4.w};
                    const int ax = k == 2 ? 2 : 0, ay = k == 1 ? 2 : 1;          // plane_ax / plane_ay, compile time
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float4 q = v[b][k * 4 + t];
                        V[ax].x = fmaf(cx[t], q.x, V[ax].x); V[ax].y = fmaf(cx[t], q.y, V[ax].y);
                        V[ax].z = fmaf(cx[t], q.z, V[ax].z); V[ax].w = fmaf(cx[t], q.w, V[ax].w);
                        V[ay].x = fmaf(cy[t], q.x, V[ay].x); V[ay].y = fmaf(cy[t], q.y, V[ay].y);
                        V[ay].z = fmaf(cy[t], q.z, V[ay].z); V[ay].w = fmaf(cy[t], q.w, V[ay].w);
                    }
                }
            }
            float* r = rows + (size_t)(pt[b] * L::NK) * SP + ch[b] * 4;
            *reinterpret_cast<float4*>(r) = e;
            if (NORMAL) {
                *reinterpret_cast<float4*>(r + SP) = V[0];
                *reinterpret_cast<float4*>(r + 2 * SP) = V[1];
                *reinterpret_cast<float4*>(r + 3 * SP) = V[2];
            }
        }
    }
}

// 32 accumulator columns -> registers
__device__ __forceinline__ void ws_ld32(const Umma& u, uint32_t col, float (&d)[32]) {
    uint32_t v[32];
    tmem_ld32(u.tmem + u.lane_base + col, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) d[j] = __uint_as_float(v[j]);
}
// 32 activation columns -> TMEM A operand (tf32 hi + exact remainder)
__device__ __forceinline__ void ws_st32_split(const Umma& u, uint32_t col, const float (&x)[32]) {
    uint32_t hi[32], lo[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float h = tf32_hi(x[j]);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(x[j] - h);
    }
    tmem_st32(u.tmem + u.lane_base + TC_COL_AHI + col, hi);
    tmem_st32(u.tmem + u.lane_base + TC_COL_ALO + col, lo);
}
__device__ __forceinline__ uint32_t ws_pos_bits(const float (&d)[32]) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) m |= (d[j] > 0.f ? 1u : 0u) << j;
    return m;
}
__device__ __forceinline__ uint32_t ws_shfl_u32(uint32_t v, int src_lane) {
    return (uint32_t)__shfl_sync(0xffffffffu, (int)v, src_lane);
}

// SDF decoder (+ analytic normal) at a list of points.  sources and outputs as k_geo_tc.
template <int C, bool NORMAL>
__global__ void __launch_bounds__(WS_THREADS, 1) k_geo_ws(const float* __restrict__ planes, const float* __restrict__ wp,
                                                         tt_config cfg, TcSrc src, int64_t N, float* sdf_o,
                                                         float* sdf_orig_o, float* grad_o, float* normal_o,
                                                         uint64_t* masks_o) {
    TT_SHARED(smem);
    using L = GeoWs<C, NORMAL>;
    constexpr int SP = L::SP, NK = L::NK, PTS = L::PTS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const WOff wo = woff(C);
    btile_fill(smem + L::W1H, smem + L::W1L, 64, C, [&](int n, int k) { return __ldg(wp + wo.w1s + n * C + k); }, tid, WS_THREADS);
    btile_fill(smem + L::W2H, smem + L::W2L, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2s + n * 64 + k); }, tid, WS_THREADS);
    if (tid < 64) smem[L::W3 + tid] = __ldg(wp + wo.w3s + tid);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BARS);
    uint64_t* full = bars; uint64_t* empty = bars + L::NSTAGE; uint64_t* mmab = bars + 2 * L::NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * L::NSTAGE + WS_CG);
    if (tid == 0) {
        for (int i = 0; i < L::NSTAGE; ++i) { mbar_init_n(full + i, WS_M); mbar_init_n(empty + i, TC_GROUP); }
        for (int g = 0; g < WS_CG; ++g) mbar_init_n(mmab + g, 1);
        mbar_init_fence();
    }
    if (warp == 0) tmem_alloc_warp(tmem_slot, 512);
    async_proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const int64_t n_live = src.count ? (int64_t)*src.count : N;
    const int64_t n_tiles = (n_live + TC_GROUP - 1) / TC_GROUP;

    if (tid < WS_M) {
        // ============================================================================ gather warps
        const int mt = tid;
        float* tab = smem + L::MTAB;
        int* tap_o = reinterpret_cast<int*>(tab + L::TAP_O);
        float* tap_w = tab + L::TAP_W; float* tap_cx = tab + L::TAP_CX; float* tap_cy = tab + L::TAP_CY;
        uint32_t* pbase = reinterpret_cast<uint32_t*>(tab + L::PBASE);
        float* mmeta = tab + L::MMETA;
        int64_t chunk = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t slot = tile * TC_GROUP + mt;
            const bool valid = slot < n_live;
            const int64_t id = valid ? (src.index ? (int64_t)src.index[slot] : slot) : 0;
            float x[3] = {0.f, 0.f, 0.f}; int prompt = 0;
            Taps tp[3];
            if (valid) {
                tc_point(src, id, x, prompt);
                float p[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
#pragma unroll
                for (int k = 0; k < 3; ++k) tp[k] = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int4 o4; float4 w4, x4, y4;
                const bool i0 = valid && tp[k].o[0] >= 0, i1 = valid && tp[k].o[1] >= 0;
                const bool i2 = valid && tp[k].o[2] >= 0, i3 = valid && tp[k].o[3] >= 0;
                o4.x = i0 ? tp[k].o[0] : 0; o4.y = i1 ? tp[k].o[1] : 0; o4.z = i2 ? tp[k].o[2] : 0; o4.w = i3 ? tp[k].o[3] : 0;
                w4.x = i0 ? tp[k].w[0] : 0.f; w4.y = i1 ? tp[k].w[1] : 0.f; w4.z = i2 ? tp[k].w[2] : 0.f; w4.w = i3 ? tp[k].w[3] : 0.f;
                *reinterpret_cast<int4*>(tap_o + mt * 12 + k * 4) = o4;
                *reinterpret_cast<float4*>(tap_w + mt * 12 + k * 4) = w4;
                if (NORMAL) {      // d e / d ix = wy0 (t1 - t0) + wy1 (t3 - t2),  d e / d iy = wx0 (t2 - t0) + wx1 (t3 - t1)
                    x4.x = i0 ? -tp[k].wy0 : 0.f; x4.y = i1 ? tp[k].wy0 : 0.f; x4.z = i2 ? -tp[k].wy1 : 0.f; x4.w = i3 ? tp[k].wy1 : 0.f;
                    y4.x = i0 ? -tp[k].wx0 : 0.f; y4.y = i1 ? -tp[k].wx1 : 0.f; y4.z = i2 ? tp[k].wx0 : 0.f; y4.w = i3 ? tp[k].wx1 : 0.f;
                    *reinterpret_cast<float4*>(tap_cx + mt * 12 + k * 4) = x4;
                    *reinterpret_cast<float4*>(tap_cy + mt * 12 + k * 4) = y4;
                }
            }
            pbase[mt] = (uint32_t)prompt;
            *reinterpret_cast<float4*>(mmeta + mt * 4) = make_float4(__uint_as_float((uint32_t)(valid ? (int)id : -1)), x[0], x[1], x[2]);
            group_sync(0);
#pragma unroll 1
            for (int q = 0; q < NK; ++q, ++chunk) {
                const int sb = (int)(chunk & 1) * WS_NBUF + (int)((chunk >> 1) & 1);
                mbar_wait_parity(smem_u32(empty + sb), (uint32_t)(((chunk >> 2) & 1) ^ 1));
                float* st = smem + L::STAGE0 + sb * L::STAGE_FLOATS;
                ws_gather_chunk<C, NORMAL>(planes, ps, tap_o, tap_w, tap_cx, tap_cy, pbase, q * PTS, st + L::ROWS, mt);
                if (mt < PTS) *reinterpret_cast<float4*>(st + L::META + mt * 4) = *reinterpret_cast<const float4*>(mmeta + (q * PTS + mt) * 4);
                mbar_arrive(smem_u32(full + sb));
            }
            group_sync(0);          // the tables are rewritten by the next tile
        }
    } else {
        // ============================================================================ consumer groups
        const int g = (tid - WS_M) / TC_GROUP, tg = (tid - WS_M) % TC_GROUP;
        Umma u;
        u.tmem = *tmem_slot + (uint32_t)g * TC_COLS_PER_GROUP;
        u.lane_base = (uint32_t)((warp & 3) * 32) << 16;
        u.mbar = smem_u32(mmab + g); u.phase = 0; u.group = 1 + g;
        const bool leader = tg == 0;
        const BTile bW1 = btile_make(smem + L::W1H, smem + L::W1L, 64, C);
        const BTile bW2 = btile_make(smem + L::W2H, smem + L::W2L, 64, 64);
        const float* w3 = smem + L::W3;
        const int prim = lane & ~3;                     // lane of this point's primal row (NORMAL)
        int64_t chunk = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
            for (int q = 0; q < NK; ++q, ++chunk) {
                if ((int)(chunk & 1) != g) continue;
                const int sb = g * WS_NBUF + (int)((chunk >> 1) & 1);
                mbar_wait_parity(smem_u32(full + sb), (uint32_t)((chunk >> 2) & 1));
                const float* st = smem + L::STAGE0 + sb * L::STAGE_FLOATS;
                const float4 mv = *reinterpret_cast<const float4*>(st + L::META + (NORMAL ? (tg >> 2) : tg) * 4);
                {
                    float e[C];
#pragma unroll
                    for (int c = 0; c < C; c += 4) {
                        const float4 v = *reinterpret_cast<const float4*>(st + L::ROWS + tg * SP + c);
                        e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
                    }
                    umma_put_A<C>(u, e);
                }
                mbar_arrive(smem_u32(empty + sb));      // row and meta are in registers / TMEM: the buffer is free
                group_sync(u.group);
                if (leader) { umma_mma<3>(u, bW1, C, false); umma_commit(u); }
                umma_wait(u);
                // ---- layer-1 epilogue: ReLU (primal row) / mask (tangent rows), split, second-layer operand
                uint32_t m1[2], m2[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float d[32];
                    ws_ld32(u, TC_COL_D + 32 * h, d);
                    uint32_t bits = ws_pos_bits(d);
                    if (NORMAL) bits = ws_shfl_u32(bits, prim);
                    m1[h] = bits;
#pragma unroll
                    for (int j = 0; j < 32; ++j) d[j] = ((bits >> j) & 1u) ? d[j] : 0.f;
                    ws_st32_split(u, 32 * h, d);
                }
                tmem_wait_st();
                tc_fence_before();
                group_sync(u.group);
                if (leader) { umma_mma<3>(u, bW2, 64, false); umma_commit(u); }
                umma_wait(u);
                // ---- layer-2 epilogue + head: primal row -> sdf, tangent row a -> d sdf / d x_a (before the index scale)
                float val = 0.f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float d[32];
                    ws_ld32(u, TC_COL_D + 32 * h, d);
                    uint32_t bits = ws_pos_bits(d);
                    if (NORMAL) bits = ws_shfl_u32(bits, prim);
                    m2[h] = bits;
#pragma unroll
                    for (int j = 0; j < 32; ++j) val = fmaf(((bits >> j) & 1u) ? d[j] : 0.f, w3[32 * h + j], val);
                }
                tc_fence_before();
                float tx = 0.f, ty = 0.f, tz = 0.f;
                if (NORMAL) {
                    tx = __shfl_sync(0xffffffffu, val, prim + 1);
                    ty = __shfl_sync(0xffffffffu, val, prim + 2);
                    tz = __shfl_sync(0xffffffffu, val, prim + 3);
                }
                const int id32 = (int)__float_as_uint(mv.x);
                if (id32 >= 0 && (!NORMAL || (lane & 3) == 0)) {
                    const int64_t id = id32;
                    const float x[3] = {mv.y, mv.z, mv.w};
                    const float nrm = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
                    if (sdf_orig_o) sdf_orig_o[id] = val;
                    if (sdf_o) sdf_o[id] = val + (nrm - cfg.sdf_bias_radius);
                    if (masks_o) {     // ReLU masks for the backward
                        masks_o[id * 4 + 2] = (uint64_t)m1[0] | ((uint64_t)m1[1] << 32);
                        masks_o[id * 4 + 3] = (uint64_t)m2[0] | ((uint64_t)m2[1] << 32);
                    }
                    if (NORMAL && (grad_o || normal_o)) {
                        const float scale = 0.5f * (float)cfg.R / cfg.radius;
                        const float inv = nrm > 0.f ? 1.f / nrm : 0.f;
                        const float gr[3] = {tx * scale + x[0] * inv, ty * scale + x[1] * inv, tz * scale + x[2] * inv};
                        if (grad_o) { grad_o[id * 3] = gr[0]; grad_o[id * 3 + 1] = gr[1]; grad_o[id * 3 + 2] = gr[2]; }
                        if (normal_o) {
                            float n[3], len; normalize3(gr, n, len);
                            normal_o[id * 3] = n[0]; normal_o[id * 3 + 1] = n[1]; normal_o[id * 3 + 2] = n[2];
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_warp(*tmem_slot, 512);
}

// ================================================================================================ colour decoder
// Feature MLP at a list of (live) samples: three texture-plane encodings concatenated (tex_interpolate v2), 3C -> 64 -> 64 -> 3.
// The gather warps hand one PLANE of a 128-point tile at a time to the consumers (stage rows of C floats); the first layer
// is the sum of three K = C MMAs, one per plane, into the same accumulator.  For C <= 32 the three A operands live in
// disjoint TMEM columns, so the consumer never waits between planes: put A_k, issue MMA_k, take the next plane.
template <int C>
struct TexWs {
    static constexpr int SP = C + 4;
    static constexpr int NREG = (6 * C <= 192) ? 3 : 1;            // disjoint A regions for the three planes
    static constexpr uint32_t COL_D = 192, COL_LO1 = 96;             // single-region layout: hi [0,C) lo [96,96+C); layer 2: hi [0,64) lo [96,160)
    static constexpr int W1H = 0, W1L = W1H + 3 * 64 * C, W2H = W1L + 3 * 64 * C, W2L = W2H + 4096, W3 = W2L + 4096;
    static constexpr int MTAB = W3 + 192;
    static constexpr int TAP_O = 0, TAP_W = TAP_O + 128 * 12, PBASE = TAP_W + 128 * 12, MIDS = PBASE + 128, MTAB_FLOATS = MIDS + 128;
    static constexpr int STAGE0 = MTAB + MTAB_FLOATS;
    static constexpr int ROWS = 0, META = ROWS + 128 * SP, STAGE_FLOATS = META + 128;
    static constexpr int NSTAGE = WS_CG * WS_NBUF;
    static constexpr int BARS = STAGE0 + NSTAGE * STAGE_FLOATS;
    static constexpr int TOTAL = BARS + 2 * (2 * NSTAGE + WS_CG) + 4;
    static_assert(BARS % 2 == 0, "mbarriers must be 8-byte aligned");
    __host__ __device__ static constexpr uint32_t col_hi(int k) { return NREG == 3 ? (uint32_t)(k * 2 * C) : 0u; }
    __host__ __device__ static constexpr uint32_t col_lo(int k) { return NREG == 3 ? (uint32_t)(k * 2 * C + C) : COL_LO1; }
};

// one texture plane of a 128-point tile: 4 taps per item, JB items per batch (24 loads in flight per lane)
template <int C>
__device__ __forceinline__ void ws_gather_plane(const float* __restrict__ planes, size_t ps, const int* tap_o, const float* tap_w,
                                                const uint32_t* pbase, int k, float* rows, int mt) {
    constexpr int U = C / 4, SP = C + 4, JB = 6, ITEMS = 128 * U;
#pragma unroll 1
    for (int i0 = mt; i0 < ITEMS; i0 += JB * WS_M) {
        float4 v[JB][4];
        int pt[JB], ch[JB];
#pragma unroll
        for (int b = 0; b < JB; ++b) {
            const int item = i0 + b * WS_M < ITEMS ? i0 + b * WS_M : ITEMS - 1;
            pt[b] = item / U; ch[b] = item - pt[b] * U;
            const float* pb = planes + ((size_t)pbase[pt[b]] * 6 + 3 + k) * ps + ch[b] * 4;
            const int4 o4 = *reinterpret_cast<const int4*>(tap_o + pt[b] * 12 + k * 4);
            v[b][0] = ldg4(pb + (size_t)o4.x * C); v[b][1] = ldg4(pb + (size_t)o4.y * C);
            v[b][2] = ldg4(pb + (size_t)o4.z * C); v[b][3] = ldg4(pb + (size_t)o4.w * C);
        }
#pragma unroll
        for (int b = 0; b < JB; ++b) {
            if (i0 + b * WS_M >= ITEMS) continue;
            const float4 w4 = *reinterpret_cast<const float4*>(tap_w + pt[b] * 12 + k * 4);
            const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float4 q = v[b][t];
                s.x = fmaf(ww[t], q.x, s.x); s.y = fmaf(ww[t], q.y, s.y); s.z = fmaf(ww[t], q.z, s.z); s.w = fmaf(ww[t], q.w, s.w);
            }
            *reinterpret_cast<float4*>(rows + pt[b] * SP + ch[b] * 4) = s;
        }
    }
}

// this thread's row x[0..K) -> TMEM columns col_hi.. (tf32 hi) and col_lo.. (exact remainder)
template <int K>
__device__ __forceinline__ void ws_put_row(const Umma& u, const float (&x)[K], uint32_t col_hi, uint32_t col_lo) {
    constexpr int K32 = K / 32 * 32;
#pragma unroll
    for (int k0 = 0; k0 < K32; k0 += 32) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float h = tf32_hi(x[k0 + j]);
            hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(x[k0 + j] - h);
        }
        tmem_st32(u.tmem + u.lane_base + col_hi + k0, hi);
        tmem_st32(u.tmem + u.lane_base + col_lo + k0, lo);
    }
#pragma unroll
    for (int k0 = K32; k0 < K; k0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float h = tf32_hi(x[k0 + j]);
            hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(x[k0 + j] - h);
        }
        tmem_st8(u.tmem + u.lane_base + col_hi + k0, hi);
        tmem_st8(u.tmem + u.lane_base + col_lo + k0, lo);
    }
    tmem_wait_st();
    tc_fence_before();
}

template <int C>
__global__ void __launch_bounds__(WS_THREADS, 1) k_tex_ws(const float* __restrict__ planes, const float* __restrict__ wp,
                                                         tt_config cfg, TcSrc src, int64_t N, float* feat_o,
                                                         uint64_t* masks_o) {
    TT_SHARED(smem);
    using L = TexWs<C>;
    constexpr int SP = L::SP;
    const int tid = threadIdx.x, warp = tid >> 5;
    const WOff wo = woff(C);
    for (int k = 0; k < 3; ++k)      // W1f as three [64][C] K-major tiles (one per texture plane)
        btile_fill(smem + L::W1H + k * 64 * C, smem + L::W1L + k * 64 * C, 64, C,
                   [&](int n, int kk) { return __ldg(wp + wo.w1f + n * 3 * C + k * C + kk); }, tid, WS_THREADS);
    btile_fill(smem + L::W2H, smem + L::W2L, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2f + n * 64 + k); }, tid, WS_THREADS);
    if (tid < 192) smem[L::W3 + tid] = __ldg(wp + wo.w3f + tid);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BARS);
    uint64_t* full = bars; uint64_t* empty = bars + L::NSTAGE; uint64_t* mmab = bars + 2 * L::NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * L::NSTAGE + WS_CG);
    if (tid == 0) {
        for (int i = 0; i < L::NSTAGE; ++i) { mbar_init_n(full + i, WS_M); mbar_init_n(empty + i, TC_GROUP); }
        for (int g = 0; g < WS_CG; ++g) mbar_init_n(mmab + g, 1);
        mbar_init_fence();
    }
    if (warp == 0) tmem_alloc_warp(tmem_slot, 512);
    async_proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const int64_t n_live = src.count ? (int64_t)*src.count : N;
    const int64_t n_tiles = (n_live + TC_GROUP - 1) / TC_GROUP;

    if (tid < WS_M) {
        // ============================================================================ gather warps
        const int mt = tid;
        float* tab = smem + L::MTAB;
        int* tap_o = reinterpret_cast<int*>(tab + L::TAP_O);
        float* tap_w = tab + L::TAP_W;
        uint32_t* pbase = reinterpret_cast<uint32_t*>(tab + L::PBASE);
        int* mids = reinterpret_cast<int*>(tab + L::MIDS);
        int64_t t = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            const int64_t slot = tile * TC_GROUP + mt;
            const bool valid = slot < n_live;
            const int64_t id = valid ? (src.index ? (int64_t)src.index[slot] : slot) : 0;
            float p[3] = {0.f, 0.f, 0.f}; int prompt = 0;
            if (valid) {
                float x[3];
                tc_point(src, id, x, prompt);
#pragma unroll
                for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const Taps tp = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
                int4 o4; float4 w4;
                const bool i0 = valid && tp.o[0] >= 0, i1 = valid && tp.o[1] >= 0, i2 = valid && tp.o[2] >= 0, i3 = valid && tp.o[3] >= 0;
                o4.x = i0 ? tp.o[0] : 0; o4.y = i1 ? tp.o[1] : 0; o4.z = i2 ? tp.o[2] : 0; o4.w = i3 ? tp.o[3] : 0;
                w4.x = i0 ? tp.w[0] : 0.f; w4.y = i1 ? tp.w[1] : 0.f; w4.z = i2 ? tp.w[2] : 0.f; w4.w = i3 ? tp.w[3] : 0.f;
                *reinterpret_cast<int4*>(tap_o + mt * 12 + k * 4) = o4;
                *reinterpret_cast<float4*>(tap_w + mt * 12 + k * 4) = w4;
            }
            pbase[mt] = (uint32_t)prompt;
            mids[mt] = valid ? (int)id : -1;
            group_sync(0);
            const int g = (int)(t & 1);
            const int64_t cg0 = (t >> 1) * 3;                    // chunk index inside the consumer group
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                const int64_t cg = cg0 + k;
                const int sb = g * WS_NBUF + (int)(cg & 1);
                mbar_wait_parity(smem_u32(empty + sb), (uint32_t)(((cg >> 1) & 1) ^ 1));
                float* st = smem + L::STAGE0 + sb * L::STAGE_FLOATS;
                ws_gather_plane<C>(planes, ps, tap_o, tap_w, pbase, k, st + L::ROWS, mt);
                reinterpret_cast<int*>(st + L::META)[mt] = mids[mt];
                mbar_arrive(smem_u32(full + sb));
            }
            group_sync(0);
        }
    } else {
        // ============================================================================ consumer groups
        const int g = (tid - WS_M) / TC_GROUP, tg = (tid - WS_M) % TC_GROUP;
        Umma u;
        u.tmem = *tmem_slot + (uint32_t)g * TC_COLS_PER_GROUP;
        u.lane_base = (uint32_t)((warp & 3) * 32) << 16;
        u.mbar = smem_u32(mmab + g); u.phase = 0; u.group = 1 + g;
        const bool leader = tg == 0;
        BTile bW1[3];
        for (int k = 0; k < 3; ++k) bW1[k] = btile_make(smem + L::W1H + k * 64 * C, smem + L::W1L + k * 64 * C, 64, C);
        const BTile bW2 = btile_make(smem + L::W2H, smem + L::W2L, 64, 64);
        const float* w3 = smem + L::W3;
        int64_t t = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            if ((int)(t & 1) != g) continue;
            const int64_t cg0 = (t >> 1) * 3;
            int id32 = -1;
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                const int64_t cg = cg0 + k;
                const int sb = g * WS_NBUF + (int)(cg & 1);
                mbar_wait_parity(smem_u32(full + sb), (uint32_t)((cg >> 1) & 1));
                const float* st = smem + L::STAGE0 + sb * L::STAGE_FLOATS;
                id32 = reinterpret_cast<const int*>(st + L::META)[tg];
                float e[C];
#pragma unroll
                for (int c = 0; c < C; c += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(st + L::ROWS + tg * SP + c);
                    e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
                }
                if (L::NREG == 1 && k > 0) umma_wait(u);          // the previous plane's MMAs are done reading A
                ws_put_row<C>(u, e, L::col_hi(k), L::col_lo(k));
                mbar_arrive(smem_u32(empty + sb));
                group_sync(u.group);
                if (leader) {
                    umma_mma_ex<3>(u, bW1[k], C, k > 0, L::col_hi(k), L::col_lo(k), L::COL_D);
                    if (L::NREG == 1 || k == 2) umma_commit(u);
                }
            }
            umma_wait(u);
            uint32_t m1[2], m2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float d[32];
                ws_ld32(u, L::COL_D + 32 * h, d);
                m1[h] = ws_pos_bits(d);
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float x = fmaxf(d[j], 0.f), hh = tf32_hi(x);
                    hi[j] = __float_as_uint(hh); lo[j] = __float_as_uint(x - hh);
                }
                tmem_st32(u.tmem + u.lane_base + 32 * h, hi);
                tmem_st32(u.tmem + u.lane_base + L::COL_LO1 + 32 * h, lo);
            }
            tmem_wait_st();
            tc_fence_before();
            group_sync(u.group);
            if (leader) { umma_mma_ex<3>(u, bW2, 64, false, 0, L::COL_LO1, L::COL_D); umma_commit(u); }
            umma_wait(u);
            float f[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float d[32];
                ws_ld32(u, L::COL_D + 32 * h, d);
                m2[h] = ws_pos_bits(d);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float x = fmaxf(d[j], 0.f);
                    f[0] = fmaf(x, w3[32 * h + j], f[0]); f[1] = fmaf(x, w3[64 + 32 * h + j], f[1]); f[2] = fmaf(x, w3[128 + 32 * h + j], f[2]);
                }
            }
            tc_fence_before();
            if (id32 >= 0) {
                const int64_t id = id32;
                if (feat_o) { feat_o[id * 3] = f[0]; feat_o[id * 3 + 1] = f[1]; feat_o[id * 3 + 2] = f[2]; }
                if (masks_o) {
                    masks_o[id * 4] = (uint64_t)m1[0] | ((uint64_t)m1[1] << 32);
                    masks_o[id * 4 + 1] = (uint64_t)m2[0] | ((uint64_t)m2[1] << 32);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_warp(*tmem_slot, 512);
}

}  // namespace tt
