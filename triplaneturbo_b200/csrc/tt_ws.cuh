// Warp-specialised tensor-core kernels of the forward path (impl = 2, the default).
//
// The round-1 kernels (tt_tc.cuh) ran every phase of a tile on the same 128 threads: gather, layer round trips and
// the normal pass were serialised per group and their times ADDED UP (DESIGN.md "Phase anatomy").  Here a CTA has a
// dedicated memory warpgroup and two consumer groups that overlap through mbarrier hand-offs:
//
//   gather warps (threads 0..127, "M group")   for each consumer group in turn: cooperative gather of the blended encoding of
//                                              the group's next 128-point tile (consecutive lanes = consecutive 16-byte
//                                              chunks of a channel-last texel, 36 loads in flight per lane) into the
//                                              group's stage; with the normal also the tangent rows V_a = d enc / d x_a
//                                              (same texels, other weights) into an L2-resident scratch.  The tap tables
//                                              (sample -> position -> 12 taps and weights) are made by the gather warps
//                                              in the kernel with the normal, by the CONSUMER groups in the others (CTAB)
//   2 consumer groups (2 x 128 threads)        thread = sample point = TMEM lane: the decoder layers on tcgen05 (3xTF32, A
//                                              operand in TMEM, accumulator read back in 32-column halves), heads,
//                                              outputs; normal = unit-seed adjoint (W2^T, W1^T layers) . tangent rows;
//                                              field query (DEFORM): both decoders' first layers as one N = 128 MMA chain
// While one group runs its layers the gather warps serve the other group, so the load path and the tensor pipe are busy
// at the same time (round 1: `mem_lock` spin lock between two groups that each did their own gathers).
//
// Measured alternatives that are NOT used:
//   * TMA producer (cp.async.bulk.tensor 2x2xC boxes, hardware zero fill = zeros padding): box issue costs ~67 cycles per
//     box and warp; 2.1-4.1 TB/s against 9.8-19 TB/s for the cooperative LDG gather (tools/tma_gather_probe.cu).
//   * tangent-mode normal (rows e, de/dx, de/dy, de/dz through two layers, no second gather, no transposed weights):
//     parity-green but 2x the epilogue rows and MMAs per point; the kernel is issue-bound and ran 20 % SLOWER than round 1
//     (profiles/r02_ws_tangent_*.txt).
//   * adjoint normal with a SECOND gather of the 12 taps once d sdf / d enc is known, two gather warpgroups, consumer-made
//     tables in the kernel with the normal, regular-grid z-lines: DESIGN.md 3.5 items 3, 6, 11, 12.
#pragma once
#include "tt_tc.cuh"

namespace tt {

constexpr int WS_M = 128;                               // gather threads
constexpr int WS_CG = 2;                                // consumer groups
constexpr int WS_THREADS = WS_M + WS_CG * TC_GROUP;     // 384
constexpr int WS_NBUF = 2;                              // stage buffers per consumer group (colour kernel)
// Register re-balancing between the roles (setmaxnreg works per warpgroup = 4 consecutive warps): the kernel starts with
// 65536 / 384 -> 168 registers per thread; the gather warpgroup grows to WS_REG_M (more loads in flight: the gather is
// bound by loads in flight, tools/tma_gather_probe.cu), the two consumer warpgroups shrink to WS_REG_C.
// 128 * 232 + 256 * 136 = 64512 = 384 * 168.
#ifndef WS_CTAB_NORMAL
#define WS_CTAB_NORMAL 0        // consumer-made tap tables also in the kernel WITH the normal (experiment: DESIGN 3.5 item 12)
#endif
#ifndef WS_GEO_NMG
#define WS_GEO_NMG 1            // gather warpgroups of k_geo_ws: 1 (serves both consumer groups in turn) or 2 (one per consumer group)
#endif
#ifndef WS_REG_M
#if WS_GEO_NMG == 2             // 512 threads start with 128 registers: 2 * 128 * (152 + 104) = 65536
#define WS_REG_M 152
#define WS_REG_C 104
#else
#define WS_REG_M 232
#define WS_REG_C 136
#endif
#endif
constexpr int WS_GEO_MT = WS_GEO_NMG * 128;                       // gather threads of k_geo_ws
constexpr int WS_GEO_THREADS = WS_GEO_MT + WS_CG * TC_GROUP;
#ifndef WS_GATHER_JB
#define WS_GATHER_JB 3          // items (12 loads each) in flight per gather lane
#endif
template <int N> __device__ __forceinline__ void ws_reg_inc() {
#ifndef TT_EMUL
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
#endif
}
template <int N> __device__ __forceinline__ void ws_reg_dec() {
#ifndef TT_EMUL
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
#endif
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
// bit j = (d[j] > 0).  relu first (one FMNMX, +0 for anything <= 0), then the sign of the negated bit pattern
__device__ __forceinline__ uint32_t ws_pos_bits(const float (&d)[32]) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 31; j >= 0; --j) {
        const uint32_t neg = 0u - __float_as_uint(fmaxf(d[j], 0.f));      // top bit set iff d[j] > 0
        m = (m << 1) | (neg >> 31);
    }
    return m;
}
// all-ones / all-zeros word from bit j of m
__device__ __forceinline__ uint32_t ws_bit_mask(uint32_t m, int j) { return 0u - ((m >> j) & 1u); }

// DEFORM (field query of the mesh paths, forward_field): the deformation decoder runs on the same encoding.  Its first
// layer is STACKED under the SDF decoder's ([128][C] weight tile, one N = 128 MMA chain, accumulator columns [128,256)), so a
// tile costs three layer round trips instead of four.
template <int C, bool NORMAL, bool DEFORM = false>
struct GeoWs {
    static_assert(!(NORMAL && DEFORM), "the field query has no normal");
    static constexpr int SP = C + 4, CP = (C + 15) / 16 * 16, U = C / 4;
    static constexpr int N1 = DEFORM ? 128 : 64;            // rows of the first-layer weight tile
    // float offsets: weight tiles (tf32 hi / lo, canonical K-major)
    static constexpr int W1H = 0, W1L = W1H + N1 * C, W2H = W1L + N1 * C, W2L = W2H + 4096;
    static constexpr int W2TH = W2L + 4096, W2TL = W2TH + (NORMAL ? 4096 : 0);
    static constexpr int W1TH = W2TL + (NORMAL ? 4096 : 0), W1TL = W1TH + (NORMAL ? CP * 64 : 0);
    static constexpr int W2DH = W1TL + (NORMAL ? CP * 64 : 0), W2DL = W2DH + (DEFORM ? 4096 : 0);
    static constexpr int W3D = W2DL + (DEFORM ? 4096 : 0);
    static constexpr int W3 = W3D + (DEFORM ? 192 : 0);
    // gather-group tables of the tile being gathered
    static constexpr int MTAB = W3 + 64;
    static constexpr int TAP_O = 0, TAP_W = TAP_O + 128 * 12, TAP_CX = TAP_W + 128 * 12;
    static constexpr int TAP_CY = TAP_CX + (NORMAL ? 128 * 12 : 0), PBASE = TAP_CY + (NORMAL ? 128 * 12 : 0);
    static constexpr int MTAB_FLOATS = PBASE + 128;
    // per consumer group: meta (id, x) and the blended encodings of its next tile
    static constexpr int GROUP0 = MTAB + WS_GEO_NMG * MTAB_FLOATS;
    static constexpr int META = 0, STAGE = META + 128 * 4, GROUP_FLOATS = STAGE + 128 * SP;
    static constexpr int BARS = GROUP0 + WS_CG * GROUP_FLOATS;     // uint64: full, stage_free, mma, tables_ready per group
    // CTAB (kernels without the normal): the CONSUMER groups compute the tap tables of their next tile (thread = point, in
    // the shadow of their MMA round trips) and hand them to the gather warps through `tables_ready`; the gather warps only
    // gather.  Group 0's tables live in MTAB, group 1's in CTAB1.
    static constexpr int CTAB1 = BARS + 2 * 4 * WS_CG + 4;
    static constexpr bool CTAB = WS_GEO_NMG == 1 && (!NORMAL || (WS_CTAB_NORMAL && (size_t)(CTAB1 + MTAB_FLOATS) * 4 <= 227 * 1024));
    static constexpr int TOTAL0 = CTAB1 + (CTAB ? MTAB_FLOATS : 0);
    static_assert(CTAB1 % 4 == 0, "table buffers are read as 16-byte vectors");
    static_assert(BARS % 2 == 0, "mbarriers must be 8-byte aligned");
    // regular-grid source (isosurface grid): z-lines of a tile, see ws_grid_segment.  Tables alias MTAB (unused there).
    static constexpr int GLMAX = 160;                              // lines per tile (all segments)
    static constexpr bool GRID = !NORMAL && WS_GEO_NMG == 1 && (size_t)(TOTAL0 + GLMAX * C) * 4 <= 227 * 1024;
    static constexpr int LINES = TOTAL0, TOTAL = TOTAL0 + (GRID ? GLMAX * C : 0);
    static constexpr int G_SEG = 0, G_ZT = G_SEG + 8 * 24, G_P0 = G_ZT + 128 * 4;      // inside MTAB
    static_assert(!GRID || G_P0 + 8 * C <= MTAB_FLOATS, "grid tables must fit in the tap tables");
    // tangent rows V_a = d enc / d x_a of a tile in the L2-resident scratch: [pt / 8][a][chunk][pt % 8][4 floats]
    // (the gather warps write 64-byte runs, the consumers read 128-byte runs)
    static constexpr int VTILE = 128 * 3 * C;                      // floats per (group, buffer)
    __host__ __device__ static constexpr int vidx(int pt, int a, int ch) { return ((((pt >> 3) * 3 + a) * U + ch) * 8 + (pt & 7)) * 4; }
};
__host__ __device__ constexpr size_t ws_vscratch_floats(int n_cta, int C) { return (size_t)n_cta * WS_CG * 2 * 128 * 3 * C; }

// Regular-grid source (src.mode == 3, the 512^3 isosurface grid of the mesh export): a tile is `128 / SL` SEGMENTS of SL
// consecutive grid points along z with (x, y) fixed, so
//   * plane 0 (x, y) contributes ONE blended vector per segment,
//   * plane 1 (x, z) and plane 2 (z, y) are bilinear in z over LINES that depend on the texel row / column z_i only:
//       line(z_i) = wx0 T1[z_i][x0] + wx1 T1[z_i][x0+1] + wy0 T2[y0][z_i] + wy1 T2[y0+1][z_i],
//     and a point is  e = P0 + wz0 line(z0) + wz1 line(z0 + 1).
// Per tile the gather warps load 4 + 4 x (lines) texels instead of 12 x 128 (config 5: 66 lines, 268 texels: 5.7x fewer
// loads and blend instructions).  Returns SL (0: not applicable -> generic gather).  Same formula on host and device.
__host__ __device__ inline int ws_grid_segment(int rr, int R, int lmax) {
    if (rr < 16) return 0;
    const int SL = rr < 128 ? rr : 128;
    if (128 % SL != 0 || rr % SL != 0) return 0;
    const int span = (int)(((long long)(SL - 1) * R) / (rr - 1)) + 3;      // texel rows a segment can touch (+ slack)
    return (128 / SL) * span <= lmax ? SL : 0;
}

// lower texel index and the two weights of a coordinate along one plane axis: the arithmetic of make_taps (tt_device.cuh)
__device__ __forceinline__ void ws_axis_tap(float g, int R, int& i0, float (&w)[2]) {
    const float fR = (float)R;
    const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), fR), 1.f), 0.5f);
    const float x0f = floorf(ix);
    w[0] = __fsub_rn(__fadd_rn(x0f, 1.f), ix); w[1] = __fsub_rn(ix, x0f);
    i0 = (int)fminf(fmaxf(x0f, -2.f), fR + 2.f);
}

// SDF decoder (+ analytic normal) at a list of points.  Sources and outputs as k_geo_tc (tt_tc.cuh); vscratch:
// ws_vscratch_floats(gridDim.x, C) floats (NORMAL only).
template <int C, bool NORMAL, bool DEFORM = false>
__global__ void __launch_bounds__(WS_GEO_THREADS, 1) k_geo_ws(const float* __restrict__ planes, const float* __restrict__ wp,
                                                         tt_config cfg, TcSrc src, int64_t N, float* sdf_o,
                                                         float* sdf_orig_o, float* grad_o, float* normal_o,
                                                         uint64_t* masks_o, float* __restrict__ vscratch,
                                                         float* deform_o) {
    TT_SHARED(smem);
    using L = GeoWs<C, NORMAL, DEFORM>;
    constexpr int SP = L::SP, CP = L::CP, U = L::U;
    const int tid = threadIdx.x, warp = tid >> 5;
    const WOff wo = woff(C);
    btile_fill(smem + L::W1H, smem + L::W1L, L::N1, C,
               [&](int n, int k) { return n < 64 ? __ldg(wp + wo.w1s + n * C + k) : __ldg(wp + wo.w1d + (n - 64) * C + k); }, tid, WS_GEO_THREADS);
    btile_fill(smem + L::W2H, smem + L::W2L, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2s + n * 64 + k); }, tid, WS_GEO_THREADS);
    if (DEFORM) {
        btile_fill(smem + L::W2DH, smem + L::W2DL, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2d + n * 64 + k); }, tid, WS_GEO_THREADS);
        if (tid < 192) smem[L::W3D + tid] = __ldg(wp + wo.w3d + tid);
    }
    if (NORMAL) {
        btile_fill(smem + L::W2TH, smem + L::W2TL, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2s + k * 64 + n); }, tid, WS_GEO_THREADS);
        btile_fill(smem + L::W1TH, smem + L::W1TL, CP, 64, [&](int n, int k) { return n < C ? __ldg(wp + wo.w1s + k * C + n) : 0.f; }, tid, WS_GEO_THREADS);
    }
    if (tid < 64) smem[L::W3 + tid] = __ldg(wp + wo.w3s + tid);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BARS);
    uint64_t* full = bars; uint64_t* sfree = bars + WS_CG; uint64_t* mmab = bars + 2 * WS_CG; uint64_t* tready = bars + 3 * WS_CG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * WS_CG);
    if (tid == 0) {
        for (int g = 0; g < WS_CG; ++g) { mbar_init_n(full + g, WS_M); mbar_init_n(sfree + g, TC_GROUP); mbar_init_n(mmab + g, 1); mbar_init_n(tready + g, TC_GROUP); }
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
    const int64_t tile_stride = (int64_t)gridDim.x * WS_CG;
    const int grid_sl = (L::GRID && src.mode == 3 && src.grid_lines && !src.index) ? ws_grid_segment(src.grid_res, cfg.R, L::GLMAX) : 0;   // regular-grid gather
    const bool ctab = L::CTAB && grid_sl == 0;          // tap tables by the consumer groups
    auto ctab_of = [&](int g) { return smem + (g == 0 ? L::MTAB : L::CTAB1); };
    float* vcta = NORMAL ? vscratch + (size_t)blockIdx.x * WS_CG * 2 * L::VTILE : nullptr;
    // gather jobs.  One gather warpgroup: job j serves consumer group j & 1 with its tile number j >> 1.  Two: warpgroup m
    // serves consumer group m, job j = its tile number j.
    const int mgrp = tid / 128;                                                      // gather warpgroup (tid < WS_GEO_MT)
    auto job_group = [&](int64_t j) { return WS_GEO_NMG == 2 ? mgrp : (int)(j & 1); };
    auto job_num = [&](int64_t j) { return WS_GEO_NMG == 2 ? j : (j >> 1); };
    auto job_tile = [&](int64_t j) { return (int64_t)blockIdx.x * WS_CG + job_group(j) + job_num(j) * tile_stride; };

    if (tid < WS_GEO_MT) {
        // ============================================================================ gather warps
        ws_reg_inc<WS_REG_M>();
        const int mt = tid % 128;
        float* tab = smem + L::MTAB + mgrp * L::MTAB_FLOATS;
        int* tap_o = reinterpret_cast<int*>(tab + L::TAP_O);
        float* tap_w = tab + L::TAP_W; float* tap_cx = tab + L::TAP_CX; float* tap_cy = tab + L::TAP_CY;
        uint32_t* pbase = reinterpret_cast<uint32_t*>(tab + L::PBASE);
        const bool prof_m = blockIdx.x == 0 && tid == 0; (void)prof_m;
        WS_T0(tm);
        // software pipeline over jobs: sample ids two jobs ahead, sample positions one job ahead (both are dependent
        // global loads whose latency would otherwise be exposed once per tile)
        if (ctab) {
            // the consumer group publishes the tables of its tile (tready), the gather warps blend it into the group's stage
            for (int64_t j = 0; job_tile(j) < n_tiles; ++j) {
                const int g = job_group(j);
                const uint32_t par = (uint32_t)(job_num(j) & 1);
                float* gs = smem + L::GROUP0 + g * L::GROUP_FLOATS;
                const float* ct = ctab_of(g);
                const int* t_o = reinterpret_cast<const int*>(ct + L::TAP_O);
                const float* t_w = ct + L::TAP_W;
                const uint32_t* t_p = reinterpret_cast<const uint32_t*>(ct + L::PBASE);
                mbar_wait_parity(smem_u32(tready + g), par);
                WS_ACC(1, tm, prof_m);
                mbar_wait_parity(smem_u32(sfree + g), par ^ 1u);        // the group has taken its previous tile out of the stage
                WS_ACC(0, tm, prof_m);
                float* stage = gs + L::STAGE;
                const float* t_cx = ct + L::TAP_CX; const float* t_cy = ct + L::TAP_CY; (void)t_cx; (void)t_cy;
                float* vt = NORMAL ? vcta + (size_t)(g * 2 + (int)par) * L::VTILE : nullptr;
                constexpr int JB = WS_GATHER_JB, ITEMS = 128 * U;
#pragma unroll 1
                for (int i0 = mt; i0 < ITEMS; i0 += JB * WS_M) {
                    float4 v[JB][12];
                    int pt[JB], ch[JB];
#pragma unroll
                    for (int b = 0; b < JB; ++b) {
                        const int item = i0 + b * WS_M < ITEMS ? i0 + b * WS_M : ITEMS - 1;      // tail: repeat the last item (not stored)
                        pt[b] = item / U; ch[b] = item - pt[b] * U;
                        const float* base = planes + (size_t)t_p[pt[b]] * 6 * ps + ch[b] * 4;
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk) {
                            const int4 o4 = *reinterpret_cast<const int4*>(t_o + pt[b] * 12 + kk * 4);
                            const float* pb = base + (size_t)kk * ps;
                            v[b][kk * 4 + 0] = ldg4(pb + (size_t)o4.x * C); v[b][kk * 4 + 1] = ldg4(pb + (size_t)o4.y * C);
                            v[b][kk * 4 + 2] = ldg4(pb + (size_t)o4.z * C); v[b][kk * 4 + 3] = ldg4(pb + (size_t)o4.w * C);
                        }
                    }
#pragma unroll
                    for (int b = 0; b < JB; ++b) {
                        if (i0 + b * WS_M >= ITEMS) continue;
                        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 V[3];
#pragma unroll
                        for (int a = 0; a < 3; ++a) V[a] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk) {
                            const float4 w4 = *reinterpret_cast<const float4*>(t_w + pt[b] * 12 + kk * 4);
                            const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
                            float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float4 q = v[b][kk * 4 + t];
                                sacc.x = fmaf(ww[t], q.x, sacc.x); sacc.y = fmaf(ww[t], q.y, sacc.y);
                                sacc.z = fmaf(ww[t], q.z, sacc.z); sacc.w = fmaf(ww[t], q.w, sacc.w);
                            }
                            e.x += sacc.x; e.y += sacc.y; e.z += sacc.z; e.w += sacc.w;
                            if (NORMAL) {
                                const float4 x4 = *reinterpret_cast<const float4*>(t_cx + pt[b] * 12 + kk * 4);
                                const float4 y4 = *reinterpret_cast<const float4*>(t_cy + pt[b] * 12 + kk * 4);
                                const float cx[4] = {x4.x, x4.y, x4.z, x4.w}, cy[4] = {y4.x, y4.y, y4.z, y4.w};
                                const int ax = kk == 2 ? 2 : 0, ay = kk == 1 ? 2 : 1;          // plane_ax / plane_ay, compile time
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const float4 q = v[b][kk * 4 + t];
                                    V[ax].x = fmaf(cx[t], q.x, V[ax].x); V[ax].y = fmaf(cx[t], q.y, V[ax].y);
                                    V[ax].z = fmaf(cx[t], q.z, V[ax].z); V[ax].w = fmaf(cx[t], q.w, V[ax].w);
                                    V[ay].x = fmaf(cy[t], q.x, V[ay].x); V[ay].y = fmaf(cy[t], q.y, V[ay].y);
                                    V[ay].z = fmaf(cy[t], q.z, V[ay].z); V[ay].w = fmaf(cy[t], q.w, V[ay].w);
                                }
                            }
                        }
                        *reinterpret_cast<float4*>(stage + pt[b] * SP + ch[b] * 4) = e;
                        if (NORMAL) {
#pragma unroll
                            for (int a = 0; a < 3; ++a) *reinterpret_cast<float4*>(vt + L::vidx(pt[b], a, ch[b])) = V[a];
                        }
                    }
                }
#ifndef TT_EMUL
                if (NORMAL) __threadfence_block();
#endif
                mbar_arrive(smem_u32(full + g));
                WS_ACC(2, tm, prof_m);
            }
        } else {
        WsRaw cur = ws_load_raw(src, ws_load_id(src, job_tile(0), mt, n_live, n_tiles));
        int id_next = ws_load_id(src, job_tile(1), mt, n_live, n_tiles);
        for (int64_t j = 0; job_tile(j) < n_tiles; ++j) {
            const int g = job_group(j);
            const uint32_t par = (uint32_t)(job_num(j) & 1);
            float* gs = smem + L::GROUP0 + g * L::GROUP_FLOATS;
            const WsRaw nxt = ws_load_raw(src, id_next);                        // in flight during this job's gather
            id_next = ws_load_id(src, job_tile(j + 2), mt, n_live, n_tiles);
            const bool valid = cur.id >= 0;
            float cx_[3]; int cprompt;
            ws_point_from_raw(src, cur, cx_, cprompt);
            if (L::GRID && grid_sl > 0) {
                // ================================================================ regular grid: z-lines (ws_grid_segment)
                const int SL = grid_sl, NSEG = 128 / SL, LSEG = L::GLMAX / NSEG;
                int* seg = reinterpret_cast<int*>(tab + L::G_SEG);
                float* zt = tab + L::G_ZT; float* p0 = tab + L::G_P0; float* lines = smem + L::LINES;
                int x0, y0, z0; float wx[2], wy[2], wz[2];
                ws_axis_tap(rescale1(cx_[0], cfg.radius), cfg.R, x0, wx);
                ws_axis_tap(rescale1(cx_[1], cfg.radius), cfg.R, y0, wy);
                ws_axis_tap(rescale1(cx_[2], cfg.radius), cfg.R, z0, wz);
                group_sync(mgrp);                   // every gather thread is done with the previous job's tables
                const int sg = mt / SL, sl = mt - sg * SL;
                if (sl == 0) {
                    int* q = seg + sg * 24;
                    const bool xv[2] = {x0 >= 0 && x0 < cfg.R, x0 + 1 >= 0 && x0 + 1 < cfg.R};
                    const bool yv[2] = {y0 >= 0 && y0 < cfg.R, y0 + 1 >= 0 && y0 + 1 < cfg.R};
                    q[0] = valid ? 1 : 0; q[1] = cprompt; q[2] = z0;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {        // plane 0 (x, y): nw, ne, sw, se as make_taps
                        const bool in = xv[t & 1] && yv[t >> 1];
                        q[4 + t] = in ? (y0 + (t >> 1)) * cfg.R + x0 + (t & 1) : -1;
                        q[8 + t] = (int)__float_as_uint(in ? __fmul_rn(wx[t & 1], wy[t >> 1]) : 0.f);
                    }
                    q[12] = x0; q[13] = y0;              // plane 1 (x, z): columns x0, x0 + 1 ; plane 2 (z, y): rows y0, y0 + 1
                    q[14] = (int)__float_as_uint(xv[0] ? wx[0] : 0.f); q[15] = (int)__float_as_uint(xv[1] ? wx[1] : 0.f);
                    q[16] = (int)__float_as_uint(yv[0] ? wy[0] : 0.f); q[17] = (int)__float_as_uint(yv[1] ? wy[1] : 0.f);
                }
                if (sl == SL - 1) seg[sg * 24 + 3] = z0 + 1;          // last line of the segment
                group_sync(mgrp);
                {   // this point's z taps relative to the segment's first line
                    const int zlo = seg[sg * 24 + 2];
                    *reinterpret_cast<float4*>(zt + mt * 4) = make_float4(__int_as_float(z0 - zlo), wz[0], wz[1], 0.f);
                }
                // ---- build: P0 per segment and the z-lines (both planes summed), item = (segment, line, 16-byte chunk).
                // All loads of the tile are issued before the first one is used (a thread holds up to 4 + 4 GB texel
                // chunks in registers): the phase costs ONE memory latency instead of one per batch.
                int nl = 0;                         // lines of the longest segment
                for (int s_ = 0; s_ < NSEG; ++s_) { const int n_ = seg[s_ * 24 + 3] - seg[s_ * 24 + 2] + 1; nl = n_ > nl ? n_ : nl; }
                nl = nl < LSEG ? nl : LSEG;
                const int n_line_items = NSEG * nl * U;
                constexpr int GB = 5;
                float4 pv[4]; float pw[4];
                const bool p0_item = mt < NSEG * U;
                {
                    const int s_ = p0_item ? mt / U : 0, ch = p0_item ? mt - s_ * U : 0;
                    const int* q = seg + s_ * 24;
                    const float* pb = planes + (size_t)q[1] * 6 * ps + ch * 4;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        pw[t] = (p0_item && q[0] && q[4 + t] >= 0) ? __uint_as_float((uint32_t)q[8 + t]) : 0.f;
                        pv[t] = pw[t] != 0.f ? ldg4(pb + (size_t)q[4 + t] * C) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll 1
                for (int it0 = mt; it0 < n_line_items; it0 += GB * WS_M) {
                    float4 v[GB][4]; float w[GB][4]; int li[GB];
#pragma unroll
                    for (int b = 0; b < GB; ++b) {
                        const int it = it0 + b * WS_M;
                        const bool ok = it < n_line_items;
                        const int itc = ok ? it : it0;
                        const int ch = itc % U, ln = itc / U, s_ = NSEG == 1 ? 0 : (int)((uint32_t)ln / (uint32_t)nl), l = ln - s_ * nl;
                        li[b] = ok ? (s_ * LSEG + l) * C + ch * 4 : -1;
                        const int* q = seg + s_ * 24;
                        const int zi = q[2] + l;
                        const bool in = ok && q[0] && zi >= 0 && zi < cfg.R && zi <= q[3];
                        const int x0 = q[12], y0 = q[13];
#pragma unroll
                        for (int t = 0; t < 4; ++t) { w[b][t] = in ? __uint_as_float((uint32_t)q[14 + t]) : 0.f; v[b][t] = make_float4(0.f, 0.f, 0.f, 0.f); }
                        const float* pb = planes + (size_t)q[1] * 6 * ps + ch * 4;
                        if (w[b][0] != 0.f) v[b][0] = ldg4(pb + ps + ((size_t)zi * cfg.R + x0) * C);
                        if (w[b][1] != 0.f) v[b][1] = ldg4(pb + ps + ((size_t)zi * cfg.R + x0 + 1) * C);
                        if (w[b][2] != 0.f) v[b][2] = ldg4(pb + 2 * ps + ((size_t)y0 * cfg.R + zi) * C);
                        if (w[b][3] != 0.f) v[b][3] = ldg4(pb + 2 * ps + ((size_t)(y0 + 1) * cfg.R + zi) * C);
                    }
#pragma unroll
                    for (int b = 0; b < GB; ++b) {
                        if (li[b] < 0) continue;
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int t = 0; t < 4; ++t) { a.x = fmaf(w[b][t], v[b][t].x, a.x); a.y = fmaf(w[b][t], v[b][t].y, a.y); a.z = fmaf(w[b][t], v[b][t].z, a.z); a.w = fmaf(w[b][t], v[b][t].w, a.w); }
                        *reinterpret_cast<float4*>(lines + li[b]) = a;
                    }
                }
                if (p0_item) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int t = 0; t < 4; ++t) { a.x = fmaf(pw[t], pv[t].x, a.x); a.y = fmaf(pw[t], pv[t].y, a.y); a.z = fmaf(pw[t], pv[t].z, a.z); a.w = fmaf(pw[t], pv[t].w, a.w); }
                    *reinterpret_cast<float4*>(p0 + mt * 4) = a;          // = p0[s_ * C + ch * 4]
                }
                WS_ACC(1, tm, prof_m);
                mbar_wait_parity(smem_u32(sfree + g), par ^ 1u);        // the group has taken its previous tile out of the stage
                WS_ACC(0, tm, prof_m);
                *reinterpret_cast<float4*>(gs + L::META + mt * 4) =
                    make_float4(__uint_as_float((uint32_t)cur.id), cx_[0], cx_[1], cx_[2]);
                group_sync(mgrp);
                // ---- blend: item = (point, 16-byte chunk)
                float* stage = gs + L::STAGE;
#pragma unroll 2
                for (int it = mt; it < 128 * U; it += WS_M) {
                    const int pt = it / U, ch = it - pt * U, s_ = pt / SL;
                    const float4 z4 = *reinterpret_cast<const float4*>(zt + pt * 4);
                    const int l0 = __float_as_int(z4.x);
                    const float* ln = lines + (s_ * LSEG + l0) * C + ch * 4;
                    float4 e = *reinterpret_cast<const float4*>(p0 + s_ * C + ch * 4);
                    if (l0 >= 0) { const float4 a = *reinterpret_cast<const float4*>(ln); e.x = fmaf(z4.y, a.x, e.x); e.y = fmaf(z4.y, a.y, e.y); e.z = fmaf(z4.y, a.z, e.z); e.w = fmaf(z4.y, a.w, e.w); }
                    if (l0 + 1 >= 0 && l0 + 1 < LSEG) { const float4 a = *reinterpret_cast<const float4*>(ln + C); e.x = fmaf(z4.z, a.x, e.x); e.y = fmaf(z4.z, a.y, e.y); e.z = fmaf(z4.z, a.z, e.z); e.w = fmaf(z4.z, a.w, e.w); }
                    *reinterpret_cast<float4*>(stage + pt * SP + ch * 4) = e;
                }
                mbar_arrive(smem_u32(full + g));
                WS_ACC(2, tm, prof_m);
                cur = nxt;
                continue;
            }
            Taps tp[3];
            {
                float p[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) p[a] = rescale1(cx_[a], cfg.radius);
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) tp[kk] = make_taps(p[plane_ax(kk)], p[plane_ay(kk)], cfg.R);
            }
            group_sync(mgrp);                       // every gather thread is done with the previous job's tables
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
                int4 o4; float4 w4;
                const bool i0 = valid && tp[kk].o[0] >= 0, i1 = valid && tp[kk].o[1] >= 0;
                const bool i2 = valid && tp[kk].o[2] >= 0, i3 = valid && tp[kk].o[3] >= 0;
                o4.x = i0 ? tp[kk].o[0] : 0; o4.y = i1 ? tp[kk].o[1] : 0; o4.z = i2 ? tp[kk].o[2] : 0; o4.w = i3 ? tp[kk].o[3] : 0;
                w4.x = i0 ? tp[kk].w[0] : 0.f; w4.y = i1 ? tp[kk].w[1] : 0.f; w4.z = i2 ? tp[kk].w[2] : 0.f; w4.w = i3 ? tp[kk].w[3] : 0.f;
                *reinterpret_cast<int4*>(tap_o + mt * 12 + kk * 4) = o4;
                *reinterpret_cast<float4*>(tap_w + mt * 12 + kk * 4) = w4;
                if (NORMAL) {   // d enc / d ix = wy0 (t1 - t0) + wy1 (t3 - t2),  d enc / d iy = wx0 (t2 - t0) + wx1 (t3 - t1)
                    float4 x4, y4;
                    x4.x = i0 ? -tp[kk].wy0 : 0.f; x4.y = i1 ? tp[kk].wy0 : 0.f; x4.z = i2 ? -tp[kk].wy1 : 0.f; x4.w = i3 ? tp[kk].wy1 : 0.f;
                    y4.x = i0 ? -tp[kk].wx0 : 0.f; y4.y = i1 ? -tp[kk].wx1 : 0.f; y4.z = i2 ? tp[kk].wx0 : 0.f; y4.w = i3 ? tp[kk].wx1 : 0.f;
                    *reinterpret_cast<float4*>(tap_cx + mt * 12 + kk * 4) = x4;
                    *reinterpret_cast<float4*>(tap_cy + mt * 12 + kk * 4) = y4;
                }
            }
            pbase[mt] = (uint32_t)cprompt;
            WS_ACC(1, tm, prof_m);
            mbar_wait_parity(smem_u32(sfree + g), par ^ 1u);        // the group has taken its previous tile out of the stage
            WS_ACC(0, tm, prof_m);
            *reinterpret_cast<float4*>(gs + L::META + mt * 4) =
                make_float4(__uint_as_float((uint32_t)cur.id), cx_[0], cx_[1], cx_[2]);
            group_sync(mgrp);
            // ---- cooperative gather: item = (point, 16-byte channel chunk); 2 items = 24 loads in flight per lane.
            // Blends the encoding e (-> stage) and, for the normal, the tangent rows V_a = d e / d x_a (-> L2 scratch).
            float* stage = gs + L::STAGE;
            float* vt = NORMAL ? vcta + (size_t)(g * 2 + (int)par) * L::VTILE : nullptr;
            constexpr int JB = WS_GATHER_JB, ITEMS = 128 * U;
#pragma unroll 1
            for (int i0 = mt; i0 < ITEMS; i0 += JB * WS_M) {
                float4 v[JB][12];
                int pt[JB], ch[JB];
#pragma unroll
                for (int b = 0; b < JB; ++b) {
                    const int item = i0 + b * WS_M < ITEMS ? i0 + b * WS_M : ITEMS - 1;      // tail: repeat the last item (not stored)
                    pt[b] = item / U; ch[b] = item - pt[b] * U;
                    const float* base = planes + (size_t)pbase[pt[b]] * 6 * ps + ch[b] * 4;
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) {
                        const int4 o4 = *reinterpret_cast<const int4*>(tap_o + pt[b] * 12 + kk * 4);
                        const float* pb = base + (size_t)kk * ps;
                        v[b][kk * 4 + 0] = ldg4(pb + (size_t)o4.x * C); v[b][kk * 4 + 1] = ldg4(pb + (size_t)o4.y * C);
                        v[b][kk * 4 + 2] = ldg4(pb + (size_t)o4.z * C); v[b][kk * 4 + 3] = ldg4(pb + (size_t)o4.w * C);
                    }
                }
#pragma unroll
                for (int b = 0; b < JB; ++b) {
                    if (i0 + b * WS_M >= ITEMS) continue;
                    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 V[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) V[a] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) {
                        const float4 w4 = *reinterpret_cast<const float4*>(tap_w + pt[b] * 12 + kk * 4);
                        const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
                        float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float4 q = v[b][kk * 4 + t];
                            sacc.x = fmaf(ww[t], q.x, sacc.x); sacc.y = fmaf(ww[t], q.y, sacc.y);
                            sacc.z = fmaf(ww[t], q.z, sacc.z); sacc.w = fmaf(ww[t], q.w, sacc.w);
                        }
                        e.x += sacc.x; e.y += sacc.y; e.z += sacc.z; e.w += sacc.w;
                        if (NORMAL) {
                            const float4 x4 = *reinterpret_cast<const float4*>(tap_cx + pt[b] * 12 + kk * 4);
                            const float4 y4 = *reinterpret_cast<const float4*>(tap_cy + pt[b] * 12 + kk * 4);
                            const float cx[4] = {x4.x, x4.y, x4.z, x4.w}, cy[4] = {y4.x, y4.y, y4.z, y4.w};
                            const int ax = kk == 2 ? 2 : 0, ay = kk == 1 ? 2 : 1;          // plane_ax / plane_ay, compile time
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float4 q = v[b][kk * 4 + t];
                                V[ax].x = fmaf(cx[t], q.x, V[ax].x); V[ax].y = fmaf(cx[t], q.y, V[ax].y);
                                V[ax].z = fmaf(cx[t], q.z, V[ax].z); V[ax].w = fmaf(cx[t], q.w, V[ax].w);
                                V[ay].x = fmaf(cy[t], q.x, V[ay].x); V[ay].y = fmaf(cy[t], q.y, V[ay].y);
                                V[ay].z = fmaf(cy[t], q.z, V[ay].z); V[ay].w = fmaf(cy[t], q.w, V[ay].w);
                            }
                        }
                    }
                    *reinterpret_cast<float4*>(stage + pt[b] * SP + ch[b] * 4) = e;
                    if (NORMAL) {
#pragma unroll
                        for (int a = 0; a < 3; ++a) *reinterpret_cast<float4*>(vt + L::vidx(pt[b], a, ch[b])) = V[a];
                    }
                }
            }
#ifndef TT_EMUL
            if (NORMAL) __threadfence_block();
#endif
            mbar_arrive(smem_u32(full + g));
            WS_ACC(2, tm, prof_m);
            cur = nxt;
        }
        }
    } else {
        // ============================================================================ consumer groups
        ws_reg_dec<WS_REG_C>();
        const int g = (tid - WS_GEO_MT) / TC_GROUP, tg = (tid - WS_GEO_MT) % TC_GROUP;
        Umma u;
        u.tmem = *tmem_slot + (uint32_t)g * TC_COLS_PER_GROUP;
        u.lane_base = (uint32_t)((warp & 3) * 32) << 16;
        u.mbar = smem_u32(mmab + g); u.phase = 0; u.group = WS_GEO_NMG + g;
        const bool leader = tg == 0;
        const BTile bW1 = btile_make(smem + L::W1H, smem + L::W1L, L::N1, C);
        const BTile bW2 = btile_make(smem + L::W2H, smem + L::W2L, 64, 64);
        const BTile bW2T = btile_make(smem + L::W2TH, smem + L::W2TL, 64, 64);
        const BTile bW1T = btile_make(smem + L::W1TH, smem + L::W1TL, CP, 64);
        const BTile bW2D = btile_make(smem + L::W2DH, smem + L::W2DL, 64, 64);
        const float* w3 = smem + L::W3;
        float* gs = smem + L::GROUP0 + g * L::GROUP_FLOATS;
        const float* stage = gs + L::STAGE;
        int64_t k = 0;
        const bool prof_c = blockIdx.x == 0 && g == 0 && tg == 0; (void)prof_c;
        WS_T0(tc);
        // CTAB: this thread's point of the group's NEXT tile — id and raw words prefetched one tile further, the tap tables
        // written to the group's buffer once the gather of the current tile is complete (full), then `tready`.
        const int64_t tile0 = (int64_t)blockIdx.x * WS_CG + g;
        float4 mv_next = make_float4(0.f, 0.f, 0.f, 0.f);
        WsRaw raw_next; raw_next.id = -1; int id_after = -1;
        auto publish_tables = [&](const WsRaw& r) {
            float x[3]; int prompt;
            ws_point_from_raw(src, r, x, prompt);
            const bool valid = r.id >= 0;
            float p[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
            float* ct = ctab_of(g);
            int* t_o = reinterpret_cast<int*>(ct + L::TAP_O); float* t_w = ct + L::TAP_W;
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
                const Taps tp = make_taps(p[plane_ax(kk)], p[plane_ay(kk)], cfg.R);
                int4 o4; float4 w4;
                const bool i0 = valid && tp.o[0] >= 0, i1 = valid && tp.o[1] >= 0, i2 = valid && tp.o[2] >= 0, i3 = valid && tp.o[3] >= 0;
                o4.x = i0 ? tp.o[0] : 0; o4.y = i1 ? tp.o[1] : 0; o4.z = i2 ? tp.o[2] : 0; o4.w = i3 ? tp.o[3] : 0;
                w4.x = i0 ? tp.w[0] : 0.f; w4.y = i1 ? tp.w[1] : 0.f; w4.z = i2 ? tp.w[2] : 0.f; w4.w = i3 ? tp.w[3] : 0.f;
                *reinterpret_cast<int4*>(t_o + tg * 12 + kk * 4) = o4;
                *reinterpret_cast<float4*>(t_w + tg * 12 + kk * 4) = w4;
                if (NORMAL) {   // d enc / d ix, d enc / d iy coefficients of the four texels (as the gather warps' tables)
                    float4 x4, y4;
                    x4.x = i0 ? -tp.wy0 : 0.f; x4.y = i1 ? tp.wy0 : 0.f; x4.z = i2 ? -tp.wy1 : 0.f; x4.w = i3 ? tp.wy1 : 0.f;
                    y4.x = i0 ? -tp.wx0 : 0.f; y4.y = i1 ? -tp.wx1 : 0.f; y4.z = i2 ? tp.wx0 : 0.f; y4.w = i3 ? tp.wx1 : 0.f;
                    *reinterpret_cast<float4*>(ct + L::TAP_CX + tg * 12 + kk * 4) = x4;
                    *reinterpret_cast<float4*>(ct + L::TAP_CY + tg * 12 + kk * 4) = y4;
                }
            }
            reinterpret_cast<uint32_t*>(ct + L::PBASE)[tg] = (uint32_t)prompt;
            mbar_arrive(smem_u32(tready + g));
            return make_float4(__uint_as_float((uint32_t)r.id), x[0], x[1], x[2]);
        };
        if (ctab && tile0 < n_tiles) {
            const WsRaw r0 = ws_load_raw(src, ws_load_id(src, tile0, tg, n_live, n_tiles));
            raw_next = ws_load_raw(src, ws_load_id(src, tile0 + tile_stride, tg, n_live, n_tiles));
            id_after = ws_load_id(src, tile0 + 2 * tile_stride, tg, n_live, n_tiles);
            mv_next = publish_tables(r0);
        }
        for (int64_t tile = tile0; tile < n_tiles; tile += tile_stride, ++k) {
            const uint32_t par = (uint32_t)(k & 1);
            WS_ACC(15, tc, prof_c);
            mbar_wait_parity(smem_u32(full + g), par);
            WS_ACC(8, tc, prof_c);
            if (prof_c) { g_ws_prof_tiles(); }
            const float4 mv = ctab ? mv_next : *reinterpret_cast<const float4*>(gs + L::META + tg * 4);
            const int id32 = (int)__float_as_uint(mv.x);
            const int64_t id = id32;
            {
                float e[C];
#pragma unroll
                for (int c = 0; c < C; c += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(stage + tg * SP + c);
                    e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
                }
                umma_put_A<C>(u, e);
            }
            mbar_arrive(smem_u32(sfree + g));       // row and meta are in registers / TMEM: the gather warps may refill the stage
            group_sync(u.group);
            if (leader) { umma_mma<3>(u, bW1, C, false); umma_commit(u); }
            if (ctab && tile + tile_stride < n_tiles) {      // tables of the next tile, in the shadow of the first layer
                const WsRaw r = raw_next;
                raw_next = ws_load_raw(src, id_after);
                id_after = ws_load_id(src, tile + 3 * tile_stride, tg, n_live, n_tiles);
                mv_next = publish_tables(r);
            }
            umma_wait(u);
            uint32_t m1[2], m2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {           // h1 = relu(W1 e)
                float d[32];
                ws_ld32(u, TC_COL_D + 32 * h, d);
                m1[h] = ws_pos_bits(d);
#pragma unroll
                for (int j = 0; j < 32; ++j) d[j] = fmaxf(d[j], 0.f);
                ws_st32_split(u, 32 * h, d);
            }
            tmem_wait_st();
            tc_fence_before();
            group_sync(u.group);
            if (leader) { umma_mma<3>(u, bW2, 64, false); umma_commit(u); }
            umma_wait(u);
            float s = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {           // sdf = w3 . relu(W2 h1); a2 = m2 ⊙ w3 is the next A operand
                float d[32];
                ws_ld32(u, TC_COL_D + 32 * h, d);
                m2[h] = ws_pos_bits(d);
#pragma unroll
                for (int j = 0; j < 32; ++j) s = fmaf(fmaxf(d[j], 0.f), w3[32 * h + j], s);
                if (NORMAL) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) d[j] = __uint_as_float(__float_as_uint(w3[32 * h + j]) & ws_bit_mask(m2[h], j));
                    ws_st32_split(u, 32 * h, d);
                }
            }
            const float nrm = sqrtf(mv.y * mv.y + mv.z * mv.z + mv.w * mv.w);
            if (id32 >= 0) {
                if (sdf_orig_o) sdf_orig_o[id] = s;
                if (sdf_o) sdf_o[id] = s + (nrm - cfg.sdf_bias_radius);
                if (masks_o) {     // ReLU masks for the backward
                    masks_o[id * 4 + 2] = (uint64_t)m1[0] | ((uint64_t)m1[1] << 32);
                    masks_o[id * 4 + 3] = (uint64_t)m2[0] | ((uint64_t)m2[1] << 32);
                }
            }
            if (DEFORM) {       // deformation decoder: its first layer sits in accumulator columns [192,256) since the stacked MMA
                tc_fence_before();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float d[32];
                    ws_ld32(u, TC_COL_D + 64 + 32 * h, d);
#pragma unroll
                    for (int j = 0; j < 32; ++j) d[j] = fmaxf(d[j], 0.f);
                    ws_st32_split(u, 32 * h, d);
                }
                tmem_wait_st();
                tc_fence_before();
                group_sync(u.group);
                if (leader) { umma_mma<3>(u, bW2D, 64, false); umma_commit(u); }
                umma_wait(u);
                const float* w3d = smem + L::W3D;
                float df[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float d[32];
                    ws_ld32(u, TC_COL_D + 32 * h, d);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x = fmaxf(d[j], 0.f);
                        df[0] = fmaf(x, w3d[32 * h + j], df[0]); df[1] = fmaf(x, w3d[64 + 32 * h + j], df[1]);
                        df[2] = fmaf(x, w3d[128 + 32 * h + j], df[2]);
                    }
                }
                if (id32 >= 0 && deform_o) { deform_o[id * 3] = df[0]; deform_o[id * 3 + 1] = df[1]; deform_o[id * 3 + 2] = df[2]; }
            }
            if (NORMAL) {       // unit-seed adjoint: a1 = m1 ⊙ (W2ᵀ a2), de = W1ᵀ a1;  d sdf / d x_a = de . V_a
                tmem_wait_st();
                tc_fence_before();
                group_sync(u.group);
                if (leader) { umma_mma<3>(u, bW2T, 64, false); umma_commit(u); }
                umma_wait(u);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float d[32];
                    ws_ld32(u, TC_COL_D + 32 * h, d);
#pragma unroll
                    for (int j = 0; j < 32; ++j) d[j] = __uint_as_float(__float_as_uint(d[j]) & ws_bit_mask(m1[h], j));
                    ws_st32_split(u, 32 * h, d);
                }
                tmem_wait_st();
                tc_fence_before();
                group_sync(u.group);
                if (leader) { umma_mma<3>(u, bW1T, 64, false); umma_commit(u); }
                umma_wait(u);
                float de[CP];
                umma_get_D<CP>(u, de);
                const float* vt = vcta + (size_t)(g * 2 + (int)par) * L::VTILE;
                float gm[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    float acc = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < U; ++c4) {
                        const float4 q = *reinterpret_cast<const float4*>(vt + L::vidx(tg, a, c4));
                        acc = fmaf(de[c4 * 4], q.x, fmaf(de[c4 * 4 + 1], q.y, fmaf(de[c4 * 4 + 2], q.z, fmaf(de[c4 * 4 + 3], q.w, acc))));
                    }
                    gm[a] = acc;
                }
                if (id32 >= 0 && (grad_o || normal_o)) {
                    const float scale = 0.5f * (float)cfg.R / cfg.radius;
                    const float inv = nrm > 0.f ? 1.f / nrm : 0.f;
                    const float gr[3] = {gm[0] * scale + mv.y * inv, gm[1] * scale + mv.z * inv, gm[2] * scale + mv.w * inv};
                    if (grad_o) { grad_o[id * 3] = gr[0]; grad_o[id * 3 + 1] = gr[1]; grad_o[id * 3 + 2] = gr[2]; }
                    if (normal_o) {
                        float n[3], len; normalize3(gr, n, len);
                        normal_o[id * 3] = n[0]; normal_o[id * 3 + 1] = n[1]; normal_o[id * 3 + 2] = n[2];
                    }
                }
            } else {
                tc_fence_before();
            }
            WS_ACC(9, tc, prof_c);
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
                umma_put_A_ex<C>(u, e, L::col_hi(k), L::col_lo(k));
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
