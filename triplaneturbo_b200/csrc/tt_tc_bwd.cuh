// Tensor-core (tcgen05) kernels of the backward path.  Two independent groups of 128 threads per CTA (thread = sample
// point = TMEM lane) share the weight tiles; persistent over 128-point tiles of the compacted backward lists.
//
// Nothing of the decoders is recomputed beyond what a gradient needs: the forward saved the ReLU masks of both decoders.
//   * adjoint / tangent layers: single-pass TF32 (gradient precision), A operand in TMEM,
//   * weight gradients dW = Σ_points a ⊗ x: both operands written by the threads into K-major shared-memory tiles
//     (K = the tile's 128 points), accumulated in TMEM across all tiles of the CTA and flushed once with atomicAdd.
//     The MMA shape is M = 128: only the first 64 rows of A are meaningful, the remaining rows read the B tile behind it
//     and land in accumulator rows that are never read.
//   * SDF branch: dW2 and dw3 both come out of one accumulator Q = Σ m2 ⊗ h̃1 (k_bwd_geo_tc),
//   * colour branch: first-layer gradients through hidden-gradient planes (k_bwd_tex_tc, k_hid_planes, k_hid_wgrad).
// Plane gradients: cooperative scatter (consecutive lanes = consecutive 16-byte chunks of one texel) with
// red.global.add.v4.f32.  The 64-wide colour scatter merges a tile's taps per texel on chip first when the sample lists are
// patch-ordered (coop_scatter_merged + k_patch_lists, the default with an image-shaped ray batch), else run-length merges
// consecutive samples of long rays (coop_scatter_rl) or scatters plainly.
#pragma once
#include "tt_tc.cuh"

namespace tt {

// Optional phase anatomy (-DTT_WS_TIMING): thread 0 of the gather warps and thread 0 of consumer group 0 of CTA 0 accumulate
// the cycles they spend in each phase into g_ws_prof (read back with tt_debug_ws_prof).
#if defined(TT_WS_TIMING) && !defined(TT_EMUL)
__device__ unsigned long long g_ws_prof[32];
__device__ __forceinline__ void g_ws_prof_tiles() { g_ws_prof[16] += 1; }
__device__ __forceinline__ void g_ws_prof_add(int slot, unsigned long long v) { g_ws_prof[slot] += v; }
#define WS_T0(var) long long var = clock64()
#define WS_ACC(slot, var, cond) do { if (cond) { const long long n_ = clock64(); g_ws_prof[slot] += (unsigned long long)(n_ - var); var = n_; } } while (0)
#else
__device__ __forceinline__ void g_ws_prof_tiles() {}
__device__ __forceinline__ void g_ws_prof_add(int, unsigned long long) {}
#define WS_T0(var)
#define WS_ACC(slot, var, cond)
#endif


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
        for (int k = 0; k < NPL; ++k) {
            const int4 o4 = *reinterpret_cast<const int4*>(tap_o + pt * NT + k * 4);
            const float4 w4 = *reinterpret_cast<const float4*>(tap_w + pt * NT + k * 4);
            float* pb = base + k * ps;
            if (w4.x != 0.f) red_add4(pb + (size_t)o4.x * C, make_float4(v.x * w4.x, v.y * w4.x, v.z * w4.x, v.w * w4.x));
            if (w4.y != 0.f) red_add4(pb + (size_t)o4.y * C, make_float4(v.x * w4.y, v.y * w4.y, v.z * w4.y, v.w * w4.y));
            if (w4.z != 0.f) red_add4(pb + (size_t)o4.z * C, make_float4(v.x * w4.z, v.y * w4.z, v.z * w4.z, v.w * w4.z));
            if (w4.w != 0.f) red_add4(pb + (size_t)o4.w * C, make_float4(v.x * w4.w, v.y * w4.w, v.z * w4.w, v.w * w4.w));
        }
    }
}

// Fused scatter + gather over the SAME taps with the SAME weights (SDF backward: scatter de . w, gather e~ = sum w . texel).
// Both use the item mapping of coop_gather, so a thread reads the de chunk of an item from the stage, issues the item's 12
// loads, sends the item's 12 vector reductions while the loads are in flight, then blends and overwrites the chunk with e~:
// the reduction stream and the load stream of a tile overlap instead of running one after the other, the tap tables are
// read once, and no barrier is needed between the two (a thread only touches its own items' stage chunks).
template <int C, int NPL>
__device__ __forceinline__ void coop_scatter_gather(const float* __restrict__ planes, float* __restrict__ gplanes, size_t ps,
                                                    const int* tap_o, const float* tap_w, const uint32_t* pbase, float* stage,
                                                    int tg) {
    constexpr int U = C / 4, NT = 4 * NPL, SP = C + 4, JB = TT_GATHER_LOADS / NT;
#pragma unroll 1
    for (int j0 = 0; j0 < U; j0 += JB) {
        float4 v[JB][NT];
        float w[JB][NT];
        int o[JB][NT];
        int sto[JB];
        float4 dv[JB];
        size_t pofs[JB];
#pragma unroll
        for (int b = 0; b < JB; ++b) {
            const int j = j0 + b < U ? j0 + b : U - 1;
            const int item = tg + TC_GROUP * j;
            const int pt = item / U, ch = item - pt * U;
            sto[b] = pt * SP + ch * 4;
            dv[b] = *reinterpret_cast<const float4*>(stage + sto[b]);
            pofs[b] = (size_t)pbase[pt] * 6 * ps + ch * 4;
            const float* base = planes + pofs[b];
#pragma unroll
            for (int q = 0; q < NT; q += 4) {
                const int4 o4 = *reinterpret_cast<const int4*>(tap_o + pt * NT + q);
                const float4 w4 = *reinterpret_cast<const float4*>(tap_w + pt * NT + q);
                const float* pb = base + (size_t)(q >> 2) * ps;
                v[b][q] = ldg4(pb + (size_t)o4.x * C); v[b][q + 1] = ldg4(pb + (size_t)o4.y * C);
                v[b][q + 2] = ldg4(pb + (size_t)o4.z * C); v[b][q + 3] = ldg4(pb + (size_t)o4.w * C);
                w[b][q] = w4.x; w[b][q + 1] = w4.y; w[b][q + 2] = w4.z; w[b][q + 3] = w4.w;
                o[b][q] = o4.x; o[b][q + 1] = o4.y; o[b][q + 2] = o4.z; o[b][q + 3] = o4.w;
            }
        }
        if (gplanes) {
#pragma unroll
            for (int b = 0; b < JB; ++b) {
                if (j0 + b >= U) continue;                   // tail batch repeats the last item: scatter it once
                float* gb = gplanes + pofs[b];
#pragma unroll
                for (int q = 0; q < NT; ++q) {
                    const float ww = w[b][q];
                    if (ww != 0.f)
                        red_add4(gb + (size_t)(q >> 2) * ps + (size_t)o[b][q] * C,
                                 make_float4(dv[b].x * ww, dv[b].y * ww, dv[b].z * ww, dv[b].w * ww));
                }
            }
        }
#pragma unroll
        for (int b = 0; b < JB; ++b) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NPL; ++k) {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float ww = w[b][k * 4 + t]; const float4 q = v[b][k * 4 + t];
                    s.x = fmaf(ww, q.x, s.x); s.y = fmaf(ww, q.y, s.y); s.z = fmaf(ww, q.z, s.z); s.w = fmaf(ww, q.w, s.w);
                }
                acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w;
            }
            *reinterpret_cast<float4*>(stage + sto[b]) = acc;
        }
    }
}

// Run-length merged scatter.  The points of a tile are consecutive samples of a ray (ray-major lists); inside the band
// the importance sampler concentrates on, ten or more consecutive samples fall into the same texel cell, so their
// contributions to a texel are summed in registers and leave the SM as ONE vector reduction.  Thread (ch, r) owns the
// 16-byte channel chunk ch of every texel and walks the r-th contiguous range of the tile's points with one pending
// (texel, float4) per tap slot; a pending entry is flushed when the slot moves to another texel.  The reductions of the
// CH lanes of a range stay one contiguous run of a texel.
//   dst + prompt * prompt_stride + k * plane_stride + texel * texel_stride + ch * 4   (floats)
template <int CH, int NPL>
__device__ __forceinline__ void coop_scatter_rl(float* __restrict__ dst, size_t prompt_stride, size_t plane_stride,
                                                int texel_stride, const int* tap_o, const float* tap_w,
                                                const uint32_t* pbase, const float* stage, int stage_stride, int tg) {
    constexpr int NR = TC_GROUP / CH, RL = (TC_GROUP + NR - 1) / NR, NT = 4 * NPL;
    const int ch = tg % CH, r = tg / CH;
    if (r >= NR) return;
    int cur[NT];
    float4 acc[NT];
#pragma unroll
    for (int s = 0; s < NT; ++s) { cur[s] = -1; acc[s] = make_float4(0.f, 0.f, 0.f, 0.f); }
    const int p0 = r * RL, p1 = p0 + RL < TC_GROUP ? p0 + RL : TC_GROUP;
    uint32_t pb_cur = p0 < TC_GROUP ? pbase[p0] : 0u;
    float* base = dst + (size_t)pb_cur * prompt_stride + ch * 4;
#pragma unroll 1
    for (int p = p0; p < p1; ++p) {
        const uint32_t pb = pbase[p];
        if (pb != pb_cur) {                      // the tile crosses into another prompt's planes (rare): flush everything
#pragma unroll
            for (int s = 0; s < NT; ++s) {
                if (cur[s] >= 0) red_add4(base + (size_t)(s >> 2) * plane_stride + (size_t)cur[s] * texel_stride, acc[s]);
                cur[s] = -1;
            }
            pb_cur = pb; base = dst + (size_t)pb_cur * prompt_stride + ch * 4;
        }
        const float4 v = *reinterpret_cast<const float4*>(stage + p * stage_stride + ch * 4);
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int4 o4 = *reinterpret_cast<const int4*>(tap_o + p * NT + k * 4);
            const float4 w4 = *reinterpret_cast<const float4*>(tap_w + p * NT + k * 4);
            const int oo[4] = {o4.x, o4.y, o4.z, o4.w};
            const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int s = k * 4 + t;
                const bool val = ww[t] != 0.f, same = val && oo[t] == cur[s];
                if (val && !same && cur[s] >= 0) red_add4(base + (size_t)k * plane_stride + (size_t)cur[s] * texel_stride, acc[s]);
                if (val) {
                    const float4 a = same ? acc[s] : make_float4(0.f, 0.f, 0.f, 0.f);
                    acc[s] = make_float4(fmaf(v.x, ww[t], a.x), fmaf(v.y, ww[t], a.y), fmaf(v.z, ww[t], a.z), fmaf(v.w, ww[t], a.w));
                    cur[s] = oo[t];
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NT; ++s)
        if (cur[s] >= 0) red_add4(base + (size_t)(s >> 2) * plane_stride + (size_t)cur[s] * texel_stride, acc[s]);
}


// ---- tile-merged scatter -------------------------------------------------------------------------------------------------
// The taps of a tile that fall on the SAME texel are summed inside the SM first: one vector reduction per (texel, 16-byte
// chunk) instead of one per (tap, chunk).  The scatter phases run at the L2 reduction rate, so the number of
// reductions is what they cost; with the patch-ordered sample lists (k_patch_lists: a tile = 4x4 neighbouring rays x 8
// consecutive samples) a texel receives ~8 of a tile's taps (measured on the config-3 samples), ~1.5 in ray order.
//   1. every thread inserts the texels of its point's taps into an open-addressing hash table in shared memory
//      (key = global texel number) and takes a rank inside the texel's bucket;
//   2. a scan over the table turns the bucket sizes into list offsets and compacts the used buckets;
//   3. the taps are written to their bucket's list;
//   4. item = (used bucket, 16-byte chunk): sum weight x staged row over the bucket's taps, ONE red.global.add.v4.f32.
// Workspace: MergeWs::TOTAL ints of shared memory per group (aliases an operand tile that is free during the scatter).
struct MergeWs {
    static constexpr int HN = 2048, MAXT = 1536;                 // hash positions; taps of a tile (128 points x 3 planes x 4)
    static constexpr int EB = 4;                                 // entries per block of the walk (list padded to it, read one block ahead; 8: no gain)
    static constexpr int KEY = 0, CNT = KEY + HN, ENT = CNT + HN, MISC = ENT + 2 * MAXT + 4 * EB;
    static constexpr int TOTAL = MISC + 8;
    static constexpr int KEY_BITS = 24;                          // texel numbers must fit (the point index rides above them)
};
// key(k, t) / weight(k, t): texel number (prompt, plane and offset folded: the destination row is dst + key * texel_stride)
// and weight of tap t of plane k of THIS thread's point; weight 0 = no tap.
template <int CH, int NPL, typename KeyF, typename WF>
__device__ __forceinline__ void coop_scatter_merged(float* __restrict__ dst, int texel_stride, int* ws, const float* stage,
                                                    int stage_stride, int tg, int group, KeyF key_of, WF w_of) {
    constexpr int NT = NPL * 4, HN = MergeWs::HN, KB = MergeWs::KEY_BITS;
    static_assert(128 * NT <= MergeWs::MAXT, "workspace");
    static_assert(TC_GROUP % CH == 0, "ranges");
    int* hkey = ws + MergeWs::KEY; int* hcnt = ws + MergeWs::CNT;
    int2* ent = reinterpret_cast<int2*>(ws + MergeWs::ENT);           // bucket after bucket: (texel | point << 24, weight)
    int* misc = ws + MergeWs::MISC;
    const bool prof = blockIdx.x == 0 && threadIdx.x == 0; (void)prof;
    WS_T0(tm_);
    for (int i = tg * 4; i < HN; i += TC_GROUP * 4) {
        *reinterpret_cast<int4*>(hkey + i) = make_int4(-1, -1, -1, -1);
        *reinterpret_cast<int4*>(hcnt + i) = make_int4(0, 0, 0, 0);
    }
    group_sync(group);
    WS_ACC(24, tm_, prof);
    // (the three passes keep the shared-memory atomics of the 12 taps independent of each other: they pipeline)
    int hpos[NT], aux[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        hpos[j] = -1; aux[j] = 0;
        if (w_of(j >> 2, j & 3) != 0.f) {
            const int key = key_of(j >> 2, j & 3);
            hpos[j] = (int)(((uint32_t)key * 2654435761u) >> 21);
            aux[j] = atomicCAS(hkey + hpos[j], -1, key);
        }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        if (hpos[j] >= 0 && aux[j] != -1) {
            const int key = key_of(j >> 2, j & 3);
            int old = aux[j];
            while (old != -1 && old != key) {                     // linear probing
                hpos[j] = (hpos[j] + 1) & (HN - 1);
                old = atomicCAS(hkey + hpos[j], -1, key);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) aux[j] = hpos[j] >= 0 ? atomicAdd(hcnt + hpos[j], 1) : 0;       // rank in the bucket
    group_sync(group);
    WS_ACC(25, tm_, prof);
    {   // exclusive scan of the bucket sizes = offsets of the buckets in the entry list, 16 positions per thread
        constexpr int PER = HN / TC_GROUP;
        int c[PER], s = 0;
#pragma unroll
        for (int i = 0; i < PER; i += 4) {
            const int4 q = *reinterpret_cast<const int4*>(hcnt + tg * PER + i);
            c[i] = q.x; c[i + 1] = q.y; c[i + 2] = q.z; c[i + 3] = q.w;
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) s += c[i];
        int v = s;
        const int lane = tg & 31, wrp = tg >> 5;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, v, off); if (lane >= off) v += o; }
        if (lane == 31) misc[wrp] = v;
        group_sync(group);
        int off = v - s;
        for (int q = 0; q < wrp; ++q) off += misc[q];
#pragma unroll
        for (int i = 0; i < PER; i += 4) {
            int4 q;
            q.x = off; off += c[i]; q.y = off; off += c[i + 1]; q.z = off; off += c[i + 2]; q.w = off; off += c[i + 3];
            *reinterpret_cast<int4*>(hcnt + tg * PER + i) = q;
        }
        if (tg == TC_GROUP - 1) misc[4] = off;            // entries of the tile
    }
    group_sync(group);
    // entry = (texel | point << 24, weight); the FIRST entry of a bucket carries the sign bit in its weight (bilinear weights
    // are >= 0), which is all the walk below needs to find the runs
#pragma unroll
    for (int j = 0; j < NT; ++j)
        if (hpos[j] >= 0)
            ent[hcnt[hpos[j]] + aux[j]] = make_int2(key_of(j >> 2, j & 3) | (tg << KB),
                                                    (int)(__float_as_uint(w_of(j >> 2, j & 3)) | (aux[j] == 0 ? 0x80000000u : 0u)));
    constexpr int EB = MergeWs::EB;
    {   // pad the list to a multiple of EB entries (weight 0, no run start)
        const int total = misc[4];
        if (tg < EB && total + tg < ((total + EB - 1) / EB) * EB) ent[total + tg] = make_int2(0, 0);
    }
    group_sync(group);
    WS_ACC(26, tm_, prof);
    // The entry list is cut into TC_GROUP / CH ranges of whole EB-entry blocks; thread (range, 16-byte chunk) walks its range,
    // sums weight x staged row while the texel stays the same and issues ONE vector reduction per run (a bucket that straddles
    // two ranges costs two).  A block's staged rows are loaded together, its entries one block ahead.
    {
        constexpr int NRANGE = TC_GROUP / CH;
        const int total = misc[4];
        if (prof) g_ws_prof_add(28, (unsigned long long)total);
        const int blocks = (total + EB - 1) / EB;
        const int r = tg / CH, ch = tg - r * CH;
        const int e0 = ((blocks * r) / NRANGE) * EB, e1 = ((blocks * (r + 1)) / NRANGE) * EB;
        const float* srow = stage + ch * 4;
        float* drow = dst + ch * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int cur = -1;
        if (e0 < e1) { const int2 f = ent[e0]; cur = f.y < 0 ? -1 : (f.x & ((1 << KB) - 1)); }      // range starts inside a bucket
        int2 en[EB];
#pragma unroll
        for (int q = 0; q < EB; ++q) en[q] = ent[e0 + q];            // (the list has 2 blocks of slack: reading ahead is safe)
#pragma unroll 1
        for (int e = e0; e < e1; e += EB) {
            float4 v[EB]; int2 nx[EB];
#pragma unroll
            for (int q = 0; q < EB; ++q) v[q] = *reinterpret_cast<const float4*>(srow + ((uint32_t)en[q].x >> KB) * stage_stride);
#pragma unroll
            for (int q = 0; q < EB; ++q) nx[q] = ent[e + EB + q];    // next block, one iteration ahead of its use
#pragma unroll
            for (int q = 0; q < EB; ++q) {
                if (en[q].y < 0) {                                   // a new texel starts: flush the finished run
                    if (cur >= 0) red_add4(drow + (size_t)cur * texel_stride, acc);
                    cur = en[q].x & ((1 << KB) - 1);
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const float w = fabsf(__uint_as_float((uint32_t)en[q].y));
                acc.x = fmaf(w, v[q].x, acc.x); acc.y = fmaf(w, v[q].y, acc.y);
                acc.z = fmaf(w, v[q].z, acc.z); acc.w = fmaf(w, v[q].w, acc.w);
            }
#pragma unroll
            for (int q = 0; q < EB; ++q) en[q] = nx[q];
        }
        if (cur >= 0) red_add4(drow + (size_t)cur * texel_stride, acc);
    }
    group_sync(group);
    WS_ACC(27, tm_, prof);
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
    int* mlock = reinterpret_cast<int*>(tmem_slot + 1);
    if (tid == 0) { for (int g = 0; g < G; ++g) mbar_init(mbars + g); *mlock = 0; }
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

    // per-sample inputs one tile ahead, sample id two tiles ahead (as k_bwd_tex_tc; -DTT_BWDGEO_PREFETCH=0: loaded in place)
#ifndef TT_BWDGEO_PREFETCH
#define TT_BWDGEO_PREFETCH 1
#endif
    struct Pre { WsRaw raw; float gs, uu[3]; uint64_t m1, m2; };
    const int64_t tstride = (int64_t)gridDim.x * G, tile0 = (int64_t)blockIdx.x * G + group;
    auto load_pre = [&](int id) {
        Pre q; q.raw = ws_load_raw(src, id); q.gs = q.uu[0] = q.uu[1] = q.uu[2] = 0.f; q.m1 = q.m2 = 0ull;
        if (id >= 0) {
            q.gs = gs_i[id]; q.uu[0] = u_i[(int64_t)id * 3]; q.uu[1] = u_i[(int64_t)id * 3 + 1]; q.uu[2] = u_i[(int64_t)id * 3 + 2];
            q.m1 = masks[(int64_t)id * 4 + 2]; q.m2 = masks[(int64_t)id * 4 + 3];
        }
        return q;
    };
    Pre pre_next; int id_after = -1;
    if (TT_BWDGEO_PREFETCH) {
        pre_next = load_pre(ws_load_id(src, tile0, tg, n_live, n_tiles));
        id_after = ws_load_id(src, tile0 + tstride, tg, n_live, n_tiles);
    }
    for (int64_t tile = tile0; tile < n_tiles; tile += tstride) {
        Pre cur;
        if (TT_BWDGEO_PREFETCH) {
            cur = pre_next;
            pre_next = load_pre(id_after);
            id_after = ws_load_id(src, tile + 2 * tstride, tg, n_live, n_tiles);
        } else cur = load_pre(ws_load_id(src, tile, tg, n_live, n_tiles));
        const bool valid = cur.raw.id >= 0;
        const float gs = cur.gs, uu[3] = {cur.uu[0], cur.uu[1], cur.uu[2]};
        const bool active = valid && (gs != 0.f || uu[0] != 0.f || uu[1] != 0.f || uu[2] != 0.f);
        const uint64_t m1 = active ? cur.m1 : 0ull, m2 = active ? cur.m2 : 0ull;
        int prompt = 0;
        {
            float x[3] = {0.f, 0.f, 0.f}, p[3];
            if (active) ws_point_from_raw(src, cur.raw, x, prompt);
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
#ifndef TT_BWDGEO_FUSED
#define TT_BWDGEO_FUSED 1
#endif
#if TT_BWDGEO_FUSED
        // d L / d texel = de . w  scattered and  e~ = sum w . texel  gathered in one pass over the taps
        mem_lock(mlock, leader, group);
        coop_scatter_gather<C, 3>(planes, gplanes, ps, tap_o, tap_om, pbase, stage, tg);
#else
        mem_lock(mlock, leader, group);
        // d L / d texel = de · ω  (plain scatter: with 12 tap slots the run-length merged variant measured 4 % slower on
        // the synthetic benchmark planes, whose noisy SDF spreads the fine samples; it wins for the 64-wide colour scatter)
        if (gplanes) coop_scatter<C, 3>(gplanes, ps, tap_o, tap_om, pbase, 0, stage, tg);
        group_sync(group);
        coop_gather<C, 3>(planes, ps, tap_o, tap_om, pbase, 0, stage, tg);                    // ẽ = Σ ω · texel
#endif
        mem_unlock(mlock, leader, group);
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
// First-layer gradients through a HIDDEN-GRADIENT PLANE.  With g1 = dL/d(pre-activation of layer 1) [64] per sample,
// e_k = Σ_t w_t · texel_k(o_t) the sampled encoding of texture plane k and W1_k the [64][C] block of the first layer:
//     dL/dtexel_k(o) = Σ_{samples, taps hitting o} w_t · W1_kᵀ g1      =  (H_k W1_k)(o)
//     dL/dW1_k        = Σ_samples g1 ⊗ e_k                              =  H_kᵀ · planes_k
// where H_k(o)[64] = Σ w_t · g1 is the scatter of the 64-wide hidden gradient.  So the per-sample work of the first layer
// is ONE scatter (no second gather of e_k, no W1ᵀ layer, no per-tile dW1 contraction), and the two products above are
// dense [R²x64]·[64xC] / [64xR²]·[R²xC] contractions done once per backward by k_hid_planes / k_hid_wgrad (fp32 FMA).
// Per tile: gather e_k (3 planes) -> h1 = relu(Σ W1_k e_k) -> z2 = W2 h1 (single-pass TF32, masks from the forward),
// g2 = m2 ⊙ W3ᵀ gf, dW2 += g2 h1ᵀ (TMEM accumulator), g1 = m1 ⊙ W2ᵀ g2, dW3 += gf h2ᵀ (SIMT from the h2 operand tile),
// scatter g1·w into H.  G independent 128-thread groups per CTA; TMEM per group: A [0,64) dW2 [64,128) D [128,192).
template <int C, bool P3 = false>
struct BwdTexSmem {
    static constexpr int HS = 68;                                               // row stride of the staged g1 [128][64]
    // weight tiles as tf32 hi + exact remainder: the recomputed activations and the adjoint layer run as 3xTF32 like the
    // forward (round 2, DESIGN 4.2: with single-pass layers the colour decoder's weight gradients were 200x less accurate
    // than an fp32 run; the kernel is bound by its vector reductions, the extra MMAs are free)
    // P3 (TT_FLAG_PRECISE_BWD): the lo tiles exist and the layers run as 3xTF32; the operand tiles of the weight-gradient
    // contraction (2 x 36 KB per group) then leave room for ONE group per CTA only: measured 651 vs 432 ms at config 3.
    static constexpr int W1H = 0, W1L = W1H + 3 * 64 * C, W2H = W1L + (P3 ? 3 * 64 * C : 0), W2L = W2H + 4096;
    static constexpr int W2TH = W2L + (P3 ? 4096 : 0), W2TL = W2TH + 4096, W3 = W2TL + (P3 ? 4096 : 0);
    static constexpr int GROUP0 = W3 + 192;
    static constexpr uint32_t COL_ALO = 192;                                    // TMEM: A hi [0,64) dW2 [64,128) D [128,192) A lo [192,256)
    static constexpr int AT = 0, BT = AT + wg_tile_floats(64), STAGE = BT;      // STAGE aliases the B tile
    static constexpr int TAP_O = BT + wg_tile_floats(64), TAP_W = TAP_O + 128 * 4, PBASE = TAP_W + 128 * 4;
    static constexpr int GF = PBASE + 128;
    static constexpr int GROUP_FLOATS = GF + 3 * 128;
    static constexpr int G = (GROUP0 + 2 * GROUP_FLOATS + 16) * 4 <= 227 * 1024 ? 2 : 1;
    static constexpr int TOTAL = GROUP0 + G * GROUP_FLOATS + 16;
    static constexpr uint32_t COL_GW2 = 64;
    static_assert(128 * (C + 4) <= wg_tile_floats(64) && 128 * HS <= wg_tile_floats(64), "stage must fit in the B tile");
};

// RL: run-length merged scatter of the hidden gradient (wins when many consecutive samples of a ray share texel cells:
// config 2, 289 samples per ray: 109 vs 113 ms) or the plain one (config 3, 193 samples per ray at 512^2: 432 vs 460 ms).
// SC: 0 plain, 1 run-length merged, 2 tile-merged (coop_scatter_merged; its workspace aliases the A tile of the dW2
// contraction, whose MMAs have completed before the scatter)
template <int C, int SC, bool P3>
__global__ void __launch_bounds__(BwdTexSmem<C, P3>::G * TC_GROUP, 1)
k_bwd_tex_tc(const float* __restrict__ planes, const float* __restrict__ wp, tt_config cfg, TcSrc src, int64_t N,
             const float* __restrict__ gf_i, const uint64_t* __restrict__ masks, float* __restrict__ hid,
             float* __restrict__ gw) {
    TT_SHARED(smem);
    using L = BwdTexSmem<C, P3>;
    constexpr int SP = C + 4, HS = L::HS, G = L::G, NT = G * TC_GROUP;
    constexpr int PASSES = P3 ? 3 : 1;
    const int tid = threadIdx.x, group = tid / TC_GROUP, tg = tid % TC_GROUP, warp = tid >> 5;
    const WOff wo = woff(C);
    const GOff go = goff(C);
    if (P3) {
        for (int k = 0; k < 3; ++k)      // W1f as three [64][C] K-major tiles (one per texture plane), tf32 hi + remainder
            btile_fill(smem + L::W1H + k * 64 * C, smem + L::W1L + k * 64 * C, 64, C,
                       [&](int n, int kk) { return __ldg(wp + wo.w1f + n * 3 * C + k * C + kk); }, tid, NT);
        btile_fill(smem + L::W2H, smem + L::W2L, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2f + n * 64 + k); }, tid, NT);
        btile_fill(smem + L::W2TH, smem + L::W2TL, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2f + k * 64 + n); }, tid, NT);
    } else {
        for (int i = tid; i < 3 * 64 * C; i += NT) {          // single pass: round-to-nearest tf32 weights
            const int k = i / (64 * C), r = i - k * 64 * C, n = r / C, kk = r % C;
            smem[L::W1H + k * 64 * C + btile_off(n, kk, C)] = tf32_rn(__ldg(wp + wo.w1f + n * 3 * C + k * C + kk));
        }
        for (int i = tid; i < 4096; i += NT) {
            const int n = i / 64, k = i % 64;
            smem[L::W2H + btile_off(n, k, 64)] = tf32_rn(__ldg(wp + wo.w2f + n * 64 + k));
            smem[L::W2TH + btile_off(n, k, 64)] = tf32_rn(__ldg(wp + wo.w2f + k * 64 + n));
        }
    }
    for (int i = tid; i < 192; i += NT) smem[L::W3 + i] = __ldg(wp + wo.w3f + i);
    float* gsm = smem + L::GROUP0 + group * L::GROUP_FLOATS;
    for (int i = tg; i < 2 * wg_tile_floats(64); i += TC_GROUP) gsm[L::AT + i] = 0.f;
    uint64_t* mbars = reinterpret_cast<uint64_t*>(smem + L::GROUP0 + G * L::GROUP_FLOATS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbars + G);
    int* mlock = reinterpret_cast<int*>(tmem_slot + 1);
    if (tid == 0) { for (int g = 0; g < G; ++g) mbar_init(mbars + g); *mlock = 0; }
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
    BTile bW1[3];
    for (int k = 0; k < 3; ++k) bW1[k] = btile_make(smem + L::W1H + k * 64 * C, smem + L::W1L + k * 64 * C, 64, C);
    const BTile bW2 = btile_make(smem + L::W2H, smem + L::W2L, 64, 64);
    const BTile bW2T = btile_make(smem + L::W2TH, smem + L::W2TL, 64, 64);
    float* At = gsm + L::AT; float* Bt = gsm + L::BT;
    const uint32_t at_addr = smem_u32(At), bt_addr = smem_u32(Bt);
    int* tap_o = reinterpret_cast<int*>(gsm + L::TAP_O);
    float* tap_w = gsm + L::TAP_W;
    uint32_t* pbase = reinterpret_cast<uint32_t*>(gsm + L::PBASE);
    float* gfs = gsm + L::GF;
    float* stage = gsm + L::STAGE;
    const float* w3 = smem + L::W3;
    const size_t ps = (size_t)cfg.R * cfg.R * C, hs = (size_t)cfg.R * cfg.R * 64;
    const int64_t n_live = src.count ? (int64_t)*src.count : N;
    const int64_t n_tiles = (n_live + TC_GROUP - 1) / TC_GROUP;
    bool any_tile = false;
    float dw3[3] = {0.f, 0.f, 0.f};            // dW3[c][j] partial: j = tg & 63 over the points of half tg >> 6
    const int j3 = tg & 63, half3 = tg >> 6;

    const bool prof = blockIdx.x == 0 && tid == 0; (void)prof;
    WS_T0(tp_);
    // Per-sample inputs of a tile (seed, ReLU masks, raw position words) are loaded ONE tile ahead and the sample id two
    // tiles ahead: their dependent global loads (id -> seed / masks / ray) would otherwise be exposed at every tile start
    // (anatomy: 3.1 k of 36 k cycles per tile).
    struct Pre { WsRaw raw; float gf[3]; uint64_t m1, m2; };
    const int64_t tstride = (int64_t)gridDim.x * G, tile0 = (int64_t)blockIdx.x * G + group;
    auto load_pre = [&](int id) {
        Pre q; q.raw = ws_load_raw(src, id); q.gf[0] = q.gf[1] = q.gf[2] = 0.f; q.m1 = q.m2 = 0ull;
        if (id >= 0) {
            q.gf[0] = gf_i[(int64_t)id * 3]; q.gf[1] = gf_i[(int64_t)id * 3 + 1]; q.gf[2] = gf_i[(int64_t)id * 3 + 2];
            q.m1 = masks[(int64_t)id * 4]; q.m2 = masks[(int64_t)id * 4 + 1];
        }
        return q;
    };
    Pre pre_next = load_pre(ws_load_id(src, tile0, tg, n_live, n_tiles));
    int id_after = ws_load_id(src, tile0 + tstride, tg, n_live, n_tiles);
    for (int64_t tile = tile0; tile < n_tiles; tile += tstride) {
        const Pre cur = pre_next;
        pre_next = load_pre(id_after);                                  // in flight during this tile
        id_after = ws_load_id(src, tile + 2 * tstride, tg, n_live, n_tiles);
        const bool valid = cur.raw.id >= 0;
        const float gf[3] = {cur.gf[0], cur.gf[1], cur.gf[2]};
        const bool active = valid && (gf[0] != 0.f || gf[1] != 0.f || gf[2] != 0.f);
        if (prof) g_ws_prof_tiles();
        // the ReLU masks are the forward's (3xTF32) masks: a single-pass recompute may flip units near zero
        const uint64_t m1 = active ? cur.m1 : 0ull, m2 = active ? cur.m2 : 0ull;
        int prompt = 0;
        float p[3];
        {
            float x[3] = {0.f, 0.f, 0.f};
            if (active) ws_point_from_raw(src, cur.raw, x, prompt);
#pragma unroll
            for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
        }
        pbase[tg] = (uint32_t)prompt;
        gfs[tg] = gf[0]; gfs[128 + tg] = gf[1]; gfs[256 + tg] = gf[2];
        float d[64];
        WS_ACC(20, tp_, prof);
        // ---- recompute (single pass): z1 = Σ_k W1f_k e_k, z2 = W2f relu(z1) -----------------------------------------
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            {
                const Taps t = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
                int4 o4; float4 w4;
                o4.x = (active && t.o[0] >= 0) ? t.o[0] : 0; w4.x = (active && t.o[0] >= 0) ? t.w[0] : 0.f;
                o4.y = (active && t.o[1] >= 0) ? t.o[1] : 0; w4.y = (active && t.o[1] >= 0) ? t.w[1] : 0.f;
                o4.z = (active && t.o[2] >= 0) ? t.o[2] : 0; w4.z = (active && t.o[2] >= 0) ? t.w[2] : 0.f;
                o4.w = (active && t.o[3] >= 0) ? t.o[3] : 0; w4.w = (active && t.o[3] >= 0) ? t.w[3] : 0.f;
                *reinterpret_cast<int4*>(tap_o + tg * 4) = o4;
                *reinterpret_cast<float4*>(tap_w + tg * 4) = w4;
            }
            group_sync(group);
            coop_gather<C, 1>(planes, ps, tap_o, tap_w, pbase, 3 + k, stage, tg);
            group_sync(group);
            float e[C];
#pragma unroll
            for (int c = 0; c < C; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(stage + tg * SP + c);
                e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
            }
            if (k > 0) umma_wait(u);          // previous chunk's MMAs are done reading A
            if (P3) umma_put_A_ex<C>(u, e, TC_COL_AHI, L::COL_ALO); else umma_put_A1<C>(u, e);
            group_sync(group);
            if (leader) { umma_mma_ex<PASSES>(u, bW1[k], C, k > 0, TC_COL_AHI, L::COL_ALO, TC_COL_D); umma_commit(u); }
        }
        umma_wait(u);
        umma_get_D<64>(u, d);
        WS_ACC(21, tp_, prof);
        // ---- g2 = m2 ⊙ W3ᵀ gf needs only the masks: dW2 += g2 h1ᵀ is issued together with z2 = W2 h1 ------------------
        {
            float h[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                h[j] = ((m1 >> j) & 1ull) ? d[j] : 0.f;
                Bt[wg_off(j, tg)] = tf32_rn(h[j]);                                   // h1 -> B tile of dW2
                const float v = gf[0] * w3[j] + gf[1] * w3[64 + j] + gf[2] * w3[128 + j];
                At[wg_off(j, tg)] = ((m2 >> j) & 1ull) ? tf32_rn(v) : 0.f;           // g2 -> A tile of dW2
            }
            if (P3) umma_put_A_ex<64>(u, h, TC_COL_AHI, L::COL_ALO); else umma_put_A1<64>(u, h);
            async_proxy_fence();
            group_sync(group);
            if (leader) {
                if (gw) umma_mma_ss(u, at_addr, bt_addr, 64, L::COL_GW2, any_tile);
                umma_mma_ex<PASSES>(u, bW2, 64, false, TC_COL_AHI, L::COL_ALO, TC_COL_D);
                umma_commit(u);
            }
            umma_wait(u);
            umma_get_D<64>(u, d);                                                    // z2
        }
        // ---- dW3 += gf h2ᵀ from the h2 operand tile (SIMT: 3 x 64 outputs, 64 points per thread) ---------------------
        if (gw) {
#pragma unroll
            for (int j = 0; j < 64; ++j) Bt[wg_off(j, tg)] = ((m2 >> j) & 1ull) ? d[j] : 0.f;      // h2
        }
        // ---- g1 = m1 ⊙ W2ᵀ g2 ---------------------------------------------------------------------------------------
        float g1[64];
        {
            float g2[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const float v = gf[0] * w3[j] + gf[1] * w3[64 + j] + gf[2] * w3[128 + j];
                g2[j] = ((m2 >> j) & 1ull) ? v : 0.f;
            }
            if (P3) umma_put_A_ex<64>(u, g2, TC_COL_AHI, L::COL_ALO); else umma_put_A1<64>(u, g2);
            group_sync(u.group);                                 // (also publishes the h2 tile)
            if (leader) { umma_mma_ex<PASSES>(u, bW2T, 64, false, TC_COL_AHI, L::COL_ALO, TC_COL_D); umma_commit(u); }
            umma_wait(u);
            umma_get_D<64>(u, g1);
        }
        if (gw) {
#pragma unroll 4
            for (int q = 0; q < 16; ++q) {
                const int p0 = half3 * 64 + q * 4;
                const float4 hv = *reinterpret_cast<const float4*>(Bt + wg_off(j3, p0));
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 gv = *reinterpret_cast<const float4*>(gfs + c * 128 + p0);
                    dw3[c] = fmaf(gv.x, hv.x, fmaf(gv.y, hv.y, fmaf(gv.z, hv.z, fmaf(gv.w, hv.w, dw3[c]))));
                }
            }
        }
        group_sync(group);                     // the B tile becomes the stage of g1
#pragma unroll
        for (int j = 0; j < 64; j += 4)
            *reinterpret_cast<float4*>(stage + tg * HS + j) =
                make_float4(((m1 >> j) & 1ull) ? g1[j] : 0.f, ((m1 >> (j + 1)) & 1ull) ? g1[j + 1] : 0.f,
                            ((m1 >> (j + 2)) & 1ull) ? g1[j + 2] : 0.f, ((m1 >> (j + 3)) & 1ull) ? g1[j + 3] : 0.f);
        WS_ACC(22, tp_, prof);
        // ---- scatter g1 · w into the hidden-gradient planes ----------------------------------------------------------
        if (SC == 2) {
            static_assert(MergeWs::TOTAL <= wg_tile_floats(64), "merge workspace must fit in the A tile");
            Taps t3[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) t3[k] = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
            const int kbase = prompt * 3 * cfg.R * cfg.R, rr = cfg.R * cfg.R;
            coop_scatter_merged<16, 3>(hid, 64, reinterpret_cast<int*>(At), stage, HS, tg, group,
                [&](int k, int t) { return kbase + k * rr + t3[k].o[t]; },
                [&](int k, int t) { return (active && t3[k].o[t] >= 0) ? t3[k].w[t] : 0.f; });
        } else
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            {
                const Taps t = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
                int4 o4; float4 w4;
                o4.x = (active && t.o[0] >= 0) ? t.o[0] : 0; w4.x = (active && t.o[0] >= 0) ? t.w[0] : 0.f;
                o4.y = (active && t.o[1] >= 0) ? t.o[1] : 0; w4.y = (active && t.o[1] >= 0) ? t.w[1] : 0.f;
                o4.z = (active && t.o[2] >= 0) ? t.o[2] : 0; w4.z = (active && t.o[2] >= 0) ? t.w[2] : 0.f;
                o4.w = (active && t.o[3] >= 0) ? t.o[3] : 0; w4.w = (active && t.o[3] >= 0) ? t.w[3] : 0.f;
                *reinterpret_cast<int4*>(tap_o + tg * 4) = o4;
                *reinterpret_cast<float4*>(tap_w + tg * 4) = w4;
            }
            group_sync(group);
            if (SC == 1) coop_scatter_rl<16, 1>(hid + (size_t)k * hs, 3 * hs, 0, 64, tap_o, tap_w, pbase, stage, HS, tg);
            else {   // plain scatter of the 64-wide hidden gradient: item = (point, 16-byte chunk), 4 vector reductions each
#pragma unroll 4
                for (int j = 0; j < 16; ++j) {
                    const int item = tg + TC_GROUP * j;
                    const int pt = item >> 4, ch = item & 15;
                    const float4 v = *reinterpret_cast<const float4*>(stage + pt * HS + ch * 4);
                    const int4 o4 = *reinterpret_cast<const int4*>(tap_o + pt * 4);
                    const float4 w4 = *reinterpret_cast<const float4*>(tap_w + pt * 4);
                    float* pb = hid + (size_t)pbase[pt] * 3 * hs + (size_t)k * hs + ch * 4;
                    if (w4.x != 0.f) red_add4(pb + (size_t)o4.x * 64, make_float4(v.x * w4.x, v.y * w4.x, v.z * w4.x, v.w * w4.x));
                    if (w4.y != 0.f) red_add4(pb + (size_t)o4.y * 64, make_float4(v.x * w4.y, v.y * w4.y, v.z * w4.y, v.w * w4.y));
                    if (w4.z != 0.f) red_add4(pb + (size_t)o4.z * 64, make_float4(v.x * w4.z, v.y * w4.z, v.z * w4.z, v.w * w4.z));
                    if (w4.w != 0.f) red_add4(pb + (size_t)o4.w * 64, make_float4(v.x * w4.w, v.y * w4.w, v.z * w4.w, v.w * w4.w));
                }
            }
            group_sync(group);
        }
        WS_ACC(23, tp_, prof);
        any_tile = true;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (gw && any_tile) {
        if (tg < 64) {
            float g[64];
            umma_get_D<64>(u, g, L::COL_GW2);
#pragma unroll
            for (int j = 0; j < 64; ++j) if (g[j] != 0.f) atomicAdd(gw + go.g2f + tg * 64 + j, g[j]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) if (dw3[c] != 0.f) atomicAdd(gw + go.g3f + c * 64 + j3, dw3[c]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_warp(*tmem_slot, G * TC_COLS_PER_GROUP);
}

// ---- dense products with the hidden-gradient planes (once per backward) ----------------------------------------------
// gplanes[p][3+k][o][c] += Σ_j hid[p][k][o][j] · W1f[j][kC + c]
template <int C>
__global__ void __launch_bounds__(256) k_hid_planes(const float* __restrict__ hid, const float* __restrict__ wp, int P, int R,
                                                   float* __restrict__ gplanes) {
    TT_SHARED(smem);                          // W1f_k as [64][C]
    constexpr int U = C / 4;
    const int k = blockIdx.y;
    const WOff wo = woff(C);
    for (int i = threadIdx.x; i < 64 * C; i += 256) { const int j = i / C, c = i % C; smem[i] = __ldg(wp + wo.w1f + j * 3 * C + k * C + c); }
    __syncthreads();
    const int64_t RR = (int64_t)R * R, items = (int64_t)P * RR * U;
    for (int64_t it = (int64_t)blockIdx.x * 256 + threadIdx.x; it < items; it += (int64_t)gridDim.x * 256) {
        const int64_t cell = it / U; const int ch = (int)(it - cell * U);
        const int64_t p = cell / RR, o = cell - p * RR;
        const float* h = hid + ((p * 3 + k) * RR + o) * 64;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        bool any = false;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
            const float4 hv = ldg4(h + j);
            any = any || hv.x != 0.f || hv.y != 0.f || hv.z != 0.f || hv.w != 0.f;
            const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(smem + (j + q) * C + ch * 4);
                acc.x = fmaf(hh[q], w.x, acc.x); acc.y = fmaf(hh[q], w.y, acc.y);
                acc.z = fmaf(hh[q], w.z, acc.z); acc.w = fmaf(hh[q], w.w, acc.w);
            }
        }
        if (any) {
            float4* dst = reinterpret_cast<float4*>(gplanes + ((p * 6 + 3 + k) * RR + o) * C + ch * 4);
            float4 g = *dst;
            g.x += acc.x; g.y += acc.y; g.z += acc.z; g.w += acc.w;
            *dst = g;
        }
    }
}
// gw.W1f[j][kC + c] += Σ_{p,o} hid[p][k][o][j] · planes[p][3+k][o][c]
template <int C>
__global__ void __launch_bounds__(256) k_hid_wgrad(const float* __restrict__ hid, const float* __restrict__ planes, int P,
                                                  int R, float* __restrict__ gw) {
    TT_SHARED(smem);
    constexpr int TX = 64, GS = 65, PS = C + 1, NC = C / 4;
    float* Gs = smem; float* Ps = smem + TX * GS;
    const int k = blockIdx.y, j = threadIdx.x & 63, cq = threadIdx.x >> 6;
    const GOff go = goff(C);
    const int64_t RR = (int64_t)R * R, spp = (RR + TX - 1) / TX, slabs = (int64_t)P * spp;
    float acc[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = 0.f;
    for (int64_t sl = blockIdx.x; sl < slabs; sl += gridDim.x) {
        const int64_t p = sl / spp, o0 = (sl - p * spp) * TX;
        const int nx = (int)(RR - o0 < TX ? RR - o0 : TX);              // texels in this slab
        const float* h = hid + ((p * 3 + k) * RR + o0) * 64;
        const float* pl = planes + ((p * 6 + 3 + k) * RR + o0) * C;
        bool any = false;
        for (int i = threadIdx.x; i < TX * 64; i += 256) {
            const float v = i < nx * 64 ? __ldg(h + i) : 0.f;
            any = any || v != 0.f; Gs[(i >> 6) * GS + (i & 63)] = v;
        }
        for (int i = threadIdx.x; i < TX * C; i += 256) Ps[(i / C) * PS + (i % C)] = i < nx * C ? __ldg(pl + i) : 0.f;
        if (__syncthreads_or(any)) {
#pragma unroll 4
            for (int x = 0; x < TX; ++x) {
                const float g = Gs[x * GS + j];
#pragma unroll
                for (int i = 0; i < NC; ++i) acc[i] = fmaf(g, Ps[x * PS + cq + 4 * i], acc[i]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) if (acc[i] != 0.f) atomicAdd(gw + go.g1f + j * 3 * C + k * C + cq + 4 * i, acc[i]);
}

}  // namespace tt
