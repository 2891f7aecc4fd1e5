// Per-ray kernels of the rendering path, one WARP per ray: the 32 lanes take 32 consecutive samples of the ray, so every
// per-sample array ([n_rays][S], ray-major) is read and written with coalesced accesses, and the two recurrences along
// the ray become warp scans:
//   forward  T_i = Π_{j<i} (1 - α_j)                       -> exclusive product scan
//   backward R_{i-1} = gw_i α_i + (1 - α_i) R_i             -> scan of affine maps R -> a R + b, far end first
// (a thread-per-ray loop reads each array with a stride of S floats between lanes: ~12x the HBM time).
//   k_weights          a12-a16 minus colour: NeuS alpha, transmittance, weights, opacity / depth / normal / z-variance /
//                      eikonal accumulators, ordered list of the live samples (T > 0, non-empty point)
//   k_accum_rgb        a10 + a16 (rgb)
//   k_render_bwd_comp  backward of a14-a16 (+ eikonal, rgb_grad_shrink, inv_std) -> per-sample seeds gs, u, gf and the
//                      ordered lists of the samples the decoder backward has to visit
// Reference: custom/triplaneturbo/models/renderers/generative_space_sdf_volume_renderer.py:397-431,466-472,
// threestudio/models/renderers/neus_volume_renderer.py:93-117, nerfacc.render_weight_from_alpha / accumulate_along_rays.
#pragma once
#include "tt_tc.cuh"

namespace tt {

constexpr int RAY_WARPS = 4;      // rays (warps) per CTA

__device__ __forceinline__ float wsum_all(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ unsigned lanes_below(int lane) { return (1u << lane) - 1u; }

// this ray's entries go to one contiguous slice of the list, in sample order; returns the slice start (all lanes)
__device__ __forceinline__ int warp_reserve(int n, int* counter, int lane) {
    int base = 0;
    if (lane == 0 && n > 0) base = atomicAdd(counter, n);
    return __shfl_sync(0xffffffffu, base, 0);
}

__global__ void __launch_bounds__(RAY_WARPS * 32) k_weights(tt_config cfg, RaySrcT rs, int64_t n_rays,
                                                          const float* __restrict__ sdf, const float* __restrict__ grad,
                                                          float* __restrict__ acc_o, float* weights_o, float* trans_o,
                                                          float* normal_o, float* feat_zero, int* live_idx,
                                                          int* live_count, int all_live) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * RAY_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;                                   // warp-uniform
    const float d[3] = {rs.rays_d[ray * 3], rs.rays_d[ray * 3 + 1], rs.rays_d[ray * 3 + 2]};
    const float o[3] = {rs.rays_o[ray * 3], rs.rays_o[ray * 3 + 1], rs.rays_o[ray * 3 + 2]};
    const float* t0p = rs.t_starts + ray * rs.t_stride;
    const float* t1p = rs.t_ends + ray * rs.t_stride;
    const int S = rs.S;
    float Tc = 1.f;                                              // transmittance entering the chunk
    float opac = 0.f, depth = 0.f, nsum[3] = {0.f, 0.f, 0.f}, eik = 0.f, wsum = 0.f, mean = 0.f, m2 = 0.f;
    int n_live = 0;
    for (int c0 = 0; c0 < S; c0 += 32) {
        const int i = c0 + lane;
        const bool in = i < S;
        const int64_t si = ray * S + (in ? i : 0);
        float alpha = 0.f, tm = 0.f, n[3] = {0.f, 0.f, 0.f};
        if (in) {
            const float t0 = t0p[i], t1 = t1p[i];
            tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
            const float g[3] = {grad[si * 3], grad[si * 3 + 1], grad[si * 3 + 2]};
            float len; normalize3(g, n, len);
            alpha = neus_alpha(sdf[si], n, d, __fsub_rn(t1, t0), cfg.inv_std, cfg.cos_anneal_ratio).alpha;
            eik += (len - 1.f) * (len - 1.f);
        }
        float inc = 1.f - alpha;                                 // inclusive product scan of (1 - α)
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const float v = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc *= v; }
        float ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) ex = 1.f;
        const float T = Tc * ex;
        Tc *= __shfl_sync(0xffffffffu, inc, 31);
        bool live = false;
        if (in) {
            const float w = T * alpha;
            opac += w; depth = fmaf(w, tm, depth);
#pragma unroll
            for (int a = 0; a < 3; ++a) nsum[a] = fmaf(w, n[a], nsum[a]);
            const float wn = wsum + w;
            if (wn > 0.f) { const float dl = tm - mean; mean += (w / wn) * dl; m2 += w * dl * (tm - mean); }
            wsum = wn;
            if (weights_o) weights_o[si] = w;
            if (trans_o) trans_o[si] = T;
            if (normal_o) { normal_o[si * 3] = n[0]; normal_o[si * 3 + 1] = n[1]; normal_o[si * 3 + 2] = n[2]; }
            const float x[3] = {__fadd_rn(o[0], __fmul_rn(d[0], tm)), __fadd_rn(o[1], __fmul_rn(d[1], tm)),
                                __fadd_rn(o[2], __fmul_rn(d[2], tm))};
            live = (all_live || T > 0.f) && !point_empty(x, cfg.radius, cfg.R);   // colour of an empty point: features = 0
            if (!live && feat_zero) { feat_zero[si * 3] = 0.f; feat_zero[si * 3 + 1] = 0.f; feat_zero[si * 3 + 2] = 0.f; }
        }
        n_live += __popc(__ballot_sync(0xffffffffu, live));
    }
    // ---- ray totals: sums, and a pairwise merge of the per-lane weighted mean / second moment -----------------------
    opac = wsum_all(opac); depth = wsum_all(depth); eik = wsum_all(eik);
#pragma unroll
    for (int a = 0; a < 3; ++a) nsum[a] = wsum_all(nsum[a]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float wb = __shfl_xor_sync(0xffffffffu, wsum, off), mb = __shfl_xor_sync(0xffffffffu, mean, off);
        const float qb = __shfl_xor_sync(0xffffffffu, m2, off);
        const float wn = wsum + wb;
        if (wn > 0.f) {
            const float dl = mb - mean;
            m2 = m2 + qb + dl * dl * (wsum * wb / wn);
            mean = (wsum * mean + wb * mb) / wn;
        }
        wsum = wn;
    }
    if (lane == 0) {
        float* a = acc_o + ray * TT_ACC;
        a[0] = opac; a[1] = depth; a[2] = 0.f; a[3] = 0.f; a[4] = 0.f;
        a[5] = m2 + wsum * (mean - depth) * (mean - depth);      // Σ w (t - depth)^2, depth un-normalised
        a[6] = nsum[0]; a[7] = nsum[1]; a[8] = nsum[2]; a[9] = eik;
    }
    if (live_idx) {     // second pass over the transmittance this warp has just written: ordered slice of the list
        int at = warp_reserve(n_live, live_count, lane);
        if (n_live > 0)
            for (int c0 = 0; c0 < S; c0 += 32) {
                const int i = c0 + lane;
                bool live = false;
                if (i < S) {
                    const float tm = __fmul_rn(__fadd_rn(t0p[i], t1p[i]), 0.5f);
                    const float x[3] = {__fadd_rn(o[0], __fmul_rn(d[0], tm)), __fadd_rn(o[1], __fmul_rn(d[1], tm)),
                                        __fadd_rn(o[2], __fmul_rn(d[2], tm))};
                    live = (all_live || trans_o[ray * S + i] > 0.f) && !point_empty(x, cfg.radius, cfg.R);
                }
                const unsigned m = __ballot_sync(0xffffffffu, live);
                if (live) live_idx[at + __popc(m & lanes_below(lane))] = (int)(ray * S + i);
                at += __popc(m);
            }
    }
}

// rgb accumulator: Σ_i (T_i alpha_i) sigmoid_mipnerf(f_i); weights are recomputed from the saved transmittance
__global__ void __launch_bounds__(RAY_WARPS * 32) k_accum_rgb(tt_config cfg, RaySrcT rs, int64_t n_rays,
                                                            const float* __restrict__ sdf, const float* __restrict__ grad,
                                                            const float* __restrict__ trans, const float* __restrict__ feat,
                                                            float* __restrict__ acc_o) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * RAY_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const float d[3] = {rs.rays_d[ray * 3], rs.rays_d[ray * 3 + 1], rs.rays_d[ray * 3 + 2]};
    const float* t0p = rs.t_starts + ray * rs.t_stride;
    const float* t1p = rs.t_ends + ray * rs.t_stride;
    const int S = rs.S;
    float rgb[3] = {0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < S; c0 += 32) {
        const int i = c0 + lane;
        const int64_t si = ray * S + i;
        const float T = i < S ? trans[si] : 0.f;
        if (__ballot_sync(0xffffffffu, T > 0.f) == 0u) break;    // transmittance is non-increasing along the ray
        if (T > 0.f) {
            const float dt = __fsub_rn(t1p[i], t0p[i]);
            const float g[3] = {grad[si * 3], grad[si * 3 + 1], grad[si * 3 + 2]};
            float n[3], len; normalize3(g, n, len);
            const float w = T * neus_alpha(sdf[si], n, d, dt, cfg.inv_std, cfg.cos_anneal_ratio).alpha;
#pragma unroll
            for (int a = 0; a < 3; ++a) rgb[a] = fmaf(w, sigmoid_mipnerf(feat[si * 3 + a]), rgb[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) rgb[a] = wsum_all(rgb[a]);
    if (lane == 0) { acc_o[ray * TT_ACC + 2] = rgb[0]; acc_o[ray * TT_ACC + 3] = rgb[1]; acc_o[ray * TT_ACC + 4] = rgb[2]; }
}

// backward of the compositing + alpha + normalisation.  Chunks of 32 samples are visited from the far end of the ray;
// inside a chunk lane 0 holds the farthest sample.  flags: [n_rays][2 * ceil(S/32)] words (nullable with the lists).
__host__ __device__ inline int ray_chunks(int S) { return (S + 31) / 32; }
__global__ void __launch_bounds__(RAY_WARPS * 32) k_render_bwd_comp(tt_config cfg, RaySrcT rs, int64_t n_rays,
        const float* __restrict__ acc, const float* __restrict__ sdf, const float* __restrict__ grad,
        const float* __restrict__ feat, const float* __restrict__ trans, const float* __restrict__ g_acc,
        const float* __restrict__ g_sdf, const float* __restrict__ g_grad, const float* __restrict__ g_normal,
        const float* __restrict__ g_feat, const float* __restrict__ g_weights, float rgb_scale,
        float* __restrict__ gs_o, float* __restrict__ u_o, float* __restrict__ gf_o, float* g_inv_std,
        int* geo_list, int* geo_count, int* tex_list, int* tex_count, uint32_t* flags, int emit_lists) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * RAY_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const float d[3] = {rs.rays_d[ray * 3], rs.rays_d[ray * 3 + 1], rs.rays_d[ray * 3 + 2]};
    const float o[3] = {rs.rays_o[ray * 3], rs.rays_o[ray * 3 + 1], rs.rays_o[ray * 3 + 2]};
    const float* t0p = rs.t_starts + ray * rs.t_stride;
    const float* t1p = rs.t_ends + ray * rs.t_stride;
    const int S = rs.S, NCH = ray_chunks(S);
    const float* ga = g_acc + ray * TT_ACC;
    const float opac = acc[ray * TT_ACC], D = acc[ray * TT_ACC + 1];
    const float gE = ga[9], gO = ga[0], gZ = ga[5];
    const float gD = ga[1] + gZ * (-2.f) * D * (1.f - opac);      // z_variance depends on depth[ray]
    const float gC[3] = {ga[2], ga[3], ga[4]}, gN[3] = {ga[6], ga[7], ga[8]};
    const float car = cfg.cos_anneal_ratio, inv_std = cfg.inv_std;
    uint32_t* fl = flags ? flags + ray * 2 * NCH : nullptr;
    float Rc = 0.f;       // R entering the chunk: Σ_{j beyond} gw_j α_j Π (1-α_k)
    float gis = 0.f;
    int ng = 0, nt = 0;
    for (int c = NCH - 1; c >= 0; --c) {
        const int i = c * 32 + 31 - lane;
        const bool in = i < S;
        const int64_t si = ray * S + (in ? i : 0);
        float s = 0.f, g[3] = {0.f, 0.f, 1.f}, f[3] = {0.f, 0.f, 0.f}, T = 0.f, tm = 0.f, dt = 0.f;
        if (in) {
            const float t0 = t0p[i], t1 = t1p[i];
            tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f); dt = __fsub_rn(t1, t0);
            s = sdf[si];
            g[0] = grad[si * 3]; g[1] = grad[si * 3 + 1]; g[2] = grad[si * 3 + 2];
            f[0] = feat[si * 3]; f[1] = feat[si * 3 + 1]; f[2] = feat[si * 3 + 2];
            T = trans[si];
        }
        float n[3], len; normalize3(g, n, len);
        const AlphaTerms at = neus_alpha(s, n, d, dt, inv_std, car);
        const float alpha = in ? at.alpha : 0.f;
        const float w = T * alpha;
        float c3[3], sg3[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { sg3[a] = sigmoidf(f[a]); c3[a] = sg3[a] * 1.002f - 0.001f; }
        float gw = gO + gD * tm + gC[0] * c3[0] + gC[1] * c3[1] + gC[2] * c3[2] + gN[0] * n[0] + gN[1] * n[1] +
                   gN[2] * n[2] + gZ * (tm - D) * (tm - D);
        if (in && g_weights) gw += g_weights[si];
        if (!in) gw = 0.f;
        // inclusive scan of the affine maps R -> A R + B (lower lanes = farther samples are applied first)
        float A = 1.f - alpha, B = gw * alpha;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float Ap = __shfl_up_sync(0xffffffffu, A, off), Bp = __shfl_up_sync(0xffffffffu, B, off);
            if (lane >= off) { B = fmaf(A, Bp, B); A *= Ap; }
        }
        float Ae = __shfl_up_sync(0xffffffffu, A, 1), Be = __shfl_up_sync(0xffffffffu, B, 1);
        if (lane == 0) { Ae = 1.f; Be = 0.f; }
        const float Rh = fmaf(Ae, Rc, Be);                        // R just beyond this sample
        Rc = fmaf(__shfl_sync(0xffffffffu, A, 31), Rc, __shfl_sync(0xffffffffu, B, 31));
        bool fg = false, ft = false;
        if (in) {
            const float galpha = T * (gw - Rh);
            float gfv[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {                          // colour
                float v = w * gC[a] * rgb_scale * 1.002f * sg3[a] * (1.f - sg3[a]);
                if (g_feat) v += g_feat[si * 3 + a];
                gfv[a] = v; gf_o[si * 3 + a] = v;
            }
            // alpha -> sdf, normal
            float gn[3] = {w * gN[0], w * gN[1], w * gN[2]};
            if (g_normal) { gn[0] += g_normal[si * 3]; gn[1] += g_normal[si * 3 + 1]; gn[2] += g_normal[si * 3 + 2]; }
            float gsdf = g_sdf ? g_sdf[si] : 0.f;
            if (at.alpha_raw >= 0.f && at.alpha_raw <= 1.f && galpha != 0.f) {
                const float den = at.prev_cdf + 1e-5f;
                const float gnum = galpha / den, gden = -galpha * at.alpha_raw / den;
                const float gpc = gnum + gden, gnc = -gnum;
                const float dp = at.prev_cdf * (1.f - at.prev_cdf), dn = at.next_cdf * (1.f - at.next_cdf);
                const float gsp = gpc * dp * inv_std, gsn = gnc * dn * inv_std;
                gis += gpc * dp * at.s_prev + gnc * dn * at.s_next;
                gsdf += gsp + gsn;
                const float giter = (gsn - gsp) * dt * 0.5f;
                const float dct = 0.5f * (1.f - car) * ((-at.true_cos * 0.5f + 0.5f) > 0.f ? 1.f : 0.f) +
                                  car * ((-at.true_cos) > 0.f ? 1.f : 0.f);
                const float gtc = giter * dct;
                gn[0] += gtc * d[0]; gn[1] += gtc * d[1]; gn[2] += gtc * d[2];
            }
            float u[3] = {0.f, 0.f, 0.f};                          // n = g / max(|g|, eps)
            if (len > 1e-12f) {
                const float dotv = n[0] * gn[0] + n[1] * gn[1] + n[2] * gn[2];
                const float il = 1.f / len;
#pragma unroll
                for (int a = 0; a < 3; ++a) u[a] = (gn[a] - n[a] * dotv) * il;
            } else {
#pragma unroll
                for (int a = 0; a < 3; ++a) u[a] = gn[a] * 1e12f;
            }
            if (g_grad) { u[0] += g_grad[si * 3]; u[1] += g_grad[si * 3 + 1]; u[2] += g_grad[si * 3 + 2]; }
            if (gE != 0.f && len > 0.f) {                          // d (|g|-1)^2 / d g = 2 (|g|-1) g / |g|
                const float ce = gE * 2.f * (len - 1.f) / len;
                u[0] += ce * g[0]; u[1] += ce * g[1]; u[2] += ce * g[2];
            }
            gs_o[si] = gsdf;
            u_o[si * 3] = u[0]; u_o[si * 3 + 1] = u[1]; u_o[si * 3 + 2] = u[2];
            if (fl) {   // samples the decoder backward has to visit: non-empty point (an empty point depends on no
                        // parameter) and a non-zero seed
                const bool sg = gsdf != 0.f || u[0] != 0.f || u[1] != 0.f || u[2] != 0.f;
                const bool st = gfv[0] != 0.f || gfv[1] != 0.f || gfv[2] != 0.f;
                if (sg || st) {
                    const float x[3] = {__fadd_rn(o[0], __fmul_rn(d[0], tm)), __fadd_rn(o[1], __fmul_rn(d[1], tm)),
                                        __fadd_rn(o[2], __fmul_rn(d[2], tm))};
                    const bool ne = !point_empty(x, cfg.radius, cfg.R);
                    fg = sg && ne; ft = st && ne;
                }
            }
        }
        if (fl) {       // lane L holds sample 32c + 31 - L: bit-reverse so that bit b = sample 32c + b
            const unsigned mg = __brev(__ballot_sync(0xffffffffu, fg)), mt = __brev(__ballot_sync(0xffffffffu, ft));
            if (lane == 0) { fl[2 * c] = mg; fl[2 * c + 1] = mt; }
            ng += __popc(mg); nt += __popc(mt);
        }
    }
    if (g_inv_std) {
        gis = wsum_all(gis);
        if (lane == 0 && gis != 0.f) atomicAdd(g_inv_std, gis);
    }
    if (fl && emit_lists) {           // ordered slices of the two lists (ray order; patch order: k_patch_lists)
        __syncwarp();
        int at_g = warp_reserve(ng, geo_count, lane), at_t = warp_reserve(nt, tex_count, lane);
        for (int c = 0; c < NCH; ++c) {
            const unsigned mg = fl[2 * c], mt = fl[2 * c + 1];
            const int si = (int)(ray * S + c * 32 + lane);
            if ((mg >> lane) & 1u) geo_list[at_g + __popc(mg & lanes_below(lane))] = si;
            if ((mt >> lane) & 1u) tex_list[at_t + __popc(mt & lanes_below(lane))] = si;
            at_g += __popc(mg); at_t += __popc(mt);
        }
    }
}

// ---- patch-ordered sample lists --------------------------------------------------------------------------------------------
// When the rays are [B][H][W] images (tt_config.image_h/w), the lists of the samples the decoder backward visits are
// emitted from the flag words of k_render_bwd_comp in PATCH order: CTA = a 4x4 patch of neighbouring pixels, entries ordered
// (chunk of 8 consecutive samples, ray of the patch, sample of the chunk).  A 128-sample tile of the backward kernels is then
// ~ 16 neighbouring rays x 8 samples, whose bilinear taps fall on few distinct texels (~8 taps per texel at config 3,
// 1.5 in ray order): coop_scatter_merged (tt_tc_bwd.cuh) sums them on chip.  The patches reserve their slices with one
// atomic per list, like the rays of the ray-ordered lists; the order of the entries never changes a result bit of the
// forward and only the (already unordered) summation order of the float reductions in the backward.
constexpr int PATCH = 4, PATCH_RAYS = PATCH * PATCH, PATCH_THREADS = 256, PATCH_MAX_CHUNKS = 32;
__global__ void __launch_bounds__(PATCH_THREADS) k_patch_lists(const uint32_t* __restrict__ flags, int H, int W, int S,
        int* geo_list, int* geo_count, int* tex_list, int* tex_count) {
    TT_SHARED(smem);
    uint32_t* fl_s = reinterpret_cast<uint32_t*>(smem);                 // [16 rays][2 NCH]
    int* misc = reinterpret_cast<int*>(smem) + PATCH_RAYS * 2 * PATCH_MAX_CHUNKS;      // warp sums [8], base
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int NCH = ray_chunks(S), PW = W / PATCH, PH = H / PATCH;
    const int img = blockIdx.x / (PH * PW), rem = blockIdx.x - img * PH * PW, py = rem / PW, px = rem - py * PW;
    auto ray_of = [&](int r) { return ((int64_t)img * H + py * PATCH + r / PATCH) * W + px * PATCH + r % PATCH; };
    for (int i = tid; i < PATCH_RAYS * 2 * NCH; i += PATCH_THREADS) {
        const int r = i / (2 * NCH), w = i - r * 2 * NCH;
        fl_s[i] = flags[ray_of(r) * 2 * NCH + w];
    }
    __syncthreads();
    const int n8 = (S + 7) / 8, cells = n8 * PATCH_RAYS, K = (cells + PATCH_THREADS - 1) / PATCH_THREADS;
    for (int which = 0; which < 2; ++which) {
        int* list = which ? tex_list : geo_list; int* count = which ? tex_count : geo_count;
        auto mask_of = [&](int q) -> uint32_t {          // cell q = (chunk of 8 samples, ray): its 8 flag bits
            const int c8 = q / PATCH_RAYS, r = q - c8 * PATCH_RAYS;
            return (fl_s[r * 2 * NCH + 2 * (c8 >> 2) + which] >> ((c8 & 3) * 8)) & 0xffu;
        };
        int sum = 0;
        for (int i = 0; i < K; ++i) { const int q = tid * K + i; if (q < cells) sum += __popc(mask_of(q)); }
        int v = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, v, off); if (lane >= off) v += o; }
        if (lane == 31) misc[wrp] = v;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int q = 0; q < PATCH_THREADS / 32; ++q) { const int t = misc[q]; misc[q] = tot; tot += t; }
            misc[8] = tot > 0 ? atomicAdd(count, tot) : 0;
        }
        __syncthreads();
        int at = misc[8] + misc[wrp] + v - sum;
        for (int i = 0; i < K; ++i) {
            const int q = tid * K + i;
            if (q >= cells) break;
            const int c8 = q / PATCH_RAYS, r = q - c8 * PATCH_RAYS;
            uint32_t m = mask_of(q);
            const int64_t si0 = ray_of(r) * S + c8 * 8;
            while (m) { const int b = __ffs(m) - 1; m &= m - 1; list[at++] = (int)(si0 + b); }
        }
        __syncthreads();
    }
}

// ---- stand-alone compositor, one WARP per ray (nerfacc.render_weight_from_alpha + accumulate_along_rays on dense rays) -----
// The 32 lanes take 32 consecutive samples, so alphas / values / weights / trans stream with coalesced accesses (a
// thread-per-ray loop strides S floats between lanes: measured 7 % of the HBM peak, this form is HBM-bound).
//   forward   T_i = Π_{j<i} (1 - α_j) as an exclusive product scan, w = T α, out[k] = Σ w v[k]
//   backward  R_{i-1} = gw_i α_i + (1 - α_i) R_i as a scan of affine maps from the far end, g_α_i = T_i (gw_i - R_i)
template <int DMAX>
__global__ void __launch_bounds__(RAY_WARPS * 32) k_composite_fwd_w(const float* __restrict__ alphas, const float* __restrict__ values,
                                                                   int64_t n_rays, int S, int D, float* weights, float* trans, float* out) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * RAY_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const int Do = D > 0 ? D : 1;
    float acc[DMAX];
#pragma unroll
    for (int k = 0; k < DMAX; ++k) acc[k] = 0.f;
    float Tc = 1.f;
    for (int c0 = 0; c0 < S; c0 += 32) {
        const int i = c0 + lane;
        const bool in = i < S;
        const int64_t si = ray * S + (in ? i : 0);
        const float a = in ? alphas[si] : 0.f;
        float inc = 1.f - a;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const float v = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc *= v; }
        float ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) ex = 1.f;
        const float T = Tc * ex;
        Tc *= __shfl_sync(0xffffffffu, inc, 31);
        if (in) {
            const float w = T * a;
            if (weights) weights[si] = w;
            if (trans) trans[si] = T;
            if (D > 0) {
#pragma unroll
                for (int k = 0; k < DMAX; ++k) if (k < D) acc[k] = fmaf(w, values[si * D + k], acc[k]);
            } else acc[0] += w;
        }
    }
#pragma unroll
    for (int k = 0; k < DMAX; ++k) if (k < Do) acc[k] = wsum_all(acc[k]);
    if (out && lane == 0)
#pragma unroll
        for (int k = 0; k < DMAX; ++k) if (k < Do) out[ray * Do + k] = acc[k];
}
template <int DMAX>
__global__ void __launch_bounds__(RAY_WARPS * 32) k_composite_bwd_w(const float* __restrict__ alphas, const float* __restrict__ values,
                                                                   const float* __restrict__ trans, const float* __restrict__ g_out,
                                                                   const float* __restrict__ g_weights, int64_t n_rays, int S, int D,
                                                                   float* g_alphas, float* g_values) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * RAY_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const int Do = D > 0 ? D : 1;
    float go[DMAX];
#pragma unroll
    for (int k = 0; k < DMAX; ++k) go[k] = (g_out && k < Do) ? g_out[ray * Do + k] : 0.f;
    const int NCH = (S + 31) / 32;
    float Rc = 0.f;
    for (int c = NCH - 1; c >= 0; --c) {
        const int i = c * 32 + 31 - lane;          // lane 0 holds the farthest sample of the chunk
        const bool in = i < S;
        const int64_t si = ray * S + (in ? i : 0);
        const float a = in ? alphas[si] : 0.f, T = in ? trans[si] : 0.f;
        float gw = (in && g_weights) ? g_weights[si] : 0.f;
        if (in) {
            if (D > 0) {
#pragma unroll
                for (int k = 0; k < DMAX; ++k)
                    if (k < D) {
                        gw = fmaf(go[k], values[si * D + k], gw);
                        if (g_values) g_values[si * D + k] = T * a * go[k];
                    }
            } else gw += go[0];
        }
        float A = 1.f - a, B = gw * a;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float Ap = __shfl_up_sync(0xffffffffu, A, off), Bp = __shfl_up_sync(0xffffffffu, B, off);
            if (lane >= off) { B = fmaf(A, Bp, B); A *= Ap; }
        }
        float Ae = __shfl_up_sync(0xffffffffu, A, 1), Be = __shfl_up_sync(0xffffffffu, B, 1);
        if (lane == 0) { Ae = 1.f; Be = 0.f; }
        const float Rh = fmaf(Ae, Rc, Be);          // R just beyond this sample
        Rc = fmaf(__shfl_sync(0xffffffffu, A, 31), Rc, __shfl_sync(0xffffffffu, B, 31));
        if (in) g_alphas[si] = T * (gw - Rh);
    }
}

}  // namespace tt
