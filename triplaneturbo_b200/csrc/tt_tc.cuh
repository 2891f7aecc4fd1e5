// Tensor-core (tcgen05) kernels of the forward path: point-list SDF decoder (+ analytic normal) and colour decoder,
// plus the light per-ray kernels around them.  See tt_umma.cuh for the execution model.
//
// Forward pipeline (tt_render_fwd, impl = 1):
//   k_classify         closed-form outputs of the empty samples, list of the others
//   k_geo_tc<C,true>   sdf, d sdf/dx at every non-empty sample, ReLU masks  (tensor cores, cooperative gathers)
//   k_weights          per ray: NeuS alpha, transmittance, weights, all non-colour accumulators, live-sample list (tt_rays.cuh)
//   k_tex_tc<C>        colour features at the live samples (T > 0), ReLU masks (tensor cores)
//   k_accum_rgb        per ray: Σ w · sigmoid_mipnerf(features)                  (tt_rays.cuh)
// Sampler (tt_importance_sample, impl = 1):  k_classify + k_geo_tc<C,false> on the proposal midpoints, then k_sampler_post.
// Field query (tt_geometry_fwd): k_geo_tc<C,false,DEFORM> also runs the deformation decoder on the same encoding.
#pragma once
#include "tt_device.cuh"
#include "tt_umma.cuh"

namespace tt {

constexpr int TC_THREADS = 256;      // two groups of 128 per CTA share the weight tiles
constexpr int TC_GROUPS = TC_THREADS / TC_GROUP;

struct RaySrcT {
    const float* rays_o; const float* rays_d;
    const float* t_starts; const float* t_ends; int64_t t_stride; int S;
};
struct TcSrc {
    int mode;                 // 0: explicit points, 1: ray samples (id = ray*S + i), 2: proposal midpoints
                              // (id = ray*n_imp + j), 3: isosurface grid (id = prompt*res^3 + vertex)
    const float* points; int64_t M;
    RaySrcT rs; int rays_per_cache;
    int n_imp; const float* jitter0; float near_plane, far_plane;
    int grid_res; int grid_lines;           // grid_lines: z-line gather for the regular grid (k_geo_ws, opt-in experiment)
    const int* index; const int* count;     // optional compaction: slot -> id, *count live slots
};

__device__ __forceinline__ float quantile_s(int j, int n, bool strat, float b) {
    return strat ? __fdiv_rn(__fadd_rn((float)j, b), (float)(n + 1)) : __fdiv_rn((float)j, (float)n);
}
__device__ __forceinline__ float stot_u(float s, float tmin, float tmax) {
    return __fadd_rn(__fmul_rn(s, tmax), __fmul_rn(__fsub_rn(1.f, s), tmin));
}
__device__ __forceinline__ float grid_coord_tc(int i, int res) {
    const float step = __fdiv_rn(1.f, (float)(res - 1));
    float v = (i < res / 2) ? __fmul_rn((float)i, step) : __fsub_rn(1.f, __fmul_rn((float)(res - 1 - i), step));
    return __fadd_rn(__fmul_rn(v, 2.f), -1.f);
}
// (every tensor-core kernel is launched with N < 2^31 sample ids: the index arithmetic is 32-bit — a 64-bit division costs
// ~100 instructions, and these kernels are bound by instructions per thread)
__device__ __forceinline__ void tc_point(const TcSrc& s, int64_t id, float (&x)[3], int& prompt) {
    const uint32_t id32 = (uint32_t)id;
    if (s.mode == 0) {
        x[0] = s.points[id * 3]; x[1] = s.points[id * 3 + 1]; x[2] = s.points[id * 3 + 2];
        prompt = s.M > 0x7fffffffLL ? 0 : (int)(id32 / (uint32_t)s.M);
    } else if (s.mode == 3) {
        const uint32_t M = (uint32_t)s.M, rr = (uint32_t)s.grid_res;
        prompt = (int)(id32 / M);
        const uint32_t v = id32 - (uint32_t)prompt * M, xi = v / (rr * rr), rem = v - xi * rr * rr, yi = rem / rr;
        x[0] = grid_coord_tc((int)xi, (int)rr); x[1] = grid_coord_tc((int)yi, (int)rr);
        x[2] = grid_coord_tc((int)(rem - yi * rr), (int)rr);
    } else {
        uint32_t ray; float tm;
        if (s.mode == 1) {
            ray = id32 / (uint32_t)s.rs.S; const int i = (int)(id32 - ray * (uint32_t)s.rs.S);
            const float t0 = s.rs.t_starts[(int64_t)ray * s.rs.t_stride + i], t1 = s.rs.t_ends[(int64_t)ray * s.rs.t_stride + i];
            tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
        } else {
            ray = id32 / (uint32_t)s.n_imp; const int j = (int)(id32 - ray * (uint32_t)s.n_imp);
            const bool strat = s.jitter0 != nullptr; const float b = strat ? s.jitter0[ray] : 0.f;
            const float t0 = stot_u(quantile_s(j, s.n_imp, strat, b), s.near_plane, s.far_plane);
            const float t1 = stot_u(quantile_s(j + 1, s.n_imp, strat, b), s.near_plane, s.far_plane);
            tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) x[a] = __fadd_rn(s.rs.rays_o[(size_t)ray * 3 + a], __fmul_rn(s.rs.rays_d[(size_t)ray * 3 + a], tm));
        prompt = (int)(ray / (uint32_t)s.rays_per_cache);
    }
}

// Sample of a tile, prefetched one tile ahead as RAW loaded words: the arithmetic that turns them into a position runs one
// tile later, so the global-load latency (sample id -> interval edges / ray) hides behind the previous tile's work
// (gather warps of k_geo_ws, consumer-made tables, the colour backward).
struct WsRaw { float v[8]; int id; };
__device__ __forceinline__ int ws_load_id(const TcSrc& src, int64_t tile, int mt, int64_t n_live, int64_t n_tiles) {
    if (tile >= n_tiles) return -1;
    const int64_t slot = tile * TC_GROUP + mt;
    if (slot >= n_live) return -1;
    return src.index ? src.index[slot] : (int)slot;
}
__device__ __forceinline__ WsRaw ws_load_raw(const TcSrc& s, int id) {
    WsRaw r; r.id = id;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = 0.f;
    if (id < 0) return r;
    if (s.mode == 0) {
        r.v[0] = s.points[(int64_t)id * 3]; r.v[1] = s.points[(int64_t)id * 3 + 1]; r.v[2] = s.points[(int64_t)id * 3 + 2];
    } else if (s.mode == 1 || s.mode == 2) {
        int64_t ray;
        if (s.mode == 1) {
            ray = id / s.rs.S; const int i = id - (int)ray * s.rs.S;
            r.v[0] = s.rs.t_starts[ray * s.rs.t_stride + i]; r.v[1] = s.rs.t_ends[ray * s.rs.t_stride + i];
        } else {
            ray = id / s.n_imp;
            r.v[0] = s.jitter0 ? s.jitter0[ray] : 0.f;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) { r.v[2 + a] = s.rs.rays_o[ray * 3 + a]; r.v[5 + a] = s.rs.rays_d[ray * 3 + a]; }
    }
    return r;
}
// same arithmetic as tc_point (tt_tc.cuh), on the prefetched words
__device__ __forceinline__ void ws_point_from_raw(const TcSrc& s, const WsRaw& r, float (&x)[3], int& prompt) {
    x[0] = x[1] = x[2] = 0.f; prompt = 0;
    if (r.id < 0) return;
    if (s.mode == 0) {
        x[0] = r.v[0]; x[1] = r.v[1]; x[2] = r.v[2];
        prompt = s.M > 0x7fffffffLL ? 0 : (int)((uint32_t)r.id / (uint32_t)s.M);
    } else if (s.mode == 3) {
        // (sample ids are 32-bit here and M = rr^3 <= N < 2^31: 32-bit divisions, a 64-bit one costs ~100 instructions)
        const uint32_t M = (uint32_t)s.M, rr = (uint32_t)s.grid_res;
        prompt = (int)((uint32_t)r.id / M);
        const uint32_t v = (uint32_t)r.id - (uint32_t)prompt * M, xi = v / (rr * rr), rem = v - xi * rr * rr, yi = rem / rr;
        x[0] = grid_coord_tc((int)xi, (int)rr); x[1] = grid_coord_tc((int)yi, (int)rr);
        x[2] = grid_coord_tc((int)(rem - yi * rr), (int)rr);
    } else {
        int ray; float tm;
        if (s.mode == 1) {
            ray = r.id / s.rs.S;
            tm = __fmul_rn(__fadd_rn(r.v[0], r.v[1]), 0.5f);
        } else {
            ray = r.id / s.n_imp; const int j = r.id - ray * s.n_imp;
            const bool strat = s.jitter0 != nullptr; const float b = r.v[0];
            const float t0 = stot_u(quantile_s(j, s.n_imp, strat, b), s.near_plane, s.far_plane);
            const float t1 = stot_u(quantile_s(j + 1, s.n_imp, strat, b), s.near_plane, s.far_plane);
            tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) x[a] = __fadd_rn(r.v[2 + a], __fmul_rn(r.v[5 + a], tm));
        prompt = ray / s.rays_per_cache;
    }
}

// A point is EMPTY when every bilinear tap of every plane is out of bounds (zeros padding): both encodings are
// exactly 0, so sdf = |x| - bias, d sdf/dx = x/|x|, features = 0 and every gradient vanishes.  Such samples
// (typically 35-50 % of a ray: the segment outside the [-r,r]^3 box) never enter the tensor-core kernels.
__device__ __forceinline__ bool point_empty(const float (&x)[3], float radius, int R) {
    float p[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], radius);
    const float fR = (float)R;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(p[plane_ax(k)], 1.f), fR), 1.f), 0.5f);
        const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(p[plane_ay(k)], 1.f), fR), 1.f), 0.5f);
        const float x0 = floorf(ix), y0 = floorf(iy);
        if (x0 >= -1.f && x0 <= fR - 1.f && y0 >= -1.f && y0 <= fR - 1.f) return false;
    }
    return true;
}
// Ordered block-level reservation in a list: every thread of the block asks for n slots; slices are handed out in
// thread order inside ONE contiguous range per block (a single atomicAdd), so that the list keeps the sample order
// of the block (ray-major).  (A warp-aggregated append interleaves 32-entry chunks of every resident warp of the GPU,
// which scatters a 128-point tile over several views: measured 123 GB of DRAM reads per backward launch.)  All threads
// must call.
template <int NTHREADS>
__device__ __forceinline__ int block_reserve(int n, int* counter) {
    __shared__ int s_cnt[NTHREADS / 32 + 1];
#ifndef TT_EMUL
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = n;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += v; }
    if (lane == 31) s_cnt[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < NTHREADS / 32; ++w) { const int c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
        s_cnt[NTHREADS / 32] = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    const int r = s_cnt[NTHREADS / 32] + s_cnt[warp] + inc - n;
    __syncthreads();
    return r;
#else
    __shared__ int s_all[NTHREADS];
    s_all[threadIdx.x] = n;
    __syncthreads();
    int pre = 0;
    for (int t = 0; t < (int)threadIdx.x; ++t) pre += s_all[t];
    if (threadIdx.x == NTHREADS - 1) s_cnt[0] = atomicAdd(counter, pre + n);
    __syncthreads();
    const int r = s_cnt[0] + pre;
    __syncthreads();
    return r;
#endif
}

// one thread per point: closed-form outputs for empty points, list of the others (block-contiguous, in point order)
__global__ void __launch_bounds__(256) k_classify(tt_config cfg, TcSrc src, int64_t N, float* sdf_o, float* sdf_orig_o,
                                                 float* grad_o, float* normal_o, int* list, int* count) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    if (id < N) {
        float x[3]; int prompt;
        tc_point(src, id, x, prompt);
        if (point_empty(x, cfg.radius, cfg.R)) {
            const float nrm = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
            const float s = 0.f;
            if (sdf_orig_o) sdf_orig_o[id] = s;
            if (sdf_o) sdf_o[id] = s + (nrm - cfg.sdf_bias_radius);
            if (grad_o || normal_o) {
                const float inv = nrm > 0.f ? 1.f / nrm : 0.f;
                const float g[3] = {x[0] * inv, x[1] * inv, x[2] * inv};
                if (grad_o) { grad_o[id * 3] = g[0]; grad_o[id * 3 + 1] = g[1]; grad_o[id * 3 + 2] = g[2]; }
                if (normal_o) {
                    float n[3], len; normalize3(g, n, len);
                    normal_o[id * 3] = n[0]; normal_o[id * 3 + 1] = n[1]; normal_o[id * 3 + 2] = n[2];
                }
            }
        } else take = true;
    }
    const int at = block_reserve<256>(take ? 1 : 0, count);
    if (take) list[at] = (int)id;
}

// ---- cooperative gather: 128 points x C channels, consecutive lanes read consecutive 16-byte chunks of a texel ---
// Tap table per point: int o[NT] (texel index, CLAMPED to a valid texel), float w[NT] (weight, 0 for out-of-bounds
// taps: zeros padding), NT = 4 * NPL; pbase = prompt index.  All NT loads of an item are issued before they are
// used, without branches, so the memory system sees NT independent 16-byte requests per lane.
#ifndef TT_GATHER_LOADS
#define TT_GATHER_LOADS 24      // independent 16-byte loads in flight per lane: one group alone must fill the load path
#endif
template <int C, int NPL>
__device__ __forceinline__ void coop_gather(const float* __restrict__ planes, size_t ps, const int* tap_o,
                                            const float* tap_w, const uint32_t* pbase, int plane0, float* stage, int tg) {
    // JB items per batch so that 12 independent 16-byte loads are in flight per lane whatever NPL is
    constexpr int U = C / 4, NT = 4 * NPL, SP = C + 4, JB = TT_GATHER_LOADS / NT;
#pragma unroll 1
    for (int j0 = 0; j0 < U; j0 += JB) {
        float4 v[JB][NT];
        float w[JB][NT];
        int sto[JB];
#pragma unroll
        for (int b = 0; b < JB; ++b) {
            const int j = j0 + b < U ? j0 + b : U - 1;          // tail batch: repeat the last item (same result)
            const int item = tg + TC_GROUP * j;
            const int pt = item / U, ch = item - pt * U;
            sto[b] = pt * SP + ch * 4;
            const float* base = planes + (size_t)pbase[pt] * 6 * ps + (size_t)plane0 * ps + ch * 4;
#pragma unroll
            for (int q = 0; q < NT; q += 4) {
                const int4 o4 = *reinterpret_cast<const int4*>(tap_o + pt * NT + q);
                const float4 w4 = *reinterpret_cast<const float4*>(tap_w + pt * NT + q);
                const float* pb = base + (size_t)(q >> 2) * ps;
                v[b][q] = ldg4(pb + (size_t)o4.x * C); v[b][q + 1] = ldg4(pb + (size_t)o4.y * C);
                v[b][q + 2] = ldg4(pb + (size_t)o4.z * C); v[b][q + 3] = ldg4(pb + (size_t)o4.w * C);
                w[b][q] = w4.x; w[b][q + 1] = w4.y; w[b][q + 2] = w4.z; w[b][q + 3] = w4.w;
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
// table entries of one tap: clamped offset, weight (0 when out of bounds)
__device__ __forceinline__ int tap_off(int o) { return o < 0 ? 0 : o; }

// shared-memory plan of k_geo_tc (float offsets)
template <int C, bool NORMAL, bool DEFORM = false>
struct GeoSmem {
    static constexpr int CP = (C + 15) / 16 * 16;
    static constexpr int W1H = 0, W1L = W1H + 64 * C, W2H = W1L + 64 * C, W2L = W2H + 4096;
    static constexpr int W2TH = W2L + 4096, W2TL = W2TH + (NORMAL ? 4096 : 0);
    static constexpr int W1TH = W2TL + (NORMAL ? 4096 : 0), W1TL = W1TH + (NORMAL ? CP * 64 : 0);
    static constexpr int W3 = W1TL + (NORMAL ? CP * 64 : 0);
    // deformation MLP (same encoding, second decoder): W1d, W2d hi/lo tiles and the [3][64] head
    static constexpr int D1H = W3 + 64, D1L = D1H + (DEFORM ? 64 * C : 0), D2H = D1L + (DEFORM ? 64 * C : 0);
    static constexpr int D2L = D2H + (DEFORM ? 4096 : 0), D3 = D2L + (DEFORM ? 4096 : 0);
    static constexpr int GROUP0 = D3 + (DEFORM ? 192 : 0);
    // per group
    static constexpr int TAP_O = 0, TAP_W = TAP_O + 128 * 12, TAP_F = TAP_W + 128 * 12;      // ints, floats, factors
    static constexpr int TAP_V = TAP_F + (NORMAL ? 128 * 12 : 0);                              // validity (1 / 0)
    static constexpr int PBASE = TAP_V + (NORMAL ? 128 * 12 : 0);
    static constexpr int NACC = PBASE + 128;
    static constexpr int STAGE = NACC + (NORMAL ? 128 * 6 : 0);
    static constexpr int GROUP_FLOATS = STAGE + 128 * (C + 4);
    static constexpr int TOTAL = GROUP0 + TC_GROUPS * GROUP_FLOATS + 16;     // + mbarriers, tmem slot
};

template <int C, bool NORMAL, bool DEFORM = false>
__global__ void __launch_bounds__(TC_THREADS, 1) k_geo_tc(const float* __restrict__ planes, const float* __restrict__ wp,
                                                         tt_config cfg, TcSrc src, int64_t N, float* sdf_o,
                                                         float* sdf_orig_o, float* grad_o, float* normal_o,
                                                         uint64_t* masks_o, float* deform_o = nullptr) {
    TT_SHARED(smem);
    using L = GeoSmem<C, NORMAL, DEFORM>;
    constexpr int CP = L::CP, SP = C + 4;
    const int tid = threadIdx.x, group = tid / TC_GROUP, tg = tid % TC_GROUP, warp = tid >> 5;
    const WOff wo = woff(C);
    // ---- one-time setup: weight tiles (tf32 hi/lo, canonical K-major), mbarriers, TMEM ------------------------
    btile_fill(smem + L::W1H, smem + L::W1L, 64, C, [&](int n, int k) { return __ldg(wp + wo.w1s + n * C + k); }, tid, TC_THREADS);
    btile_fill(smem + L::W2H, smem + L::W2L, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2s + n * 64 + k); }, tid, TC_THREADS);
    if (NORMAL) {
        btile_fill(smem + L::W2TH, smem + L::W2TL, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2s + k * 64 + n); }, tid, TC_THREADS);
        btile_fill(smem + L::W1TH, smem + L::W1TL, CP, 64, [&](int n, int k) { return n < C ? __ldg(wp + wo.w1s + k * C + n) : 0.f; }, tid, TC_THREADS);
    }
    if (tid < 64) smem[L::W3 + tid] = __ldg(wp + wo.w3s + tid);
    if (DEFORM) {
        btile_fill(smem + L::D1H, smem + L::D1L, 64, C, [&](int n, int k) { return __ldg(wp + wo.w1d + n * C + k); }, tid, TC_THREADS);
        btile_fill(smem + L::D2H, smem + L::D2L, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2d + n * 64 + k); }, tid, TC_THREADS);
        if (tid < 192) smem[L::D3 + tid] = __ldg(wp + wo.w3d + tid);
    }
    uint64_t* mbars = reinterpret_cast<uint64_t*>(smem + L::GROUP0 + TC_GROUPS * L::GROUP_FLOATS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbars + TC_GROUPS);
    int* mlock = reinterpret_cast<int*>(tmem_slot + 1);
    if (tid == 0) { for (int g = 0; g < TC_GROUPS; ++g) mbar_init(mbars + g); *mlock = 0; }
    if (warp == 0) tmem_alloc_warp(tmem_slot, 512);
    async_proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    Umma u;
    u.tmem = *tmem_slot + (uint32_t)group * TC_COLS_PER_GROUP;
    u.lane_base = (uint32_t)((warp & 3) * 32) << 16;
    u.mbar = smem_u32(mbars + group); u.phase = 0; u.group = group;
    const bool leader = tg == 0;
    const BTile bD1 = btile_make(smem + L::D1H, smem + L::D1L, 64, C);
    const BTile bD2 = btile_make(smem + L::D2H, smem + L::D2L, 64, 64);
    const BTile bW1 = btile_make(smem + L::W1H, smem + L::W1L, 64, C);
    const BTile bW2 = btile_make(smem + L::W2H, smem + L::W2L, 64, 64);
    const BTile bW2T = btile_make(smem + L::W2TH, smem + L::W2TL, 64, 64);
    const BTile bW1T = btile_make(smem + L::W1TH, smem + L::W1TL, CP, 64);
    float* gs = smem + L::GROUP0 + group * L::GROUP_FLOATS;
    int* tap_o = reinterpret_cast<int*>(gs + L::TAP_O);
    float* tap_w = gs + L::TAP_W;
    float* tap_f = gs + L::TAP_F;
    float* tap_v = gs + L::TAP_V;
    uint32_t* pbase = reinterpret_cast<uint32_t*>(gs + L::PBASE);
    float* nacc = gs + L::NACC;
    float* stage = gs + L::STAGE;
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const int64_t n_live = src.count ? (int64_t)*src.count : N;
    const int64_t n_tiles = (n_live + TC_GROUP - 1) / TC_GROUP;
    const float* w3 = smem + L::W3;

    for (int64_t tile = (int64_t)blockIdx.x * TC_GROUPS + group; tile < n_tiles; tile += (int64_t)gridDim.x * TC_GROUPS) {
        const int64_t slot = tile * TC_GROUP + tg;
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
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const bool in = valid && tp[k].o[t] >= 0;
                tap_o[tg * 12 + k * 4 + t] = in ? tp[k].o[t] : 0;
                tap_w[tg * 12 + k * 4 + t] = in ? tp[k].w[t] : 0.f;
                if (NORMAL) tap_v[tg * 12 + k * 4 + t] = in ? 1.f : 0.f;
            }
            if (NORMAL) {
                tap_f[tg * 12 + k * 4 + 0] = valid ? tp[k].wx0 : 0.f; tap_f[tg * 12 + k * 4 + 1] = valid ? tp[k].wx1 : 0.f;
                tap_f[tg * 12 + k * 4 + 2] = valid ? tp[k].wy0 : 0.f; tap_f[tg * 12 + k * 4 + 3] = valid ? tp[k].wy1 : 0.f;
            }
        }
        pbase[tg] = (uint32_t)prompt;
        mem_lock(mlock, leader, group);
        coop_gather<C, 3>(planes, ps, tap_o, tap_w, pbase, 0, stage, tg);
        mem_unlock(mlock, leader, group);
        // ---- SDF MLP on tensor cores ----------------------------------------------------------------------------
        float d[64];
        uint64_t m1 = 0, m2 = 0;
        {
            float e[C];
#pragma unroll
            for (int c = 0; c < C; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(stage + tg * SP + c);
                e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
            }
            umma_layer<C, 64, 3>(u, leader, e, bW1, d);
        }
#pragma unroll
        for (int j = 0; j < 64; ++j) { const bool on = d[j] > 0.f; m1 |= (uint64_t)on << j; d[j] = on ? d[j] : 0.f; }
        {
            float h[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) h[j] = d[j];
            umma_layer<64, 64, 3>(u, leader, h, bW2, d);
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) { const bool on = d[j] > 0.f; m2 |= (uint64_t)on << j; s = fmaf(on ? d[j] : 0.f, w3[j], s); }
        const float nrm = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        if (valid) {
            if (sdf_orig_o) sdf_orig_o[id] = s;
            if (sdf_o) sdf_o[id] = s + (nrm - cfg.sdf_bias_radius);
            if (masks_o) { masks_o[id * 4 + 2] = m1; masks_o[id * 4 + 3] = m2; }     // ReLU masks for the backward
        }
        if (DEFORM) {     // deformation decoder on the same encoding (still in the stage): 32->64->64->3, 3xTF32
            const float* w3d = smem + L::D3;
            {
                float e[C];
#pragma unroll
                for (int c = 0; c < C; c += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(stage + tg * SP + c);
                    e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
                }
                umma_layer<C, 64, 3>(u, leader, e, bD1, d);
            }
            {
                float h[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) h[j] = fmaxf(d[j], 0.f);
                umma_layer<64, 64, 3>(u, leader, h, bD2, d);
            }
            float df[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const float h = fmaxf(d[j], 0.f);
                df[0] = fmaf(h, w3d[j], df[0]); df[1] = fmaf(h, w3d[64 + j], df[1]); df[2] = fmaf(h, w3d[128 + j], df[2]);
            }
            if (valid && deform_o) { deform_o[id * 3] = df[0]; deform_o[id * 3 + 1] = df[1]; deform_o[id * 3 + 2] = df[2]; }
        }
        if (NORMAL) {
            {   // unit-seed adjoint: a2 = m2 ⊙ w3 ; a1 = m1 ⊙ (W2ᵀ a2) ; de = W1ᵀ a1
                float a[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) a[j] = ((m2 >> j) & 1ull) ? w3[j] : 0.f;
                umma_layer<64, 64, 3>(u, leader, a, bW2T, d);
#pragma unroll
                for (int j = 0; j < 64; ++j) a[j] = ((m1 >> j) & 1ull) ? d[j] : 0.f;
                float de[CP];
                umma_layer<64, CP, 3>(u, leader, a, bW1T, de);
#pragma unroll
                for (int c = 0; c < C; c += 4)
                    *reinterpret_cast<float4*>(stage + tg * SP + c) = make_float4(de[c], de[c + 1], de[c + 2], de[c + 3]);
            }
            mem_lock(mlock, leader, group);
            {   // cooperative pass: d(de · f_k)/d(ix,iy) per plane.  4 lanes share one point: lane sl takes the 16-byte
                // channel chunks sl, sl+4, ... of every tap and accumulates its partial dot products before the
                // two-step shuffle reduction over the 4 lanes (8 points per warp iteration).
                constexpr int U = C / 4, STEPS = (U + 3) / 4;
                const int lane = tid & 31, seg = lane >> 2, sl = lane & 3;
                const int wpt0 = (tg >> 5) * 32;
#pragma unroll 1
                for (int r = 0; r < 4; ++r) {
                    const int pt = wpt0 + r * 8 + seg;
                    int o[12];
#pragma unroll
                    for (int t = 0; t < 12; t += 4) {
                        const int4 o4 = *reinterpret_cast<const int4*>(tap_o + pt * 12 + t);
                        o[t] = o4.x; o[t + 1] = o4.y; o[t + 2] = o4.z; o[t + 3] = o4.w;
                    }
                    const float* pbp = planes + (size_t)pbase[pt] * 6 * ps;
                    float A[12];
#pragma unroll
                    for (int t = 0; t < 12; ++t) A[t] = 0.f;
#pragma unroll
                    for (int st = 0; st < STEPS; ++st) {
                        const int ch = sl + 4 * st;
                        if (ch < U) {
                            const float4 dv = *reinterpret_cast<const float4*>(stage + pt * SP + ch * 4);
                            const float* base = pbp + ch * 4;
                            float4 q[12];
#pragma unroll
                            for (int t = 0; t < 12; ++t) q[t] = ldg4(base + (size_t)(t >> 2) * ps + (size_t)o[t] * C);
#pragma unroll
                            for (int t = 0; t < 12; ++t)
                                A[t] = fmaf(dv.x, q[t].x, fmaf(dv.y, q[t].y, fmaf(dv.z, q[t].z, fmaf(dv.w, q[t].w, A[t]))));
                        }
                    }
                    float part[6];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float4 tv = *reinterpret_cast<const float4*>(tap_v + pt * 12 + k * 4);
                        const float4 tf = *reinterpret_cast<const float4*>(tap_f + pt * 12 + k * 4);     // wx0 wx1 wy0 wy1
                        const float a0 = A[k * 4] * tv.x, a1 = A[k * 4 + 1] * tv.y, a2 = A[k * 4 + 2] * tv.z, a3 = A[k * 4 + 3] * tv.w;
                        part[k * 2] = (a1 - a0) * tf.z + (a3 - a2) * tf.w;
                        part[k * 2 + 1] = (a2 - a0) * tf.x + (a3 - a1) * tf.y;
                    }
#pragma unroll
                    for (int qd = 0; qd < 6; ++qd) {
                        float v = part[qd];
                        v += __shfl_xor_sync(0xffffffffu, v, 1);
                        v += __shfl_xor_sync(0xffffffffu, v, 2);
                        if (sl == 0) nacc[pt * 6 + qd] = v;
                    }
                }
            }
            mem_unlock(mlock, leader, group);
            if (valid && (grad_o || normal_o)) {
                float gm[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 3; ++k) { gm[plane_ax(k)] += nacc[tg * 6 + k * 2]; gm[plane_ay(k)] += nacc[tg * 6 + k * 2 + 1]; }
                const float scale = 0.5f * (float)cfg.R / cfg.radius;
                const float inv = nrm > 0.f ? 1.f / nrm : 0.f;
                float g[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) g[a] = gm[a] * scale + x[a] * inv;
                if (grad_o) { grad_o[id * 3] = g[0]; grad_o[id * 3 + 1] = g[1]; grad_o[id * 3 + 2] = g[2]; }
                if (normal_o) {
                    float n[3], len; normalize3(g, n, len);
                    normal_o[id * 3] = n[0]; normal_o[id * 3 + 1] = n[1]; normal_o[id * 3 + 2] = n[2];
                }
            }
        }
        group_sync(group);     // tap table / stage are rewritten by the next tile
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_warp(*tmem_slot, 512);
}

// ---- colour decoder at a list of samples -----------------------------------------------------------------------
template <int C>
struct TexSmem {
    static constexpr int W1H = 0, W1L = W1H + 3 * 64 * C, W2H = W1L + 3 * 64 * C, W2L = W2H + 4096, W3 = W2L + 4096;
    static constexpr int GROUP0 = W3 + 192;
    static constexpr int TAP_O = 0, TAP_W = TAP_O + 128 * 4, PBASE = TAP_W + 128 * 4, STAGE = PBASE + 128;
    static constexpr int GROUP_FLOATS = STAGE + 128 * (C + 4);
    static constexpr int TOTAL = GROUP0 + TC_GROUPS * GROUP_FLOATS + 16;
};

template <int C>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tex_tc(const float* __restrict__ planes, const float* __restrict__ wp,
                                                         tt_config cfg, TcSrc src, int64_t N, float* feat_o,
                                                         uint64_t* masks_o) {
    TT_SHARED(smem);
    using L = TexSmem<C>;
    constexpr int SP = C + 4;
    const int tid = threadIdx.x, group = tid / TC_GROUP, tg = tid % TC_GROUP, warp = tid >> 5;
    const WOff wo = woff(C);
    for (int k = 0; k < 3; ++k)      // W1f as three [64][C] K-major tiles (one per texture plane)
        btile_fill(smem + L::W1H + k * 64 * C, smem + L::W1L + k * 64 * C, 64, C,
                   [&](int n, int kk) { return __ldg(wp + wo.w1f + n * 3 * C + k * C + kk); }, tid, TC_THREADS);
    btile_fill(smem + L::W2H, smem + L::W2L, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2f + n * 64 + k); }, tid, TC_THREADS);
    if (tid < 192) smem[L::W3 + tid] = __ldg(wp + wo.w3f + tid);
    uint64_t* mbars = reinterpret_cast<uint64_t*>(smem + L::GROUP0 + TC_GROUPS * L::GROUP_FLOATS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbars + TC_GROUPS);
    int* mlock = reinterpret_cast<int*>(tmem_slot + 1);
    if (tid == 0) { for (int g = 0; g < TC_GROUPS; ++g) mbar_init(mbars + g); *mlock = 0; }
    if (warp == 0) tmem_alloc_warp(tmem_slot, 512);
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
    float* gs = smem + L::GROUP0 + group * L::GROUP_FLOATS;
    int* tap_o = reinterpret_cast<int*>(gs + L::TAP_O);
    float* tap_w = gs + L::TAP_W;
    uint32_t* pbase = reinterpret_cast<uint32_t*>(gs + L::PBASE);
    float* stage = gs + L::STAGE;
    const float* w3 = smem + L::W3;
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const int64_t n_live = src.count ? (int64_t)*src.count : N;
    const int64_t n_tiles = (n_live + TC_GROUP - 1) / TC_GROUP;

    // raw position words one tile ahead, sample id two tiles ahead (as k_tex_tc1)
    const int64_t tstride = (int64_t)gridDim.x * TC_GROUPS, tile0 = (int64_t)blockIdx.x * TC_GROUPS + group;
    WsRaw raw_next = ws_load_raw(src, ws_load_id(src, tile0, tg, n_live, n_tiles));
    int id_after = ws_load_id(src, tile0 + tstride, tg, n_live, n_tiles);
    for (int64_t tile = tile0; tile < n_tiles; tile += tstride) {
        const WsRaw raw = raw_next;
        raw_next = ws_load_raw(src, id_after);
        id_after = ws_load_id(src, tile + 2 * tstride, tg, n_live, n_tiles);
        const bool valid = raw.id >= 0;
        const int64_t id = valid ? raw.id : 0;
        float p[3] = {0.f, 0.f, 0.f}; int prompt = 0;
        if (valid) {
            float x[3];
            ws_point_from_raw(src, raw, x, prompt);
#pragma unroll
            for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
        }
        pbase[tg] = (uint32_t)prompt;
        float d[64];
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            Taps t;
            if (valid) t = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool in = valid && t.o[q] >= 0;
                tap_o[tg * 4 + q] = in ? t.o[q] : 0; tap_w[tg * 4 + q] = in ? t.w[q] : 0.f;
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
            umma_put_A<C>(u, e);
            group_sync(group);
            if (leader) { umma_mma<3>(u, bW1[k], C, k > 0); umma_commit(u); }
        }
        umma_wait(u);
        umma_get_D<64>(u, d);
        uint64_t m1 = 0, m2 = 0;
        {
            float h[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) { m1 |= (uint64_t)(d[j] > 0.f) << j; h[j] = fmaxf(d[j], 0.f); }
            umma_layer<64, 64, 3>(u, leader, h, bW2, d);
        }
        float f[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            const float h = fmaxf(d[j], 0.f);
            m2 |= (uint64_t)(d[j] > 0.f) << j;
            f[0] = fmaf(h, w3[j], f[0]); f[1] = fmaf(h, w3[64 + j], f[1]); f[2] = fmaf(h, w3[128 + j], f[2]);
        }
        if (valid) {
            if (feat_o) { feat_o[id * 3] = f[0]; feat_o[id * 3 + 1] = f[1]; feat_o[id * 3 + 2] = f[2]; }
            if (masks_o) { masks_o[id * 4] = m1; masks_o[id * 4 + 1] = m2; }     // ReLU masks for the backward
        }
        group_sync(group);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_warp(*tmem_slot, 512);
}

// ---- colour decoder, one gather per tile (C <= 32) ------------------------------------------------------------------------
// k_tex_tc above walks the three texture planes one after the other: taps, barrier, 4-tap gather, barrier, put A, wait for
// the previous plane's MMAs, barrier, issue -- three dependent round trips per tile before the first layer is done.  For
// C <= 32 the three A operands fit into disjoint TMEM columns (hi_k = 2kC, lo_k = 2kC + C, accumulator at 192), so this
// variant gathers all 12 taps of a point in ONE cooperative pass (24 loads in flight per lane like the geometry gather),
// puts the three operands, and issues the three MMA chains behind a single barrier.
template <int C>
struct Tex1Smem {
    static constexpr int SP3 = 3 * C + 4;
    static constexpr uint32_t COL_D = 192, COL_LO2 = 96;            // layer 2: hi [0,64) lo [96,160)
    static constexpr int W1H = 0, W1L = W1H + 3 * 64 * C, W2H = W1L + 3 * 64 * C, W2L = W2H + 4096, W3 = W2L + 4096;
    static constexpr int GROUP0 = W3 + 192;
    static constexpr int TAP_O = 0, TAP_W = TAP_O + 128 * 12, PBASE = TAP_W + 128 * 12, STAGE = PBASE + 128;
    static constexpr int GROUP_FLOATS = STAGE + 128 * SP3;
    static constexpr int TOTAL = GROUP0 + TC_GROUPS * GROUP_FLOATS + 16;
    static constexpr bool OK = 6 * C <= 192;
};

template <int C>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tex_tc1(const float* __restrict__ planes, const float* __restrict__ wp,
                                                          tt_config cfg, TcSrc src, int64_t N, float* feat_o,
                                                          uint64_t* masks_o) {
    TT_SHARED(smem);
    using L = Tex1Smem<C>;
    constexpr int SP3 = L::SP3, U = C / 4;
    const int tid = threadIdx.x, group = tid / TC_GROUP, tg = tid % TC_GROUP, warp = tid >> 5;
    const WOff wo = woff(C);
    for (int k = 0; k < 3; ++k)
        btile_fill(smem + L::W1H + k * 64 * C, smem + L::W1L + k * 64 * C, 64, C,
                   [&](int n, int kk) { return __ldg(wp + wo.w1f + n * 3 * C + k * C + kk); }, tid, TC_THREADS);
    btile_fill(smem + L::W2H, smem + L::W2L, 64, 64, [&](int n, int k) { return __ldg(wp + wo.w2f + n * 64 + k); }, tid, TC_THREADS);
    if (tid < 192) smem[L::W3 + tid] = __ldg(wp + wo.w3f + tid);
    uint64_t* mbars = reinterpret_cast<uint64_t*>(smem + L::GROUP0 + TC_GROUPS * L::GROUP_FLOATS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbars + TC_GROUPS);
    if (tid == 0) for (int g = 0; g < TC_GROUPS; ++g) mbar_init(mbars + g);
    if (warp == 0) tmem_alloc_warp(tmem_slot, 512);
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
    float* gs = smem + L::GROUP0 + group * L::GROUP_FLOATS;
    int* tap_o = reinterpret_cast<int*>(gs + L::TAP_O);
    float* tap_w = gs + L::TAP_W;
    uint32_t* pbase = reinterpret_cast<uint32_t*>(gs + L::PBASE);
    float* stage = gs + L::STAGE;
    const float* w3 = smem + L::W3;
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const int64_t n_live = src.count ? (int64_t)*src.count : N;
    const int64_t n_tiles = (n_live + TC_GROUP - 1) / TC_GROUP;

    // the sample's raw position words are loaded one tile ahead, its id two tiles ahead (dependent global loads that
    // would otherwise be exposed at every tile start)
    const int64_t tstride = (int64_t)gridDim.x * TC_GROUPS, tile0 = (int64_t)blockIdx.x * TC_GROUPS + group;
    WsRaw raw_next = ws_load_raw(src, ws_load_id(src, tile0, tg, n_live, n_tiles));
    int id_after = ws_load_id(src, tile0 + tstride, tg, n_live, n_tiles);
    for (int64_t tile = tile0; tile < n_tiles; tile += tstride) {
        const WsRaw raw = raw_next;
        raw_next = ws_load_raw(src, id_after);
        id_after = ws_load_id(src, tile + 2 * tstride, tg, n_live, n_tiles);
        const bool valid = raw.id >= 0;
        const int64_t id = valid ? raw.id : 0;
        float p[3] = {0.f, 0.f, 0.f}; int prompt = 0;
        if (valid) {
            float x[3];
            ws_point_from_raw(src, raw, x, prompt);
#pragma unroll
            for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const Taps t = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
            int4 o4; float4 w4;
            const bool i0 = valid && t.o[0] >= 0, i1 = valid && t.o[1] >= 0, i2 = valid && t.o[2] >= 0, i3 = valid && t.o[3] >= 0;
            o4.x = i0 ? t.o[0] : 0; o4.y = i1 ? t.o[1] : 0; o4.z = i2 ? t.o[2] : 0; o4.w = i3 ? t.o[3] : 0;
            w4.x = i0 ? t.w[0] : 0.f; w4.y = i1 ? t.w[1] : 0.f; w4.z = i2 ? t.w[2] : 0.f; w4.w = i3 ? t.w[3] : 0.f;
            *reinterpret_cast<int4*>(tap_o + tg * 12 + k * 4) = o4;
            *reinterpret_cast<float4*>(tap_w + tg * 12 + k * 4) = w4;
        }
        pbase[tg] = (uint32_t)prompt;
        group_sync(group);
        {   // cooperative gather of the three texture planes: item = (point, chunk), 12 loads each, 2 items in flight
            constexpr int JB = 2;
#pragma unroll 1
            for (int j0 = 0; j0 < U; j0 += JB) {
                float4 v[JB][12];
                int pt[JB], ch[JB];
#pragma unroll
                for (int b = 0; b < JB; ++b) {
                    const int j = j0 + b < U ? j0 + b : U - 1;
                    const int item = tg + TC_GROUP * j;
                    pt[b] = item / U; ch[b] = item - pt[b] * U;
                    const float* base = planes + ((size_t)pbase[pt[b]] * 6 + 3) * ps + ch[b] * 4;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int4 o4 = *reinterpret_cast<const int4*>(tap_o + pt[b] * 12 + k * 4);
                        const float* pb = base + (size_t)k * ps;
                        v[b][k * 4 + 0] = ldg4(pb + (size_t)o4.x * C); v[b][k * 4 + 1] = ldg4(pb + (size_t)o4.y * C);
                        v[b][k * 4 + 2] = ldg4(pb + (size_t)o4.z * C); v[b][k * 4 + 3] = ldg4(pb + (size_t)o4.w * C);
                    }
                }
#pragma unroll
                for (int b = 0; b < JB; ++b) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float4 w4 = *reinterpret_cast<const float4*>(tap_w + pt[b] * 12 + k * 4);
                        const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
                        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float4 q = v[b][k * 4 + t];
                            s.x = fmaf(ww[t], q.x, s.x); s.y = fmaf(ww[t], q.y, s.y); s.z = fmaf(ww[t], q.z, s.z); s.w = fmaf(ww[t], q.w, s.w);
                        }
                        *reinterpret_cast<float4*>(stage + pt[b] * SP3 + k * C + ch[b] * 4) = s;
                    }
                }
            }
        }
        group_sync(group);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float e[C];
#pragma unroll
            for (int c = 0; c < C; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(stage + tg * SP3 + k * C + c);
                e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
            }
            umma_put_A_ex<C>(u, e, (uint32_t)(k * 2 * C), (uint32_t)(k * 2 * C + C));
        }
        group_sync(group);
        if (leader) {
#pragma unroll
            for (int k = 0; k < 3; ++k) umma_mma_ex<3>(u, bW1[k], C, k > 0, (uint32_t)(k * 2 * C), (uint32_t)(k * 2 * C + C), L::COL_D);
            umma_commit(u);
        }
        umma_wait(u);
        float d[64];
        umma_get_D<64>(u, d, L::COL_D);
        uint64_t m1 = 0, m2 = 0;
        {
            float h[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) { m1 |= (uint64_t)(d[j] > 0.f) << j; h[j] = fmaxf(d[j], 0.f); }
            umma_put_A_ex<64>(u, h, 0u, L::COL_LO2);
        }
        group_sync(group);
        if (leader) { umma_mma_ex<3>(u, bW2, 64, false, 0u, L::COL_LO2, L::COL_D); umma_commit(u); }
        umma_wait(u);
        umma_get_D<64>(u, d, L::COL_D);
        float f[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            const float h = fmaxf(d[j], 0.f);
            m2 |= (uint64_t)(d[j] > 0.f) << j;
            f[0] = fmaf(h, w3[j], f[0]); f[1] = fmaf(h, w3[64 + j], f[1]); f[2] = fmaf(h, w3[128 + j], f[2]);
        }
        if (valid) {
            if (feat_o) { feat_o[id * 3] = f[0]; feat_o[id * 3 + 1] = f[1]; feat_o[id * 3 + 2] = f[2]; }
            if (masks_o) { masks_o[id * 4] = m1; masks_o[id * 4 + 1] = m2; }     // ReLU masks for the backward
        }
        group_sync(group);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_warp(*tmem_slot, 512);
}

// importance sampler, stage 2: proposal sdf [n_rays][n_imp] -> density -> cdf -> inverse-CDF draws -> sorted edges
__global__ void __launch_bounds__(128) k_sampler_post(tt_config cfg, int64_t n_rays, int n_imp, int n_fine,
                                                     const float* __restrict__ sdf, const float* __restrict__ jit0,
                                                     const float* __restrict__ jit1, float* __restrict__ cdf,
                                                     float* __restrict__ t_vals) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const bool strat = jit0 != nullptr;
    const float b0 = strat ? jit0[ray] : 0.f, b1 = strat ? jit1[ray] : 0.f;
    const float step = cfg.render_step_size;
    float run = 0.f;
    float t_lo = stot_u(quantile_s(0, n_imp, strat, b0), cfg.near_plane, cfg.far_plane);
    for (int j = 0; j < n_imp; ++j) {
        const float t_hi = stot_u(quantile_s(j + 1, n_imp, strat, b0), cfg.near_plane, cfg.far_plane);
        cdf[(size_t)j * n_rays + ray] = 1.f - expf(-run);
        const float s = sdf[ray * n_imp + j];
        const float pc = sigmoidf((s + step * 0.5f) * cfg.inv_std), nc = sigmoidf((s - step * 0.5f) * cfg.inv_std);
        const float alpha = fminf(fmaxf((pc - nc + 1e-5f) / (pc + 1e-5f), 0.f), 1.f);
        run += (alpha / step) * (t_hi - t_lo);
        t_lo = t_hi;
    }
    cdf[(size_t)n_imp * n_rays + ray] = 1.f;
    const int n_in = n_imp + 1, n_out = n_imp + n_fine + 2;
    float* out = t_vals + (size_t)ray * n_out;
    int k = 0; float last = -3.0e38f;
    auto emit = [&](float v) {
        if (v >= last) { out[k] = v; last = v; }
        else { int j = k; while (j > 0 && out[j - 1] > v) { out[j] = out[j - 1]; --j; } out[j] = v; }
        ++k;
    };
    int a = 0;
    float ta = stot_u(quantile_s(0, n_imp, strat, b0), cfg.near_plane, cfg.far_plane);
    int p = 0;
    for (int j = 0; j <= n_fine; ++j) {
        const float uq = quantile_s(j, n_fine, strat, b1);
        while (p < n_in && cdf[(size_t)p * n_rays + ray] <= uq) ++p;
        const int pc = min(max(p, 1), n_in - 1);
        const float c0 = cdf[(size_t)(pc - 1) * n_rays + ray], c1 = cdf[(size_t)pc * n_rays + ray];
        const float v0 = quantile_s(pc - 1, n_imp, strat, b0), v1 = quantile_s(pc, n_imp, strat, b0);
        const float den = __fsub_rn(c1, c0);
        float fr = den > 0.f ? __fdiv_rn(__fsub_rn(uq, c0), den) : 0.f;
        fr = fminf(fmaxf(fr, 0.f), 1.f);
        const float tf = stot_u(__fadd_rn(v0, __fmul_rn(fr, __fsub_rn(v1, v0))), cfg.near_plane, cfg.far_plane);
        while (a < n_in && ta <= tf) {
            emit(ta); ++a;
            if (a < n_in) ta = stot_u(quantile_s(a, n_imp, strat, b0), cfg.near_plane, cfg.far_plane);
        }
        emit(tf);
    }
    while (a < n_in) {
        emit(ta); ++a;
        if (a < n_in) ta = stot_u(quantile_s(a, n_imp, strat, b0), cfg.near_plane, cfg.far_plane);
    }
}

}  // namespace tt
