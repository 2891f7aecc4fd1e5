// Kernels + C ABI of libtriplane_b200.so (see include/triplane_b200.h).  sm_100a only, no CPU path.
#include "../../include/triplane_b200.h"
#include "tt_device.cuh"
#include "tt_tc.cuh"
#include "tt_tc_bwd.cuh"
#include "tt_ws.cuh"
#include "tt_rays.cuh"
#include "tt_sampler.cuh"

static_assert(sizeof(tt_config) == 56, "tt_config is part of the C ABI: 14 x 4 bytes (INTEGRATION.md section 2)");
#include <atomic>
#include <cstdlib>
#include <cstdio>
#include <cstring>

using namespace tt;

// =====================================================================================================
// host-side helpers
// =====================================================================================================
static thread_local char g_err[512] = "";
static int g_impl = 2;      // 2: warp-specialised tcgen05 kernels (tt_ws.cuh; default), 1: round-1 tcgen05 kernels
                            // (tt_tc.cuh), 0: SIMT reference kernels (this file)
static const size_t kMaxSmem = 227 * 1024;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* fmt, const char* a = "", long long b = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
static int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return TT_E_CUDA;
    }
    return TT_OK;
}
// ---- optional per-launch timing (CUDA events on the launching stream), off by default ------------------
#ifndef TT_EMUL
#include <string>
#include <vector>
struct ProfRec { std::string name; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static cudaEvent_t g_prof_pending = nullptr;
void tt_prof_pre(cudaStream_t st) {
    if (!g_prof_on) return;
    cudaEventCreate(&g_prof_pending);
    cudaEventRecord(g_prof_pending, st);
}
void tt_prof_post(const char* name, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfRec r; r.name = name; r.a = g_prof_pending;
    cudaEventCreate(&r.b);
    cudaEventRecord(r.b, st);
    g_prof.push_back(r);
}
#endif

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool supported_C(int C) { return C == 8 || C == 16 || C == 32 || C == 40 || C == 64; }

#define TT_DISPATCH_C(C, ...)                                                    \
    switch (C) {                                                                 \
        case 8:  { constexpr int kC = 8;  __VA_ARGS__; } break;                  \
        case 16: { constexpr int kC = 16; __VA_ARGS__; } break;                  \
        case 32: { constexpr int kC = 32; __VA_ARGS__; } break;                  \
        case 40: { constexpr int kC = 40; __VA_ARGS__; } break;                  \
        case 64: { constexpr int kC = 64; __VA_ARGS__; } break;                  \
        default: return fail(TT_E_ARG, "unsupported channel count%s %lld (8,16,32,40,64)", "", (long long)(C)); \
    }

static int check_cfg(const tt_config* cfg) {
    if (!cfg) return fail(TT_E_ARG, "cfg is NULL%s", "");
    if (!supported_C(cfg->C)) return fail(TT_E_ARG, "unsupported channel count%s %lld (8,16,32,40,64)", "", cfg->C);
    if (cfg->R < 2 || cfg->R > 8192) return fail(TT_E_ARG, "plane resolution out of range%s: %lld", "", cfg->R);
    if (cfg->P < 1) return fail(TT_E_ARG, "P must be >= 1%s (got %lld)", "", cfg->P);
    if (!(cfg->radius > 0.f)) return fail(TT_E_ARG, "radius must be > 0%s", "");
    return TT_OK;
}
template <typename K>
static int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(%zu B smem): %s", bytes, cudaGetErrorString(e)); return TT_E_CUDA; }
    return TT_OK;
}
static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}
static unsigned tc_grid(int64_t n_points) {
    const int64_t ctas = (n_points + TC_THREADS - 1) / TC_THREADS;
    return (unsigned)(ctas < (int64_t)num_sms() ? (ctas < 1 ? 1 : ctas) : num_sms());
}
// SDF decoder on a point list: warp-specialised kernel (impl 2) or the round-1 kernel (impl 1).  vscratch: the L2-resident
// tangent-row buffer of the warp-specialised normal path (ws_vscratch_floats(num_sms(), C) floats), may be NULL.
template <int kC, bool NORMAL>
static int launch_geo_decoder(const float* planes, const float* wpack, const tt_config* cfg, const TcSrc& src, int64_t N,
                              float* sdf, float* sdf_orig, float* grad, float* normal, uint64_t* masks, float* vscratch,
                              cudaStream_t st) {
    const size_t smw = (size_t)GeoWs<kC, NORMAL>::TOTAL * 4;
    if (g_impl == 2 && smw <= kMaxSmem && (!NORMAL || vscratch)) {
        if (int e = set_smem(k_geo_ws<kC, NORMAL>, smw)) return e;
        const int64_t ctas = (N + WS_CG * TC_GROUP - 1) / (WS_CG * TC_GROUP);
        const unsigned grid = (unsigned)(ctas < (int64_t)num_sms() ? (ctas < 1 ? 1 : ctas) : num_sms());
        TT_LAUNCH((k_geo_ws<kC, NORMAL>), grid, WS_GEO_THREADS, smw, st, planes, wpack, *cfg, src, N, sdf, sdf_orig, grad, normal, masks, vscratch, (float*)nullptr);
        return check_launch("k_geo_ws");
    }
    const size_t smg = (size_t)GeoSmem<kC, NORMAL>::TOTAL * 4;
    if (int e = set_smem(k_geo_tc<kC, NORMAL>, smg)) return e;
    TT_LAUNCH((k_geo_tc<kC, NORMAL>), tc_grid(N), TC_THREADS, smg, st, planes, wpack, *cfg, src, N, sdf, sdf_orig, grad, normal, masks, (float*)nullptr);
    return check_launch("k_geo_tc");
}
static size_t round4(size_t n) { return (n + 3) / 4 * 4; }
static size_t ws_scratch_floats() { return ws_vscratch_floats(num_sms(), 40) + 16; }      // largest C of the ws kernels

// colour decoder on a point list
template <int kC>
static int launch_tex_decoder(const float* planes, const float* wpack, const tt_config* cfg, const TcSrc& src, int64_t N,
                              float* features, uint64_t* masks, cudaStream_t st) {
    // The warp-specialised colour kernel is parity-green but slower than the two-group kernel at both bench configurations
    // (config 3: 185.6 vs 164.7 ms, config 2: 47.6 vs 43.6 ms, profiles/r02_bench_ws3_*.json): one gather warpgroup issues
    // fewer loads than two groups gathering for themselves.  Opt in with -DTT_TEX_WS=1.
#ifndef TT_TEX_WS
#define TT_TEX_WS 0
#endif
    const size_t smw = (size_t)TexWs<kC>::TOTAL * 4;
    if (TT_TEX_WS && g_impl == 2 && smw <= kMaxSmem) {
        if (int e = set_smem(k_tex_ws<kC>, smw)) return e;
        const int64_t tiles = (N + TC_GROUP - 1) / TC_GROUP;
        const unsigned grid = (unsigned)(tiles < (int64_t)num_sms() ? (tiles < 1 ? 1 : tiles) : num_sms());
        TT_LAUNCH(k_tex_ws<kC>, grid, WS_THREADS, smw, st, planes, wpack, *cfg, src, N, features, masks);
        return check_launch("k_tex_ws");
    }
    const size_t sm1 = (size_t)Tex1Smem<kC>::TOTAL * 4;
    if (g_impl == 2 && Tex1Smem<kC>::OK && sm1 <= kMaxSmem) {      // one gather per tile, three A operands in disjoint TMEM columns
        if (int e = set_smem(k_tex_tc1<kC>, sm1)) return e;
        TT_LAUNCH(k_tex_tc1<kC>, tc_grid(N), TC_THREADS, sm1, st, planes, wpack, *cfg, src, N, features, masks);
        return check_launch("k_tex_tc1");
    }
    const size_t smt = (size_t)TexSmem<kC>::TOTAL * 4;
    if (int e = set_smem(k_tex_tc<kC>, smt)) return e;
    TT_LAUNCH(k_tex_tc<kC>, tc_grid(N), TC_THREADS, smt, st, planes, wpack, *cfg, src, N, features, masks);
    return check_launch("k_tex_tc");
}

// switches (tt_set_option; initial values from the environment): "scatter" = -1 auto, 0 plain, 1 run-length,
// 2 tile-merged hidden-gradient scatter; "patch_lists" = 1 (default): patch-ordered sample lists when the image shape is known;
// "grid_lines" = 1: z-line gather of the regular isosurface grid (ws_grid_segment, tt_ws.cuh)
static int g_opt_scatter = -2, g_opt_patch = -1, g_opt_grid_lines = 0;
static int scatter_mode() {
    if (g_opt_scatter == -2) { const char* e = getenv("TT_SCATTER"); g_opt_scatter = e ? atoi(e) : -1; }
    return g_opt_scatter;
}
static bool patch_mode() {
    if (g_opt_patch < 0) { const char* e = getenv("TT_PATCH_LISTS"); g_opt_patch = e ? (atoi(e) != 0) : 1; }
    return g_opt_patch != 0;
}
// colour backward: SC = scatter variant of the hidden gradient, P3 = 3xTF32 layers (TT_FLAG_PRECISE_BWD)
template <int kC, int SC, bool P3>
static int launch_bwd_tex(const float* planes, const float* wpack, const tt_config* cfg, const TcSrc& ts, int64_t N, int64_t tiles,
                          const float* gf, const uint64_t* tex_masks, float* hid, float* gw, cudaStream_t st) {
    constexpr int GT = BwdTexSmem<kC, P3>::G;
    const size_t smx = (size_t)BwdTexSmem<kC, P3>::TOTAL * 4;
    const int64_t ctas_t = (tiles + GT - 1) / GT;
    const unsigned grid_t = (unsigned)(ctas_t < (int64_t)num_sms() ? ctas_t : num_sms());
    if (int e = set_smem((k_bwd_tex_tc<kC, SC, P3>), smx)) return e;
    TT_LAUNCH((k_bwd_tex_tc<kC, SC, P3>), grid_t, GT * TC_GROUP, smx, st, planes, wpack, *cfg, ts, N, gf, tex_masks, hid, gw);
    return check_launch("k_bwd_tex_tc");
}

static inline size_t slab_bytes(int rows) { return (size_t)rows * ST * sizeof(float); }
static inline int imax(int a, int b) { return a > b ? a : b; }

// =====================================================================================================
// weights
// =====================================================================================================
__global__ void k_pack_weights(const float* s0, const float* s1, const float* s2, const float* f0, const float* f1,
                               const float* f2, const float* d0, const float* d1, const float* d2, int C, float* wp) {
    const WOff wo = woff(C);
    const int n = blockDim.x * gridDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    auto both = [&](const float* src, int rows, int cols, int dst, int dstT) {   // src [rows][cols]
        if (!src) return;
        for (int i = t; i < rows * cols; i += n) {
            const int r = i / cols, c = i % cols;
            const float v = src[i];
            wp[dst + i] = v;
            wp[dstT + c * rows + r] = v;
        }
    };
    auto plain = [&](const float* src, int cnt, int dst) {
        if (!src) return;
        for (int i = t; i < cnt; i += n) wp[dst + i] = src[i];
    };
    both(s0, 64, C, wo.w1s, wo.w1sT); both(s1, 64, 64, wo.w2s, wo.w2sT); plain(s2, 64, wo.w3s);
    both(f0, 64, 3 * C, wo.w1f, wo.w1fT); both(f1, 64, 64, wo.w2f, wo.w2fT); plain(f2, 192, wo.w3f);
    both(d0, 64, C, wo.w1d, wo.w1dT); both(d1, 64, 64, wo.w2d, wo.w2dT); plain(d2, 192, wo.w3d);
}

// =====================================================================================================
// plane repack: NCHW (un-rotated) <-> channel-last (rotated)
// rotated[h][w] = src[sh][sw]:  plane%3==0: transpose (sh=w, sw=h); ==1: rot90 k=2 (sh=R-1-h, sw=R-1-w);
// ==2: rot90 k=-1 (sh=R-1-w, sw=h)            (few_step…diffusion.py:212-225)
// =====================================================================================================
__device__ __forceinline__ void rot_dst(int k3, int R, int sh, int sw, int& h, int& w) {
    if (k3 == 0) { h = sw; w = sh; }
    else if (k3 == 1) { h = R - 1 - sh; w = R - 1 - sw; }
    else { h = sw; w = R - 1 - sh; }
}
// grid: (ceil(R/32), R, P*6); block (32, 8).  Tile = one source row segment of 32 texels, all channels.
template <bool BWD>
__global__ void k_repack(const float* __restrict__ src, float* __restrict__ dst, int Csrc, int off_geo, int off_tex,
                         int C, int R) {
    TT_SHARED(tile);                         // [C][33]
    const int pk = blockIdx.z, k = pk % 6, sh = blockIdx.y, sw0 = blockIdx.x * 32;
    const int coff = (k < 3) ? off_geo : off_tex;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const size_t nchw_plane = (size_t)pk * Csrc * R * R;
    const size_t cl_plane = (size_t)pk * R * R * C;
    if (!BWD) {
        for (int c = ty; c < C; c += 8)
            if (sw0 + tx < R) tile[c * 33 + tx] = src[nchw_plane + ((size_t)(coff + c) * R + sh) * R + sw0 + tx];
        __syncthreads();
        for (int i = ty; i < 32 && sw0 + i < R; i += 8) {
            int h, w; rot_dst(k % 3, R, sh, sw0 + i, h, w);
            for (int c = tx; c < C; c += 32) dst[cl_plane + ((size_t)h * R + w) * C + c] = tile[c * 33 + i];
        }
    } else {   // src = channel-last gradient, dst = NCHW gradient (Csrc == C, offsets 0)
        for (int i = ty; i < 32 && sw0 + i < R; i += 8) {
            int h, w; rot_dst(k % 3, R, sh, sw0 + i, h, w);
            for (int c = tx; c < C; c += 32) tile[c * 33 + i] = src[cl_plane + ((size_t)h * R + w) * C + c];
        }
        __syncthreads();
        for (int c = ty; c < C; c += 8)
            if (sw0 + tx < R) dst[nchw_plane + ((size_t)(coff + c) * R + sh) * R + sw0 + tx] = tile[c * 33 + tx];
    }
}

// =====================================================================================================
// point sources
// =====================================================================================================
struct RaySrc {
    const float* rays_o; const float* rays_d;
    const float* t_starts; const float* t_ends; int64_t t_stride; int S;
};
// isosurface grid vertex (threestudio/models/isosurface.py:37-51): linspace(0,1,res) -> scale to (-1,1)
__device__ __forceinline__ float grid_coord(int i, int res) {
    // torch.linspace(0, 1, res): step = 1/(res-1); first half start + i*step, second half end - (res-1-i)*step
    const float step = __fdiv_rn(1.f, (float)(res - 1));
    float v = (i < res / 2) ? __fmul_rn((float)i, step) : __fsub_rn(1.f, __fmul_rn((float)(res - 1 - i), step));
    v = __fadd_rn(__fmul_rn(v, 1.f), 0.f);
    return __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v, 0.f), 1.f), 2.f), -1.f);
}

// =====================================================================================================
// geometry on point lists (forward / forward_sdf / forward_field / export)
// =====================================================================================================
template <int C>
__global__ void __launch_bounds__(TPB) k_geometry_fwd(const float* __restrict__ planes, const float* __restrict__ wp,
                                                     tt_config cfg, const float* __restrict__ points, int64_t M,
                                                     int grid_res, float* sdf_o, float* sdf_orig_o, float* feat_o,
                                                     float* normal_o, float* grad_o, float* deform_o) {
    TT_SHARED(smem);
    constexpr int RX = C > HID ? C : HID;
    float* slotX = smem + threadIdx.x;
    float* slotB = smem + RX * ST + threadIdx.x;
    const WOff wo = woff(C);
    const int64_t N = (int64_t)cfg.P * M;
    const int64_t idx = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (idx >= N) return;
    const int prompt = (int)(idx / M);
    float x[3];
    if (points) { x[0] = points[idx * 3]; x[1] = points[idx * 3 + 1]; x[2] = points[idx * 3 + 2]; }
    else {
        const int64_t v = idx % M;
        const int iz = (int)(v % grid_res), iy = (int)((v / grid_res) % grid_res), ixx = (int)(v / ((int64_t)grid_res * grid_res));
        x[0] = grid_coord(ixx, grid_res); x[1] = grid_coord(iy, grid_res); x[2] = grid_coord(iz, grid_res);
    }
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const float* geo = planes + (size_t)prompt * 6 * ps;
    const bool want_n = normal_o || grad_o;
    float so, s, g[3] = {0.f, 0.f, 0.f};
    if (want_n) geo_eval<C, true>(geo, cfg.R, wp, wo, x, cfg.radius, cfg.sdf_bias_radius, slotX, slotB, so, s, g);
    else geo_eval<C, false>(geo, cfg.R, wp, wo, x, cfg.radius, cfg.sdf_bias_radius, slotX, slotB, so, s, g);
    if (sdf_o) sdf_o[idx] = s;
    if (sdf_orig_o) sdf_orig_o[idx] = so;
    if (grad_o) { grad_o[idx * 3] = g[0]; grad_o[idx * 3 + 1] = g[1]; grad_o[idx * 3 + 2] = g[2]; }
    if (normal_o) {
        float n[3], len; normalize3(g, n, len);
        normal_o[idx * 3] = n[0]; normal_o[idx * 3 + 1] = n[1]; normal_o[idx * 3 + 2] = n[2];
    }
    if (deform_o) {   // deformation MLP on the same geometry encoding (few_step…diffusion.py:375-394); slotX still holds it
        float acc[HID];
        zero64(acc); layer64_acc(acc, wp + wo.w1dT, C, slotX); store_relu64(acc, slotB);
        zero64(acc); layer64_acc(acc, wp + wo.w2dT, HID, slotB);
        float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < HID; ++j) {
            const float h = fmaxf(acc[j], 0.f);
            d[0] = fmaf(h, __ldg(wp + wo.w3d + j), d[0]);
            d[1] = fmaf(h, __ldg(wp + wo.w3d + HID + j), d[1]);
            d[2] = fmaf(h, __ldg(wp + wo.w3d + 2 * HID + j), d[2]);
        }
        deform_o[idx * 3] = d[0]; deform_o[idx * 3 + 1] = d[1]; deform_o[idx * 3 + 2] = d[2];
    }
    if (feat_o) {
        float p[3], f[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
        tex_eval<C>(geo + 3 * ps, cfg.R, wp, wo, p, slotX, slotB, f);
        feat_o[idx * 3] = f[0]; feat_o[idx * 3 + 1] = f[1]; feat_o[idx * 3 + 2] = f[2];
    }
}

// =====================================================================================================
// importance sampling (one thread per ray)
// =====================================================================================================
__device__ __forceinline__ float quantile(int j, int n, bool strat, float b) {
    return strat ? __fdiv_rn(__fadd_rn((float)j, b), (float)(n + 1)) : __fdiv_rn((float)j, (float)n);
}
__device__ __forceinline__ float stot(float s, float tmin, float tmax) {    // estimators.py:104-118 (uniform)
    return __fadd_rn(__fmul_rn(s, tmax), __fmul_rn(__fsub_rn(1.f, s), tmin));
}
template <int C>
__global__ void __launch_bounds__(TPB) k_importance_sample(const float* __restrict__ planes, const float* __restrict__ wp,
                                                          tt_config cfg, const float* __restrict__ rays_o,
                                                          const float* __restrict__ rays_d, int64_t n_rays, int n_imp,
                                                          int n_fine, const float* __restrict__ jit0,
                                                          const float* __restrict__ jit1, float* __restrict__ cdf,
                                                          float* __restrict__ t_vals) {
    TT_SHARED(smem);
    constexpr int RX = C > HID ? C : HID;
    float* slotX = smem + threadIdx.x;
    float* slotB = smem + RX * ST + threadIdx.x;
    const WOff wo = woff(C);
    const int64_t ray = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (ray >= n_rays) return;
    const int prompt = (int)(ray / cfg.rays_per_cache);
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const float* geo = planes + (size_t)prompt * 6 * ps;
    const float o[3] = {rays_o[ray * 3], rays_o[ray * 3 + 1], rays_o[ray * 3 + 2]};
    const float d[3] = {rays_d[ray * 3], rays_d[ray * 3 + 1], rays_d[ray * 3 + 2]};
    const bool strat = jit0 != nullptr;
    const float b0 = strat ? jit0[ray] : 0.f, b1 = strat ? jit1[ray] : 0.f;
    const float step = cfg.render_step_size;
    // level 0: the [0,1] cdf maps quantile u to s = u (nerfacc importance_sampling on the unit interval)
    float run = 0.f;                              // exclusive sum of sigma*dt
    float s_lo = quantile(0, n_imp, strat, b0);
    float t_lo = stot(s_lo, cfg.near_plane, cfg.far_plane);
    for (int j = 0; j < n_imp; ++j) {
        const float s_hi = quantile(j + 1, n_imp, strat, b0);
        const float t_hi = stot(s_hi, cfg.near_plane, cfg.far_plane);
        cdf[(size_t)j * n_rays + ray] = 1.f - expf(-run);
        const float tm = __fmul_rn(__fadd_rn(t_lo, t_hi), 0.5f);
        const float x[3] = {__fadd_rn(o[0], __fmul_rn(d[0], tm)), __fadd_rn(o[1], __fmul_rn(d[1], tm)),
                            __fadd_rn(o[2], __fmul_rn(d[2], tm))};
        float so, s, g[3];
        geo_eval<C, false>(geo, cfg.R, wp, wo, x, cfg.radius, cfg.sdf_bias_radius, slotX, slotB, so, s, g);
        // proposal density (…sdf_volume_renderer.py:289-297)
        const float pc = sigmoidf((s + step * 0.5f) * cfg.inv_std), nc = sigmoidf((s - step * 0.5f) * cfg.inv_std);
        const float alpha = fminf(fmaxf((pc - nc + 1e-5f) / (pc + 1e-5f), 0.f), 1.f);
        const float sigma = alpha / step;
        run += sigma * (t_hi - t_lo);
        t_lo = t_hi;
    }
    cdf[(size_t)n_imp * n_rays + ray] = 1.f;
    // fine edges by inverse CDF + merge with the coarse edges into one sorted row
    const int n_in = n_imp + 1, n_out = n_imp + n_fine + 2;
    float* out = t_vals + (size_t)ray * n_out;
    int k = 0; float last = -3.0e38f;
    auto emit = [&](float v) {
        if (v >= last) { out[k] = v; last = v; }
        else { int j = k; while (j > 0 && out[j - 1] > v) { out[j] = out[j - 1]; --j; } out[j] = v; }
        ++k;
    };
    int a = 0;                                    // next coarse edge to emit
    float ta = stot(quantile(0, n_imp, strat, b0), cfg.near_plane, cfg.far_plane);
    int p = 0;                                    // searchsorted pointer (count of cdf <= u)
    for (int j = 0; j <= n_fine; ++j) {
        const float u = quantile(j, n_fine, strat, b1);
        while (p < n_in && cdf[(size_t)p * n_rays + ray] <= u) ++p;
        const int pc = min(max(p, 1), n_in - 1);
        const float c0 = cdf[(size_t)(pc - 1) * n_rays + ray], c1 = cdf[(size_t)pc * n_rays + ray];
        const float v0 = quantile(pc - 1, n_imp, strat, b0), v1 = quantile(pc, n_imp, strat, b0);
        const float den = __fsub_rn(c1, c0);
        float fr = den > 0.f ? __fdiv_rn(__fsub_rn(u, c0), den) : 0.f;
        fr = fminf(fmaxf(fr, 0.f), 1.f);
        const float sv = __fadd_rn(v0, __fmul_rn(fr, __fsub_rn(v1, v0)));
        const float tf = stot(sv, cfg.near_plane, cfg.far_plane);
        while (a < n_in && ta <= tf) {
            emit(ta); ++a;
            if (a < n_in) ta = stot(quantile(a, n_imp, strat, b0), cfg.near_plane, cfg.far_plane);
        }
        emit(tf);
    }
    while (a < n_in) {
        emit(ta); ++a;
        if (a < n_in) ta = stot(quantile(a, n_imp, strat, b0), cfg.near_plane, cfg.far_plane);
    }
}

// =====================================================================================================
// fused march + decoders + alpha + compositing (one thread per ray, lock-step over the S intervals)
// =====================================================================================================
template <int C>
__global__ void __launch_bounds__(TPB) k_render_fwd(const float* __restrict__ planes, const float* __restrict__ wp,
                                                   tt_config cfg, RaySrc rs, int64_t n_rays, float* __restrict__ acc_o,
                                                   float* sdf_o, float* sdf_orig_o, float* grad_o, float* normal_o,
                                                   float* feat_o, float* weights_o, float* trans_o) {
    TT_SHARED(smem);
    constexpr int RX = C > HID ? C : HID;
    float* slotX = smem + threadIdx.x;
    float* slotB = smem + RX * ST + threadIdx.x;
    const WOff wo = woff(C);
    const int64_t ray = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (ray >= n_rays) return;
    const int prompt = (int)(ray / cfg.rays_per_cache);
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const float* geo = planes + (size_t)prompt * 6 * ps;
    const float o[3] = {rs.rays_o[ray * 3], rs.rays_o[ray * 3 + 1], rs.rays_o[ray * 3 + 2]};
    const float d[3] = {rs.rays_d[ray * 3], rs.rays_d[ray * 3 + 1], rs.rays_d[ray * 3 + 2]};
    const float* t0p = rs.t_starts + ray * rs.t_stride;
    const float* t1p = rs.t_ends + ray * rs.t_stride;
    const int S = rs.S;
    float T = 1.f, opac = 0.f, depth = 0.f, rgb[3] = {0.f, 0.f, 0.f}, nsum[3] = {0.f, 0.f, 0.f};
    float wsum = 0.f, mean = 0.f, m2 = 0.f;      // weighted Welford for z_variance
    float eik = 0.f;                             // Σ (|sdf_grad| - 1)^2 (eikonal term, un-weighted)
    for (int i = 0; i < S; ++i) {
        const float t0 = t0p[i], t1 = t1p[i];
        const float tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f), dt = __fsub_rn(t1, t0);
        const float x[3] = {__fadd_rn(o[0], __fmul_rn(d[0], tm)), __fadd_rn(o[1], __fmul_rn(d[1], tm)),
                            __fadd_rn(o[2], __fmul_rn(d[2], tm))};
        float so, s, g[3];
        geo_eval<C, true>(geo, cfg.R, wp, wo, x, cfg.radius, cfg.sdf_bias_radius, slotX, slotB, so, s, g);
        float n[3], len; normalize3(g, n, len);
        const AlphaTerms at = neus_alpha(s, n, d, dt, cfg.inv_std, cfg.cos_anneal_ratio);
        eik += (len - 1.f) * (len - 1.f);
        float p[3], f[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
        tex_eval<C>(geo + 3 * ps, cfg.R, wp, wo, p, slotX, slotB, f);
        const float w = T * at.alpha;
        opac += w; depth = fmaf(w, tm, depth);
#pragma unroll
        for (int a = 0; a < 3; ++a) { rgb[a] = fmaf(w, sigmoid_mipnerf(f[a]), rgb[a]); nsum[a] = fmaf(w, n[a], nsum[a]); }
        const float wn = wsum + w;
        if (wn > 0.f) { const float dl = tm - mean; mean += (w / wn) * dl; m2 += w * dl * (tm - mean); }
        wsum = wn;
        const int64_t si = ray * S + i;
        if (sdf_o) sdf_o[si] = s;
        if (sdf_orig_o) sdf_orig_o[si] = so;
        if (grad_o) { grad_o[si * 3] = g[0]; grad_o[si * 3 + 1] = g[1]; grad_o[si * 3 + 2] = g[2]; }
        if (normal_o) { normal_o[si * 3] = n[0]; normal_o[si * 3 + 1] = n[1]; normal_o[si * 3 + 2] = n[2]; }
        if (feat_o) { feat_o[si * 3] = f[0]; feat_o[si * 3 + 1] = f[1]; feat_o[si * 3 + 2] = f[2]; }
        if (weights_o) weights_o[si] = w;
        if (trans_o) trans_o[si] = T;
        T *= (1.f - at.alpha);
    }
    float* a = acc_o + ray * TT_ACC;
    a[0] = opac; a[1] = depth; a[2] = rgb[0]; a[3] = rgb[1]; a[4] = rgb[2];
    a[5] = m2 + wsum * (mean - depth) * (mean - depth);      // Σ w (t - depth)^2, depth un-normalised
    a[6] = nsum[0]; a[7] = nsum[1]; a[8] = nsum[2];
    a[9] = eik;
}

// backward, stage 1 (compositing + alpha + normalisation -> per-sample seeds): k_render_bwd_comp in tt_rays.cuh

// seeds for the stand-alone geometry backward: gs = g_sdf, u = g_sdf_grad + d normalize, gf = g_features
__global__ void k_geometry_bwd_seed(int64_t N, const float* __restrict__ grad /*sdf_grad, needed iff g_normal*/,
                                    const float* g_sdf, const float* g_feat, const float* g_normal,
                                    const float* g_grad, float* gs_o, float* u_o, float* gf_o) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    gs_o[i] = g_sdf ? g_sdf[i] : 0.f;
    float u[3] = {0.f, 0.f, 0.f};
    if (g_normal) {
        const float g[3] = {grad[i * 3], grad[i * 3 + 1], grad[i * 3 + 2]};
        const float gn[3] = {g_normal[i * 3], g_normal[i * 3 + 1], g_normal[i * 3 + 2]};
        float n[3], len; normalize3(g, n, len);
        if (len > 1e-12f) {
            const float dotv = n[0] * gn[0] + n[1] * gn[1] + n[2] * gn[2];
#pragma unroll
            for (int a = 0; a < 3; ++a) u[a] = (gn[a] - n[a] * dotv) / len;
        } else {
#pragma unroll
            for (int a = 0; a < 3; ++a) u[a] = gn[a] * 1e12f;
        }
    }
    if (g_grad) { u[0] += g_grad[i * 3]; u[1] += g_grad[i * 3 + 1]; u[2] += g_grad[i * 3 + 2]; }
    u_o[i * 3] = u[0]; u_o[i * 3 + 1] = u[1]; u_o[i * 3 + 2] = u[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) gf_o[i * 3 + a] = g_feat ? g_feat[i * 3 + a] : 0.f;
}

// =====================================================================================================
// backward, stage 2a: SDF decoder + geometry planes, one thread per sample point.
//   L depends on sdf (seed gs) and on sdf_grad = d sdf/d x (seed u).  With the tap weights
//   ω = gs·w + ẇ (ẇ = directional derivative of the bilinear weights along u) and ẽ = Σ ω·texel:
//     dL/dW1 = a1 ẽᵀ, dL/dW2 = a2 h̃1ᵀ, dL/dw3 = h̃2,  h̃1 = m1⊙W1ẽ, h̃2 = m2⊙W2h̃1,
//     dL/dtexel = (W1ᵀa1)·ω          (a1, a2: unit-seed adjoints of the SDF MLP)
//   which is the function of grid_sample_gradfix/gridsample_cuda.cu:87-209 fused with the MLP terms.
// =====================================================================================================
struct PtSrc {
    const float* points; int64_t M;     // explicit points [P*M][3] (prompt = idx / M) or NULL
    RaySrc rs; int rays_per_cache;      // else sample idx = ray*S + i
};
__device__ __forceinline__ void point_of(const PtSrc& src, int64_t idx, float (&x)[3], int& prompt) {
    if (src.points) {
        x[0] = src.points[idx * 3]; x[1] = src.points[idx * 3 + 1]; x[2] = src.points[idx * 3 + 2];
        prompt = (int)(idx / src.M);
    } else {
        const int64_t ray = idx / src.rs.S; const int i = (int)(idx % src.rs.S);
        const float t0 = src.rs.t_starts[ray * src.rs.t_stride + i], t1 = src.rs.t_ends[ray * src.rs.t_stride + i];
        const float tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
#pragma unroll
        for (int a = 0; a < 3; ++a) x[a] = __fadd_rn(src.rs.rays_o[ray * 3 + a], __fmul_rn(src.rs.rays_d[ray * 3 + a], tm));
        prompt = (int)(ray / src.rays_per_cache);
    }
}

template <int C>
__global__ void __launch_bounds__(TPB) k_bwd_geo(const float* __restrict__ planes, const float* __restrict__ wp,
                                                tt_config cfg, PtSrc src, int64_t N, const float* __restrict__ gs_i,
                                                const float* __restrict__ u_i, float* __restrict__ gplanes,
                                                float* __restrict__ gw) {
    TT_SHARED(smem);
    constexpr int RE = C;
    float* sE = smem;                       // C rows:  enc -> ẽ
    float* sB = sE + RE * ST;               // 64 rows: h1 -> a2 -> a1 -> a2 -> h̃2
    float* sD = sB + HID * ST;              // C rows:  W1ᵀ a1
    float* sT = sD + C * ST;                // 64 rows: h̃1
    const int tid = threadIdx.x;
    const WOff wo = woff(C);
    const GOff go = goff(C);
    const int64_t idx = (int64_t)blockIdx.x * TPB + tid;
    float gs = 0.f, u[3] = {0.f, 0.f, 0.f};
    bool active = idx < N;
    if (active) {
        gs = gs_i[idx]; u[0] = u_i[idx * 3]; u[1] = u_i[idx * 3 + 1]; u[2] = u_i[idx * 3 + 2];
        active = (gs != 0.f) || (u[0] != 0.f) || (u[1] != 0.f) || (u[2] != 0.f);
    }
    if (!__syncthreads_or(active)) return;
    PointTaps pt; float om[12]; uint64_t m1 = 0, m2 = 0;
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const float* geo = planes; float* ggeo = gplanes;
    if (active) {
        float x[3]; int prompt;
        point_of(src, idx, x, prompt);
        geo = planes + (size_t)prompt * 6 * ps; ggeo = gplanes + (size_t)prompt * 6 * ps;
        float p[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
        Taps tp[3];
        point_taps(p, cfg.R, pt, tp);
        const float sc = 0.5f * (float)cfg.R / cfg.radius;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float ixd = u[plane_ax(k)] * sc, iyd = u[plane_ay(k)] * sc;
            om[k * 4 + 0] = gs * tp[k].w[0] + (-tp[k].wy0 * ixd - tp[k].wx0 * iyd);
            om[k * 4 + 1] = gs * tp[k].w[1] + (tp[k].wy0 * ixd - tp[k].wx1 * iyd);
            om[k * 4 + 2] = gs * tp[k].w[2] + (-tp[k].wy1 * ixd + tp[k].wx0 * iyd);
            om[k * 4 + 3] = gs * tp[k].w[3] + (tp[k].wy1 * ixd + tp[k].wx1 * iyd);
        }
        gather<C, 3>(geo, ps, pt.o, pt.w, sE + tid);
        float acc[HID];
        zero64(acc); layer64_acc(acc, wp + wo.w1sT, C, sE + tid); m1 = store_relu64(acc, sB + tid);
        zero64(acc); layer64_acc(acc, wp + wo.w2sT, HID, sB + tid); m2 = mask64(acc);
#pragma unroll
        for (int j = 0; j < HID; ++j) sB[j * ST + tid] = ((m2 >> j) & 1ull) ? __ldg(wp + wo.w3s + j) : 0.f;
        zero64(acc); layerT_acc<HID>(acc, wp + wo.w2s, HID, sB + tid); store_masked64(acc, m1, sB + tid);   // a1
        float de[C];
#pragma unroll
        for (int c = 0; c < C; ++c) de[c] = 0.f;
        layerT_acc<C>(de, wp + wo.w1s, C, sB + tid);
#pragma unroll
        for (int c = 0; c < C; ++c) sD[c * ST + tid] = de[c];
        gather<C, 3>(geo, ps, pt.o, om, sE + tid);            // ẽ
    } else {
        zero_col(sE + tid, RE); zero_col(sB + tid, HID); zero_col(sT + tid, HID);
    }
    __syncthreads();
    if (gw) wgrad(gw + go.g1s, C, sB, HID, sE, C);            // dW1 += a1 ẽᵀ
    if (active) {
        float acc[HID];
        zero64(acc); layer64_acc(acc, wp + wo.w1sT, C, sE + tid); store_masked64(acc, m1, sT + tid);   // h̃1
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int j = 0; j < HID; ++j) sB[j * ST + tid] = ((m2 >> j) & 1ull) ? __ldg(wp + wo.w3s + j) : 0.f;   // a2
    }
    __syncthreads();
    if (gw) wgrad(gw + go.g2s, HID, sB, HID, sT, HID);        // dW2 += a2 h̃1ᵀ
    {
        float acc[HID];
        zero64(acc);
        if (active) layer64_acc(acc, wp + wo.w2sT, HID, sT + tid);
        __syncthreads();
        store_masked64(acc, m2, sB + tid);                     // h̃2 (zero for inactive threads: m2 = 0)
    }
    __syncthreads();
    if (gw && tid < HID) {                                     // dw3 += Σ_p h̃2
        float s = 0.f;
        for (int p = 0; p < TPB; ++p) s += sB[tid * ST + p];
        if (s != 0.f) atomicAdd(gw + go.g3s + tid, s);
    }
    if (active && gplanes) scatter_col<C, 3>(ggeo, ps, pt.o, om, sD + tid);
}

// =====================================================================================================
// backward, stage 2b: colour decoder + texture planes, one thread per sample point (seed gf[3])
// =====================================================================================================
template <int C>
__global__ void __launch_bounds__(TPB) k_bwd_tex(const float* __restrict__ planes, const float* __restrict__ wp,
                                                tt_config cfg, PtSrc src, int64_t N, const float* __restrict__ gf_i,
                                                float* __restrict__ gplanes, float* __restrict__ gw) {
    TT_SHARED(smem);
    float* sX = smem;                       // C rows: one texture plane's encoding
    float* sA = sX + C * ST;                // 64 rows: h1
    float* sB = sA + HID * ST;              // 64 rows: h2 -> g_h2 -> g_h1
    float* sG = sB + HID * ST;              // 4 rows: gf
    const int tid = threadIdx.x;
    const WOff wo = woff(C);
    const GOff go = goff(C);
    const int64_t idx = (int64_t)blockIdx.x * TPB + tid;
    float gf[3] = {0.f, 0.f, 0.f};
    bool active = idx < N;
    if (active) {
        gf[0] = gf_i[idx * 3]; gf[1] = gf_i[idx * 3 + 1]; gf[2] = gf_i[idx * 3 + 2];
        active = (gf[0] != 0.f) || (gf[1] != 0.f) || (gf[2] != 0.f);
    }
    if (!__syncthreads_or(active)) return;
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const float* tex = planes; float* gtex = gplanes;
    float p[3] = {0.f, 0.f, 0.f};
    uint64_t m1 = 0, m2 = 0;
    if (active) {
        float x[3]; int prompt;
        point_of(src, idx, x, prompt);
        tex = planes + ((size_t)prompt * 6 + 3) * ps; gtex = gplanes + ((size_t)prompt * 6 + 3) * ps;
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
        float acc[HID];
        zero64(acc);
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            const Taps t = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
            gather<C, 1>(tex + k * ps, 0, t.o, t.w, sX + tid);
            layer64_acc(acc, wp + wo.w1fT + k * C * HID, C, sX + tid);
        }
        m1 = store_relu64(acc, sA + tid);
        zero64(acc); layer64_acc(acc, wp + wo.w2fT, HID, sA + tid); m2 = store_relu64(acc, sB + tid);
    } else {
        zero_col(sA + tid, HID); zero_col(sB + tid, HID); zero_col(sX + tid, C);
    }
    sG[tid] = gf[0]; sG[ST + tid] = gf[1]; sG[2 * ST + tid] = gf[2]; sG[3 * ST + tid] = 0.f;
    __syncthreads();
    if (gw) wgrad(gw + go.g3f, HID, sG, 3, sB, HID);          // dW3 += gf h2ᵀ
    __syncthreads();
    if (active) {
#pragma unroll
        for (int j = 0; j < HID; ++j) {
            const float v = gf[0] * __ldg(wp + wo.w3f + j) + gf[1] * __ldg(wp + wo.w3f + HID + j) +
                            gf[2] * __ldg(wp + wo.w3f + 2 * HID + j);
            sB[j * ST + tid] = ((m2 >> j) & 1ull) ? v : 0.f;  // g_h2
        }
    }
    __syncthreads();
    if (gw) wgrad(gw + go.g2f, HID, sB, HID, sA, HID);        // dW2 += g_h2 h1ᵀ
    __syncthreads();
    if (active) {
        float acc[HID];
        zero64(acc); layerT_acc<HID>(acc, wp + wo.w2f, HID, sB + tid); store_masked64(acc, m1, sB + tid);   // g_h1
    }
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        Taps t;
        if (active) {
            t = make_taps(p[plane_ax(k)], p[plane_ay(k)], cfg.R);
            gather<C, 1>(tex + k * ps, 0, t.o, t.w, sX + tid);
        }
        __syncthreads();
        if (gw) wgrad(gw + go.g1f + k * C, 3 * C, sB, HID, sX, C);   // dW1[:, kC:(k+1)C] += g_h1 enc_kᵀ
        __syncthreads();
        if (active && gplanes) {
            float ge[C];
#pragma unroll
            for (int c = 0; c < C; ++c) ge[c] = 0.f;
            layerT_acc<C>(ge, wp + wo.w1f + k * C, 3 * C, sB + tid);
#pragma unroll
            for (int c = 0; c < C; ++c) sX[c * ST + tid] = ge[c];
            scatter_col<C, 1>(gtex + k * ps, 0, t.o, t.w, sX + tid);
        }
    }
}

// =====================================================================================================
// backward of the deformation decoder of forward_field (few_step…diffusion.py:375-394): same geometry encoding
// (three planes summed), head [3][64], seed gd[3] per point.  One thread per point, fp32.
// gwd: [64][C] | [64][64] | [3][64]  (nn.Linear layout, tt_wgrad_def_floats)
// =====================================================================================================
template <int C>
__global__ void __launch_bounds__(TPB) k_bwd_def(const float* __restrict__ planes, const float* __restrict__ wp,
                                                tt_config cfg, PtSrc src, int64_t N, const float* __restrict__ gd_i,
                                                float* __restrict__ gplanes, float* __restrict__ gwd) {
    TT_SHARED(smem);
    float* sX = smem;                       // C rows: geometry encoding -> W1dᵀ g_h1
    float* sA = sX + C * ST;                // 64 rows: h1
    float* sB = sA + HID * ST;              // 64 rows: h2 -> g_h2 -> g_h1
    float* sG = sB + HID * ST;              // 4 rows: gd
    const int tid = threadIdx.x;
    const WOff wo = woff(C);
    const int64_t idx = (int64_t)blockIdx.x * TPB + tid;
    float gd[3] = {0.f, 0.f, 0.f};
    bool active = idx < N;
    if (active) {
        gd[0] = gd_i[idx * 3]; gd[1] = gd_i[idx * 3 + 1]; gd[2] = gd_i[idx * 3 + 2];
        active = (gd[0] != 0.f) || (gd[1] != 0.f) || (gd[2] != 0.f);
    }
    if (!__syncthreads_or(active)) return;
    const size_t ps = (size_t)cfg.R * cfg.R * C;
    const float* geo = planes; float* ggeo = gplanes;
    PointTaps pt; uint64_t m1 = 0, m2 = 0;
    if (active) {
        float x[3], p[3]; int prompt;
        point_of(src, idx, x, prompt);
        geo = planes + (size_t)prompt * 6 * ps; ggeo = gplanes + (size_t)prompt * 6 * ps;
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], cfg.radius);
        Taps tp[3];
        point_taps(p, cfg.R, pt, tp);
        gather<C, 3>(geo, ps, pt.o, pt.w, sX + tid);
        float acc[HID];
        zero64(acc); layer64_acc(acc, wp + wo.w1dT, C, sX + tid); m1 = store_relu64(acc, sA + tid);
        zero64(acc); layer64_acc(acc, wp + wo.w2dT, HID, sA + tid); m2 = store_relu64(acc, sB + tid);
    } else {
        zero_col(sX + tid, C); zero_col(sA + tid, HID); zero_col(sB + tid, HID);
    }
    sG[tid] = gd[0]; sG[ST + tid] = gd[1]; sG[2 * ST + tid] = gd[2]; sG[3 * ST + tid] = 0.f;
    __syncthreads();
    if (gwd) wgrad(gwd + 64 * C + 4096, HID, sG, 3, sB, HID);          // dW3 += gd h2ᵀ
    __syncthreads();
    if (active) {
#pragma unroll
        for (int j = 0; j < HID; ++j) {
            const float v = gd[0] * __ldg(wp + wo.w3d + j) + gd[1] * __ldg(wp + wo.w3d + HID + j) +
                            gd[2] * __ldg(wp + wo.w3d + 2 * HID + j);
            sB[j * ST + tid] = ((m2 >> j) & 1ull) ? v : 0.f;           // g_h2
        }
    }
    __syncthreads();
    if (gwd) wgrad(gwd + 64 * C, HID, sB, HID, sA, HID);               // dW2 += g_h2 h1ᵀ
    __syncthreads();
    if (active) {
        float acc[HID];
        zero64(acc); layerT_acc<HID>(acc, wp + wo.w2d, HID, sB + tid); store_masked64(acc, m1, sB + tid);   // g_h1
    }
    __syncthreads();
    if (gwd) wgrad(gwd, C, sB, HID, sX, C);                            // dW1 += g_h1 eᵀ
    __syncthreads();
    if (active && gplanes) {
        float ge[C];
#pragma unroll
        for (int c = 0; c < C; ++c) ge[c] = 0.f;
        layerT_acc<C>(ge, wp + wo.w1d, C, sB + tid);
#pragma unroll
        for (int c = 0; c < C; ++c) sX[c * ST + tid] = ge[c];
        scatter_col<C, 3>(ggeo, ps, pt.o, pt.w, sX + tid);
    }
}

// =====================================================================================================
// stand-alone compositor (nerfacc.render_weight_from_alpha + accumulate_along_rays, dense rays)
// =====================================================================================================
__global__ void k_composite_fwd(const float* __restrict__ alphas, const float* __restrict__ values, int64_t n_rays,
                                int S, int D, float* weights, float* trans, float* out) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    float T = 1.f, acc[8];
    const int Do = D > 0 ? D : 1;
    for (int k = 0; k < Do; ++k) acc[k] = 0.f;
    for (int i = 0; i < S; ++i) {
        const int64_t si = ray * S + i;
        const float a = alphas[si], w = T * a;
        if (weights) weights[si] = w;
        if (trans) trans[si] = T;
        if (D > 0) { for (int k = 0; k < D; ++k) acc[k] = fmaf(w, values[si * D + k], acc[k]); }
        else acc[0] += w;
        T *= (1.f - a);
    }
    if (out) for (int k = 0; k < Do; ++k) out[ray * Do + k] = acc[k];
}
__global__ void k_composite_bwd(const float* __restrict__ alphas, const float* __restrict__ values,
                                const float* __restrict__ trans, const float* __restrict__ g_out,
                                const float* __restrict__ g_weights, int64_t n_rays, int S, int D, float* g_alphas,
                                float* g_values) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const int Do = D > 0 ? D : 1;
    float go[8];
    for (int k = 0; k < Do; ++k) go[k] = g_out ? g_out[ray * Do + k] : 0.f;
    float Rh = 0.f;
    for (int i = S - 1; i >= 0; --i) {
        const int64_t si = ray * S + i;
        const float a = alphas[si], T = trans[si], w = T * a;
        float gw = g_weights ? g_weights[si] : 0.f;
        if (D > 0) {
            for (int k = 0; k < D; ++k) {
                gw = fmaf(go[k], values[si * D + k], gw);
                if (g_values) g_values[si * D + k] = w * go[k];
            }
        } else gw += go[0];
        g_alphas[si] = T * (gw - Rh);
        Rh = gw * a + (1.f - a) * Rh;
    }
}

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int tt_version(void) { return TT_VERSION; }
#if defined(TT_WS_TIMING) && !defined(TT_EMUL)
// debug builds only: phase anatomy of k_geo_ws (cycles accumulated by CTA 0), reset after reading
extern "C" int tt_debug_ws_prof(unsigned long long* out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, tt::g_ws_prof, sizeof(unsigned long long) * 32);
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(tt::g_ws_prof, z, sizeof(z));
    return 0;
}
#endif
const char* tt_last_error(void) { return g_err; }
int64_t tt_launch_count(void) { return g_launches.load(); }
int tt_profile_begin(void) {
#ifndef TT_EMUL
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = true;
#endif
    return TT_OK;
}
int tt_profile_end(char* buf, size_t cap) {
    if (buf && cap) buf[0] = 0;
#ifndef TT_EMUL
    g_prof_on = false;
    size_t used = 0;
    for (auto& r : g_prof) {
        float ms = 0.f;
        cudaEventSynchronize(r.b);
        cudaEventElapsedTime(&ms, r.a, r.b);
        if (buf) {
            const int n = snprintf(buf + used, used < cap ? cap - used : 0, "%s:%.6f;", r.name.c_str(), ms);
            if (n > 0 && used + (size_t)n < cap) used += (size_t)n;
        }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
#endif
    return TT_OK;
}
int tt_set_option(const char* name, int value) {
    if (!name) return fail(TT_E_ARG, "tt_set_option: null name%s", "");
    if (!strcmp(name, "scatter")) { if (value < -1 || value > 2) return fail(TT_E_ARG, "tt_set_option: scatter must be -1..2%s", ""); g_opt_scatter = value; return TT_OK; }
    if (!strcmp(name, "patch_lists")) { g_opt_patch = value != 0; return TT_OK; }
    if (!strcmp(name, "grid_lines")) { g_opt_grid_lines = value != 0; return TT_OK; }
    return fail(TT_E_ARG, "tt_set_option: unknown option '%s'", name);
}
int tt_set_impl(int impl) {
    if (impl < 0 || impl > 2) return fail(TT_E_ARG, "tt_set_impl: impl must be 0 (SIMT), 1 (round-1 tcgen05) or 2 (warp-specialised tcgen05)%s", "");
    g_impl = impl;
    return TT_OK;
}
int tt_get_impl(void) { return g_impl; }
int tt_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) { cudaGetLastError(); return 0; }
    return 1;
}
size_t tt_wpack_floats(int C) { return supported_C(C) ? (size_t)woff(C).total : 0; }
size_t tt_wgrad_floats(int C) { return supported_C(C) ? (size_t)goff(C).total : 0; }
int tt_wgrad_offsets(int C, int64_t off[6]) {
    if (!supported_C(C) || !off) return fail(TT_E_ARG, "tt_wgrad_offsets: bad arguments%s", "");
    const GOff g = goff(C);
    off[0] = g.g1s; off[1] = g.g2s; off[2] = g.g3s; off[3] = g.g1f; off[4] = g.g2f; off[5] = g.g3f;
    return TT_OK;
}

int tt_pack_weights(const float* s0, const float* s1, const float* s2, const float* f0, const float* f1,
                    const float* f2, const float* d0, const float* d1, const float* d2, int C, float* wpack,
                    void* stream) {
    if (!supported_C(C)) return fail(TT_E_ARG, "unsupported channel count%s %lld (8,16,32,40,64)", "", C);
    if (!s0 || !s1 || !s2 || !wpack) return fail(TT_E_ARG, "tt_pack_weights: sdf weights and wpack are required%s", "");
    if ((f0 || f1 || f2) && !(f0 && f1 && f2)) return fail(TT_E_ARG, "tt_pack_weights: feature weights must be all set or all NULL%s", "");
    if ((d0 || d1 || d2) && !(d0 && d1 && d2)) return fail(TT_E_ARG, "tt_pack_weights: deformation weights must be all set or all NULL%s", "");
    if (!aligned16(wpack)) return fail(TT_E_ALIGN, "wpack must be 16-byte aligned%s", "");
    TT_LAUNCH(k_pack_weights, 32, 256, 0, (cudaStream_t)stream, s0, s1, s2, f0, f1, f2, d0, d1, d2, C, wpack);
    return check_launch("tt_pack_weights");
}

int tt_repack_planes(const float* src, int P, int Csrc, int off_geo, int off_tex, int C, int R, float* dst, void* stream) {
    if (!src || !dst) return fail(TT_E_ARG, "tt_repack_planes: NULL pointer%s", "");
    if (P < 1 || R < 1 || C < 1 || off_geo < 0 || off_tex < 0 || off_geo + C > Csrc || off_tex + C > Csrc)
        return fail(TT_E_ARG, "tt_repack_planes: bad shape%s", "");
    if (C > 256) return fail(TT_E_ARG, "tt_repack_planes: C too large%s (%lld)", "", C);
    dim3 grid((R + 31) / 32, R, P * 6), block(32, 8);
    TT_LAUNCH(k_repack<false>, grid, block, (size_t)C * 33 * 4, (cudaStream_t)stream, src, dst, Csrc, off_geo, off_tex, C, R);
    return check_launch("tt_repack_planes");
}
int tt_repack_planes_bwd(const float* gplanes, int P, int C, int R, float* gsrc, void* stream) {
    if (!gplanes || !gsrc) return fail(TT_E_ARG, "tt_repack_planes_bwd: NULL pointer%s", "");
    if (P < 1 || R < 1 || C < 1 || C > 256) return fail(TT_E_ARG, "tt_repack_planes_bwd: bad shape%s", "");
    dim3 grid((R + 31) / 32, R, P * 6), block(32, 8);
    TT_LAUNCH(k_repack<true>, grid, block, (size_t)C * 33 * 4, (cudaStream_t)stream, gplanes, gsrc, C, 0, 0, C, R);
    return check_launch("tt_repack_planes_bwd");
}

int tt_repack_planes_bwd_split(const float* gplanes, int P, int Cdst, int off_geo, int off_tex, int C, int R, float* gsrc,
                               void* stream) {
    if (!gplanes || !gsrc) return fail(TT_E_ARG, "tt_repack_planes_bwd_split: NULL pointer%s", "");
    if (P < 1 || R < 1 || C < 1 || C > 256 || off_geo < 0 || off_tex < 0 || off_geo + C > Cdst || off_tex + C > Cdst)
        return fail(TT_E_ARG, "tt_repack_planes_bwd_split: bad shape%s", "");
    dim3 grid((R + 31) / 32, R, P * 6), block(32, 8);
    TT_LAUNCH(k_repack<true>, grid, block, (size_t)C * 33 * 4, (cudaStream_t)stream, gplanes, gsrc, Cdst, off_geo, off_tex, C, R);
    return check_launch("tt_repack_planes_bwd_split");
}

int tt_geometry_fwd(const float* planes, const float* wpack, const tt_config* cfg, const float* points, int64_t M,
                    int grid_res, float* sdf, float* sdf_orig, float* features, float* normal, float* sdf_grad,
                    float* deformation, void* stream) {
    if (int e = check_cfg(cfg)) return e;
    if (!planes || !wpack) return fail(TT_E_ARG, "tt_geometry_fwd: planes/wpack NULL%s", "");
    if (!points) {
        if (grid_res < 2) return fail(TT_E_ARG, "tt_geometry_fwd: points NULL needs grid_res >= 2%s", "");
        M = (int64_t)grid_res * grid_res * grid_res;
    }
    if (M < 0) return fail(TT_E_ARG, "tt_geometry_fwd: M < 0%s", "");
    if (!aligned16(planes) || !aligned16(wpack)) return fail(TT_E_ALIGN, "planes/wpack must be 16-byte aligned%s", "");
    const int64_t N = (int64_t)cfg->P * M;
    if (N == 0) return TT_OK;
    if (g_impl >= 1 && !(deformation && (normal || sdf_grad)) && N < 2147483647LL) {
        bool done = false;
        TT_DISPATCH_C(cfg->C, {
            const size_t smg_n = (size_t)GeoSmem<kC, true>::TOTAL * 4, smg = (size_t)GeoSmem<kC, false>::TOTAL * 4;
            const size_t smg_d = (size_t)GeoSmem<kC, false, true>::TOTAL * 4;
            const size_t smt = (size_t)TexSmem<kC>::TOTAL * 4;
            const bool want_n = normal || sdf_grad;
            if ((deformation ? smg_d : want_n ? smg_n : smg) <= kMaxSmem && smt <= kMaxSmem) {
                TcSrc src{}; src.mode = points ? 0 : 3; src.points = points; src.M = M; src.grid_res = grid_res; src.grid_lines = g_opt_grid_lines;
                if (sdf || sdf_orig || want_n || deformation) {
                    if (deformation) {      // field query of the mesh paths: SDF + deformation decoders on one gather
                        const size_t smw_d = (size_t)GeoWs<kC, false, true>::TOTAL * 4;
                        if (g_impl == 2 && smw_d <= kMaxSmem) {
                            if (int e = set_smem((k_geo_ws<kC, false, true>), smw_d)) return e;
                            const int64_t ctas = (N + WS_CG * TC_GROUP - 1) / (WS_CG * TC_GROUP);
                            const unsigned grid = (unsigned)(ctas < (int64_t)num_sms() ? (ctas < 1 ? 1 : ctas) : num_sms());
                            TT_LAUNCH((k_geo_ws<kC, false, true>), grid, WS_GEO_THREADS, smw_d, (cudaStream_t)stream, planes, wpack, *cfg, src, N, sdf, sdf_orig, (float*)nullptr, (float*)nullptr, (uint64_t*)nullptr, (float*)nullptr, deformation);
                            if (int e = check_launch("k_geo_ws")) return e;
                        } else {
                        if (int e = set_smem(k_geo_tc<kC, false, true>, smg_d)) return e;
                        TT_LAUNCH((k_geo_tc<kC, false, true>), tc_grid(N), TC_THREADS, smg_d, (cudaStream_t)stream, planes, wpack, *cfg, src, N, sdf, sdf_orig, sdf_grad, normal, (uint64_t*)nullptr, deformation);
                        if (int e = check_launch("k_geo_tc")) return e;
                        }
                    } else if (want_n) {
                        if (int e = launch_geo_decoder<kC, true>(planes, wpack, cfg, src, N, sdf, sdf_orig, sdf_grad, normal, nullptr, nullptr, (cudaStream_t)stream)) return e;
                    } else {
                        if (int e = launch_geo_decoder<kC, false>(planes, wpack, cfg, src, N, sdf, sdf_orig, sdf_grad, normal, nullptr, nullptr, (cudaStream_t)stream)) return e;
                    }
                }
                if (features) {
                    if (int e = launch_tex_decoder<kC>(planes, wpack, cfg, src, N, features, nullptr, (cudaStream_t)stream)) return e;
                }
                done = true;
            }
        });
        if (done) return TT_OK;
    }
    const int64_t blocks = (N + TPB - 1) / TPB;
    if (blocks > 2147483647LL) return fail(TT_E_ARG, "tt_geometry_fwd: too many points%s (%lld)", "", N);
    TT_DISPATCH_C(cfg->C, {
        const size_t sm = slab_bytes(imax(kC, HID) + HID);
        if (int e = set_smem(k_geometry_fwd<kC>, sm)) return e;
        TT_LAUNCH(k_geometry_fwd<kC>, (unsigned)blocks, TPB, sm, (cudaStream_t)stream, planes, wpack, *cfg, points, M, grid_res,
            sdf, sdf_orig, features, normal, sdf_grad, deformation);
    });
    return check_launch("tt_geometry_fwd");
}

size_t tt_sample_scratch_floats(int64_t n_rays, int n_imp) { return 3 * (size_t)n_rays * (size_t)(n_imp + 1) + 16; }

int tt_importance_sample(const float* planes, const float* wpack, const tt_config* cfg, const float* rays_o,
                         const float* rays_d, int64_t n_rays, int n_imp, int n_fine, const float* jitter0,
                         const float* jitter1, float* scratch, float* t_vals, void* stream) {
    if (int e = check_cfg(cfg)) return e;
    if (!planes || !wpack || !rays_o || !rays_d || !scratch || !t_vals) return fail(TT_E_ARG, "tt_importance_sample: NULL pointer%s", "");
    if (n_imp < 1 || n_fine < 1) return fail(TT_E_ARG, "tt_importance_sample: n_imp and n_fine must be >= 1%s", "");
    if ((jitter0 == nullptr) != (jitter1 == nullptr)) return fail(TT_E_ARG, "tt_importance_sample: both jitters or none%s", "");
    if (cfg->rays_per_cache < 1 || n_rays > (int64_t)cfg->P * cfg->rays_per_cache)
        return fail(TT_E_ARG, "tt_importance_sample: n_rays exceeds P*rays_per_cache%s (%lld)", "", n_rays);
    if (!aligned16(planes) || !aligned16(wpack)) return fail(TT_E_ALIGN, "planes/wpack must be 16-byte aligned%s", "");
    if (n_rays <= 0) return TT_OK;
    const int64_t blocks = (n_rays + TPB - 1) / TPB;
    if (g_impl >= 1 && n_rays * n_imp < 2147483647LL) {
        bool done = false;
        TT_DISPATCH_C(cfg->C, {
            const size_t smg = (size_t)GeoSmem<kC, false>::TOTAL * 4;
            if (smg <= kMaxSmem) {
                float* cdf = scratch;
                float* sdf = scratch + (size_t)n_rays * (size_t)(n_imp + 1);
                TcSrc src{}; src.mode = 2; src.rs = RaySrcT{rays_o, rays_d, nullptr, nullptr, 0, 1};
                src.rays_per_cache = cfg->rays_per_cache; src.n_imp = n_imp; src.jitter0 = jitter0;
                src.near_plane = cfg->near_plane; src.far_plane = cfg->far_plane;
                const int64_t N = n_rays * n_imp;
                int* list = reinterpret_cast<int*>(scratch + 2 * (size_t)n_rays * (size_t)(n_imp + 1));
                int* lcount = list + N;
                if (cudaMemsetAsync(lcount, 0, sizeof(int), (cudaStream_t)stream) != cudaSuccess) return fail(TT_E_CUDA, "cudaMemsetAsync failed%s", "");
                TT_LAUNCH(k_classify, (unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream, *cfg, src, N, sdf, (float*)nullptr, (float*)nullptr, (float*)nullptr, list, lcount);
                if (int e = check_launch("k_classify")) return e;
                src.index = list; src.count = lcount;
                if (int e = launch_geo_decoder<kC, false>(planes, wpack, cfg, src, N, sdf, nullptr, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream)) return e;
                TT_LAUNCH(k_sampler_post, (unsigned)blocks, TPB, 0, (cudaStream_t)stream, *cfg, n_rays, n_imp, n_fine, (const float*)sdf, jitter0, jitter1, cdf, t_vals);
                if (int e = check_launch("k_sampler_post")) return e;
                done = true;
            }
        });
        if (done) return TT_OK;
    }
    TT_DISPATCH_C(cfg->C, {
        const size_t sm = slab_bytes(imax(kC, HID) + HID);
        if (int e = set_smem(k_importance_sample<kC>, sm)) return e;
        TT_LAUNCH(k_importance_sample<kC>, (unsigned)blocks, TPB, sm, (cudaStream_t)stream, planes, wpack, *cfg, rays_o, rays_d,
            n_rays, n_imp, n_fine, jitter0, jitter1, scratch, t_vals);
    });
    return check_launch("tt_importance_sample");
}

static int check_rays(const char* who, const tt_config* cfg, const float* rays_o, const float* rays_d, int64_t n_rays,
                      const float* t0, const float* t1, int64_t t_stride, int S) {
    if (!rays_o || !rays_d || !t0 || !t1) return fail(TT_E_ARG, "%s: NULL ray/interval pointer", who);
    if (S < 1 || t_stride < S) return fail(TT_E_ARG, "%s: bad S / t_stride", who);
    if (cfg->rays_per_cache < 1 || n_rays > (int64_t)cfg->P * cfg->rays_per_cache)
        return fail(TT_E_ARG, "%s: n_rays exceeds P*rays_per_cache (%lld)", who, n_rays);
    return TT_OK;
}

size_t tt_render_fwd_scratch_floats(int64_t n_rays, int S) { return round4((size_t)n_rays * (size_t)S * 9 + 16) + ws_scratch_floats(); }

int tt_render_fwd(const float* planes, const float* wpack, const tt_config* cfg, const float* rays_o,
                  const float* rays_d, int64_t n_rays, const float* t_starts, const float* t_ends, int64_t t_stride,
                  int S, float* acc, float* sdf, float* sdf_orig, float* sdf_grad, float* normal, float* features,
                  float* weights, float* trans, uint64_t* masks, float* scratch, void* stream) {
    if (int e = check_cfg(cfg)) return e;
    if (!planes || !wpack || !acc) return fail(TT_E_ARG, "tt_render_fwd: NULL pointer%s", "");
    if (int e = check_rays("tt_render_fwd", cfg, rays_o, rays_d, n_rays, t_starts, t_ends, t_stride, S)) return e;
    if (!aligned16(planes) || !aligned16(wpack)) return fail(TT_E_ALIGN, "planes/wpack must be 16-byte aligned%s", "");
    if (n_rays <= 0) return TT_OK;
    const RaySrc rs{rays_o, rays_d, t_starts, t_ends, t_stride, S};
    const int64_t blocks = (n_rays + TPB - 1) / TPB;
    if (g_impl >= 1 && scratch && n_rays * S < 2147483647LL) {
        bool done = false;
        TT_DISPATCH_C(cfg->C, {
            const size_t smg = (size_t)GeoSmem<kC, true>::TOTAL * 4, smt = (size_t)TexSmem<kC>::TOTAL * 4;
            if (smg <= kMaxSmem && smt <= kMaxSmem) {
                cudaStream_t st = (cudaStream_t)stream;
                const int64_t N = n_rays * S;
                float* p_sdf = sdf ? sdf : scratch;
                float* p_grad = sdf_grad ? sdf_grad : scratch + N;
                float* p_trans = trans ? trans : scratch + 4 * N;
                float* p_feat = features ? features : scratch + 5 * N;
                int* live = reinterpret_cast<int*>(scratch + 8 * N);
                int* count = live + N;
                const int all_live = (cfg->flags & TT_FLAG_ALL_FEATURES) ? 1 : 0;
                const RaySrcT rt{rays_o, rays_d, t_starts, t_ends, t_stride, S};
                TcSrc src{}; src.mode = 1; src.rs = rt; src.rays_per_cache = cfg->rays_per_cache;
                if (cudaMemsetAsync(count, 0, 2 * sizeof(int), st) != cudaSuccess) return fail(TT_E_CUDA, "cudaMemsetAsync failed%s", "");
                // empty points (every tap out of bounds) get their closed-form outputs here; the decoder runs on the rest.
                // The list shares the live-sample list's storage (it is dead once k_geo_tc has run).
                TT_LAUNCH(k_classify, (unsigned)((N + 255) / 256), 256, 0, st, *cfg, src, N, p_sdf, sdf_orig, p_grad, (float*)nullptr, live, count + 1);
                if (int e = check_launch("k_classify")) return e;
                src.index = live; src.count = count + 1;
                if (int e = launch_geo_decoder<kC, true>(planes, wpack, cfg, src, N, p_sdf, sdf_orig, p_grad, nullptr, masks, scratch + round4((size_t)N * 9 + 16), st)) return e;
                src.index = nullptr; src.count = nullptr;
                TT_LAUNCH(k_weights, (unsigned)((n_rays + RAY_WARPS - 1) / RAY_WARPS), RAY_WARPS * 32, 0, st, *cfg, rt, n_rays, (const float*)p_sdf, (const float*)p_grad, acc, weights, p_trans, normal,
                          all_live ? (float*)nullptr : p_feat, all_live ? (int*)nullptr : live, count, all_live);
                if (int e = check_launch("k_weights")) return e;
                if (!all_live) { src.index = live; src.count = count; }
                if (int e = launch_tex_decoder<kC>(planes, wpack, cfg, src, N, p_feat, masks, st)) return e;
                TT_LAUNCH(k_accum_rgb, (unsigned)((n_rays + RAY_WARPS - 1) / RAY_WARPS), RAY_WARPS * 32, 0, st, *cfg, rt, n_rays, (const float*)p_sdf, (const float*)p_grad, (const float*)p_trans, (const float*)p_feat, acc);
                if (int e = check_launch("k_accum_rgb")) return e;
                done = true;
            }
        });
        if (done) return TT_OK;
    }
    TT_DISPATCH_C(cfg->C, {
        const size_t sm = slab_bytes(imax(kC, HID) + HID);
        if (int e = set_smem(k_render_fwd<kC>, sm)) return e;
        TT_LAUNCH(k_render_fwd<kC>, (unsigned)blocks, TPB, sm, (cudaStream_t)stream, planes, wpack, *cfg, rs, n_rays, acc, sdf,
            sdf_orig, sdf_grad, normal, features, weights, trans);
    });
    return check_launch("tt_render_fwd");
}

static size_t hid_floats(const tt_config* cfg) {      // hidden-gradient planes of the colour backward, [P][3][R*R][64]
    return cfg ? (size_t)cfg->P * 3 * (size_t)cfg->R * (size_t)cfg->R * 64 : 0;
}

static int launch_point_bwd(const float* planes, const float* wpack, const tt_config* cfg, const PtSrc& src, int64_t N,
                            const float* gs, const float* u, const float* gf, const uint64_t* tex_masks, float* gplanes,
                            float* gw, float* hid, cudaStream_t st, const int* geo_list = nullptr,
                            const int* geo_count = nullptr, const int* tex_list = nullptr, const int* tex_count = nullptr,
                            bool patch_ordered = false) {
    const int64_t blocks = (N + TPB - 1) / TPB;
    if (blocks > 2147483647LL) return fail(TT_E_ARG, "too many sample points%s (%lld)", "", N);
    if (g_impl >= 1 && N < 2147483647LL) {
        bool done = false;
        TT_DISPATCH_C(cfg->C, {
            {
                const size_t smg = (size_t)BwdGeoSmem<kC>::TOTAL * 4, smt = (size_t)BwdTexSmem<kC>::TOTAL * 4;
                if (tex_masks && hid && smg <= kMaxSmem && smt <= kMaxSmem) {
                    TcSrc ts{};
                    if (src.points) { ts.mode = 0; ts.points = src.points; ts.M = src.M; }
                    else {
                        ts.mode = 1; ts.rs = RaySrcT{src.rs.rays_o, src.rs.rays_d, src.rs.t_starts, src.rs.t_ends, src.rs.t_stride, src.rs.S};
                        ts.rays_per_cache = src.rays_per_cache;
                    }
                    const int64_t tiles = (N + TC_GROUP - 1) / TC_GROUP;
                    const unsigned grid = (unsigned)(tiles < (int64_t)num_sms() ? tiles : num_sms());
                    constexpr int GG = BwdGeoSmem<kC>::G;
                    const int64_t ctas_g = (tiles + GG - 1) / GG;
                    const unsigned grid_g = (unsigned)(ctas_g < (int64_t)num_sms() ? ctas_g : num_sms());
                    if (int e = set_smem(k_bwd_geo_tc<kC>, smg)) return e;
                    ts.index = geo_list; ts.count = geo_count;
                    TT_LAUNCH(k_bwd_geo_tc<kC>, grid_g, GG * TC_GROUP, smg, st, planes, wpack, *cfg, ts, N, gs, u, tex_masks, gplanes, gw);
                    if (int e = check_launch("k_bwd_geo_tc")) return e;
                    // colour branch: per-sample kernel scatters the 64-wide hidden gradient, then two dense products
                    if (cudaMemsetAsync(hid, 0, hid_floats(cfg) * sizeof(float), st) != cudaSuccess) return fail(TT_E_CUDA, "cudaMemsetAsync failed%s", "");
                    ts.index = tex_list; ts.count = tex_count;
                    // scatter of the hidden gradient: tile-merged when the lists are patch-ordered (a texel then receives ~6 of a
                    // tile's taps: DESIGN 3.5 item 10), else run-length merged for long rays (consecutive samples share texel
                    // cells) or plain; tt_set_option("scatter", 0|1|2) forces a variant
                    int sc = scatter_mode();
                    if (sc < 0) sc = patch_ordered ? 2 : ((!src.points && src.rs.S >= 256) ? 1 : 0);
                    if ((int64_t)cfg->P * 3 * cfg->R * cfg->R >= (1LL << MergeWs::KEY_BITS) && sc == 2) sc = 0;       // texel numbers must fit the merge keys
                    const bool p3 = (cfg->flags & TT_FLAG_PRECISE_BWD) != 0 && (size_t)BwdTexSmem<kC, true>::TOTAL * 4 <= kMaxSmem;
                    int e = 0;
                    if (p3) e = sc == 2 ? launch_bwd_tex<kC, 2, true>(planes, wpack, cfg, ts, N, tiles, gf, tex_masks, hid, gw, st)
                              : sc == 1 ? launch_bwd_tex<kC, 1, true>(planes, wpack, cfg, ts, N, tiles, gf, tex_masks, hid, gw, st)
                                        : launch_bwd_tex<kC, 0, true>(planes, wpack, cfg, ts, N, tiles, gf, tex_masks, hid, gw, st);
                    else e = sc == 2 ? launch_bwd_tex<kC, 2, false>(planes, wpack, cfg, ts, N, tiles, gf, tex_masks, hid, gw, st)
                           : sc == 1 ? launch_bwd_tex<kC, 1, false>(planes, wpack, cfg, ts, N, tiles, gf, tex_masks, hid, gw, st)
                                     : launch_bwd_tex<kC, 0, false>(planes, wpack, cfg, ts, N, tiles, gf, tex_masks, hid, gw, st);
                    if (e) return e;
                    if (gplanes) {
                        const size_t smh = (size_t)64 * kC * 4;
                        const int64_t items = (int64_t)cfg->P * cfg->R * cfg->R * (kC / 4);
                        const int64_t nb = (items + 255) / 256;
                        TT_LAUNCH(k_hid_planes<kC>, dim3((unsigned)(nb < 8 * num_sms() ? nb : 8 * num_sms()), 3), 256, smh, st, (const float*)hid, wpack, cfg->P, cfg->R, gplanes);
                        if (int e = check_launch("k_hid_planes")) return e;
                    }
                    if (gw) {
                        const size_t smw = (size_t)(64 * 65 + 64 * (kC + 1)) * 4;
                        const int64_t slabs = (int64_t)cfg->P * (((int64_t)cfg->R * cfg->R + 63) / 64);
                        TT_LAUNCH(k_hid_wgrad<kC>, dim3((unsigned)(slabs < num_sms() ? slabs : num_sms()), 3), 256, smw, st, (const float*)hid, planes, cfg->P, cfg->R, gw);
                        if (int e = check_launch("k_hid_wgrad")) return e;
                    }
                    (void)grid;
                    done = true;
                }
            }
        });
        if (done) return TT_OK;
    }
    TT_DISPATCH_C(cfg->C, {
        const size_t smg = slab_bytes(kC + HID + kC + HID);
        if (int e = set_smem(k_bwd_geo<kC>, smg)) return e;
        TT_LAUNCH(k_bwd_geo<kC>, (unsigned)blocks, TPB, smg, st, planes, wpack, *cfg, src, N, gs, u, gplanes, gw);
        if (int e = check_launch("k_bwd_geo")) return e;
        const size_t smt = slab_bytes(kC + HID + HID + 4);
        if (int e = set_smem(k_bwd_tex<kC>, smt)) return e;
        TT_LAUNCH(k_bwd_tex<kC>, (unsigned)blocks, TPB, smt, st, planes, wpack, *cfg, src, N, gf, gplanes, gw);
        if (int e = check_launch("k_bwd_tex")) return e;
    });
    return TT_OK;
}

size_t tt_render_bwd_scratch_floats(const tt_config* cfg, int64_t n_rays, int S) {
    // seeds, lists, counters, flags | hidden-gradient planes
    return round4((size_t)n_rays * (size_t)S * 9 + 16 + (size_t)n_rays * 2 * (size_t)ray_chunks(S)) + hid_floats(cfg);
}

int tt_render_bwd(const float* planes, const float* wpack, const tt_config* cfg, const float* rays_o,
                  const float* rays_d, int64_t n_rays, const float* t_starts, const float* t_ends, int64_t t_stride,
                  int S, const float* acc, const float* sdf, const float* sdf_grad, const float* features,
                  const float* trans, const uint64_t* masks, const float* g_acc, const float* g_sdf,
                  const float* g_sdf_grad, const float* g_normal, const float* g_features, const float* g_weights,
                  float rgb_grad_scale, float* scratch, float* gplanes, float* gw, float* g_inv_std, void* stream) {
    if (int e = check_cfg(cfg)) return e;
    if (!planes || !wpack || !acc || !sdf || !sdf_grad || !features || !trans || !g_acc || !scratch)
        return fail(TT_E_ARG, "tt_render_bwd: NULL pointer%s", "");
    if (int e = check_rays("tt_render_bwd", cfg, rays_o, rays_d, n_rays, t_starts, t_ends, t_stride, S)) return e;
    if (!aligned16(planes) || !aligned16(wpack) || (gplanes && !aligned16(gplanes)))
        return fail(TT_E_ALIGN, "planes/wpack/gplanes must be 16-byte aligned%s", "");
    if (n_rays <= 0) return TT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const RaySrc rs{rays_o, rays_d, t_starts, t_ends, t_stride, S};
    const RaySrcT rst{rays_o, rays_d, t_starts, t_ends, t_stride, S};
    const int64_t N = n_rays * S;
    float* gs = scratch; float* u = scratch + N; float* gf = scratch + 4 * N;
    // tensor-core family: compacted lists of the samples that can contribute (non-empty point, non-zero seed)
    int* geo_list = nullptr; int* tex_list = nullptr; int* counts = nullptr;
    bool fwd_tc = false;     // did tt_render_fwd take the tensor-core path (and write the ReLU masks)?  Same test as there.
    TT_DISPATCH_C(cfg->C, { fwd_tc = (size_t)GeoSmem<kC, true>::TOTAL * 4 <= kMaxSmem && (size_t)TexSmem<kC>::TOTAL * 4 <= kMaxSmem; });
    if (g_impl >= 1 && fwd_tc && masks && N < 2147483647LL && (gplanes || gw)) {
        geo_list = reinterpret_cast<int*>(scratch + 7 * N); tex_list = geo_list + N; counts = tex_list + N;
        if (cudaMemsetAsync(counts, 0, 2 * sizeof(int), st) != cudaSuccess) return fail(TT_E_CUDA, "cudaMemsetAsync failed%s", "");
    }
    // rays given as [B][H][W] images: patch-ordered lists (k_patch_lists) so that the scatter kernels can merge a tile's taps
    const bool patch_lists = geo_list && cfg->image_h > 0 && cfg->image_w > 0 && cfg->image_h % PATCH == 0 && cfg->image_w % PATCH == 0 &&
                             n_rays % ((int64_t)cfg->image_h * cfg->image_w) == 0 && ray_chunks(S) <= PATCH_MAX_CHUNKS && patch_mode();
    TT_LAUNCH(k_render_bwd_comp, (unsigned)((n_rays + RAY_WARPS - 1) / RAY_WARPS), RAY_WARPS * 32, 0, st, *cfg, rst, n_rays, acc, sdf, sdf_grad, features,
        trans, g_acc, g_sdf, g_sdf_grad, g_normal, g_features, g_weights, rgb_grad_scale, gs, u, gf, g_inv_std,
        geo_list, counts, tex_list, counts ? counts + 1 : (int*)nullptr, counts ? reinterpret_cast<uint32_t*>(counts + 16) : (uint32_t*)nullptr,
        patch_lists ? 0 : 1);
    if (int e = check_launch("k_render_bwd_comp")) return e;
    if (patch_lists) {
        const size_t smp = (size_t)(PATCH_RAYS * 2 * PATCH_MAX_CHUNKS + 16) * 4;
        TT_LAUNCH(k_patch_lists, (unsigned)(n_rays / PATCH_RAYS), PATCH_THREADS, smp, st, (const uint32_t*)reinterpret_cast<uint32_t*>(counts + 16),
                  (int)cfg->image_h, (int)cfg->image_w, S, geo_list, counts, tex_list, counts + 1);
        if (int e = check_launch("k_patch_lists")) return e;
    }
    if (!gplanes && !gw) return TT_OK;
    PtSrc src; src.points = nullptr; src.M = 0; src.rs = rs; src.rays_per_cache = cfg->rays_per_cache;
    float* hid = scratch + round4((size_t)N * 9 + 16 + (size_t)n_rays * 2 * (size_t)ray_chunks(S));
    if (!geo_list) masks = nullptr;      // the forward ran the SIMT kernels (no masks were written): SIMT backward
    return launch_point_bwd(planes, wpack, cfg, src, N, gs, u, gf, masks, gplanes, gw, hid, st, geo_list, counts, tex_list,
                            counts ? counts + 1 : (const int*)nullptr, patch_lists);
}

size_t tt_geometry_bwd_scratch_floats(const tt_config* cfg, int64_t n_points) { return round4((size_t)n_points * 18 + 16) + hid_floats(cfg); }

int tt_geometry_bwd(const float* planes, const float* wpack, const tt_config* cfg, const float* points, int64_t M,
                       const float* g_sdf, const float* g_features, const float* g_normal, const float* g_sdf_grad,
                       float* scratch, float* gplanes, float* gw, void* stream) {
    if (int e = check_cfg(cfg)) return e;
    if (!planes || !wpack || !points || !scratch) return fail(TT_E_ARG, "tt_geometry_bwd: NULL pointer%s", "");
    if (!aligned16(planes) || !aligned16(wpack) || (gplanes && !aligned16(gplanes)))
        return fail(TT_E_ALIGN, "planes/wpack/gplanes must be 16-byte aligned%s", "");
    const int64_t N = (int64_t)cfg->P * M;
    if (N <= 0) return TT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float* gs = scratch; float* u = scratch + N; float* gf = scratch + 4 * N; float* grad = scratch + 7 * N;
    if (g_normal) {   // the normalisation Jacobian needs sdf_grad: recompute it
        if (int e = tt_geometry_fwd(planes, wpack, cfg, points, M, 0, nullptr, nullptr, nullptr, nullptr, grad, nullptr, stream)) return e;
    }
    TT_LAUNCH(k_geometry_bwd_seed, (unsigned)((N + 255) / 256), 256, 0, st, N, grad, g_sdf, g_features, g_normal, g_sdf_grad, gs, u, gf);
    if (int e = check_launch("k_geometry_bwd_seed")) return e;
    PtSrc src; src.points = points; src.M = M; src.rs = RaySrc{nullptr, nullptr, nullptr, nullptr, 0, 1}; src.rays_per_cache = 1;
    uint64_t* masks = nullptr;
    if (g_impl >= 1 && N < 2147483647LL) {      // forward ReLU masks of both decoders (tensor-core passes, 3xTF32)
        bool done = false;
        TT_DISPATCH_C(cfg->C, {
            const size_t smt = (size_t)TexSmem<kC>::TOTAL * 4, smg = (size_t)GeoSmem<kC, false>::TOTAL * 4;
            if (smt <= kMaxSmem && smg <= kMaxSmem) {
                masks = reinterpret_cast<uint64_t*>(scratch + ((10 * N + 3) / 4) * 4);
                TcSrc ts{}; ts.mode = 0; ts.points = points; ts.M = M;
                if (int e = launch_geo_decoder<kC, false>(planes, wpack, cfg, ts, N, nullptr, nullptr, nullptr, nullptr, masks, nullptr, st)) return e;
                if (int e = launch_tex_decoder<kC>(planes, wpack, cfg, ts, N, nullptr, masks, st)) return e;
                done = true;
            }
        });
        (void)done;
    }
    return launch_point_bwd(planes, wpack, cfg, src, N, gs, u, gf, masks, gplanes, gw, scratch + round4((size_t)N * 18 + 16), st);
}

size_t tt_wgrad_def_floats(int C) { return (size_t)64 * C + 4096 + 192; }

int tt_field_bwd(const float* planes, const float* wpack, const tt_config* cfg, const float* points, int64_t M,
                 const float* g_sdf, const float* g_deformation, float* scratch, float* gplanes, float* gw, float* gw_def,
                 void* stream) {
    if (int e = check_cfg(cfg)) return e;
    if (!planes || !wpack || !points || !scratch) return fail(TT_E_ARG, "tt_field_bwd: NULL pointer%s", "");
    if (!aligned16(planes) || !aligned16(wpack) || (gplanes && !aligned16(gplanes)))
        return fail(TT_E_ALIGN, "planes/wpack/gplanes must be 16-byte aligned%s", "");
    const int64_t N = (int64_t)cfg->P * M;
    if (N <= 0) return TT_OK;
    if (g_sdf && (gplanes || gw))
        if (int e = tt_geometry_bwd(planes, wpack, cfg, points, M, g_sdf, nullptr, nullptr, nullptr, scratch, gplanes, gw, stream)) return e;
    if (g_deformation && (gplanes || gw_def)) {
        const int64_t blocks = (N + TPB - 1) / TPB;
        if (blocks > 2147483647LL) return fail(TT_E_ARG, "tt_field_bwd: too many points%s (%lld)", "", N);
        PtSrc src; src.points = points; src.M = M; src.rs = RaySrc{nullptr, nullptr, nullptr, nullptr, 0, 1}; src.rays_per_cache = 1;
        TT_DISPATCH_C(cfg->C, {
            const size_t sm = slab_bytes(kC + HID + HID + 4);
            if (int e = set_smem(k_bwd_def<kC>, sm)) return e;
            TT_LAUNCH(k_bwd_def<kC>, (unsigned)blocks, TPB, sm, (cudaStream_t)stream, planes, wpack, *cfg, src, N, g_deformation, gplanes, gw_def);
        });
        if (int e = check_launch("k_bwd_def")) return e;
    }
    return TT_OK;
}

int tt_composite_fwd(const float* alphas, const float* values, int64_t n_rays, int S, int D, float* weights,
                     float* trans, float* out, void* stream) {
    if (!alphas || S < 1 || D < 0 || D > 8 || (D > 0 && !values)) return fail(TT_E_ARG, "tt_composite_fwd: bad arguments%s", "");
    if (n_rays <= 0) return TT_OK;
    if (g_impl >= 1) {      // one warp per ray (coalesced streams); the thread-per-ray kernel stays as the SIMT cross-check
        const unsigned grid = (unsigned)((n_rays + RAY_WARPS - 1) / RAY_WARPS);
        if (D <= 4) TT_LAUNCH(k_composite_fwd_w<4>, grid, RAY_WARPS * 32, 0, (cudaStream_t)stream, alphas, values, n_rays, S, D, weights, trans, out);
        else TT_LAUNCH(k_composite_fwd_w<8>, grid, RAY_WARPS * 32, 0, (cudaStream_t)stream, alphas, values, n_rays, S, D, weights, trans, out);
        return check_launch("tt_composite_fwd");
    }
    TT_LAUNCH(k_composite_fwd, (unsigned)((n_rays + 127) / 128), 128, 0, (cudaStream_t)stream, alphas, values, n_rays, S, D, weights, trans, out);
    return check_launch("tt_composite_fwd");
}
int tt_composite_bwd(const float* alphas, const float* values, const float* trans, const float* g_out,
                     const float* g_weights, int64_t n_rays, int S, int D, float* g_alphas, float* g_values,
                     void* stream) {
    if (!alphas || !trans || !g_alphas || S < 1 || D < 0 || D > 8 || (D > 0 && !values))
        return fail(TT_E_ARG, "tt_composite_bwd: bad arguments%s", "");
    if (n_rays <= 0) return TT_OK;
    if (g_impl >= 1) {
        const unsigned grid = (unsigned)((n_rays + RAY_WARPS - 1) / RAY_WARPS);
        if (D <= 4) TT_LAUNCH(k_composite_bwd_w<4>, grid, RAY_WARPS * 32, 0, (cudaStream_t)stream, alphas, values, trans, g_out, g_weights, n_rays, S, D, g_alphas, g_values);
        else TT_LAUNCH(k_composite_bwd_w<8>, grid, RAY_WARPS * 32, 0, (cudaStream_t)stream, alphas, values, trans, g_out, g_weights, n_rays, S, D, g_alphas, g_values);
        return check_launch("tt_composite_bwd");
    }
    TT_LAUNCH(k_composite_bwd, (unsigned)((n_rays + 127) / 128), 128, 0, (cudaStream_t)stream, alphas, values, trans, g_out, g_weights,
        n_rays, S, D, g_alphas, g_values);
    return check_launch("tt_composite_bwd");
}

// ---- stand-alone plane sampler (tt_sampler.cuh) -----------------------------------------------------------------------
static int sample_args(const char* who, const float* planes, int N, int K, int C, int H, int W, const float* grid, int64_t M,
                       int concat, SampleArgs* a) {
    if (!planes || !grid) return fail(TT_E_ARG, "%s: planes/grid NULL", who);
    if (N < 0 || K < 1 || K > SMP_MAXK || C < 4 || (C & 3) || H < 1 || W < 1 || M < 0)
        return fail(TT_E_ARG, "%s: bad sizes (1 <= K <= 4, C a multiple of 4)", who);
    if (!aligned16(planes)) return fail(TT_E_ALIGN, "%s: planes must be 16-byte aligned", who);
    int seg = 1;
    while (seg < (C >> 2) && seg < 32) seg <<= 1;
    const int lanes = seg > (C >> 2) ? seg : (C >> 2);
    *a = SampleArgs{planes, grid, N, K, C, H, W, M, concat ? 1 : 0, seg, ((int64_t)N * M * lanes + 256 < 0xffffffffLL && M < 0x7fffffffLL) ? 1 : 0};
    return TT_OK;
}
static unsigned sample_grid_dim(const SampleArgs& a) { return (unsigned)(((int64_t)a.N * a.M * a.seg + 255) / 256); }
static unsigned sample_grid_flat(const SampleArgs& a) { return (unsigned)(((int64_t)a.N * a.M * (a.C >> 2) + 255) / 256); }

int tt_sample_planes_fwd(const float* planes, int N, int K, int C, int H, int W, const float* grid, int64_t M, int concat,
                         float* out, void* stream) {
    SampleArgs a;
    if (int e = sample_args("tt_sample_planes_fwd", planes, N, K, C, H, W, grid, M, concat, &a)) return e;
    if (!out || !aligned16(out)) return fail(TT_E_ARG, "tt_sample_planes_fwd: out NULL or misaligned%s", "");
    if ((int64_t)N * M == 0) return TT_OK;
    const bool i32 = a.i32 != 0;       // item index fits 32 bits: 32-bit divisions
    switch (K) {
        case 1: if (i32) TT_LAUNCH((k_sample_fwd<1, true>), sample_grid_flat(a), 256, 0, (cudaStream_t)stream, a, out); else TT_LAUNCH((k_sample_fwd<1, false>), sample_grid_flat(a), 256, 0, (cudaStream_t)stream, a, out); break;
        case 2: if (i32) TT_LAUNCH((k_sample_fwd<2, true>), sample_grid_flat(a), 256, 0, (cudaStream_t)stream, a, out); else TT_LAUNCH((k_sample_fwd<2, false>), sample_grid_flat(a), 256, 0, (cudaStream_t)stream, a, out); break;
        case 3: if (i32) TT_LAUNCH((k_sample_fwd<3, true>), sample_grid_flat(a), 256, 0, (cudaStream_t)stream, a, out); else TT_LAUNCH((k_sample_fwd<3, false>), sample_grid_flat(a), 256, 0, (cudaStream_t)stream, a, out); break;
        default: if (i32) TT_LAUNCH((k_sample_fwd<4, true>), sample_grid_flat(a), 256, 0, (cudaStream_t)stream, a, out); else TT_LAUNCH((k_sample_fwd<4, false>), sample_grid_flat(a), 256, 0, (cudaStream_t)stream, a, out); break;
    }
    return check_launch("tt_sample_planes_fwd");
}
int tt_sample_planes_bwd(const float* planes, int N, int K, int C, int H, int W, const float* grid, int64_t M, int concat,
                         const float* g_out, float* g_planes, float* g_grid, void* stream) {
    SampleArgs a;
    if (int e = sample_args("tt_sample_planes_bwd", planes, N, K, C, H, W, grid, M, concat, &a)) return e;
    if (!g_out || !aligned16(g_out) || (g_planes && !aligned16(g_planes)))
        return fail(TT_E_ARG, "tt_sample_planes_bwd: g_out NULL or misaligned pointer%s", "");
    if ((int64_t)N * M == 0 || (!g_planes && !g_grid)) return TT_OK;
    TT_LAUNCH(k_sample_bwd, sample_grid_dim(a), 256, 0, (cudaStream_t)stream, a, g_out, g_planes, g_grid);
    return check_launch("tt_sample_planes_bwd");
}
int tt_sample_planes_bwdbwd(const float* planes, int N, int K, int C, int H, int W, const float* grid, int64_t M, int concat,
                            const float* g_out, const float* gg_planes, const float* gg_grid, float* gg_out,
                            float* g_planes, float* g_grid, void* stream) {
    SampleArgs a;
    if (int e = sample_args("tt_sample_planes_bwdbwd", planes, N, K, C, H, W, grid, M, concat, &a)) return e;
    if (!g_out || !aligned16(g_out) || (gg_planes && !aligned16(gg_planes)) || (gg_out && !aligned16(gg_out)) ||
        (g_planes && !aligned16(g_planes)))
        return fail(TT_E_ARG, "tt_sample_planes_bwdbwd: g_out NULL or misaligned pointer%s", "");
    if ((int64_t)N * M == 0 || (!gg_out && !g_planes && !g_grid)) return TT_OK;
    switch (K) {
        case 1: TT_LAUNCH(k_sample_bwdbwd<1>, sample_grid_dim(a), 256, 0, (cudaStream_t)stream, a, g_out, gg_planes, gg_grid, gg_out, g_planes, g_grid); break;
        case 2: TT_LAUNCH(k_sample_bwdbwd<2>, sample_grid_dim(a), 256, 0, (cudaStream_t)stream, a, g_out, gg_planes, gg_grid, gg_out, g_planes, g_grid); break;
        case 3: TT_LAUNCH(k_sample_bwdbwd<3>, sample_grid_dim(a), 256, 0, (cudaStream_t)stream, a, g_out, gg_planes, gg_grid, gg_out, g_planes, g_grid); break;
        default: TT_LAUNCH(k_sample_bwdbwd<4>, sample_grid_dim(a), 256, 0, (cudaStream_t)stream, a, g_out, gg_planes, gg_grid, gg_out, g_planes, g_grid); break;
    }
    return check_launch("tt_sample_planes_bwdbwd");
}
static int transpose_batched(const char* who, const float* src, int64_t B, int64_t rows, int64_t cols, float* dst, void* stream) {
    if (!src || !dst) return fail(TT_E_ARG, "%s: NULL pointer", who);
    if (B < 0 || rows < 1 || cols < 1 || B > 65535 || (rows + 31) / 32 > 65535) return fail(TT_E_ARG, "%s: bad sizes", who);
    if (B == 0) return TT_OK;
    TT_LAUNCH(k_transpose, dim3((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)B), 256, 0, (cudaStream_t)stream, src, dst, rows, cols);
    return check_launch(who);
}
int tt_to_channel_last(const float* src, int64_t B, int C, int64_t HW, float* dst, void* stream) {
    return transpose_batched("tt_to_channel_last", src, B, C, HW, dst, stream);
}
int tt_from_channel_last(const float* src, int64_t B, int C, int64_t HW, float* dst, void* stream) {
    return transpose_batched("tt_from_channel_last", src, B, HW, C, dst, stream);
}

}  // extern "C"
