// Probe: bilinear triplane gather at ray-coherent sample positions on sm_100a -- TMA tensor boxes against the LDG path.
//
//   planes  [6P][R][R][C] fp32, channel-last (the library's layout); one bilinear footprint = a 2 x 2 x C box
//   points  ray-major lists of in-box samples (uniform + a cluster around a "surface", sorted along the ray)
//
//   kernel "ldg":  the library's cooperative gather (tt::coop_gather, TT_GATHER_LOADS loads in flight per lane) run by
//                  G groups of 128 threads per SM, blended in registers, written to a point-major stage
//   kernel "tma":  NP producer warps issue ONE cp.async.bulk.tensor.4d per (point, plane) -- box {C,2,2,1}, out-of-bounds
//                  elements zero-filled by the hardware (= grid_sample's zeros padding) -- into a ring of NS stages of PT
//                  points; NB blend warps wait on the stage's mbarrier, read the raw texels from shared memory, blend and
//                  release the stage.  BLEND = 0 measures the bare TMA path (consumers only release).
// Every kernel accumulates  sum_points sum_c (c+1) * e[c]  and is checked against a plain one-thread-per-point kernel.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/tma_gather_probe tools/tma_gather_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../include/triplane_b200.h"
#include "../triplaneturbo_b200/csrc/tt_tc.cuh"

void tt_prof_pre(cudaStream_t) {}
void tt_prof_post(const char*, cudaStream_t) {}

#define CK(x)                                                                                              \
    do {                                                                                                   \
        cudaError_t e_ = (x);                                                                              \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
    } while (0)

using namespace tt;

__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (!done) __trap();
}
__device__ __forceinline__ void tma_box(void* dst, const CUtensorMap* tm, int x0, int y0, int plane, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(0), "r"(x0), "r"(y0), "r"(plane), "r"(smem_u32(bar)) : "memory");
}

struct Pt { float x, y, z; int prompt; };

// ---- reference: one thread per point ------------------------------------------------------------------------------
template <int C>
__global__ void k_ref(const float* __restrict__ planes, const Pt* __restrict__ pts, int n, int R, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (i < n) {
        const Pt q = pts[i];
        const float p[3] = {rescale1(q.x, 1.f), rescale1(q.y, 1.f), rescale1(q.z, 1.f)};
        const size_t ps = (size_t)R * R * C;
        for (int k = 0; k < 3; ++k) {
            const Taps t = make_taps(p[plane_ax(k)], p[plane_ay(k)], R);
            for (int a = 0; a < 4; ++a) {
                if (t.o[a] < 0) continue;
                const float* tex = planes + ((size_t)q.prompt * 6 + k) * ps + (size_t)t.o[a] * C;
                for (int c = 0; c < C; ++c) s += (double)(t.w[a] * tex[c]) * (c + 1);
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(out, s);
}

// ---- LDG path: G groups of 128 threads, the library's cooperative gather ------------------------------------------
template <int C, int G>
__global__ void __launch_bounds__(G * 128, 1) k_ldg(const float* __restrict__ planes, const Pt* __restrict__ pts, int n, int R,
                                                    double* out) {
    extern __shared__ __align__(16) float smem[];
    constexpr int SP = C + 4, GF = 128 * 12 * 2 + 128 + 128 * SP;
    const int tid = threadIdx.x, group = tid / 128, tg = tid % 128;
    float* gs = smem + group * GF;
    int* tap_o = reinterpret_cast<int*>(gs);
    float* tap_w = gs + 128 * 12;
    uint32_t* pbase = reinterpret_cast<uint32_t*>(gs + 128 * 24);
    float* stage = gs + 128 * 24 + 128;
    const size_t ps = (size_t)R * R * C;
    const int n_tiles = (n + 127) / 128;
    float cs = 0.f;
    for (int tile = blockIdx.x * G + group; tile < n_tiles; tile += gridDim.x * G) {
        const int i = tile * 128 + tg;
        const bool valid = i < n;
        Pt q = valid ? pts[i] : Pt{0.f, 0.f, 0.f, 0};
        const float p[3] = {rescale1(q.x, 1.f), rescale1(q.y, 1.f), rescale1(q.z, 1.f)};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const Taps t = make_taps(p[plane_ax(k)], p[plane_ay(k)], R);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const bool in = valid && t.o[a] >= 0;
                tap_o[tg * 12 + k * 4 + a] = in ? t.o[a] : 0;
                tap_w[tg * 12 + k * 4 + a] = in ? t.w[a] : 0.f;
            }
        }
        pbase[tg] = (uint32_t)q.prompt;
        group_sync(group);
        coop_gather<C, 3>(planes, ps, tap_o, tap_w, pbase, 0, stage, tg);
        group_sync(group);
#pragma unroll
        for (int c = 0; c < C; c += 4) {
            const float4 v = *reinterpret_cast<const float4*>(stage + tg * SP + c);
            cs += v.x * (c + 1) + v.y * (c + 2) + v.z * (c + 3) + v.w * (c + 4);
        }
        group_sync(group);
    }
    double s = cs;
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((tid & 31) == 0 && s != 0.0) atomicAdd(out, s);
}

// ---- TMA path -----------------------------------------------------------------------------------------------------
template <int C, int PT, int NS, int NP, int NB, int BLEND>
__global__ void __launch_bounds__((NP + NB) * 32, 1)
k_tma(const __grid_constant__ CUtensorMap tm, const Pt* __restrict__ pts, int n, int R, double* out) {
    extern __shared__ __align__(16) float smem_raw[];
    float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    constexpr int BOX = 4 * C;                       // floats per (point, plane) box
    constexpr int RAW = PT * 3 * BOX;                // floats per stage
    float* raw = smem;
    float* taps = smem + NS * RAW;                   // [NS][PT][12]
    uint64_t* full = reinterpret_cast<uint64_t*>(taps + NS * PT * 12);
    uint64_t* empty = full + NS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mb_init(full + s, 1); mb_init(empty + s, NB); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_chunks = (n + PT - 1) / PT;
    float cs = 0.f;
    if (warp < NP) {
        // ===== producers: chunk `it` of this CTA goes to stage it % NS; producer warp it % NP issues it
        int it = 0;
        for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
            if (it % NP != warp) continue;
            const int s = it % NS;
            mb_wait(empty + s, ((it / NS) & 1) ^ 1);
            for (int l0 = 0; l0 < PT; l0 += 32) {
                const int pl = l0 + lane;
                const int i = chunk * PT + pl;
                const bool valid = pl < PT && i < n;
                Pt q = valid ? pts[i] : Pt{0.f, 0.f, 0.f, 0};
                const float p[3] = {rescale1(q.x, 1.f), rescale1(q.y, 1.f), rescale1(q.z, 1.f)};
                int x0[3], y0[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float gx = p[plane_ax(k)], gy = p[plane_ay(k)];
                    const Taps t = make_taps(gx, gy, R);
                    const float fR = (float)R;
                    const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), fR), 1.f), 0.5f);
                    const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), fR), 1.f), 0.5f);
                    x0[k] = (int)fminf(fmaxf(floorf(ix), -2.f), fR);
                    y0[k] = (int)fminf(fmaxf(floorf(iy), -2.f), fR);
                    if (pl < PT)
#pragma unroll
                        for (int a = 0; a < 4; ++a) taps[(s * PT + pl) * 12 + k * 4 + a] = valid ? t.w[a] : 0.f;
                }
                __syncwarp();
                if (l0 == 0 && lane == 0) mb_expect_tx(full + s, (uint32_t)(RAW * 4));
                __syncwarp();
                if (pl < PT)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        tma_box(raw + (size_t)s * RAW + (pl * 3 + k) * BOX, &tm, x0[k], y0[k], q.prompt * 6 + k, full + s);
            }
        }
    } else {
        // ===== blend warps
        const int cw = warp - NP;
        constexpr int U = C / 4;
        int it = 0;
        for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
            const int s = it % NS;
            mb_wait(full + s, (it / NS) & 1);
            if (BLEND) {
                for (int item = cw * 32 + lane; item < PT * U; item += NB * 32) {
                    const int pt = item / U, ch = item - pt * U;
                    const float* w = taps + (s * PT + pt) * 12;
                    const float* rp = raw + (size_t)s * RAW + pt * 3 * BOX + ch * 4;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float4 w4 = *reinterpret_cast<const float4*>(w + k * 4);
                        const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
                        float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const float4 q = *reinterpret_cast<const float4*>(rp + k * BOX + a * C);
                            sm.x = fmaf(ww[a], q.x, sm.x); sm.y = fmaf(ww[a], q.y, sm.y);
                            sm.z = fmaf(ww[a], q.z, sm.z); sm.w = fmaf(ww[a], q.w, sm.w);
                        }
                        acc.x += sm.x; acc.y += sm.y; acc.z += sm.z; acc.w += sm.w;
                    }
                    const int c = ch * 4;
                    cs += acc.x * (c + 1) + acc.y * (c + 2) + acc.z * (c + 3) + acc.w * (c + 4);
                }
            }
            __syncwarp();
            if (lane == 0) mb_arrive(empty + s);
        }
    }
    double s = cs;
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0 && s != 0.0) atomicAdd(out, s);
}

// ---- host -----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// image-space variant: rays of a pinhole camera through a W x W image; order = (16x8 pixel patch, sample, pixel in patch), so a
// 128-point tile holds ONE sample index of 128 ADJACENT rays (their 2x2 footprints overlap) instead of 128 samples of one ray
static std::vector<Pt> make_points_patch(int W, int S, int P, bool patch_major, unsigned seed) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::vector<Pt> pts;
    const float o[3] = {2.0f, 0.6f, 0.9f};
    float f[3] = {-o[0], -o[1], -o[2]}, fl = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (auto& v : f) v /= fl;
    float r[3] = {f[1], -f[0], 0.f}, rl = std::sqrt(r[0] * r[0] + r[1] * r[1]);
    for (auto& v : r) v /= rl;
    const float u[3] = {r[1] * f[2] - r[2] * f[1], r[2] * f[0] - r[0] * f[2], r[0] * f[1] - r[1] * f[0]};
    const float tanh_ = 0.5774f;
    auto ray_pts = [&](int px, int py, std::vector<Pt>& out) {
        const float sx = ((px + 0.5f) / W * 2 - 1) * tanh_, sy = ((py + 0.5f) / W * 2 - 1) * tanh_;
        float d[3]; float dl = 0;
        for (int a = 0; a < 3; ++a) { d[a] = f[a] + sx * r[a] + sy * u[a]; dl += d[a] * d[a]; }
        dl = std::sqrt(dl);
        for (auto& v : d) v /= dl;
        for (int i = 0; i < S; ++i) {
            const float t = 1.2f + 2.0f * (i + 0.5f) / S;
            out.push_back(Pt{o[0] + d[0] * t, o[1] + d[1] * t, o[2] + d[2] * t, 0});
        }
    };
    if (!patch_major) {
        for (int py = 0; py < W; ++py) for (int px = 0; px < W; ++px) ray_pts(px, py, pts);
    } else {
        for (int by = 0; by < W; by += 8) for (int bx = 0; bx < W; bx += 16) {
            std::vector<Pt> tmp;
            for (int y = 0; y < 8; ++y) for (int x = 0; x < 16; ++x) ray_pts(bx + x, by + y, tmp);
            for (int i = 0; i < S; ++i) for (int q = 0; q < 128; ++q) pts.push_back(tmp[(size_t)q * S + i]);
        }
    }
    // keep in-box points only (the kernels skip the others)
    std::vector<Pt> in;
    for (auto& p : pts) if (std::fabs(p.x) < 1.004f && std::fabs(p.y) < 1.004f && std::fabs(p.z) < 1.004f) in.push_back(p);
    (void)P; (void)U; (void)rng;
    return in;
}

static std::vector<Pt> make_points(int n_rays, int S, int P, float spread, unsigned seed) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::normal_distribution<float> N(0.f, 1.f);
    std::vector<Pt> pts;
    pts.reserve((size_t)n_rays * S);
    for (int r = 0; r < n_rays; ++r) {
        // camera on a sphere of radius 2.5 looking at a point near the origin
        float o[3], d[3], len = 0.f;
        for (int a = 0; a < 3; ++a) { o[a] = N(rng); len += o[a] * o[a]; }
        len = std::sqrt(len);
        for (int a = 0; a < 3; ++a) o[a] *= 2.5f / len;
        len = 0.f;
        for (int a = 0; a < 3; ++a) { d[a] = -o[a] + 0.9f * (U(rng) - 0.5f); len += d[a] * d[a]; }
        len = std::sqrt(len);
        for (int a = 0; a < 3; ++a) d[a] /= len;
        float t0 = 0.f, t1 = 1e9f;           // slab test against [-1.004, 1.004]^3 (a few taps fall out of bounds)
        for (int a = 0; a < 3; ++a) {
            const float ta = (-1.004f - o[a]) / d[a], tb = (1.004f - o[a]) / d[a];
            t0 = std::max(t0, std::min(ta, tb)); t1 = std::min(t1, std::max(ta, tb));
        }
        if (t1 <= t0) { --r; continue; }
        std::vector<float> ts(S);
        const float tsurf = t0 + (t1 - t0) * (0.25f + 0.5f * U(rng));
        for (int i = 0; i < S; ++i) {
            if (i < (2 * S) / 3) ts[i] = t0 + (t1 - t0) * (i + U(rng)) / ((2 * S) / 3);
            else ts[i] = std::min(std::max(tsurf + spread * N(rng), t0), t1);
        }
        std::sort(ts.begin(), ts.end());
        const int prompt = (int)((int64_t)r * P / n_rays);
        for (int i = 0; i < S; ++i) pts.push_back(Pt{o[0] + d[0] * ts[i], o[1] + d[1] * ts[i], o[2] + d[2] * ts[i], prompt});
    }
    return pts;
}

template <typename F>
static float time_ms(F f, int warm = 2, int iters = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < warm; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

template <int C>
static void run(int P, int R, int n_rays, int S, float spread, int patch = -1) {
    const size_t ps = (size_t)R * R * C, nplane = (size_t)P * 6 * ps;
    std::vector<float> h(nplane);
    std::mt19937 rng(1);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    for (auto& v : h) v = U(rng);
    float* planes; CK(cudaMalloc(&planes, nplane * 4));
    CK(cudaMemcpy(planes, h.data(), nplane * 4, cudaMemcpyHostToDevice));
    std::vector<Pt> pts = patch < 0 ? make_points(n_rays, S, P, spread, 7) : make_points_patch(512, 128, P, patch == 1, 7);
    if (patch >= 0) printf("-- 512^2 pinhole rays x 128 samples, order: %s\n", patch ? "(16x8 pixel patch, sample, pixel)" : "ray-major");
    const int n = (int)pts.size();
    Pt* dp; CK(cudaMalloc(&dp, (size_t)n * sizeof(Pt)));
    CK(cudaMemcpy(dp, pts.data(), (size_t)n * sizeof(Pt), cudaMemcpyHostToDevice));
    double* out; CK(cudaMalloc(&out, 8));
    int sms = 148; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const double logical = (double)n * 12.0 * 4.0 * C;
    printf("== C=%d P=%d R=%d rays=%d S=%d spread=%.3f points=%d  logical texel bytes %.2f GB, %d SMs\n", C, P, R, n_rays, S,
           spread, n, logical / 1e9, sms);

    auto result = [&]() { double v; CK(cudaMemcpy(&v, out, 8, cudaMemcpyDeviceToHost)); return v; };
    CK(cudaMemset(out, 0, 8));
    k_ref<C><<<(n + 255) / 256, 256>>>(planes, dp, n, R, out);
    CK(cudaDeviceSynchronize());
    const double ref = result();
    printf("   reference checksum %.6e\n", ref);

    auto report = [&](const char* name, float ms, double got, int iters_total) {
        const double per = got / iters_total;
        printf("   %-44s %8.3f ms  %6.2f TB/s logical   checksum rel err %.2e\n", name, ms, logical / (ms * 1e-3) / 1e12,
               std::fabs(per - ref) / (std::fabs(ref) + 1e-30));
    };
    // ---- LDG path
    auto ldg = [&](auto gtag, const char* name, size_t extra = 0) {
        constexpr int G = decltype(gtag)::value;
        const size_t sm = (size_t)G * (128 * 24 + 128 + 128 * (C + 4)) * 4 + extra;
        CK(cudaFuncSetAttribute(k_ldg<C, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaMemset(out, 0, 8));
        const float ms = time_ms([&]() { k_ldg<C, G><<<sms, G * 128, sm>>>(planes, dp, n, R, out); });
        CK(cudaGetLastError());
        report(name, ms, result(), 7);
    };
    ldg(std::integral_constant<int, 1>{}, "ldg  4 warps/SM (24 loads in flight)");
    ldg(std::integral_constant<int, 2>{}, "ldg  8 warps/SM");
    ldg(std::integral_constant<int, 3>{}, "ldg 12 warps/SM");
    // the decoder kernels keep 160-215 KB of shared memory: what the same gather gets with the L1 that is left
    ldg(std::integral_constant<int, 1>{}, "ldg  4 warps/SM, +130 KB smem (L1 ~ 64 KB)", 130 * 1024);
    ldg(std::integral_constant<int, 1>{}, "ldg  4 warps/SM, +170 KB smem (L1 ~ 28 KB)", 170 * 1024);
    ldg(std::integral_constant<int, 2>{}, "ldg  8 warps/SM, +100 KB smem (L1 ~ 64 KB)", 100 * 1024);
    ldg(std::integral_constant<int, 2>{}, "ldg  8 warps/SM, +130 KB smem (L1 ~ 28 KB)", 130 * 1024);

    if (patch >= 0) { CK(cudaFree(planes)); CK(cudaFree(dp)); CK(cudaFree(out)); return; }
    // ---- TMA path
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled unavailable\n"); exit(1); }
    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)R, (cuuint64_t)R, (cuuint64_t)P * 6};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)R * C * 4, (cuuint64_t)ps * 4};
    const cuuint32_t box[4] = {(cuuint32_t)C, 2, 2, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, planes, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); exit(1); }

    auto tma = [&](auto pt_, auto ns_, auto np_, auto nb_, auto bl_, const char* name) {
        constexpr int PT = decltype(pt_)::value, NS = decltype(ns_)::value, NP = decltype(np_)::value,
                      NB = decltype(nb_)::value, BL = decltype(bl_)::value;
        const size_t sm = (size_t)(NS * PT * 3 * 4 * C + NS * PT * 12) * 4 + 2 * NS * 8 + 128;
        CK(cudaFuncSetAttribute((k_tma<C, PT, NS, NP, NB, BL>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaMemset(out, 0, 8));
        const float ms = time_ms([&]() { k_tma<C, PT, NS, NP, NB, BL><<<sms, (NP + NB) * 32, sm>>>(tm, dp, n, R, out); });
        CK(cudaGetLastError());
        report(name, ms, BL ? result() : ref * 7, 7);
    };
    using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
    using I3 = std::integral_constant<int, 3>; using I4 = std::integral_constant<int, 4>; using I8 = std::integral_constant<int, 8>;
    using I16 = std::integral_constant<int, 16>; using I32 = std::integral_constant<int, 32>;
    tma(I32{}, I2{}, I1{}, I4{}, I0{}, "tma  PT=32 NS=2 1 prod, no blend (bare TMA)");
    tma(I32{}, I3{}, I1{}, I4{}, I0{}, "tma  PT=32 NS=3 1 prod, no blend (bare TMA)");
    tma(I32{}, I3{}, I2{}, I4{}, I0{}, "tma  PT=32 NS=3 2 prod, no blend (bare TMA)");
    tma(I16{}, I4{}, I1{}, I4{}, I0{}, "tma  PT=16 NS=4 1 prod, no blend (bare TMA)");
    tma(I32{}, I2{}, I1{}, I4{}, I1{}, "tma  PT=32 NS=2 1 prod, 4 blend warps");
    tma(I32{}, I3{}, I1{}, I4{}, I1{}, "tma  PT=32 NS=3 1 prod, 4 blend warps");
    tma(I32{}, I3{}, I2{}, I4{}, I1{}, "tma  PT=32 NS=3 2 prod, 4 blend warps");
    tma(I32{}, I3{}, I2{}, I8{}, I1{}, "tma  PT=32 NS=3 2 prod, 8 blend warps");
    tma(I16{}, I4{}, I1{}, I4{}, I1{}, "tma  PT=16 NS=4 1 prod, 4 blend warps");
    tma(I16{}, I4{}, I2{}, I8{}, I1{}, "tma  PT=16 NS=4 2 prod, 8 blend warps");
    CK(cudaFree(planes)); CK(cudaFree(dp)); CK(cudaFree(out));
}

int main(int argc, char** argv) {
    const int n_rays = argc > 1 ? atoi(argv[1]) : 32768;
    run<32>(4, 256, n_rays, 128, 0.05f);
    run<32>(4, 256, n_rays, 128, 0.4f);
    run<40>(4, 256, n_rays, 128, 0.05f);
    run<32>(4, 256, n_rays, 128, 0.05f, 0);
    run<32>(4, 256, n_rays, 128, 0.05f, 1);
    return 0;
}
