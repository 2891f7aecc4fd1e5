// Device-side building blocks shared by every kernel of libtriplane_b200 (sm_100a).
//
// Execution model ("column slabs"): one thread owns one sample point.  Its activation vectors live in
// shared memory as private columns  slot[row * ST + tid]  (row = channel / hidden unit), so the
// per-point decoder layers need no synchronisation, and the CTA-wide weight-gradient contractions
// (sum over the CTA's points) read the same slabs row-wise after a __syncthreads().
//
// Reference arithmetic restated here (paths relative to the reference root):
//   rescale_points / scale_tensor      threestudio/utils/ops.py:27-38
//   plane projection                   custom/triplaneturbo/models/geometry/utils.py:46-63,111-125
//   bilinear taps (zeros, !align)      ATen grid_sampler_2d as called at geometry/utils.py:21-24
//   VanillaMLP                         threestudio/models/networks.py:67-104
//   shifted SDF                        …/few_step_triplane_dual_stable_diffusion.py:131-154
//   analytic normal                    …/few_step_triplane_dual_stable_diffusion.py:329-339
//   second derivative of the sampler   extern/grid_sample_gradfix/gridsample_cuda.cu:87-209
//   NeuS alpha                         threestudio/models/renderers/neus_volume_renderer.py:93-117
#pragma once
#ifndef TT_EMUL   // TT_EMUL: host emulation used only by tests/emul (see tests/emul/cuda_emul.h)
#include <cuda_runtime.h>
void tt_prof_pre(cudaStream_t st);
void tt_prof_post(const char* name, cudaStream_t st);
#define TT_LAUNCH(kernel, grid, block, smem, stream, ...)                \
    do {                                                                 \
        tt_prof_pre(stream);                                             \
        kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);              \
        tt_prof_post(#kernel, stream);                                   \
    } while (0)
#define TT_SHARED(name) extern __shared__ __align__(16) float name[]
#endif
#include <stdint.h>

namespace tt {

constexpr int HID = 64;       // hidden units of every decoder MLP (networks.py:74-88, n_neurons 64)
constexpr int TPB = 128;      // threads (= sample points) per CTA
constexpr int ST = TPB + 4;   // column-slab row stride in floats (+4: conflict-free float4 row reads)

// ---- packed decoder weights -------------------------------------------------------------------
struct WOff {
    int w1s, w1sT, w2s, w2sT, w3s;   // sdf:      [64][C], [C][64], [64][64], [64][64]^T, [64]
    int w1f, w1fT, w2f, w2fT, w3f;   // feature:  [64][3C], [3C][64], [64][64], ^T, [3][64]
    int w1d, w1dT, w2d, w2dT, w3d;   // deform.:  like sdf, head [3][64]
    int total;
};
__host__ __device__ inline WOff woff(int C) {
    WOff o; int p = 0;
    o.w1s = p; p += 64 * C; o.w1sT = p; p += 64 * C; o.w2s = p; p += 4096; o.w2sT = p; p += 4096; o.w3s = p; p += 64;
    o.w1f = p; p += 192 * C; o.w1fT = p; p += 192 * C; o.w2f = p; p += 4096; o.w2fT = p; p += 4096; o.w3f = p; p += 192;
    o.w1d = p; p += 64 * C; o.w1dT = p; p += 64 * C; o.w2d = p; p += 4096; o.w2dT = p; p += 4096; o.w3d = p; p += 192;
    o.total = p;
    return o;
}
struct GOff { int g1s, g2s, g3s, g1f, g2f, g3f, total; };   // [out][in] blocks, nn.Linear layout
__host__ __device__ inline GOff goff(int C) {
    GOff o; int p = 0;
    o.g1s = p; p += 64 * C; o.g2s = p; p += 4096; o.g3s = p; p += 64;
    o.g1f = p; p += 192 * C; o.g2f = p; p += 4096; o.g3f = p; p += 192;
    o.total = p;
    return o;
}

// ---- coordinates --------------------------------------------------------------------------------
// scale_tensor(x, bbox=(-r, r), (-1, 1)) with the reference's operation order and no FMA contraction,
// so that the integer corner indices below are bit-exact against the oracle.
__device__ __forceinline__ float rescale1(float x, float r) {
    float lo = -r;
    float t = __fdiv_rn(__fsub_rn(x, lo), __fsub_rn(r, lo));
    return __fadd_rn(__fmul_rn(t, 2.f), -1.f);
}

// plane k samples (gx, gy) = (p[AX[k]], p[AY[k]])  (geometry/utils.py:46-63 after the bmm with the inverse)
__device__ __forceinline__ int plane_ax(int k) { return k == 2 ? 2 : 0; }
__device__ __forceinline__ int plane_ay(int k) { return k == 1 ? 2 : 1; }

struct Taps {
    int o[4];                   // texel index y*R+x of nw, ne, sw, se; -1 when out of bounds (zeros padding)
    float w[4];                 // bilinear weights nw, ne, sw, se
    float wx0, wx1, wy0, wy1;   // x1-ix, ix-x0, y1-iy, iy-y0
};
__device__ __forceinline__ Taps make_taps(float gx, float gy, int R) {
    Taps t;
    const float fR = (float)R;
    const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), fR), 1.f), 0.5f);
    const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), fR), 1.f), 0.5f);
    const float x0f = floorf(ix), y0f = floorf(iy);
    t.wx0 = __fsub_rn(__fadd_rn(x0f, 1.f), ix); t.wx1 = __fsub_rn(ix, x0f);
    t.wy0 = __fsub_rn(__fadd_rn(y0f, 1.f), iy); t.wy1 = __fsub_rn(iy, y0f);
    t.w[0] = __fmul_rn(t.wx0, t.wy0); t.w[1] = __fmul_rn(t.wx1, t.wy0);
    t.w[2] = __fmul_rn(t.wx0, t.wy1); t.w[3] = __fmul_rn(t.wx1, t.wy1);
    // clamp before the int conversion so far-away points cannot overflow
    const float lim = fR + 2.f;
    const int x0 = (int)fminf(fmaxf(x0f, -2.f), lim), y0 = (int)fminf(fmaxf(y0f, -2.f), lim);
    const bool xa = x0 >= 0 && x0 < R, xb = x0 + 1 >= 0 && x0 + 1 < R;
    const bool ya = y0 >= 0 && y0 < R, yb = y0 + 1 >= 0 && y0 + 1 < R;
    t.o[0] = (xa && ya) ? y0 * R + x0 : -1;
    t.o[1] = (xb && ya) ? y0 * R + x0 + 1 : -1;
    t.o[2] = (xa && yb) ? (y0 + 1) * R + x0 : -1;
    t.o[3] = (xb && yb) ? (y0 + 1) * R + x0 + 1 : -1;
    return t;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---- gather: Σ_planes Σ_taps w * texel  ->  column (C rows) ---------------------------------------
// base points at plane 0 of the group; plane k is at base + k*plane_stride.  Channel-last texels.
template <int C, int NPL>
__device__ __forceinline__ void gather(const float* __restrict__ base, size_t plane_stride,
                                       const int* o, const float* w, float* col) {
#pragma unroll 2
    for (int c = 0; c < C; c += 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (o[k * 4 + t] >= 0) {
                    const float4 v = ldg4(base + k * plane_stride + (size_t)o[k * 4 + t] * C + c);
                    const float ww = w[k * 4 + t];
                    s.x = fmaf(ww, v.x, s.x); s.y = fmaf(ww, v.y, s.y);
                    s.z = fmaf(ww, v.z, s.z); s.w = fmaf(ww, v.w, s.w);
                }
            }
            acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w;
        }
        col[(c + 0) * ST] = acc.x; col[(c + 1) * ST] = acc.y;
        col[(c + 2) * ST] = acc.z; col[(c + 3) * ST] = acc.w;
    }
}

// scatter-add  v[c] * w  into the channel-last gradient planes (vector reductions, 16 B each)
__device__ __forceinline__ void red_add4(float* addr, float4 v) {
#ifdef TT_EMUL
    atomicAdd(addr, v.x); atomicAdd(addr + 1, v.y); atomicAdd(addr + 2, v.z); atomicAdd(addr + 3, v.w);
    return;
#else
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
#endif
}
template <int C, int NPL>
__device__ __forceinline__ void scatter_col(float* __restrict__ gbase, size_t plane_stride, const int* o,
                                            const float* w, const float* col) {
    for (int c = 0; c < C; c += 4) {
        const float4 v = make_float4(col[(c + 0) * ST], col[(c + 1) * ST], col[(c + 2) * ST], col[(c + 3) * ST]);
#pragma unroll
        for (int k = 0; k < NPL; ++k)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float ww = w[k * 4 + t];
                if (o[k * 4 + t] >= 0 && ww != 0.f)
                    red_add4(gbase + k * plane_stride + (size_t)o[k * 4 + t] * C + c,
                             make_float4(v.x * ww, v.y * ww, v.z * ww, v.w * ww));
            }
    }
}

// ---- per-point decoder layers -----------------------------------------------------------------------
// acc[0..63] += Σ_k WT[k][0..63] * x[k]    (WT: transposed nn.Linear weight, rows of 64)
__device__ __forceinline__ void layer64_acc(float (&acc)[HID], const float* __restrict__ WT, int K, const float* xcol) {
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const float x = xcol[k * ST];
        const float* wr = WT + k * HID;
#pragma unroll
        for (int j = 0; j < HID / 4; ++j) {
            const float4 w = ldg4(wr + 4 * j);
            acc[4 * j + 0] = fmaf(x, w.x, acc[4 * j + 0]); acc[4 * j + 1] = fmaf(x, w.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(x, w.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(x, w.w, acc[4 * j + 3]);
        }
    }
}
__device__ __forceinline__ void zero64(float (&acc)[HID]) {
#pragma unroll
    for (int j = 0; j < HID; ++j) acc[j] = 0.f;
}
// ReLU + store to a column; returns the activation mask (bit o = pre-activation > 0)
__device__ __forceinline__ uint64_t store_relu64(const float (&acc)[HID], float* ycol) {
    uint64_t m = 0;
#pragma unroll
    for (int j = 0; j < HID; ++j) {
        const bool on = acc[j] > 0.f;
        m |= (uint64_t)on << j;
        ycol[j * ST] = on ? acc[j] : 0.f;
    }
    return m;
}
__device__ __forceinline__ uint64_t mask64(const float (&acc)[HID]) {
    uint64_t m = 0;
#pragma unroll
    for (int j = 0; j < HID; ++j) m |= (uint64_t)(acc[j] > 0.f) << j;
    return m;
}
// masked store: ycol[j] = mask_j ? acc[j] : 0
__device__ __forceinline__ void store_masked64(const float (&acc)[HID], uint64_t m, float* ycol) {
#pragma unroll
    for (int j = 0; j < HID; ++j) ycol[j * ST] = ((m >> j) & 1ull) ? acc[j] : 0.f;
}
// transposed product: acc[0..N) += Σ_o W[o*ld + 0..N) * a[o]   (o over the 64 hidden units)
template <int N>
__device__ __forceinline__ void layerT_acc(float (&acc)[N], const float* __restrict__ W, int ld, const float* acol) {
#pragma unroll 2
    for (int o = 0; o < HID; ++o) {
        const float a = acol[o * ST];
        const float* wr = W + o * ld;
#pragma unroll
        for (int j = 0; j < N / 4; ++j) {
            const float4 w = ldg4(wr + 4 * j);
            acc[4 * j + 0] = fmaf(a, w.x, acc[4 * j + 0]); acc[4 * j + 1] = fmaf(a, w.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(a, w.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(a, w.w, acc[4 * j + 3]);
        }
    }
}
__device__ __forceinline__ void zero_col(float* col, int rows) {
    for (int r = 0; r < rows; ++r) col[r * ST] = 0.f;
}

// ---- CTA-wide weight gradient:  G[n*ldg + k] += Σ_p A[n][p] * X[k][p]  ---------------------------------
// A, X: slab bases (column 0).  N rows of A (any N >= 1), K rows of X (multiple of 4).  Call between
// __syncthreads(); every thread of the CTA must call it.
__device__ __forceinline__ void wgrad(float* __restrict__ G, int ldg, const float* A, int N, const float* X, int K) {
    const int tk = K >> 2, tn = (N + 3) >> 2;
    for (int tile = threadIdx.x; tile < tn * tk; tile += TPB) {
        const int n0 = (tile / tk) << 2, k0 = (tile % tk) << 2;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int p = 0; p < TPB; p += 4) {
            float4 a[4], x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                a[i] = (n0 + i < N) ? *reinterpret_cast<const float4*>(A + (n0 + i) * ST + p) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = *reinterpret_cast<const float4*>(X + (k0 + j) * ST + p);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] += a[i].x * x[j].x + a[i].y * x[j].y + a[i].z * x[j].z + a[i].w * x[j].w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (n0 + i < N)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (acc[i][j] != 0.f) atomicAdd(G + (size_t)(n0 + i) * ldg + k0 + j, acc[i][j]);
    }
}

// ---- geometry at one point -------------------------------------------------------------------------------
struct PointTaps {           // taps of the three planes of one group for one point (index k*4 + tap)
    int o[12];
    float w[12];
};
__device__ __forceinline__ void point_taps(const float (&p)[3], int R, PointTaps& pt, Taps (&tp)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        tp[k] = make_taps(p[plane_ax(k)], p[plane_ay(k)], R);
#pragma unroll
        for (int t = 0; t < 4; ++t) { pt.o[k * 4 + t] = tp[k].o[t]; pt.w[k * 4 + t] = tp[k].w[t]; }
    }
}

// d(Σ_c de[c] * f_k[c]) / d(ix, iy) of plane k from the tap dot products A_t = Σ_c de[c] * texel_t[c]
__device__ __forceinline__ void tap_grad(const Taps& t, const float (&A)[4], float& dix, float& diy) {
    dix = (A[1] - A[0]) * t.wy0 + (A[3] - A[2]) * t.wy1;
    diy = (A[2] - A[0]) * t.wx0 + (A[3] - A[1]) * t.wx1;
}

// SDF decoder at world point x: sdf_orig, shifted sdf and (NORMAL) sdf_grad = d sdf / d x.
// slotX: >= C rows (receives the geometry encoding), slotB: 64 rows (scratch).
template <int C, bool NORMAL>
__device__ __forceinline__ void geo_eval(const float* __restrict__ gplanes3, int R, const float* __restrict__ wp,
                                         const WOff& wo, const float (&x)[3], float radius, float bias_r,
                                         float* slotX, float* slotB, float& sdf_orig, float& sdf, float (&g)[3]) {
    float p[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = rescale1(x[a], radius);
    PointTaps pt; Taps tp[3];
    point_taps(p, R, pt, tp);
    const size_t ps = (size_t)R * R * C;
    gather<C, 3>(gplanes3, ps, pt.o, pt.w, slotX);
    float acc[HID];
    zero64(acc);
    layer64_acc(acc, wp + wo.w1sT, C, slotX);
    const uint64_t m1 = store_relu64(acc, slotB);
    zero64(acc);
    layer64_acc(acc, wp + wo.w2sT, HID, slotB);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < HID; ++j) {
        const float w3 = __ldg(wp + wo.w3s + j);
        const bool on = acc[j] > 0.f;
        s = fmaf(on ? acc[j] : 0.f, w3, s);
        if (NORMAL) slotB[j * ST] = on ? w3 : 0.f;      // a2 = m2 ⊙ w3
    }
    const float nrm = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    sdf_orig = s;
    sdf = s + (nrm - bias_r);
    if (NORMAL) {
        zero64(acc);
        layerT_acc<HID>(acc, wp + wo.w2s, HID, slotB);  // W2^T a2
        store_masked64(acc, m1, slotB);                 // a1
        float de[C];
#pragma unroll
        for (int c = 0; c < C; ++c) de[c] = 0.f;
        layerT_acc<C>(de, wp + wo.w1s, C, slotB);       // d sdf_orig / d enc
        float gm[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float A[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float d = 0.f;
                if (pt.o[k * 4 + t] >= 0) {
                    const float* tex = gplanes3 + k * ps + (size_t)pt.o[k * 4 + t] * C;
#pragma unroll
                    for (int c = 0; c < C; c += 4) {
                        const float4 v = ldg4(tex + c);
                        d = fmaf(de[c], v.x, d); d = fmaf(de[c + 1], v.y, d);
                        d = fmaf(de[c + 2], v.z, d); d = fmaf(de[c + 3], v.w, d);
                    }
                }
                A[t] = d;
            }
            float dix, diy;
            tap_grad(tp[k], A, dix, diy);
            gm[plane_ax(k)] += dix;
            gm[plane_ay(k)] += diy;
        }
        const float scale = 0.5f * (float)R / radius;   // d ix / d x_world
        const float inv = nrm > 0.f ? 1.f / nrm : 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) g[a] = gm[a] * scale + x[a] * inv;
    }
}

// colour decoder (tex planes concatenated, feature MLP) at rescaled point p -> 3 features (pre-activation)
template <int C>
__device__ __forceinline__ void tex_eval(const float* __restrict__ tplanes3, int R, const float* __restrict__ wp,
                                         const WOff& wo, const float (&p)[3], float* slotX, float* slotB,
                                         float (&f)[3]) {
    const size_t ps = (size_t)R * R * C;
    float acc[HID];
    zero64(acc);
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        const Taps t = make_taps(p[plane_ax(k)], p[plane_ay(k)], R);
        gather<C, 1>(tplanes3 + k * ps, 0, t.o, t.w, slotX);
        layer64_acc(acc, wp + wo.w1fT + k * C * HID, C, slotX);
    }
    store_relu64(acc, slotB);
    zero64(acc);
    layer64_acc(acc, wp + wo.w2fT, HID, slotB);
    f[0] = f[1] = f[2] = 0.f;
#pragma unroll
    for (int j = 0; j < HID; ++j) {
        const float h = fmaxf(acc[j], 0.f);
        f[0] = fmaf(h, __ldg(wp + wo.w3f + j), f[0]);
        f[1] = fmaf(h, __ldg(wp + wo.w3f + HID + j), f[1]);
        f[2] = fmaf(h, __ldg(wp + wo.w3f + 2 * HID + j), f[2]);
    }
}

// ---- NeuS alpha (neus_volume_renderer.py:93-117, use_volsdf = False) ---------------------------------------
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

struct AlphaTerms { float alpha_raw, alpha, prev_cdf, next_cdf, s_prev, s_next, true_cos, iter_cos; };
__device__ __forceinline__ AlphaTerms neus_alpha(float sdf, const float (&n)[3], const float (&d)[3], float dt,
                                                 float inv_std, float car) {
    AlphaTerms a;
    a.true_cos = d[0] * n[0] + d[1] * n[1] + d[2] * n[2];
    a.iter_cos = -(fmaxf(-a.true_cos * 0.5f + 0.5f, 0.f) * (1.f - car) + fmaxf(-a.true_cos, 0.f) * car);
    a.s_next = sdf + a.iter_cos * dt * 0.5f;
    a.s_prev = sdf - a.iter_cos * dt * 0.5f;
    a.prev_cdf = sigmoidf(a.s_prev * inv_std);
    a.next_cdf = sigmoidf(a.s_next * inv_std);
    a.alpha_raw = (a.prev_cdf - a.next_cdf + 1e-5f) / (a.prev_cdf + 1e-5f);
    a.alpha = fminf(fmaxf(a.alpha_raw, 0.f), 1.f);
    return a;
}
__device__ __forceinline__ void normalize3(const float (&g)[3], float (&n)[3], float& len) {
    len = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const float inv = 1.f / fmaxf(len, 1e-12f);     // F.normalize eps
    n[0] = g[0] * inv; n[1] = g[1] * inv; n[2] = g[2] * inv;
}
__device__ __forceinline__ float sigmoid_mipnerf(float f) { return sigmoidf(f) * 1.002f - 0.001f; }

}  // namespace tt
