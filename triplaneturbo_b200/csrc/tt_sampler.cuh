// Stand-alone plane sampler: bilinear grid_sample (zeros padding, align_corners = False) of channel-last feature planes,
// with its first and second derivative.  Replaces, as an operator of its own,
//   custom/triplaneturbo/models/geometry/utils.py:21-24,127-145   grid_sample / sample_from_planes
//   custom/triplaneturbo/extern/grid_sample_gradfix/cuda_gridsample.py:22-79   _GridSample2dForward / _GridSample2dBackward
//   custom/triplaneturbo/extern/grid_sample_gradfix/gridsample_cuda.cu:27-210  grid_sampler_2d_grad2_kernel
// (inside the renderer the same arithmetic is fused into the decoder kernels; this is the functional API of SURVEY 8b).
//
// Layout: planes [N*K][H][W][C] channel-last (K planes per sample set, e.g. 3), grid [N*K][M][2] normalised (x -> W,
// y -> H), out [N][M][OS] point-major: OS = C and the K planes are SUMMED (interpolate_feat v1), or OS = K*C and they are
// concatenated (v2).  HBM-bound: M*(8K + 4*OS) bytes + the touched texels; no contraction -> no tensor cores.
// Threads: SEG = min(32, pow2 >= C/4) consecutive lanes share one point, lane l takes the 16-byte channel chunks
// l, l+SEG, ...: one tap of one point is one contiguous 4C-byte read, the output row one contiguous write, and the two
// grid-gradient components are reduced over the SEG lanes with shuffles.
#pragma once
#include "tt_device.cuh"

namespace tt {

constexpr int SMP_MAXK = 4;

struct TapsHW {
    int o[4];                    // texel index y*W+x of nw, ne, sw, se; -1 when out of bounds
    float w[4];                  // bilinear weights
    float dwx[4], dwy[4];        // d w / d ix, d w / d iy
};
__device__ __forceinline__ TapsHW make_taps_hw(float gx, float gy, int H, int W) {
    TapsHW t;
    // ATen grid_sampler_unnormalize (align_corners = False): ((g + 1) * size - 1) / 2, no FMA contraction
    const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 0.5f);
    const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 0.5f);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wx0 = __fsub_rn(__fadd_rn(x0f, 1.f), ix), wx1 = __fsub_rn(ix, x0f);
    const float wy0 = __fsub_rn(__fadd_rn(y0f, 1.f), iy), wy1 = __fsub_rn(iy, y0f);
    t.w[0] = __fmul_rn(wx0, wy0); t.w[1] = __fmul_rn(wx1, wy0); t.w[2] = __fmul_rn(wx0, wy1); t.w[3] = __fmul_rn(wx1, wy1);
    t.dwx[0] = -wy0; t.dwx[1] = wy0; t.dwx[2] = -wy1; t.dwx[3] = wy1;
    t.dwy[0] = -wx0; t.dwy[1] = -wx1; t.dwy[2] = wx0; t.dwy[3] = wx1;
    const int x0 = (int)fminf(fmaxf(x0f, -2.f), (float)W + 2.f), y0 = (int)fminf(fmaxf(y0f, -2.f), (float)H + 2.f);
    const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W;
    const bool ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
    t.o[0] = (xa && ya) ? y0 * W + x0 : -1;
    t.o[1] = (xb && ya) ? y0 * W + x0 + 1 : -1;
    t.o[2] = (xa && yb) ? (y0 + 1) * W + x0 : -1;
    t.o[3] = (xb && yb) ? (y0 + 1) * W + x0 + 1 : -1;
    return t;
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float seg_sum(float v, int seg) {
    for (int off = seg >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

struct SampleArgs {
    const float* planes; const float* grid;
    int N, K, C, H, W; int64_t M; int concat; int seg;
    int i32;                     // N * M * max(seg, C/4) + 256 < 2^32: thread and point indices split with 32-bit divisions
};
// thread -> (point, lane of the point's segment), point -> (sample set, point of the set)
__device__ __forceinline__ void sample_split(const SampleArgs& a, int64_t& pt0, int& l) {
    if (a.i32) { const uint32_t t = blockIdx.x * 256u + threadIdx.x, p = t / (uint32_t)a.seg; pt0 = p; l = (int)(t - p * (uint32_t)a.seg); }
    else { const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; pt0 = t / a.seg; l = (int)(t % a.seg); }
}
__device__ __forceinline__ void sample_set(const SampleArgs& a, int64_t pt, int64_t& n, int64_t& m) {
    if (a.i32) { const uint32_t n32 = (uint32_t)pt / (uint32_t)a.M; n = n32; m = (uint32_t)pt - n32 * (uint32_t)a.M; }
    else { n = pt / a.M; m = pt - n * a.M; }
}

// I32: the item index (point, chunk) fits 32 bits (the usual case): 32-bit divisions.  All 4K taps of an item are loaded
// unconditionally (an out-of-bounds tap reads texel 0 with weight 0) so that they are in flight together, and the output
// row is written with streaming stores: it is never read again and must not evict the planes from L2.
template <int K, bool I32>
__global__ void __launch_bounds__(256) k_sample_fwd(SampleArgs a, float* __restrict__ out) {
    const int U = a.C >> 2, OS = a.concat ? K * a.C : a.C;
    int64_t pt, n, m; int ch;
    if (I32) {
        const uint32_t t = blockIdx.x * 256u + threadIdx.x;
        const uint32_t p32 = t / (uint32_t)U; ch = (int)(t - p32 * (uint32_t)U);
        if ((int64_t)p32 >= (int64_t)a.N * a.M) return;
        const uint32_t n32 = p32 / (uint32_t)a.M;
        pt = p32; n = n32; m = p32 - n32 * (uint32_t)a.M;
    } else {
        const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
        pt = t / U; ch = (int)(t - pt * U);
        if (pt >= (int64_t)a.N * a.M) return;
        n = pt / a.M; m = pt - n * a.M;
    }
    const size_t ps = (size_t)a.H * a.W * a.C;
    float4 v[K][4]; float w[K][4]; unsigned in_mask = 0u;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float2 g = __ldg(reinterpret_cast<const float2*>(a.grid + (((size_t)n * K + k) * a.M + m) * 2));
        const TapsHW tp = make_taps_hw(g.x, g.y, a.H, a.W);
        const float* base = a.planes + ((size_t)n * K + k) * ps + ch * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool in = tp.o[q] >= 0;
            in_mask |= in ? (1u << (k * 4 + q)) : 0u;
            w[k][q] = tp.w[q];
            v[k][q] = ldg4(base + (size_t)(in ? tp.o[q] : 0) * a.C);
        }
    }
    float* orow = out + (size_t)pt * OS;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if ((in_mask >> (k * 4 + q)) & 1u) {      // out-of-bounds taps were loaded from texel 0 and are dropped here
                s.x = fmaf(w[k][q], v[k][q].x, s.x); s.y = fmaf(w[k][q], v[k][q].y, s.y);
                s.z = fmaf(w[k][q], v[k][q].z, s.z); s.w = fmaf(w[k][q], v[k][q].w, s.w);
            }
        if (a.concat) __stcs(reinterpret_cast<float4*>(orow + k * a.C + ch * 4), s);
        else { acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w; }
    }
    if (!a.concat) __stcs(reinterpret_cast<float4*>(orow + ch * 4), acc);
}

// d/d planes (accumulated, nullable) and d/d grid (written, nullable) for the upstream gradient g_out [N][M][OS]
__global__ void __launch_bounds__(256) k_sample_bwd(SampleArgs a, const float* __restrict__ g_out,
                                                   float* __restrict__ g_planes, float* __restrict__ g_grid) {
    int64_t pt0; int l;
    sample_split(a, pt0, l);
    const bool act = pt0 < (int64_t)a.N * a.M;
    const int64_t pt = act ? pt0 : 0;
    int64_t n, m;
    sample_set(a, pt, n, m);
    const int U = a.C >> 2, OS = a.concat ? a.K * a.C : a.C;
    const size_t ps = (size_t)a.H * a.W * a.C;
    for (int k = 0; k < a.K; ++k) {
        const size_t gi = (((size_t)n * a.K + k) * a.M + m) * 2;
        const TapsHW tp = make_taps_hw(a.grid[gi], a.grid[gi + 1], a.H, a.W);
        float A[4] = {0.f, 0.f, 0.f, 0.f};
        if (act)
            for (int ch = l; ch < U; ch += a.seg) {
                const float4 go = ldg4(g_out + (size_t)pt * OS + (a.concat ? k * a.C : 0) + ch * 4);
                const size_t off = ((size_t)n * a.K + k) * ps + ch * 4;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (tp.o[q] >= 0) {
                        const size_t at = off + (size_t)tp.o[q] * a.C; const float w = tp.w[q];
                        if (g_planes) red_add4(g_planes + at, make_float4(go.x * w, go.y * w, go.z * w, go.w * w));
                        if (g_grid) A[q] += dot4(go, ldg4(a.planes + at));
                    }
            }
        if (g_grid) {
            float gx = 0.f, gy = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) { gx = fmaf(A[q], tp.dwx[q], gx); gy = fmaf(A[q], tp.dwy[q], gy); }
            gx = seg_sum(gx, a.seg); gy = seg_sum(gy, a.seg);
            if (act && l == 0) { g_grid[gi] = gx * (0.5f * (float)a.W); g_grid[gi + 1] = gy * (0.5f * (float)a.H); }
        }
    }
}

// Backward of the backward (gridsample_cuda.cu:87-209).  Inputs: gg_planes = gradient arriving at d/d planes
// (channel-last, nullable), gg_grid = gradient arriving at d/d grid ([N*K][M][2], nullable).  Outputs (nullable):
//   gg_out   [N][M][OS]   d/d g_out    = Σ_t w_t gg_planes[o_t] + Σ_t ẇ_t planes[o_t],  ẇ_t = dw_t/d(ix,iy) · (gg_grid ⊙ size/2)
//   g_planes (accumulated) d/d planes  = ẇ_t · g_out
//   g_grid   (written)     d/d grid    = size/2 ⊙ [Σ_t dw_t B_t + cross-term Σ_t c_t A_t],  A_t = g_out·planes[o_t],
//                                        B_t = g_out·gg_planes[o_t], c = (+1, -1, -1, +1) (the only non-zero second
//                                        derivative of the bilinear weights is the mixed one)
template <int K>
__global__ void __launch_bounds__(256) k_sample_bwdbwd(SampleArgs a, const float* __restrict__ g_out,
                                                      const float* __restrict__ gg_planes, const float* __restrict__ gg_grid,
                                                      float* __restrict__ gg_out, float* __restrict__ g_planes,
                                                      float* __restrict__ g_grid) {
    int64_t pt0; int l;
    sample_split(a, pt0, l);
    const bool act = pt0 < (int64_t)a.N * a.M;
    const int64_t pt = act ? pt0 : 0;
    int64_t n, m;
    sample_set(a, pt, n, m);
    const int U = a.C >> 2, OS = a.concat ? K * a.C : a.C;
    const size_t ps = (size_t)a.H * a.W * a.C;
    const float sx = 0.5f * (float)a.W, sy = 0.5f * (float)a.H;
    const float cr[4] = {1.f, -1.f, -1.f, 1.f};
    TapsHW tp[K];
    float wd[K][4], vx[K], vy[K];
    float A[K][4], B[K][4];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const size_t gi = (((size_t)n * K + k) * a.M + m) * 2;
        tp[k] = make_taps_hw(a.grid[gi], a.grid[gi + 1], a.H, a.W);
        vx[k] = gg_grid ? gg_grid[gi] * sx : 0.f; vy[k] = gg_grid ? gg_grid[gi + 1] * sy : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) { wd[k][q] = tp[k].dwx[q] * vx[k] + tp[k].dwy[q] * vy[k]; A[k][q] = 0.f; B[k][q] = 0.f; }
    }
    if (act)
        for (int ch = l; ch < U; ch += a.seg) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float4 go = ldg4(g_out + (size_t)pt * OS + (a.concat ? k * a.C : 0) + ch * 4);
                const size_t off = ((size_t)n * K + k) * ps + ch * 4;
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (tp[k].o[q] >= 0) {
                        const size_t at = off + (size_t)tp[k].o[q] * a.C;
                        const float4 I = ldg4(a.planes + at);
                        const float w = tp[k].w[q], d = wd[k][q];
                        s.x = fmaf(d, I.x, s.x); s.y = fmaf(d, I.y, s.y); s.z = fmaf(d, I.z, s.z); s.w = fmaf(d, I.w, s.w);
                        A[k][q] += dot4(go, I);
                        if (gg_planes) {
                            const float4 J = ldg4(gg_planes + at);
                            s.x = fmaf(w, J.x, s.x); s.y = fmaf(w, J.y, s.y); s.z = fmaf(w, J.z, s.z); s.w = fmaf(w, J.w, s.w);
                            B[k][q] += dot4(go, J);
                        }
                        if (g_planes && d != 0.f) red_add4(g_planes + at, make_float4(go.x * d, go.y * d, go.z * d, go.w * d));
                    }
                if (a.concat) { if (gg_out) *reinterpret_cast<float4*>(gg_out + (size_t)pt * OS + k * a.C + ch * 4) = s; }
                else { acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w; }
            }
            if (!a.concat && gg_out) *reinterpret_cast<float4*>(gg_out + (size_t)pt * OS + ch * 4) = acc;
        }
    if (g_grid)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float gx = 0.f, gy = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                gx += tp[k].dwx[q] * B[k][q] + vy[k] * cr[q] * A[k][q];
                gy += tp[k].dwy[q] * B[k][q] + vx[k] * cr[q] * A[k][q];
            }
            gx = seg_sum(gx, a.seg); gy = seg_sum(gy, a.seg);
            const size_t gi = (((size_t)n * K + k) * a.M + m) * 2;
            if (act && l == 0) { g_grid[gi] = gx * sx; g_grid[gi + 1] = gy * sy; }
        }
}

// batched transpose dst[b][c][r] = src[b][r][c] (NCHW <-> channel-last), 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) k_transpose(const float* __restrict__ src, float* __restrict__ dst, int64_t rows,
                                                  int64_t cols) {
    __shared__ float tile[32][33];
    const int64_t b = blockIdx.z, c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* s = src + (size_t)b * rows * cols;
    float* d = dst + (size_t)b * rows * cols;
    for (int j = ty; j < 32; j += 8)
        if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = s[(size_t)(r0 + j) * cols + c0 + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (c0 + j < cols && r0 + tx < rows) d[(size_t)(c0 + j) * rows + r0 + tx] = tile[tx][j];
}

}  // namespace tt
