// Probe: throughput of the plane-gradient scatter patterns on sm_100a.
//   texel = one contiguous C*4-byte run of a channel-last gradient plane; a tile scatters NTAP texels.
//   mode 0: red.global.add.v4.f32, U = C/4 consecutive lanes cover one texel (the kernels' coop_scatter pattern)
//   mode 1: red.global.add.f32, C consecutive lanes cover one texel
//   mode 2: cp.reduce.async.bulk.global.shared::cta.add.f32 of C*4 bytes, one thread per texel (TMA reduce)
//   mode 3: plain st.global.v4.f32 (no reduction: upper bound of the store path)
//   loc = number of consecutive taps that hit the same texel (1 = all distinct / random, 8 = runs of 8)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_probe red_probe.cu && ./red_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int C>
__global__ void __launch_bounds__(256) k_probe(float* g, uint32_t n_texels, int iters, int mode, int loc) {
    constexpr int U = C / 4, SP = C + 4;
    extern __shared__ __align__(16) float stage[];      // [blockDim][SP]
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < nt * SP; i += nt) stage[i] = 1.0f;
    __syncthreads();
    const uint32_t cta_seed = blockIdx.x * 7919u;
    for (int it = 0; it < iters; ++it) {
        // one "tile": nt points x 12 taps
        if (mode == 0 || mode == 3) {
#pragma unroll 1
            for (int j = 0; j < U; ++j) {
                const int item = tid + nt * j;
                const int pt = item / U, ch = item - pt * U;
                const float4 v = *reinterpret_cast<const float4*>(stage + pt * SP + ch * 4);
#pragma unroll
                for (int t = 0; t < 12; ++t) {
                    const uint32_t key = (cta_seed + it) * 4096u + (uint32_t)(pt / loc) * 16u + t;
                    const uint32_t tex = hash32(key) % n_texels;
                    float* a = g + (size_t)tex * C + ch * 4;
                    if (mode == 0)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                    else
                        asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
            }
        } else if (mode == 1) {
#pragma unroll 1
            for (int j = 0; j < C; ++j) {
                const int item = tid + nt * j;
                const int pt = item / C, ch = item - pt * C;
                const float v = stage[pt * SP + ch];
#pragma unroll
                for (int t = 0; t < 12; ++t) {
                    const uint32_t key = (cta_seed + it) * 4096u + (uint32_t)(pt / loc) * 16u + t;
                    const uint32_t tex = hash32(key) % n_texels;
                    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g + (size_t)tex * C + ch), "f"(v) : "memory");
                }
            }
        } else {
            const int pt = tid;
#pragma unroll
            for (int t = 0; t < 12; ++t) {
                const uint32_t key = (cta_seed + it) * 4096u + (uint32_t)(pt / loc) * 16u + t;
                const uint32_t tex = hash32(key) % n_texels;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             ::"l"(g + (size_t)tex * C), "r"(smem_u32(stage + pt * SP)), "r"(C * 4) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
}

template <int C>
static void run(float* g, uint32_t n_texels, int mode, int loc, int threads, int ctas_per_sm, int sms) {
    const int iters = 64;
    const size_t smem = (size_t)threads * (C + 4) * 4;
    cudaFuncSetAttribute(k_probe<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = sms * ctas_per_sm;
    k_probe<C><<<grid, threads, smem>>>(g, n_texels, 4, mode, loc);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_probe<C><<<grid, threads, smem>>>(g, n_texels, iters, mode, loc);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f; cudaEventElapsedTime(&ms, a, b);
    const double texels = (double)grid * iters * threads * 12.0;
    const char* names[4] = {"red.v4.f32", "red.f32", "cp.reduce.bulk", "st.v4.f32"};
    printf("C=%d %-15s loc=%-2d thr=%-4d cta/sm=%d  %8.3f ms  %7.2f Gtexel/s  %8.1f Gfloat/s  %7.1f GB/s  (%s)\n", C,
           names[mode], loc, threads, ctas_per_sm, ms, texels / ms * 1e-6, texels * C / ms * 1e-6,
           texels * C * 4 / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs=%d\n", p.name, p.multiProcessorCount);
    const int C = 40; const uint32_t n_texels = 4u * 3u * 256u * 256u;      // geometry planes of config 2
    float* g; cudaMalloc(&g, (size_t)n_texels * C * 4); cudaMemset(g, 0, (size_t)n_texels * C * 4);
    const int sms = p.multiProcessorCount;
    for (int mode = 0; mode < 4; ++mode)
        for (int loc : {1, 4, 16})
            for (int cfg = 0; cfg < 3; ++cfg) {
                const int threads = cfg == 0 ? 128 : 256, cps = cfg == 2 ? 2 : 1;
                run<40>(g, n_texels, mode, loc, threads, cps, sms);
            }
    // small working set (fits L2 comfortably): 1 prompt, 1 plane
    printf("-- small working set (one plane)\n");
    for (int mode = 0; mode < 4; ++mode) run<40>(g, 256u * 256u, mode, 1, 256, 2, sms);
    printf("-- C=32\n");
    for (int mode = 0; mode < 4; ++mode) run<32>(g, n_texels, mode, 1, 256, 2, sms);
    cudaFree(g);
    return 0;
}
