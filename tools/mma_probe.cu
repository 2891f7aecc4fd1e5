// Probe: throughput of legacy warp-level mma.sync (tf32 / bf16) and FFMA on sm_100a.  Build + run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_tf32(float* out, int iters) {
    float c[8][4] = {};
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f810000u, 0x3f820000u, 0x3f830000u}, b[2] = {0x3f840000u, 0x3f850000u + threadIdx.x};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16(float* out, int iters) {
    float c[8][4] = {};
    unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f813f81u, 0x3f823f82u, 0x3f833f83u}, b[2] = {0x3f843f84u, 0x3f853f85u + threadIdx.x};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters) {
    float c[32]; for (int j = 0; j < 32; ++j) c[j] = threadIdx.x * 1e-3f + j;
    float a = 1.0001f + threadIdx.x * 1e-7f, b = 0.9999f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) c[j] = fmaf(c[j], a, b);
    }
    float s = 0; for (int j = 0; j < 32; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K> double run(K k, int blocks, int threads, int iters, double flop_per_thread_iter, const char* name) {
    float* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    k<<<blocks, threads>>>(out, 16);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaDeviceSynchronize();
    cudaEventRecord(a); k<<<blocks, threads>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double tf = flop_per_thread_iter * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    printf("%-28s blocks=%d thr=%d  %.3f ms  %.1f TFLOP/s  (err=%s)\n", name, blocks, threads, ms, tf, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); return tf;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const int it = 20000;
    for (int thr : {128, 256, 512, 1024}) {
        int blocks = p.multiProcessorCount * (2048 / thr >= 2 ? 2 : 1);
        // per warp per iter: 8 mma of 16*8*8*2 flop -> per thread /32
        run(k_tf32, blocks, thr, it, 8.0 * 16 * 8 * 8 * 2 / 32, "mma.sync m16n8k8 tf32");
        run(k_bf16, blocks, thr, it, 8.0 * 16 * 8 * 16 * 2 / 32, "mma.sync m16n8k16 bf16");
        run(k_ffma, blocks, thr, it, 32.0 * 2, "ffma");
    }
    return 0;
}
