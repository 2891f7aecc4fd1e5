// Probe: tcgen05.mma kind::tf32 with A in TMEM (written by tcgen05.st, one row per thread), B in shared memory
// (K-major, no swizzle), D in TMEM, 3xTF32 split.  Validates descriptor / layout conventions used by the kernels.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu && ./umma_probe
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
    return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr) : "memory");
}

// mode 0: 1xTF32, mode 1: 3xTF32
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int K, int N, int mode, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* Bhi = reinterpret_cast<float*>(smem);
    float* Blo = Bhi + N * K;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t LBO = 128, SBO = (uint32_t)(K / 4) * 128;
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const float w = B[i];
        const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
        const int off = ((n % 8) * 16 + (n / 8) * SBO + (k / 4) * LBO + (k % 4) * 4) / 4;
        Bhi[off] = hi;
        Blo[off] = w - hi;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t colAhi = 0, colAlo = 64, colD = 128;
    // A row of this thread -> TMEM
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t hi[8], lo[8];
        for (int j = 0; j < 8; ++j) {
            const float a = A[tid * K + k0 + j];
            const float h = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
            hi[j] = __float_as_uint(h);
            lo[j] = __float_as_uint(a - h);
        }
        tmem_st8(tb + lane_base + colAhi + k0, hi);
        tmem_st8(tb + lane_base + colAlo + k0, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint32_t bh = smem_u32(Bhi), bl = smem_u32(Blo);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t dh = make_desc(bh + ks * 2 * LBO, LBO, SBO), dl = make_desc(bl + ks * 2 * LBO, LBO, SBO);
            if (mode == 1) {
                umma_tf32_ts(tb + colD, tb + colAlo + ks * 8, dh, idesc, ks > 0);
                umma_tf32_ts(tb + colD, tb + colAhi + ks * 8, dl, idesc, 1);
                umma_tf32_ts(tb + colD, tb + colAhi + ks * 8, dh, idesc, 1);
            } else {
                umma_tf32_ts(tb + colD, tb + colAhi + ks * 8, dh, idesc, ks > 0);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // wait (bounded)
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    if (!done) { if (tid == 0) *status = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done) {
        for (int n0 = 0; n0 < N; n0 += 8) {
            uint32_t v[8];
            tmem_ld8(tb + lane_base + colD + n0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) D[tid * N + n0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}

int main() {
    int cases[][2] = {{64, 64}, {40, 64}, {64, 48}, {8, 16}, {120, 64}};
    for (auto& c : cases) {
        const int K = c[0], N = c[1];
        if (K > 64) continue;   // A columns limited to 64 in this probe
        std::vector<float> A(128 * K), B(N * K), D(128 * N);
        srand(1);
        for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
        for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f);
        float *dA, *dB, *dD; int* dS;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        for (int mode = 0; mode < 2; ++mode) {
            cudaMemset(dD, 0, D.size() * 4); cudaMemset(dS, 0, 4);
            probe<<<1, 128, 2 * N * K * 4 + 1024>>>(dA, dB, dD, K, N, mode, dS);
            cudaError_t e = cudaDeviceSynchronize();
            int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0, maxref = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double r = 0;
                    for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
                    maxerr = fmax(maxerr, fabs(r - D[m * N + n])); maxref = fmax(maxref, fabs(r));
                }
            printf("K=%3d N=%3d mode=%s  cuda=%s status=%d  max|err|=%.3e  max|ref|=%.3f  rel=%.3e\n", K, N,
                   mode ? "3xTF32" : "1xTF32", cudaGetErrorString(e), st, maxerr, maxref, maxerr / maxref);
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
    }
    return 0;
}
