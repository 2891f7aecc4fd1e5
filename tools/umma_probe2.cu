// Probe 2 (sm_100a): two questions the backward kernels depend on.
//   (1) MN-major B operand for kind::tf32: a weight tile stored ONCE as the K-major canonical tile of W[j][c]
//       (rows j, K = c) is re-read as the transposed operand (N = c, K = j) by swapping the roles of LBO / SBO and
//       setting the b_major bit of the instruction descriptor  ->  D[m][c] = sum_j A[m][j] W[j][c].
//   (2) accumulating MMAs into the SAME TMEM accumulator issued concurrently by two different threads of a CTA:
//       integer-valued operands make the fp32 sum exact, so a single lost / torn update shows up.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_probe2 umma_probe2.cu && ./umma_probe2
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t phase) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 24) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    return done != 0;
}

// ---- (1) transposed read of a K-major tile ----------------------------------------------------------------------
// W: [64][C] row-major in global.  A: [128][64].  D: [128][NP] (NP = C rounded up to 16; columns >= C are garbage).
__global__ void __launch_bounds__(128) probe_mn(const float* A, const float* W, float* D, int C, int NP, int variant, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* Wt = reinterpret_cast<float*>(smem);
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t LBO_K = 128, SBO_K = (uint32_t)(C / 4) * 128;     // K-major tile of W[j][c]: rows j, K = c
    for (int i = tid; i < 64 * C; i += 128) {
        const int j = i / C, c = i % C;
        Wt[((j % 8) * 16 + (j / 8) * SBO_K + (c / 4) * LBO_K + (c % 4) * 4) / 4] = __uint_as_float(__float_as_uint(W[i]) & 0xffffe000u);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int k0 = 0; k0 < 64; k0 += 8) {
        uint32_t hi[8];
        for (int j = 0; j < 8; ++j) hi[j] = __float_as_uint(A[tid * 64 + k0 + j]) & 0xffffe000u;
        tmem_st8(tb + lane_base + k0, hi);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc_tf32(128, NP, 1);
        // MN-major view: element (n = c, k = j) at (c%4)*4 + (j%8)*16 + (c/4)*LBO_K + (j/8)*SBO_K
        //   -> stride between 4-wide MN chunks = LBO_K (goes into the SBO field), stride between 8-deep K blocks = SBO_K
        for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(tb + 128, tb + ks * 8, variant == 0 ? make_desc(smem_u32(Wt) + ks * SBO_K, /*lbo=*/SBO_K, /*sbo=*/LBO_K)
                                                              : make_desc(smem_u32(Wt) + ks * SBO_K, /*lbo=*/LBO_K, /*sbo=*/SBO_K), idesc, ks > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    const bool done = mbar_wait(smem_u32(&mbar), 0);
    if (!done && tid == 0) *status = 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done)
        for (int n0 = 0; n0 < NP; n0 += 8) {
            uint32_t v[8];
            tmem_ld8(tb + lane_base + 128 + n0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) D[tid * NP + n0 + j] = __uint_as_float(v[j]);
        }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}

// ---- (2) two issuing threads, one accumulator ---------------------------------------------------------------------
// 256 threads = two groups; thread 0 and thread 128 each issue `reps` accumulating MMAs (K = 8, N = 64) into the same
// 64 TMEM columns.  A_g[m][k] = ((m + k + g) % 3), B[n][k] = ((n + 2k) % 4): expected D[m][n] = reps * sum_g sum_k A_g B.
__global__ void __launch_bounds__(256) probe_xthread(float* D, int reps, int* status) {
    __shared__ __align__(1024) float Bt[64 * 8];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, group = tid >> 7, tg = tid & 127;
    for (int i = tid; i < 64 * 8; i += 256) {
        const int n = i / 8, k = i % 8;
        Bt[((n % 8) * 16 + (n / 8) * 256 + (k / 4) * 128 + (k % 4) * 4) / 4] = (float)((n + 2 * k) % 4);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t a[8], z[8];
    for (int k = 0; k < 8; ++k) { a[k] = __float_as_uint((float)((tg + k + group) % 3)); z[k] = 0u; }
    tmem_st8(tb + lane_base + group * 8, a);                 // A_g at columns [8g, 8g+8)
    if (group == 0) for (int n0 = 0; n0 < 64; n0 += 8) tmem_st8(tb + lane_base + 128 + n0, z);   // D = 0
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tg == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc_tf32(128, 64, 0);
        const uint64_t bd = make_desc(smem_u32(Bt), 128, 256);
        for (int r = 0; r < reps; ++r) umma_tf32_ts(tb + 128, tb + group * 8, bd, idesc, 1);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[group])) : "memory");
    }
    const bool d0 = mbar_wait(smem_u32(&mbar[0]), 0), d1 = mbar_wait(smem_u32(&mbar[1]), 0);
    if (!(d0 && d1) && tid == 0) *status = 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (group == 0 && d0 && d1)
        for (int n0 = 0; n0 < 64; n0 += 8) {
            uint32_t v[8];
            tmem_ld8(tb + lane_base + 128 + n0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) D[((size_t)blockIdx.x * 128 + tg) * 64 + n0 + j] = __uint_as_float(v[j]);
        }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}

int main() {
    for (int C : {40, 32, 64, 16}) {
        const int NP = (C + 15) / 16 * 16;
        std::vector<float> A(128 * 64), W(64 * C), D(128 * NP);
        srand(2);
        for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
        for (auto& x : W) x = (rand() / (float)RAND_MAX - 0.5f);
        float *dA, *dW, *dD; int* dS;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
      for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(dD, 0xff, D.size() * 4); cudaMemset(dS, 0, 4);
        probe_mn<<<1, 128, 64 * C * 4 + 4096>>>(dA, dW, dD, C, NP, variant, dS);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < 128; ++m)
            for (int c = 0; c < C; ++c) {
                double r = 0;
                for (int j = 0; j < 64; ++j) r += (double)A[m * 64 + j] * W[j * C + c];
                maxerr = fmax(maxerr, fabs(r - D[m * NP + c])); maxref = fmax(maxref, fabs(r));
            }
        printf("MN-major B  C=%2d NP=%2d variant=%d cuda=%s status=%d  max|err|=%.3e max|ref|=%.3f rel=%.3e\n", C, NP, variant,
               cudaGetErrorString(e), st, maxerr, maxref, maxerr / maxref);
        printf("   D[5][0..7] ="); for (int c = 0; c < 8; ++c) printf(" %8.4f", D[5 * NP + c]);
        printf("\n   ref        ="); for (int c = 0; c < 8; ++c) { double r = 0; for (int j = 0; j < 64; ++j) r += (double)A[5 * 64 + j] * W[j * C + c]; printf(" %8.4f", r); }
        printf("\n");
      }
        cudaFree(dA); cudaFree(dW); cudaFree(dD); cudaFree(dS);
    }
    {
        const int blocks = 148 * 4, reps = 4096;
        std::vector<float> D((size_t)blocks * 128 * 64);
        float* dD; int* dS;
        cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
        cudaMemset(dD, 0, D.size() * 4); cudaMemset(dS, 0, 4);
        probe_xthread<<<blocks, 256>>>(dD, reps, dS);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        long long bad = 0; double worst = 0;
        for (int b = 0; b < blocks; ++b)
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    double r = 0;
                    for (int g = 0; g < 2; ++g)
                        for (int k = 0; k < 8; ++k) r += (double)((m + k + g) % 3) * ((n + 2 * k) % 4);
                    r *= reps;
                    const double d = fabs(r - D[((size_t)b * 128 + m) * 64 + n]);
                    if (d != 0) ++bad;
                    worst = fmax(worst, d);
                }
        printf("two issuing threads, one accumulator: %d CTAs x %d MMAs/thread  cuda=%s status=%d  mismatches=%lld  worst=%.1f\n",
               blocks, reps, cudaGetErrorString(e), st, bad, worst);
        cudaFree(dD); cudaFree(dS);
    }
    return 0;
}
