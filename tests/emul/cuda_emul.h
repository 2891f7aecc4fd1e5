// TEST TOOLING ONLY — never part of the product.
//
// A minimal host emulation of the CUDA execution model, so that the *same kernel sources*
// (triplaneturbo_b200/csrc/*.cu, compiled with g++ -DTT_EMUL) can be executed on the CPU-only build
// container to debug kernel logic before spending GPU time.  The resulting tests/emul/libtt_emul.so is loaded
// exclusively by tests/test_emul_kernels.py through raw ctypes + numpy; the package triplaneturbo_b200 never
// loads it and has no CPU path.
//
// Model: one OS thread per CUDA thread of a block; blocks run one after another; __syncthreads() is a
// std::barrier; shared memory is one heap buffer per launch; atomics use std::atomic_ref.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static          /* one block at a time: a function-local static is block-shared */

struct uint3_e { unsigned x = 0, y = 0, z = 0; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct alignas(8) float2 { float x, y; };
inline void __stcs(float4* p, float4 v) { *p = v; }
struct alignas(8) int2 { int x, y; };
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

namespace tt_emul {
inline thread_local uint3_e threadIdx_, blockIdx_;
inline dim3 blockDim_, gridDim_;
inline float* smem_ = nullptr;
inline std::barrier<>* bar_ = nullptr;
inline std::atomic<int> or_flag_{0};
inline std::barrier<>* gbar_[8] = {};
inline std::barrier<>* wbar_[32] = {};
inline uint32_t shfl_buf_[32][32];
inline float tmem_[128][512];

// ---- tcgen05 emulation (see triplaneturbo_b200/csrc/tt_umma.cuh) ------------------------------------------
inline uint32_t smem_off(const void* p) { return (uint32_t)((const char*)p - (const char*)smem_); }
inline void group_sync(int g) { gbar_[g]->arrive_and_wait(); }
// mbarrier: bits [0,20) completed phases, [20,40) arrivals of the current phase, [40,60) expected arrivals per phase
inline void mbar_init(uint64_t* bar, uint32_t count = 1) { *bar = (uint64_t)count << 40; }
inline void mbar_arrive(uint32_t off) {
    std::atomic_ref<uint64_t> r(*(uint64_t*)((char*)smem_ + off));
    uint64_t old = r.load();
    for (;;) {
        const uint64_t expect = old >> 40, arr = ((old >> 20) & 0xfffff) + 1, ph = old & 0xfffff;
        const uint64_t nw = arr >= expect ? ((expect << 40) | ((ph + 1) & 0xfffff)) : ((expect << 40) | (arr << 20) | ph);
        if (r.compare_exchange_weak(old, nw)) break;
    }
}
inline void mbar_wait(uint32_t off, uint32_t parity) {
    std::atomic_ref<uint64_t> r(*(uint64_t*)((char*)smem_ + off));
    while ((r.load() & 1u) == parity) std::this_thread::yield();
}
inline int cas(int* addr, int expect, int desired) {
    std::atomic_ref<int> r(*addr); int e = expect; r.compare_exchange_strong(e, desired); return e;
}
inline float tf32(float x) { uint32_t u; std::memcpy(&u, &x, 4); u &= 0xffffe000u; std::memcpy(&x, &u, 4); return x; }
inline void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
    const int lane = (int)(addr >> 16) + (int)(threadIdx_.x & 31), col = (int)(addr & 0xffff);
    for (int j = 0; j < 8; ++j) std::memcpy(&tmem_[lane][col + j], &v[j], 4);
}
inline void tmem_ld8(uint32_t addr, uint32_t (&v)[8]) {
    const int lane = (int)(addr >> 16) + (int)(threadIdx_.x & 31), col = (int)(addr & 0xffff);
    for (int j = 0; j < 8; ++j) std::memcpy(&v[j], &tmem_[lane][col + j], 4);
}
// D[128][N] (+)= A[128][K] * B[N][K]^T ; A hi / lo / D at explicit columns of the group's region
inline void umma_ex(uint32_t tmem, uint32_t col_hi, uint32_t col_lo, uint32_t col_d, uint32_t bhi, uint32_t blo, uint32_t sbo,
                    int N, int K, bool acc, int passes) {
    const int c0 = (int)(tmem & 0xffff);
    auto B = [&](uint32_t base, int n, int k) {
        const uint32_t off = base + (n & 7) * 16 + (n >> 3) * sbo + (k >> 2) * 128 + (k & 3) * 4;
        return tf32(*(const float*)((const char*)smem_ + off));
    };
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float d = acc ? tmem_[m][c0 + col_d + n] : 0.f;
            for (int k0 = 0; k0 < K; k0 += 8) {
                float s1 = 0.f, s2 = 0.f, s3 = 0.f;
                for (int k = k0; k < k0 + 8; ++k) {
                    const float ah = tf32(tmem_[m][c0 + col_hi + k]);
                    s3 += ah * B(bhi, n, k);
                    if (passes == 3) { const float al = tf32(tmem_[m][c0 + col_lo + k]); s1 += al * B(bhi, n, k); s2 += ah * B(blo, n, k); }
                }
                d = ((d + s1) + s2) + s3;
            }
            tmem_[m][c0 + col_d + n] = d;
        }
}
inline void umma(uint32_t tmem, uint32_t a_col0, uint32_t bhi, uint32_t blo, uint32_t sbo, int N, int K, bool acc, int passes) {
    umma_ex(tmem, a_col0, 64 + a_col0, 128, bhi, blo, sbo, N, K, acc, passes);
}
// G[128][N] (+)= At * Bt^T over 128 points; operand tiles with LBO 144 / SBO 4608 (tt_umma.cuh wg_off)
inline void umma_ss(uint32_t d_addr, uint32_t a, uint32_t b, int N, bool acc) {
    const int c0 = (int)(d_addr & 0xffff);
    auto E = [&](uint32_t base, int r, int p) {
        const uint32_t off = base + (r & 7) * 16 + (r >> 3) * 4608 + (p >> 2) * 144 + (p & 3) * 4;
        return tf32(*(const float*)((const char*)smem_ + off));
    };
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float d = acc ? tmem_[m][c0 + n] : 0.f;
            for (int p = 0; p < 128; ++p) d += E(a, m, p) * E(b, n, p);
            tmem_[m][c0 + n] = d;
        }
}
}  // namespace tt_emul
#define threadIdx tt_emul::threadIdx_
#define blockIdx tt_emul::blockIdx_
#define blockDim tt_emul::blockDim_
#define gridDim tt_emul::gridDim_

inline void __syncthreads() { tt_emul::bar_->arrive_and_wait(); }
// warp collectives (all 32 lanes must participate)
inline uint32_t tt_emul_shfl_u32(uint32_t v, int src_lane) {
    const int w = (int)(threadIdx.x >> 5), l = (int)(threadIdx.x & 31);
    tt_emul::shfl_buf_[w][l] = v;
    tt_emul::wbar_[w]->arrive_and_wait();
    const uint32_t r = tt_emul::shfl_buf_[w][src_lane & 31];
    tt_emul::wbar_[w]->arrive_and_wait();
    return r;
}
inline float __shfl_xor_sync(unsigned, float v, int m) {
    uint32_t u; std::memcpy(&u, &v, 4); u = tt_emul_shfl_u32(u, (int)(threadIdx.x & 31) ^ m); std::memcpy(&v, &u, 4); return v;
}
inline int __shfl_sync(unsigned, int v, int lane) { return (int)tt_emul_shfl_u32((uint32_t)v, lane); }
inline float __shfl_sync(unsigned, float v, int lane) {
    uint32_t u; std::memcpy(&u, &v, 4); u = tt_emul_shfl_u32(u, lane); std::memcpy(&v, &u, 4); return v;
}
inline int __shfl_up_sync(unsigned, int v, int delta) {
    const int l = (int)(threadIdx.x & 31);
    return (int)tt_emul_shfl_u32((uint32_t)v, l >= delta ? l - delta : l);
}
inline float __shfl_up_sync(unsigned, float v, int delta) {
    const int l = (int)(threadIdx.x & 31);
    uint32_t u; std::memcpy(&u, &v, 4); u = tt_emul_shfl_u32(u, l >= delta ? l - delta : l); std::memcpy(&v, &u, 4); return v;
}
inline unsigned __ballot_sync(unsigned, bool pred) {
    const int w = (int)(threadIdx.x >> 5), l = (int)(threadIdx.x & 31);
    tt_emul::shfl_buf_[w][l] = pred ? 1u : 0u;
    tt_emul::wbar_[w]->arrive_and_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (tt_emul::shfl_buf_[w][i] & 1u) << i;
    tt_emul::wbar_[w]->arrive_and_wait();
    return m;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline void __syncwarp() { tt_emul::wbar_[threadIdx.x >> 5]->arrive_and_wait(); }
inline int __syncthreads_or(int pred) {
    if (pred) tt_emul::or_flag_.store(1);
    __syncthreads();
    const int r = tt_emul::or_flag_.load();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) tt_emul::or_flag_.store(0);
    __syncthreads();
    return r;
}
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float atomicAdd(float* addr, float v) {
    std::atomic_ref<float> r(*addr);
    float old = r.load();
    while (!r.compare_exchange_weak(old, old + v)) {}
    return old;
}
inline int atomicAdd(int* addr, int v) { std::atomic_ref<int> r(*addr); return r.fetch_add(v); }
inline int atomicCAS(int* addr, int expect, int desired) { return tt_emul::cas(addr, expect, desired); }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline void __trap() { std::abort(); }
using std::max;
using std::min;

// ---- runtime stubs -----------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
enum { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return cudaSuccess; }
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }

namespace tt_emul {
template <typename K, typename... A>
void launch(K kernel, dim3 grid, dim3 block, size_t smem_bytes, A... args) {
    const unsigned nthreads = block.x * block.y * block.z;
    std::vector<float> smem((smem_bytes + 64) / sizeof(float) + 16);
    void* p = smem.data(); size_t space = smem.size() * sizeof(float);
    smem_ = static_cast<float*>(std::align(16, smem_bytes, p, space));
    std::barrier<> bar(nthreads);
    bar_ = &bar; blockDim_ = block; gridDim_ = grid;
    std::vector<std::unique_ptr<std::barrier<>>> warps;
    for (unsigned w = 0; w < (nthreads + 31) / 32 && w < 32; ++w) {
        warps.emplace_back(new std::barrier<>(std::min(32u, nthreads - w * 32)));
        wbar_[w] = warps.back().get();
    }
    std::vector<std::unique_ptr<std::barrier<>>> groups;
    for (unsigned g = 0; g < nthreads / 128 && g < 8; ++g) {
        groups.emplace_back(new std::barrier<>(128));
        gbar_[g] = groups.back().get();
    }
    auto worker = [&](unsigned t) {
        threadIdx_.x = t % block.x; threadIdx_.y = (t / block.x) % block.y; threadIdx_.z = t / (block.x * block.y);
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    blockIdx_.x = bx; blockIdx_.y = by; blockIdx_.z = bz;
                    kernel(args...);
                    bar.arrive_and_wait();      // block boundary: shared memory is reused
                }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
    bar_ = nullptr; smem_ = nullptr;
}
}  // namespace tt_emul
#define TT_LAUNCH(kernel, grid, block, smem, stream, ...) tt_emul::launch(kernel, dim3(grid), dim3(block), smem, __VA_ARGS__)
#define TT_SHARED(name) float* name = tt_emul::smem_
