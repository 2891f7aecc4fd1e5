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

struct uint3_e { unsigned x = 0, y = 0, z = 0; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace tt_emul {
inline thread_local uint3_e threadIdx_, blockIdx_;
inline dim3 blockDim_, gridDim_;
inline float* smem_ = nullptr;
inline std::barrier<>* bar_ = nullptr;
inline std::atomic<int> or_flag_{0};
}  // namespace tt_emul
#define threadIdx tt_emul::threadIdx_
#define blockIdx tt_emul::blockIdx_
#define blockDim tt_emul::blockDim_
#define gridDim tt_emul::gridDim_

inline void __syncthreads() { tt_emul::bar_->arrive_and_wait(); }
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
using std::max;
using std::min;

// ---- runtime stubs -----------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
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
