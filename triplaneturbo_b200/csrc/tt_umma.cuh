// tcgen05 (5th-gen tensor core) primitives for the decoder-MLP layers, sm_100a.
//
// Usage model ("thread = point, tensor core = layer"): a group of 128 threads owns a tile of 128 sample points;
// thread i owns point i == TMEM lane i.  One decoder layer  Y[128][N] = X[128][K] · Wᵀ  is
//   1. every thread splits its activation row into tf32 hi/lo parts and writes them to TMEM (tcgen05.st),
//   2. one thread issues the K/8 x 3 tcgen05.mma (3xTF32: lo·hi + hi·lo + hi·hi, A from TMEM, W from shared memory
//      in the canonical K-major no-swizzle layout) and commits to an mbarrier,
//   3. every thread waits on the mbarrier and reads its output row back (tcgen05.ld).
// The 3xTF32 split keeps ~1e-6 relative accuracy (measured, tools/umma_probe.cu), which the NeuS alpha needs
// (inv_std = 100 multiplies SDF error inside a sigmoid).
//
// Under TT_EMUL (host emulation for tests/emul) the same API is implemented with plain loops.
#pragma once
#include <stdint.h>

namespace tt {

constexpr int TC_GROUP = 128;          // threads (= points = TMEM lanes) per tile
constexpr uint32_t TC_COL_AHI = 0;     // TMEM column offsets inside a group's 256-column region
constexpr uint32_t TC_COL_ALO = 64;
constexpr uint32_t TC_COL_D = 128;
constexpr uint32_t TC_COLS_PER_GROUP = 256;
constexpr uint32_t TC_LBO = 128;       // bytes between the 16-byte K-chunks of a core-matrix row group

// A weight matrix as MMA operand B: element (n, k) of the N x K (K-major) tile, tf32 hi and lo parts.
struct BTile {
    uint32_t hi, lo;     // shared-memory byte addresses (emulation: byte offsets into the block's smem)
    uint32_t sbo;        // bytes between 8-row groups
    int N, K;
};
__host__ __device__ inline int btile_floats(int N, int K) { return N * K; }
__device__ __forceinline__ int btile_off(int n, int k, int K) {     // float offset of element (n,k)
    return ((n & 7) * 16 + (n >> 3) * ((K >> 2) * 128) + (k >> 2) * 128 + (k & 3) * 4) >> 2;
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// round-to-nearest tf32 (single-pass operands: gradients), ties away from zero.  (Two full-rate integer ops;
// cvt.rna.tf32.f32 is one instruction but runs on the conversion pipe: measured 5 % slower in k_bwd_geo_tc.)
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

// Operand tiles of the weight-gradient contractions (K = the tile's 128 points): element (row r, point p) of an
// [R rows][128 points] K-major tile.  The 16-byte K-chunks are 144 B apart so that the 32 lanes of a warp (32
// consecutive points, same row) write 32 different banks.
constexpr uint32_t WG_LBO = 144, WG_SBO = 32 * 144;
__device__ __forceinline__ int wg_off(int r, int p) {
    return (int)(((uint32_t)(r & 7) * 16 + (uint32_t)(r >> 3) * WG_SBO + (uint32_t)(p >> 2) * WG_LBO + (uint32_t)(p & 3) * 4) >> 2);
}
__host__ __device__ constexpr int wg_tile_floats(int rows) { return (rows / 8) * (int)(WG_SBO / 4); }

struct Umma {
    uint32_t tmem;        // TMEM address of column 0 of this group's region (lane field 0)
    uint32_t lane_base;   // (warp % 4) * 32 << 16 : the TMEM lanes this warp may touch
    uint32_t mbar;        // shared address of the group's mbarrier
    uint32_t phase;
    int group;
};

#ifndef TT_EMUL
// ======================================================================================== device (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void group_sync(int group) {      // named barrier over the group's 128 threads
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(TC_GROUP) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void async_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc_warp(uint32_t* slot, uint32_t ncols) {     // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_warp(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// general mbarriers (producer / consumer hand-off between warp roles): `count` arrivals complete a phase
__device__ __forceinline__ void mbar_init_n(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {        // release at CTA scope: prior shared-memory writes are visible
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 24) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done) __trap();                                       // never hang the GPU: abort the kernel instead
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t umma_idesc_tf32(int N) {      // M = 128, fp32 accumulate, K-major A and B
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_issue(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
// D (+)= A · Bᵀ over K (multiple of 8); issued by ONE thread.  PASSES = 3: 3xTF32, 1: hi·hi only.
template <int PASSES>
__device__ __forceinline__ void umma_mma(const Umma& u, const BTile& b, int K, bool accumulate, uint32_t a_col0 = 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_tf32(b.N);
    const uint32_t d = u.tmem + TC_COL_D;
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t acc0 = (accumulate || ks > 0) ? 1u : 0u;
        const uint64_t dh = umma_desc(b.hi + ks * 2 * TC_LBO, TC_LBO, b.sbo);
        if (PASSES == 3) {
            const uint64_t dl = umma_desc(b.lo + ks * 2 * TC_LBO, TC_LBO, b.sbo);
            umma_issue(d, u.tmem + TC_COL_ALO + a_col0 + ks * 8, dh, idesc, acc0);
            umma_issue(d, u.tmem + TC_COL_AHI + a_col0 + ks * 8, dl, idesc, 1u);
            umma_issue(d, u.tmem + TC_COL_AHI + a_col0 + ks * 8, dh, idesc, 1u);
        } else {
            umma_issue(d, u.tmem + TC_COL_AHI + a_col0 + ks * 8, dh, idesc, acc0);
        }
    }
}
// same with explicit TMEM columns of the A hi / lo parts and of the accumulator (inside the group's region)
template <int PASSES>
__device__ __forceinline__ void umma_mma_ex(const Umma& u, const BTile& b, int K, bool accumulate, uint32_t col_hi,
                                            uint32_t col_lo, uint32_t col_d) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_tf32(b.N);
    const uint32_t d = u.tmem + col_d;
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t acc0 = (accumulate || ks > 0) ? 1u : 0u;
        const uint64_t dh = umma_desc(b.hi + ks * 2 * TC_LBO, TC_LBO, b.sbo);
        if (PASSES == 3) {
            const uint64_t dl = umma_desc(b.lo + ks * 2 * TC_LBO, TC_LBO, b.sbo);
            umma_issue(d, u.tmem + col_lo + ks * 8, dh, idesc, acc0);
            umma_issue(d, u.tmem + col_hi + ks * 8, dl, idesc, 1u);
            umma_issue(d, u.tmem + col_hi + ks * 8, dh, idesc, 1u);
        } else {
            umma_issue(d, u.tmem + col_hi + ks * 8, dh, idesc, acc0);
        }
    }
}
// G[128][N] (+)= At · Btᵀ over the 128 points of a tile (both operands in shared memory, 1xTF32); one thread.
__device__ __forceinline__ void umma_mma_ss(const Umma& u, uint32_t a_saddr, uint32_t b_saddr, int N, uint32_t d_col,
                                            bool accumulate) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_tf32(N);
    for (int ks = 0; ks < TC_GROUP / 8; ++ks) {
        const uint64_t da = umma_desc(a_saddr + ks * 2 * WG_LBO, WG_LBO, WG_SBO);
        const uint64_t db = umma_desc(b_saddr + ks * 2 * WG_LBO, WG_LBO, WG_SBO);
        const uint32_t acc = (accumulate || ks > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
            ::"r"(u.tmem + d_col), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
    }
}
__device__ __forceinline__ void umma_commit(const Umma& u) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(u.mbar) : "memory");
}
__device__ __forceinline__ void umma_wait(Umma& u) {          // all threads of the group
    uint32_t done = 0;
    for (int it = 0; it < (1 << 24) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(u.mbar), "r"(u.phase) : "memory");
    if (!done) __trap();                                       // never hang the GPU: abort the kernel instead
    u.phase ^= 1u;
    tc_fence_after();
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#else
// ======================================================================================== host emulation
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return tt_emul::smem_off(p); }
__device__ __forceinline__ void group_sync(int group) { tt_emul::group_sync(group); }
__device__ __forceinline__ void tc_fence_before() {}
__device__ __forceinline__ void tc_fence_after() {}
__device__ __forceinline__ void async_proxy_fence() {}
__device__ __forceinline__ void tmem_alloc_warp(uint32_t* slot, uint32_t) { *slot = 0; }
__device__ __forceinline__ void tmem_dealloc_warp(uint32_t, uint32_t) {}
__device__ __forceinline__ void mbar_init(uint64_t* bar) { tt_emul::mbar_init(bar); }
__device__ __forceinline__ void mbar_init_n(uint64_t* bar, uint32_t count) { tt_emul::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_init_fence() {}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { tt_emul::mbar_arrive(bar); }
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) { tt_emul::mbar_wait(bar, parity); }
template <int PASSES>
__device__ __forceinline__ void umma_mma(const Umma& u, const BTile& b, int K, bool accumulate, uint32_t a_col0 = 0) {
    tt_emul::umma(u.tmem, a_col0, b.hi, b.lo, b.sbo, b.N, K, accumulate, PASSES);
}
template <int PASSES>
__device__ __forceinline__ void umma_mma_ex(const Umma& u, const BTile& b, int K, bool accumulate, uint32_t col_hi,
                                            uint32_t col_lo, uint32_t col_d) {
    tt_emul::umma_ex(u.tmem, col_hi, col_lo, col_d, b.hi, b.lo, b.sbo, b.N, K, accumulate, PASSES);
}
__device__ __forceinline__ void umma_mma_ss(const Umma& u, uint32_t a_saddr, uint32_t b_saddr, int N, uint32_t d_col,
                                            bool accumulate) {
    tt_emul::umma_ss(u.tmem + d_col, a_saddr, b_saddr, N, accumulate);
}
__device__ __forceinline__ void umma_commit(const Umma& u) { tt_emul::mbar_arrive(u.mbar); }
__device__ __forceinline__ void umma_wait(Umma& u) { tt_emul::mbar_wait(u.mbar, u.phase); u.phase ^= 1u; }
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) { tt_emul::tmem_st8(addr, v); }
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&v)[8]) { tt_emul::tmem_ld8(addr, v); }
__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t* v) {
    for (int i = 0; i < 32; i += 8) { uint32_t t[8]; for (int j = 0; j < 8; ++j) t[j] = v[i + j]; tt_emul::tmem_st8(addr + i, t); }
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t* v) {
    for (int i = 0; i < 32; i += 8) { uint32_t t[8]; tt_emul::tmem_ld8(addr + i, t); for (int j = 0; j < 8; ++j) v[i + j] = t[j]; }
}
__device__ __forceinline__ void tmem_wait_st() {}
__device__ __forceinline__ void tmem_wait_ld() {}
#endif

// ---- memory-phase lock -------------------------------------------------------------------------------------------------
// The groups of a CTA run identical code on identical tile sizes, so they stay in lockstep: both gather at the same time,
// then both sit in tensor-core round trips while the load/store path idles (measured: the phase times of a kernel add up
// to its duration).  A CTA-wide lock around the gather / scatter phases forces them out of phase: while one group owns
// the load/store path, the other runs its layers.  (The owner keeps enough loads in flight to fill the path on its own.)
// Measured at config 2 with the lock in k_geo_tc (both passes) and k_bwd_geo_tc: 522 -> 490 ms per step (it hurts the
// reduction-bound colour backward, which therefore never takes it).  -DTT_MEMLOCK=0 builds without it.  Validation: see
// DESIGN.md 3.4 (racecheck clean, bit-identical repeated field queries, GPU suite soaked on several boxes).
#ifndef TT_MEMLOCK
#define TT_MEMLOCK 1
#endif
__device__ __forceinline__ void mem_lock(int* lock, bool leader, int group) {
#if TT_MEMLOCK
    if (leader) {
#ifndef TT_EMUL
        while (atomicCAS(lock, 0, 1) != 0) __nanosleep(100);
#else
        while (tt_emul::cas(lock, 0, 1) != 0) std::this_thread::yield();
#endif
    }
#endif
    group_sync(group);
}
__device__ __forceinline__ void mem_unlock(int* lock, bool leader, int group) {
    group_sync(group);
#if TT_MEMLOCK
#ifndef TT_EMUL
    if (leader) atomicExch(lock, 0);
#else
    if (leader) tt_emul::cas(lock, 1, 0);
#endif
#endif
}

// ---- group-level layer steps ---------------------------------------------------------------------------------
// this thread's activation row x[0..K) -> TMEM (tf32 hi at A_hi, exact remainder at A_lo)
template <int K>
__device__ __forceinline__ void umma_put_A(const Umma& u, const float (&x)[K], uint32_t col0 = 0) {
    constexpr int K32 = K / 32 * 32;
#pragma unroll
    for (int k0 = 0; k0 < K32; k0 += 32) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float h = tf32_hi(x[k0 + j]);
            hi[j] = __float_as_uint(h);
            lo[j] = __float_as_uint(x[k0 + j] - h);
        }
        tmem_st32(u.tmem + u.lane_base + TC_COL_AHI + col0 + k0, hi);
        tmem_st32(u.tmem + u.lane_base + TC_COL_ALO + col0 + k0, lo);
    }
#pragma unroll
    for (int k0 = K32; k0 < K; k0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float h = tf32_hi(x[k0 + j]);
            hi[j] = __float_as_uint(h);
            lo[j] = __float_as_uint(x[k0 + j] - h);
        }
        tmem_st8(u.tmem + u.lane_base + TC_COL_AHI + col0 + k0, hi);
        tmem_st8(u.tmem + u.lane_base + TC_COL_ALO + col0 + k0, lo);
    }
    tmem_wait_st();
    tc_fence_before();
}
// single-pass variant: only the (rounded) hi part
template <int K>
__device__ __forceinline__ void umma_put_A1(const Umma& u, const float (&x)[K]) {
    constexpr int K32 = K / 32 * 32;
#pragma unroll
    for (int k0 = 0; k0 < K32; k0 += 32) {
        uint32_t hi[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) hi[j] = __float_as_uint(tf32_rn(x[k0 + j]));
        tmem_st32(u.tmem + u.lane_base + TC_COL_AHI + k0, hi);
    }
#pragma unroll
    for (int k0 = K32; k0 < K; k0 += 8) {
        uint32_t hi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) hi[j] = __float_as_uint(tf32_rn(x[k0 + j]));
        tmem_st8(u.tmem + u.lane_base + TC_COL_AHI + k0, hi);
    }
    tmem_wait_st();
    tc_fence_before();
}
// this thread's output row d[0..N) <- TMEM
template <int N>
__device__ __forceinline__ void umma_get_D(const Umma& u, float (&d)[N], uint32_t col = TC_COL_D) {
    constexpr int N32 = N / 32 * 32;
#pragma unroll
    for (int n0 = 0; n0 < N32; n0 += 32) {
        uint32_t v[32];
        tmem_ld32(u.tmem + u.lane_base + col + n0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) d[n0 + j] = __uint_as_float(v[j]);
    }
#pragma unroll
    for (int n0 = N32; n0 < N; n0 += 8) {
        uint32_t v[8];
        tmem_ld8(u.tmem + u.lane_base + col + n0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[n0 + j] = __uint_as_float(v[j]);
    }
    tmem_wait_ld();
    tc_fence_before();
}
// this thread's row x[0..K) -> TMEM columns col_hi.. (tf32 hi) and col_lo.. (exact remainder), explicit columns
template <int K>
__device__ __forceinline__ void umma_put_A_ex(const Umma& u, const float (&x)[K], uint32_t col_hi, uint32_t col_lo) {
    constexpr int K32 = K / 32 * 32;
#pragma unroll
    for (int k0 = 0; k0 < K32; k0 += 32) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float h = tf32_hi(x[k0 + j]);
            hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(x[k0 + j] - h);
        }
        tmem_st32(u.tmem + u.lane_base + col_hi + k0, hi);
        tmem_st32(u.tmem + u.lane_base + col_lo + k0, lo);
    }
#pragma unroll
    for (int k0 = K32; k0 < K; k0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float h = tf32_hi(x[k0 + j]);
            hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(x[k0 + j] - h);
        }
        tmem_st8(u.tmem + u.lane_base + col_hi + k0, hi);
        tmem_st8(u.tmem + u.lane_base + col_lo + k0, lo);
    }
    tmem_wait_st();
    tc_fence_before();
}
// one full layer: put A, barrier, one thread issues + commits, everybody waits, get D
template <int K, int N, int PASSES>
__device__ __forceinline__ void umma_layer(Umma& u, bool leader, const float (&x)[K], const BTile& b, float (&d)[N]) {
    if (PASSES == 3) umma_put_A<K>(u, x); else umma_put_A1<K>(u, x);
    group_sync(u.group);
    if (leader) { umma_mma<PASSES>(u, b, K, false); umma_commit(u); }
    umma_wait(u);
    umma_get_D<N>(u, d);
}

// cooperative fill of a B tile from a row-major source: element (n,k) = src(n,k)
template <typename F>
__device__ __forceinline__ void btile_fill(float* hi, float* lo, int N, int K, F src, int tid, int nthreads) {
    for (int i = tid; i < N * K; i += nthreads) {
        const int n = i / K, k = i - n * K;
        const float w = src(n, k);
        const float h = tf32_hi(w);
        const int off = btile_off(n, k, K);
        hi[off] = h;
        lo[off] = w - h;
    }
}
__device__ __forceinline__ BTile btile_make(const float* hi, const float* lo, int N, int K) {
    BTile b; b.hi = smem_u32(hi); b.lo = smem_u32(lo); b.sbo = (uint32_t)(K >> 2) * 128u; b.N = N; b.K = K;
    return b;
}

}  // namespace tt
