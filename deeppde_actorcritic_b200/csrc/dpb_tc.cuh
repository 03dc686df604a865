// dpb_tc.cuh -- sm_100a tensor-core primitives used by the tensor path (inline PTX only):
// mbarrier, bulk async copy (TMA engine, 1-D), TMEM allocation, tcgen05.mma / commit / ld / st,
// UMMA shared-memory and instruction descriptors.
//
// Operand images.  Every bf16 matrix operand [R rows][K] (R = the M or N extent, K = the contraction
// extent of the product it was built for) is stored as 8x8 "core matrices" of 128 contiguous bytes,
//      byte(r, k) = (k/8) * (R/8)*128 + (r/8) * 128 + (r%8) * 16 + (k%8) * 2
// i.e. [k-group][row-group][8 rows][8 k].  A 16-wide K slice of all rows is contiguous (one bulk copy),
// a thread that owns row r writes 16-byte pieces that are contiguous across the warp (no bank
// conflicts), and the same bytes serve two descriptor views (SWIZZLE_NONE):
//   K-major  (rows = M/N, contraction = K):  LBO = (R/8)*128 (next k-group), SBO = 128 (next row-group)
//   MN-major (M/N = k index, contraction = r): SBO = (R/8)*128, LBO = 128
// (cute/atom/mma_traits_sm100.hpp: K-major INTERLEAVE ((8,n),2):((1,SBO),LBO); MN-major INTERLEAVE
//  ((1,n),(8,k)):((X,SBO),(1,LBO)), in 16-byte units.)
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace dpb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }   // generic -> async proxy, global space
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }    // generic -> async proxy, all spaces
// 16-byte vector reduction (no return value) into global memory
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar_saddr) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { mbar_arrive_u32(smem_u32(bar)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Waits for the phase with the given parity to complete.  A wait that lasts ~4 s of GPU clocks is a
// protocol bug: trap (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait_u32(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;                                   // fast path: no clock reads when the phase is already complete
    const long long t0 = clock64();
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (clock64() - t0 > 8000000000LL) asm volatile("trap;");
    }
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_u32(smem_u32(bar), parity); }

// one non-blocking probe of the phase with the given parity
__device__ __forceinline__ bool mbar_test_u32(uint32_t addr, uint32_t parity) {         // non-blocking
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}

// CTA barriers over the first `nthreads` threads only (named barrier 1): the producer warp never joins them
__device__ __forceinline__ void bar_work_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
__device__ __forceinline__ int bar_work_or(int pred, int nthreads) {
    int r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbar.red.or.pred q, 1, %2, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
                 : "=r"(r) : "r"(pred), "r"(nthreads) : "memory");
    return r;
}

// ------------------------------------------------------------------------------- bulk copy (TMA 1-D)
// global -> shared, completion counted in bytes on an mbarrier.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------- TMEM
// One warp allocates `ncols` (power of two >= 32) columns; the base address is written to *slot (smem).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of the (converged) warp: the control warp runs its protocol on all 32 lanes with identical, provably
// warp-uniform values and issues tcgen05.mma / commit / arrive from the elected lane only -- the operands then sit in
// uniform registers; issued from `if (lane == 0)` code every MMA costs a broadcast loop of ~100 cycles.
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p)::"memory");
    return p != 0;
}
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// all previously issued tcgen05.mma of this thread complete -> arrive(1) on the mbarrier
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tc_commit_u32(uint32_t bar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4)..+31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------- descriptors
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100): cute/arch/mma_sm100_desc.hpp
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor for kind::f16 with BF16 operands and FP32 accumulation
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// the same with FP16 operands (a_format = b_format = 0)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]   (one thread issues)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: M lanes x K/2 columns, two bf16 per 32-bit column)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// One contraction chunk of a bf16x3 product (ah*bh, ah*bl, al*bh into D) and the commit that releases its weight slot, then
// wait until two mbarrier phases are complete (the next weight slot, the next operand chunk).  Both barriers are TESTED before
// the MMAs are issued and the results are only looked at (by branches, which the assembler cannot move) after the commit: a
// test has a latency of 80-250 cycles even when the phase is long complete, which the issue of the MMAs then hides.
// Executed by the whole (converged) warp on warp-uniform operands; the elected lane issues the MMAs and the commit.
__device__ __forceinline__ void mma_ts3_commit_wait2(uint32_t d_tmem, uint32_t ahi, uint32_t alo, uint64_t bhi, uint64_t blo, uint32_t idesc,
                                                      uint32_t accumulate, uint32_t release_bar, uint32_t bar1, uint32_t parity1,
                                                      uint32_t bar2, uint32_t parity2) {
    asm volatile("{\n\t.reg .pred pe, pa, pt, pu, p1;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 pt, [%8], %9;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 pu, [%10], %11;\n\t"
                 "elect.sync _|pe, 0xffffffff;\n\t"
                 "setp.ne.b32 pa, %4, 0;\n\t"
                 "setp.eq.b32 p1, %4, %4;\n\t"
                 "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %5, %3, pa;\n\t"
                 "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %6, %3, p1;\n\t"
                 "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %5, %3, p1;\n\t"
                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t"
                 "W1:\n\t"
                 "@pt bra D1;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 pt, [%8], %9;\n\t"
                 "bra W1;\n\t"
                 "D1:\n\t"
                 "@pu bra D2;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 pu, [%10], %11;\n\t"
                 "bra D1;\n\t"
                 "D2:\n\t"
                 "tcgen05.fence::after_thread_sync;\n\t}"
                 :
                 : "r"(d_tmem), "r"(ahi), "r"(alo), "r"(idesc), "r"(accumulate), "l"(bhi), "l"(blo), "r"(release_bar), "r"(bar1), "r"(parity1),
                   "r"(bar2), "r"(parity2)
                 : "memory");
}

// ------------------------------------------------------------------------------ bf16 hi/lo split
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 significant bits; a*w ~= ah*wh + ah*wl + al*wh.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {      // a -> bits 0..15 (lower index)
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// byte offset of element (r, k) in an operand image with R rows
__host__ __device__ constexpr uint32_t img_off(int r, int k, int R) {
    return (uint32_t)((k >> 3) * (R >> 3) * 128 + (r >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

}  // namespace tc
}  // namespace dpb
