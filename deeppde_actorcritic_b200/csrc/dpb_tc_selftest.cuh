// dpb_tc_selftest.cuh -- one-CTA check of the three tcgen05 product forms the tensor path uses:
//   D1 = A * B^T          A[128][K], B[208][K] both K-major images in shared memory   (forward / dX)
//   D2 = A * B^T          A read from TMEM (packed bf16 pairs written with tcgen05.st) (forward / dX)
//   D3 = A^T * C          A[128 paths][K feats], C[128 paths][208 feats] read MN-major (dW)
// Inputs are float (expected bf16-representable), outputs float [3][128][208].
#pragma once
#include "dpb_tc.cuh"
#include "dpb_tc_nets.cuh"

namespace dpb {
namespace tc {

constexpr int ST_N = 208;

__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* Aimg = smem;                              // R = 128
    unsigned char* Bimg = Aimg + 128 * K * 2;                // R = 208
    unsigned char* Cimg = Bimg + ST_N * K * 2;               // R = 128, K = 208 columns (features g)
    uint64_t* bar = reinterpret_cast<uint64_t*>(Cimg + 128 * ST_N * 2);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);

    for (int i = tid; i < 128 * K; i += 128) {
        int r = i / K, k = i - r * K;
        *reinterpret_cast<__nv_bfloat16*>(Aimg + img_off(r, k, 128)) = __float2bfloat16_rn(A[i]);
    }
    for (int i = tid; i < ST_N * K; i += 128) {
        int r = i / K, k = i - r * K;
        *reinterpret_cast<__nv_bfloat16*>(Bimg + img_off(r, k, ST_N)) = __float2bfloat16_rn(B[i]);
    }
    for (int i = tid; i < 128 * ST_N; i += 128) {
        int p = i / ST_N, g = i - p * ST_N;
        float v = (g < K) ? B[p * K + g] : 0.f;              // C[p][g] = B[p][g] (first 128 rows of B)
        *reinterpret_cast<__nv_bfloat16*>(Cimg + img_off(p, g, 128)) = __float2bfloat16_rn(v);
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;
    const uint32_t idK = idesc_bf16(128, ST_N, 0, 0), idMN = idesc_bf16(128, ST_N, 1, 1);

    // ---- test 1: SS, K-major
    if (tid == 0) {
        for (int s = 0; s < K / 16; ++s) {
            uint64_t ad = smem_desc(smem_u32(Aimg) + s * 2 * (128 / 8) * 128, (128 / 8) * 128, 128);
            uint64_t bd = smem_desc(smem_u32(Bimg) + s * 2 * (ST_N / 8) * 128, (ST_N / 8) * 128, 128);
            mma_ss(tbase, ad, bd, idK, s > 0);
        }
        tc_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    for (int c = 0; c < ST_N / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(lane_base + c * 16, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[(0 * 128 + tid) * ST_N + c * 16 + j] = __uint_as_float(v[j]);
    }
    // ---- test 2: TS (A row of this thread -> TMEM columns 256.. as packed pairs)
    for (int c = 0; c < K / 16; ++c) {
        uint32_t v[8];
        for (int j = 0; j < 8; ++j)
            v[j] = pack2(__float2bfloat16_rn(A[tid * K + c * 16 + 2 * j]), __float2bfloat16_rn(A[tid * K + c * 16 + 2 * j + 1]));
        tmem_st8(lane_base + 256 + c * 8, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        for (int s = 0; s < K / 16; ++s) {
            uint64_t bd = smem_desc(smem_u32(Bimg) + s * 2 * (ST_N / 8) * 128, (ST_N / 8) * 128, 128);
            mma_ts(tbase, tbase + 256 + s * 8, bd, idK, s > 0);
        }
        tc_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    for (int c = 0; c < ST_N / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(lane_base + c * 16, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[(1 * 128 + tid) * ST_N + c * 16 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    // ---- test 3: SS, both MN-major: D3[f][g] = sum_p A[p][f] * C[p][g],  f < 128 (needs K >= 128)
    if (tid == 0) {
        tc_fence_after();
        for (int s = 0; s < 128 / 16; ++s) {
            uint64_t ad = smem_desc(smem_u32(Aimg) + s * 2 * 128, 128, (128 / 8) * 128);
            uint64_t bd = smem_desc(smem_u32(Cimg) + s * 2 * 128, 128, (128 / 8) * 128);
            mma_ss(tbase, ad, bd, idMN, s > 0);
        }
        tc_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    for (int c = 0; c < ST_N / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(lane_base + c * 16, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[(2 * 128 + tid) * ST_N + c * 16 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}


// Latency of one control-thread <-> path-thread hand-off of the tensor path with (almost) no work in it:
// control thread: wait a_ready -> one tcgen05.mma (M=128, N=16, K=16) -> tcgen05.commit(acc_full);
// 8 path warps: wait acc_full -> tcgen05.ld of 16 columns -> tcgen05.st of 8 -> publish (one arrival per warp).
// out[0] = cycles per round trip (thread 0), out[1] = rounds.
__global__ void __launch_bounds__(288, 1) tc_handshake_kernel(long long* out, int rounds) {
    __shared__ __align__(128) unsigned char img[2 * 128 * 16 * 2];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (int)sizeof(img) / 4; i += 288) reinterpret_cast<uint32_t*>(img)[i] = 0u;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 8); fence_barrier_init(); }
    if (warp == 8) tmem_alloc(&tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    const uint32_t tl = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t idesc = idesc_bf16(128, 16, 0, 0);
    const long long t0 = clock64();
    if (warp == 8) {
        if (lane == 0) {
            for (int r = 0; r < rounds; ++r) {
                mbar_wait(&bars[1], r & 1);
                tc_fence_after();
                mma_ss(tbase, smem_desc(smem_u32(img), 2048, 128), smem_desc(smem_u32(img) + 4096, 256, 128), idesc, 0);
                tc_commit(&bars[0]);
            }
        }
    } else {
        for (int r = 0; r < rounds; ++r) {
            uint32_t v[16];
            if (r > 0) { mbar_wait(&bars[0], (r - 1) & 1); tc_fence_after(); tmem_ld16(tl, v); tmem_ld_wait(); } else { for (int j = 0; j < 16; ++j) v[j] = 0u; }
            tmem_st8(tl + 256, v);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[1]);
        }
        mbar_wait(&bars[0], (rounds - 1) & 1);
    }
    if (tid == 0) { out[0] = (clock64() - t0) / rounds; out[1] = rounds; }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tbase, 512);
}


// Cost of one hidden-layer epilogue (13 chunks of 16 columns: TMEM load, affine + y+relu(y), bf16 hi/lo split,
// two TMEM stores in place) when `ngroups` groups of 4 warps share the chunks of the 128 lanes.  No hand-off barriers: the
// pure path-thread work.  with_mma: one more warp issues back-to-back N=208 tcgen05.mma (A and D in the other TMEM region)
// for the whole time -- what the epilogue costs while the tensor pipe is busy with the next layer.  publish: 1 = every
// chunk is followed by the publish sequence of the kernels (wait::st, fence, __syncwarp, mbarrier arrive).
// out[0] = cycles per epilogue.
__global__ void tc_epilogue_bench_kernel(long long* out, int rounds, int ngroups, int with_mma, int publish) {
    __shared__ __align__(16) float gcbb[2 * 208];
    __shared__ __align__(1024) unsigned char bimg[208 * 32];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tslot;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nepi = 4 * ngroups;                            // epilogue warps; warp nepi (if present) issues the MMAs
    for (int i = tid; i < 416; i += blockDim.x) gcbb[i] = 0.5f + 0.001f * i;
    for (int i = tid; i < 208 * 32 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(bimg)[i] = 0u;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], (1u << 20) - 1u); stop = 0; fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    const uint32_t tl = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const int grp = warp >> 2;
    if (warp < nepi) {   // defined accumulator contents
        uint32_t z[8];
        for (int j = 0; j < 8; ++j) z[j] = __float_as_uint(0.25f * (j - 3));
        for (int c = 0; c < 64; ++c) tmem_st8(tl + 8 * c, z);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == nepi) {
        if (with_mma) {
            const uint32_t idesc = idesc_bf16(128, 208, 0, 0);
            const uint64_t bd = smem_desc(smem_u32(bimg), (208 >> 3) * 128, 128);
            int it = 0;
            while (!stop) {
                if (elect_one()) {
                    for (int j = 0; j < 13; ++j) mma_ts(tbase + 256, tbase + 256 + (j & 7) * 16, bd, idesc, 1);
                    tc_commit(&bars[0]);
                }
                __syncwarp();
                mbar_wait(&bars[0], it & 1);                 // at most 13 MMAs in flight
                ++it;
            }
        }
    } else if (warp < nepi) {
        const long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {                       // (after round 0 the columns hold bf16 pairs: same work)
            const float* gc = gcbb;
            const float* bb = gcbb + 208;
            uint32_t ra[16];
            for (int c = grp; c < 13; c += ngroups) {
                tmem_ld16(tl + 16 * c, ra);
                tmem_ld_wait();
                float v[16];
                affine16(ra, gc + 16 * c, bb + 16 * c, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = v[j] + fmaxf(v[j], 0.f);
                if (publish & 2) {                               // round-1 layout: separate plane regions
                    uint32_t h[8], l[8];
                    split16(v, h, l);
                    tmem_st8(tl + 256 + 8 * c, h);
                    tmem_st8(tl + 384 + 8 * c, l);
                } else {
                    put16(tl + 16 * c, v);                       // in place, as the kernels do
                }
                if (publish & 1) {
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(&bars[1]);
                }
            }
            tmem_st_wait();
            asm volatile("bar.sync 1, %0;" ::"r"(nepi * 32) : "memory");
        }
        if (tid == 0) { out[0] = (clock64() - t0) / rounds; out[1] = rounds; stop = 1; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// Cost of one tcgen05.mma (kind::f16, M=128, K=16, FP32 accumulation) as a function of N, issued back to back by one thread
// with both operands in place: ts = 1 reads A from tensor memory, 0 from shared memory.  `per_commit` MMAs are followed by
// one tcgen05.commit (as the weight-ring loop does).  out[0] = cycles per MMA (issue of the first to completion of the last).
__global__ void __launch_bounds__(128, 1) tc_mma_bench_kernel(long long* out, int n, int rounds, int ts, int per_commit) {
    extern __shared__ __align__(1024) unsigned char smem[];            // A image 128 x 16 (4 KB) + B image n x 16
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (4096 + n * 32) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    if (warp == 0) {
        const uint32_t idesc = idesc_bf16(128, n, 0, 0);
        const uint64_t ad = smem_desc(smem_u32(smem), 2048, 128), bd = smem_desc(smem_u32(smem) + 4096, (n >> 3) * 128, 128);
        const long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            if (elect_one()) {
                for (int j = 0; j < per_commit; ++j) {
                    if (ts) mma_ts(tbase, tbase + 256 + (j & 7) * 8, bd, idesc, 1);
                    else mma_ss(tbase, ad, bd, idesc, 1);
                }
                tc_commit(&bars[1]);
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(&bars[0]);
        const long long t1 = clock64();
        mbar_wait(&bars[0], 0);
        const long long t2 = clock64();
        if (tid == 0) { out[0] = (t2 - t0) / ((long long)rounds * per_commit); out[1] = (t1 - t0) / ((long long)rounds * per_commit); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace tc
}  // namespace dpb
