// dpb_tile.cuh -- CTA-wide building blocks of the exact (CUDA-core FMA) path.
//
// A CTA of NTHREADS threads owns a tile of P = 8*TP paths.  Activations live TRANSPOSED in shared
// memory, [feature][path] with the padded path stride LDP, so that the tile advances in lock-step
// and every per-path scalar of dpb_eqn.h is a column.  GEMM inner loops call fma() explicitly (the
// translation unit is compiled with -fmad=false, so nothing else is contracted).
#pragma once
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <stdint.h>
#include "dpb_eqn.h"

namespace dpb {

constexpr int NTHREADS = 256;
constexpr int MAXLIN = 7;            // linear layers per network: <= 6 hidden + last
constexpr int WS_NMAX = 256;         // widest layer the weight stage holds
constexpr int SMALL_ROWS = 32;       // rows of the small per-tile arrays (>= round8(max(dim, control_dim + 1)))

template <typename real> struct RT;
template <> struct RT<float>  { static constexpr int KC = 8; static constexpr int VEC = 4; static constexpr int PADP = 4; static constexpr int TP = 4; };
template <> struct RT<double> { static constexpr int KC = 4; static constexpr int VEC = 2; static constexpr int PADP = 2; static constexpr int TP = 2; };

__host__ __device__ inline int round8(int x) { return (x + 7) & ~7; }

// Aligned vector loads/stores on shared memory (alignment guaranteed by layout: LDP, npad and tile
// offsets are multiples of the vector width).
template <int N> __device__ __forceinline__ void ldv(float* d, const float* s) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            float4 v = reinterpret_cast<const float4*>(s)[i];
            d[4 * i] = v.x; d[4 * i + 1] = v.y; d[4 * i + 2] = v.z; d[4 * i + 3] = v.w;
        }
    } else if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            float2 v = reinterpret_cast<const float2*>(s)[i];
            d[2 * i] = v.x; d[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = s[i];
    }
}
template <int N> __device__ __forceinline__ void ldv(double* d, const double* s) {
    if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            double2 v = reinterpret_cast<const double2*>(s)[i];
            d[2 * i] = v.x; d[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = s[i];
    }
}
template <int N> __device__ __forceinline__ void stv(float* d, const float* s) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) reinterpret_cast<float4*>(d)[i] = make_float4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = s[i];
    }
}
template <int N> __device__ __forceinline__ void stv(double* d, const double* s) {
    if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) reinterpret_cast<double2*>(d)[i] = make_double2(s[2 * i], s[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = s[i];
    }
}

// cp.async staging of one K-chunk of a packed weight matrix (rows contiguous: row stride == npad)
template <typename real>
__device__ __forceinline__ void stage_chunk(real* dst, const real* __restrict__ src, int nelem) {
    constexpr int VEC = RT<real>::VEC;
    const int nvec = nelem / VEC;                                   // nelem is a multiple of 8
    for (int v = threadIdx.x; v < nvec; v += NTHREADS)
        __pipeline_memcpy_async(dst + v * VEC, src + v * VEC, 16);
    __pipeline_commit();
}

// out^T[n][p] = epi( sum_k W[k][n] * in^T[k][p] ),  k < kpad (multiple of 8), n < npad (multiple of 8).
//   W: global, packed [kpad][npad] (zero padded).  in: smem [kpad][LDP].  Ws: smem 2 x KC x WS_NMAX.
//   TN = 8: thread tile 8 features x TP paths (wide layers); TN = 1: one feature x TP paths (npad <= 32).
//   epi(n, pbase, acc[TP]) is called once per owned feature row.  Starts and ends with __syncthreads().
template <typename real, int TN, class Epi>
__device__ __forceinline__ void gemm_AW(const real* __restrict__ Wg, int kpad, int npad, const real* in, real* Ws, Epi epi) {
    constexpr int KC = RT<real>::KC;
    constexpr int TP = RT<real>::TP;
    constexpr int LDP = 8 * TP + RT<real>::PADP;
    const int tid = threadIdx.x;
    const int pg = tid & 7, ng = tid >> 3;
    const int n0 = (TN == 8) ? ng * 8 : ng;
    const bool active = n0 < npad;
    real acc[TN][TP];
#pragma unroll
    for (int i = 0; i < TN; ++i)
#pragma unroll
        for (int j = 0; j < TP; ++j) acc[i][j] = (real)0;

    const int nchunks = kpad / KC;
    const int chunk_elems = KC * npad;
    __syncthreads();                                                // `in` complete, Ws free
    stage_chunk(Ws, Wg, chunk_elems);
    for (int c = 0; c < nchunks; ++c) {
        __pipeline_wait_prior(0);
        __syncthreads();
        if (c + 1 < nchunks) stage_chunk(Ws + ((c + 1) & 1) * (KC * WS_NMAX), Wg + (size_t)(c + 1) * chunk_elems, chunk_elems);
        if (active) {
            const real* w = Ws + (c & 1) * (KC * WS_NMAX) + n0;
            const real* a = in + (size_t)(c * KC) * LDP + pg * TP;
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                real wv[TN], av[TP];
                ldv<TN>(wv, w + k * npad);
                ldv<TP>(av, a + k * LDP);
#pragma unroll
                for (int i = 0; i < TN; ++i)
#pragma unroll
                    for (int j = 0; j < TP; ++j) acc[i][j] = fma(wv[i], av[j], acc[i][j]);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < TN; ++i) epi(n0 + i, pg * TP, acc[i]);
    }
    __syncthreads();
}

// g[i*nl + j] += sum_p A^T[i][p] * dY^T[j][p]   (i < kl, j < nl), accumulated into this CTA's private
// gradient slab with RED (each address has one owner thread per call and calls are separated by
// barriers => deterministic).  A, dY: smem [..][LDP] with rows allocated up to multiples of 8.
template <typename real>
__device__ __forceinline__ void gemm_dW(const real* A, int kl, const real* dY, int nl, real* g) {
    constexpr int TP = RT<real>::TP;
    constexpr int P = 8 * TP;
    constexpr int LDP = P + RT<real>::PADP;
    constexpr int VEC = RT<real>::VEC;
    const int TI = (kl + 7) >> 3, TJ = (nl + 7) >> 3;
    for (int tt = threadIdx.x; tt < TI * TJ; tt += NTHREADS) {
        const int ti = tt / TJ, tj = tt - ti * TJ;
        real acc[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int s = 0; s < 8; ++s) acc[r][s] = (real)0;
        for (int p = 0; p < P; p += VEC) {
            real av[8][VEC], bv[8][VEC];
#pragma unroll
            for (int r = 0; r < 8; ++r) ldv<VEC>(av[r], A + (ti + r * TI) * LDP + p);
#pragma unroll
            for (int s = 0; s < 8; ++s) ldv<VEC>(bv[s], dY + (tj + s * TJ) * LDP + p);
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int s = 0; s < 8; ++s) acc[r][s] = fma(av[r][v], bv[s][v], acc[r][s]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = ti + r * TI;
            if (i < kl) {
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int j = tj + s * TJ;
                    if (j < nl) atomicAdd(g + (size_t)i * nl + j, acc[r][s]);
                }
            }
        }
    }
}

// g[j] += sum_p dY^T[j][p]
template <typename real>
__device__ __forceinline__ void colsum_dY(const real* dY, int nl, real* g) {
    constexpr int P = 8 * RT<real>::TP;
    constexpr int LDP = P + RT<real>::PADP;
    for (int j = threadIdx.x; j < nl; j += NTHREADS) {
        real s = (real)0;
        for (int p = 0; p < P; ++p) s = s + dY[j * LDP + p];
        atomicAdd(g + j, s);
    }
}

// Sum over the CTA of one value per thread; result valid in thread 0.  red: smem [NTHREADS/32].
template <typename real>
__device__ __forceinline__ real block_sum(real v, real* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    real s = (real)0;
    if (threadIdx.x == 0)
        for (int i = 0; i < NTHREADS / 32; ++i) s = s + red[i];
    return s;
}

}  // namespace dpb
