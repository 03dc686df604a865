// dpb_device.cuh -- device building blocks of the exact (CUDA-core FMA) path.
//
//  * Eq<real>: closed forms of the four equations (reference equation.py:144-311) and the partial
//    derivatives the actor's reverse sweep needs.
//  * Philox4x32-10 counter-based increments.
//  * CTA-wide GEMMs on activations kept TRANSPOSED in shared memory, [feature][path] with a padded
//    path stride LDP, so that a tile of P paths advances in lock-step.
//
// Compiled with -fmad=false: elementwise arithmetic is uncontracted IEEE (a numpy float32 mirror
// reproduces the step schedule bit for bit); the GEMM inner loops call fma() explicitly.
#pragma once
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <stdint.h>

namespace dpb {

constexpr int NTHREADS = 256;
constexpr int MAXLIN = 7;            // linear layers per network: <= 6 hidden + last
constexpr int WS_NMAX = 256;         // widest layer the weight stage holds

template <typename real> struct RT;
template <> struct RT<float>  { static constexpr int KC = 8; static constexpr int VEC = 4; static constexpr int PADP = 4; };
template <> struct RT<double> { static constexpr int KC = 4; static constexpr int VEC = 2; static constexpr int PADP = 2; };

__host__ __device__ inline int round8(int x) { return (x + 7) & ~7; }

// ------------------------------------------------------------------------------------------------
// Equation constants.  The host fills the double struct; kernels convert once to `real`.
struct EqnD {
    int eqn, d, m, scheme, td;
    double R, R2, gamma, sig;
    double p, q, beta, k, a, eps, a2, a3;
    double cu;        // LQR: -beta*k/q                                    (equation.py:164)
    double wconst;    // LQR/LQR_var: 2*k*d ; VDP: 2*a*d                    (equation.py:155,199,290)
    double ZR;        // LQR/LQR_var: k*R^2                                 (equation.py:158,293)
    double C0;        // ekn: 3(d+1)a3/(2 a2 d)                             (equation.py:272)
    double lv_num;    // LQR_var: k^2 (beta+2eps)^2                         (equation.py:289)
    double lv_den;    // LQR_var: 2 k eps^2
    double lv_un;     // LQR_var: beta + 2 eps                              (equation.py:299)
    double lv_ud;     // LQR_var: q/k
    double lv_ue;     // LQR_var: 2 eps^2
    double lv_gk;     // LQR_var: gamma*k
    // scheme constants (equation.py:48-49,75,80,85-86)
    double delta_t, sqrt_delta_t, hb, c3, hmin;
};

enum { EQ_LQR = 0, EQ_VDP = 1, EQ_EKN = 2, EQ_LQRVAR = 3 };

template <typename real>
struct Eq {
    int eqn, d, m, scheme, td;
    real R, R2, gamma, sig, p, q, beta, k, a, eps, a2, a3, cu, wconst, ZR, C0;
    real lv_num, lv_den, lv_un, lv_ud, lv_ue, lv_gk;
    real delta_t, sqrt_delta_t, hb, c3, hmin;
    __device__ explicit Eq(const EqnD& e)
        : eqn(e.eqn), d(e.d), m(e.m), scheme(e.scheme), td(e.td),
          R((real)e.R), R2((real)e.R2), gamma((real)e.gamma), sig((real)e.sig), p((real)e.p), q((real)e.q),
          beta((real)e.beta), k((real)e.k), a((real)e.a), eps((real)e.eps), a2((real)e.a2), a3((real)e.a3),
          cu((real)e.cu), wconst((real)e.wconst), ZR((real)e.ZR), C0((real)e.C0),
          lv_num((real)e.lv_num), lv_den((real)e.lv_den), lv_un((real)e.lv_un), lv_ud((real)e.lv_ud),
          lv_ue((real)e.lv_ue), lv_gk((real)e.lv_gk),
          delta_t((real)e.delta_t), sqrt_delta_t((real)e.sqrt_delta_t), hb((real)e.hb), c3((real)e.c3),
          hmin((real)e.hmin) {}
};

// All per-path functions address smem column arrays as v[k*ld + p].
#define DPB_AT(v, k) ((v)[(k) * ld + p])

template <typename real>
__device__ inline real norm2_path(const real* x, int d, int ld, int p) {
    real s = (real)0;
    for (int k = 0; k < d; ++k) s = s + DPB_AT(x, k) * DPB_AT(x, k);
    return s;
}

// u_true (equation.py:163-164, 212-217, 259-261, 298-299)
template <typename real>
__device__ inline void eq_u_true(const Eq<real>& E, const real* x, real* u, int ld, int p) {
    const int d = E.d, m = E.m;
    switch (E.eqn) {
    case EQ_LQR:
        for (int k = 0; k < d; ++k) DPB_AT(u, k) = E.cu * DPB_AT(x, k);
        break;
    case EQ_VDP:
        for (int j = 0; j < m; ++j) {
            real x2 = DPB_AT(x, m + j);
            real px2 = DPB_AT(x, m + (j + 1 == m ? 0 : j + 1));
            real nx2 = DPB_AT(x, m + (j == 0 ? m - 1 : j - 1));
            DPB_AT(u, j) = -((real)2 * E.a * x2 - E.eps * (px2 + nx2)) / (real)2 / E.q;
        }
        break;
    case EQ_EKN: {
        real r = sqrt(norm2_path(x, d, ld, p));
        for (int k = 0; k < d; ++k) DPB_AT(u, k) = DPB_AT(x, k) / r;
        break;
    }
    default:
        for (int k = 0; k < d; ++k) {
            real xk = DPB_AT(x, k);
            DPB_AT(u, k) = -E.lv_un * xk / (E.lv_ud + E.lv_ue * xk * xk);
        }
    }
}

// V_true (equation.py:160,204-210,255-257,295)
template <typename real>
__device__ inline real eq_V_true(const Eq<real>& E, const real* x, int ld, int p) {
    const int d = E.d, m = E.m;
    real n2 = norm2_path(x, d, ld, p);
    switch (E.eqn) {
    case EQ_LQR:
    case EQ_LQRVAR:
        return n2 * E.k;
    case EQ_VDP: {
        real s = (real)0;
        for (int j = 0; j < m; ++j) {
            int jn = (j + 1 == m ? 0 : j + 1);
            s = s + (DPB_AT(x, j) * DPB_AT(x, jn) + DPB_AT(x, m + j) * DPB_AT(x, m + jn));
        }
        return E.a * n2 - E.eps * s;
    }
    default: {
        real r = sqrt(n2);
        return E.a3 * r * r * r - E.a2 * r * r;
    }
    }
}

// Z_tf on the boundary (equation.py:157,201,252,292)
template <typename real>
__device__ inline real eq_Z(const Eq<real>& E, const real* x, int ld, int p) {
    if (E.eqn == EQ_LQR || E.eqn == EQ_LQRVAR) return E.ZR;
    return eq_V_true(E, x, ld, p);
}

// V_grad_true (equation.py:166,219-227,263-265,301) -> g[k]
template <typename real>
__device__ inline void eq_V_grad_true(const Eq<real>& E, const real* x, real* g, int ld, int p) {
    const int d = E.d, m = E.m;
    switch (E.eqn) {
    case EQ_LQR:
    case EQ_LQRVAR:
        for (int k = 0; k < d; ++k) DPB_AT(g, k) = (real)2 * E.k * DPB_AT(x, k);
        break;
    case EQ_VDP:
        for (int j = 0; j < m; ++j) {
            int jn = (j + 1 == m ? 0 : j + 1), jp = (j == 0 ? m - 1 : j - 1);
            DPB_AT(g, j) = (real)2 * E.a * DPB_AT(x, j) - E.eps * (DPB_AT(x, jn) + DPB_AT(x, jp));
            DPB_AT(g, m + j) = (real)2 * E.a * DPB_AT(x, m + j) - E.eps * (DPB_AT(x, m + jn) + DPB_AT(x, m + jp));
        }
        break;
    default: {
        real r = sqrt(norm2_path(x, d, ld, p));
        real c = (real)3 * E.a3 * r - (real)2 * E.a2;
        for (int k = 0; k < d; ++k) DPB_AT(g, k) = c * DPB_AT(x, k);
    }
    }
}

// running cost w_tf (equation.py:154,188-199,249,288-290)
template <typename real>
__device__ inline real eq_w(const Eq<real>& E, const real* x, const real* u, int ld, int p) {
    const int d = E.d, m = E.m;
    switch (E.eqn) {
    case EQ_LQR: {
        real s = (real)0;
        for (int k = 0; k < d; ++k) s = s + (E.p * DPB_AT(x, k) * DPB_AT(x, k) + E.q * DPB_AT(u, k) * DPB_AT(u, k));
        return s - E.wconst;
    }
    case EQ_VDP: {
        real s = (real)0, n2 = (real)0;
        for (int j = 0; j < m; ++j) {
            int jn = (j + 1 == m ? 0 : j + 1), jp = (j == 0 ? m - 1 : j - 1);
            real x1 = DPB_AT(x, j), x2 = DPB_AT(x, m + j);
            real px1 = DPB_AT(x, jn), px2 = DPB_AT(x, m + jn), nx1 = DPB_AT(x, jp), nx2 = DPB_AT(x, m + jp);
            real dv1 = (real)2 * E.a * x1 - E.eps * (px1 + nx1);
            real dv2 = (real)2 * E.a * x2 - E.eps * (px2 + nx2);
            real uj = DPB_AT(u, j);
            real t = -E.gamma * E.eps * (x1 * px1 + x2 * px2) + (dv2 * dv2) / (real)4 / E.q - x2 * dv1
                     - (((real)1 - x1 * x1) * x2 - x1) * dv2;
            s = s + (t + E.q * (uj * uj));
            n2 = n2 + (x1 * x1 + x2 * x2);
        }
        return s + E.gamma * E.a * n2 - E.wconst;
    }
    case EQ_EKN:
        return (real)1;
    default: {
        real s1 = (real)0, s2 = (real)0;
        for (int k = 0; k < d; ++k) {
            real xk = DPB_AT(x, k), uk = DPB_AT(u, k);
            s1 = s1 + E.lv_num * xk * xk / (E.q + E.lv_den * xk * xk);
            s2 = s2 + (E.lv_gk * xk * xk + E.q * uk * uk);
        }
        return s1 + s2 - E.wconst;
    }
    }
}

// Per-path state shared by drift evaluations (ekn needs |x|).
template <typename real>
struct DriftCtx { real c; real r; };

template <typename real>
__device__ inline DriftCtx<real> eq_drift_ctx(const Eq<real>& E, const real* x, int ld, int p) {
    DriftCtx<real> c;
    c.c = (real)0; c.r = (real)0;
    if (E.eqn == EQ_EKN) {
        c.r = sqrt(norm2_path(x, E.d, ld, p));
        c.c = E.C0 / ((real)2 * E.a2 - (real)3 * E.a3 * c.r);          // equation.py:272
    }
    return c;
}

// drift component k (equation.py:172,232-235,270-273,307)
template <typename real>
__device__ inline real eq_drift(const Eq<real>& E, const DriftCtx<real>& C, const real* x, const real* u, int k, int ld, int p) {
    switch (E.eqn) {
    case EQ_LQR:
    case EQ_LQRVAR:
        return E.beta * DPB_AT(u, k);
    case EQ_VDP: {
        const int m = E.m;
        if (k < m) return DPB_AT(x, m + k);
        int j = k - m;
        real x1 = DPB_AT(x, j), x2 = DPB_AT(x, m + j);
        return ((real)1 - x1 * x1) * x2 - x1 + DPB_AT(u, j);
    }
    default:
        return C.c * DPB_AT(u, k);
    }
}

// diagonal of sigma, component k (equation.py:170,230,268,305)
template <typename real>
__device__ inline real eq_sigma(const Eq<real>& E, const real* x, const real* u, int k, int ld, int p) {
    if (E.eqn == EQ_LQRVAR) return E.sig * ((real)1 + E.eps * DPB_AT(x, k) * DPB_AT(u, k));
    return E.sig;
}

// adaptive-scheme flag of a point with norm nrm (equation.py:80-82,94-95):
//   1 + floor((sign(R-n-hb) + sign(R-n))/2)  ==  (R-n > 0) ? ((R-n-hb > 0) ? 2 : 1) : 0
template <typename real>
__device__ inline int eq_flag(const Eq<real>& E, real nrm) {
    real t2 = E.R - nrm;
    real t1 = E.R - nrm - E.hb;
    return (t2 > (real)0) ? ((t1 > (real)0) ? 2 : 1) : 0;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  counter = (path_lo, path_hi, step, chunk), key = seed
// xor-folded with the stream id.
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}

// Four increments for (global path, step, chunk): N(0,1) by Box-Muller, or the 3-point law
// {-sqrt3, 0, +sqrt3} with probabilities 1/6, 4/6, 1/6 (equation.py:31-32).  Always float.
__device__ inline void philox_increments(int mode_bounded, uint64_t seed, uint64_t stream, uint64_t path,
                                         uint32_t step, uint32_t chunk, float out[4]) {
    uint32_t c[4] = {(uint32_t)path, (uint32_t)(path >> 32), step, chunk};
    uint32_t k0 = (uint32_t)seed ^ (uint32_t)(stream * 0x9E3779B97F4A7C15ull >> 32);
    uint32_t k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(stream * 0x9E3779B97F4A7C15ull);
    philox4x32_10(c, k0, k1);
    if (mode_bounded) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // randint(6) via multiply-high; floor((k-1)/4): k=0 -> -1, k=1..4 -> 0, k=5 -> +1
            uint32_t k = (uint32_t)(((uint64_t)c[i] * 6u) >> 32);
            out[i] = (k == 0u) ? -1.7320508075688772f : ((k == 5u) ? 1.7320508075688772f : 0.0f);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float u1 = ((float)(c[2 * i] >> 8) + 0.5f) * (1.0f / 16777216.0f);       // (0,1)
            float u2 = ((float)(c[2 * i + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
            float r = sqrtf(-2.0f * logf(u1));
            float s, co;
            sincospif(2.0f * u2, &s, &co);
            out[2 * i] = r * co;
            out[2 * i + 1] = r * s;
        }
    }
}


// Aligned vector loads from shared memory into register arrays (alignment guaranteed by layout:
// LDP, npad and tile offsets are multiples of the vector width).
template <int N> __device__ inline void ldv(float* d, const float* s) {
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
template <int N> __device__ inline void ldv(double* d, const double* s) {
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

// ------------------------------------------------------------------------------------------------
// cp.async staging of one K-chunk of a packed weight matrix (rows are contiguous: row stride == npad)
template <typename real>
__device__ inline void stage_chunk(real* dst, const real* __restrict__ src, int nelem) {
    constexpr int VEC = RT<real>::VEC;
    const int nvec = nelem / VEC;                                   // nelem is a multiple of 8
    for (int v = threadIdx.x; v < nvec; v += NTHREADS)
        __pipeline_memcpy_async(dst + v * VEC, src + v * VEC, 16);
    __pipeline_commit();
}

// out^T[n][p] = epi( sum_k W[k][n] * in^T[k][p] ),  k < kpad (multiple of 8), n < npad (multiple of 8).
//   W: global, packed [kpad][npad] (zero padded).  in: smem [kpad][LDP].  Ws: smem 2 x KC x WS_NMAX.
//   TN = 8: thread tile 8 features x TP paths (wide layers); TN = 1: one feature x TP paths (npad <= 32).
//   epi(n, pbase, acc[TP]) is called once per owned feature row.  Ends with __syncthreads().
template <typename real, int TP, int TN, class Epi>
__device__ inline void gemm_AW(const real* __restrict__ Wg, int kpad, int npad, const real* in, real* Ws, Epi epi) {
    constexpr int KC = RT<real>::KC;
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
// gradient slab with RED (program-ordered per address => deterministic).  A, dY: smem [..][LDP],
// allocated with rows padded to multiples of 8.
template <typename real, int TP>
__device__ inline void gemm_dW(const real* A, int kl, const real* dY, int nl, real* g) {
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
template <typename real, int TP>
__device__ inline void colsum_dY(const real* dY, int nl, real* g) {
    constexpr int P = 8 * TP;
    constexpr int LDP = P + RT<real>::PADP;
    for (int j = threadIdx.x; j < nl; j += NTHREADS) {
        real s = (real)0;
        for (int p = 0; p < P; ++p) s = s + dY[j * LDP + p];
        atomicAdd(g + j, s);
    }
}

}  // namespace dpb
