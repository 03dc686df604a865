// dpb_nets.cuh -- DeepNN (reference solver.py:227-278) on a tile of paths: packed weights,
// forward, reverse-mode backward, and the raw-gradient -> flat-gradient conversion.
//
// With training=False at every call site of the reference each BatchNormalization is
// the fixed affine map z -> z*gamma*c + beta, c = 1/sqrt(1+1e-6).  Per network the pack kernel
// forms, once per train step:
//   g0c[k] = gamma0[k]*c, b0[k]                                   (bn0, solver.py:265)
//   W_l[kpad][npad]            zero padded                        (dense_l, solver.py:267,270)
//   WTg_l[npad][kpad] = W_l[k][n]*gc_l[n]                         (operand of the dX product)
//   gc_l[n] = gamma_{l+1}[n]*c ;  bb_l[n] = beta_{l+1}[n]  (last layer: bias*gc + beta)
// The backward accumulates RAW sums G_l[k][n] = sum a_{l-1}[k]*dz_l[n], C_l[n] = sum dz_l[n],
// SX[k] = sum x[k]*dy0[k], S0[k] = sum dy0[k] and finalize_grad turns them into the gradient of
// every trainable variable in the flat layout of include/deeppde_b200.h:
//   dW = G*gc ; dgamma = c*sum_k W[k][n]*G[k][n] (+ c*bias*C for the last layer) ; dbeta = C ;
//   dbias = gc*C ; dgamma0 = c*SX ; dbeta0 = S0.
#pragma once
#include "dpb_tile.cuh"

namespace dpb {

struct NetDev {
    int L;                          // hidden layers
    int in, out;                    // logical input / raw output width (ekn actor: control_dim + 1)
    int ekn_head, mctrl;
    int kl[MAXLIN], nl[MAXLIN];     // logical dims of linear layer l (l = 0..L)
    int kp[MAXLIN], np[MAXLIN];     // padded to multiples of 8
    // packed buffer (elements)
    long long offW[MAXLIN], offWT[MAXLIN], offg[MAXLIN], offb[MAXLIN], offg0, offb0, ptotal;
    // flat layout (elements)
    long long fW[MAXLIN], fg[MAXLIN], fb[MAXLIN], fbias, fg0, fb0, ftotal;
    // raw-gradient slab (elements)
    long long gW[MAXLIN], gC[MAXLIN], gX, g0, gtotal;
};

inline long long align8(long long x) { return (x + 7) & ~7LL; }

// in -> hid[0..L-1] -> out
inline void netdev_init(NetDev& nd, int in, const int* hid, int L, int out, int ekn_head, int mctrl) {
    nd.L = L; nd.in = in; nd.out = out; nd.ekn_head = ekn_head; nd.mctrl = mctrl;
    int prev = in;
    long long f = 0, p = 0, g = 0;
    nd.fg0 = f; f += in; nd.fb0 = f; f += in;
    nd.offg0 = p; p += round8(in); nd.offb0 = p; p += round8(in);
    for (int l = 0; l <= L; ++l) {
        int n = (l < L) ? hid[l] : out;
        nd.kl[l] = prev; nd.nl[l] = n; nd.kp[l] = round8(prev); nd.np[l] = round8(n);
        nd.fW[l] = f; f += (long long)prev * n;
        if (l == L) { nd.fbias = f; f += n; }
        nd.fg[l] = f; f += n; nd.fb[l] = f; f += n;
        nd.offW[l] = p; p += (long long)nd.kp[l] * nd.np[l];
        nd.offWT[l] = p; p += (long long)nd.kp[l] * nd.np[l];
        nd.offg[l] = p; p += nd.np[l]; nd.offb[l] = p; p += nd.np[l];
        nd.gW[l] = g; g += align8((long long)prev * n);
        nd.gC[l] = g; g += align8(n);
        prev = n;
    }
    nd.gX = g; g += align8(in); nd.g0 = g; g += align8(in);
    nd.ftotal = f; nd.ptotal = align8(p); nd.gtotal = g;
}

// ---------------------------------------------------------------------------------------- pack
template <typename real>
__global__ void pack_net_kernel(NetDev nd, const real* __restrict__ th, real* __restrict__ pk, real c) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = t0; i < round8(nd.in); i += stride) {
        pk[nd.offg0 + i] = i < nd.in ? th[nd.fg0 + i] * c : (real)0;
        pk[nd.offb0 + i] = i < nd.in ? th[nd.fb0 + i] : (real)0;
    }
    for (int l = 0; l <= nd.L; ++l) {
        const int kl = nd.kl[l], nl = nd.nl[l], kp = nd.kp[l], np = nd.np[l];
        for (long long i = t0; i < np; i += stride) {
            real gc = i < nl ? th[nd.fg[l] + i] * c : (real)0;
            real bb = i < nl ? th[nd.fb[l] + i] : (real)0;
            if (l == nd.L && i < nl) bb = th[nd.fbias + i] * gc + bb;
            pk[nd.offg[l] + i] = gc;
            pk[nd.offb[l] + i] = bb;
        }
        for (long long i = t0; i < (long long)kp * np; i += stride) {
            int k = (int)(i / np), n = (int)(i - (long long)k * np);
            pk[nd.offW[l] + i] = (k < kl && n < nl) ? th[nd.fW[l] + (long long)k * nl + n] : (real)0;
            int n2 = (int)(i / kp), k2 = (int)(i - (long long)n2 * kp);
            pk[nd.offWT[l] + i] = (k2 < kl && n2 < nl) ? th[nd.fW[l] + (long long)k2 * nl + n2] * (th[nd.fg[l] + n2] * c) : (real)0;
        }
    }
}

// ------------------------------------------------------------------------------------- finalize
// raw[gtotal] = sum over the `nslab` per-CTA slabs (fixed order => deterministic)
template <typename real>
__global__ void reduce_slabs_kernel(const real* __restrict__ slabs, int nslab, long long gtotal, real* __restrict__ raw) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < gtotal; i += stride) {
        real s = (real)0;
        for (int c = 0; c < nslab; ++c) s = s + slabs[(size_t)c * gtotal + i];
        raw[i] = s;
    }
}

template <typename real>
__global__ void finalize_grad_kernel(NetDev nd, const real* __restrict__ th, const real* __restrict__ raw, real* __restrict__ grad, real c) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = t0; i < nd.in; i += stride) {
        grad[nd.fg0 + i] = c * raw[nd.gX + i];
        grad[nd.fb0 + i] = raw[nd.g0 + i];
    }
    for (int l = 0; l <= nd.L; ++l) {
        const int kl = nd.kl[l], nl = nd.nl[l];
        for (long long i = t0; i < (long long)kl * nl; i += stride) {
            int n = (int)(i % nl);
            grad[nd.fW[l] + i] = raw[nd.gW[l] + i] * (th[nd.fg[l] + n] * c);
        }
        for (long long n = t0; n < nl; n += stride) {
            real s = (real)0;
            for (int k = 0; k < kl; ++k) s = fma(th[nd.fW[l] + (long long)k * nl + n], raw[nd.gW[l] + (long long)k * nl + n], s);
            real C = raw[nd.gC[l] + n];
            if (l == nd.L) {
                s = s + th[nd.fbias + n] * C;
                grad[nd.fbias + n] = (th[nd.fg[l] + n] * c) * C;
            }
            grad[nd.fg[l] + n] = c * s;
            grad[nd.fb[l] + n] = C;
        }
    }
}

// ------------------------------------------------------------------------------------- forward
// xin: smem [>= in rows][LDP].  y0: smem [kp[0]][LDP].  hb[l]: output buffer of hidden layer l
// (distinct buffers when the activations are kept for the backward, two alternating ones
// otherwise).  out: smem [np[L]][LDP], the raw network output (before the ekn head).
template <typename real>
__device__ __forceinline__ void net_forward(const NetDev& nd, const real* __restrict__ pk, const real* xin, real* y0,
                                            real* const* hb, real* out, real* Ws) {
    constexpr int TP = RT<real>::TP;
    constexpr int P = 8 * TP, LDP = P + RT<real>::PADP;
    __syncthreads();                                                // xin complete
    {
        const real* g0c = pk + nd.offg0;
        const real* b0 = pk + nd.offb0;
        for (int idx = threadIdx.x; idx < nd.kp[0] * P; idx += NTHREADS) {
            int k = idx / P, p = idx - k * P;
            y0[k * LDP + p] = (k < nd.in) ? xin[k * LDP + p] * g0c[k] + b0[k] : (real)0;      // solver.py:265
        }
    }
    const real* cur = y0;
    for (int l = 0; l < nd.L; ++l) {
        const real* gc = pk + nd.offg[l];
        const real* bb = pk + nd.offb[l];
        real* o = hb[l];
        gemm_AW<real, 8>(pk + nd.offW[l], nd.kp[l], nd.np[l], cur, Ws, [&](int n, int p0, const real* acc) {
            const real g = gc[n], b = bb[n];
            real v[TP];
#pragma unroll
            for (int j = 0; j < TP; ++j) {
                real z = acc[j] * g + b;                                                     // solver.py:267-268
                v[j] = z + dpb_max(z, (real)0);                                              // solver.py:269
            }
            stv<TP>(o + n * LDP + p0, v);
        });
        cur = o;
    }
    {
        const int l = nd.L;
        const real* gc = pk + nd.offg[l];
        const real* bb = pk + nd.offb[l];
        gemm_AW<real, 1>(pk + nd.offW[l], nd.kp[l], nd.np[l], cur, Ws, [&](int n, int p0, const real* acc) {
            const real g = gc[n], b = bb[n];
            real v[TP];
#pragma unroll
            for (int j = 0; j < TP; ++j) v[j] = acc[j] * g + b;                               // solver.py:270-271
            stv<TP>(out + n * LDP + p0, v);
        });
    }
}

// ------------------------------------------------------------------------------------ backward
// dOut: smem [np[L]][LDP] cotangent of the raw output (rows >= out are zero).  hb[0..L-1] hold the
// kept activations, dzA/dzB are two more hidden-size buffers, dy0: smem [kp[0]][LDP].
// gs: this CTA's raw-gradient slab (NULL: no parameter gradients).  dx: smem [>= in rows][LDP]
// receives the input gradient (NULL: not needed).
template <typename real>
__device__ __forceinline__ void net_backward(const NetDev& nd, const real* __restrict__ pk, const real* xin, const real* y0,
                                             real* const* hb, const real* dOut, real* dzA, real* dzB, real* dy0,
                                             real* gs, real* dx, real* Ws) {
    constexpr int TP = RT<real>::TP;
    constexpr int P = 8 * TP, LDP = P + RT<real>::PADP;
    const int L = nd.L;
    __syncthreads();
    if (gs) {
        gemm_dW<real>(hb[L - 1], nd.kl[L], dOut, nd.nl[L], gs + nd.gW[L]);
        colsum_dY<real>(dOut, nd.nl[L], gs + nd.gC[L]);
    }
    const real* dzin = dOut;
    real* dzcur = dzA;
    real* dzoth = dzB;
    for (int l = L; l >= 1; --l) {
        const real* am = hb[l - 1];
        real* o = dzcur;
        gemm_AW<real, 8>(pk + nd.offWT[l], nd.np[l], nd.kp[l], dzin, Ws, [&](int n, int p0, const real* acc) {
            real a[TP], v[TP];
            ldv<TP>(a, am + n * LDP + p0);
#pragma unroll
            for (int j = 0; j < TP; ++j) v[j] = (a[j] > (real)0) ? (real)2 * acc[j] : acc[j];   // d(z + relu z)
            stv<TP>(o + n * LDP + p0, v);
        });
        if (gs) {
            const real* aprev = (l - 1 > 0) ? hb[l - 2] : y0;
            gemm_dW<real>(aprev, nd.kl[l - 1], dzcur, nd.nl[l - 1], gs + nd.gW[l - 1]);
            colsum_dY<real>(dzcur, nd.nl[l - 1], gs + nd.gC[l - 1]);
        }
        dzin = dzcur;
        real* t = dzcur; dzcur = dzoth; dzoth = t;
    }
    if (gs || dx) {
        gemm_AW<real, 1>(pk + nd.offWT[0], nd.np[0], nd.kp[0], dzin, Ws, [&](int n, int p0, const real* acc) {
            stv<TP>(dy0 + n * LDP + p0, acc);
        });
        if (gs) {
            for (int k = threadIdx.x; k < nd.in; k += NTHREADS) {
                real sx = (real)0, s0 = (real)0;
                for (int p = 0; p < P; ++p) {
                    real v = dy0[k * LDP + p];
                    sx = sx + xin[k * LDP + p] * v;
                    s0 = s0 + v;
                }
                atomicAdd(gs + nd.gX + k, sx);
                atomicAdd(gs + nd.g0 + k, s0);
            }
        }
        if (dx) {
            const real* g0c = pk + nd.offg0;
            for (int idx = threadIdx.x; idx < nd.in * P; idx += NTHREADS) {
                int k = idx / P, p = idx - k * P;
                dx[k * LDP + p] = dy0[k * LDP + p] * g0c[k];
            }
        }
    }
    __syncthreads();
}

}  // namespace dpb
