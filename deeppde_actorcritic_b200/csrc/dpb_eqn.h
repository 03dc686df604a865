// dpb_eqn.h -- per-path arithmetic of the rollout and of its reverse sweep.
//
//  * Eq<real>: closed forms of the four equations (reference equation.py:144-311) and the partial
//    derivatives the actor's reverse sweep needs.
//  * fwd_*: one Euler-Maruyama step of propagate_naive / propagate_adaptive (equation.py:46-106)
//    and the running sums of CriticModel.call / ActorModel.call (solver.py:159-224).
//  * adj_step: one step of the reverse recursion of d mean(y) / d (path state).
//
// Everything here is __host__ __device__ and addresses per-path vectors as v[k*ld + p] (column p of
// a [feature][path] array), so the very same code runs in the CUDA kernels (ld = padded tile
// width) and in the g++-compiled test harness of tests/ (ld = 1, p = 0).
//
// Compiled with -fmad=false / -ffp-contract=off: elementwise arithmetic is uncontracted IEEE in the
// written order, so that the step schedule (dt, coef, exit index) is reproducible bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DPB_HD __host__ __device__ __forceinline__
#define DPB_UNROLL _Pragma("unroll (DMAX > 0 ? DMAX : 4)")
#else
#define DPB_HD inline
#define DPB_UNROLL
#endif
// Loop over the first n components.  With a compile-time bound DMAX > 0 (the tensor-path kernels are
// instantiated per padded dimension) the loop is fully unrolled and guarded, so that the per-path vectors
// live in registers; DMAX == 0 is the generic run-time loop.
#define DPB_LOOP(k, n) DPB_UNROLL for (int k = 0; k < (DMAX > 0 ? DMAX : (n)); ++k) if (DMAX == 0 || k < (n))
#define DPB_TPL template <typename real, int DMAX = 0, int EQN = -1, int MV = 0>
// MV > 0: control_dim of the VDP instantiations (makes the cyclic neighbour indices of equation.py:192-195 static)
#define DPB_EQN(E) (EQN >= 0 ? EQN : (E).eqn)

namespace dpb {

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
    // scheme constants (equation.py:48-49,75,80,85-86); filled per call from (T, N)
    double delta_t, sqrt_delta_t, hb, c3, hmin;
};

enum { EQ_LQR = 0, EQ_VDP = 1, EQ_EKN = 2, EQ_LQRVAR = 3 };
enum { SCHEME_NAIVE = 0, SCHEME_ADAPTIVE = 1 };

template <typename real>
struct Eq {
    int eqn, d, m, scheme, td;
    real R, R2, gamma, sig, p, q, beta, k, a, eps, a2, a3, cu, wconst, ZR, C0;
    real lv_num, lv_den, lv_un, lv_ud, lv_ue, lv_gk;
    real delta_t, sqrt_delta_t, hb, c3, hmin;
    Eq() = default;
    DPB_HD explicit Eq(const EqnD& e)
        : eqn(e.eqn), d(e.d), m(e.m), scheme(e.scheme), td(e.td),
          R((real)e.R), R2((real)e.R2), gamma((real)e.gamma), sig((real)e.sig), p((real)e.p), q((real)e.q),
          beta((real)e.beta), k((real)e.k), a((real)e.a), eps((real)e.eps), a2((real)e.a2), a3((real)e.a3),
          cu((real)e.cu), wconst((real)e.wconst), ZR((real)e.ZR), C0((real)e.C0),
          lv_num((real)e.lv_num), lv_den((real)e.lv_den), lv_un((real)e.lv_un), lv_ud((real)e.lv_ud),
          lv_ue((real)e.lv_ue), lv_gk((real)e.lv_gk),
          delta_t((real)e.delta_t), sqrt_delta_t((real)e.sqrt_delta_t), hb((real)e.hb), c3((real)e.c3),
          hmin((real)e.hmin) {}
};

DPB_HD float dpb_sqrt(float x) { return sqrtf(x); }
DPB_HD double dpb_sqrt(double x) { return sqrt(x); }
DPB_HD float dpb_exp(float x) { return expf(x); }
DPB_HD double dpb_exp(double x) { return exp(x); }
DPB_HD float dpb_max(float a, float b) { return fmaxf(a, b); }
DPB_HD double dpb_max(double a, double b) { return fmax(a, b); }

// All per-path functions address column arrays as v[k*ld + p].
#define DPB_AT(v, k) ((v)[(k) * ld + p])

DPB_TPL
DPB_HD real norm2_path(const real* x, int d, int ld, int p) {
    real s = (real)0;
    DPB_LOOP(k, d) s = s + DPB_AT(x, k) * DPB_AT(x, k);
    return s;
}

// u_true (equation.py:163-164, 212-217, 259-261, 298-299)
DPB_TPL
DPB_HD void eq_u_true(const Eq<real>& E, const real* x, real* u, int ld, int p) {
    const int d = E.d, m = (MV > 0 ? MV : E.m);
    switch (DPB_EQN(E)) {
    case EQ_LQR:
        DPB_LOOP(k, d) DPB_AT(u, k) = E.cu * DPB_AT(x, k);
        break;
    case EQ_VDP:
        DPB_LOOP(j, m) {
            real x2 = DPB_AT(x, m + j);
            real px2 = DPB_AT(x, m + (j + 1 == m ? 0 : j + 1));
            real nx2 = DPB_AT(x, m + (j == 0 ? m - 1 : j - 1));
            DPB_AT(u, j) = -((real)2 * E.a * x2 - E.eps * (px2 + nx2)) / (real)2 / E.q;
        }
        break;
    case EQ_EKN: {
        real r = dpb_sqrt(norm2_path<real, DMAX, EQN, MV>(x, d, ld, p));
        DPB_LOOP(k, d) DPB_AT(u, k) = DPB_AT(x, k) / r;
        break;
    }
    default:
        DPB_LOOP(k, d) {
            real xk = DPB_AT(x, k);
            DPB_AT(u, k) = -E.lv_un * xk / (E.lv_ud + E.lv_ue * xk * xk);
        }
    }
}

// V_true (equation.py:160,204-210,255-257,295)
DPB_TPL
DPB_HD real eq_V_true(const Eq<real>& E, const real* x, int ld, int p) {
    const int d = E.d, m = (MV > 0 ? MV : E.m);
    real n2 = norm2_path<real, DMAX, EQN, MV>(x, d, ld, p);
    switch (DPB_EQN(E)) {
    case EQ_LQR:
    case EQ_LQRVAR:
        return n2 * E.k;
    case EQ_VDP: {
        real s = (real)0;
        DPB_LOOP(j, m) {
            int jn = (j + 1 == m ? 0 : j + 1);
            s = s + (DPB_AT(x, j) * DPB_AT(x, jn) + DPB_AT(x, m + j) * DPB_AT(x, m + jn));
        }
        return E.a * n2 - E.eps * s;
    }
    default: {
        real r = dpb_sqrt(n2);
        return E.a3 * r * r * r - E.a2 * r * r;
    }
    }
}

// Z_tf on the boundary (equation.py:157,201,252,292)
DPB_TPL
DPB_HD real eq_Z(const Eq<real>& E, const real* x, int ld, int p) {
    if (DPB_EQN(E) == EQ_LQR || DPB_EQN(E) == EQ_LQRVAR) return E.ZR;
    return eq_V_true<real, DMAX, EQN, MV>(E, x, ld, p);
}

// V_grad_true (equation.py:166,219-227,263-265,301) -> g[k]
DPB_TPL
DPB_HD void eq_V_grad_true(const Eq<real>& E, const real* x, real* g, int ld, int p) {
    const int d = E.d, m = (MV > 0 ? MV : E.m);
    switch (DPB_EQN(E)) {
    case EQ_LQR:
    case EQ_LQRVAR:
        DPB_LOOP(k, d) DPB_AT(g, k) = (real)2 * E.k * DPB_AT(x, k);
        break;
    case EQ_VDP:
        DPB_LOOP(j, m) {
            int jn = (j + 1 == m ? 0 : j + 1), jp = (j == 0 ? m - 1 : j - 1);
            DPB_AT(g, j) = (real)2 * E.a * DPB_AT(x, j) - E.eps * (DPB_AT(x, jn) + DPB_AT(x, jp));
            DPB_AT(g, m + j) = (real)2 * E.a * DPB_AT(x, m + j) - E.eps * (DPB_AT(x, m + jn) + DPB_AT(x, m + jp));
        }
        break;
    default: {
        real r = dpb_sqrt(norm2_path<real, DMAX, EQN, MV>(x, d, ld, p));
        real c = (real)3 * E.a3 * r - (real)2 * E.a2;
        DPB_LOOP(k, d) DPB_AT(g, k) = c * DPB_AT(x, k);
    }
    }
}

// running cost w_tf (equation.py:154,188-199,249,288-290)
DPB_TPL
DPB_HD real eq_w(const Eq<real>& E, const real* x, const real* u, int ld, int p) {
    const int d = E.d, m = (MV > 0 ? MV : E.m);
    switch (DPB_EQN(E)) {
    case EQ_LQR: {
        real s1 = (real)0, s2 = (real)0;
        DPB_LOOP(k, d) {
            s1 = s1 + E.p * (DPB_AT(x, k) * DPB_AT(x, k));
            s2 = s2 + E.q * (DPB_AT(u, k) * DPB_AT(u, k));
        }
        return s1 + s2 - E.wconst;
    }
    case EQ_VDP: {
        real s = (real)0, n2 = (real)0;
        DPB_LOOP(j, m) {
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
        DPB_LOOP(k, d) {
            real xk = DPB_AT(x, k), uk = DPB_AT(u, k);
            s1 = s1 + E.lv_num * (xk * xk) / (E.q + E.lv_den * (xk * xk));
            s2 = s2 + (E.lv_gk * (xk * xk) + E.q * (uk * uk));
        }
        return s1 + s2 - E.wconst;
    }
    }
}

// ekn drift factor c(|x|) (equation.py:270-272); 0 for the other equations.
DPB_TPL
DPB_HD real eq_drift_c(const Eq<real>& E, real r) {
    return E.C0 / ((real)2 * E.a2 - (real)3 * E.a3 * r);
}

// drift component k (equation.py:172,232-235,270-273,307).  cc = eq_drift_c<real, DMAX, EQN, MV>(|x|) for ekn.
DPB_TPL
DPB_HD real eq_drift(const Eq<real>& E, real cc, const real* x, const real* u, int k, int ld, int p) {
    switch (DPB_EQN(E)) {
    case EQ_LQR:
    case EQ_LQRVAR:
        return E.beta * DPB_AT(u, k);
    case EQ_VDP: {
        const int m = (MV > 0 ? MV : E.m);
        if (k < m) return DPB_AT(x, m + k);
        int j = k - m;
        real x1 = DPB_AT(x, j), x2 = DPB_AT(x, m + j);
        return ((real)1 - x1 * x1) * x2 - x1 + DPB_AT(u, j);
    }
    default:
        return cc * DPB_AT(u, k);
    }
}

// diagonal of sigma, component k (equation.py:170,230,268,305)
DPB_TPL
DPB_HD real eq_sigma(const Eq<real>& E, const real* x, const real* u, int k, int ld, int p) {
    if (DPB_EQN(E) == EQ_LQRVAR) return E.sig * ((real)1 + E.eps * DPB_AT(x, k) * DPB_AT(u, k));
    return E.sig;
}

// adaptive-scheme flag of a point with norm nrm (equation.py:80-82,94-95):
//   1 + floor((sign(R-n-hb) + sign(R-n))/2)  ==  (R-n > 0) ? ((R-n-hb > 0) ? 2 : 1) : 0
DPB_TPL
DPB_HD int eq_flag(const Eq<real>& E, real nrm) {
    real t2 = E.R - nrm;
    real t1 = E.R - nrm - E.hb;
    return (t2 > (real)0) ? ((t1 > (real)0) ? 2 : 1) : 0;
}

// ------------------------------------------------------------------------------------------------
// Forward step.  `flag`: naive 1 = alive, 0 = frozen (equation.py:51,69); adaptive 2 = inner,
// 1 = boundary layer, 0 = out (equation.py:80-82).
DPB_TPL
DPB_HD int fwd_initial_flag(const Eq<real>& E, const real* x, int ld, int p) {
    if (E.scheme == SCHEME_NAIVE) return 1;
    return eq_flag<real, DMAX, EQN, MV>(E, dpb_sqrt(norm2_path<real, DMAX, EQN, MV>(x, E.d, ld, p)));
}

// step size of this step (equation.py:49 | 84-86).  `clamped` tells the reverse sweep whether the
// maximum() selected the constant (zero gradient).
DPB_TPL
DPB_HD void fwd_dt(const Eq<real>& E, const real* x, int flag, int ld, int p, real& dt, real& sqdt, real& xnorm, int& dt_grad) {
    dt_grad = 0;
    xnorm = (real)0;
    if (E.scheme == SCHEME_NAIVE) {
        dt = E.delta_t;
        sqdt = E.sqrt_delta_t;
        return;
    }
    xnorm = dpb_sqrt(norm2_path<real, DMAX, EQN, MV>(x, E.d, ld, p));
    if (flag == 1) {
        real g = E.R - xnorm;
        dt = g * g / E.c3;
        dt_grad = 1;
    } else {
        dt = E.delta_t;
    }
    if (!(dt > E.hmin)) { dt = E.hmin; dt_grad = 0; }
    sqdt = dpb_sqrt(dt);
}

// Proposal + exit logic + in-place state update (equation.py:58-69 | 91-105).
//   x, u, dw: columns;  xdw_out (optional): sigma_k*dw_k per component (the diffusion vector).
// Returns coef in {0,1}; updates x (only if coef) and flag.
DPB_TPL
DPB_HD int fwd_move(const Eq<real>& E, real* x, const real* u, const real* dw, real dt, real sqdt, real xnorm, int& flag,
                    real* sdw_out, int ld, int p) {
    const int d = E.d;
    real cc = (real)0;
    if (DPB_EQN(E) == EQ_EKN) {
        real r = (E.scheme == SCHEME_ADAPTIVE) ? xnorm : dpb_sqrt(norm2_path<real, DMAX, EQN, MV>(x, d, ld, p));
        cc = eq_drift_c<real, DMAX, EQN, MV>(E, r);
    }
    real n2 = (real)0;
    DPB_LOOP(k, d) {
        real sd = eq_sigma<real, DMAX, EQN, MV>(E, x, u, k, ld, p) * DPB_AT(dw, k);
        if (sdw_out) DPB_AT(sdw_out, k) = sd;
        real dk = eq_drift<real, DMAX, EQN, MV>(E, cc, x, u, k, ld, p) * dt + sd * sqdt;
        real pk = DPB_AT(x, k) + dk;
        n2 = n2 + pk * pk;
    }
    int coef, newflag;
    if (E.scheme == SCHEME_NAIVE) {
        int ex = (n2 - E.R2 >= (real)0) ? 1 : 0;            // ceil((sign(b)+1)/2)
        coef = flag * (1 - ex);
        newflag = coef;
    } else {
        int nf = eq_flag<real, DMAX, EQN, MV>(E, dpb_sqrt(n2));
        newflag = (flag > 0) ? nf : 0;                      // flag(p) * sign(flag)
        coef = (flag > 0 && newflag > 0) ? 1 : 0;           // sign(flag) * sign(new_flag)
    }
    if (coef) {
        // the increment is evaluated again (same expression, same bits) rather than kept in d registers; every
        // component reads the OLD state, so the new one is staged before it is written back
        real xn[DMAX > 0 ? DMAX : 32];
        DPB_LOOP(k, d) {
            real sd = eq_sigma<real, DMAX, EQN, MV>(E, x, u, k, ld, p) * DPB_AT(dw, k);
            xn[k] = DPB_AT(x, k) + (eq_drift<real, DMAX, EQN, MV>(E, cc, x, u, k, ld, p) * dt + sd * sqdt);
        }
        DPB_LOOP(k, d) DPB_AT(x, k) = xn[k];
    }
    flag = newflag;
    return coef;
}

// ------------------------------------------------------------------------------------------------
// Reverse step of d mean(y)/d(state).  Inputs: state at the START of
// step t (x, u, dw), the step's (dt, sqdt, coef, dt_grad), D_t, and the adjoints of the state
// AFTER the step: lam[k] (in/out -> becomes the direct part xbar), Dbar (in/out).
// Outputs ubar[j] (cotangent of the control).  invB = 1/B_global.
DPB_TPL
DPB_HD void adj_step(const Eq<real>& E, const real* x, const real* u, const real* dw, real dt, real sqdt, int coef,
                     int dt_grad, real xnorm, real D_t, real invB, real* lam, real& Dbar, real* ubar, int ld, int p) {
    const int d = E.d, m = (MV > 0 ? MV : E.m);
    if (!coef) {                                            // identity step: contributes nothing
        DPB_LOOP(j, m) DPB_AT(ubar, j) = (real)0;
        return;
    }
    const real w = eq_w<real, DMAX, EQN, MV>(E, x, u, ld, p);
    const real ed = dpb_exp(-E.gamma * dt);
    const real D_next = D_t * ed;
    const real cw = dt * D_t * invB;                        // weight of dw/d(.) terms
    real cc = (real)0, r = (real)0;
    if (DPB_EQN(E) == EQ_EKN) {
        r = (E.scheme == SCHEME_ADAPTIVE) ? xnorm : dpb_sqrt(norm2_path<real, DMAX, EQN, MV>(x, d, ld, p));
        cc = eq_drift_c<real, DMAX, EQN, MV>(E, r);
    }
    // hbar = D_t w /B - gamma Dbar D_{t+1} + <lam, mu + s*xi/(2 sqrt h)>
    real hbar = (real)0;
    if (dt_grad) {
        real acc = (real)0;
        DPB_LOOP(k, d) {
            real mu = eq_drift<real, DMAX, EQN, MV>(E, cc, x, u, k, ld, p);
            real s = eq_sigma<real, DMAX, EQN, MV>(E, x, u, k, ld, p);
            acc = acc + DPB_AT(lam, k) * (mu + s * DPB_AT(dw, k) / ((real)2 * sqdt));
        }
        hbar = D_t * w * invB - E.gamma * Dbar * D_next + acc;
    }
    // Dbar_t = Dbar * exp(-gamma h) + h w / B
    Dbar = Dbar * ed + dt * w * invB;

    // ubar and the direct part of xbar
    real xb[DMAX > 0 ? DMAX : 32];
    DPB_LOOP(k, d) xb[k] = DPB_AT(lam, k);
    switch (DPB_EQN(E)) {
    case EQ_LQR:
        DPB_LOOP(k, d) {
            DPB_AT(ubar, k) = cw * ((real)2 * E.q * DPB_AT(u, k)) + dt * E.beta * DPB_AT(lam, k);
            xb[k] = xb[k] + cw * ((real)2 * E.p * DPB_AT(x, k));
        }
        break;
    case EQ_LQRVAR:
        DPB_LOOP(k, d) {
            real xk = DPB_AT(x, k), uk = DPB_AT(u, k), lk = DPB_AT(lam, k), xi = DPB_AT(dw, k);
            real den = E.q + E.lv_den * xk * xk;
            real dwdx = E.lv_num * (real)2 * xk * E.q / (den * den) + (real)2 * E.lv_gk * xk;
            DPB_AT(ubar, k) = cw * ((real)2 * E.q * uk) + dt * E.beta * lk + sqdt * (E.sig * E.eps * xk) * xi * lk;
            xb[k] = xb[k] + cw * dwdx + sqdt * (E.sig * E.eps * uk) * xi * lk;
        }
        break;
    case EQ_EKN: {
        real lu = (real)0;
        DPB_LOOP(k, d) lu = lu + DPB_AT(lam, k) * DPB_AT(u, k);
        real dc = cc * (real)3 * E.a3 / ((real)2 * E.a2 - (real)3 * E.a3 * r);      // dc/dr
        DPB_LOOP(k, d) {
            DPB_AT(ubar, k) = dt * cc * DPB_AT(lam, k);
            xb[k] = xb[k] + dt * lu * dc * DPB_AT(x, k) / r;
        }
        break;
    }
    default: {                                              // VDP
        real dv1[16], dv2[16], f[16];
        DPB_LOOP(j, m) {
            int jn = (j + 1 == m ? 0 : j + 1), jp = (j == 0 ? m - 1 : j - 1);
            real x1 = DPB_AT(x, j), x2 = DPB_AT(x, m + j);
            dv1[j] = (real)2 * E.a * x1 - E.eps * (DPB_AT(x, jn) + DPB_AT(x, jp));
            dv2[j] = (real)2 * E.a * x2 - E.eps * (DPB_AT(x, m + jn) + DPB_AT(x, m + jp));
            f[j] = ((real)1 - x1 * x1) * x2 - x1;
        }
        DPB_LOOP(j, m) {
            int jn = (j + 1 == m ? 0 : j + 1), jp = (j == 0 ? m - 1 : j - 1);
            real x1 = DPB_AT(x, j), x2 = DPB_AT(x, m + j);
            real l1 = DPB_AT(lam, j), l2 = DPB_AT(lam, m + j);
            real dw1 = -E.gamma * E.eps * (DPB_AT(x, jn) + DPB_AT(x, jp)) - dv2[j]
                       + ((real)2 * x1 * x2 + (real)1) * dv2[j] + (real)2 * E.gamma * E.a * x1;
            real dw2 = -E.gamma * E.eps * (DPB_AT(x, m + jn) + DPB_AT(x, m + jp))
                       + ((real)2 * E.a * dv2[j] - E.eps * (dv2[jp] + dv2[jn])) / ((real)2 * E.q)
                       - dv1[j] - ((real)1 - x1 * x1) * dv2[j]
                       - ((real)2 * E.a * f[j] - E.eps * (f[jp] + f[jn])) + (real)2 * E.gamma * E.a * x2;
            DPB_AT(ubar, j) = cw * ((real)2 * E.q * DPB_AT(u, j)) + dt * l2;
            xb[j] = xb[j] + cw * dw1 + dt * (l2 * (-(real)2 * x1 * x2 - (real)1));
            xb[m + j] = xb[m + j] + cw * dw2 + dt * (l1 + ((real)1 - x1 * x1) * l2);
        }
    }
    }
    if (dt_grad) {
        // dh/dx = -2 (R-|x|)/(3 d sigU^2) * x/|x|      (equation.py:85)
        real g = -(real)2 * (E.R - xnorm) / E.c3 / xnorm * hbar;
        DPB_LOOP(k, d) xb[k] = xb[k] + g * DPB_AT(x, k);
    }
    DPB_LOOP(k, d) DPB_AT(lam, k) = xb[k];
}

// Gradient of clipped square rho (solver.py:76-77): 2 z if |z| < 50 else 100 sign(z).
DPB_TPL
DPB_HD real rho(real z, real clip) {
    real az = z < (real)0 ? -z : z;
    return az < clip ? z * z : (real)2 * clip * az - clip * clip;
}
DPB_TPL
DPB_HD real rho_grad(real z, real clip) {
    real az = z < (real)0 ? -z : z;
    if (az < clip) return (real)2 * z;
    return z > (real)0 ? (real)2 * clip : (z < (real)0 ? -(real)2 * clip : (real)0);
}

// ekn actor head (solver.py:272-274): u = y[:m] / (1e-15 + relu(y[m]) + ||y[:m]||)
DPB_TPL
DPB_HD void ekn_head_fwd(const real* y, real* u, int m, int ld, int p) {
    real n = dpb_sqrt(norm2_path<real, DMAX, EQN, MV>(y, m, ld, p));
    real ym = DPB_AT(y, m);
    real D = (real)0.000000000000001 + (ym > (real)0 ? ym : (real)0) + n;
    DPB_LOOP(k, m) DPB_AT(u, k) = DPB_AT(y, k) / D;
}
// cotangent ubar[m] -> ybar[m+1] (in place allowed when ubar and ybar are distinct buffers)
DPB_TPL
DPB_HD void ekn_head_bwd(const real* y, const real* ubar, real* ybar, int m, int ld, int p) {
    real n = dpb_sqrt(norm2_path<real, DMAX, EQN, MV>(y, m, ld, p));
    real ym = DPB_AT(y, m);
    real D = (real)0.000000000000001 + (ym > (real)0 ? ym : (real)0) + n;
    real s = (real)0;
    DPB_LOOP(k, m) s = s + DPB_AT(ubar, k) * DPB_AT(y, k);
    real c = s / (D * D);
    DPB_LOOP(k, m) DPB_AT(ybar, k) = DPB_AT(ubar, k) / D - c * (DPB_AT(y, k) / n);
    DPB_AT(ybar, m) = (ym > (real)0) ? -c : (real)0;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  counter = (path_lo, path_hi, step, chunk), key = seed
// xor-folded with the stream id.
DPB_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}

DPB_HD void philox_key(uint64_t seed, uint64_t stream, uint32_t& k0, uint32_t& k1) {
    uint64_t s = stream * 0x9E3779B97F4A7C15ull;
    k0 = (uint32_t)seed ^ (uint32_t)(s >> 32);
    k1 = (uint32_t)(seed >> 32) ^ (uint32_t)s;
}

// 3-point law {-sqrt3, 0, +sqrt3} with probabilities 1/6, 4/6, 1/6 (equation.py:31-32):
// randint(6) via multiply-high; floor((k-1)/4): k=0 -> -1, k=1..4 -> 0, k=5 -> +1.
DPB_HD float philox_bounded(uint32_t r) {
    uint32_t k = (uint32_t)(((uint64_t)r * 6u) >> 32);
    return (k == 0u) ? -1.7320508075688772f : ((k == 5u) ? 1.7320508075688772f : 0.0f);
}
DPB_HD float philox_u01(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)

}  // namespace dpb
