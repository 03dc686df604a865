// host_harness.cpp -- TEST INFRASTRUCTURE (never shipped, never loaded by the package).
// Compiles the __host__ __device__ per-path arithmetic of csrc/dpb_eqn.h with g++ so that the
// step schedule and the reverse recursion can be checked against the oracle on a box without a
// GPU.  The control is the affine map u = A x + b (so that its VJP is exact and trivial).
#include "../../deeppde_actorcritic_b200/csrc/dpb_eqn.h"
#include "../../deeppde_actorcritic_b200/csrc/dpb_host.h"
#include <vector>
#include <cstring>

using namespace dpb;

template <typename real>
static void affine(const real* A, const real* b, const real* x, real* u, int m, int d) {
    for (int j = 0; j < m; ++j) {
        real s = b[j];
        for (int k = 0; k < d; ++k) s = s + A[j * d + k] * x[k];
        u[j] = s;
    }
}

// One path: forward rollout storing the trajectory, actor cost y = sum c w h D + D_N V_true(x_N),
// then the reverse sweep producing dy/dA, dy/db (scaled by invB).  cheat: 1 -> u_true.
template <typename real>
static void run_path(const dpb_config* cfg, int N, double T, const real* A, const real* b, const real* x0, const real* dw /*[d][N]*/,
                     int cheat, real invB, real* xs /*[d][N+1]*/, real* dts, real* coefs, real* y_out, real* gA, real* gb) {
    EqnD ed;
    fill_eqn(*cfg, N, T, ed);
    Eq<real> E(ed);
    const int d = E.d, m = E.m;
    std::vector<real> x(x0, x0 + d), u(32), xi(32), us((size_t)N * 32), Ds(N + 1), sq(N), xn(N);
    std::vector<int> dg(N), cf(N);
    int flag = fwd_initial_flag(E, x.data(), 1, 0);
    real D = 1, y = 0;
    for (int k = 0; k < d; ++k) xs[k * (N + 1)] = x[k];
    for (int t = 0; t < N; ++t) {
        real dt, sqdt, xnorm; int dtg;
        fwd_dt(E, x.data(), flag, 1, 0, dt, sqdt, xnorm, dtg);
        if (cheat) eq_u_true(E, x.data(), u.data(), 1, 0); else affine(A, b, x.data(), u.data(), m, d);
        for (int k = 0; k < d; ++k) xi[k] = dw[k * N + t];
        real w = eq_w(E, x.data(), u.data(), 1, 0);
        for (int j = 0; j < m; ++j) us[(size_t)t * 32 + j] = u[j];
        int c = fwd_move(E, x.data(), u.data(), xi.data(), dt, sqdt, xnorm, flag, (real*)nullptr, 1, 0);
        Ds[t] = D; dts[t] = dt; coefs[t] = (real)c; sq[t] = sqdt; xn[t] = xnorm; dg[t] = dtg; cf[t] = c;
        y = y + (real)c * w * dt * D;                                   // solver.py:218
        D = D * dpb_exp(-E.gamma * dt * (real)c);                       // solver.py:219
        for (int k = 0; k < d; ++k) xs[k * (N + 1) + t + 1] = x[k];
    }
    Ds[N] = D;
    real VN = eq_V_true(E, x.data(), 1, 0);
    y = y + VN * D;                                                     // solver.py:223
    *y_out = y;
    if (!gA) return;
    // reverse sweep (cheat_value seed)
    std::vector<real> lam(32), ubar(32), xt(32);
    eq_V_grad_true(E, x.data(), lam.data(), 1, 0);
    for (int k = 0; k < d; ++k) lam[k] = lam[k] * D * invB;
    real Dbar = VN * invB;
    for (int i = 0; i < m * d; ++i) gA[i] = 0;
    for (int j = 0; j < m; ++j) gb[j] = 0;
    for (int t = N - 1; t >= 0; --t) {
        for (int k = 0; k < d; ++k) { xt[k] = xs[k * (N + 1) + t]; xi[k] = dw[k * N + t]; }
        adj_step(E, xt.data(), &us[(size_t)t * 32], xi.data(), dts[t], sq[t], cf[t], dg[t], xn[t], Ds[t], invB,
                 lam.data(), Dbar, ubar.data(), 1, 0);
        for (int j = 0; j < m; ++j) {
            gb[j] += ubar[j];
            for (int k = 0; k < d; ++k) { gA[j * d + k] += ubar[j] * xt[k]; lam[k] += A[j * d + k] * ubar[j]; }
        }
    }
}

extern "C" {
int hh_run_path_f64(const dpb_config* cfg, int N, double T, const double* A, const double* b, const double* x0, const double* dw,
                    int cheat, double invB, double* xs, double* dts, double* coefs, double* y, double* gA, double* gb) {
    run_path<double>(cfg, N, T, A, b, x0, dw, cheat, invB, xs, dts, coefs, y, gA, gb);
    return 0;
}
int hh_run_path_f32(const dpb_config* cfg, int N, double T, const float* A, const float* b, const float* x0, const float* dw,
                    int cheat, float invB, float* xs, float* dts, float* coefs, float* y, float* gA, float* gb) {
    run_path<float>(cfg, N, T, A, b, x0, dw, cheat, invB, xs, dts, coefs, y, gA, gb);
    return 0;
}
// closed forms at one point: out = [V_true, Z, w(x,u_in), u_true[m], V_grad_true[d]]
int hh_closed_forms_f64(const dpb_config* cfg, const double* x, const double* u_in, double* out) {
    EqnD ed; fill_eqn(*cfg, 10, 1.0, ed);
    Eq<double> E(ed);
    out[0] = eq_V_true(E, x, 1, 0); out[1] = eq_Z(E, x, 1, 0); out[2] = eq_w(E, x, u_in, 1, 0);
    eq_u_true(E, x, out + 3, 1, 0);
    eq_V_grad_true(E, x, out + 3 + E.m, 1, 0);
    return 0;
}
// ekn head forward/backward at one point
int hh_ekn_head_f64(const double* y, const double* ubar, int m, double* u, double* ybar) {
    ekn_head_fwd(y, u, m, 1, 0);
    ekn_head_bwd(y, ubar, ybar, m, 1, 0);
    return 0;
}
}
