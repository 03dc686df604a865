// dpb_kernels.cuh -- the fused rollout + TD kernels of the exact path.
//
//   critic_kernel: CriticModel.call + loss_critic + grad_critic (reference solver.py:73-78,85-90,
//                  159-191) for a shard of paths.  Per tile of P paths: sweep 1 rolls the SDE out under
//                  the actor (equation.py:46-106), evaluates NN_value_grad along the trajectory and
//                  forms delta; then NN_value at x_0, x_N, x_bdry forward+backward; then sweep 2
//                  re-evaluates NN_value_grad at the stored x_t and back-propagates the per-step
//                  cotangent rho'(delta) * coef*sqrt(dt)*disc*(sigma.dw) (two sweeps because every
//                  step's cotangent needs the path's delta).
//   actor_kernel:  ActorModel.call + loss_actor + grad_actor (solver.py:80-83,92-97,207-224): forward
//                  rollout storing (x_t, dw_t, dt, coef, D_t), then the reverse sweep of SURVEY 3.4
//                  with the actor re-evaluated (activations recomputed, never stored in HBM).
//
// Only x_t / dw_t / a few scalars per step go through the per-CTA scratch (L2-resident), never
// activations.  Parameter gradients accumulate in per-CTA raw slabs (dpb_nets.cuh).
#pragma once
#include "dpb_nets.cuh"

namespace dpb {

enum { F_CHEAT_CONTROL = 1, F_CHEAT_VALUE = 2, F_NEED_GRAD = 4, F_PROPAGATE_ONLY = 8 };
enum { DW_EXTERNAL = 0, DW_PHILOX_NORMAL = 1, DW_PHILOX_BOUNDED = 2 };

// rows of the per-path scalar array S
enum { S_FLAG = 0, S_DT, S_SQDT, S_XN, S_DTG, S_COEF, S_DISC, S_Y, S_V0, S_VN, S_DELTA, S_RHOG, S_DBAR, S_NACC, S_VALID, S_TMP, S_ROWS };
// scalars per step in the actor scratch
enum { A_DT = 0, A_SQDT, A_COEF, A_DISC, A_XN, A_DTG, A_NSCAL = 8 };

template <typename real>
struct StepArgs {
    EqnD eq;
    NetDev nA, nV, nG;
    const real *pkA, *pkV, *pkG;
    const real *x0, *dw, *xb;
    int dw_mode;
    unsigned long long seed, stream;
    const unsigned long long* stream_base;   // optional device word added to `stream` (CUDA-graph replays: the iteration lives in memory)
    long long B_local, path_offset;
    real invB;                      // 1 / B_global
    int N;
    unsigned flags;
    real* loss_part;                // [grid][2]
    real *slabV, *slabG, *slabA;    // per-CTA raw-gradient slabs
    real* scratch;                  // per-CTA trajectory scratch
    long long scratch_per_cta;      // elements
    int sr;                         // rows of the small arrays: round8(max(dim, control_dim + 1))
    int hrows;                      // rows of a hidden buffer
    int nhb;                        // hidden buffers available
    real *o_x, *o_dt, *o_coef, *o_delta, *o_delta_b;
    int* o_exit;
};

template <typename real>
struct Carve {
    real *Ws, *X, *U, *DW, *SDW, *GO, *OUTA, *Y0, *DY0, *DOUT, *LAM, *S, *red;
    real* hb[MAXLIN + 2];
    real* hpp[MAXLIN];              // ping-pong view for forward-only evaluations
};

template <typename real>
__host__ __device__ inline size_t carve_elems(int sr, int hrows, int nhb) {
    constexpr int LDP = 8 * RT<real>::TP + RT<real>::PADP;
    return (size_t)2 * RT<real>::KC * WS_NMAX + (size_t)10 * sr * LDP + (size_t)S_ROWS * LDP + 32 + (size_t)nhb * hrows * LDP;
}

template <typename real>
__device__ __forceinline__ void carve(Carve<real>& c, real* base, int sr, int hrows, int nhb) {
    constexpr int LDP = 8 * RT<real>::TP + RT<real>::PADP;
    real* p = base;
    c.Ws = p; p += 2 * RT<real>::KC * WS_NMAX;
    real** small[10] = {&c.X, &c.U, &c.DW, &c.SDW, &c.GO, &c.OUTA, &c.Y0, &c.DY0, &c.DOUT, &c.LAM};
    for (int i = 0; i < 10; ++i) { *small[i] = p; p += sr * LDP; }
    c.S = p; p += S_ROWS * LDP;
    c.red = p; p += 32;
    for (int i = 0; i < MAXLIN + 2; ++i) c.hb[i] = nullptr;
    for (int i = 0; i < nhb; ++i) { c.hb[i] = p; p += hrows * LDP; }
    for (int i = 0; i < MAXLIN; ++i) c.hpp[i] = c.hb[i & 1];
}

// tile column load: dst[k][p] = src[(base+p)*d + k] (valid paths) else fill
template <typename real>
__device__ __forceinline__ void load_cols(real* dst, const real* __restrict__ src, long long base, int nvalid, int d, int rows, real fill) {
    constexpr int P = 8 * RT<real>::TP, LDP = P + RT<real>::PADP;
    for (int idx = threadIdx.x; idx < rows * P; idx += NTHREADS) {
        int p = idx / rows, k = idx - p * rows;                     // k fastest: contiguous in global memory
        real v = (real)0;
        if (k < d) v = (p < nvalid) ? src[(base + p) * d + k] : fill;
        dst[k * LDP + p] = v;
    }
}

// Brownian increments of step t for the tile (equation.py:19 | 31-32): external tensor dw[B][d][N]
// or Philox4x32-10 keyed by (seed, stream) with counter (global path, step, component/4).
template <typename real>
__device__ __forceinline__ void load_dw(const StepArgs<real>& a, long long base, int nvalid, int t, real* DW) {
    constexpr int P = 8 * RT<real>::TP, LDP = P + RT<real>::PADP;
    const int d = a.eq.d;
    if (a.dw_mode == DW_EXTERNAL) {
        for (int idx = threadIdx.x; idx < d * P; idx += NTHREADS) {
            int p = idx / d, k = idx - p * d;
            DW[k * LDP + p] = (p < nvalid) ? a.dw[((base + p) * d + k) * (long long)a.N + t] : (real)0;
        }
    } else {
        const int nch = (d + 3) >> 2;
        uint32_t k0, k1;
        philox_key(a.seed, a.stream + (a.stream_base ? *a.stream_base : 0ull), k0, k1);
        for (int idx = threadIdx.x; idx < nch * P; idx += NTHREADS) {
            int ch = idx / P, p = idx - ch * P;
            unsigned long long gp = (unsigned long long)(a.path_offset + base + p);
            uint32_t c[4] = {(uint32_t)gp, (uint32_t)(gp >> 32), (uint32_t)t, (uint32_t)ch};
            philox4x32_10(c, k0, k1);
            float o[4];
            if (a.dw_mode == DW_PHILOX_BOUNDED) {
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = philox_bounded(c[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 2; ++i) {                                       // Box-Muller
                    float r = sqrtf(-2.0f * logf(philox_u01(c[2 * i])));
                    float s, co;
                    sincospif(2.0f * philox_u01(c[2 * i + 1]), &s, &co);
                    o[2 * i] = r * co;
                    o[2 * i + 1] = r * s;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (4 * ch + i < d) DW[(4 * ch + i) * LDP + p] = (real)o[i];
        }
    }
}

// u = NN_control(x) or u_true(x) for the tile (solver.py:153-157); keep: activations in c.hb[0..L-1]
template <typename real>
__device__ __forceinline__ void eval_control(const StepArgs<real>& a, const Eq<real>& E, Carve<real>& c, bool cheat, bool keep) {
    constexpr int P = 8 * RT<real>::TP, LDP = P + RT<real>::PADP;
    if (cheat) {
        __syncthreads();
        if (threadIdx.x < P) eq_u_true(E, c.X, c.U, LDP, threadIdx.x);
    } else {
        real* const* hb = keep ? c.hb : c.hpp;
        if (a.nA.ekn_head) {
            net_forward<real>(a.nA, a.pkA, c.X, c.Y0, hb, c.OUTA, c.Ws);
            if (threadIdx.x < P) ekn_head_fwd(c.OUTA, c.U, a.nA.mctrl, LDP, threadIdx.x);
        } else {
            net_forward<real>(a.nA, a.pkA, c.X, c.Y0, hb, c.U, c.Ws);
        }
    }
    __syncthreads();
}

template <typename real>
__device__ __forceinline__ void zero_rows(real* buf, int rows) {
    constexpr int LDP = 8 * RT<real>::TP + RT<real>::PADP;
    for (int idx = threadIdx.x; idx < rows * LDP; idx += NTHREADS) buf[idx] = (real)0;
}

// ================================================================================== critic kernel
template <typename real>
__global__ void __launch_bounds__(NTHREADS, 1) critic_kernel(const StepArgs<real> a) {
    constexpr int P = 8 * RT<real>::TP, LDP = P + RT<real>::PADP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carve<real> c;
    carve<real>(c, reinterpret_cast<real*>(smem_raw), a.sr, a.hrows, a.nhb);
    {
        real* base = reinterpret_cast<real*>(smem_raw);
        const size_t n = carve_elems<real>(a.sr, a.hrows, a.nhb);
        for (size_t i = threadIdx.x; i < n; i += NTHREADS) base[i] = (real)0;
    }
    __syncthreads();
    const Eq<real> E(a.eq);
    const int d = E.d, N = a.N, sr = a.sr;
    const int tid = threadIdx.x;
    const bool cheat = a.flags & F_CHEAT_CONTROL, need_grad = a.flags & F_NEED_GRAD, prop_only = a.flags & F_PROPAGATE_ONLY;
    const bool td1 = (E.td == 1) && !prop_only;
    const real scale = (real)100 * a.invB;                       // d loss / d mean(rho)   (solver.py:76-78)
    const real fill = (real)0.5 * E.R / dpb_sqrt((real)d);       // padded paths sit at an interior point
    real* traj = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;      // [N][2*sr][P]
    real* gsV = a.slabV ? a.slabV + (size_t)blockIdx.x * a.nV.gtotal : nullptr;
    real* gsG = a.slabG ? a.slabG + (size_t)blockIdx.x * a.nG.gtotal : nullptr;
    real loss0 = (real)0, loss1 = (real)0;
    real* S = c.S;

    const long long ntiles = (a.B_local + P - 1) / P;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * P;
        const int nvalid = (int)((a.B_local - base < P) ? (a.B_local - base) : P);
        __syncthreads();
        load_cols<real>(c.X, a.x0, base, nvalid, d, sr, fill);
        __syncthreads();
        if (tid < P) {
            const int p = tid;
            S[S_FLAG * LDP + p] = (real)fwd_initial_flag(E, c.X, LDP, p);
            S[S_DISC * LDP + p] = (real)1;
            S[S_Y * LDP + p] = (real)0;
            S[S_NACC * LDP + p] = (real)0;
            S[S_VALID * LDP + p] = (p < nvalid) ? (real)1 : (real)0;
            if (a.o_x && p < nvalid)
                for (int k = 0; k < d; ++k) a.o_x[((base + p) * d + k) * (long long)(N + 1)] = c.X[k * LDP + p];
        }
        // ------------------------------------------------------------------ sweep 1: rollout
        int tlive = 0;
        for (int t = 0; t < N; ++t) {
            const int alive = __syncthreads_or(tid < P && tid < nvalid && S[S_FLAG * LDP + tid] > (real)0);
            if (!alive) break;
            tlive = t + 1;
            load_dw<real>(a, base, nvalid, t, c.DW);
            if (tid < P) {
                real dt, sqdt, xn; int dtg;
                fwd_dt(E, c.X, (int)S[S_FLAG * LDP + tid], LDP, tid, dt, sqdt, xn, dtg);
                S[S_DT * LDP + tid] = dt; S[S_SQDT * LDP + tid] = sqdt; S[S_XN * LDP + tid] = xn;
            }
            eval_control<real>(a, E, c, cheat, false);
            if (td1) net_forward<real>(a.nG, a.pkG, c.X, c.Y0, c.hpp, c.GO, c.Ws);
            __syncthreads();
            if (tid < P) {
                const int p = tid;
                const real dt = S[S_DT * LDP + p], sqdt = S[S_SQDT * LDP + p], xn = S[S_XN * LDP + p];
                int flag = (int)S[S_FLAG * LDP + p];
                real* tr = traj + (size_t)t * 2 * sr * P;
                for (int k = 0; k < d; ++k) tr[k * P + p] = c.X[k * LDP + p];
                real w = (real)0;
                if (!prop_only) w = eq_w(E, c.X, c.U, LDP, p);
                const int coef = fwd_move(E, c.X, c.U, c.DW, dt, sqdt, xn, flag, c.SDW, LDP, p);
                const real cf = (real)coef;
                real disc = S[S_DISC * LDP + p], y = S[S_Y * LDP + p];
                y = y + w * disc * cf * dt;                                          // solver.py:170-174
                if (td1) {
                    real dif = (real)0;
                    for (int k = 0; k < d; ++k) dif = dif + c.SDW[k * LDP + p] * c.GO[k * LDP + p];   // solver.py:177-182
                    dif = dif * disc;
                    y = y - dif * cf * sqdt;                                         // solver.py:184
                    const real q = disc * cf * sqdt;
                    for (int k = 0; k < d; ++k) tr[(sr + k) * P + p] = c.SDW[k * LDP + p] * q;
                }
                disc = disc * dpb_exp(-E.gamma * dt * cf);                          // solver.py:187
                S[S_DISC * LDP + p] = disc; S[S_Y * LDP + p] = y; S[S_FLAG * LDP + p] = (real)flag;
                S[S_NACC * LDP + p] = S[S_NACC * LDP + p] + cf;
                if (p < nvalid) {
                    if (a.o_dt) a.o_dt[(base + p) * N + t] = dt;
                    if (a.o_coef) a.o_coef[(base + p) * N + t] = cf;
                    if (a.o_x)
                        for (int k = 0; k < d; ++k) a.o_x[((base + p) * d + k) * (long long)(N + 1) + t + 1] = c.X[k * LDP + p];
                }
            }
        }
        __syncthreads();
        if (tid < nvalid) {                                      // steps after every path of the tile is frozen
            const int p = tid;
            for (int t = tlive; t < N; ++t) {
                if (a.o_dt) a.o_dt[(base + p) * N + t] = E.delta_t;
                if (a.o_coef) a.o_coef[(base + p) * N + t] = (real)0;
                if (a.o_x)
                    for (int k = 0; k < d; ++k) a.o_x[((base + p) * d + k) * (long long)(N + 1) + t + 1] = c.X[k * LDP + p];
            }
            if (a.o_exit) a.o_exit[base + p] = (int)S[S_NACC * LDP + p];
        }
        if (prop_only) continue;

        // ------------------------------------------------------------------ NN_value at x_N, x_0
        // v_N with activations kept, v_0 forward-only first (delta needs both)
        net_forward<real>(a.nV, a.pkV, c.X, c.Y0, c.hb, c.OUTA, c.Ws);          // V(x_N), kept
        __syncthreads();
        if (tid < P) S[S_VN * LDP + tid] = c.OUTA[tid];
        __syncthreads();
        load_cols<real>(c.LAM, a.x0, base, nvalid, d, sr, fill);                 // LAM used as x_0 staging
        {
            real* hbv[MAXLIN];
            for (int i = 0; i < MAXLIN; ++i) hbv[i] = c.hb[a.nV.L + (i & 1)];    // two spare buffers
            net_forward<real>(a.nV, a.pkV, c.LAM, c.DY0, hbv, c.GO, c.Ws);       // V(x_0), forward only
        }
        __syncthreads();
        real rho_v = (real)0;
        if (tid < P) {
            const int p = tid;
            const real v0 = c.GO[p];
            const real delta = v0 - S[S_Y * LDP + p] - S[S_VN * LDP + p] * S[S_DISC * LDP + p];   // solver.py:189
            const bool valid = p < nvalid;
            S[S_DELTA * LDP + p] = delta;
            S[S_RHOG * LDP + p] = valid ? rho_grad(delta, (real)50) * scale : (real)0;
            if (valid) {
                rho_v = rho(delta, (real)50);
                if (a.o_delta) a.o_delta[base + p] = delta;
            }
        }
        loss0 = loss0 + block_sum(rho_v, c.red);
        if (need_grad) {
            zero_rows<real>(c.DOUT, sr);
            __syncthreads();
            if (tid < P) c.DOUT[tid] = -S[S_RHOG * LDP + tid] * S[S_DISC * LDP + tid];
            net_backward<real>(a.nV, a.pkV, c.X, c.Y0, c.hb, c.DOUT, c.hb[a.nV.L], c.hb[a.nV.L + 1], c.DY0, gsV, (real*)nullptr, c.Ws);
            net_forward<real>(a.nV, a.pkV, c.LAM, c.Y0, c.hb, c.OUTA, c.Ws);     // V(x_0), kept
            __syncthreads();
            if (tid < P) c.DOUT[tid] = S[S_RHOG * LDP + tid];
            net_backward<real>(a.nV, a.pkV, c.LAM, c.Y0, c.hb, c.DOUT, c.hb[a.nV.L], c.hb[a.nV.L + 1], c.DY0, gsV, (real*)nullptr, c.Ws);
        }
        // ------------------------------------------------------------------ boundary term
        __syncthreads();
        load_cols<real>(c.X, a.xb, base, nvalid, d, sr, fill);
        net_forward<real>(a.nV, a.pkV, c.X, c.Y0, c.hb, c.OUTA, c.Ws);
        __syncthreads();
        real rho_b = (real)0;
        if (tid < P) {
            const int p = tid;
            const real db = c.OUTA[p] - eq_Z(E, c.X, LDP, p);                      // solver.py:190
            const bool valid = p < nvalid;
            S[S_TMP * LDP + p] = valid ? rho_grad(db, (real)50) * scale : (real)0;
            if (valid) {
                rho_b = rho(db, (real)50);
                if (a.o_delta_b) a.o_delta_b[base + p] = db;
            }
        }
        loss1 = loss1 + block_sum(rho_b, c.red);
        if (need_grad) {
            zero_rows<real>(c.DOUT, sr);
            __syncthreads();
            if (tid < P) c.DOUT[tid] = S[S_TMP * LDP + tid];
            net_backward<real>(a.nV, a.pkV, c.X, c.Y0, c.hb, c.DOUT, c.hb[a.nV.L], c.hb[a.nV.L + 1], c.DY0, gsV, (real*)nullptr, c.Ws);
        }
        // ------------------------------------------------------------------ sweep 2: NN_value_grad backward
        if (need_grad && td1) {
            for (int t = 0; t < tlive; ++t) {
                const real* tr = traj + (size_t)t * 2 * sr * P;
                __syncthreads();
                for (int idx = tid; idx < sr * P; idx += NTHREADS) {
                    int k = idx / P, p = idx - k * P;
                    c.X[k * LDP + p] = (k < d) ? tr[k * P + p] : (real)0;
                    c.DOUT[k * LDP + p] = (k < d) ? tr[(sr + k) * P + p] * S[S_RHOG * LDP + p] : (real)0;
                }
                net_forward<real>(a.nG, a.pkG, c.X, c.Y0, c.hb, c.GO, c.Ws);
                net_backward<real>(a.nG, a.pkG, c.X, c.Y0, c.hb, c.DOUT, c.hb[a.nG.L], c.hb[a.nG.L + 1], c.DY0, gsG, (real*)nullptr, c.Ws);
            }
        }
    }
    if (tid == 0 && a.loss_part) {
        a.loss_part[blockIdx.x * 2] = loss0;
        a.loss_part[blockIdx.x * 2 + 1] = loss1;
    }
}

// =================================================================================== actor kernel
template <typename real>
__global__ void __launch_bounds__(NTHREADS, 1) actor_kernel(const StepArgs<real> a) {
    constexpr int P = 8 * RT<real>::TP, LDP = P + RT<real>::PADP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carve<real> c;
    carve<real>(c, reinterpret_cast<real*>(smem_raw), a.sr, a.hrows, a.nhb);
    {
        real* base = reinterpret_cast<real*>(smem_raw);
        const size_t n = carve_elems<real>(a.sr, a.hrows, a.nhb);
        for (size_t i = threadIdx.x; i < n; i += NTHREADS) base[i] = (real)0;
    }
    __syncthreads();
    const Eq<real> E(a.eq);
    const int d = E.d, m = E.m, N = a.N, sr = a.sr;
    const int tid = threadIdx.x;
    const bool cheat = a.flags & F_CHEAT_CONTROL, cheat_v = a.flags & F_CHEAT_VALUE;
    const bool need_grad = (a.flags & F_NEED_GRAD) && !cheat;
    const real fill = (real)0.5 * E.R / dpb_sqrt((real)d);
    real* traj = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;      // [N][2*sr + A_NSCAL][P]
    const int trs = 2 * sr + A_NSCAL;
    real* gsA = a.slabA ? a.slabA + (size_t)blockIdx.x * a.nA.gtotal : nullptr;
    real loss0 = (real)0;
    real* S = c.S;

    const long long ntiles = (a.B_local + P - 1) / P;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * P;
        const int nvalid = (int)((a.B_local - base < P) ? (a.B_local - base) : P);
        __syncthreads();
        load_cols<real>(c.X, a.x0, base, nvalid, d, sr, fill);
        __syncthreads();
        if (tid < P) {
            const int p = tid;
            S[S_FLAG * LDP + p] = (real)fwd_initial_flag(E, c.X, LDP, p);
            S[S_DISC * LDP + p] = (real)1;
            S[S_Y * LDP + p] = (real)0;
            S[S_NACC * LDP + p] = (real)0;
            if (a.o_x && p < nvalid)
                for (int k = 0; k < d; ++k) a.o_x[((base + p) * d + k) * (long long)(N + 1)] = c.X[k * LDP + p];
        }
        // ------------------------------------------------------------------ forward rollout
        int tlive = 0;
        for (int t = 0; t < N; ++t) {
            const int alive = __syncthreads_or(tid < P && tid < nvalid && S[S_FLAG * LDP + tid] > (real)0);
            if (!alive) break;
            tlive = t + 1;
            load_dw<real>(a, base, nvalid, t, c.DW);
            if (tid < P) {
                real dt, sqdt, xn; int dtg;
                fwd_dt(E, c.X, (int)S[S_FLAG * LDP + tid], LDP, tid, dt, sqdt, xn, dtg);
                S[S_DT * LDP + tid] = dt; S[S_SQDT * LDP + tid] = sqdt; S[S_XN * LDP + tid] = xn; S[S_DTG * LDP + tid] = (real)dtg;
            }
            eval_control<real>(a, E, c, cheat, false);
            if (tid < P) {
                const int p = tid;
                const real dt = S[S_DT * LDP + p], sqdt = S[S_SQDT * LDP + p], xn = S[S_XN * LDP + p];
                int flag = (int)S[S_FLAG * LDP + p];
                real disc = S[S_DISC * LDP + p], y = S[S_Y * LDP + p];
                real* tr = traj + (size_t)t * trs * P;
                for (int k = 0; k < d; ++k) { tr[k * P + p] = c.X[k * LDP + p]; tr[(sr + k) * P + p] = c.DW[k * LDP + p]; }
                const real w = eq_w(E, c.X, c.U, LDP, p);
                const int coef = fwd_move(E, c.X, c.U, c.DW, dt, sqdt, xn, flag, (real*)nullptr, LDP, p);
                const real cf = (real)coef;
                real* sc = tr + (size_t)2 * sr * P;
                sc[A_DT * P + p] = dt; sc[A_SQDT * P + p] = sqdt; sc[A_COEF * P + p] = cf; sc[A_DISC * P + p] = disc;
                sc[A_XN * P + p] = xn; sc[A_DTG * P + p] = S[S_DTG * LDP + p];
                y = y + cf * w * dt * disc;                                          // solver.py:218
                disc = disc * dpb_exp(-E.gamma * dt * cf);                          // solver.py:219
                S[S_DISC * LDP + p] = disc; S[S_Y * LDP + p] = y; S[S_FLAG * LDP + p] = (real)flag;
                S[S_NACC * LDP + p] = S[S_NACC * LDP + p] + cf;
                if (p < nvalid) {
                    if (a.o_dt) a.o_dt[(base + p) * N + t] = dt;
                    if (a.o_coef) a.o_coef[(base + p) * N + t] = cf;
                    if (a.o_x)
                        for (int k = 0; k < d; ++k) a.o_x[((base + p) * d + k) * (long long)(N + 1) + t + 1] = c.X[k * LDP + p];
                }
            }
        }
        __syncthreads();
        if (tid < nvalid) {
            const int p = tid;
            for (int t = tlive; t < N; ++t) {
                if (a.o_dt) a.o_dt[(base + p) * N + t] = E.delta_t;
                if (a.o_coef) a.o_coef[(base + p) * N + t] = (real)0;
                if (a.o_x)
                    for (int k = 0; k < d; ++k) a.o_x[((base + p) * d + k) * (long long)(N + 1) + t + 1] = c.X[k * LDP + p];
            }
            if (a.o_exit) a.o_exit[base + p] = (int)S[S_NACC * LDP + p];
        }
        // ------------------------------------------------------------------ terminal value
        if (cheat_v) {
            __syncthreads();
            if (tid < P) S[S_VN * LDP + tid] = eq_V_true(E, c.X, LDP, tid);             // solver.py:223
        } else {
            net_forward<real>(a.nV, a.pkV, c.X, c.Y0, c.hb, c.OUTA, c.Ws);              // solver.py:221
            __syncthreads();
            if (tid < P) S[S_VN * LDP + tid] = c.OUTA[tid];
        }
        __syncthreads();
        real yv = (real)0;
        if (tid < P) {
            const int p = tid;
            const real y = S[S_Y * LDP + p] + S[S_VN * LDP + p] * S[S_DISC * LDP + p];
            S[S_Y * LDP + p] = y;
            if (p < nvalid) {
                yv = y;
                if (a.o_delta) a.o_delta[base + p] = y;
            }
        }
        loss0 = loss0 + block_sum(yv, c.red);
        if (!need_grad) continue;
        // ------------------------------------------------------------------ reverse sweep
        // seed: lam = D_N/B * grad V(x_N) ; Dbar = V(x_N)/B      (zero for padded paths)
        if (cheat_v) {
            if (tid < P) {
                const int p = tid;
                eq_V_grad_true(E, c.X, c.LAM, LDP, p);
                const real s = (p < nvalid) ? S[S_DISC * LDP + p] * a.invB : (real)0;
                for (int k = 0; k < d; ++k) c.LAM[k * LDP + p] = c.LAM[k * LDP + p] * s;
            }
        } else {
            zero_rows<real>(c.DOUT, sr);
            __syncthreads();
            if (tid < P) c.DOUT[tid] = (tid < nvalid) ? S[S_DISC * LDP + tid] * a.invB : (real)0;
            net_backward<real>(a.nV, a.pkV, c.X, c.Y0, c.hb, c.DOUT, c.hb[a.nV.L], c.hb[a.nV.L + 1], c.DY0, (real*)nullptr, c.LAM, c.Ws);
        }
        if (tid < P) S[S_DBAR * LDP + tid] = (tid < nvalid) ? S[S_VN * LDP + tid] * a.invB : (real)0;
        __syncthreads();
        for (int t = tlive - 1; t >= 0; --t) {
            const real* tr = traj + (size_t)t * trs * P;
            const real* sc = tr + (size_t)2 * sr * P;
            const int any = __syncthreads_or(tid < nvalid && sc[A_COEF * P + (tid < P ? tid : 0)] > (real)0);
            if (!any) continue;
            for (int idx = tid; idx < sr * P; idx += NTHREADS) {
                int k = idx / P, p = idx - k * P;
                c.X[k * LDP + p] = (k < d) ? tr[k * P + p] : (real)0;
                c.DW[k * LDP + p] = (k < d) ? tr[(sr + k) * P + p] : (real)0;
            }
            eval_control<real>(a, E, c, false, true);                               // activations kept in c.hb[0..L-1]
            zero_rows<real>(c.DOUT, sr);
            __syncthreads();
            if (tid < P) {
                const int p = tid;
                const int coef = (p < nvalid && sc[A_COEF * P + p] > (real)0) ? 1 : 0;
                real Dbar = S[S_DBAR * LDP + p];
                adj_step(E, c.X, c.U, c.DW, sc[A_DT * P + p], sc[A_SQDT * P + p], coef, (int)sc[A_DTG * P + p], sc[A_XN * P + p],
                         sc[A_DISC * P + p], a.invB, c.LAM, Dbar, c.SDW, LDP, p);   // SDW <- ubar
                S[S_DBAR * LDP + p] = Dbar;
                if (a.nA.ekn_head) ekn_head_bwd(c.OUTA, c.SDW, c.DOUT, m, LDP, p);
                else for (int j = 0; j < m; ++j) c.DOUT[j * LDP + p] = c.SDW[j * LDP + p];
            }
            net_backward<real>(a.nA, a.pkA, c.X, c.Y0, c.hb, c.DOUT, c.hb[a.nA.L], c.hb[a.nA.L + 1], c.DY0, gsA, c.GO, c.Ws);   // GO <- dx
            for (int idx = tid; idx < d * P; idx += NTHREADS) {
                int k = idx / P, p = idx - k * P;
                c.LAM[k * LDP + p] = c.LAM[k * LDP + p] + c.GO[k * LDP + p];
            }
            __syncthreads();
        }
    }
    if (tid == 0 && a.loss_part) {
        a.loss_part[blockIdx.x * 2] = loss0;
        a.loss_part[blockIdx.x * 2 + 1] = (real)0;
    }
}

// ===================================================================================== aux kernels
// out_loss[i] = scale_i * sum over CTAs of loss_part[cta][i]   (fixed order)
template <typename real>
__global__ void reduce_loss_kernel(const real* __restrict__ part, int nparts, real s0, real s1, real* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        real a0 = (real)0, a1 = (real)0;
        for (int i = 0; i < nparts; ++i) { a0 = a0 + part[2 * i]; a1 = a1 + part[2 * i + 1]; }
        out[0] = a0 * s0;
        out[1] = a1 * s1;
    }
}

// DeepNN.call on n points (solver.py:260-278): x[n][in] -> out[n][out_dim]
template <typename real>
__global__ void __launch_bounds__(NTHREADS, 1) mlp_forward_kernel(NetDev nd, const real* __restrict__ pk, const real* __restrict__ x, long long n,
                                                                  real* __restrict__ out, int sr, int hrows) {
    constexpr int P = 8 * RT<real>::TP, LDP = P + RT<real>::PADP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carve<real> c;
    carve<real>(c, reinterpret_cast<real*>(smem_raw), sr, hrows, 2);
    {
        real* base = reinterpret_cast<real*>(smem_raw);
        const size_t ne = carve_elems<real>(sr, hrows, 2);
        for (size_t i = threadIdx.x; i < ne; i += NTHREADS) base[i] = (real)0;
    }
    __syncthreads();
    const long long ntiles = (n + P - 1) / P;
    const int od = nd.ekn_head ? nd.mctrl : nd.out;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * P;
        const int nvalid = (int)((n - base < P) ? (n - base) : P);
        __syncthreads();
        load_cols<real>(c.X, x, base, nvalid, nd.in, sr, (real)0.25);
        net_forward<real>(nd, pk, c.X, c.Y0, c.hpp, c.OUTA, c.Ws);
        __syncthreads();
        real* res = c.OUTA;
        if (nd.ekn_head) {
            if (threadIdx.x < P) ekn_head_fwd(c.OUTA, c.U, nd.mctrl, LDP, threadIdx.x);
            __syncthreads();
            res = c.U;
        }
        for (int idx = threadIdx.x; idx < od * P; idx += NTHREADS) {
            int p = idx / od, k = idx - p * od;
            if (p < nvalid) out[(base + p) * od + k] = res[k * LDP + p];
        }
    }
}

// closed forms on n points (equation.py:157-167,201-227,252-265,292-302).
enum { CF_V_TRUE = 0, CF_U_TRUE = 1, CF_V_GRAD_TRUE = 2, CF_Z = 3, CF_W = 4, CF_DRIFT = 5, CF_SIGMA = 6 };
template <typename real>
__global__ void closed_form_kernel(EqnD eq, int which, const real* __restrict__ x, const real* __restrict__ u, long long n, real* __restrict__ out) {
    const Eq<real> E(eq);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const real* xi = x + i * E.d;
    real tmp[32];
    switch (which) {
    case CF_V_TRUE: out[i] = eq_V_true(E, xi, 1, 0); break;
    case CF_Z: out[i] = eq_Z(E, xi, 1, 0); break;
    case CF_W: out[i] = eq_w(E, xi, u + i * E.m, 1, 0); break;
    case CF_U_TRUE:
        eq_u_true(E, xi, tmp, 1, 0);
        for (int k = 0; k < E.m; ++k) out[i * E.m + k] = tmp[k];
        break;
    case CF_DRIFT: {                                     // Equation.drift (equation.py:172,232-235,270-273,307)
        const real cc = (E.eqn == EQ_EKN) ? eq_drift_c(E, dpb_sqrt(norm2_path(xi, E.d, 1, 0))) : (real)0;
        for (int k = 0; k < E.d; ++k) out[i * E.d + k] = eq_drift(E, cc, xi, u + i * E.m, k, 1, 0);
        break;
    }
    case CF_SIGMA:                                       // Equation.sigma (equation.py:170,230,268,305): [n][d][d], diagonal
        for (int k = 0; k < E.d; ++k)
            for (int j = 0; j < E.d; ++j) out[(i * E.d + k) * E.d + j] = (j == k) ? eq_sigma(E, xi, u ? u + i * E.m : xi, k, 1, 0) : (real)0;
        break;
    default:
        eq_V_grad_true(E, xi, tmp, 1, 0);
        for (int k = 0; k < E.d; ++k) out[i * E.d + k] = tmp[k];
    }
}

// Equation.diffusion (equation.py:175-176,237-238,275-276,310-311): sigma(x,u) . dw on n points (sigma is diagonal)
template <typename real>
__global__ void diffusion_kernel(EqnD eq, const real* __restrict__ x, const real* __restrict__ u, const real* __restrict__ dw, long long n,
                                 real* __restrict__ out) {
    const Eq<real> E(eq);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const real* xi = x + i * E.d;
    for (int k = 0; k < E.d; ++k) out[i * E.d + k] = eq_sigma(E, xi, u ? u + i * E.m : xi, k, 1, 0) * dw[i * E.d + k];
}

// error metrics of solver.py:109-130 on n values: out[0] = sum (t-a)^2, out[1] = sum t^2, out[2] = max |t-a|
// (one CTA, fixed order => deterministic; validation sets are a few thousand points)
template <typename real>
__global__ void err_metrics_kernel(const real* __restrict__ t, const real* __restrict__ a, long long n, real* __restrict__ out) {
    __shared__ real sh[3][256];
    real s0 = (real)0, s1 = (real)0, mx = (real)0;
    for (long long i = threadIdx.x; i < n; i += 256) {
        const real e = t[i] - a[i];
        s0 = s0 + e * e;
        s1 = s1 + t[i] * t[i];
        const real ae = e < (real)0 ? -e : e;
        mx = ae > mx ? ae : mx;
    }
    sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1; sh[2][threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sh[0][threadIdx.x] = sh[0][threadIdx.x] + sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] = sh[1][threadIdx.x] + sh[1][threadIdx.x + o];
            const real b = sh[2][threadIdx.x + o];
            if (b > sh[2][threadIdx.x]) sh[2][threadIdx.x] = b;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sh[0][0]; out[1] = sh[1][0]; out[2] = sh[2][0]; }
}

// tf.keras Adam (solver.py:16-21): m,v EMA then theta -= lr_t * m / (sqrt(v) + eps)
template <typename real>
__global__ void adam_kernel(real* __restrict__ th, const real* __restrict__ g, real* __restrict__ m, real* __restrict__ v, long long n,
                            real lr_t, const double* __restrict__ lr_t_dev, real b1, real b2, real eps) {
    if (lr_t_dev) lr_t = (real)*lr_t_dev;                  // (CUDA-graph replays: the host writes this step's rate into device memory)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const real gi = g[i];
        const real mi = m[i] + (gi - m[i]) * ((real)1 - b1);
        const real vi = v[i] + (gi * gi - v[i]) * ((real)1 - b2);
        m[i] = mi; v[i] = vi;
        th[i] = th[i] - lr_t * mi / (dpb_sqrt(vi) + eps);
    }
}

// Lifetime sort of the naive scheme (counting sort of the paths by their number of accepted steps, longest first):
// hist[N - nacc]++  ->  exclusive scan (one block)  ->  perm[offset[bin]++] = path
static __global__ void lifetime_hist_kernel(const int* __restrict__ nacc, long long B, int N, int* __restrict__ hist) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += stride) {
        int k = nacc[i]; k = k < 0 ? 0 : (k > N ? N : k);
        atomicAdd(&hist[N - k], 1);
    }
}
static __global__ void lifetime_scan_kernel(int* __restrict__ hist, int nbins) {       // in place: counts -> start offsets
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < nbins; ++b) { const int c = hist[b]; hist[b] = run; run += c; }
    }
}
static __global__ void lifetime_scatter_kernel(const int* __restrict__ nacc, long long B, int N, int* __restrict__ offs, int* __restrict__ perm) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += stride) {
        int k = nacc[i]; k = k < 0 ? 0 : (k > N ? N : k);
        perm[atomicAdd(&offs[N - k], 1)] = (int)i;
    }
}

// the increments the PHILOX modes generate, materialised as dw[B][d][N]
template <typename real>
__global__ void philox_dw_kernel(int mode, unsigned long long seed, unsigned long long stream, long long path_offset, long long B, int d, int N,
                                 real* __restrict__ dw) {
    const int nch = (d + 3) >> 2;
    const long long total = B * N * nch;
    uint32_t k0, k1;
    philox_key(seed, stream, k0, k1);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int t = (int)(i % N);
        const long long r = i / N;
        const int ch = (int)(r % nch);
        const long long p = r / nch;
        const unsigned long long gp = (unsigned long long)(path_offset + p);
        uint32_t c[4] = {(uint32_t)gp, (uint32_t)(gp >> 32), (uint32_t)t, (uint32_t)ch};
        philox4x32_10(c, k0, k1);
        float o[4];
        if (mode == DW_PHILOX_BOUNDED) {
            for (int j = 0; j < 4; ++j) o[j] = philox_bounded(c[j]);
        } else {
            for (int j = 0; j < 2; ++j) {
                float rr = sqrtf(-2.0f * logf(philox_u01(c[2 * j])));
                float s, co;
                sincospif(2.0f * philox_u01(c[2 * j + 1]), &s, &co);
                o[2 * j] = rr * co;
                o[2 * j + 1] = rr * s;
            }
        }
        for (int j = 0; j < 4; ++j)
            if (4 * ch + j < d) dw[(p * d + 4 * ch + j) * (long long)N + t] = (real)o[j];
    }
}

// x0 uniform in the ball of radius R, x_bdry uniform on the sphere (equation.py:14-22), from Philox
// streams keyed by the GLOBAL path index (counter word 2: 0x80000000 | chunk, word 3: 1 = x0, 2 = x_bdry).
template <typename real>
__global__ void sample_x_kernel(unsigned long long seed, unsigned long long stream, const unsigned long long* __restrict__ stream_base,
                                long long path_offset, long long B, int d, real R, real* __restrict__ x0, real* __restrict__ xb) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    if (stream_base) stream += *stream_base;
    uint32_t k0, k1;
    philox_key(seed, stream, k0, k1);
    const unsigned long long gp = (unsigned long long)(path_offset + p);
    const int nch = (d + 3) >> 2;
    for (int which = 1; which <= 2; ++which) {
        real* dst = (which == 1) ? x0 : xb;
        if (!dst) continue;
        float g[32];
        float n2 = 0.f;
        for (int ch = 0; ch < nch; ++ch) {
            uint32_t c[4] = {(uint32_t)gp, (uint32_t)(gp >> 32), 0x80000000u | (uint32_t)ch, (uint32_t)which};
            philox4x32_10(c, k0, k1);
            for (int j = 0; j < 2; ++j) {
                float rr = sqrtf(-2.0f * logf(philox_u01(c[2 * j])));
                float s, co;
                sincospif(2.0f * philox_u01(c[2 * j + 1]), &s, &co);
                g[4 * ch + 2 * j] = rr * co;
                g[4 * ch + 2 * j + 1] = rr * s;
            }
        }
        for (int k = 0; k < d; ++k) n2 += g[k] * g[k];
        float rad = (float)R;
        if (which == 1) {
            uint32_t c[4] = {(uint32_t)gp, (uint32_t)(gp >> 32), 0xC0000000u, 1u};
            philox4x32_10(c, k0, k1);
            rad = (float)R * powf(philox_u01(c[0]), 1.0f / (float)d);      // r = R * U^(1/d)
        }
        const float s = rad / sqrtf(n2);
        for (int k = 0; k < d; ++k) dst[p * d + k] = (real)(g[k] * s);
    }
}

}  // namespace dpb
