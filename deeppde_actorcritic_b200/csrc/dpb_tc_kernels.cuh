// dpb_tc_kernels.cuh -- the fused rollout + TD kernels with the MLP layers on tcgen05 (impl = tensor).
// Same algorithm and per-path arithmetic (dpb_eqn.h) as dpb_kernels.cuh; thread t of warps 0-3 owns
// path t of the 128-path tile entirely in registers, warp 4 lane 0 drives the tensor pipe.
#pragma once
#include "dpb_kernels.cuh"
#include "dpb_tc_nets.cuh"

namespace dpb {
namespace tc {

struct TcArgs {
    EqnD eq;
    TcNet nA, nV, nG;
    const unsigned char *imgA, *imgV, *imgG;
    const float *vecA, *vecV, *vecG;
    const float *x0, *dw, *xb;
    int dw_mode;
    unsigned long long seed, stream;
    long long B_local, path_offset;
    float invB;
    int N;
    unsigned flags;
    float* loss_part;               // [grid][2]
    float *slabV, *slabG, *slabA;   // per-CTA raw-gradient slabs
    float* scratch;
    long long scratch_per_cta;
    int sr;
    float *o_x, *o_dt, *o_coef, *o_delta, *o_delta_b;
    int* o_exit;
};

struct TcSmem {
    unsigned char* ring;
    float *vecA, *vecV, *vecG;
    uint64_t *full, *empty, *acc_full, *a_ready;
    uint32_t* tslot;
    Sched* sch;
    float* red;
};

__host__ __device__ inline size_t tc_smem_bytes(int vfA, int vfV, int vfG) {
    return (size_t)NSLOT * SLOT_BYTES + (size_t)(vfA + vfV + vfG) * 4 + (2 * NSLOT + 2) * 8 + 64 + sizeof(Sched) + 64 + 1024;
}

__device__ __forceinline__ void tc_carve(TcSmem& s, unsigned char* base, int vfA, int vfV, int vfG) {
    unsigned char* p = reinterpret_cast<unsigned char*>(((uintptr_t)base + 1023) & ~(uintptr_t)1023);
    s.ring = p; p += (size_t)NSLOT * SLOT_BYTES;
    s.vecA = reinterpret_cast<float*>(p); p += (size_t)vfA * 4;
    s.vecV = reinterpret_cast<float*>(p); p += (size_t)vfV * 4;
    s.vecG = reinterpret_cast<float*>(p); p += (size_t)vfG * 4;
    s.full = reinterpret_cast<uint64_t*>(p); p += NSLOT * 8;
    s.empty = reinterpret_cast<uint64_t*>(p); p += NSLOT * 8;
    s.acc_full = reinterpret_cast<uint64_t*>(p); p += 8;
    s.a_ready = reinterpret_cast<uint64_t*>(p); p += 8;
    s.tslot = reinterpret_cast<uint32_t*>(p); p += 64;
    s.sch = reinterpret_cast<Sched*>(p); p += sizeof(Sched);
    s.red = reinterpret_cast<float*>(((uintptr_t)p + 15) & ~(uintptr_t)15);
}

// common prologue: barriers, TMEM, vector blocks -> shared memory.  Returns the TMEM base.
__device__ __forceinline__ uint32_t tc_setup(TcSmem& s, const TcArgs& a) {
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < a.nA.vec_floats; i += TC_THREADS) s.vecA[i] = a.vecA ? a.vecA[i] : 0.f;
    for (int i = tid; i < a.nV.vec_floats; i += TC_THREADS) s.vecV[i] = a.vecV ? a.vecV[i] : 0.f;
    for (int i = tid; i < a.nG.vec_floats; i += TC_THREADS) s.vecG[i] = a.vecG ? a.vecG[i] : 0.f;
    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
        mbar_init(s.acc_full, 1);
        mbar_init(s.a_ready, TC_PATHS);
        s.sch->nops = 0;
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(s.tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *s.tslot;
}

__device__ __forceinline__ float tc_block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0)
        for (int i = 0; i < TC_THREADS / 32; ++i) s += red[i];
    return s;
}

// increments of step t for one path (same generator and bits as load_dw of the exact path)
__device__ __forceinline__ void path_dw(const TcArgs& a, long long gpath_local, bool valid, int t, float* dw) {
    const int d = a.eq.d;
    if (a.dw_mode == DW_EXTERNAL) {
        for (int k = 0; k < d; ++k) dw[k] = valid ? a.dw[(gpath_local * d + k) * (long long)a.N + t] : 0.f;
        return;
    }
    uint32_t k0, k1;
    philox_key(a.seed, a.stream, k0, k1);
    const unsigned long long gp = (unsigned long long)(a.path_offset + gpath_local);
    const int nch = (d + 3) >> 2;
    for (int ch = 0; ch < nch; ++ch) {
        uint32_t c[4] = {(uint32_t)gp, (uint32_t)(gp >> 32), (uint32_t)t, (uint32_t)ch};
        philox4x32_10(c, k0, k1);
        float o[4];
        if (a.dw_mode == DW_PHILOX_BOUNDED) {
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = philox_bounded(c[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float r = sqrtf(-2.0f * logf(philox_u01(c[2 * i])));
                float s, co;
                sincospif(2.0f * philox_u01(c[2 * i + 1]), &s, &co);
                o[2 * i] = r * co;
                o[2 * i + 1] = r * s;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (4 * ch + i < d) dw[4 * ch + i] = o[i];
    }
}

// ================================================================================== critic (tensor)
__global__ void __launch_bounds__(TC_THREADS, 1) critic_tc_kernel(const TcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem S;
    tc_carve(S, smem_raw, a.nA.vec_floats, a.nV.vec_floats, a.nG.vec_floats);
    const uint32_t tmem = tc_setup(S, a);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_path = warp < 4;
    const bool is_ctrl = (warp == 4 && lane == 0);
    const Eq<float> E(a.eq);
    const int d = E.d, N = a.N;
    const bool cheat = a.flags & F_CHEAT_CONTROL, prop_only = a.flags & F_PROPAGATE_ONLY;
    const bool td1 = (E.td == 1) && !prop_only;
    const float scale = 100.f * a.invB;
    const float fill = 0.5f * E.R / sqrtf((float)d);

    Ctrl C;
    C.ring = S.ring; C.full = S.full; C.empty = S.empty; C.acc_full = S.acc_full; C.a_ready = S.a_ready; C.sch = S.sch;
    C.pf_op = 0; C.pf_ch = 0; C.n_loaded = 0; C.n_consumed = 0; C.op_count = 0; C.tmem = tmem;
    PathCtx P;
    P.tl = tmem + ((uint32_t)(warp * 32) << 16);
    P.acc_full = S.acc_full; P.a_ready = S.a_ready; P.op_count = 0;

    float loss0 = 0.f, loss1 = 0.f;
    const long long ntiles = (a.B_local + TC_PATHS - 1) / TC_PATHS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * TC_PATHS;
        const long long gp = base + tid;                          // local path index of this thread
        const bool valid = is_path && gp < a.B_local;
        float x[32], u[32], dwv[32], sdw[32], g[32], raw[32];
        int flag = 0, nacc = 0;
        float disc = 1.f, y = 0.f;
        if (is_path) {
            for (int k = 0; k < d; ++k) x[k] = valid ? a.x0[gp * d + k] : fill;
            flag = fwd_initial_flag(E, x, 1, 0);
            if (a.o_x && valid)
                for (int k = 0; k < d; ++k) a.o_x[(gp * d + k) * (long long)(N + 1)] = x[k];
        }
        if (is_ctrl) {                                            // schedule of the rollout
            ctrl_flush(C);
            if (!cheat) sched_add_fwd(C.sch, a.nA, a.imgA, a.nA.L);
            if (td1) sched_add_fwd(C.sch, a.nG, a.imgG, a.nG.L);
        }
        // ------------------------------------------------------------------ sweep 1: rollout
        int tlive = 0;
        for (int t = 0; t < N; ++t) {
            const int alive = __syncthreads_or(valid && flag > 0);
            if (!alive) break;
            tlive = t + 1;
            if (is_ctrl) {
                if (!cheat) ctrl_net_forward(C, a.nA, a.nA.L);
                if (td1) ctrl_net_forward(C, a.nG, a.nG.L);
            } else if (is_path) {
                path_dw(a, gp, valid, t, dwv);
                float dt, sqdt, xn; int dtg;
                fwd_dt(E, x, flag, 1, 0, dt, sqdt, xn, dtg);
                if (cheat) {
                    eq_u_true(E, x, u, 1, 0);
                } else {
                    path_net_forward(P, a.nA, S.vecA, x, raw);
                    if (a.nA.ekn_head) ekn_head_fwd(raw, u, a.nA.mctrl, 1, 0);
                    else for (int j = 0; j < E.m; ++j) u[j] = raw[j];
                }
                if (td1) path_net_forward(P, a.nG, S.vecG, x, g);
                float w = 0.f;
                if (!prop_only) w = eq_w(E, x, u, 1, 0);
                const int coef = fwd_move(E, x, u, dwv, dt, sqdt, xn, flag, sdw, 1, 0);
                const float cf = (float)coef;
                y = y + w * disc * cf * dt;                                       // solver.py:170-174
                if (td1) {
                    float dif = 0.f;
                    for (int k = 0; k < d; ++k) dif = dif + sdw[k] * g[k];        // solver.py:177-182
                    dif = dif * disc;
                    y = y - dif * cf * sqdt;                                      // solver.py:184
                }
                disc = disc * expf(-E.gamma * dt * cf);                          // solver.py:187
                nacc += coef;
                if (valid) {
                    if (a.o_dt) a.o_dt[gp * N + t] = dt;
                    if (a.o_coef) a.o_coef[gp * N + t] = cf;
                    if (a.o_x)
                        for (int k = 0; k < d; ++k) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
                }
            }
        }
        if (valid) {
            for (int t = tlive; t < N; ++t) {
                if (a.o_dt) a.o_dt[gp * N + t] = E.delta_t;
                if (a.o_coef) a.o_coef[gp * N + t] = 0.f;
                if (a.o_x)
                    for (int k = 0; k < d; ++k) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
            }
            if (a.o_exit) a.o_exit[gp] = nacc;
        }
        if (prop_only) continue;
        // ------------------------------------------------------------------ NN_value at x_N, x_0, x_bdry
        float rho_v = 0.f, rho_b = 0.f;
        if (is_ctrl) {
            ctrl_flush(C);
            sched_add_fwd(C.sch, a.nV, a.imgV, a.nV.L);
            for (int i = 0; i < 3; ++i) ctrl_net_forward(C, a.nV, a.nV.L);
        } else if (is_path) {
            float vN[1], v0[1], vb[1], x0v[32], xbv[32];
            path_net_forward(P, a.nV, S.vecV, x, vN);
            for (int k = 0; k < d; ++k) x0v[k] = valid ? a.x0[gp * d + k] : fill;
            path_net_forward(P, a.nV, S.vecV, x0v, v0);
            for (int k = 0; k < d; ++k) xbv[k] = valid ? a.xb[gp * d + k] : fill;
            path_net_forward(P, a.nV, S.vecV, xbv, vb);
            const float delta = v0[0] - y - vN[0] * disc;                         // solver.py:189
            const float db = vb[0] - eq_Z(E, xbv, 1, 0);                          // solver.py:190
            if (valid) {
                rho_v = rho(delta, 50.f);
                rho_b = rho(db, 50.f);
                if (a.o_delta) a.o_delta[gp] = delta;
                if (a.o_delta_b) a.o_delta_b[gp] = db;
            }
        }
        loss0 += tc_block_sum(rho_v, S.red);
        loss1 += tc_block_sum(rho_b, S.red);
        (void)scale;
    }
    if (is_ctrl) ctrl_flush(C);
    if (tid == 0 && a.loss_part) {
        a.loss_part[blockIdx.x * 2] = loss0;
        a.loss_part[blockIdx.x * 2 + 1] = loss1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 512);
}

// =================================================================================== actor (tensor)
__global__ void __launch_bounds__(TC_THREADS, 1) actor_tc_kernel(const TcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem S;
    tc_carve(S, smem_raw, a.nA.vec_floats, a.nV.vec_floats, a.nG.vec_floats);
    const uint32_t tmem = tc_setup(S, a);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_path = warp < 4;
    const bool is_ctrl = (warp == 4 && lane == 0);
    const Eq<float> E(a.eq);
    const int d = E.d, N = a.N;
    const bool cheat = a.flags & F_CHEAT_CONTROL, cheat_v = a.flags & F_CHEAT_VALUE;
    const float fill = 0.5f * E.R / sqrtf((float)d);

    Ctrl C;
    C.ring = S.ring; C.full = S.full; C.empty = S.empty; C.acc_full = S.acc_full; C.a_ready = S.a_ready; C.sch = S.sch;
    C.pf_op = 0; C.pf_ch = 0; C.n_loaded = 0; C.n_consumed = 0; C.op_count = 0; C.tmem = tmem;
    PathCtx P;
    P.tl = tmem + ((uint32_t)(warp * 32) << 16);
    P.acc_full = S.acc_full; P.a_ready = S.a_ready; P.op_count = 0;

    float loss0 = 0.f;
    const long long ntiles = (a.B_local + TC_PATHS - 1) / TC_PATHS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * TC_PATHS;
        const long long gp = base + tid;
        const bool valid = is_path && gp < a.B_local;
        float x[32], u[32], dwv[32], raw[32];
        int flag = 0, nacc = 0;
        float disc = 1.f, y = 0.f;
        if (is_path) {
            for (int k = 0; k < d; ++k) x[k] = valid ? a.x0[gp * d + k] : fill;
            flag = fwd_initial_flag(E, x, 1, 0);
            if (a.o_x && valid)
                for (int k = 0; k < d; ++k) a.o_x[(gp * d + k) * (long long)(N + 1)] = x[k];
        }
        if (is_ctrl) {
            ctrl_flush(C);
            if (!cheat) sched_add_fwd(C.sch, a.nA, a.imgA, a.nA.L);
        }
        int tlive = 0;
        for (int t = 0; t < N; ++t) {
            const int alive = __syncthreads_or(valid && flag > 0);
            if (!alive) break;
            tlive = t + 1;
            if (is_ctrl) {
                if (!cheat) ctrl_net_forward(C, a.nA, a.nA.L);
            } else if (is_path) {
                path_dw(a, gp, valid, t, dwv);
                float dt, sqdt, xn; int dtg;
                fwd_dt(E, x, flag, 1, 0, dt, sqdt, xn, dtg);
                if (cheat) {
                    eq_u_true(E, x, u, 1, 0);
                } else {
                    path_net_forward(P, a.nA, S.vecA, x, raw);
                    if (a.nA.ekn_head) ekn_head_fwd(raw, u, a.nA.mctrl, 1, 0);
                    else for (int j = 0; j < E.m; ++j) u[j] = raw[j];
                }
                const float w = eq_w(E, x, u, 1, 0);
                const int coef = fwd_move(E, x, u, dwv, dt, sqdt, xn, flag, (float*)nullptr, 1, 0);
                const float cf = (float)coef;
                y = y + cf * w * dt * disc;                                       // solver.py:218
                disc = disc * expf(-E.gamma * dt * cf);                          // solver.py:219
                nacc += coef;
                if (valid) {
                    if (a.o_dt) a.o_dt[gp * N + t] = dt;
                    if (a.o_coef) a.o_coef[gp * N + t] = cf;
                    if (a.o_x)
                        for (int k = 0; k < d; ++k) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
                }
            }
        }
        if (valid) {
            for (int t = tlive; t < N; ++t) {
                if (a.o_dt) a.o_dt[gp * N + t] = E.delta_t;
                if (a.o_coef) a.o_coef[gp * N + t] = 0.f;
                if (a.o_x)
                    for (int k = 0; k < d; ++k) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
            }
            if (a.o_exit) a.o_exit[gp] = nacc;
        }
        float yv = 0.f;
        if (is_ctrl) {
            ctrl_flush(C);
            if (!cheat_v) {
                sched_add_fwd(C.sch, a.nV, a.imgV, a.nV.L);
                ctrl_net_forward(C, a.nV, a.nV.L);
            }
        } else if (is_path) {
            float vN[1];
            if (cheat_v) vN[0] = eq_V_true(E, x, 1, 0);                           // solver.py:223
            else path_net_forward(P, a.nV, S.vecV, x, vN);                        // solver.py:221
            y = y + vN[0] * disc;
            if (valid) {
                yv = y;
                if (a.o_delta) a.o_delta[gp] = y;
            }
        }
        loss0 += tc_block_sum(yv, S.red);
    }
    if (is_ctrl) ctrl_flush(C);
    if (tid == 0 && a.loss_part) {
        a.loss_part[blockIdx.x * 2] = loss0;
        a.loss_part[blockIdx.x * 2 + 1] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 512);
}

}  // namespace tc
}  // namespace dpb
