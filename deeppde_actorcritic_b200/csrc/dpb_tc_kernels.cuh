// dpb_tc_kernels.cuh -- the fused rollout + TD kernels with the MLP layers on tcgen05 (impl = tensor).
// Same algorithm and per-path arithmetic (dpb_eqn.h) as dpb_kernels.cuh.  CTA = one tile of 128 paths = the 128
// TMEM lanes; 4 * TC_NGRP path warps (TC_NGRP = 2, critic kernels: threads t and t+128 own path t, state in registers,
// epilogue chunks split between them; TC_NGRP = 1, actor kernels: one thread per path with 255 registers), then the
// control warp (all lanes run the protocol, the elected lane issues every tcgen05.mma) and the producer warp (lane 0
// streams the weights).  The group count is a per-translation-unit constant (DPB_TC_NGRP, dpb_tc_nets.cuh).
// Phases per tile -- critic: rollout (actor + NN_value_grad forward) -> NN_value at x_N, x_0, x_bdry (+ backward)
// -> second sweep re-evaluating NN_value_grad at the stored x_t and back-propagating; actor: rollout -> terminal
// value (+ input gradient) -> reverse sweep (re-evaluate the actor, adjoint step, back-propagate).
#pragma once
#include "dpb_kernels.cuh"
#include "dpb_tc_nets.cuh"

namespace dpb {
namespace tc {

struct TcArgs {
    Eq<float> eqf;                  // equation + scheme constants in the arithmetic type of the tensor path
    TcNet nA, nV, nG;
    const unsigned char *imgA, *imgV, *imgG;
    const float *vecA, *vecV, *vecG;
    const float *x0, *dw, *xb;
    int dw_mode;
    unsigned long long seed, stream;
    const unsigned long long* stream_base;   // optional device word added to `stream` (CUDA-graph replays)
    long long B_local, path_offset;
    float invB;
    int N;
    unsigned flags;
    float* loss_part;               // [grid][2]
    float *slabV, *slabG, *slabA;   // per-CTA raw-gradient slabs (layout TcSlab)
    TcSlab gA, gV, gG;
    float* scratch;                 // per-CTA trajectory scratch (floats)
    long long scratch_per_cta;
    unsigned char* copies;          // per-CTA activation-copy scratch (bytes)
    long long copies_per_cta;
    int sr;
    int nslot, slot_bytes, actdz_bytes;   // ring geometry; bytes of each of the ACT / DZ images (0: no gradients)
    int nslab;                      // gradient slabs shared by the CTAs (CTA b reduces into slab b % nslab)
    float *o_x, *o_dt, *o_coef, *o_delta, *o_delta_b;
    int* o_exit;
    long long* stats;               // [grid][16] cycle counters (diagnostics), may be NULL
    int* tile_counter;              // zeroed before the launch: CTAs take tile blockIdx.x first, then gridDim.x + counter++
    const int* perm;                // optional: slot i of the tiling works on local path perm[i] (naive scheme: paths sorted by
                                    // lifetime so that the tiles die as a whole; NULL: identity)
};

// the kernels are instantiated in their own translation units (dpb_tc_inst_*.cu)
typedef void (*TcKernelFn)(const TcArgs);
#define DPB_TC_FOR_INSTANCES(X) X(lqr) X(ekn) X(lqrvar) X(vdp2) X(vdp5) X(vdp10) X(generic)
#define DPB_TC_DECL_GETTERS(n) TcKernelFn tc_get_critic_##n(); TcKernelFn tc_get_actor_##n();
DPB_TC_FOR_INSTANCES(DPB_TC_DECL_GETTERS)

struct TcSmem {
    unsigned char *act, *dz;
    unsigned char* ring;
    float *vecA, *vecV, *vecG;
    uint64_t *full, *empty, *acc_full, *a_all, *a_chunk, *act_full;
    uint32_t* tslot;
    int* tile;                       // the tile a CTA works on next (dynamic tile scheduler)
    Sched* sch;
    ProdCtl* pc;
    float* red;
    uint32_t* dzmax;                 // [2][8]: exchange of the tile's largest |cotangent| among the path warps (path_put_dz)
    TcNet *nA, *nV, *nG;             // shared-memory copies of the network descriptors
    TcSlab *gA, *gV, *gG;
};

// everything except the ring
__host__ __device__ inline size_t tc_smem_fixed(int vfA, int vfV, int vfG, int actdz_bytes) {
    return 2 * (size_t)actdz_bytes + (size_t)(vfA + vfV + vfG) * 4 + (2 * MAX_NSLOT + 3 + MAX_CHUNK) * 8 + 64 + sizeof(Sched) + 64 + 64 + 64 + 3 * sizeof(TcNet) + 3 * sizeof(TcSlab) + 64 + 1024;
}
__host__ __device__ inline size_t tc_smem_bytes(int vfA, int vfV, int vfG, int actdz_bytes, int nslot, int slot_bytes) {
    return tc_smem_fixed(vfA, vfV, vfG, actdz_bytes) + (size_t)nslot * slot_bytes;
}

__device__ __forceinline__ void tc_carve(TcSmem& s, unsigned char* base, const TcArgs& a) {
    const int vfA = a.nA.vec_floats, vfV = a.nV.vec_floats, vfG = a.nG.vec_floats;
    unsigned char* p = reinterpret_cast<unsigned char*>(((uintptr_t)base + 1023) & ~(uintptr_t)1023);
    s.act = p; p += a.actdz_bytes;
    s.dz = p; p += a.actdz_bytes;
    s.ring = p; p += (size_t)a.nslot * a.slot_bytes;
    s.vecA = reinterpret_cast<float*>(p); p += (size_t)vfA * 4;
    s.vecV = reinterpret_cast<float*>(p); p += (size_t)vfV * 4;
    s.vecG = reinterpret_cast<float*>(p); p += (size_t)vfG * 4;
    s.full = reinterpret_cast<uint64_t*>(p); p += MAX_NSLOT * 8;
    s.empty = reinterpret_cast<uint64_t*>(p); p += MAX_NSLOT * 8;
    s.acc_full = reinterpret_cast<uint64_t*>(p); p += 8;           // acc_full, a_all, a_chunk[16]: contiguous (PathCtx::bars)
    s.a_all = reinterpret_cast<uint64_t*>(p); p += 8;
    s.a_chunk = reinterpret_cast<uint64_t*>(p); p += 8 * MAX_CHUNK;
    s.act_full = reinterpret_cast<uint64_t*>(p); p += 8;
    s.tslot = reinterpret_cast<uint32_t*>(p); s.tile = reinterpret_cast<int*>(p) + 4; p += 64;
    s.sch = reinterpret_cast<Sched*>(p); p += sizeof(Sched);
    s.pc = reinterpret_cast<ProdCtl*>(p); p += 64;
    s.red = reinterpret_cast<float*>(((uintptr_t)p + 15) & ~(uintptr_t)15);
    p = reinterpret_cast<unsigned char*>(s.red) + 64;
    s.dzmax = reinterpret_cast<uint32_t*>(p); p += 64;
    s.nA = reinterpret_cast<TcNet*>(p); p += sizeof(TcNet);
    s.nV = reinterpret_cast<TcNet*>(p); p += sizeof(TcNet);
    s.nG = reinterpret_cast<TcNet*>(p); p += sizeof(TcNet);
    s.gA = reinterpret_cast<TcSlab*>(p); p += sizeof(TcSlab);
    s.gV = reinterpret_cast<TcSlab*>(p); p += sizeof(TcSlab);
    s.gG = reinterpret_cast<TcSlab*>(p);
}

// common prologue: barriers, TMEM, vector blocks -> shared memory.  Returns the TMEM base.
__device__ __forceinline__ uint32_t tc_setup(TcSmem& s, const TcArgs& a) {
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < a.nA.vec_floats; i += TC_THREADS) s.vecA[i] = a.vecA ? a.vecA[i] : 0.f;
    for (int i = tid; i < a.nV.vec_floats; i += TC_THREADS) s.vecV[i] = a.vecV ? a.vecV[i] : 0.f;
    for (int i = tid; i < a.nG.vec_floats; i += TC_THREADS) s.vecG[i] = a.vecG ? a.vecG[i] : 0.f;
    for (int i = tid; i < 2 * a.actdz_bytes / 4; i += TC_THREADS) reinterpret_cast<uint32_t*>(s.act)[i] = 0u;
    if (tid == 0) { *s.nA = a.nA; *s.nV = a.nV; *s.nG = a.nG; *s.gA = a.gA; *s.gV = a.gV; *s.gG = a.gG; }
    if (tid == 0) {
        for (int i = 0; i < MAX_NSLOT; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
        mbar_init(s.acc_full, 1);
        mbar_init(s.a_all, TC_PATH_THREADS / 32);
        mbar_init(&s.a_chunk[0], TC_PATH_THREADS / 32);                      // chunk 0: every path warp (see for_acc_chunks)
        for (int i = 1; i < MAX_CHUNK; ++i) mbar_init(&s.a_chunk[i], 4);     // the four warps of the group that owns the chunk
        mbar_init(s.act_full, 1);
        s.sch->nops = 0;
        s.pc->req = 0; s.pc->gen = 0; s.pc->quit = 0;
        fence_barrier_init();
    }
    if (warp == TC_CTRL_WARP) tmem_alloc(s.tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *s.tslot;
}

// sum over the work threads (warps 0..8); result valid in thread 0
__device__ __forceinline__ float tc_block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    bar_work_sync(TC_WORK_THREADS);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    bar_work_sync(TC_WORK_THREADS);
    float s = 0.f;
    if (threadIdx.x == 0)
        for (int i = 0; i < TC_WORK_THREADS / 32; ++i) s += red[i];
    return s;
}

// next tile of this CTA (all work threads call it): tiles are handed out through a global counter, so a CTA whose tiles
// ended early (every path of a tile can leave the domain before step N) takes more of them
__device__ __forceinline__ long long tc_next_tile(const TcSmem& S, const TcArgs& a) {
    bar_work_sync(TC_WORK_THREADS);
    if (threadIdx.x == 0) *S.tile = (int)gridDim.x + atomicAdd(a.tile_counter, 1);
    bar_work_sync(TC_WORK_THREADS);
    return *S.tile;
}

// loop over the first n (<= DPX) components with static indices
#define KLOOP(k, n) _Pragma("unroll") for (int k = 0; k < DPX; ++k) if (k < (n))

// increments of step t for one path (same generator and bits as load_dw of the exact path)
template <int DPX>
__device__ __forceinline__ void path_dw(const TcArgs& a, long long gpath_local, bool valid, int t, float (&dw)[DPX]) {
    const int d = a.eqf.d;
    if (a.dw_mode == DW_EXTERNAL) {
        KLOOP(k, d) dw[k] = valid ? a.dw[(gpath_local * d + k) * (long long)a.N + t] : 0.f;
        return;
    }
    uint32_t k0, k1;
    philox_key(a.seed, a.stream + (a.stream_base ? *a.stream_base : 0ull), k0, k1);
    const unsigned long long gp = (unsigned long long)(a.path_offset + gpath_local);
#pragma unroll
    for (int ch = 0; ch < (DPX + 3) / 4; ++ch) {
        if (4 * ch < d) {
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
                if (4 * ch + i < DPX && 4 * ch + i < d) dw[4 * ch + i < DPX ? 4 * ch + i : 0] = o[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// role contexts shared by both kernels
struct Roles {
    Ctrl C;
    PathCtx P;
    bool is_path, is_ctrl, primary;      // primary: the path thread of group 0 (does the global stores of its path)
    int row;
};
__device__ __forceinline__ void roles_init(Roles& r, const TcSmem& S, const TcArgs& a, uint32_t tmem) {
    const int tid = threadIdx.x, warp = (int)warp_uniform(tid >> 5);
    r.is_path = warp < TC_CTRL_WARP;
    r.is_ctrl = (warp == TC_CTRL_WARP);          // the whole warp runs the control protocol (elect_one() issues)
    r.primary = warp < 4;
    r.row = tid & 127;
    Ctrl& C = r.C;
    C.ring = S.ring; C.full = S.full; C.empty = S.empty; C.acc_full = S.acc_full; C.a_all = S.a_all; C.a_chunk = S.a_chunk; C.sch = S.sch;
    C.pc = S.pc; C.n_req = 0; C.n_consumed = 0; C.op_count = 0; C.sync = 0; C.tmem = warp_uniform(tmem); C.gen = 0;
    C.act_full = S.act_full; C.act_count = 0; C.nslot = a.nslot; C.slot_bytes = a.slot_bytes; C.act = S.act; C.dz = S.dz;
    r.P.tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    r.P.grp = (warp >> 2) % TC_NGRP;
    r.P.bars = smem_u32(S.acc_full); r.P.sync = 0; r.P.dexp = 0;
    TC_STAT(r.P.t_accw = 0; r.P.t_epi = 0; r.P.t_hid = 0; r.P.t_drain = 0; r.P.t_mark = clock64();)
    TC_STAT(C.t_aready = 0; C.t_issue = 0; C.t_accw = 0; C.t_dw_ready = 0; C.t_act = 0;)
    C.n_ops = 0;
    C.mm_slot = 0; C.mm_use = 0;
}
// stats row: [0] kernel cycles, ctrl: [1] waiting for the path threads, [2] waiting for weights, [3] ops;
// path thread 0: [4] waiting for the tensor pipe, [5] epilogue (wake-up -> publish)
__device__ __forceinline__ void roles_stats(const Roles& r, const TcArgs& a, long long t_start) {
    if (!a.stats) return;
    long long* st = a.stats + (size_t)blockIdx.x * 16;
    if (r.is_ctrl && (threadIdx.x & 31) == 0) {
        st[0] = clock64() - t_start; st[3] = r.C.n_ops;
        TC_STAT(st[1] = r.C.t_aready; st[2] = r.C.t_dw_ready; st[7] = r.C.t_issue; st[8] = r.C.t_accw;)
    }
    TC_STAT(if (threadIdx.x == 0) { st[4] = r.P.t_accw; st[5] = r.P.t_epi; st[6] = r.P.t_hid; })
}

// Input-layer sums SX[k] = sum x_k dy0_k, S0[k] = sum dy0_k.  The two threads of a path split the components: group g
// (0: threads 0..127, 1: threads 128..255) accumulates k in [g*DH, g*DH + DH), DH = DPX/2, in DH registers each (all
// indices static -- a run-time index anywhere would put the accumulators in local memory).
template <int DPX>
__device__ __forceinline__ void acc_input_sums(float (&sx)[DPX / TC_NGRP], float (&s0)[DPX / TC_NGRP], const float (&x)[DPX], const float (&dy0)[DPX], int grp, int d) {
    constexpr int DH = DPX / TC_NGRP;
#pragma unroll
    for (int j = 0; j < DH; ++j) {
        const float xs = (TC_NGRP > 1 && grp) ? x[(DH + j) % DPX] : x[j], ds = (TC_NGRP > 1 && grp) ? dy0[(DH + j) % DPX] : dy0[j];
        if (grp * DH + j < d) { sx[j] += xs * ds; s0[j] += ds; }
    }
}
// kernel end: sum over the path threads of each group -> atomicAdd into dst[g*DH + j]
template <int DH>
__device__ __forceinline__ void reduce_rows_to(float* dst, const float (&acc)[DH], int d, int grp, bool is_path) {
#pragma unroll
    for (int j = 0; j < DH; ++j) {
        float v = is_path ? acc[j] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int k = grp * DH + j;
        if (is_path && k < d && (threadIdx.x & 31) == 0) atomicAdd(dst + k, v);
    }
}

// ================================================================================== critic (tensor)
// DP > 0: instantiation for dim (and control_dim + 1) <= DP with equation EQN fixed at compile time -- the
// per-path vectors are register arrays and every d-loop is unrolled (MV > 0: VDP with control_dim = MV, which makes its
// cyclic neighbour indices static); <0,-1,0>: generic run-time version.
// The body is compiled once per role (IS_PATH: a path warp, otherwise the control warp): the same source, so both roles run
// the same sequence of CTA barriers by construction, but each role's code is a branch of its own -- its registers are
// allocated separately, which is what lets `setmaxnreg` give the path warps more than the launch bound.
template <int DP, int EQN, int MV, bool IS_PATH>
__device__ __forceinline__ void critic_tc_body(const TcArgs& a, const TcSmem& S, Roles& R) {
    constexpr int DPX = DP > 0 ? DP : 32;
    Ctrl& C = R.C;
    PathCtx& P = R.P;
    const TcNet& nA = *S.nA;
    const TcNet& nV = *S.nV;
    const TcNet& nG = *S.nG;
    const TcSlab& gA = *S.gA;
    const TcSlab& gV = *S.gV;
    const TcSlab& gG = *S.gG;
    (void)gA; (void)gV; (void)gG;
    const int tid = threadIdx.x;
    constexpr bool is_path = IS_PATH, is_ctrl = !IS_PATH;
    const bool primary = R.primary;
    const int row = R.row;
    const Eq<float>& E = a.eqf;                                   // kernel-parameter space: fields are constant-bank operands
    const int d = E.d, N = a.N, sr = a.sr;
    const bool cheat = a.flags & F_CHEAT_CONTROL, prop_only = a.flags & F_PROPAGATE_ONLY;
    const bool need_grad = (a.flags & F_NEED_GRAD) && !prop_only;
    const bool td1 = (E.td == 1) && !prop_only;
    const float scale = 100.f * a.invB;
    const float fill = 0.5f * E.R / sqrtf((float)d);
    float* traj = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;               // [N][2*sr][128]
    unsigned char* copies = a.copies ? a.copies + (size_t)blockIdx.x * a.copies_per_cta : nullptr;
    float* gsV = a.slabV ? a.slabV + (size_t)(blockIdx.x % a.nslab) * gV.gtotal : nullptr;
    float* gsG = a.slabG ? a.slabG + (size_t)(blockIdx.x % a.nslab) * gG.gtotal : nullptr;

    float loss0 = 0.f, loss1 = 0.f;
    TC_STAT(long long ph_roll = 0, ph_val = 0, ph_grad = 0;)          // cycles per phase (diagnostics)
    TC_STAT(long long seg_dw = 0, seg_A = 0, seg_mv = 0, seg_G = 0;)
    const long long ntiles = (a.B_local + TC_PATHS - 1) / TC_PATHS;
    for (long long tile = blockIdx.x; tile < ntiles; tile = tc_next_tile(S, a)) {
        const long long base = tile * TC_PATHS;
        const long long slot = base + row;                        // position of this thread's path in the tiling
        const bool valid = is_path && slot < a.B_local;
        const long long gp = (valid && a.perm) ? (long long)a.perm[slot] : slot;      // local path index of this thread
        const bool wr = valid && primary;               // this thread does the global stores of its path
        TC_STAT(const long long tp0 = clock64();)
        float x[DPX], u[DPX], dwv[DPX], sdw[DPX], g[DPX], raw[DPX];
        int flag = 0, nacc = 0;
        float disc = 1.f, y = 0.f;
        if (is_path) {
            KLOOP(k, d) x[k] = valid ? a.x0[gp * d + k] : fill;
            flag = fwd_initial_flag<float, DP, EQN, MV>(E, x, 1, 0);
            if (a.o_x && wr)
                KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1)] = x[k];
        }
        if (is_ctrl) {                                            // schedule of the rollout
            ctrl_flush(C);
            if (!cheat) sched_add_fwd(C, nA, a.imgA, nA.L);
            if (td1) sched_add_fwd(C, nG, a.imgG, nG.L);
            ctrl_sched_ready(C);
        }
        // ------------------------------------------------------------------ sweep 1: rollout
        int tlive = 0;
        for (int t = 0; t < N; ++t) {
            const int alive = bar_work_or(valid && flag > 0, TC_WORK_THREADS);
            if (!alive) break;
            tlive = t + 1;
            if (is_ctrl) {
                if (!cheat) ctrl_net_forward(C, nA, nA.L);
                if (td1) ctrl_net_forward(C, nG, nG.L);
            } else if (is_path) {
                // per-path arithmetic is placed where the tensor pipe is busy with a first layer
                if (!cheat) { path_net_begin(P, nA, S.vecA, x); path_hidden_range(P, nA, S.vecA, 0, 1); }
                TC_STAT(const long long q1 = clock64();)
                path_dw(a, gp, valid, t, dwv);
                TC_STAT(const long long q1b = clock64(); seg_dw += q1b - q1;)
                if (!cheat) path_hidden_range(P, nA, S.vecA, 1, 2);
                float dt, sqdt, xn; int dtg;
                TC_STAT(const long long q2 = clock64();)
                fwd_dt<float, DP, EQN, MV>(E, x, flag, 1, 0, dt, sqdt, xn, dtg);
                TC_STAT(seg_A += clock64() - q2;)
                if (cheat) {
                    eq_u_true<float, DP, EQN, MV>(E, x, u, 1, 0);
                } else {
                    path_net_finish(P, nA, S.vecA, raw, 2);
                    if (nA.ekn_head) ekn_head_fwd<float, DP, EQN, MV>(raw, u, nA.mctrl, 1, 0);
                    else KLOOP(j, E.m) u[j] = raw[j];
                }
                if (td1) path_net_begin(P, nG, S.vecG, x);                        // NN_value_grad at x_t (before the move)
                TC_STAT(const long long q3 = clock64();)
                float* tr = traj + (size_t)t * 2 * sr * TC_PATHS;
                if (need_grad && td1 && primary)
                    KLOOP(k, d) __stcs(&tr[k * TC_PATHS + row], x[k]);
                float w = 0.f;
                if (!prop_only) w = eq_w<float, DP, EQN, MV>(E, x, u, 1, 0);
                TC_STAT(const long long q3b = clock64(); seg_mv += q3b - q3;)
                if (td1) path_hidden_range(P, nG, S.vecG, 0, 1);
                TC_STAT(const long long q3c = clock64();)
                const int coef = fwd_move<float, DP, EQN, MV>(E, x, u, dwv, dt, sqdt, xn, flag, sdw, 1, 0);
                const float cf = (float)coef;
                TC_STAT(const long long q4 = clock64(); seg_mv += q4 - q3c;)
                if (td1) path_net_finish(P, nG, S.vecG, g, 1);
                TC_STAT(const long long q5 = clock64();)
                y = y + w * disc * cf * dt;                                       // solver.py:170-174
                if (td1) {
                    float dif = 0.f;
                    KLOOP(k, d) dif = dif + sdw[k] * g[k];        // solver.py:177-182
                    dif = dif * disc;
                    y = y - dif * cf * sqdt;                                      // solver.py:184
                    if (need_grad && primary) {
                        const float q = disc * cf * sqdt;
                        KLOOP(k, d) __stcs(&tr[(sr + k) * TC_PATHS + row], sdw[k] * q);
                    }
                }
                disc = disc * expf(-E.gamma * dt * cf);                          // solver.py:187
                TC_STAT(seg_G += clock64() - q5;)
                nacc += coef;
                if (wr) {
                    if (a.o_dt) a.o_dt[gp * N + t] = dt;
                    if (a.o_coef) a.o_coef[gp * N + t] = cf;
                    if (a.o_x)
                        KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
                }
            }
        }
        if (wr) {
            for (int t = tlive; t < N; ++t) {
                if (a.o_dt) a.o_dt[gp * N + t] = E.delta_t;
                if (a.o_coef) a.o_coef[gp * N + t] = 0.f;
                if (a.o_x)
                    KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
            }
            if (a.o_exit) a.o_exit[gp] = nacc;
        }
        TC_STAT(const long long tp1 = clock64(); ph_roll += tp1 - tp0;)
        if (prop_only) continue;
        // ------------------------------------------------------------------ NN_value at x_0, x_N, x_bdry
        float rho_v = 0.f, rho_b = 0.f, rhog = 0.f;
        if (is_ctrl) {
            ctrl_flush(C);
            if (!need_grad) {
                sched_add_fwd(C, nV, a.imgV, nV.L);
                ctrl_sched_ready(C);
                for (int i = 0; i < 3; ++i) ctrl_net_forward(C, nV, nV.L);
            } else {
                sched_add_fwd(C, nV, a.imgV, nV.L);                       // V(x_0), forward only
                for (int i = 0; i < 3; ++i) { sched_add_fwd(C, nV, a.imgV, nV.L); sched_add_bwd(C, nV, a.imgV); }
                ctrl_sched_ready(C);
                ctrl_net_forward(C, nV, nV.L);
                for (int i = 0; i < 3; ++i) { ctrl_net_forward(C, nV, nV.L); ctrl_net_backward(C, nV, true, copies); }
            }
        } else if (is_path) {
            float vN[1], v0[1], vb[1], x0v[DPX], xbv[DPX], dy0[DPX], cot[1];
            KLOOP(k, d) x0v[k] = valid ? a.x0[gp * d + k] : fill;
            KLOOP(k, d) xbv[k] = valid ? a.xb[gp * d + k] : fill;
            if (!need_grad) {
                path_net_forward(P, nV, S.vecV, x0v, v0);
                path_net_forward(P, nV, S.vecV, x, vN);
                path_net_forward(P, nV, S.vecV, xbv, vb);
            } else {
                Masks mk;
                // (the input-layer sums of NN_value are flushed once per tile: three updates do not justify registers that
                //  stay live through both sweeps)
                float sxV[DPX / TC_NGRP], s0V[DPX / TC_NGRP];
#pragma unroll
                for (int k = 0; k < DPX / TC_NGRP; ++k) { sxV[k] = 0.f; s0V[k] = 0.f; }
                path_net_forward(P, nV, S.vecV, x0v, v0);
                path_net_forward_keep(P, nV, S.vecV, x, vN, mk, copies, S.act, row, false);
                const float delta = v0[0] - y - vN[0] * disc;
                rhog = valid ? rho_grad(delta, 50.f) * scale : 0.f;
                cot[0] = -rhog * disc;
                path_net_backward(P, nV, gV, mk, cot, true, gsV, S.dz, row, dy0, S.dzmax);
                acc_input_sums<DPX>(sxV, s0V, x, dy0, P.grp, d);
                path_net_forward_keep(P, nV, S.vecV, x0v, v0, mk, copies, S.act, row, false);
                cot[0] = rhog;
                path_net_backward(P, nV, gV, mk, cot, true, gsV, S.dz, row, dy0, S.dzmax);
                acc_input_sums<DPX>(sxV, s0V, x0v, dy0, P.grp, d);
                path_net_forward_keep(P, nV, S.vecV, xbv, vb, mk, copies, S.act, row, false);
                const float dbb = vb[0] - eq_Z<float, DP, EQN, MV>(E, xbv, 1, 0);
                cot[0] = valid ? rho_grad(dbb, 50.f) * scale : 0.f;
                path_net_backward(P, nV, gV, mk, cot, true, gsV, S.dz, row, dy0, S.dzmax);
                acc_input_sums<DPX>(sxV, s0V, xbv, dy0, P.grp, d);
                reduce_rows_to(gsV + gV.gX, sxV, d, P.grp, true);
                reduce_rows_to(gsV + gV.g0, s0V, d, P.grp, true);
            }
            const float delta = v0[0] - y - vN[0] * disc;                         // solver.py:189
            const float db = vb[0] - eq_Z<float, DP, EQN, MV>(E, xbv, 1, 0);                          // solver.py:190
            if (wr) {
                rho_v = rho(delta, 50.f);
                rho_b = rho(db, 50.f);
                if (a.o_delta) a.o_delta[gp] = delta;
                if (a.o_delta_b) a.o_delta_b[gp] = db;
            }
        }
        loss0 += tc_block_sum(rho_v, S.red);
        loss1 += tc_block_sum(rho_b, S.red);
        TC_STAT(const long long tp2 = clock64(); ph_val += tp2 - tp1;)
        // ------------------------------------------------------------------ sweep 2: NN_value_grad backward
        if (need_grad && td1) {
            if (is_ctrl) {
                ctrl_flush(C);
                sched_add_fwd(C, nG, a.imgG, nG.L - 1);
                sched_add_bwd(C, nG, a.imgG);
                ctrl_sched_ready(C);
                for (int t = 0; t < tlive; ++t) {
                    ctrl_net_forward(C, nG, nG.L - 1);
                    ctrl_net_backward(C, nG, true, copies);
                }
            } else if (is_path) {
                Masks mk;
                float xt[DPX], cot[DPX], dy0[DPX];
                // (input-layer sums of this tile: live in this sweep only, flushed into the slab at its end)
                float sxG[DPX / TC_NGRP], s0G[DPX / TC_NGRP];
#pragma unroll
                for (int k = 0; k < DPX / TC_NGRP; ++k) { sxG[k] = 0.f; s0G[k] = 0.f; }
                for (int t = 0; t < tlive; ++t) {
                    const float* tr = traj + (size_t)t * 2 * sr * TC_PATHS;
                    KLOOP(k, d) { xt[k] = __ldcs(&tr[k * TC_PATHS + row]); cot[k] = __ldcs(&tr[(sr + k) * TC_PATHS + row]) * rhog; }
                    float unused[1];
                    path_net_forward_keep(P, nG, S.vecG, xt, unused, mk, copies, S.act, row, true);
                    path_net_backward(P, nG, gG, mk, cot, true, gsG, S.dz, row, dy0, S.dzmax);
                    acc_input_sums<DPX>(sxG, s0G, xt, dy0, P.grp, d);
                }
                reduce_rows_to(gsG + gG.gX, sxG, d, P.grp, true);
                reduce_rows_to(gsG + gG.g0, s0G, d, P.grp, true);
            }
        }
        TC_STAT(ph_grad += clock64() - tp2;)
    }
    TC_STAT(if (a.stats && tid == 0) { long long* st = a.stats + (size_t)blockIdx.x * 16; st[9] = ph_roll; st[10] = ph_val; st[11] = ph_grad; st[12] = seg_dw; st[13] = seg_A; st[14] = seg_mv; st[15] = seg_G; })
    if (is_ctrl) { ctrl_flush(C); C.pc->quit = 1; }
    if (tid == 0 && a.loss_part) {
        a.loss_part[blockIdx.x * 2] = loss0;
        a.loss_part[blockIdx.x * 2 + 1] = loss1;
    }
}

// common prologue / role dispatch / epilogue of both kernels
#define DPB_TC_KERNEL(NAME)                                                                                                   \
    template <int DP, int EQN, int MV>                                                                                        \
    __global__ void __launch_bounds__(TC_THREADS, 1) NAME##_tc_kernel(const TcArgs a) {                                       \
        extern __shared__ __align__(1024) unsigned char smem_raw[];                                                           \
        TcSmem S;                                                                                                             \
        tc_carve(S, smem_raw, a);                                                                                             \
        const uint32_t tmem = tc_setup(S, a);                                                                                 \
        Roles R;                                                                                                              \
        roles_init(R, S, a, tmem);                                                                                            \
        const long long t_start = clock64();                                                                                  \
        const int warp = threadIdx.x >> 5;                                                                                    \
        if (warp >= TC_CTRL_WARP) {                      /* control, producer (lane 0 streams the weights), idle warps */     \
            tc_regs_release();                                                                                                \
            if (warp == TC_PROD_WARP) {                                                                                       \
                if ((threadIdx.x & 31) == 0) producer_loop(S.ring, S.full, S.empty, S.sch, S.pc, a.nslot, a.slot_bytes);      \
            } else if (warp == TC_CTRL_WARP) {                                                                                \
                NAME##_tc_body<DP, EQN, MV, false>(a, S, R);                                                                  \
            }                                                                                                                 \
        } else {                                                                                                              \
            tc_regs_take();                                                                                                   \
            NAME##_tc_body<DP, EQN, MV, true>(a, S, R);                                                                       \
        }                                                                                                                     \
        roles_stats(R, a, t_start);                                                                                           \
        tc_fence_before();                                                                                                    \
        __syncthreads();                                                                                                      \
        if (warp == TC_CTRL_WARP) tmem_dealloc(tmem, 512);                                                                    \
    }

// =================================================================================== actor (tensor)
// The body is compiled once per role (IS_PATH: a path warp, otherwise the control warp): the same source, so both roles run
// the same sequence of CTA barriers by construction, but each role's code is a branch of its own -- its registers are
// allocated separately, which is what lets `setmaxnreg` give the path warps more than the launch bound.
template <int DP, int EQN, int MV, bool IS_PATH>
__device__ __forceinline__ void actor_tc_body(const TcArgs& a, const TcSmem& S, Roles& R) {
    constexpr int DPX = DP > 0 ? DP : 32;
    Ctrl& C = R.C;
    PathCtx& P = R.P;
    const TcNet& nA = *S.nA;
    const TcNet& nV = *S.nV;
    const TcNet& nG = *S.nG;
    const TcSlab& gA = *S.gA;
    const TcSlab& gV = *S.gV;
    const TcSlab& gG = *S.gG;
    (void)gA; (void)gV; (void)gG; (void)nG;
    const int tid = threadIdx.x;
    constexpr bool is_path = IS_PATH, is_ctrl = !IS_PATH;
    const bool primary = R.primary;
    const int row = R.row;
    const Eq<float>& E = a.eqf;                                   // kernel-parameter space: fields are constant-bank operands
    const int d = E.d, m = E.m, N = a.N, sr = a.sr;
    const bool cheat = a.flags & F_CHEAT_CONTROL, cheat_v = a.flags & F_CHEAT_VALUE;
    const bool need_grad = (a.flags & F_NEED_GRAD) && !cheat;
    const float fill = 0.5f * E.R / sqrtf((float)d);
    const int trs = 2 * sr + A_NSCAL;
    float* traj = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;               // [N][2*sr + A_NSCAL][128]
    unsigned char* copies = a.copies ? a.copies + (size_t)blockIdx.x * a.copies_per_cta : nullptr;
    float* gsA = a.slabA ? a.slabA + (size_t)(blockIdx.x % a.nslab) * gA.gtotal : nullptr;

    // (kept for the whole kernel: scoping them to the reverse sweep of a tile, as the critic does with its sums, made the
    //  actor 1.6 % slower)
    float sxA[DPX / TC_NGRP], s0A[DPX / TC_NGRP];
#pragma unroll
    for (int k = 0; k < DPX / TC_NGRP; ++k) { sxA[k] = 0.f; s0A[k] = 0.f; }

    float loss0 = 0.f;
    TC_STAT(long long seg_fk = 0, seg_adj = 0, seg_bwd = 0, seg_fwd = 0;)   // reverse step: forward_keep / adjoint step / backward; forward rollout
    const long long ntiles = (a.B_local + TC_PATHS - 1) / TC_PATHS;
    for (long long tile = blockIdx.x; tile < ntiles; tile = tc_next_tile(S, a)) {
        const long long base = tile * TC_PATHS;
        const long long slot = base + row;
        const bool valid = is_path && slot < a.B_local;
        const long long gp = (valid && a.perm) ? (long long)a.perm[slot] : slot;
        TC_STAT(const long long f0 = clock64();)
        const bool wr = valid && primary;               // this thread does the global stores of its path
        float x[DPX], u[DPX], dwv[DPX], raw[DPX];
        int flag = 0, nacc = 0;
        float disc = 1.f, y = 0.f;
        if (is_path) {
            KLOOP(k, d) x[k] = valid ? a.x0[gp * d + k] : fill;
            flag = fwd_initial_flag<float, DP, EQN, MV>(E, x, 1, 0);
            if (a.o_x && wr)
                KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1)] = x[k];
        }
        if (is_ctrl) {
            ctrl_flush(C);
            if (!cheat) sched_add_fwd(C, nA, a.imgA, nA.L);
            ctrl_sched_ready(C);
        }
        // ------------------------------------------------------------------ forward rollout
        int tlive = 0;
        for (int t = 0; t < N; ++t) {
            const int alive = bar_work_or(valid && flag > 0, TC_WORK_THREADS);
            if (!alive) break;
            tlive = t + 1;
            if (is_ctrl) {
                if (!cheat) ctrl_net_forward(C, nA, nA.L);
            } else if (is_path) {
                if (!cheat) { path_net_begin(P, nA, S.vecA, x); path_hidden_range(P, nA, S.vecA, 0, 1); }
                path_dw(a, gp, valid, t, dwv);
                if (!cheat) path_hidden_range(P, nA, S.vecA, 1, 2);
                float dt, sqdt, xn; int dtg;
                fwd_dt<float, DP, EQN, MV>(E, x, flag, 1, 0, dt, sqdt, xn, dtg);
                if (cheat) {
                    eq_u_true<float, DP, EQN, MV>(E, x, u, 1, 0);
                } else {
                    path_net_finish(P, nA, S.vecA, raw, 2);
                    if (nA.ekn_head) ekn_head_fwd<float, DP, EQN, MV>(raw, u, nA.mctrl, 1, 0);
                    else KLOOP(j, m) u[j] = raw[j];
                }
                float* tr = traj + (size_t)t * trs * TC_PATHS;
                if (need_grad && primary)
                    KLOOP(k, d) { __stcs(&tr[k * TC_PATHS + row], x[k]); __stcs(&tr[(sr + k) * TC_PATHS + row], dwv[k]); }
                const float w = eq_w<float, DP, EQN, MV>(E, x, u, 1, 0);
                const int coef = fwd_move<float, DP, EQN, MV>(E, x, u, dwv, dt, sqdt, xn, flag, (float*)nullptr, 1, 0);
                const float cf = (float)coef;
                if (need_grad && primary) {
                    float* sc = tr + (size_t)2 * sr * TC_PATHS;
                    sc[A_DT * TC_PATHS + row] = dt; sc[A_SQDT * TC_PATHS + row] = sqdt; sc[A_COEF * TC_PATHS + row] = valid ? cf : 0.f;
                    sc[A_DISC * TC_PATHS + row] = disc; sc[A_XN * TC_PATHS + row] = xn; sc[A_DTG * TC_PATHS + row] = (float)dtg;
                }
                y = y + cf * w * dt * disc;                                       // solver.py:218
                disc = disc * expf(-E.gamma * dt * cf);                          // solver.py:219
                nacc += coef;
                if (wr) {
                    if (a.o_dt) a.o_dt[gp * N + t] = dt;
                    if (a.o_coef) a.o_coef[gp * N + t] = cf;
                    if (a.o_x)
                        KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
                }
            }
        }
        if (wr) {
            for (int t = tlive; t < N; ++t) {
                if (a.o_dt) a.o_dt[gp * N + t] = E.delta_t;
                if (a.o_coef) a.o_coef[gp * N + t] = 0.f;
                if (a.o_x)
                    KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
            }
            if (a.o_exit) a.o_exit[gp] = nacc;
        }
        TC_STAT(seg_fwd += clock64() - f0;)
        // ------------------------------------------------------------------ terminal value (+ its input gradient)
        float yv = 0.f;
        float lam[DPX];
        float Dbar = 0.f;
        if (is_ctrl) {
            ctrl_flush(C);
            if (!cheat_v) {
                sched_add_fwd(C, nV, a.imgV, nV.L);
                if (need_grad) sched_add_bwd(C, nV, a.imgV);
                ctrl_sched_ready(C);
                ctrl_net_forward(C, nV, nV.L);
                if (need_grad) ctrl_net_backward(C, nV, false, nullptr);
            }
        } else if (is_path) {
            float vN[1];
            const float seed = valid ? disc * a.invB : 0.f;
            if (cheat_v) {
                vN[0] = eq_V_true<float, DP, EQN, MV>(E, x, 1, 0);                                    // solver.py:223
                if (need_grad) {
                    eq_V_grad_true<float, DP, EQN, MV>(E, x, lam, 1, 0);
                    KLOOP(k, d) lam[k] = lam[k] * seed;
                }
            } else if (!need_grad) {
                path_net_forward(P, nV, S.vecV, x, vN);                         // solver.py:221
            } else {
                Masks mk;
                float cot[1], dy0[DPX];
                path_net_forward_keep(P, nV, S.vecV, x, vN, mk, nullptr, nullptr, row, false);
                cot[0] = seed;
                path_net_backward(P, nV, gV, mk, cot, false, nullptr, nullptr, row, dy0, S.dzmax);
                const float* g0c = S.vecV + nV.vec_g0;
                KLOOP(k, d) lam[k] = dy0[k] * g0c[k];
            }
            Dbar = valid ? vN[0] * a.invB : 0.f;
            y = y + vN[0] * disc;
            if (wr) {
                yv = y;
                if (a.o_delta) a.o_delta[gp] = y;
            }
        }
        loss0 += tc_block_sum(yv, S.red);
        if (!need_grad) continue;
        // ------------------------------------------------------------------ reverse sweep
        if (is_ctrl) {
            ctrl_flush(C);
            sched_add_fwd(C, nA, a.imgA, nA.L);
            sched_add_bwd(C, nA, a.imgA);
            ctrl_sched_ready(C);
        }
        for (int t = tlive - 1; t >= 0; --t) {
            const float* tr = traj + (size_t)t * trs * TC_PATHS;
            const float* sc = tr + (size_t)2 * sr * TC_PATHS;
            const int any = bar_work_or(valid && sc[A_COEF * TC_PATHS + row] > 0.f, TC_WORK_THREADS);
            if (!any) continue;
            if (is_ctrl) {
                ctrl_net_forward(C, nA, nA.L);
                ctrl_net_backward(C, nA, true, copies);
            } else if (is_path) {
                Masks mk;
                float xt[DPX], ubar[DPX], cot[DPX], dy0[DPX];
                TC_STAT(const long long r0 = clock64();)
                KLOOP(k, d) { xt[k] = __ldcs(&tr[k * TC_PATHS + row]); dwv[k] = __ldcs(&tr[(sr + k) * TC_PATHS + row]); }
                path_net_forward_keep(P, nA, S.vecA, xt, raw, mk, copies, S.act, row, false);
                TC_STAT(const long long r1 = clock64(); seg_fk += r1 - r0;)
                if (nA.ekn_head) ekn_head_fwd<float, DP, EQN, MV>(raw, u, nA.mctrl, 1, 0);
                else KLOOP(j, m) u[j] = raw[j];
                const int coef = (valid && sc[A_COEF * TC_PATHS + row] > 0.f) ? 1 : 0;
                adj_step<float, DP, EQN, MV>(E, xt, u, dwv, sc[A_DT * TC_PATHS + row], sc[A_SQDT * TC_PATHS + row], coef, (int)sc[A_DTG * TC_PATHS + row],
                         sc[A_XN * TC_PATHS + row], sc[A_DISC * TC_PATHS + row], a.invB, lam, Dbar, ubar, 1, 0);
                if (nA.ekn_head) {
                    if (coef) ekn_head_bwd<float, DP, EQN, MV>(raw, ubar, cot, m, 1, 0);
                    else KLOOP(j, m + 1) cot[j] = 0.f;
                } else {
                    KLOOP(j, m) cot[j] = ubar[j];
                }
                TC_STAT(const long long r2 = clock64(); seg_adj += r2 - r1;)
                path_net_backward(P, nA, gA, mk, cot, true, gsA, S.dz, row, dy0, S.dzmax);
                TC_STAT(seg_bwd += clock64() - r2;)
                const float* g0c = S.vecA + nA.vec_g0;
                acc_input_sums<DPX>(sxA, s0A, xt, dy0, P.grp, d);
                KLOOP(k, d) lam[k] = lam[k] + dy0[k] * g0c[k];
            }
        }
    }
    if (is_ctrl) { ctrl_flush(C); C.pc->quit = 1; }
    if (need_grad) {
        reduce_rows_to(gsA + gA.gX, sxA, d, P.grp, is_path);
        reduce_rows_to(gsA + gA.g0, s0A, d, P.grp, is_path);
    }
    if (tid == 0 && a.loss_part) {
        a.loss_part[blockIdx.x * 2] = loss0;
        a.loss_part[blockIdx.x * 2 + 1] = 0.f;
    }
    TC_STAT(if (a.stats) { long long* st = a.stats + (size_t)blockIdx.x * 16;
                           if (is_ctrl && (tid & 31) == 0) st[9] = C.t_act;
                           if (tid == 0) { st[11] = P.t_drain; st[12] = seg_fwd; st[13] = seg_fk; st[14] = seg_adj; st[15] = seg_bwd; } })
}

DPB_TC_KERNEL(critic)
DPB_TC_KERNEL(actor)

}  // namespace tc
}  // namespace dpb
