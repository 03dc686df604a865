// dpb_tc_kernels.cuh -- the fused rollout + TD kernels with the MLP layers on tcgen05 (impl = tensor).
// Same algorithm and per-path arithmetic (dpb_eqn.h) as dpb_kernels.cuh.  CTA = one tile of 128 paths = the 128
// TMEM lanes.  Warp roles (dpb_tc_nets.cuh): 4 owner warps (thread t owns path t: state in registers, per-path SDE
// arithmetic, network inputs / outputs), 4 * TC_NGRP stateless helper warps (hidden-layer epilogues, dW drains), the
// control warp (all lanes run the protocol, the elected lane issues every tcgen05.mma) and the producer warp (lane 0
// streams the weights).  The body of each kernel is one source compiled once per role, so that all roles walk through the
// same sequence of products and CTA barriers by construction.
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
    int img_rep;                    // copies of each image, img_stride* bytes apart: CTA i streams copy i % img_rep
    long long img_strideA, img_strideV, img_strideG;
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
    unsigned long long* trace;      // [3][TC_TRACE_CAP] event trace of CTA 0 (stats builds), may be NULL
    int* tile_counter;              // zeroed before the launch: CTAs take tile blockIdx.x first, then gridDim.x + counter++
    const int* perm;                // optional: slot i of the tiling works on local path perm[i] (naive scheme: paths sorted by
                                    // lifetime so that the tiles die as a whole; NULL: identity)
    // Small batches (fewer tiles than half the SMs): the second sweep of the critic -- one independent backward per stored
    // step -- is cut out of the rollout launch (F_S2_DEFER) and run by a second launch of the same kernel (F_S2_ONLY) that
    // deals (tile, block of s2_chunk steps) items out to ALL SMs.
    float* s2_rhog;                 // [tiles][128] rho'(delta) * 100 / B of every path (written by the first launch)
    int* s2_tlive;                  // [tiles] steps the tile was alive
    int s2_chunk;
};
constexpr unsigned F_S2_DEFER = 0x100u, F_S2_ONLY = 0x200u;       // internal flags (above the DPB_FLAG_* bits)

// the kernels are instantiated in their own translation units (dpb_tc_inst_*.cu)
typedef void (*TcKernelFn)(const TcArgs);
#define DPB_TC_FOR_INSTANCES(X) X(lqr) X(ekn) X(lqrvar) X(vdp2) X(vdp5) X(vdp10) X(lqr12) X(ekn12) X(lqrvar12) X(generic)
#define DPB_TC_DECL_GETTERS(n) TcKernelFn tc_get_critic_##n(); TcKernelFn tc_get_actor_##n();
DPB_TC_FOR_INSTANCES(DPB_TC_DECL_GETTERS)

struct TcSmem {
    unsigned char *act, *dz;
    unsigned char* ring;
    float *vecA, *vecV, *vecG;
    uint64_t *full, *empty, *bars, *act_full;      // bars: the hand-off barriers (BAR_* in dpb_tc_nets.cuh)
    uint32_t* tslot;
    int* tile;                       // the tile a CTA works on next (dynamic tile scheduler)
    Sched* sch;
    ProdCtl* pc;
    float* red;
    uint32_t* dzmax;                 // [2][8]: exchange of the tile's largest |cotangent| among the owner warps (own_put_dz); [16]: its exponent
    TcNet *nA, *nV, *nG;             // shared-memory copies of the network descriptors
    TcSlab *gA, *gV, *gG;
};

// everything except the ring
__host__ __device__ inline size_t tc_smem_fixed(int vfA, int vfV, int vfG, int actdz_bytes) {
    return 2 * (size_t)actdz_bytes + (size_t)(vfA + vfV + vfG) * 4 + (2 * MAX_NSLOT + 1 + NUM_HANDOFF_BARS) * 8 + 64 + sizeof(Sched) + 64 + 64 + 128 + 3 * sizeof(TcNet) + 3 * sizeof(TcSlab) + 64 + 1024;
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
    s.bars = reinterpret_cast<uint64_t*>(p); p += 8 * NUM_HANDOFF_BARS;
    s.act_full = reinterpret_cast<uint64_t*>(p); p += 8;
    s.tslot = reinterpret_cast<uint32_t*>(p); s.tile = reinterpret_cast<int*>(p) + 4; p += 64;
    s.sch = reinterpret_cast<Sched*>(p); p += sizeof(Sched);
    s.pc = reinterpret_cast<ProdCtl*>(p); p += 64;
    s.red = reinterpret_cast<float*>(((uintptr_t)p + 15) & ~(uintptr_t)15);
    p = reinterpret_cast<unsigned char*>(s.red) + 64;
    s.dzmax = reinterpret_cast<uint32_t*>(p); p += 128;            // [2][8] maxima + the exponent word
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
        mbar_init(&s.bars[BAR_ACC], 1);
        mbar_init(&s.bars[BAR_FIN], 1);
        mbar_init(&s.bars[BAR_DW], 1);
        mbar_init(&s.bars[BAR_HELP], TC_EPI_WARPS);
        mbar_init(&s.bars[BAR_OWN], TC_OWN_THREADS / 32);
        mbar_init(&s.bars[BAR_CHUNK], TC_EPI_WARPS);                                    // chunk 0: every helper warp (see for_acc_chunks)
        for (int i = 1; i < MAX_CHUNK; ++i) mbar_init(&s.bars[BAR_CHUNK + i], 4);       // the four warps of the group that owns the chunk
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
enum { ROLE_OWN = 0, ROLE_HELP = 1, ROLE_CTRL = 2 };
struct Roles {
    Ctrl C;
    PathCtx P;
    int row;                             // owners / helpers: lane of the tile = path slot this thread works on
};
__device__ __forceinline__ void roles_init(Roles& r, const TcSmem& S, const TcArgs& a, uint32_t tmem) {
    const int tid = threadIdx.x, warp = (int)warp_uniform(tid >> 5);
    r.row = tid & 127;
    Ctrl& C = r.C;
    C.ring = S.ring; C.full = S.full; C.empty = S.empty; C.bars = S.bars; C.sch = S.sch;
    C.pc = S.pc; C.n_req = 0; C.n_consumed = 0; C.op_count = 0; C.dw_count = 0; C.sync = 0; C.tmem = warp_uniform(tmem); C.gen = 0;
    C.act_full = S.act_full; C.act_count = 0; C.nslot = a.nslot; C.slot_bytes = a.slot_bytes; C.act = S.act; C.dz = S.dz;
    C.dexp = reinterpret_cast<volatile int*>(S.dzmax + 16);
    r.P.tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    r.P.grp = warp >= 4 ? (((warp - 4) >> 2) + (TC_COMBINED ? 1 : 0)) % TC_EGRP : 0;
    r.P.bars = smem_u32(S.bars); r.P.sync = 0; r.P.dexp = 0;
    TC_STAT(r.P.t_accw = 0; r.P.t_epi = 0; r.P.t_hid = 0; r.P.t_drain = 0; r.P.t_mark = clock64();)
    TC_STAT(C.t_aready = 0; C.t_issue = 0; C.t_accw = 0; C.t_dw_ready = 0; C.t_act = 0;)
    TC_STAT(const bool tr0 = a.trace && blockIdx.x == 0;)
    TC_STAT(r.P.tn = 0; r.P.tr = (tr0 && tid == 0) ? a.trace : ((tr0 && tid == 128 && TC_HELP_WARPS > 0) ? a.trace + TC_TRACE_CAP : nullptr);)
    TC_STAT(C.tn = 0; C.tr = (tr0 && warp == TC_CTRL_WARP && (tid & 31) == 0) ? a.trace + 2 * TC_TRACE_CAP : nullptr;)
    C.n_ops = 0;
    C.mm_slot = 0; C.mm_use = 0;
}
// stats row: [0] kernel cycles, ctrl: [1] waiting for owners / helpers, [2] dW products waiting for their operands, [3] ops,
// [7] inside issue loops, [8] waiting for the MMAs that read ACT; owner thread 0: [4] waiting for network results, [5] writing
// inputs; helper thread 128: [6] inside hidden-layer epilogues, [9] waiting for the tensor pipe, [10] in dW drains
__device__ __forceinline__ void roles_stats(const Roles& r, const TcArgs& a, long long t_start) {
    if (!a.stats) return;
    long long* st = a.stats + (size_t)blockIdx.x * 16;
    if ((threadIdx.x >> 5) == TC_CTRL_WARP && (threadIdx.x & 31) == 0) {
        st[0] = clock64() - t_start; st[3] = r.C.n_ops;
        TC_STAT(st[1] = r.C.t_aready; st[2] = r.C.t_dw_ready; st[7] = r.C.t_issue; st[8] = r.C.t_accw;)
    }
    TC_STAT(if (threadIdx.x == 0) { st[4] = r.P.t_accw; st[5] = r.P.t_epi; })
    TC_STAT(if (threadIdx.x == 128) { st[6] = r.P.t_hid; st[9] = r.P.t_accw; st[10] = r.P.t_drain; })
}

// ================================================================================== critic (tensor)
// DP > 0: instantiation for dim (and control_dim + 1) <= DP with equation EQN fixed at compile time -- the
// per-path vectors are register arrays and every d-loop is unrolled (MV > 0: VDP with control_dim = MV, which makes its
// cyclic neighbour indices static); <0,-1,0>: generic run-time version.
// The body is compiled once per role (owner / helper / control warp): the same source, so all roles run the same sequence of
// products and CTA barriers by construction, but each role's code is a branch of its own with its own register allocation.
template <int DP, int EQN, int MV, int ROLE>
__device__ __forceinline__ void critic_tc_body(const TcArgs& a, const TcSmem& S, Roles& R) {
    constexpr int DPX = DP > 0 ? DP : 32;
    constexpr bool is_own = ROLE == ROLE_OWN, is_help = ROLE == ROLE_HELP, is_ctrl = ROLE == ROLE_CTRL;
    Ctrl& C = R.C;
    PathCtx& P = R.P;
    const TcNet& nA = *S.nA;
    const TcNet& nV = *S.nV;
    const TcNet& nG = *S.nG;
    const TcSlab& gV = *S.gV;
    const TcSlab& gG = *S.gG;
    (void)gV; (void)gG;
    const int rep = (int)(blockIdx.x % (unsigned)a.img_rep);
    const unsigned char* imgA = a.imgA + rep * a.img_strideA;
    const unsigned char* imgV = a.imgV + rep * a.img_strideV;
    const unsigned char* imgG = a.imgG + rep * a.img_strideG;
    (void)imgA; (void)imgV; (void)imgG;
    const int tid = threadIdx.x;
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
    uint32_t* mxbuf = S.dzmax;
    volatile int* dexp = reinterpret_cast<volatile int*>(S.dzmax + 16);
    (void)traj; (void)copies; (void)gsV; (void)gsG; (void)mxbuf; (void)dexp; (void)scale; (void)fill; (void)sr; (void)row;

    float loss0 = 0.f, loss1 = 0.f;
    const long long ntiles = (a.B_local + TC_PATHS - 1) / TC_PATHS;
    if (a.flags & F_S2_ONLY) {
        // second launch of a small batch: (tile, block of steps) items of the second sweep, dealt out to all CTAs.  Tile i was
        // rolled out by CTA i of the first launch (fewer tiles than CTAs), so its trajectory sits in that CTA's scratch.
        const int nblk = (N + a.s2_chunk - 1) / a.s2_chunk;
        bool first = true;
        for (long long it = blockIdx.x; it < ntiles * nblk; it += gridDim.x) {
            const long long tile = it / nblk;
            const int t0 = (int)(it % nblk) * a.s2_chunk;
            const int tl = a.s2_tlive[tile];
            const int t1 = (t0 + a.s2_chunk < tl) ? t0 + a.s2_chunk : tl;
            if (t0 >= t1) continue;                                     // (the same decision in every thread of the CTA)
            const float* trj = a.scratch + (size_t)tile * a.scratch_per_cta;
            if (is_ctrl) {
                if (first) {                                            // one cyclic schedule serves every item
                    ctrl_flush(C);
                    sched_add_fwd(C, nG, imgG, nG.L - 1);
                    sched_add_bwd(C, nG, imgG, false);
                    ctrl_sched_ready(C);
                }
                for (int t = t0; t < t1; ++t) {
                    ctrl_net_forward(C, nG, nG.L - 1);
                    ctrl_net_backward(C, nG, true, copies, true, false);
                }
            } else if (is_help) {
                Masks mk;
                for (int t = t0; t < t1; ++t) {
                    help_forward_keep(P, nG, S.vecG, mk, copies, S.act, row, true);
                    help_backward(P, nG, gG, mk, true, gsG, S.dz, row, dexp);
                }
            } else {
                float xt[DPX], cot[DPX];
                Masks mkc;                                              // (combined mode only)
                const HelpArgs hg = {&mkc, copies, S.act, &gG, gsG};
                const float rg = a.s2_rhog[tile * TC_PATHS + row];
                float xn[DPX], cn[DPX];
                {
                    const float* tr = trj + (size_t)t0 * 2 * sr * TC_PATHS;
                    KLOOP(k, d) { xn[k] = __ldcs(&tr[k * TC_PATHS + row]); cn[k] = __ldcs(&tr[(sr + k) * TC_PATHS + row]); }
                }
                for (int t = t0; t < t1; ++t) {
                    KLOOP(k, d) { xt[k] = xn[k]; cot[k] = cn[k] * rg; }
                    float none[1];
                    own_net_forward_keep(P, nG, S.vecG, xt, none, copies, row, true, hg);
                    if (t + 1 < t1) {
                        const float* tr = trj + (size_t)(t + 1) * 2 * sr * TC_PATHS;
                        KLOOP(k, d) { xn[k] = __ldcs(&tr[k * TC_PATHS + row]); cn[k] = __ldcs(&tr[(sr + k) * TC_PATHS + row]); }
                    }
                    own_net_backward_nody0(P, nG, cot, S.dz, row, mxbuf, dexp, true, hg);
                }
            }
            first = false;
        }
        if (is_ctrl) { ctrl_flush(C); C.pc->quit = 1; }
        return;
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile = tc_next_tile(S, a)) {
        const long long base = tile * TC_PATHS;
        const long long slot = base + row;                        // position of this thread's path in the tiling
        const bool valid = is_own && slot < a.B_local;
        const long long gp = (valid && a.perm) ? (long long)a.perm[slot] : slot;      // local path index of this thread
        float x[DPX], u[DPX], dwv[DPX], sdw[DPX], g[DPX], raw[DPX];
        int flag = 0, nacc = 0;
        float disc = 1.f, y = 0.f;
        if (is_own) {
            KLOOP(k, d) x[k] = valid ? a.x0[gp * d + k] : fill;
            flag = fwd_initial_flag<float, DP, EQN, MV>(E, x, 1, 0);
            if (a.o_x && valid)
                KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1)] = x[k];
        }
        if (is_ctrl) {                                            // schedule of the rollout
            ctrl_flush(C);
            if (!cheat) sched_add_fwd(C, nA, imgA, nA.L);
            if (td1) sched_add_fwd(C, nG, imgG, nG.L);
            ctrl_sched_ready(C);
        }
        // ------------------------------------------------------------------ sweep 1: rollout
        // One CTA barrier per step decides whether any path of the tile is still inside.  The control warp and the helpers
        // run it at the top of the step; the owners run the one of step t+1 INSIDE step t, as soon as they have moved the
        // state: the networks of a step are one serial chain (the tile is the only one this SM works on), so what the owners
        // do between the end of one network and the published input of the next is exposed.  They therefore prepare the
        // next input (BatchNorm, bf16 split) while the current network still runs, and after its last product only read
        // its output, store the prepared planes and publish; the rest of the step's arithmetic follows under the next network.
        int tlive = 0;
        if (!is_own) {
            for (int t = 0; t < N; ++t) {
                const int alive = bar_work_or(false, TC_WORK_THREADS);
                if (!alive) break;
                tlive = t + 1;
                if (is_ctrl) {
                    if (!cheat) ctrl_net_forward(C, nA, nA.L);
                    if (td1) ctrl_net_forward(C, nG, nG.L);
                } else if (is_help) {
                    if (!cheat) help_forward(P, nA, S.vecA);
                    if (td1) help_forward(P, nG, S.vecG);
                }
            }
        } else {
            const bool hasA = !cheat, hasG = td1;
            const TcNet& n1 = hasA ? nA : nG;                             // first network of a step
            const float* vec1 = hasA ? S.vecA : S.vecG;
            uint32_t yh[2][8], yl[2][8];
            int alive = bar_work_or(valid && flag > 0, TC_WORK_THREADS);
            if (alive && (hasA || hasG)) {
                own_prep_y0(n1, vec1, x, yh, yl);
                own_store_y0(P, n1, yh, yl);
                own_publish(P);
            }
            for (int t = 0; t < N && alive; ++t) {
                tlive = t + 1;
                if (hasA && hasG) own_prep_y0(nG, S.vecG, x, yh, yl);            // NN_value_grad at x_t (before the move)
                path_dw(a, gp, valid, t, dwv);
                float dt, sqdt, xn; int dtg;
                fwd_dt<float, DP, EQN, MV>(E, x, flag, 1, 0, dt, sqdt, xn, dtg);
                if (cheat) {
                    eq_u_true<float, DP, EQN, MV>(E, x, u, 1, 0);
                } else {
                    if (TC_COMBINED) help_forward(P, nA, S.vecA);
                    own_last(P, nA, S.vecA, raw);
                    if (hasG) { own_store_y0(P, nG, yh, yl); own_publish(P); }
                    if (nA.ekn_head) ekn_head_fwd<float, DP, EQN, MV>(raw, u, nA.mctrl, 1, 0);
                    else KLOOP(j, E.m) u[j] = raw[j];
                }
                float* tr = traj + (size_t)t * 2 * sr * TC_PATHS;
                if (need_grad && td1)
                    KLOOP(k, d) __stcs(&tr[k * TC_PATHS + row], x[k]);
                float w = 0.f;
                if (!prop_only) w = eq_w<float, DP, EQN, MV>(E, x, u, 1, 0);
                const int coef = fwd_move<float, DP, EQN, MV>(E, x, u, dwv, dt, sqdt, xn, flag, sdw, 1, 0);
                const float cf = (float)coef;
                if (td1 && TC_COMBINED) help_forward(P, nG, S.vecG);              // (before the barrier: the control warp reaches it
                                                                                  //  only after it has issued the whole network)
                int alive_next = 0;
                if (t + 1 < N) alive_next = bar_work_or(valid && flag > 0, TC_WORK_THREADS);
                const bool pub = alive_next && (hasA || hasG);
                if (pub) own_prep_y0(n1, vec1, x, yh, yl);                        // input of the next step's first network
                if (td1) own_last(P, nG, S.vecG, g);
                if (pub) { own_store_y0(P, n1, yh, yl); own_publish(P); }
                y = y + w * disc * cf * dt;                                       // solver.py:170-174
                if (td1) {
                    float dif = 0.f;
                    KLOOP(k, d) dif = dif + sdw[k] * g[k];        // solver.py:177-182
                    dif = dif * disc;
                    y = y - dif * cf * sqdt;                                      // solver.py:184
                    if (need_grad) {
                        const float q = disc * cf * sqdt;
                        KLOOP(k, d) __stcs(&tr[(sr + k) * TC_PATHS + row], sdw[k] * q);
                    }
                }
                disc = disc * expf(-E.gamma * dt * cf);                          // solver.py:187
                nacc += coef;
                if (valid) {
                    if (a.o_dt) a.o_dt[gp * N + t] = dt;
                    if (a.o_coef) a.o_coef[gp * N + t] = cf;
                    if (a.o_x)
                        KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
                }
                alive = alive_next;
            }
        }
        if (valid) {
            for (int t = tlive; t < N; ++t) {
                if (a.o_dt) a.o_dt[gp * N + t] = E.delta_t;
                if (a.o_coef) a.o_coef[gp * N + t] = 0.f;
                if (a.o_x)
                    KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
            }
            if (a.o_exit) a.o_exit[gp] = nacc;
        }
        if (prop_only) continue;
        // ------------------------------------------------------------------ NN_value at x_0, x_N, x_bdry
        float rho_v = 0.f, rho_b = 0.f, rhog = 0.f;
        if (is_ctrl) {
            ctrl_flush(C);
            if (!need_grad) {
                sched_add_fwd(C, nV, imgV, nV.L);
                ctrl_sched_ready(C);
                for (int i = 0; i < 3; ++i) ctrl_net_forward(C, nV, nV.L);
            } else {
                sched_add_fwd(C, nV, imgV, nV.L);                       // V(x_0), forward only
                for (int i = 0; i < 3; ++i) { sched_add_fwd(C, nV, imgV, nV.L); sched_add_bwd(C, nV, imgV, false); }
                ctrl_sched_ready(C);
                ctrl_net_forward(C, nV, nV.L);
                for (int i = 0; i < 3; ++i) { ctrl_net_forward(C, nV, nV.L); ctrl_net_backward(C, nV, true, copies, false, false); }
            }
        } else if (is_help) {
            if (!need_grad) {
                for (int i = 0; i < 3; ++i) help_forward(P, nV, S.vecV);
            } else {
                Masks mk;
                help_forward(P, nV, S.vecV);
                for (int i = 0; i < 3; ++i) {
                    help_forward_keep(P, nV, S.vecV, mk, copies, S.act, row, false);
                    help_backward(P, nV, gV, mk, true, gsV, S.dz, row, dexp);
                }
            }
        } else {
            float vN[1], v0[1], vb[1], x0v[DPX], xbv[DPX], cot[1];
            KLOOP(k, d) x0v[k] = valid ? a.x0[gp * d + k] : fill;
            KLOOP(k, d) xbv[k] = valid ? a.xb[gp * d + k] : fill;
            if (!need_grad) {
                own_net_forward(P, nV, S.vecV, x0v, v0);
                own_net_forward(P, nV, S.vecV, x, vN);
                own_net_forward(P, nV, S.vecV, xbv, vb);
            } else {
                Masks mkc;                                                        // (combined mode only)
                const HelpArgs hv = {&mkc, copies, S.act, &gV, gsV};
                own_net_forward(P, nV, S.vecV, x0v, v0);
                own_net_forward_keep(P, nV, S.vecV, x, vN, copies, row, false, hv);
                const float delta = v0[0] - y - vN[0] * disc;
                rhog = valid ? rho_grad(delta, 50.f) * scale : 0.f;
                cot[0] = -rhog * disc;
                own_net_backward_nody0(P, nV, cot, S.dz, row, mxbuf, dexp, false, hv);
                own_net_forward_keep(P, nV, S.vecV, x0v, v0, copies, row, false, hv);
                cot[0] = rhog;
                own_net_backward_nody0(P, nV, cot, S.dz, row, mxbuf, dexp, false, hv);
                own_net_forward_keep(P, nV, S.vecV, xbv, vb, copies, row, false, hv);
                const float dbb = vb[0] - eq_Z<float, DP, EQN, MV>(E, xbv, 1, 0);
                cot[0] = valid ? rho_grad(dbb, 50.f) * scale : 0.f;
                own_net_backward_nody0(P, nV, cot, S.dz, row, mxbuf, dexp, false, hv);
            }
            const float delta = v0[0] - y - vN[0] * disc;                         // solver.py:189
            const float db = vb[0] - eq_Z<float, DP, EQN, MV>(E, xbv, 1, 0);                          // solver.py:190
            if (valid) {
                rho_v = rho(delta, 50.f);
                rho_b = rho(db, 50.f);
                if (a.o_delta) a.o_delta[gp] = delta;
                if (a.o_delta_b) a.o_delta_b[gp] = db;
            }
        }
        loss0 += tc_block_sum(rho_v, S.red);
        loss1 += tc_block_sum(rho_b, S.red);
        // ------------------------------------------------------------------ sweep 2: NN_value_grad backward
        if (need_grad && td1 && (a.flags & F_S2_DEFER)) {             // small batch: left to the second launch (see F_S2_ONLY)
            if (is_own) a.s2_rhog[tile * TC_PATHS + row] = rhog;
            if (tid == 0) a.s2_tlive[tile] = tlive;
        } else if (need_grad && td1) {
            if (is_ctrl) {
                ctrl_flush(C);
                sched_add_fwd(C, nG, imgG, nG.L - 1);
                sched_add_bwd(C, nG, imgG, false);
                ctrl_sched_ready(C);
                for (int t = 0; t < tlive; ++t) {
                    ctrl_net_forward(C, nG, nG.L - 1);
                    ctrl_net_backward(C, nG, true, copies, true, false);
                }
            } else if (is_help) {
                Masks mk;
                for (int t = 0; t < tlive; ++t) {
                    help_forward_keep(P, nG, S.vecG, mk, copies, S.act, row, true);
                    help_backward(P, nG, gG, mk, true, gsG, S.dz, row, dexp);
                }
            } else {
                float xt[DPX], cot[DPX];
                Masks mkc;                                                        // (combined mode only)
                const HelpArgs hg = {&mkc, copies, S.act, &gG, gsG};
                // (the trajectory of step t+1 is fetched while the backward pass of step t runs)
                float xn[DPX], cn[DPX];
                if (tlive > 0) KLOOP(k, d) { xn[k] = __ldcs(&traj[k * TC_PATHS + row]); cn[k] = __ldcs(&traj[(sr + k) * TC_PATHS + row]); }
                for (int t = 0; t < tlive; ++t) {
                    KLOOP(k, d) { xt[k] = xn[k]; cot[k] = cn[k] * rhog; }
                    float none[1];
                    own_net_forward_keep(P, nG, S.vecG, xt, none, copies, row, true, hg);
                    if (t + 1 < tlive) {
                        const float* tr = traj + (size_t)(t + 1) * 2 * sr * TC_PATHS;
                        KLOOP(k, d) { xn[k] = __ldcs(&tr[k * TC_PATHS + row]); cn[k] = __ldcs(&tr[(sr + k) * TC_PATHS + row]); }
                    }
                    own_net_backward_nody0(P, nG, cot, S.dz, row, mxbuf, dexp, true, hg);
                }
            }
        }
    }
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
        if (warp >= TC_CTRL_WARP) {                      /* control, producer (lane 0 streams the weights) */                  \
            if (warp == TC_PROD_WARP) {                                                                                       \
                if ((threadIdx.x & 31) == 0) producer_loop(S.ring, S.full, S.empty, S.sch, S.pc, a.nslot, a.slot_bytes);      \
            } else {                                                                                                          \
                NAME##_tc_body<DP, EQN, MV, ROLE_CTRL>(a, S, R);                                                              \
            }                                                                                                                 \
        } else if (warp >= 4) {                                                                                               \
            NAME##_tc_body<DP, EQN, MV, ROLE_HELP>(a, S, R);                                                                  \
        } else {                                                                                                              \
            NAME##_tc_body<DP, EQN, MV, ROLE_OWN>(a, S, R);                                                                   \
        }                                                                                                                     \
        roles_stats(R, a, t_start);                                                                                           \
        tc_fence_before();                                                                                                    \
        __syncthreads();                                                                                                      \
        if (warp == TC_CTRL_WARP) tmem_dealloc(tmem, 512);                                                                    \
    }

// =================================================================================== actor (tensor)
template <int DP, int EQN, int MV, int ROLE>
__device__ __forceinline__ void actor_tc_body(const TcArgs& a, const TcSmem& S, Roles& R) {
    constexpr int DPX = DP > 0 ? DP : 32;
    constexpr bool is_own = ROLE == ROLE_OWN, is_help = ROLE == ROLE_HELP, is_ctrl = ROLE == ROLE_CTRL;
    Ctrl& C = R.C;
    PathCtx& P = R.P;
    const TcNet& nA = *S.nA;
    const TcNet& nV = *S.nV;
    const TcSlab& gA = *S.gA;
    const TcSlab& gV = *S.gV;
    (void)gA; (void)gV;
    const int rep = (int)(blockIdx.x % (unsigned)a.img_rep);
    const unsigned char* imgA = a.imgA + rep * a.img_strideA;
    const unsigned char* imgV = a.imgV + rep * a.img_strideV;
    (void)imgA; (void)imgV;
    const int tid = threadIdx.x;
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
    uint32_t* mxbuf = S.dzmax;
    volatile int* dexp = reinterpret_cast<volatile int*>(S.dzmax + 16);
    (void)traj; (void)copies; (void)gsA; (void)mxbuf; (void)dexp; (void)fill; (void)m; (void)row; (void)trs;

    float loss0 = 0.f;
    const long long ntiles = (a.B_local + TC_PATHS - 1) / TC_PATHS;
    for (long long tile = blockIdx.x; tile < ntiles; tile = tc_next_tile(S, a)) {
        const long long base = tile * TC_PATHS;
        const long long slot = base + row;
        const bool valid = is_own && slot < a.B_local;
        const long long gp = (valid && a.perm) ? (long long)a.perm[slot] : slot;
        float x[DPX], u[DPX], dwv[DPX], raw[DPX];
        int flag = 0, nacc = 0;
        float disc = 1.f, y = 0.f;
        if (is_own) {
            KLOOP(k, d) x[k] = valid ? a.x0[gp * d + k] : fill;
            flag = fwd_initial_flag<float, DP, EQN, MV>(E, x, 1, 0);
            if (a.o_x && valid)
                KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1)] = x[k];
        }
        if (is_ctrl) {
            ctrl_flush(C);
            if (!cheat) sched_add_fwd(C, nA, imgA, nA.L);
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
            } else if (is_help) {
                if (!cheat) help_forward(P, nA, S.vecA);
            } else {
                if (!cheat) own_put_y0(P, nA, S.vecA, x, nullptr, 0);
                if (TC_COMBINED && !cheat) help_forward(P, nA, S.vecA, 0, 1);          // (the first layer's product is short)
                path_dw(a, gp, valid, t, dwv);
                if (TC_COMBINED && !cheat) help_forward(P, nA, S.vecA, 1, 2);
                float dt, sqdt, xn; int dtg;
                fwd_dt<float, DP, EQN, MV>(E, x, flag, 1, 0, dt, sqdt, xn, dtg);
                float* tr = traj + (size_t)t * trs * TC_PATHS;
                if (need_grad)
                    KLOOP(k, d) { __stcs(&tr[k * TC_PATHS + row], x[k]); __stcs(&tr[(sr + k) * TC_PATHS + row], dwv[k]); }
                if (cheat) {
                    eq_u_true<float, DP, EQN, MV>(E, x, u, 1, 0);
                } else {
                    if (TC_COMBINED) help_forward(P, nA, S.vecA, 2);
                    own_last(P, nA, S.vecA, raw);
                    if (nA.ekn_head) ekn_head_fwd<float, DP, EQN, MV>(raw, u, nA.mctrl, 1, 0);
                    else KLOOP(j, m) u[j] = raw[j];
                }
                const float w = eq_w<float, DP, EQN, MV>(E, x, u, 1, 0);
                const int coef = fwd_move<float, DP, EQN, MV>(E, x, u, dwv, dt, sqdt, xn, flag, (float*)nullptr, 1, 0);
                const float cf = (float)coef;
                if (need_grad) {
                    float* sc = tr + (size_t)2 * sr * TC_PATHS;
                    sc[A_DT * TC_PATHS + row] = dt; sc[A_SQDT * TC_PATHS + row] = sqdt; sc[A_COEF * TC_PATHS + row] = valid ? cf : 0.f;
                    sc[A_DISC * TC_PATHS + row] = disc; sc[A_XN * TC_PATHS + row] = xn; sc[A_DTG * TC_PATHS + row] = (float)dtg;
                }
                y = y + cf * w * dt * disc;                                       // solver.py:218
                disc = disc * expf(-E.gamma * dt * cf);                          // solver.py:219
                nacc += coef;
                if (valid) {
                    if (a.o_dt) a.o_dt[gp * N + t] = dt;
                    if (a.o_coef) a.o_coef[gp * N + t] = cf;
                    if (a.o_x)
                        KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
                }
            }
        }
        if (valid) {
            for (int t = tlive; t < N; ++t) {
                if (a.o_dt) a.o_dt[gp * N + t] = E.delta_t;
                if (a.o_coef) a.o_coef[gp * N + t] = 0.f;
                if (a.o_x)
                    KLOOP(k, d) a.o_x[(gp * d + k) * (long long)(N + 1) + t + 1] = x[k];
            }
            if (a.o_exit) a.o_exit[gp] = nacc;
        }
        // ------------------------------------------------------------------ terminal value (+ its input gradient)
        float yv = 0.f;
        float lam[DPX];
        float Dbar = 0.f;
        if (is_ctrl) {
            ctrl_flush(C);
            if (!cheat_v) {
                sched_add_fwd(C, nV, imgV, nV.L);
                if (need_grad) sched_add_bwd(C, nV, imgV);
                ctrl_sched_ready(C);
                ctrl_net_forward(C, nV, nV.L);
                if (need_grad) ctrl_net_backward(C, nV, false, nullptr, false, true);
            }
        } else if (is_help) {
            if (!cheat_v) {
                if (!need_grad) {
                    help_forward(P, nV, S.vecV);
                } else {
                    Masks mk;
                    help_forward_keep(P, nV, S.vecV, mk, nullptr, nullptr, row, false);
                    help_backward(P, nV, gV, mk, false, nullptr, nullptr, row, dexp);
                }
            }
        } else {
            float vN[1];
            const float seed = valid ? disc * a.invB : 0.f;
            if (cheat_v) {
                vN[0] = eq_V_true<float, DP, EQN, MV>(E, x, 1, 0);                                    // solver.py:223
                if (need_grad) {
                    eq_V_grad_true<float, DP, EQN, MV>(E, x, lam, 1, 0);
                    KLOOP(k, d) lam[k] = lam[k] * seed;
                }
            } else if (!need_grad) {
                own_net_forward(P, nV, S.vecV, x, vN);                          // solver.py:221
            } else {
                float cot[1], dy0[DPX];
                Masks mkc;                                                        // (combined mode only)
                const HelpArgs hv = {&mkc, nullptr, nullptr, &gV, nullptr};
                own_net_forward_keep(P, nV, S.vecV, x, vN, nullptr, row, false, hv);
                cot[0] = seed;
                own_net_backward(P, nV, cot, false, nullptr, row, dy0, mxbuf, dexp, false, hv);
                const float* g0c = S.vecV + nV.vec_g0;
                KLOOP(k, d) lam[k] = dy0[k] * g0c[k];
            }
            Dbar = valid ? vN[0] * a.invB : 0.f;
            y = y + vN[0] * disc;
            if (valid) {
                yv = y;
                if (a.o_delta) a.o_delta[gp] = y;
            }
        }
        loss0 += tc_block_sum(yv, S.red);
        if (!need_grad) continue;
        // ------------------------------------------------------------------ reverse sweep
        if (is_ctrl) {
            ctrl_flush(C);
            sched_add_fwd(C, nA, imgA, nA.L);
            sched_add_bwd(C, nA, imgA);
            ctrl_sched_ready(C);
        }
        for (int t = tlive - 1; t >= 0; --t) {
            const float* tr = traj + (size_t)t * trs * TC_PATHS;
            const float* sc = tr + (size_t)2 * sr * TC_PATHS;
            if (is_own && t > 0) {                                       // the record of the step before: on its way to L2 by the time it is read
                const char* nb = reinterpret_cast<const char*>(tr - (size_t)trs * TC_PATHS);
                for (int i = row; i < trs * 4; i += TC_OWN_THREADS) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + (size_t)i * 128));
            }
            const int any = bar_work_or(valid && sc[A_COEF * TC_PATHS + row] > 0.f, TC_WORK_THREADS);
            if (!any) continue;
            if (is_ctrl) {
                ctrl_net_forward(C, nA, nA.L);
                ctrl_net_backward(C, nA, true, copies, false, true);
            } else if (is_help) {
                Masks mk;
                help_forward_keep(P, nA, S.vecA, mk, copies, S.act, row, false);
                help_backward(P, nA, gA, mk, true, gsA, S.dz, row, dexp);
            } else {
                float xt[DPX], ubar[DPX], cot[DPX], dy0[DPX];
                KLOOP(k, d) { xt[k] = __ldcs(&tr[k * TC_PATHS + row]); dwv[k] = __ldcs(&tr[(sr + k) * TC_PATHS + row]); }
                Masks mkc;                                                        // (combined mode only)
                const HelpArgs ha = {&mkc, copies, S.act, &gA, gsA};
                own_net_forward_keep(P, nA, S.vecA, xt, raw, copies, row, false, ha);
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
                own_net_backward(P, nA, cot, true, S.dz, row, dy0, mxbuf, dexp, false, ha);
                const float* g0c = S.vecA + nA.vec_g0;
                KLOOP(k, d) lam[k] = lam[k] + dy0[k] * g0c[k];
            }
        }
    }
    if (is_ctrl) { ctrl_flush(C); C.pc->quit = 1; }
    if (tid == 0 && a.loss_part) {
        a.loss_part[blockIdx.x * 2] = loss0;
        a.loss_part[blockIdx.x * 2 + 1] = 0.f;
    }
}

DPB_TC_KERNEL(critic)
DPB_TC_KERNEL(actor)

}  // namespace tc
}  // namespace dpb
