// dpb_api.cu -- C ABI of libdeeppde_b200.so (include/deeppde_b200.h): handle, workspace layout and
// the launch sequences of the exact path.  No torch types, no allocation per call, no stream sync
// (except the *_host convenience entry points, which must deliver host-visible losses).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include "../../include/deeppde_b200.h"
#include "dpb_host.h"
#include "dpb_kernels.cuh"
#include "dpb_tc_selftest.cuh"
#include "dpb_tc_kernels.cuh"

using namespace dpb;

struct dpb_handle {
    dpb_config cfg;
    NetDev nA, nV, nG;
    tc::TcNet tA, tV, tG;
    tc::TcSlab sA, sV, sG;
    int tc_maxw16;                  // widest padded layer extent of the three networks
    int num_sms;
    int img_rep;                    // copies of every operand image of the tensor path (experiment knob DPB_TC_IMG_REP, default 1)
    int max_smem;
    int sr, hrows, nhb;
    long long launches;
    cudaEvent_t ev0, ev1;
    bool have_ev, timed;
    bool timing;                    // record CUDA events around the rollout kernels (off while a CUDA graph is captured)
    std::string err;
};

static void ev_begin(dpb_handle* h, cudaStream_t st) {
    if (!h->timing) return;
    if (!h->have_ev) {
        if (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) { cudaGetLastError(); return; }
        h->have_ev = true;
    }
    cudaEventRecord(h->ev0, st);
}
static void ev_end(dpb_handle* h, cudaStream_t st) {
    if (h->timing && h->have_ev) { cudaEventRecord(h->ev1, st); h->timed = true; }
}

static thread_local std::string g_err;

static int fail(dpb_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    g_err = msg;
    return code;
}

#define DPB_CUDA(h, call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return fail(h, DPB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

static inline size_t esize(const dpb_handle* h) { return h->cfg.dtype == DPB_F64 ? 8 : 4; }
static inline size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int tileP(const dpb_handle* h) { return h->cfg.dtype == DPB_F64 ? 8 * RT<double>::TP : 8 * RT<float>::TP; }
static const double BN_C = 1.0 / sqrt(1.0 + 1e-6);       // solver.py:242 (epsilon), moving var = 1

struct Layout {
    int grid;
    int nslab;              // gradient slabs: one per CTA on the exact path (deterministic), <= 16 shared ones on the tensor path
    size_t pkA, pkV, pkG, loss_part, loss_out, scratch, slabs, raw, total;
    size_t imgA, imgV, imgG, vecA, vecV, vecG, copies, stats, trace;      // tensor path: operand images, vector blocks, activation copies, counters
    size_t life_nacc, life_perm, life_hist;                        // tensor path, naive scheme: lifetime sort (exit counts, permutation, bins)
    size_t s2_rhog, s2_tlive;                                      // tensor path, small batches: hand-over to the second-sweep launch
    long long scratch_per_cta, copies_per_cta;
};

static Layout make_layout(const dpb_handle* h, long long B_local, int N) {
    Layout L;
    const size_t es = esize(h);
    const bool tensor = h->cfg.impl == DPB_IMPL_TENSOR;
    const int P = tensor ? tc::TC_PATHS : tileP(h);
    long long ntiles = (B_local + P - 1) / P;
    L.grid = (int)(ntiles < h->num_sms ? (ntiles < 1 ? 1 : ntiles) : h->num_sms);
    L.nslab = tensor ? (L.grid < 16 ? L.grid : 16) : L.grid;
    size_t o = 0;
    L.imgA = L.imgV = L.imgG = L.vecA = L.vecV = L.vecG = 0;
    if (tensor) {
        // img_rep copies of every operand image (experiment: all CTAs stream the same few hundred KB of weights, roughly in
        // step -- would one copy keep a handful of L2 slices busy while the others idle?  Measured: 1, 4, 8, 16 copies run
        // the same 2^17-path step to 0.1 %, so the weight stream is not bound by L2 slices; the default stays 1)
        L.imgA = o; o += (size_t)h->img_rep * a256(h->tA.img_bytes);
        L.imgV = o; o += (size_t)h->img_rep * a256(h->tV.img_bytes);
        L.imgG = o; o += (size_t)h->img_rep * a256(h->tG.img_bytes);
        L.vecA = o; o += a256((size_t)h->tA.vec_floats * 4);
        L.vecV = o; o += a256((size_t)h->tV.vec_floats * 4);
        L.vecG = o; o += a256((size_t)h->tG.vec_floats * 4);
    }
    L.pkA = o; o += a256(h->nA.ptotal * es);
    L.pkV = o; o += a256(h->nV.ptotal * es);
    L.pkG = o; o += a256(h->nG.ptotal * es);
    L.loss_part = o; o += a256((size_t)L.grid * 2 * es);
    L.loss_out = o; o += 256;
    L.scratch_per_cta = (long long)N * (2 * h->sr + A_NSCAL) * P;
    L.scratch = o; o += a256((size_t)L.grid * L.scratch_per_cta * es);
    L.copies = 0; L.copies_per_cta = 0; L.stats = 0; L.trace = 0; L.life_nacc = L.life_perm = L.life_hist = 0; L.s2_rhog = L.s2_tlive = 0;
    if (tensor) {
        long long cb = tc::tc_copy_bytes(h->tA);
        if (tc::tc_copy_bytes(h->tV) > cb) cb = tc::tc_copy_bytes(h->tV);
        if (tc::tc_copy_bytes(h->tG) > cb) cb = tc::tc_copy_bytes(h->tG);
        L.copies_per_cta = (long long)a256((size_t)cb);
        // (sized for every SM: the second-sweep launch of a small batch runs more CTAs than the batch has tiles)
        L.copies = o; o += a256((size_t)h->num_sms * L.copies_per_cta);
        L.stats = o; o += a256((size_t)h->num_sms * 16 * 8 + 64);       // + the tile counter of the dynamic tile scheduler
        L.trace = o; o += a256((size_t)3 * tc::TC_TRACE_CAP * 8);        // event trace of CTA 0 (written by stats builds only)
        L.life_nacc = o; o += a256((size_t)B_local * 4);
        L.life_perm = o; o += a256((size_t)B_local * 4);
        L.life_hist = o; o += a256((size_t)(N + 2) * 4);
        L.s2_rhog = o; o += a256((size_t)ntiles * tc::TC_PATHS * 4);
        L.s2_tlive = o; o += a256((size_t)ntiles * 4);
    }
    const long long gc = tensor ? h->sV.gtotal + h->sG.gtotal : h->nV.gtotal + h->nG.gtotal, ga = tensor ? h->sA.gtotal : h->nA.gtotal;
    const long long gmax = gc > ga ? gc : ga;
    L.slabs = o; o += a256((size_t)L.nslab * gmax * es);
    L.raw = o; o += a256((size_t)gmax * es);
    L.total = o;
    return L;
}

extern "C" {

const char* dpb_version(void) { return "deeppde_b200 0.2 (sm_100a: tcgen05 tensor path + exact CUDA-core path)"; }

const char* dpb_last_error(const dpb_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int dpb_create(dpb_handle** out, const dpb_config* cfg) {
    if (!out || !cfg) return fail(nullptr, DPB_ERR_ARG, "dpb_create: null argument");
    *out = nullptr;
    const dpb_config& c = *cfg;
    if (c.dtype != DPB_F32 && c.dtype != DPB_F64) return fail(nullptr, DPB_ERR_ARG, "dtype must be DPB_F32 or DPB_F64");
    if (c.eqn < DPB_EQN_LQR || c.eqn > DPB_EQN_LQR_VAR) return fail(nullptr, DPB_ERR_ARG, "unknown equation id");
    if (c.dim < 1 || c.dim > DPB_MAX_DIM - 1 || c.control_dim < 1 || c.control_dim > DPB_MAX_DIM - 1)
        return fail(nullptr, DPB_ERR_ARG, "dim / control_dim out of range (1..31)");
    if (c.scheme != DPB_SCHEME_NAIVE && c.scheme != DPB_SCHEME_ADAPTIVE) return fail(nullptr, DPB_ERR_ARG, "unknown scheme");
    if (c.td_type != DPB_TD1 && c.td_type != DPB_TD2) return fail(nullptr, DPB_ERR_ARG, "TD_type must be 1 or 2");
    if (c.eqn == DPB_EQN_VDP && c.dim != 2 * c.control_dim) return fail(nullptr, DPB_ERR_ARG, "VDP needs dim == 2*control_dim (equation.py:186-187)");
    if ((c.eqn == DPB_EQN_LQR || c.eqn == DPB_EQN_LQR_VAR || c.eqn == DPB_EQN_EKN) && c.dim != c.control_dim)
        return fail(nullptr, DPB_ERR_ARG, "LQR / LQR_var / ekn need control_dim == dim (equation.py:164,261,305)");
    if (c.n_hidden_actor < 1 || c.n_hidden_actor > DPB_MAX_HIDDEN || c.n_hidden_critic < 1 || c.n_hidden_critic > DPB_MAX_HIDDEN)
        return fail(nullptr, DPB_ERR_ARG, "1..6 hidden layers per network");
    int hmax = 0;
    for (int i = 0; i < c.n_hidden_actor; ++i) { if (c.hidden_actor[i] < 1 || c.hidden_actor[i] > WS_NMAX) return fail(nullptr, DPB_ERR_ARG, "hidden width must be 1..256"); hmax = hmax > c.hidden_actor[i] ? hmax : c.hidden_actor[i]; }
    for (int i = 0; i < c.n_hidden_critic; ++i) { if (c.hidden_critic[i] < 1 || c.hidden_critic[i] > WS_NMAX) return fail(nullptr, DPB_ERR_ARG, "hidden width must be 1..256"); hmax = hmax > c.hidden_critic[i] ? hmax : c.hidden_critic[i]; }
    if (!(c.R > 0)) return fail(nullptr, DPB_ERR_ARG, "R must be positive");

    dpb_handle* h = new dpb_handle();
    h->cfg = c;
    h->launches = 0;
    h->have_ev = false;
    h->timed = false;
    h->timing = true;
    const int ekn = (c.eqn == DPB_EQN_EKN);
    netdev_init(h->nA, c.dim, c.hidden_actor, c.n_hidden_actor, ekn ? c.control_dim + 1 : c.control_dim, ekn, c.control_dim);   // solver.py:255-258
    netdev_init(h->nV, c.dim, c.hidden_critic, c.n_hidden_critic, 1, 0, 0);                                                  // solver.py:251-252
    netdev_init(h->nG, c.dim, c.hidden_critic, c.n_hidden_critic, c.dim, 0, 0);                                              // solver.py:253-254
    tc::tcnet_init(h->tA, c.dim, c.hidden_actor, c.n_hidden_actor, ekn ? c.control_dim + 1 : c.control_dim, ekn, c.control_dim);
    tc::tcnet_init(h->tV, c.dim, c.hidden_critic, c.n_hidden_critic, 1, 0, 0);
    tc::tcnet_init(h->tG, c.dim, c.hidden_critic, c.n_hidden_critic, c.dim, 0, 0);
    tc::tcslab_init(h->sA, h->tA); tc::tcslab_init(h->sV, h->tV); tc::tcslab_init(h->sG, h->tG);
    h->tc_maxw16 = 16;
    for (const tc::TcNet* t : {&h->tA, &h->tV, &h->tG})
        for (int l = 0; l <= t->L; ++l) {
            if (t->ly[l].K16 > h->tc_maxw16) h->tc_maxw16 = t->ly[l].K16;
            if (t->ly[l].N16 > h->tc_maxw16) h->tc_maxw16 = t->ly[l].N16;
        }
    if (c.impl != DPB_IMPL_EXACT && c.impl != DPB_IMPL_TENSOR) { delete h; return fail(nullptr, DPB_ERR_ARG, "impl must be DPB_IMPL_EXACT or DPB_IMPL_TENSOR"); }
    if (c.impl == DPB_IMPL_TENSOR) {
        if (c.dtype != DPB_F32) { delete h; return fail(nullptr, DPB_ERR_ARG, "the tensor path computes in float32 (bf16x3 products, FP32 accumulation): dtype must be DPB_F32"); }
        if (!tc::tcnet_supported(h->tA) || !tc::tcnet_supported(h->tV) || !tc::tcnet_supported(h->tG)) { delete h; return fail(nullptr, DPB_ERR_ARG, "tensor path: layer width must be <= 255"); }
        // longest load schedule of a phase: V forward + 3 x (V forward + backward), one entry per product (dpb_tc_kernels.cuh)
        const int lmx = c.n_hidden_actor > c.n_hidden_critic ? c.n_hidden_actor : c.n_hidden_critic;
        if (7 * (lmx + 1) > tc::MAXOPS) { delete h; return fail(nullptr, DPB_ERR_ARG, "tensor path: too many layers for the chunk schedule"); }
    }
    const int mx = c.dim > c.control_dim + 1 ? c.dim : c.control_dim + 1;
    h->sr = round8(mx);
    h->hrows = round8(hmax);
    const int lmax = c.n_hidden_actor > c.n_hidden_critic ? c.n_hidden_actor : c.n_hidden_critic;
    h->nhb = lmax + 2;
    h->num_sms = 0;
    h->max_smem = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&h->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    } else {
        cudaGetLastError();
    }
    if (h->num_sms <= 0) h->num_sms = 148;                 // layout only; compute calls fail without a device
    h->img_rep = 1;
    if (const char* e = getenv("DPB_TC_IMG_REP")) { const int v = atoi(e); if (v >= 1 && v <= 64) h->img_rep = v; }
    const size_t need = (c.dtype == DPB_F64 ? carve_elems<double>(h->sr, h->hrows, h->nhb) * 8 : carve_elems<float>(h->sr, h->hrows, h->nhb) * 4);
    if (need > 227 * 1024) {
        delete h;
        return fail(nullptr, DPB_ERR_ARG, "networks too large for the on-chip activation tile (need " + std::to_string(need) + " B of shared memory)");
    }
    *out = h;
    return DPB_OK;
}

int dpb_destroy(dpb_handle* h) {
    if (h && h->have_ev) { cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); }
    delete h;
    return DPB_OK;
}

int64_t dpb_param_count(const dpb_handle* h, int which) {
    if (!h) return -1;
    switch (which) {
    case DPB_NET_ACTOR: return h->nA.ftotal;
    case DPB_NET_CRITIC: return h->nV.ftotal;
    case DPB_NET_CRITIC_GRAD: return h->nG.ftotal;
    }
    return -1;
}

int64_t dpb_workspace_bytes(const dpb_handle* h, int64_t B_local, int32_t N) {
    if (!h || B_local < 1 || N < 1) return -1;
    return (int64_t)make_layout(h, B_local, N).total;
}

int64_t dpb_staging_bytes(const dpb_handle* h, int64_t B_local, int32_t N, int32_t dw_mode) {
    if (!h || B_local < 1 || N < 1) return -1;
    const size_t es = esize(h);
    size_t o = 2 * a256((size_t)B_local * h->cfg.dim * es);
    if (dw_mode == DPB_DW_EXTERNAL) o += a256((size_t)B_local * h->cfg.dim * N * es);
    return (int64_t)o;
}

int64_t dpb_launch_count(const dpb_handle* h) { return h ? h->launches : -1; }

int dpb_set_timing(dpb_handle* h, int enable) {
    if (!h) return fail(nullptr, DPB_ERR_ARG, "dpb_set_timing: null handle");
    h->timing = enable != 0;
    if (!h->timing) h->timed = false;
    return DPB_OK;
}

double dpb_last_kernel_ms(dpb_handle* h) {
    if (!h || !h->timed) return -1.0;
    float ms = -1.f;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess || cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    return (double)ms;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
template <typename real>
static int pack_net(dpb_handle* h, const NetDev& nd, const void* theta, void* pk, cudaStream_t st) {
    const long long work = nd.ptotal;
    int blocks = (int)((work + 255) / 256);
    if (blocks > 296) blocks = 296;
    pack_net_kernel<real><<<blocks, 256, 0, st>>>(nd, (const real*)theta, (real*)pk, (real)BN_C);
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

template <typename real>
static int finalize_net(dpb_handle* h, const NetDev& nd, const void* theta, const real* slabs, int nslab, real* raw, void* grad, cudaStream_t st) {
    int blocks = (int)((nd.gtotal + 255) / 256);
    if (blocks > 592) blocks = 592;
    reduce_slabs_kernel<real><<<blocks, 256, 0, st>>>(slabs, nslab, nd.gtotal, raw);
    finalize_grad_kernel<real><<<blocks, 256, 0, st>>>(nd, (const real*)theta, raw, (real*)grad, (real)BN_C);
    h->launches += 2;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

template <typename real>
static void fill_common(dpb_handle* h, StepArgs<real>& a, const Layout& L, char* ws, const dpb_inputs* in, int64_t B_local, int64_t path_offset,
                        int64_t B_global, int32_t N, double T, uint32_t flags, const dpb_path_outputs* outs) {
    memset(&a, 0, sizeof(a));
    fill_eqn(h->cfg, N, T, a.eq);
    a.nA = h->nA; a.nV = h->nV; a.nG = h->nG;
    a.pkA = (const real*)(ws + L.pkA); a.pkV = (const real*)(ws + L.pkV); a.pkG = (const real*)(ws + L.pkG);
    a.x0 = (const real*)in->x0; a.dw = (const real*)in->dw; a.xb = (const real*)in->x_bdry;
    a.dw_mode = in->dw_mode; a.seed = in->seed; a.stream = in->stream; a.stream_base = (const unsigned long long*)in->stream_base;
    a.B_local = B_local; a.path_offset = path_offset;
    a.invB = (real)(1.0 / (double)B_global);
    a.N = N; a.flags = flags;
    a.loss_part = (real*)(ws + L.loss_part);
    a.scratch = (real*)(ws + L.scratch);
    a.scratch_per_cta = L.scratch_per_cta;
    a.sr = h->sr; a.hrows = h->hrows; a.nhb = h->nhb;
    if (outs) {
        a.o_x = (real*)outs->x_smp; a.o_dt = (real*)outs->dt; a.o_coef = (real*)outs->coef;
        a.o_delta = (real*)outs->delta; a.o_delta_b = (real*)outs->delta_bdry; a.o_exit = outs->exit_index;
    }
}

static int check_step_args(dpb_handle* h, const dpb_inputs* in, int64_t B_local, int64_t B_global, int32_t N, double T,
                           void* ws, int64_t ws_bytes, const char* who) {
    if (!h) return fail(nullptr, DPB_ERR_ARG, std::string(who) + ": null handle");
    if (!in || !in->x0) return fail(h, DPB_ERR_ARG, std::string(who) + ": x0 is required");
    if (B_local < 1 || B_global < B_local || N < 1 || !(T > 0)) return fail(h, DPB_ERR_ARG, std::string(who) + ": bad B_local/B_global/N/T");
    if (in->dw_mode == DPB_DW_EXTERNAL && !in->dw) return fail(h, DPB_ERR_ARG, std::string(who) + ": dw is NULL but dw_mode is EXTERNAL");
    if (in->dw_mode < DPB_DW_EXTERNAL || in->dw_mode > DPB_DW_PHILOX_BOUNDED) return fail(h, DPB_ERR_ARG, std::string(who) + ": bad dw_mode");
    if (!ws) return fail(h, DPB_ERR_WORKSPACE, std::string(who) + ": workspace is NULL");
    const int64_t need = dpb_workspace_bytes(h, B_local, N);
    if (ws_bytes < need) return fail(h, DPB_ERR_WORKSPACE, std::string(who) + ": workspace too small, need " + std::to_string(need) + " bytes");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(h, DPB_ERR_CUDA, std::string(who) + ": no CUDA device (there is no CPU fallback)");
    }
    return DPB_OK;
}

template <typename real>
static int critic_step_t(dpb_handle* h, const void* thA, const void* thV, const void* thG, const dpb_inputs* in, int64_t B_local,
                         int64_t path_offset, int64_t B_global, int32_t N, double T, uint32_t flags, void* out_loss, void* grad_V,
                         void* grad_G, const dpb_path_outputs* outs, void* workspace, cudaStream_t st) {
    char* ws = (char*)workspace;
    const Layout L = make_layout(h, B_local, N);
    StepArgs<real> a;
    fill_common<real>(h, a, L, ws, in, B_local, path_offset, B_global, N, T, flags, outs);
    const bool cheat = flags & DPB_FLAG_CHEAT_CONTROL, need_grad = (flags & DPB_FLAG_NEED_GRAD) && !(flags & DPB_FLAG_PROPAGATE_ONLY);
    const bool prop_only = flags & DPB_FLAG_PROPAGATE_ONLY;
    const bool td1 = h->cfg.td_type == DPB_TD1;
    int rc;
    if (!cheat) { if ((rc = pack_net<real>(h, h->nA, thA, ws + L.pkA, st))) return rc; }
    if (!prop_only) {
        if ((rc = pack_net<real>(h, h->nV, thV, ws + L.pkV, st))) return rc;
        if (td1) { if ((rc = pack_net<real>(h, h->nG, thG, ws + L.pkG, st))) return rc; }
    }
    if (need_grad) {
        a.slabV = (real*)(ws + L.slabs);
        a.slabG = a.slabV + (size_t)L.grid * h->nV.gtotal;
        DPB_CUDA(h, cudaMemsetAsync(ws + L.slabs, 0, (size_t)L.grid * (h->nV.gtotal + h->nG.gtotal) * sizeof(real), st));
    }
    const size_t smem = carve_elems<real>(h->sr, h->hrows, h->nhb) * sizeof(real);
    DPB_CUDA(h, cudaFuncSetAttribute(critic_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ev_begin(h, st);
    critic_kernel<real><<<L.grid, NTHREADS, smem, st>>>(a);
    ev_end(h, st);
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    if (out_loss && !prop_only) {
        const real s = (real)(100.0 / (double)B_global);
        reduce_loss_kernel<real><<<1, 32, 0, st>>>(a.loss_part, L.grid, s, s, (real*)out_loss);
        h->launches++;
    }
    if (need_grad) {
        real* raw = (real*)(ws + L.raw);
        if (grad_V) { if ((rc = finalize_net<real>(h, h->nV, thV, a.slabV, L.grid, raw, grad_V, st))) return rc; }
        if (grad_G) {
            if (td1) { if ((rc = finalize_net<real>(h, h->nG, thG, a.slabG, L.grid, raw + h->nV.gtotal, grad_G, st))) return rc; }
            else DPB_CUDA(h, cudaMemsetAsync(grad_G, 0, (size_t)h->nG.ftotal * sizeof(real), st));      // Keras skips None grads
        }
    }
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

template <typename real>
static int actor_step_t(dpb_handle* h, const void* thA, const void* thV, const dpb_inputs* in, int64_t B_local, int64_t path_offset,
                        int64_t B_global, int32_t N, double T, uint32_t flags, void* out_loss, void* grad_A,
                        const dpb_path_outputs* outs, void* workspace, cudaStream_t st) {
    char* ws = (char*)workspace;
    const Layout L = make_layout(h, B_local, N);
    StepArgs<real> a;
    fill_common<real>(h, a, L, ws, in, B_local, path_offset, B_global, N, T, flags, outs);
    const bool cheat = flags & DPB_FLAG_CHEAT_CONTROL, cheat_v = flags & DPB_FLAG_CHEAT_VALUE;
    const bool need_grad = (flags & DPB_FLAG_NEED_GRAD) && !cheat;
    int rc;
    if (!cheat) { if ((rc = pack_net<real>(h, h->nA, thA, ws + L.pkA, st))) return rc; }
    if (!cheat_v) { if ((rc = pack_net<real>(h, h->nV, thV, ws + L.pkV, st))) return rc; }
    if (need_grad) {
        a.slabA = (real*)(ws + L.slabs);
        DPB_CUDA(h, cudaMemsetAsync(ws + L.slabs, 0, (size_t)L.grid * h->nA.gtotal * sizeof(real), st));
    }
    const size_t smem = carve_elems<real>(h->sr, h->hrows, h->nhb) * sizeof(real);
    DPB_CUDA(h, cudaFuncSetAttribute(actor_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ev_begin(h, st);
    actor_kernel<real><<<L.grid, NTHREADS, smem, st>>>(a);
    ev_end(h, st);
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    if (out_loss) {
        reduce_loss_kernel<real><<<1, 32, 0, st>>>(a.loss_part, L.grid, (real)(1.0 / (double)B_global), (real)0, (real*)out_loss);
        h->launches++;
    }
    if (grad_A) {
        if (need_grad) { if ((rc = finalize_net<real>(h, h->nA, thA, a.slabA, L.grid, (real*)(ws + L.raw), grad_A, st))) return rc; }
        else if (flags & DPB_FLAG_NEED_GRAD) DPB_CUDA(h, cudaMemsetAsync(grad_A, 0, (size_t)h->nA.ftotal * sizeof(real), st));
    }
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}


// ------------------------------------------------------------------------------------------------ tensor path
static int tc_pack(dpb_handle* h, const tc::TcNet& t, const void* theta, char* ws, size_t img, size_t vec, cudaStream_t st) {
    long long work = 0;
    for (int l = 0; l <= t.L; ++l) work += (long long)t.ly[l].K16 * t.ly[l].N16;
    int blocks = (int)((work + 255) / 256);
    if (blocks > 592) blocks = 592;
    tc::tc_pack_kernel<<<blocks, 256, 0, st>>>(t, (const float*)theta, (unsigned char*)(ws + img), (float*)(ws + vec), (float)BN_C, h->img_rep,
                                               (long long)a256(t.img_bytes));
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

// The <DP, EQN, MV> instantiations keep the per-path vectors in registers (DP entries each, every loop unrolled, the
// equation folded in at compile time); the fewer entries, the fewer registers the owner threads spill: DP = 12 serves the
// d <= 12 configs (d5 / d10, vdp_d4 / vdp_d10), DP = 20 the d = 20 ones (LQR, LQR_var, vdp_d20), DP = 24 ekn (its actor has
// control_dim + 1 outputs) up to d = 23.  VDP's cyclic neighbours are static only for a compile-time control_dim: the shipped
// 2, 5, 10 are instantiated.  Everything else runs the generic run-time-loop kernels; DPB_TC_GENERIC=1 forces them.
enum { TCK_GENERIC = 0, TCK_LQR, TCK_LQR12, TCK_LQRVAR, TCK_LQRVAR12, TCK_EKN, TCK_EKN12, TCK_VDP2, TCK_VDP5, TCK_VDP10 };
static int tc_pick(const dpb_handle* h) {
    static const bool force_generic = getenv("DPB_TC_GENERIC") != nullptr;
    if (force_generic) return TCK_GENERIC;
    const int d = h->cfg.dim, m = h->cfg.control_dim;
    switch (h->cfg.eqn) {
    case DPB_EQN_LQR: return d <= 12 ? TCK_LQR12 : (d <= 20 ? TCK_LQR : TCK_GENERIC);
    case DPB_EQN_LQR_VAR: return d <= 12 ? TCK_LQRVAR12 : (d <= 20 ? TCK_LQRVAR : TCK_GENERIC);
    case DPB_EQN_EKN: return m + 1 <= 12 ? TCK_EKN12 : (m + 1 <= 24 && d <= 23 ? TCK_EKN : TCK_GENERIC);
    case DPB_EQN_VDP: return (m == 2 && d <= 12) ? TCK_VDP2 : (m == 5 && d <= 12) ? TCK_VDP5 : (m == 10 && d <= 20) ? TCK_VDP10 : TCK_GENERIC;
    }
    return TCK_GENERIC;
}
static tc::TcKernelFn tc_pick_critic(const dpb_handle* h) {
    switch (tc_pick(h)) {
    case TCK_LQR: return tc::tc_get_critic_lqr();
    case TCK_LQR12: return tc::tc_get_critic_lqr12();
    case TCK_LQRVAR: return tc::tc_get_critic_lqrvar();
    case TCK_LQRVAR12: return tc::tc_get_critic_lqrvar12();
    case TCK_EKN: return tc::tc_get_critic_ekn();
    case TCK_EKN12: return tc::tc_get_critic_ekn12();
    case TCK_VDP2: return tc::tc_get_critic_vdp2();
    case TCK_VDP5: return tc::tc_get_critic_vdp5();
    case TCK_VDP10: return tc::tc_get_critic_vdp10();
    }
    return tc::tc_get_critic_generic();
}
static tc::TcKernelFn tc_pick_actor(const dpb_handle* h) {
    switch (tc_pick(h)) {
    case TCK_LQR: return tc::tc_get_actor_lqr();
    case TCK_LQR12: return tc::tc_get_actor_lqr12();
    case TCK_LQRVAR: return tc::tc_get_actor_lqrvar();
    case TCK_LQRVAR12: return tc::tc_get_actor_lqrvar12();
    case TCK_EKN: return tc::tc_get_actor_ekn();
    case TCK_EKN12: return tc::tc_get_actor_ekn12();
    case TCK_VDP2: return tc::tc_get_actor_vdp2();
    case TCK_VDP5: return tc::tc_get_actor_vdp5();
    case TCK_VDP10: return tc::tc_get_actor_vdp10();
    }
    return tc::tc_get_actor_generic();
}

// ring geometry for a launch: slots of `slot_bytes` (largest chunk of the images used) in what is left of 227 KB
static int tc_ring(dpb_handle* h, tc::TcArgs& a, bool grads) {
    a.actdz_bytes = grads ? tc::TC_PATHS * h->tc_maxw16 * 2 : 0;
    a.slot_bytes = h->tc_maxw16 * 64;                              // largest chunk: 16 contraction rows x the widest output, hi + lo
    const size_t fixed = tc::tc_smem_fixed(h->tA.vec_floats, h->tV.vec_floats, h->tG.vec_floats, a.actdz_bytes);
    const size_t avail = (size_t)227 * 1024;
    if (fixed + 4 * (size_t)a.slot_bytes > avail) return fail(h, DPB_ERR_ARG, "tensor path: networks too wide for the shared-memory operand images");
    size_t ns = (avail - fixed) / a.slot_bytes;
    a.nslot = (int)(ns > tc::MAX_NSLOT ? tc::MAX_NSLOT : ns);
    // Five slots already keep the weight stream ahead of the MMAs; if that fits in 196 KB the SM's unified array is carved
    // 196 KB shared / 60 KB L1 instead of 228 / 28, which the path threads' local-memory traffic (register spills, relu
    // masks) feels: critic 68.0 -> 66.5 ms, actor 66.9 -> 65.6 ms at 2^17 paths (3 or 4 slots: slower again).
    if (!getenv("DPB_TC_SMEM_MAX")) {                                   // (experiment knob: keep all 227 KB for the ring)
        const size_t cap = (size_t)196 * 1024 - 1024;
        if (fixed + 5 * (size_t)a.slot_bytes <= cap) { const int v = (int)((cap - fixed) / a.slot_bytes); if (v < a.nslot) a.nslot = v; }
    }
    // experiment knob: fewer ring slots leave more of the 256 KB unified array to L1 (local-memory traffic of the path threads)
    if (const char* e = getenv("DPB_TC_NSLOT")) { const int v = atoi(e); if (v >= 2 && v < a.nslot) a.nslot = v; }
    return DPB_OK;
}

static void tc_fill(dpb_handle* h, tc::TcArgs& a, const Layout& L, char* ws, const dpb_inputs* in, int64_t B_local, int64_t path_offset,
                    int64_t B_global, int32_t N, double T, uint32_t flags, const dpb_path_outputs* outs) {
    memset(&a, 0, sizeof(a));
    { EqnD e; fill_eqn(h->cfg, N, T, e); a.eqf = Eq<float>(e); }
    a.nA = h->tA; a.nV = h->tV; a.nG = h->tG;
    a.gA = h->sA; a.gV = h->sV; a.gG = h->sG;
    a.imgA = (const unsigned char*)(ws + L.imgA); a.imgV = (const unsigned char*)(ws + L.imgV); a.imgG = (const unsigned char*)(ws + L.imgG);
    a.img_rep = h->img_rep;
    a.img_strideA = (long long)a256(h->tA.img_bytes); a.img_strideV = (long long)a256(h->tV.img_bytes); a.img_strideG = (long long)a256(h->tG.img_bytes);
    a.x0 = (const float*)in->x0; a.dw = (const float*)in->dw; a.xb = (const float*)in->x_bdry;
    a.dw_mode = in->dw_mode; a.seed = in->seed; a.stream = in->stream; a.stream_base = (const unsigned long long*)in->stream_base;
    a.B_local = B_local; a.path_offset = path_offset;
    a.invB = (float)(1.0 / (double)B_global);
    a.N = N; a.flags = flags;
    a.loss_part = (float*)(ws + L.loss_part);
    a.scratch = (float*)(ws + L.scratch);
    a.scratch_per_cta = L.scratch_per_cta;
    a.copies = (unsigned char*)(ws + L.copies);
    a.copies_per_cta = L.copies_per_cta;
    a.stats = (long long*)(ws + L.stats);
    a.tile_counter = (int*)(ws + L.stats + (size_t)h->num_sms * 16 * 8);
#ifdef DPB_TC_STATS
    a.trace = (unsigned long long*)(ws + L.trace);
#else
    a.trace = nullptr;
#endif
    a.nslab = L.nslab;
    a.sr = h->sr;
    if (outs) {
        a.o_x = (float*)outs->x_smp; a.o_dt = (float*)outs->dt; a.o_coef = (float*)outs->coef;
        a.o_delta = (float*)outs->delta; a.o_delta_b = (float*)outs->delta_bdry; a.o_exit = outs->exit_index;
    }
}

static int tc_finalize(dpb_handle* h, const tc::TcNet& t, const tc::TcSlab& g, const void* theta, const float* slabs, int nslab, float* raw,
                       void* grad, cudaStream_t st) {
    int blocks = (int)((g.gtotal + 255) / 256);
    if (blocks > 592) blocks = 592;
    reduce_slabs_kernel<float><<<blocks, 256, 0, st>>>(slabs, nslab, g.gtotal, raw);
    tc::tc_finalize_grad_kernel<<<blocks, 256, 0, st>>>(t, g, (const float*)theta, raw, (float*)grad, (float)BN_C);
    h->launches += 2;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

// Naive scheme (equation.py:46-71): a path that has left the domain stays frozen for the rest of the N steps, and a tile of 128
// paths can stop only when all of them have (live fraction 0.2 at lqr_d5, 0.015 expected at d=20: SURVEY 8d).  So the paths are
// first rolled out forward-only (the actor network alone: ~1/5 of a training step), sorted by their number of accepted steps
// (longest first) and then tiled in that order: the tiles die as a whole and the dynamic tile scheduler hands the long ones out
// first.  The permutation is a pure relabelling -- every path sees the same x0 / increments (Philox is keyed by the path index)
// and writes its outputs at its own index -- so any permutation gives the same per-path results.
static bool tc_lifetime_sort_wanted(const dpb_handle* h, int64_t B_local, uint32_t flags) {
    static const bool off = getenv("DPB_NO_LIFETIME_SORT") != nullptr;
    // (only when there are more tiles than SMs: in a single wave every tile runs at once and the launch lasts as long as the
    //  longest-lived path whatever the tiling -- the pre-pass would only add its own latency: lqr_d5 at 1024 paths 4.09 ms per
    //  iteration with the sort, 3.25 ms without)
    return !off && !(h->cfg.reserved[0] & 1) && h->cfg.scheme == DPB_SCHEME_NAIVE && !(flags & DPB_FLAG_PROPAGATE_ONLY) &&
           (B_local + tc::TC_PATHS - 1) / tc::TC_PATHS > h->num_sms;
}
static int tc_lifetime_sort(dpb_handle* h, tc::TcKernelFn critic_kern, tc::TcArgs a /* by value: a forward-only copy */, const Layout& L, char* ws,
                            size_t smem, int64_t B_local, int32_t N, cudaStream_t st) {
    a.flags = (a.flags & DPB_FLAG_CHEAT_CONTROL) | DPB_FLAG_PROPAGATE_ONLY;
    a.o_x = a.o_dt = a.o_coef = a.o_delta = a.o_delta_b = nullptr;
    a.o_exit = (int*)(ws + L.life_nacc);
    a.perm = nullptr;
    a.slabV = a.slabG = a.slabA = nullptr;
    a.vecV = a.vecG = nullptr;
    DPB_CUDA(h, cudaMemsetAsync(a.tile_counter, 0, sizeof(int), st));
    DPB_CUDA(h, cudaMemsetAsync(ws + L.life_hist, 0, (size_t)(N + 2) * 4, st));
    cudaFuncAttributes fa;
    DPB_CUDA(h, cudaFuncGetAttributes(&fa, critic_kern));
    critic_kern<<<L.grid, fa.maxThreadsPerBlock, smem, st>>>(a);
    int blocks = (int)((B_local + 255) / 256);
    if (blocks > 592) blocks = 592;
    lifetime_hist_kernel<<<blocks, 256, 0, st>>>((const int*)(ws + L.life_nacc), B_local, N, (int*)(ws + L.life_hist));
    lifetime_scan_kernel<<<1, 32, 0, st>>>((int*)(ws + L.life_hist), N + 1);
    lifetime_scatter_kernel<<<blocks, 256, 0, st>>>((const int*)(ws + L.life_nacc), B_local, N, (int*)(ws + L.life_hist), (int*)(ws + L.life_perm));
    h->launches += 4;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

static int critic_step_tc(dpb_handle* h, const void* thA, const void* thV, const void* thG, const dpb_inputs* in, int64_t B_local,
                          int64_t path_offset, int64_t B_global, int32_t N, double T, uint32_t flags, void* out_loss, void* grad_V,
                          void* grad_G, const dpb_path_outputs* outs, void* workspace, cudaStream_t st) {
    char* ws = (char*)workspace;
    const Layout L = make_layout(h, B_local, N);
    tc::TcArgs a;
    tc_fill(h, a, L, ws, in, B_local, path_offset, B_global, N, T, flags, outs);
    const bool cheat = flags & DPB_FLAG_CHEAT_CONTROL, prop_only = flags & DPB_FLAG_PROPAGATE_ONLY;
    const bool need_grad = (flags & DPB_FLAG_NEED_GRAD) && !prop_only;
    const bool td1 = h->cfg.td_type == DPB_TD1;
    int rc;
    if ((rc = tc_ring(h, a, need_grad))) return rc;
    if (!cheat) { if ((rc = tc_pack(h, h->tA, thA, ws, L.imgA, L.vecA, st))) return rc; a.vecA = (const float*)(ws + L.vecA); }
    if (!prop_only) {
        if ((rc = tc_pack(h, h->tV, thV, ws, L.imgV, L.vecV, st))) return rc;
        a.vecV = (const float*)(ws + L.vecV);
        if (td1) { if ((rc = tc_pack(h, h->tG, thG, ws, L.imgG, L.vecG, st))) return rc; a.vecG = (const float*)(ws + L.vecG); }
    }
    if (need_grad) {
        a.slabV = (float*)(ws + L.slabs);
        a.slabG = a.slabV + (size_t)L.nslab * h->sV.gtotal;
        DPB_CUDA(h, cudaMemsetAsync(ws + L.slabs, 0, (size_t)L.nslab * (h->sV.gtotal + h->sG.gtotal) * sizeof(float), st));
    }
    const size_t smem = tc::tc_smem_bytes(h->tA.vec_floats, h->tV.vec_floats, h->tG.vec_floats, a.actdz_bytes, a.nslot, a.slot_bytes);
    tc::TcKernelFn kern = tc_pick_critic(h);
    DPB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DPB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)((smem + 1024) * 100 / (228 * 1024) + 1 > 100 ? 100 : (smem + 1024) * 100 / (228 * 1024) + 1)));
    if (tc_lifetime_sort_wanted(h, B_local, flags)) {
        if ((rc = tc_lifetime_sort(h, kern, a, L, ws, smem, B_local, N, st))) return rc;
        a.perm = (const int*)(ws + L.life_perm);
    }
    DPB_CUDA(h, cudaMemsetAsync(a.tile_counter, 0, sizeof(int), st));
    // Small batch (the tiles occupy fewer than half of the SMs): the second sweep -- one independent backward per stored step --
    // runs as a second launch that spreads (tile, block of steps) items over all SMs (dpb_tc_kernels.cuh, F_S2_ONLY)
    static const bool no_split = getenv("DPB_NO_SWEEP_SPLIT") != nullptr;
    const long long ntl = (B_local + tc::TC_PATHS - 1) / tc::TC_PATHS;
    const bool split = !no_split && need_grad && td1 && ntl * 2 <= h->num_sms;
    if (split) {
        a.flags |= tc::F_S2_DEFER;
        a.s2_rhog = (float*)(ws + L.s2_rhog);
        a.s2_tlive = (int*)(ws + L.s2_tlive);
        int nblk = (int)((h->num_sms + ntl - 1) / ntl);
        if (nblk > N) nblk = N;
        a.s2_chunk = (N + nblk - 1) / nblk;
    }
    ev_begin(h, st);
    cudaFuncAttributes fa;                                       // threads per CTA = the launch bound of the instantiation
    DPB_CUDA(h, cudaFuncGetAttributes(&fa, kern));               // (its translation unit chooses the number of path-thread groups)
    kern<<<L.grid, fa.maxThreadsPerBlock, smem, st>>>(a);
    h->launches++;
    if (split) {
        tc::TcArgs b = a;
        b.flags = (a.flags & ~tc::F_S2_DEFER) | tc::F_S2_ONLY;
        b.loss_part = nullptr;
        const long long items = ntl * ((N + a.s2_chunk - 1) / a.s2_chunk);
        const int grid2 = (int)(items < h->num_sms ? items : h->num_sms);
        kern<<<grid2, fa.maxThreadsPerBlock, smem, st>>>(b);
        h->launches++;
    }
    ev_end(h, st);
    DPB_CUDA(h, cudaGetLastError());
    if (out_loss && !prop_only) {
        const float s = (float)(100.0 / (double)B_global);
        reduce_loss_kernel<float><<<1, 32, 0, st>>>(a.loss_part, L.grid, s, s, (float*)out_loss);
        h->launches++;
    }
    if (need_grad) {
        float* raw = (float*)(ws + L.raw);
        if (grad_V) { if ((rc = tc_finalize(h, h->tV, h->sV, thV, a.slabV, L.nslab, raw, grad_V, st))) return rc; }
        if (grad_G) {
            if (td1) { if ((rc = tc_finalize(h, h->tG, h->sG, thG, a.slabG, L.nslab, raw + h->sV.gtotal, grad_G, st))) return rc; }
            else DPB_CUDA(h, cudaMemsetAsync(grad_G, 0, (size_t)h->nG.ftotal * sizeof(float), st));
        }
    }
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

static int actor_step_tc(dpb_handle* h, const void* thA, const void* thV, const dpb_inputs* in, int64_t B_local, int64_t path_offset,
                         int64_t B_global, int32_t N, double T, uint32_t flags, void* out_loss, void* grad_A,
                         const dpb_path_outputs* outs, void* workspace, cudaStream_t st) {
    char* ws = (char*)workspace;
    const Layout L = make_layout(h, B_local, N);
    tc::TcArgs a;
    tc_fill(h, a, L, ws, in, B_local, path_offset, B_global, N, T, flags, outs);
    const bool cheat = flags & DPB_FLAG_CHEAT_CONTROL, cheat_v = flags & DPB_FLAG_CHEAT_VALUE;
    const bool need_grad = (flags & DPB_FLAG_NEED_GRAD) && !cheat;
    int rc;
    if ((rc = tc_ring(h, a, need_grad))) return rc;
    if (!cheat) { if ((rc = tc_pack(h, h->tA, thA, ws, L.imgA, L.vecA, st))) return rc; a.vecA = (const float*)(ws + L.vecA); }
    if (!cheat_v) { if ((rc = tc_pack(h, h->tV, thV, ws, L.imgV, L.vecV, st))) return rc; a.vecV = (const float*)(ws + L.vecV); }
    if (need_grad) {
        a.slabA = (float*)(ws + L.slabs);
        DPB_CUDA(h, cudaMemsetAsync(ws + L.slabs, 0, (size_t)L.nslab * h->sA.gtotal * sizeof(float), st));
    }
    const size_t smem = tc::tc_smem_bytes(h->tA.vec_floats, h->tV.vec_floats, h->tG.vec_floats, a.actdz_bytes, a.nslot, a.slot_bytes);
    tc::TcKernelFn kern = tc_pick_actor(h);
    DPB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DPB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)((smem + 1024) * 100 / (228 * 1024) + 1 > 100 ? 100 : (smem + 1024) * 100 / (228 * 1024) + 1)));
    if (tc_lifetime_sort_wanted(h, B_local, flags)) {
        tc::TcKernelFn ck = tc_pick_critic(h);                   // the forward-only rollout of the critic kernel (same per-path arithmetic)
        DPB_CUDA(h, cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if ((rc = tc_lifetime_sort(h, ck, a, L, ws, smem, B_local, N, st))) return rc;
        a.perm = (const int*)(ws + L.life_perm);
    }
    DPB_CUDA(h, cudaMemsetAsync(a.tile_counter, 0, sizeof(int), st));
    ev_begin(h, st);
    cudaFuncAttributes fa;                                       // threads per CTA = the launch bound of the instantiation
    DPB_CUDA(h, cudaFuncGetAttributes(&fa, kern));               // (its translation unit chooses the number of path-thread groups)
    kern<<<L.grid, fa.maxThreadsPerBlock, smem, st>>>(a);
    ev_end(h, st);
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    if (out_loss) {
        reduce_loss_kernel<float><<<1, 32, 0, st>>>(a.loss_part, L.grid, (float)(1.0 / (double)B_global), 0.f, (float*)out_loss);
        h->launches++;
    }
    if (grad_A) {
        if (need_grad) { if ((rc = tc_finalize(h, h->tA, h->sA, thA, a.slabA, L.nslab, (float*)(ws + L.raw), grad_A, st))) return rc; }
        else if (flags & DPB_FLAG_NEED_GRAD) DPB_CUDA(h, cudaMemsetAsync(grad_A, 0, (size_t)h->nA.ftotal * sizeof(float), st));
    }
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

extern "C" {

int dpb_critic_step(dpb_handle* h, const void* theta_actor, const void* theta_V, const void* theta_G, const dpb_inputs* in,
                    int64_t B_local, int64_t path_offset, int64_t B_global, int32_t N, double T, uint32_t flags, void* out_loss,
                    void* grad_V, void* grad_G, const dpb_path_outputs* outs, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_step_args(h, in, B_local, B_global, N, T, workspace, workspace_bytes, "dpb_critic_step");
    if (rc) return rc;
    const bool prop_only = flags & DPB_FLAG_PROPAGATE_ONLY, cheat = flags & DPB_FLAG_CHEAT_CONTROL;
    if (!cheat && !theta_actor) return fail(h, DPB_ERR_ARG, "dpb_critic_step: theta_actor is NULL");
    if (!prop_only && (!theta_V || !in->x_bdry)) return fail(h, DPB_ERR_ARG, "dpb_critic_step: theta_V and x_bdry are required");
    if (!prop_only && h->cfg.td_type == DPB_TD1 && !theta_G) return fail(h, DPB_ERR_ARG, "dpb_critic_step: theta_G is NULL under TD1");
    cudaStream_t st = (cudaStream_t)stream;
    if (h->cfg.impl == DPB_IMPL_TENSOR)
        return critic_step_tc(h, theta_actor, theta_V, theta_G, in, B_local, path_offset, B_global, N, T, flags, out_loss, grad_V, grad_G, outs, workspace, st);
    if (h->cfg.dtype == DPB_F64)
        return critic_step_t<double>(h, theta_actor, theta_V, theta_G, in, B_local, path_offset, B_global, N, T, flags, out_loss, grad_V, grad_G, outs, workspace, st);
    return critic_step_t<float>(h, theta_actor, theta_V, theta_G, in, B_local, path_offset, B_global, N, T, flags, out_loss, grad_V, grad_G, outs, workspace, st);
}

int dpb_actor_step(dpb_handle* h, const void* theta_actor, const void* theta_V, const dpb_inputs* in, int64_t B_local,
                   int64_t path_offset, int64_t B_global, int32_t N, double T, uint32_t flags, void* out_loss, void* grad_actor,
                   const dpb_path_outputs* outs, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_step_args(h, in, B_local, B_global, N, T, workspace, workspace_bytes, "dpb_actor_step");
    if (rc) return rc;
    if (!(flags & DPB_FLAG_CHEAT_CONTROL) && !theta_actor) return fail(h, DPB_ERR_ARG, "dpb_actor_step: theta_actor is NULL");
    if (!(flags & DPB_FLAG_CHEAT_VALUE) && !theta_V) return fail(h, DPB_ERR_ARG, "dpb_actor_step: theta_V is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (h->cfg.impl == DPB_IMPL_TENSOR)
        return actor_step_tc(h, theta_actor, theta_V, in, B_local, path_offset, B_global, N, T, flags, out_loss, grad_actor, outs, workspace, st);
    if (h->cfg.dtype == DPB_F64)
        return actor_step_t<double>(h, theta_actor, theta_V, in, B_local, path_offset, B_global, N, T, flags, out_loss, grad_actor, outs, workspace, st);
    return actor_step_t<float>(h, theta_actor, theta_V, in, B_local, path_offset, B_global, N, T, flags, out_loss, grad_actor, outs, workspace, st);
}

static int host_stage(dpb_handle* h, const dpb_inputs* in_host, dpb_inputs* dev, int64_t B_local, int32_t N, char* ws, int64_t ws_bytes,
                      int64_t* used, bool need_xb, cudaStream_t st) {
    const size_t es = esize(h);
    const int64_t base = dpb_workspace_bytes(h, B_local, N);
    const int64_t stg = dpb_staging_bytes(h, B_local, N, in_host->dw_mode);
    if (ws_bytes < base + stg) return fail(h, DPB_ERR_WORKSPACE, "host entry point: workspace too small, need " + std::to_string(base + stg) + " bytes");
    char* p = ws + base;
    const size_t nx = (size_t)B_local * h->cfg.dim * es;
    *dev = *in_host;
    DPB_CUDA(h, cudaMemcpyAsync(p, in_host->x0, nx, cudaMemcpyHostToDevice, st));
    dev->x0 = p; p += a256(nx);
    if (need_xb && in_host->x_bdry) {
        DPB_CUDA(h, cudaMemcpyAsync(p, in_host->x_bdry, nx, cudaMemcpyHostToDevice, st));
        dev->x_bdry = p;
    }
    p += a256(nx);
    if (in_host->dw_mode == DPB_DW_EXTERNAL) {
        DPB_CUDA(h, cudaMemcpyAsync(p, in_host->dw, nx * N, cudaMemcpyHostToDevice, st));
        dev->dw = p;
    }
    *used = base;
    return DPB_OK;
}

int dpb_critic_step_host(dpb_handle* h, const void* theta_actor, const void* theta_V, const void* theta_G, const dpb_inputs* in_host,
                         int64_t B_local, int64_t path_offset, int64_t B_global, int32_t N, double T, uint32_t flags,
                         void* out_loss_host, void* grad_V, void* grad_G, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!h || !in_host || !in_host->x0 || !workspace) return fail(h, DPB_ERR_ARG, "dpb_critic_step_host: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    dpb_inputs dev;
    int64_t base = 0;
    int rc = host_stage(h, in_host, &dev, B_local, N, (char*)workspace, workspace_bytes, &base, true, st);
    if (rc) return rc;
    const Layout L = make_layout(h, B_local, N);
    void* loss_dev = (char*)workspace + L.loss_out;
    rc = dpb_critic_step(h, theta_actor, theta_V, theta_G, &dev, B_local, path_offset, B_global, N, T, flags, loss_dev, grad_V, grad_G,
                         nullptr, workspace, base, stream);
    if (rc) return rc;
    if (out_loss_host) DPB_CUDA(h, cudaMemcpyAsync(out_loss_host, loss_dev, 2 * esize(h), cudaMemcpyDeviceToHost, st));
    DPB_CUDA(h, cudaStreamSynchronize(st));
    return DPB_OK;
}

int dpb_actor_step_host(dpb_handle* h, const void* theta_actor, const void* theta_V, const dpb_inputs* in_host, int64_t B_local,
                        int64_t path_offset, int64_t B_global, int32_t N, double T, uint32_t flags, void* out_loss_host,
                        void* grad_actor, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!h || !in_host || !in_host->x0 || !workspace) return fail(h, DPB_ERR_ARG, "dpb_actor_step_host: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    dpb_inputs dev;
    int64_t base = 0;
    int rc = host_stage(h, in_host, &dev, B_local, N, (char*)workspace, workspace_bytes, &base, false, st);
    if (rc) return rc;
    const Layout L = make_layout(h, B_local, N);
    void* loss_dev = (char*)workspace + L.loss_out;
    rc = dpb_actor_step(h, theta_actor, theta_V, &dev, B_local, path_offset, B_global, N, T, flags, loss_dev, grad_actor, nullptr,
                        workspace, base, stream);
    if (rc) return rc;
    if (out_loss_host) DPB_CUDA(h, cudaMemcpyAsync(out_loss_host, loss_dev, 2 * esize(h), cudaMemcpyDeviceToHost, st));
    DPB_CUDA(h, cudaStreamSynchronize(st));
    return DPB_OK;
}

int dpb_mlp_forward(dpb_handle* h, int which, const void* theta, const void* x, int64_t n, void* out, void* workspace,
                    int64_t workspace_bytes, void* stream) {
    if (!h) return fail(nullptr, DPB_ERR_ARG, "dpb_mlp_forward: null handle");
    if (!theta || !x || !out || n < 1) return fail(h, DPB_ERR_ARG, "dpb_mlp_forward: null argument or n < 1");
    const NetDev* nd = which == DPB_NET_ACTOR ? &h->nA : which == DPB_NET_CRITIC ? &h->nV : which == DPB_NET_CRITIC_GRAD ? &h->nG : nullptr;
    if (!nd) return fail(h, DPB_ERR_ARG, "dpb_mlp_forward: unknown network id");
    const size_t es = esize(h);
    if (!workspace || workspace_bytes < (int64_t)a256(nd->ptotal * es)) return fail(h, DPB_ERR_WORKSPACE, "dpb_mlp_forward: workspace too small");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(h, DPB_ERR_CUDA, "dpb_mlp_forward: no CUDA device (there is no CPU fallback)"); }
    cudaStream_t st = (cudaStream_t)stream;
    const int P = tileP(h);
    long long ntiles = (n + P - 1) / P;
    const int grid = (int)(ntiles < h->num_sms ? ntiles : h->num_sms);
    int rc;
    if (h->cfg.dtype == DPB_F64) {
        if ((rc = pack_net<double>(h, *nd, theta, workspace, st))) return rc;
        const size_t smem = carve_elems<double>(h->sr, h->hrows, 2) * 8;
        DPB_CUDA(h, cudaFuncSetAttribute(mlp_forward_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_forward_kernel<double><<<grid, NTHREADS, smem, st>>>(*nd, (const double*)workspace, (const double*)x, n, (double*)out, h->sr, h->hrows);
    } else {
        if ((rc = pack_net<float>(h, *nd, theta, workspace, st))) return rc;
        const size_t smem = carve_elems<float>(h->sr, h->hrows, 2) * 4;
        DPB_CUDA(h, cudaFuncSetAttribute(mlp_forward_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_forward_kernel<float><<<grid, NTHREADS, smem, st>>>(*nd, (const float*)workspace, (const float*)x, n, (float*)out, h->sr, h->hrows);
    }
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

int dpb_closed_form(dpb_handle* h, int which, const void* x, const void* u, int64_t n, void* out, void* stream) {
    if (!h) return fail(nullptr, DPB_ERR_ARG, "dpb_closed_form: null handle");
    if (!x || !out || n < 1 || which < 0 || which > 6 || ((which == 4 || which == 5) && !u)) return fail(h, DPB_ERR_ARG, "dpb_closed_form: bad argument");
    if (which == 6 && !u && h->cfg.eqn == DPB_EQN_LQR_VAR) return fail(h, DPB_ERR_ARG, "dpb_closed_form: sigma of LQR_var depends on u");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(h, DPB_ERR_CUDA, "dpb_closed_form: no CUDA device (there is no CPU fallback)"); }
    EqnD e;
    fill_eqn(h->cfg, 1, 1.0, e);
    const int blocks = (int)((n + 127) / 128);
    if (h->cfg.dtype == DPB_F64) closed_form_kernel<double><<<blocks, 128, 0, (cudaStream_t)stream>>>(e, which, (const double*)x, (const double*)u, n, (double*)out);
    else closed_form_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>(e, which, (const float*)x, (const float*)u, n, (float*)out);
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

int dpb_diffusion(dpb_handle* h, const void* x, const void* u, const void* dw, int64_t n, void* out, void* stream) {
    if (!h) return fail(nullptr, DPB_ERR_ARG, "dpb_diffusion: null handle");
    if (!x || !dw || !out || n < 1 || (!u && h->cfg.eqn == DPB_EQN_LQR_VAR)) return fail(h, DPB_ERR_ARG, "dpb_diffusion: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(h, DPB_ERR_CUDA, "dpb_diffusion: no CUDA device (there is no CPU fallback)"); }
    EqnD e;
    fill_eqn(h->cfg, 1, 1.0, e);
    const int blocks = (int)((n + 127) / 128);
    if (h->cfg.dtype == DPB_F64) diffusion_kernel<double><<<blocks, 128, 0, (cudaStream_t)stream>>>(e, (const double*)x, (const double*)u, (const double*)dw, n, (double*)out);
    else diffusion_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>(e, (const float*)x, (const float*)u, (const float*)dw, n, (float*)out);
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

int dpb_adam_step(dpb_handle* h, void* theta, const void* grad, void* m, void* v, int64_t n, double lr_t, const double* lr_t_dev,
                  double beta1, double beta2, double eps, void* stream) {
    if (!h) return fail(nullptr, DPB_ERR_ARG, "dpb_adam_step: null handle");
    if (!theta || !grad || !m || !v || n < 1) return fail(h, DPB_ERR_ARG, "dpb_adam_step: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(h, DPB_ERR_CUDA, "dpb_adam_step: no CUDA device (there is no CPU fallback)"); }
    int blocks = (int)((n + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    if (h->cfg.dtype == DPB_F64) adam_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((double*)theta, (const double*)grad, (double*)m, (double*)v, n, lr_t, lr_t_dev, beta1, beta2, eps);
    else adam_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)theta, (const float*)grad, (float*)m, (float*)v, n, (float)lr_t, lr_t_dev, (float)beta1, (float)beta2, (float)eps);
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

int dpb_philox_dw(dpb_handle* h, int32_t dw_mode, uint64_t seed, uint64_t stream_id, int64_t path_offset, int64_t B_local, int32_t N,
                  void* dw_out, void* stream) {
    if (!h) return fail(nullptr, DPB_ERR_ARG, "dpb_philox_dw: null handle");
    if (!dw_out || B_local < 1 || N < 1 || (dw_mode != DPB_DW_PHILOX_NORMAL && dw_mode != DPB_DW_PHILOX_BOUNDED)) return fail(h, DPB_ERR_ARG, "dpb_philox_dw: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(h, DPB_ERR_CUDA, "dpb_philox_dw: no CUDA device (there is no CPU fallback)"); }
    const int d = h->cfg.dim;
    long long total = (long long)B_local * N * ((d + 3) / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 2368) blocks = 2368;
    if (h->cfg.dtype == DPB_F64) philox_dw_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(dw_mode, seed, stream_id, path_offset, B_local, d, N, (double*)dw_out);
    else philox_dw_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(dw_mode, seed, stream_id, path_offset, B_local, d, N, (float*)dw_out);
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

int dpb_sample_x(dpb_handle* h, uint64_t seed, uint64_t stream_id, const uint64_t* stream_base, int64_t path_offset, int64_t B_local,
                 void* x0_out, void* xb_out, void* stream) {
    if (!h) return fail(nullptr, DPB_ERR_ARG, "dpb_sample_x: null handle");
    if ((!x0_out && !xb_out) || B_local < 1) return fail(h, DPB_ERR_ARG, "dpb_sample_x: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(h, DPB_ERR_CUDA, "dpb_sample_x: no CUDA device (there is no CPU fallback)"); }
    const int blocks = (int)((B_local + 127) / 128);
    if (h->cfg.dtype == DPB_F64) sample_x_kernel<double><<<blocks, 128, 0, (cudaStream_t)stream>>>(seed, stream_id, (const unsigned long long*)stream_base, path_offset, B_local, h->cfg.dim, h->cfg.R, (double*)x0_out, (double*)xb_out);
    else sample_x_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>(seed, stream_id, (const unsigned long long*)stream_base, path_offset, B_local, h->cfg.dim, (float)h->cfg.R, (float*)x0_out, (float*)xb_out);
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

}  // extern "C"

extern "C" int dpb_tc_selftest(const float* A, const float* B, float* D, int K, void* stream) {
    if (!A || !B || !D || K < 128 || K > 208 || (K % 16)) return fail(nullptr, DPB_ERR_ARG, "dpb_tc_selftest: need 128 <= K <= 208, K % 16 == 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(nullptr, DPB_ERR_CUDA, "dpb_tc_selftest: no CUDA device"); }
    const int smem = 128 * K * 2 + tc::ST_N * K * 2 + 128 * tc::ST_N * 2 + 64;
    DPB_CUDA(nullptr, cudaFuncSetAttribute(tc::tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc::tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, K);
    DPB_CUDA(nullptr, cudaGetLastError());
    return DPB_OK;
}

extern "C" int dpb_tc_stats(dpb_handle* h, const void* workspace, int64_t B_local, int32_t N, int64_t* out_host) {
    if (!h || !workspace || !out_host || h->cfg.impl != DPB_IMPL_TENSOR) return fail(h, DPB_ERR_ARG, "dpb_tc_stats: tensor-path handle and workspace required");
    const Layout L = make_layout(h, B_local, N);
    DPB_CUDA(h, cudaMemcpy(out_host, (const char*)workspace + L.stats, 16 * 8, cudaMemcpyDeviceToHost));
    return DPB_OK;
}

extern "C" int dpb_tc_trace(dpb_handle* h, const void* workspace, int64_t B_local, int32_t N, uint64_t* out_host) {
    if (!h || !workspace || !out_host || h->cfg.impl != DPB_IMPL_TENSOR) return fail(h, DPB_ERR_ARG, "dpb_tc_trace: tensor-path handle and workspace required");
#ifndef DPB_TC_STATS
    return fail(h, DPB_ERR_ARG, "dpb_tc_trace: the library was built without DPB_TC_STATS");
#else
    const Layout L = make_layout(h, B_local, N);
    DPB_CUDA(h, cudaMemcpy(out_host, (const char*)workspace + L.trace, (size_t)3 * tc::TC_TRACE_CAP * 8, cudaMemcpyDeviceToHost));
    return DPB_OK;
#endif
}

extern "C" int dpb_err_metrics(dpb_handle* h, const void* truth, const void* approx, int64_t n, void* out3, void* stream) {
    if (!h) return fail(nullptr, DPB_ERR_ARG, "dpb_err_metrics: null handle");
    if (!truth || !approx || !out3 || n < 1) return fail(h, DPB_ERR_ARG, "dpb_err_metrics: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(h, DPB_ERR_CUDA, "dpb_err_metrics: no CUDA device (there is no CPU fallback)"); }
    if (h->cfg.dtype == DPB_F64) err_metrics_kernel<double><<<1, 256, 0, (cudaStream_t)stream>>>((const double*)truth, (const double*)approx, n, (double*)out3);
    else err_metrics_kernel<float><<<1, 256, 0, (cudaStream_t)stream>>>((const float*)truth, (const float*)approx, n, (float*)out3);
    h->launches++;
    DPB_CUDA(h, cudaGetLastError());
    return DPB_OK;
}

extern "C" int dpb_tc_handshake_cycles(int64_t* out_host, int rounds) {
    if (!out_host || rounds < 1) return fail(nullptr, DPB_ERR_ARG, "dpb_tc_handshake_cycles: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(nullptr, DPB_ERR_CUDA, "dpb_tc_handshake_cycles: no CUDA device"); }
    long long* d = nullptr;
    DPB_CUDA(nullptr, cudaMalloc(&d, 16));
    tc::tc_handshake_kernel<<<1, 288>>>(d, rounds);
    cudaError_t e = cudaMemcpy(out_host, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(nullptr, DPB_ERR_CUDA, std::string("dpb_tc_handshake_cycles: ") + cudaGetErrorString(e));
    return DPB_OK;
}

extern "C" int dpb_tc_epilogue_cycles(int64_t* out_host, int rounds, int ngroups, int with_mma, int publish) {
    if (!out_host || rounds < 1 || ngroups < 1 || ngroups > 4) return fail(nullptr, DPB_ERR_ARG, "dpb_tc_epilogue_cycles: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(nullptr, DPB_ERR_CUDA, "dpb_tc_epilogue_cycles: no CUDA device"); }
    long long* d = nullptr;
    DPB_CUDA(nullptr, cudaMalloc(&d, 16));
    tc::tc_epilogue_bench_kernel<<<1, 128 * ngroups + 32>>>(d, rounds, ngroups, with_mma, publish);
    cudaError_t e = cudaMemcpy(out_host, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(nullptr, DPB_ERR_CUDA, std::string("dpb_tc_epilogue_cycles: ") + cudaGetErrorString(e));
    return DPB_OK;
}

extern "C" int dpb_tc_mma_cycles(int64_t* out_host, int n, int rounds, int ts, int per_commit) {
    if (!out_host || n < 16 || n > 256 || (n % 16) || rounds < 1 || per_commit < 1) return fail(nullptr, DPB_ERR_ARG, "dpb_tc_mma_cycles: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(nullptr, DPB_ERR_CUDA, "dpb_tc_mma_cycles: no CUDA device"); }
    long long* d = nullptr;
    DPB_CUDA(nullptr, cudaMalloc(&d, 16));
    const int smem = 4096 + n * 32 + 1024;
    tc::tc_mma_bench_kernel<<<1, 128, smem>>>(d, n, rounds, ts, per_commit);
    cudaError_t e = cudaMemcpy(out_host, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(nullptr, DPB_ERR_CUDA, std::string("dpb_tc_mma_cycles: ") + cudaGetErrorString(e));
    return DPB_OK;
}
