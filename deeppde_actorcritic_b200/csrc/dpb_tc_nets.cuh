// dpb_tc_nets.cuh -- DeepNN (reference solver.py:227-278) on the 5th-generation tensor cores.
//
// One CTA owns a tile of 128 paths; path t = TMEM lane t.  Warp roles:
//   * warps 0-3, the OWNERS: thread t owns path t -- its state lives in that thread's registers -- and does everything that
//     needs the state: the per-path SDE arithmetic, writing a network's input (y0) or output cotangent as A planes, reading
//     a network's output (or the input cotangent dy0) back;
//   * warps 4 .. 4+4*TC_NGRP-1, the HELPERS (TC_NGRP groups of four warps, one warp per TMEM lane quadrant): stateless.
//     They run every hidden-layer epilogue (accumulator -> planes of the next product, relu masks, FP16 copies for the dW
//     products) and drain the dW accumulators into the gradient slab; chunk c of an epilogue belongs to group c % TC_NGRP;
//   * then the control warp (issues every tcgen05.mma) and the producer warp (streams the weights).
// The owners' arithmetic (Philox increments, step-size rule, Euler-Maruyama move, adjoint step) therefore runs WHILE the
// helpers and the tensor pipe work through a network; in round 1 the same threads did both, one after the other.
//
//   * dW products (a_l^T dz_l over the 128 paths of a tile) read both operands from shared memory as FP16 images (11-bit
//     significands: 8x finer than bf16 at the same size and MMA rate).  FP16's narrow exponent range is handled per backward
//     evaluation: path_put_dz takes the largest |cotangent| of the tile (one warp reduction + one named barrier), multiplies
//     every output cotangent by the power of two that brings it to [2^11, 2^12), and -- the whole backward chain being
//     linear -- the drains of the dW accumulators and the read-out of dy0 divide it out again (exact: powers of two).
//   * Precision: FP32 emulated with bf16 pairs.  Every operand x is carried as hi = bf16(x) and
//     lo = bf16(x - hi) and a product a*w is formed as ah*wh + ah*wl + al*wh with FP32 accumulation in
//     tensor memory (error ~2^-16 relative per term; the dropped al*wl term is ~2^-18).
//   * Activations never touch shared memory on the forward / dX products: the epilogue reads the FP32
//     accumulator of layer l from TMEM (tcgen05.ld), applies the BatchNorm affine + y+relu(y), splits
//     and writes the two bf16 planes back to TMEM (tcgen05.st), from where layer l+1 reads them as the
//     A operand (tcgen05.mma with A in tensor memory).
//   * BatchNorm folded into the products: the forward image of layer l holds W[k][n] * gamma_n c (c = 1/sqrt(1+eps)), its row
//     k = kl (the spare input feature, which every activation carries as a constant 1) holds beta_n (+ bias_n gamma_n c for
//     the last layer), and its entry (kl, nl) is 0.5, so that the accumulator IS z = a W gamma c + beta and the spare
//     output feature becomes z + relu(z) = 1 again: the epilogues neither load gamma/beta nor multiply.
//   * Weights: the pack kernel turns the flat FP32 parameters into bf16 hi/lo operand images (layout in
//     dpb_tc.cuh), cut into 16-wide contraction chunks of N16*64 bytes.  The control thread streams the
//     chunks L2 -> shared memory through a ring of slots with 1-D bulk copies (cp.async.bulk,
//     completion on an mbarrier); a slot is released by the tcgen05.commit of the MMAs that read it.
//     The chunk sequence of a phase is a cyclic schedule known in advance, so the stream runs ahead of
//     the MMAs across layers, networks and time steps.
//   * TMEM columns: two regions of 256 columns.  A product reads its A planes from one region and accumulates into the
//     other; the epilogue converts the FP32 accumulator IN PLACE into the planes of the next product (16 accumulator
//     columns of a chunk -> 8 columns of packed hi pairs + 8 columns of packed lo pairs), so the regions swap roles
//     from layer to layer.
//   * Chunk-granular hand-off: the epilogue publishes every 16-column chunk on its own mbarrier (a_chunk[c]) and the
//     next product starts on contraction chunk c as soon as it is there -- the tensor pipe works on layer l+1 while
//     the path threads are still converting the rest of layer l.  (Splitting the OUTPUT columns of a product into
//     two halves with separate commits was tried first: parity-green but slower -- a tcgen05.commit stalls the issuing
//     thread for ~300 cycles, which three N=112 MMAs (68 cycles each) do not cover, and one MMA costs 22 + 0.45 N
//     cycles, not N/2: profiles/r02_tcgen05_*.txt.)  Inputs that are written by the path threads themselves (y0, the
//     output cotangent) or that feed a dW product are published once, on a_all.
//   * Small weight chunks (narrow outputs: 2 KB for a 32-wide layer) are streamed several per ring slot, so that one
//     tcgen05.commit releases up to eight of them.
#pragma once
#include "dpb_nets.cuh"
#include "dpb_tc.cuh"

namespace dpb {
namespace tc {

constexpr int TC_PATHS = 128;
constexpr int TC_TRACE_CAP = 4096;             // events per role of the diagnostic trace (stats builds)
#ifndef DPB_TC_NGRP
#define DPB_TC_NGRP 2
#endif
constexpr int TC_NGRP = DPB_TC_NGRP;            // helper groups of 4 warps (chunk c of an epilogue -> group c % TC_NGRP).  0: no helper
                                                // warps -- the owners run the helpers' code themselves, one thread per path with up to
                                                // 255 registers (the actor kernels: their reverse sweep is one serial dependency chain,
                                                // so separate helpers only add hand-offs and cost the owners registers)
#ifndef DPB_TC_OWNHELP
#define DPB_TC_OWNHELP 0
#endif
// the owners take part in the epilogues and drains as group 0 (always when there are no helper warps; DPB_TC_OWNHELP=1: next
// to TC_NGRP helper groups, which are then groups 1..TC_NGRP)
constexpr bool TC_COMBINED = TC_NGRP == 0 || DPB_TC_OWNHELP != 0;
constexpr int TC_EGRP = TC_NGRP + (TC_COMBINED ? 1 : 0);      // groups the chunks of an epilogue are dealt out to
constexpr int TC_OWN_THREADS = 128;             // warps 0-3: thread t owns path t
constexpr int TC_HELP_WARPS = 4 * TC_NGRP;
constexpr int TC_EPI_WARPS = 4 * TC_EGRP;       // warps that arrive on a_help / a_chunk[0]
constexpr int TC_CTRL_WARP = 4 + TC_HELP_WARPS; // waits for operands, issues every tcgen05.mma
constexpr int TC_PROD_WARP = TC_CTRL_WARP + 1;  // lane 0: streams the weight chunks (bulk copies) on its own
constexpr int TC_WORK_THREADS = 32 * (TC_CTRL_WARP + 1);   // owner + helper + control warps take part in the in-loop CTA barriers (named barrier 1)
constexpr int TC_THREADS = TC_WORK_THREADS + 32;
// Register budgets per role (setmaxnreg, warpgroup-wide): the kernel is compiled for the launch bound (65536 / threads), the
// helper, control and producer warps hand registers back and the owner warps -- whose per-path state is what spills --
// take them.  DPB_TC_SETMAXNREG=0 compiles the hooks out.
#ifndef DPB_TC_SETMAXNREG
#define DPB_TC_SETMAXNREG 1
#endif
#ifndef DPB_TC_REGS_OWNER
#define DPB_TC_REGS_OWNER 216
#endif
#ifndef DPB_TC_REGS_HELPER
#define DPB_TC_REGS_HELPER 96
#endif
#define DPB_STR2(x) #x
#define DPB_STR(x) DPB_STR2(x)
__device__ __forceinline__ void tc_regs_owner() {
#if DPB_TC_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 " DPB_STR(DPB_TC_REGS_OWNER) ";");
#endif
}
__device__ __forceinline__ void tc_regs_helper() {
#if DPB_TC_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 " DPB_STR(DPB_TC_REGS_HELPER) ";");
#endif
}
__device__ __forceinline__ void tc_regs_small() {
#if DPB_TC_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
#endif
}
constexpr int MAX_NSLOT = 16;                   // ring slots (runtime count: whatever shared memory is left)
constexpr uint32_t COL_REG = 256;               // TMEM region r = columns [256 r, 256 r + 256)
constexpr int MAXOPS = 64;                      // schedule entries of one phase: <= 7 (L+1) products, L <= 6 hidden layers

constexpr int MAX_CHUNK = 16;                   // 16-column chunks of the widest layer (256 / 16): one a_chunk barrier each
constexpr int MAX_GROUP = 8;                    // weight chunks per ring slot (narrow layers)
// weight chunks streamed per ring slot for a product with R output rows (chunk = R*64 bytes)
__host__ __device__ __forceinline__ int tc_group(int R, int slot_bytes) {
    const int g = slot_bytes / (R * 64);
    return g < 1 ? 1 : (g > MAX_GROUP ? MAX_GROUP : g);
}

__host__ __device__ inline int round16(int x) { return (x + 15) & ~15; }

struct TcLayer {
    int kl, nl;            // logical dims of linear layer l
    int K16, N16;          // padded: K16 = round16(kl + 1) keeps one spare input feature (the constant 1 that
                           // turns the bias column sums into one more row of the dW product), N16 = K16 of
                           // the next layer, round16(out) for the last one
    long long img_f;       // byte offset of the forward image  (rows N16, contraction K16): K16/16 chunks of N16*64 B
                           // (hi plane N16*32 B, then lo plane)
    long long img_b;       // byte offset of the backward image (rows K16, contraction N16): N16/16 chunks of K16*64 B
    int vec;               // float offset of gc[N16], bb[N16] in the vector block
};

struct TcNet {
    int L, in, out, ekn_head, mctrl;
    TcLayer ly[MAXLIN];
    int vec_g0;            // g0c[K16_0], b0[K16_0]
    int vec_floats;
    long long img_bytes;
    NetDev flat;           // flat parameter layout
};

inline void tcnet_init(TcNet& t, int in, const int* hid, int L, int out, int ekn_head, int mctrl) {
    t.L = L; t.in = in; t.out = out; t.ekn_head = ekn_head; t.mctrl = mctrl;
    netdev_init(t.flat, in, hid, L, out, ekn_head, mctrl);
    long long img = 0;
    int vec = 0;
    int prev = in;
    t.vec_g0 = vec; vec += 2 * round16(in + 1);
    for (int l = 0; l <= L; ++l) {
        TcLayer& y = t.ly[l];
        y.kl = prev; y.nl = (l < L) ? hid[l] : out;
        y.K16 = round16(y.kl + 1);
        y.N16 = (l < L) ? round16(y.nl + 1) : round16(y.nl);
        y.img_f = img; img += (long long)y.K16 * y.N16 * 4;          // hi + lo planes, 2 B each
        y.img_b = img; img += (long long)y.K16 * y.N16 * 4;
        y.vec = vec; vec += 2 * y.N16;
        prev = y.nl;
    }
    t.img_bytes = (img + 255) & ~255LL;
    t.vec_floats = (vec + 63) & ~63;
}

inline bool tcnet_supported(const TcNet& t) {
    for (int l = 0; l <= t.L; ++l)
        if (t.ly[l].K16 > 256 || t.ly[l].N16 > 256) return false;
    return true;
}

// ------------------------------------------------------------------------------------------- pack
// flat FP32 parameters -> bf16 hi/lo operand images + the FP32 vector block (gamma*c, beta, ...)
// (nrep copies of the image, rep_stride bytes apart: CTA i streams copy i % nrep -- see TcArgs::img_rep)
static __global__ void tc_pack_kernel(TcNet t, const float* __restrict__ th, unsigned char* __restrict__ img, float* __restrict__ vec, float c,
                                      int nrep, long long rep_stride) {
    const NetDev& nd = t.flat;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int K0 = t.ly[0].K16;
    for (long long i = t0; i < K0; i += stride) {
        vec[t.vec_g0 + i] = i < t.in ? th[nd.fg0 + i] * c : 0.f;
        vec[t.vec_g0 + K0 + i] = i < t.in ? th[nd.fb0 + i] : (i == t.in ? 1.f : 0.f);      // (i = in: the constant-1 feature)
    }
    for (int l = 0; l <= t.L; ++l) {
        const TcLayer y = t.ly[l];
        for (long long n = t0; n < y.N16; n += stride) {
            float gc = n < y.nl ? th[nd.fg[l] + n] * c : 0.f;
            float bb = n < y.nl ? th[nd.fb[l] + n] : 0.f;
            if (l == t.L && n < y.nl) bb = th[nd.fbias + n] * gc + bb;
            vec[y.vec + n] = gc;
            vec[y.vec + y.N16 + n] = bb;
        }
        const long long tot = (long long)y.K16 * y.N16;
        for (long long i = t0; i < tot; i += stride) {
            const int k = (int)(i / y.N16), n = (int)(i - (long long)k * y.N16);
            const bool in_rng = (k < y.kl && n < y.nl);
            const float w = in_rng ? th[nd.fW[l] + (long long)k * y.nl + n] : 0.f;
            __nv_bfloat16 hi, lo;
            // forward image: rows n (N16), contraction k; chunk = k/16.  BatchNorm folded in (see the header): W gamma c, row kl =
            // beta (+ bias gamma c), entry (kl, nl) = 0.5 keeps the constant-1 feature alive through z + relu(z)
            float wf = 0.f;
            if (n < y.nl) {
                const float gcn = th[nd.fg[l] + n] * c;
                if (k < y.kl) wf = w * gcn;
                else if (k == y.kl) wf = (l == t.L) ? th[nd.fbias + n] * gcn + th[nd.fb[l] + n] : th[nd.fb[l] + n];
            } else if (n == y.nl && k == y.kl && l < t.L) {
                wf = 0.5f;
            }
            split_bf16(wf, hi, lo);
            {
                const long long cb = (long long)y.N16 * 64;
                unsigned char* base = img + y.img_f + (k >> 4) * cb + (((k & 15) >> 3) * (y.N16 >> 3) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 7) * 2;
                for (int r = 0; r < nrep; ++r) {
                    *reinterpret_cast<__nv_bfloat16*>(base + r * rep_stride) = hi;
                    *reinterpret_cast<__nv_bfloat16*>(base + r * rep_stride + y.N16 * 32) = lo;
                }
            }
            // backward image: rows k (K16), contraction n, value W[k][n] * gamma[n]*c; chunk = n/16
            const float wg = in_rng ? w * (th[nd.fg[l] + n] * c) : 0.f;
            split_bf16(wg, hi, lo);
            {
                const long long cb = (long long)y.K16 * 64;
                unsigned char* base = img + y.img_b + (n >> 4) * cb + (((n & 15) >> 3) * (y.K16 >> 3) + (k >> 3)) * 128 + (k & 7) * 16 + (n & 7) * 2;
                for (int r = 0; r < nrep; ++r) {
                    *reinterpret_cast<__nv_bfloat16*>(base + r * rep_stride) = hi;
                    *reinterpret_cast<__nv_bfloat16*>(base + r * rep_stride + y.K16 * 32) = lo;
                }
            }
        }
    }
}

// cycle-counter diagnostics (tools/time_forward.py): compiled in only with -DDPB_TC_STATS -- the counters live in
// per-thread local memory across the noinline helpers and cost ~6 % of the kernel when enabled
#ifdef DPB_TC_STATS
#define TC_STAT(...) __VA_ARGS__
// event trace of CTA 0 (owner thread 0, helper thread 128, control lane 0): (clock64 << 8) | event id, TC_TRACE_CAP events per
// role (dpb_tc_trace, tools/trace_timeline.py).  ids: control 1 operands of a TS product ready, 2 product issued and committed,
// 3 / 4 the same for a dW block, 5 product entered (before the operand wait); helpers 10 epilogue entered, 11 accumulator
// committed, 12 epilogue done, 13 / 14 / 15 the same for a drain; owners 20 published, 21 waits for a result, 22 has it
// (DPB_TC_TRACE_FINE: also 6 weight slot landed, 7 operand chunk published, 8 slot's MMAs issued and slot release committed)
#ifdef DPB_TC_TRACE_FINE
#define TC_TRACE2(ctx, id) TC_TRACE(ctx, id)
#else
#define TC_TRACE2(ctx, id)
#endif
#define TC_TRACE(ctx, id) do { if ((ctx).tr && (ctx).tn < TC_TRACE_CAP) (ctx).tr[(ctx).tn++] = ((unsigned long long)clock64() << 8) | (unsigned)(id); } while (0)
#else
#define TC_STAT(...)
#define TC_TRACE(ctx, id)
#define TC_TRACE2(ctx, id)
#endif

// ------------------------------------------------------------------------------ control-thread side
struct Sched {                       // cyclic schedule of ring-slot loads of the current phase (shared memory)
    const unsigned char* ptr[MAXOPS];
    int nld[MAXOPS];                 // loads of this product (each fills one ring slot with `lb` bytes, the last with `lb_last`)
    int lb[MAXOPS], lb_last[MAXOPS];
    int nops;
};

// mailbox between the control thread and the producer thread (shared memory)
struct ProdCtl {
    volatile uint32_t req;       // loads requested so far (monotone); the producer loads until loaded == req
    volatile uint32_t gen;       // bumped whenever a new schedule has been written (the producer restarts its cursor)
    volatile uint32_t quit;
};

// Hand-off barriers (shared memory, contiguous, 8 bytes each; index in brackets):
//   [0]      acc_full   count 1   tcgen05.commit of a product the HELPERS consume (hidden layers, dX of layers >= 1, dW blocks)
//   [1]      a_help     count = helper warps: an epilogue that publishes at once (EPI_ALL) / a dW accumulator has been drained
//   [2..17]  a_chunk[c] count 4   the four helper warps that converted chunk c published its planes (a_chunk[0]: every helper warp)
//   [18]     acc_fin    count 1   tcgen05.commit of a product the OWNERS read (a network's output, the input cotangent dy0;
//                                 also "the forward products are done" before a skip-last backward)
//   [19]     a_own      count 4   the owner warps wrote a network input (y0) or an output cotangent
//   [20]     acc_dw     count 1   tcgen05.commit of a dW block (drained by the helpers; its own barrier because a dW block is
//                                 committed right behind the dX product of its layer -- two commits the helpers have not yet looked
//                                 at would wrap the parity of a shared barrier)
// Every side tracks the parities of the barriers it waits on in one word, together with the TMEM region of the next A planes.
constexpr uint32_t SY_HELP = 1u << 16, SY_ACC = 1u << 17, SY_DZ = 1u << 18, SY_OWN = 1u << 19, SY_FIN = 1u << 20, SY_DW = 1u << 21, SY_REG = 1u << 31;
constexpr int BAR_ACC = 0, BAR_HELP = 1, BAR_CHUNK = 2, BAR_FIN = 18, BAR_OWN = 19, BAR_DW = 20, NUM_HANDOFF_BARS = 21;
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((127 + e) << 23); }            // 2^e, -126 <= e <= 127
enum { IN_OWN = 0, IN_HELP = 1, IN_CHUNKS = 2, IN_BOTH = 3, IN_NONE = 4 };     // what a product's inputs were published on (IN_NONE: already seen)

struct Ctrl {
    unsigned char* ring;
    uint64_t *full, *empty, *bars, *act_full;     // bars: the hand-off barriers above
    Sched* sch;
    ProdCtl* pc;
    uint32_t nslot, slot_bytes;
    uint32_t n_req, n_consumed, op_count, dw_count, act_count, tmem, gen;      // dw_count: dW blocks committed on acc_dw so far
    uint32_t sync;                   // parity bits (see above); op_count = products committed on acc_full so far
    volatile int* dexp;              // shared word: the exponent the owners scaled the current backward evaluation by
    uint32_t mm_slot, mm_use;        // ring cursor (slot index, wrap count) of the MMA issuer
    unsigned char *act, *dz;         // shared-memory operand images of the dW products (128 paths x K16 features, bf16)
    long long n_ops;
    TC_STAT(long long t_aready, t_issue, t_accw;)   // cycle counters (diagnostics)
    TC_STAT(long long t_dw_ready, t_act;)           // dW products: waiting for their operands / for the ACT bulk copy
    TC_STAT(unsigned long long* tr; int tn;)        // event trace (lane 0 of CTA 0 only)
};

// let the producer run up to nslot loads ahead of the MMAs (one shared-memory store, never waits)
__device__ __forceinline__ void ctrl_request(Ctrl& c) {
    const uint32_t want = c.n_consumed + c.nslot;
    if (want != c.n_req) { c.n_req = want; c.pc->req = want; }
}

// drop every load that was requested ahead but will not be used (end of a phase / dead tile)
__device__ __forceinline__ void ctrl_flush(Ctrl& c) {
    while (c.n_consumed != c.n_req) {
        mbar_wait(&c.full[c.mm_slot], c.mm_use & 1);
        if (elect_one()) mbar_arrive(&c.empty[c.mm_slot]);        // (elect.sync: every lane has seen the phase complete)
        ++c.n_consumed;
        if (++c.mm_slot == c.nslot) { c.mm_slot = 0; ++c.mm_use; }
    }
    if ((threadIdx.x & 31) == 0) c.sch->nops = 0;
    __syncwarp();
}
// call after the schedule of a phase has been written (the producer is idle: everything requested was flushed)
__device__ __forceinline__ void ctrl_sched_ready(Ctrl& c) {
    __syncwarp();                                                 // lane 0 wrote the schedule
    __threadfence_block();
    c.pc->gen = ++c.gen;
    __threadfence_block();
}

// the producer thread: streams the cyclic load schedule through the ring as far as requested
__device__ __forceinline__ void producer_loop(unsigned char* ring, uint64_t* full, uint64_t* empty, Sched* sch, ProdCtl* pc,
                                              uint32_t nslot, uint32_t slot_bytes) {
    uint32_t loaded = 0, my_gen = 0, slot = 0, use = 0;
    int op = 0, ld = 0, nops = 0, cur_nld = 0;
    uint32_t cur_lb = 0, cur_last = 0;
    const unsigned char* cur_ptr = nullptr;
    for (;;) {
        const uint32_t r = pc->req;
        if (loaded == r) {
            if (pc->quit) break;
            __nanosleep(32);
            continue;
        }
        __threadfence_block();
        const uint32_t g = pc->gen;
        if (g != my_gen) {
            my_gen = g; op = 0; ld = 0; nops = sch->nops;
            cur_ptr = sch->ptr[0]; cur_nld = sch->nld[0]; cur_lb = (uint32_t)sch->lb[0]; cur_last = (uint32_t)sch->lb_last[0];
        }
        while (loaded != r) {
            if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
            const uint32_t bytes = (ld == cur_nld - 1) ? cur_last : cur_lb;
            mbar_arrive_expect_tx(&full[slot], bytes);
            bulk_g2s(ring + (size_t)slot * slot_bytes, cur_ptr + (size_t)ld * cur_lb, bytes, &full[slot]);
            if (++ld == cur_nld) {
                ld = 0;
                if (++op == nops) op = 0;
                cur_ptr = sch->ptr[op]; cur_nld = sch->nld[op]; cur_lb = (uint32_t)sch->lb[op]; cur_last = (uint32_t)sch->lb_last[op];
            }
            ++loaded;
            if (++slot == nslot) { slot = 0; ++use; }
        }
    }
}

// one product with nch_in contraction chunks of R_out rows each: ceil(nch_in / g) loads of g chunks
__device__ __forceinline__ void sched_add_op(Sched* s, const unsigned char* img, int nch_in, int R_out, int slot_bytes) {
    if ((threadIdx.x & 31) == 0) {
        if (s->nops >= MAXOPS) asm volatile("trap;");                 // (dpb_create rejects networks whose phases do not fit)
        const int g = tc_group(R_out, slot_bytes), nld = (nch_in + g - 1) / g, i = s->nops;
        s->ptr[i] = img; s->nld[i] = nld; s->lb[i] = g * R_out * 64; s->lb_last[i] = (nch_in - (nld - 1) * g) * R_out * 64;
        s->nops = i + 1;
    }
}
__device__ __forceinline__ void sched_add_fwd(Ctrl& c, const TcNet& t, const unsigned char* img, int upto /*layers 0..upto*/) {
    for (int l = 0; l <= upto; ++l) sched_add_op(c.sch, img + t.ly[l].img_f, t.ly[l].K16 / 16, t.ly[l].N16, (int)c.slot_bytes);
}

// wait for the inputs of a product (see IN_*), flipping the tracked parities
__device__ __forceinline__ void ctrl_wait_inputs(uint32_t bars0, uint32_t& sync, int in_kind) {
    if (in_kind == IN_OWN || in_kind == IN_BOTH) { mbar_wait_u32(bars0 + 8 * BAR_OWN, (sync >> 19) & 1u); sync ^= SY_OWN; }
    if (in_kind == IN_HELP || in_kind == IN_BOTH) { mbar_wait_u32(bars0 + 8 * BAR_HELP, (sync >> 16) & 1u); sync ^= SY_HELP; }
    if (in_kind != IN_CHUNKS && in_kind != IN_NONE) tc_fence_after();
}

// D[acc region] = A(planes in the other region) x B(streamed image with R rows): nchunks contraction chunks, 3 split
// products per chunk.  in_kind: where the A planes were published (IN_CHUNKS: chunk by chunk by an in-place epilogue).
// fin: the owners read the result (commit on acc_fin), otherwise the helpers convert it (acc_full); also_fin: commit on
// both.  toggles: the epilogue converts the accumulator in place into the planes of the next product (the regions swap).
__device__ __forceinline__ void ctrl_gemm_ts(Ctrl& cref, int nchunks_, int R_, int in_kind_, bool fin_, bool also_fin_, bool toggles) {
    Ctrl c = cref;                                                       // registers for the issue loop
    // everything an MMA / commit operand is computed from goes through a lane-0 broadcast: the compiler then knows the
    // values are warp-uniform (the state lives in per-thread local memory across the noinline callers)
    const int nchunks = (int)warp_uniform((uint32_t)nchunks_), R = (int)warp_uniform((uint32_t)R_);
    const int in_kind = (int)warp_uniform((uint32_t)in_kind_);
    const bool chunked = in_kind == IN_CHUNKS, fin = warp_uniform(fin_ ? 1u : 0u) != 0, also_fin = warp_uniform(also_fin_ ? 1u : 0u) != 0;
    const uint32_t tmem = warp_uniform(c.tmem), nslot = warp_uniform(c.nslot), slot_bytes = warp_uniform(c.slot_bytes);
    const uint32_t ring0 = warp_uniform(smem_u32(c.ring)), full0 = warp_uniform(smem_u32(c.full)), empty0 = warp_uniform(smem_u32(c.empty));
    const uint32_t bars0 = warp_uniform(smem_u32(c.bars));
    uint32_t mm_slot = warp_uniform(c.mm_slot), mm_use = warp_uniform(c.mm_use);
    uint32_t sync = warp_uniform(c.sync);
    const uint32_t reg = sync >> 31;
    const uint32_t t_in = tmem + COL_REG * reg, t_out = tmem + COL_REG * (reg ^ 1u);
    const int g = tc_group(R, (int)slot_bytes);
    const uint32_t idesc = idesc_bf16(128, R, 0, 0);
    const uint32_t lbo = (R >> 3) * 128;
    const uint64_t dlo = ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    ctrl_request(c);
    TC_STAT(const long long t0 = clock64();)
    TC_TRACE(c, 5);
    ctrl_wait_inputs(bars0, sync, in_kind);
    TC_STAT(const long long ti0 = clock64(); c.t_aready += ti0 - t0;)
    TC_TRACE(c, 1);
    ++c.n_ops;
    // Issue loop (the whole warp runs it on warp-uniform values; the elected lane issues).  Per slot: wait for its weights (ring
    // slot landed), per chunk for its operand planes, three MMAs, the commit that releases the slot.  Round 2 tried three other
    // shapes of this loop, all measured slower or equal at 2^17 paths (profiles/r02_issue_loop_variants.txt):
    // polling all barriers at once from the 32 lanes (an mbarrier test costs ~80 cycles PER DISTINCT ADDRESS of the warp
    // instruction), byte counters in shared memory instead of the chunk barriers, and testing the next slot's / chunk's barrier
    // under the MMAs of the current one (mma_ts3_commit_wait2 in dpb_tc.cuh: no change -- with the latencies hidden the MMAs
    // themselves still need ~520 cycles per slot next to a running epilogue).
    for (int s0 = 0; s0 < nchunks; s0 += g) {
        mbar_wait_u32(full0 + mm_slot * 8, mm_use & 1);
        TC_TRACE2(c, 6);
        const uint32_t sb0 = ring0 + mm_slot * slot_bytes;
        const int n = (nchunks - s0) < g ? (nchunks - s0) : g;
        for (int j = 0; j < n; ++j) {
            const int s = s0 + j;
            if (chunked) {
                TC_STAT(const long long tw = clock64();)
                mbar_wait_u32(bars0 + 8 * (BAR_CHUNK + s), (sync >> s) & 1u);
                TC_STAT(c.t_aready += clock64() - tw;)
                sync ^= 1u << s;
                tc_fence_after();
                TC_TRACE2(c, 7);
            }
            const uint32_t sb = sb0 + j * R * 64;
            const uint64_t bhi = dlo | (uint64_t)((sb >> 4) & 0x3FFF), blo = dlo | (uint64_t)(((sb + R * 32) >> 4) & 0x3FFF);
            const uint32_t ahi = t_in + s * 16, alo = t_in + s * 16 + 8;
            if (elect_one()) {
                mma_ts(t_out, ahi, bhi, idesc, s > 0);
                mma_ts(t_out, ahi, blo, idesc, 1);
                mma_ts(t_out, alo, bhi, idesc, 1);
            }
        }
        if (elect_one()) tc_commit_u32(empty0 + mm_slot * 8);
        TC_TRACE2(c, 8);
        ++c.n_consumed;
        if (++mm_slot == nslot) { mm_slot = 0; ++mm_use; }
        ctrl_request(c);
    }
    if (elect_one()) {
        if (!fin) tc_commit_u32(bars0 + 8 * BAR_ACC);
        if (fin || also_fin) tc_commit_u32(bars0 + 8 * BAR_FIN);
    }
    TC_STAT(c.t_issue += clock64() - ti0;)
    TC_TRACE(c, 2);
    if (!fin) ++c.op_count;
    if (toggles) sync ^= SY_REG;
    c.sync = sync;
    c.mm_slot = mm_slot; c.mm_use = mm_use;
    cref = c;
}

// forward products of layers 0..upto: the first reads the planes the owners wrote (y0), the others follow an in-place
// epilogue; the output of layer L goes to the owners.  upto = L-1 (the backward starts from the last hidden layer): its
// accumulator is not converted into planes (no region swap) and the owners are told when the products are done.
static __device__ __noinline__ void ctrl_net_forward(Ctrl& c, const TcNet& t, int upto) {
    for (int l = 0; l <= upto; ++l) {
        const bool last_of_skip = (l == upto && upto < t.L);
        ctrl_gemm_ts(c, t.ly[l].K16 / 16, t.ly[l].N16, l > 0 ? IN_CHUNKS : IN_OWN, l == t.L, last_of_skip, l < t.L && !last_of_skip);
    }
}

// ---------------------------------------------------------------------------------- owner / helper side
struct PathCtx {
    uint32_t tl;                   // TMEM address of this thread's lane (column 0)
    uint32_t bars;                 // shared-memory address of the hand-off barriers (BAR_*)
    uint32_t sync;                 // parity bits of the barriers this side waits on + TMEM region bit (SY_*)
    int grp;                       // helpers: group (chunk c belongs to group c % TC_NGRP); owners: 0
    int dexp;                      // owners: the current backward evaluation runs scaled by 2^dexp (see the header: FP16 dW operands)
    TC_STAT(long long t_accw, t_mark;)      // cycles spent waiting for the tensor pipe; time of the last wake-up
    TC_STAT(long long t_epi, t_hid;)        // t_hid: cycles inside hidden-layer epilogues only
    TC_STAT(long long t_drain;)             // cycles inside the dW drains (after the accumulator wait)
    TC_STAT(unsigned long long* tr; int tn;)  // event trace (owner thread 0 / helper thread 128 of CTA 0 only)
};

// The out-of-line helpers take the context BY VALUE (registers) and return the one field they change: passed by reference it
// lives in local memory, and with 28-60 KB of L1 every epilogue then starts with a chain of L2-latency loads.  (The
// cycle-counter build keeps the reference: its counters are updated inside the helpers.)
#ifdef DPB_TC_STATS
typedef PathCtx& PathArg;
#else
typedef PathCtx PathArg;
#endif

__device__ __forceinline__ uint32_t path_planes(const PathCtx& p) { return p.tl + COL_REG * (p.sync >> 31); }             // A planes of the next product
__device__ __forceinline__ uint32_t path_acc(const PathCtx& p) { return p.tl + COL_REG * ((p.sync >> 31) ^ 1u); }         // its accumulator

// ---- helpers ------------------------------------------------------------------------------------------------------
// an epilogue that publishes at once / a drained accumulator: one arrival per helper warp on a_help
__device__ __forceinline__ void help_publish(PathCtx& p) {
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_u32(p.bars + 8 * BAR_HELP);
    p.sync ^= SY_HELP;
    TC_STAT(p.t_epi += clock64() - p.t_mark;)
}
// planes of chunk c of an in-place epilogue are written: one arrival per warp of the group that owns the chunk
__device__ __forceinline__ void help_publish_chunk(const PathCtx& p, int c) {
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_u32(p.bars + 8 * (BAR_CHUNK + c));
}
// wait for the commit of the current helper-consumed product
__device__ __forceinline__ void help_wait_acc(PathCtx& p) {
    TC_STAT(const long long t0 = clock64();)
    mbar_wait_u32(p.bars + 8 * BAR_ACC, (p.sync >> 17) & 1u);
    TC_STAT(p.t_mark = clock64(); p.t_accw += p.t_mark - t0;)
    p.sync ^= SY_ACC;
    tc_fence_after();
}

// Packed FP32 pairs (sm_100: FFMA2 / FADD2, one instruction for two IEEE operations -- same bits as the scalar forms).
// The epilogues are bound by the instructions they issue, so every pairwise step is packed.
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// 16 floats -> 8 packed hi words + 8 packed lo words (hi = bf16(x), lo = bf16(x - hi); element 2j in bits 0..15)
__device__ __forceinline__ void split16(const float* v, uint32_t* h, uint32_t* l) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hh);
        float r0, r1;
        up2(sub2(pk2(v[2 * j], v[2 * j + 1]), pk2(__uint_as_float(hb << 16), __uint_as_float(hb & 0xffff0000u))), r0, r1);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(r0, r1);
        h[j] = hb;
        l[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
}
// 16 features of this thread's row -> hi / lo planes of their chunk: tc = TMEM address of the chunk's 16 columns
__device__ __forceinline__ void put16(uint32_t tc, const float* v) {
    uint32_t h[8], l[8];
    split16(v, h, l);
    tmem_st8(tc, h);
    tmem_st8(tc + 8, l);
}
// In-place epilogue of a product with `nco` output chunks (helpers): wait for its commit, hand the chunks of this thread's
// group (c % TC_NGRP == grp) to f(c, r[16], tc) -- r = the 16 accumulator columns, tc = their TMEM address, where f stores
// the planes of the next product.  mode: EPI_CHUNKS publishes every chunk on its own barrier (the next product starts on
// it at once); EPI_ALL publishes once at the end on a_help (the next product is a dW product, which needs everything).
// fence: what f wrote besides TMEM and the async proxy reads later (dW operands, bulk copies): 0 nothing, 1 shared memory,
// 2 global memory; the proxy fence precedes the thread's LAST arrival of the epilogue, and whoever consumes those bytes
// has waited for every chunk (or for a_help).  toggle: the accumulator became the next planes (the regions swap roles).
enum { EPI_CHUNKS = 1, EPI_ALL = 2 };
template <class F>
__device__ __forceinline__ void for_acc_chunks(PathCtx& p, int nco, int mode, int fence, bool toggle, F f) {
    // (the TMEM->register path is the bound of every epilogue -- 64 B/clk/SM, see DESIGN.md -- so a plain loop does as
    //  well as a software-pipelined one and needs 16 registers fewer)
    TC_TRACE(p, 10);
    help_wait_acc(p);
    TC_TRACE(p, 11);
    // Every helper warp takes part in the hand-off of chunk 0 (its owners when the planes are written, the others here, as
    // soon as they have seen the commit): the next product cannot be committed before all warps have observed this one.
    // Without it a warp that owns no chunk of a narrow output (one chunk: the other group's) could fall two phases behind
    // on acc_full and wait for a parity that has already come round again.
    if (TC_EGRP > 1 && mode == EPI_CHUNKS && p.grp != 0) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive_u32(p.bars + 8 * BAR_CHUNK);
    }
    const uint32_t acc = path_acc(p);
    uint32_t ra[16];
    int pend = -1;                                      // chunk whose stores are in flight (published after the next load)
    for (int c = p.grp; c < nco; c += TC_EGRP) {
        tmem_ld16(acc + 16 * c, ra);
        tmem_ld_wait();
        if (pend >= 0) help_publish_chunk(p, pend);     // (its stores completed under the latency of the load)
        f(c, ra, acc + 16 * c);
        if (mode == EPI_CHUNKS) {
            if (c < TC_EGRP) {                                                // first chunk: at once -- the next product starts on it
                if (fence == 1) fence_proxy_async(); else if (fence == 2) fence_proxy_async_global();     // (it may be this thread's only arrival)
                help_publish_chunk(p, c);
                pend = -1;
            } else {
                pend = c;
            }
        }
    }
    if (fence == 1) fence_proxy_async(); else if (fence == 2) fence_proxy_async_global();
    if (pend >= 0) help_publish_chunk(p, pend);
    if (mode == EPI_CHUNKS) p.sync ^= (1u << nco) - 1u;                    // every a_chunk[c], c < nco, completed a phase
    if (toggle) p.sync ^= SY_REG;
    if (mode == EPI_ALL) help_publish(p);
    TC_STAT(else p.t_epi += clock64() - p.t_mark;)
    TC_TRACE(p, 12);
}

// z = acc * gc + bb for 16 features (gc, bb: 16-byte aligned SHARED memory; read with ld.shared -- through generic pointers the
// compiler emitted eight generic 128-bit loads per chunk)
__device__ __forceinline__ float4 lds4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void affine16(const uint32_t* r, const float* gc, const float* bb, float* z) {
    const uint32_t ga = smem_u32(gc), ba = smem_u32(bb);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 g = lds4(ga + 16 * q), b = lds4(ba + 16 * q);
        up2(fma2(pk2(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])), pk2(g.x, g.y), pk2(b.x, b.y)), z[4 * q], z[4 * q + 1]);
        up2(fma2(pk2(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])), pk2(g.z, g.w), pk2(b.z, b.w)), z[4 * q + 2], z[4 * q + 3]);
    }
}

// 16 features 16c.. of this thread's row -> FP16 operand image of the dW products (R = 128 rows) at `img` (shared or global)
__device__ __forceinline__ void copy16f(unsigned char* img, int row, int c, const float* v, int one_at /* feature index set to 1, or -1 */) {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)                                      // (saturating: a wild value becomes +-65504, not inf)
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
    const int o = one_at - 16 * c;
    if (o >= 0 && o < 16) {                                          // fp16(1.0) = 0x3C00
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (o == 2 * j) w[j] = (w[j] & 0xffff0000u) | 0x3C00u;
            if (o == 2 * j + 1) w[j] = (w[j] & 0x0000ffffu) | 0x3C000000u;
        }
    }
    unsigned char* p = img + (size_t)(2 * c) * 2048 + (row >> 3) * 128 + (row & 7) * 16;
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(p + 2048) = make_uint4(w[4], w[5], w[6], w[7]);
}

// hidden layer epilogue: a = z + relu(z), z = acc * gc + bb (solver.py:267-269) -> planes
__device__ __forceinline__ void help_epi_hidden(PathCtx& p, const float* gcbb, int N16) {
    TC_STAT(const long long th0 = clock64();)
    (void)gcbb;                                                           // (BatchNorm is folded into the product: the accumulator is z)
    for_acc_chunks(p, N16 / 16, EPI_CHUNKS, 0, true, [&](int c, const uint32_t* r, uint32_t tc) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            up2(add2(pk2(v[2 * j], v[2 * j + 1]), pk2(fmaxf(v[2 * j], 0.f), fmaxf(v[2 * j + 1], 0.f))), v[2 * j], v[2 * j + 1]);
        put16(tc, v);
    });
    TC_STAT(p.t_hid += clock64() - th0;)
}
// the hidden-layer epilogues l0 .. l1-1 of one forward-only evaluation (helpers; in combined mode the owner places its own
// per-path arithmetic between two ranges: it then runs while the tensor pipe works on the layer just published)
static __device__ __noinline__ uint32_t help_forward_(PathArg p, const TcNet& t, const float* vec, int l0, int l1) {
    for (int l = l0; l < l1 && l < t.L; ++l) help_epi_hidden(p, vec + t.ly[l].vec, t.ly[l].N16);
    return p.sync;
}
__device__ __forceinline__ void help_forward(PathCtx& p, const TcNet& t, const float* vec, int l0 = 0, int l1 = MAXLIN) { p.sync = help_forward_(p, t, vec, l0, l1); }

// ---- owners -------------------------------------------------------------------------------------------------------
// a network input / output cotangent is in place: one arrival per owner warp on a_own
__device__ __forceinline__ void own_publish(PathCtx& p) {
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_u32(p.bars + 8 * BAR_OWN);
    TC_STAT(p.t_epi += clock64() - p.t_mark;)
    TC_TRACE(p, 20);
}
// wait for a product whose result the owners read (or for the end of the forward products of a skip-last evaluation)
__device__ __forceinline__ void own_wait_fin(PathCtx& p) {
    TC_STAT(const long long t0 = clock64();)
    TC_TRACE(p, 21);
    mbar_wait_u32(p.bars + 8 * BAR_FIN, (p.sync >> 20) & 1u);
    TC_STAT(p.t_mark = clock64(); p.t_accw += p.t_mark - t0;)
    TC_TRACE(p, 22);
    p.sync ^= SY_FIN;
    tc_fence_after();
}
// the helpers converted `n` accumulators in place since the owners last touched tensor memory: the regions swapped n times
// (combined mode: the owner ran those conversions itself and its region bit is already up to date)
__device__ __forceinline__ void own_swaps(PathCtx& p, int n) { if (!TC_COMBINED && (n & 1)) p.sync ^= SY_REG; }

// y0 = x * g0c + b0 (solver.py:265) -> planes (+ FP16 copy with the constant-1 feature when `copies`), then
// publish.  Everything indexed statically (K16_0 <= 32) so that x can live in registers.
// y0 = x * g0c + b0 (input BatchNorm; feature `in` is the constant 1: g0c = 0, b0 = 1 there) as hi / lo plane words.  Split
// from the stores so that the owners can prepare the next input while the networks still run and only have to store and
// publish it once the plane region is free (critic rollout).
template <int NX>
__device__ __forceinline__ void own_prep_y0_chunk(const TcNet& t, const float* vec, const float (&x)[NX], int c, uint32_t (&h)[8], uint32_t (&l)[8]) {
    const int K0 = t.ly[0].K16;
    const uint32_t ga = smem_u32(vec + t.vec_g0), ba = ga + 4u * (uint32_t)K0;
    float v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 g = lds4(ga + 64 * c + 16 * q), b = lds4(ba + 64 * c + 16 * q);
        const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = 16 * c + 4 * q + i;
            const float xv = (k < NX && k < t.in) ? x[k < NX ? k : 0] : 0.f;       // (entries past d are not initialised)
            v[4 * q + i] = xv * gg[i] + bb[i];
        }
    }
    split16(v, h, l);
}
template <int NX>
__device__ __forceinline__ void own_prep_y0(const TcNet& t, const float* vec, const float (&x)[NX], uint32_t (&yh)[2][8], uint32_t (&yl)[2][8]) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
        if (c < t.ly[0].K16 / 16) own_prep_y0_chunk(t, vec, x, c, yh[c], yl[c]);
}
__device__ __forceinline__ void own_store_y0(PathCtx& p, const TcNet& t, const uint32_t (&yh)[2][8], const uint32_t (&yl)[2][8]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c < t.ly[0].K16 / 16) {
            tmem_st8(path_planes(p) + 16 * c, yh[c]);
            tmem_st8(path_planes(p) + 16 * c + 8, yl[c]);
        }
    }
}
// y0 -> planes (+ the FP16 copy of the RAW input for the dW product of layer 0), publish
template <int NX>
__device__ __forceinline__ void own_put_y0(PathCtx& p, const TcNet& t, const float* vec, const float (&x)[NX], unsigned char* copies, int row) {
    const int K0 = t.ly[0].K16;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c < K0 / 16) {
            uint32_t h[8], l[8];
            own_prep_y0_chunk(t, vec, x, c, h, l);
            tmem_st8(path_planes(p) + 16 * c, h);
            tmem_st8(path_planes(p) + 16 * c + 8, l);
        }
    }
    if (copies) {
        // the dW operand of layer 0 is the RAW input x (+ the constant 1): x^T dz_0 gives the weight gradient AND the
        // gradients of the input BatchNorm (finalize kernel), so no per-thread input sums are kept
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (c < K0 / 16) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int k = 16 * c + j;
                    v[j] = (k < NX && k < t.in) ? x[k < NX ? k : 0] : 0.f;
                }
                copy16f(copies, row, c, v, t.ly[0].kl);
            }
        }
        fence_proxy_async_global();
    }
    own_publish(p);
}

// last layer: out[n] = acc * gc + bb (solver.py:270-271), n < nl <= 32; static indexing.  The helpers converted the L
// hidden accumulators in the meantime.
template <int NO>
__device__ __forceinline__ void own_last(PathCtx& p, const TcNet& t, const float* vec, float (&out)[NO]) {
    own_wait_fin(p);
    own_swaps(p, t.L);
    const uint32_t acc = path_acc(p);
    const int N16 = t.ly[t.L].N16, nl = t.ly[t.L].nl;
    (void)vec;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c < N16 / 16) {
            uint32_t r[16];
            tmem_ld16(acc + 16 * c, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = 16 * c + j;
                if (n < NO && n < nl) out[n < NO ? n : 0] = __uint_as_float(r[j]);        // (affine map and bias folded into the product)
            }
        }
    }
}
// one whole forward-only evaluation as the owners see it (combined mode: including the hidden-layer epilogues)
template <int NX, int NO>
__device__ __forceinline__ void own_net_forward(PathCtx& p, const TcNet& t, const float* vec, const float (&x)[NX], float (&out)[NO]) {
    own_put_y0(p, t, vec, x, nullptr, 0);
    if (TC_COMBINED) help_forward(p, t, vec);
    own_last(p, t, vec, out);
}

}  // namespace tc
}  // namespace dpb
namespace dpb {
namespace tc {

// ===================================================================================== backward pass
// Gradient slab of one network on the tensor path (per CTA, FP32):
//   layer l: (kl + 1) rows x nl columns; rows 0..kl-1 = G_l = a_l^T dz_l, row kl = column sums of dz_l (the
//   activation copies carry a constant 1 in feature kl), stored as [column group of 4][row][4] (tcslab_idx) so that the
//   32 lanes of a warp -- 32 consecutive rows of the accumulator -- reduce 512 contiguous bytes per instruction;
//   layer 0 is stored for the RAW input: rows k < in = GX = x^T dz_0, row `in` = column sums C of dz_0; the finalize kernel forms
//   y0^T dz_0 = g0c_k GX + b0_k C from it, and the input-BatchNorm gradients SX_k = sum_p x_k dy0_k = sum_n W gc GX[k][n],
//   S0_k = sum_p dy0_k = sum_n W gc C[n].  (gX / g0 below are no longer used.)
struct TcSlab { long long gW[MAXLIN], gX, g0, gtotal; };
__host__ __device__ __forceinline__ long long tcslab_idx(int k, int n, int kl) { return ((long long)(n >> 2) * (kl + 1) + k) * 4 + (n & 3); }
inline void tcslab_init(TcSlab& g, const TcNet& t) {
    long long o = 0;
    for (int l = 0; l <= t.L; ++l) { g.gW[l] = o; o += (long long)(t.ly[l].kl + 1) * ((t.ly[l].nl + 3) & ~3); }
    g.gX = o; o += (t.in + 3) & ~3; g.g0 = o; o += (t.in + 3) & ~3;
    g.gtotal = o;
}
// byte offset of the bf16 copy of activation a_l (l = 0..L-1) inside the per-CTA copy scratch
inline __host__ __device__ long long tc_copy_off(const TcNet& t, int l) {
    long long o = 0;
    for (int i = 0; i < l; ++i) o += (long long)TC_PATHS * t.ly[i].K16 * 2;
    return o;
}
inline __host__ __device__ long long tc_copy_bytes(const TcNet& t) { return tc_copy_off(t, t.L); }

// raw slab -> flat gradient (same formulas as finalize_grad_kernel of the exact path)
static __global__ void tc_finalize_grad_kernel(TcNet t, TcSlab g, const float* __restrict__ th, const float* __restrict__ raw, float* __restrict__ grad, float c) {
    const NetDev& nd = t.flat;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    {   // input BatchNorm (solver.py:265): d gamma0_k = c * sum_p x_k dy0_k, d beta0_k = sum_p dy0_k, dy0 = dz_0 (W_0 gamma_1 c)^T
        const int kl = t.ly[0].kl, nl = t.ly[0].nl;
        for (long long k = t0; k < t.in; k += stride) {
            float sx = 0.f, s0 = 0.f;
            for (int n = 0; n < nl; ++n) {
                const float wg = th[nd.fW[0] + k * nl + n] * (th[nd.fg[0] + n] * c);
                sx = fmaf(wg, raw[g.gW[0] + tcslab_idx((int)k, n, kl)], sx);
                s0 = fmaf(wg, raw[g.gW[0] + tcslab_idx(kl, n, kl)], s0);
            }
            grad[nd.fg0 + k] = c * sx;
            grad[nd.fb0 + k] = s0;
        }
    }
    for (int l = 0; l <= t.L; ++l) {
        const int kl = t.ly[l].kl, nl = t.ly[l].nl;
        // G[k][n] = (a_l^T dz_l)[k][n]; for layer 0 the slab holds x^T dz_0: y0 = x g0c + b0
        auto G = [&](int k, int n) -> float {
            const float r = raw[g.gW[l] + tcslab_idx(k, n, kl)];
            if (l > 0) return r;
            return fmaf(th[nd.fg0 + k] * c, r, th[nd.fb0 + k] * raw[g.gW[0] + tcslab_idx(kl, n, kl)]);
        };
        for (long long i = t0; i < (long long)kl * nl; i += stride) {
            int n = (int)(i % nl);
            grad[nd.fW[l] + i] = G((int)(i / nl), n) * (th[nd.fg[l] + n] * c);
        }
        for (long long n = t0; n < nl; n += stride) {
            float s = 0.f;
            for (int k = 0; k < kl; ++k) s = fmaf(th[nd.fW[l] + (long long)k * nl + n], G(k, (int)n), s);
            const float C = raw[g.gW[l] + tcslab_idx(kl, n, kl)];
            if (l == t.L) {
                s = s + th[nd.fbias + n] * C;
                grad[nd.fbias + n] = (th[nd.fg[l] + n] * c) * C;
            }
            grad[nd.fg[l] + n] = c * s;
            grad[nd.fb[l] + n] = C;
        }
    }
}

// ---- control thread ----------------------------------------------------------------------------------
__device__ __forceinline__ void sched_add_bwd(Ctrl& c, const TcNet& t, const unsigned char* img, bool need_dy0 = true) {
    for (int l = t.L; l >= (need_dy0 ? 0 : 1); --l) sched_add_op(c.sch, img + t.ly[l].img_b, t.ly[l].N16 / 16, t.ly[l].K16, (int)c.slot_bytes);
}

// global copy scratch -> ACT (bulk copy; waits until it has landed)
__device__ __forceinline__ void ctrl_act_load(Ctrl& c, const unsigned char* src, uint32_t bytes) {
    if (elect_one()) {
        mbar_arrive_expect_tx(c.act_full, bytes);
        bulk_g2s(c.act, src, bytes, c.act_full);
    }
}
__device__ __forceinline__ void ctrl_act_wait(Ctrl& c) {
    TC_STAT(const long long t0 = clock64();)
    mbar_wait(c.act_full, c.act_count & 1);
    TC_STAT(c.t_act += clock64() - t0;)
    ++c.act_count;
}

// D[acc] (rows = features 128*blk .. of ACT, cols = N16 features of DZ) = ACT^T DZ over the 128 paths.  The accumulator is
// the region the NEXT TS product will write: issued behind the dX product of its layer (layers >= 1) that is the region
// whose dz planes that product has just consumed; issued before it (layer 0) the region the dX product will write.
__device__ __forceinline__ void ctrl_gemm_dw(Ctrl& c, int blk_, int N16_, int in_kind_, bool also_fin_ = false) {
    const bool also_fin = warp_uniform(also_fin_ ? 1u : 0u) != 0;
    const int blk = (int)warp_uniform((uint32_t)blk_), N16 = (int)warp_uniform((uint32_t)N16_), in_kind = (int)warp_uniform((uint32_t)in_kind_);
    const uint32_t idesc = idesc_f16(128, N16, 1, 1);                   // FP16 operand images
    uint32_t sync = warp_uniform(c.sync);
    const uint32_t bars0 = warp_uniform(smem_u32(c.bars));
    TC_STAT(const long long t0 = clock64();)
    TC_TRACE(c, 5);
    ctrl_wait_inputs(bars0, sync, in_kind);                             // operands complete / previous block drained
    TC_STAT(c.t_dw_ready += clock64() - t0;)
    TC_TRACE(c, 3);
    const uint32_t a0 = warp_uniform(smem_u32(c.act)) + blk * 16 * 2048, b0 = warp_uniform(smem_u32(c.dz));
    const uint32_t tmem = warp_uniform(c.tmem);
    const uint32_t dcol = tmem + COL_REG * ((sync >> 31) ^ 1u);
    if (elect_one()) {
#pragma unroll
        for (int s = 0; s < TC_PATHS / 16; ++s) {
            const uint64_t ad = smem_desc(a0 + s * 256, 128, 2048), bd = smem_desc(b0 + s * 256, 128, 2048);
            mma_ss(dcol, ad, bd, idesc, s > 0);
        }
        tc_commit_u32(bars0 + 8 * BAR_DW);
        if (also_fin) tc_commit_u32(bars0 + 8 * BAR_FIN);
    }
    TC_TRACE(c, 4);
    c.sync = sync;
    ++c.dw_count;
}

// backward of one network evaluation.  need_w: dW products (+ drains on the helper side);
// copies: per-CTA global scratch holding the FP16 copies of a_0..a_{L-1} (a_L is already in ACT);
// skip_last: the forward stopped at the last hidden layer, whose epilogue published on a_help (the output cotangent
// comes from the owners either way).
// need_dy0 = false (the critic's networks: nothing consumes the cotangent of their input): the product dX of layer 0 is
// left out; the owners are told when the last dW product is done (from then on the copy of a_0 and the plane region are
// theirs again) and the control warp itself waits for the last drain.
// Order within a layer l >= 1 (round 2): dX FIRST -- it needs only the dz planes, which the epilogue of layer l+1 hands over
// chunk by chunk, so the tensor pipe works on dA_l while the helpers still convert dA_{l+1} -- then the dW blocks of the
// layer, accumulated in the region whose planes dX has just consumed (DZ still holds dz_l: the helpers write dz_{l-1} into
// it only after they have drained those blocks).  Layer 0 keeps dW before dX: its dX result goes to the owners, who reuse
// the plane region at once.
static __device__ __noinline__ void ctrl_net_backward(Ctrl& c, const TcNet& t, bool need_w, const unsigned char* copies, bool skip_last, bool need_dy0) {
    for (int l = t.L; l >= 1; --l) {
        // dA_l = dz_l x (W_l gamma c)^T.  (skip_last: its accumulator is the one the helpers read a_L from -- wait for them too)
        ctrl_gemm_ts(c, t.ly[l].N16 / 16, t.ly[l].K16, l == t.L ? (skip_last ? IN_BOTH : IN_OWN) : IN_CHUNKS, false, false, true);
        if (need_w) {
            if (l < t.L) ctrl_act_wait(c);
            const int nblk = (t.ly[l].kl + 1 + 127) >> 7;
            for (int b = 0; b < nblk; ++b)
                ctrl_gemm_dw(c, b, t.ly[l].N16, b > 0 ? IN_HELP : IN_NONE);
            TC_STAT(const long long t0 = clock64();)
            mbar_wait(&c.bars[BAR_DW], (c.dw_count - 1) & 1);            // the MMAs reading ACT are done
            TC_STAT(c.t_accw += clock64() - t0;)
            ctrl_act_load(c, copies + tc_copy_off(t, l - 1), (uint32_t)(TC_PATHS * t.ly[l - 1].K16 * 2));
            uint32_t sync = warp_uniform(c.sync);
            ctrl_wait_inputs(warp_uniform(smem_u32(c.bars)), sync, IN_HELP);         // the last drain of the layer: its region is the next accumulator
            c.sync = sync;
        }
    }
    if (need_w) {                                                        // layer 0
        ctrl_act_wait(c);
        const int nblk = (t.ly[0].kl + 1 + 127) >> 7;
        for (int b = 0; b < nblk; ++b) ctrl_gemm_dw(c, b, t.ly[0].N16, IN_HELP, !need_dy0 && b == nblk - 1);
        if (!need_dy0) {
            uint32_t sync = warp_uniform(c.sync);
            ctrl_wait_inputs(warp_uniform(smem_u32(c.bars)), sync, IN_HELP);         // the last drain
            c.sync = sync;
            return;
        }
    }
    // dy0 goes to the owners; after dW blocks its inputs were published at once, in a chain without them chunk by chunk
    ctrl_gemm_ts(c, t.ly[0].N16 / 16, t.ly[0].K16, need_w ? IN_HELP : IN_CHUNKS, true, false, false);
}

// ---- helpers ------------------------------------------------------------------------------------------------------
// relu masks of one evaluation: h[l][c] bit j: a_l[16c + j] > 0  (l = 1..L, K16 <= 256).  Per-thread local memory of the
// helper that converts chunk c of its lane in the forward AND in the backward (the chunk -> group assignment is static).
struct Masks { uint32_t h[MAXLIN][16]; };

// hidden-layer epilogues of a forward evaluation that keeps what the backward needs: relu masks (bits), FP16 copies of
// a_1..a_{L-1} (global scratch `copies`, NULL: none) and of a_L (shared ACT image, NULL: none).
// skip_last: the raw output is not needed -- the last hidden accumulator is not converted into planes (no region swap) and
// the epilogue publishes at once on a_help (the dW product of the last layer waits for it).
static __device__ __noinline__ uint32_t help_forward_keep_(PathArg p, const TcNet& t, const float* vec, Masks& mk, unsigned char* copies,
                                                           unsigned char* act, int row, bool skip_last) {
    for (int l = 0; l < t.L; ++l) {
        const int N16 = t.ly[l].N16;
        (void)vec;
        const bool last_hidden = (l == t.L - 1);
        unsigned char* dst = last_hidden ? act : (copies ? copies + tc_copy_off(t, l + 1) : nullptr);
        const int one_at = t.ly[l + 1].kl;
        const bool planes = !(last_hidden && skip_last);                 // (skip_last: nothing reads a_L as an MMA operand)
        for_acc_chunks(p, N16 / 16, planes ? EPI_CHUNKS : EPI_ALL, dst ? (last_hidden ? 1 : 2) : 0, planes, [&](int c, const uint32_t* r, uint32_t tc) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);       // (z itself: BatchNorm is folded into the product)
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) bits |= (v[j] > 0.f ? 1u : 0u) << j;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                up2(add2(pk2(v[2 * j], v[2 * j + 1]), pk2(fmaxf(v[2 * j], 0.f), fmaxf(v[2 * j + 1], 0.f))), v[2 * j], v[2 * j + 1]);
            mk.h[l + 1][c] = bits;
            if (planes) put16(tc, v);
            if (dst) copy16f(dst, row, c, v, one_at);                    // (shared ACT image / global copy scratch)
        });
    }
    return p.sync;
}
__device__ __forceinline__ void help_forward_keep(PathCtx& p, const TcNet& t, const float* vec, Masks& mk, unsigned char* copies,
                                                  unsigned char* act, int row, bool skip_last) {
    p.sync = help_forward_keep_(p, t, vec, mk, copies, act, row, skip_last);
}

// this thread's row of a dW block -> RED into the slab:  rows f = 128*blk + row (f <= kl), cols n < nl.  The evaluation ran
// scaled by 2^dexp (the owners leave the exponent in shared memory before they publish the output cotangent).
__device__ __forceinline__ void help_drain(PathCtx& p, int blk, int row, int kl, int nl, int N16, float* slab, const volatile int* dexp, bool in_planes_region) {
    {
        TC_STAT(const long long t0 = clock64();)
        TC_TRACE(p, 13);
        mbar_wait_u32(p.bars + 8 * BAR_DW, (p.sync >> 21) & 1u);
        TC_STAT(p.t_mark = clock64(); p.t_accw += p.t_mark - t0;)
        TC_TRACE(p, 14);
        p.sync ^= SY_DW;
        tc_fence_after();
    }
    const uint32_t acc = in_planes_region ? path_planes(p) : path_acc(p);
    TC_STAT(const long long td0 = clock64();)
    const int f = 128 * blk + row;
    float* dst = slab + (long long)f * 4;
    const long long gstride = (long long)(kl + 1) * 4;                  // next column group
    const bool rowok = f <= kl;
    // a warp whose 32 rows all lie past the last feature row skips its TMEM loads altogether (the TMEM->register path
    // is the bound of the drain): block 1 of a 200-wide layer has 73 live rows, the input layer's block 21
    const bool warp_live = (128 * blk + (row & ~31)) <= kl;
    const float inv = pow2f(-*dexp);
    for (int c = p.grp; warp_live && c < N16 / 16; c += TC_EGRP) {
        uint32_t r[16];
        tmem_ld16(acc + 16 * c, r);
        tmem_ld_wait();
        if (!rowok) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                     // (columns >= nl of the accumulator are zero)
            const int n = 16 * c + 4 * q;
            if (n < nl) red_add_v4(dst + (n >> 2) * gstride, __uint_as_float(r[4 * q]) * inv, __uint_as_float(r[4 * q + 1]) * inv,
                                   __uint_as_float(r[4 * q + 2]) * inv, __uint_as_float(r[4 * q + 3]) * inv);
        }
    }
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_u32(p.bars + 8 * BAR_HELP);    // accumulator drained
    p.sync ^= SY_HELP;
    TC_STAT(p.t_drain += clock64() - td0;)
    TC_TRACE(p, 15);
}

// the helpers' part of a backward evaluation (same order as ctrl_net_backward): for l = L..1 the drains of the layer's dW
// blocks (need_w; they sit in the region whose planes dX_l consumed), then the epilogue dz_{l-1} = dA_l (.) slope(a_l) ->
// planes (+ DZ image), handed over chunk by chunk -- except for l = 1 with dW products: layer 0's dW block comes before its
// dX and needs the whole image; finally the drains of layer 0
static __device__ __noinline__ uint32_t help_backward_(PathArg p, const TcNet& t, const TcSlab& g, const Masks& mk, bool need_w, float* slab,
                                                       unsigned char* dzimg, int row, const volatile int* dexp) {
    for (int l = t.L; l >= 1; --l) {
        if (need_w) {
            const int nblk = (t.ly[l].kl + 1 + 127) >> 7;
            for (int b = 0; b < nblk; ++b) help_drain(p, b, row, t.ly[l].kl, t.ly[l].nl, t.ly[l].N16, slab + g.gW[l], dexp, true);
        }
        const int K16 = t.ly[l].K16;
        // dA_l in the accumulator (K16_l columns) -> dz_{l-1} planes in place (+ DZ image)
        for_acc_chunks(p, K16 / 16, (need_w && l == 1) ? EPI_ALL : EPI_CHUNKS, need_w ? 1 : 0, true, [&](int c, const uint32_t* r, uint32_t tc) {
            const uint32_t bits = mk.h[l][c];
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float a = __uint_as_float(r[j]);
                v[j] = ((bits >> j) & 1u) ? 2.f * a : a;                     // d(z + relu z)
            }
            put16(tc, v);
            if (need_w) copy16f(dzimg, row, c, v, -1);
        });
    }
    if (need_w) {
        const int nblk = (t.ly[0].kl + 1 + 127) >> 7;
        for (int b = 0; b < nblk; ++b) help_drain(p, b, row, t.ly[0].kl, t.ly[0].nl, t.ly[0].N16, slab + g.gW[0], dexp, false);
    }
    return p.sync;
}
__device__ __forceinline__ void help_backward(PathCtx& p, const TcNet& t, const TcSlab& g, const Masks& mk, bool need_w, float* slab,
                                              unsigned char* dzimg, int row, const volatile int* dexp) {
    p.sync = help_backward_(p, t, g, mk, need_w, slab, dzimg, row, dexp);
}

// ---- owners -------------------------------------------------------------------------------------------------------
// cotangent of the raw output (nl <= 32 values, static indexing) -> planes (+ DZ image), publish.  The evaluation runs
// scaled by 2^dexp, chosen so that the largest |cotangent| of the tile lands in [2^11, 2^12): mxbuf = shared memory, two
// rows of one word per owner warp (alternating, so that one barrier per call is enough); dexp_out = the shared word the
// helpers read the exponent from when they drain.  skip_last: the forward stopped at the last hidden layer -- wait until
// its products are done (then the region that held the inputs of that layer is free for the dz planes).
template <int NO>
__device__ __forceinline__ void own_put_dz(PathCtx& p, const TcNet& t, const float (&dout)[NO], unsigned char* dzimg, int row, uint32_t* mxbuf,
                                           volatile int* dexp_out, bool skip_last) {
    const int N16 = t.ly[t.L].N16, nl = t.ly[t.L].nl;
    float m = 0.f;
#pragma unroll
    for (int n = 0; n < NO; ++n)
        if (n < nl) m = fmaxf(m, fabsf(dout[n]));
    uint32_t mu = __reduce_max_sync(0xffffffffu, __float_as_uint(m));        // (non-negative floats order like their bit patterns)
    uint32_t* mb = mxbuf + ((p.sync & SY_DZ) ? 8 : 0);
    p.sync ^= SY_DZ;
    if ((threadIdx.x & 31) == 0) mb[threadIdx.x >> 5] = mu;
    asm volatile("bar.sync 2, %0;" ::"r"(TC_OWN_THREADS) : "memory");
#pragma unroll
    for (int w = 0; w < TC_OWN_THREADS / 32; ++w) mu = max(mu, mb[w]);
    const int ex = (int)((mu >> 23) & 0xffu);
    p.dexp = (ex == 0 || ex == 255) ? 0 : (11 - (ex - 127));                  // zero / denormal / non-finite maximum: no scaling
    p.dexp = p.dexp < -100 ? -100 : (p.dexp > 100 ? 100 : p.dexp);
    if (threadIdx.x == 0) *dexp_out = p.dexp;
    const float sc = pow2f(p.dexp);
    if (skip_last) { own_wait_fin(p); own_swaps(p, t.L - 1); }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c < N16 / 16) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = 16 * c + j;
                v[j] = (n < NO && n < nl) ? dout[n < NO ? n : 0] * sc : 0.f;
            }
            put16(path_planes(p) + 16 * c, v);
            if (dzimg) copy16f(dzimg, row, c, v, -1);
        }
    }
    if (dzimg) fence_proxy_async();
    own_publish(p);
}

// cotangent of y0 (in <= 31 values, static indexing); the helpers converted the L accumulators dA_L..dA_1 in the meantime
template <int NX>
__device__ __forceinline__ void own_get_dy0(PathCtx& p, const TcNet& t, float (&dy0)[NX]) {
    own_wait_fin(p);
    own_swaps(p, t.L);
    const uint32_t acc = path_acc(p);
    const float inv = pow2f(-p.dexp);                                      // the evaluation ran scaled by 2^dexp
    const int K16 = t.ly[0].K16;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c < K16 / 16) {
            uint32_t r[16];
            tmem_ld16(acc + 16 * c, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = 16 * c + j;
                if (k < NX && k < t.in) dy0[k < NX ? k : 0] = __uint_as_float(r[j]) * inv;
            }
        }
    }
}
// What the combined mode (no helper warps) needs besides the owner's own arguments: the helpers' arguments.
struct HelpArgs {
    Masks* mk; unsigned char* copies; unsigned char* act; const TcSlab* g; float* slab;
};
// forward evaluation that keeps what the backward needs, as the owners see it: y0 (+ its FP16 copy) -> [hidden layers] ->
// raw output (skip_last: no output, the backward starts from the last hidden layer)
template <int NX, int NO>
__device__ __forceinline__ void own_net_forward_keep(PathCtx& p, const TcNet& t, const float* vec, const float (&x)[NX], float (&out)[NO],
                                                     unsigned char* copies, int row, bool skip_last, const HelpArgs& h) {
    own_put_y0(p, t, vec, x, copies, row);
    if (TC_COMBINED) help_forward_keep(p, t, vec, *h.mk, h.copies, h.act, row, skip_last);
    if (!skip_last) own_last(p, t, vec, out);
}
// backward of one network evaluation as the owners see it.  dout: cotangent of the raw output; dy0 receives the cotangent of y0.
template <int NO, int NX>
__device__ __forceinline__ void own_net_backward(PathCtx& p, const TcNet& t, const float (&dout)[NO], bool need_w, unsigned char* dzimg, int row,
                                                 float (&dy0)[NX], uint32_t* mxbuf, volatile int* dexp_out, bool skip_last, const HelpArgs& h) {
    own_put_dz(p, t, dout, need_w ? dzimg : nullptr, row, mxbuf, dexp_out, skip_last);
    if (TC_COMBINED) help_backward(p, t, *h.g, *h.mk, need_w, h.slab, dzimg, row, dexp_out);
    own_get_dy0(p, t, dy0);
}
// the same when nothing consumes dy0 (the critic's networks): wait until the last dW product is done
template <int NO>
__device__ __forceinline__ void own_net_backward_nody0(PathCtx& p, const TcNet& t, const float (&dout)[NO], unsigned char* dzimg, int row,
                                                       uint32_t* mxbuf, volatile int* dexp_out, bool skip_last, const HelpArgs& h) {
    own_put_dz(p, t, dout, dzimg, row, mxbuf, dexp_out, skip_last);
    if (TC_COMBINED) help_backward(p, t, *h.g, *h.mk, true, h.slab, dzimg, row, dexp_out);
    own_wait_fin(p);
    own_swaps(p, t.L);
}

}  // namespace tc
}  // namespace dpb
