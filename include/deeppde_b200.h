/* deeppde_b200.h -- C ABI of the B200-native rollout + TD-gradient library (libdeeppde_b200.so).
 *
 * The reference (MoZhou1995/DeepPDE_ActorCritic) has no FFI layer: its boundary is the Python
 * call surface of solver.py / equation.py.  Each entry point below replaces the arithmetic of
 * the reference functions named beside it; the Python host in deeppde_actorcritic_b200/ keeps
 * the reference's names and argument meaning and binds these symbols with ctypes
 * (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions: plain C, int return (0 = ok, else error; text via dpb_last_error); the caller
 * owns every buffer; all data pointers are DEVICE pointers unless the name ends in _host; all
 * arrays are contiguous, 16-byte aligned, element type given by dpb_config.dtype; the library
 * never allocates per call and never synchronises the stream it is given (cudaStream_t passed
 * as void*).  There is no CPU fallback: without a CUDA device every compute entry point
 * returns DPB_ERR_CUDA.
 *
 * Flat parameter layout of one network (in_dim -> hidden[0..L-1] -> out_dim), identical for
 * weights, gradients and Adam slots (reference variables: solver.py:239-258):
 *     bn0.gamma[in] bn0.beta[in]
 *     for i in 0..L-1:  W_i[in_i][hidden_i] (row-major, Keras kernel orientation)
 *                       bn_{i+1}.gamma[hidden_i] bn_{i+1}.beta[hidden_i]
 *     W_last[hidden_{L-1}][out]  b_last[out]  bn_last.gamma[out]  bn_last.beta[out]
 */
#ifndef DEEPPDE_B200_H
#define DEEPPDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPB_MAX_HIDDEN 6
#define DPB_MAX_DIM 32

enum { DPB_OK = 0, DPB_ERR_ARG = 1, DPB_ERR_CUDA = 2, DPB_ERR_WORKSPACE = 3 };
enum { DPB_EQN_LQR = 0, DPB_EQN_VDP = 1, DPB_EQN_EKN = 2, DPB_EQN_LQR_VAR = 3 };   /* equation.py:144,179,240,278 */
enum { DPB_SCHEME_NAIVE = 0, DPB_SCHEME_ADAPTIVE = 1 };                            /* equation.py:46,73 */
enum { DPB_TD1 = 1, DPB_TD2 = 2 };                                                  /* solver.py:177 */
enum { DPB_F32 = 0, DPB_F64 = 1 };
enum { DPB_NET_ACTOR = 0, DPB_NET_CRITIC = 1, DPB_NET_CRITIC_GRAD = 2 };            /* solver.py:145-146,200 */
enum { DPB_DW_EXTERNAL = 0, DPB_DW_PHILOX_NORMAL = 1, DPB_DW_PHILOX_BOUNDED = 2 };  /* equation.py:19,31-32 */
enum { DPB_IMPL_EXACT = 0, DPB_IMPL_TENSOR = 1 };
enum { DPB_CF_V_TRUE = 0, DPB_CF_U_TRUE = 1, DPB_CF_V_GRAD_TRUE = 2, DPB_CF_Z = 3, DPB_CF_W = 4, DPB_CF_DRIFT = 5, DPB_CF_SIGMA = 6 };

/* flags of dpb_critic_step / dpb_actor_step */
enum {
    DPB_FLAG_CHEAT_CONTROL = 1,   /* use u_true instead of NN_control (solver.py:153-157, equation.py:54,87) */
    DPB_FLAG_CHEAT_VALUE = 2,     /* use V_true instead of NN_value at x_N (solver.py:220-223) */
    DPB_FLAG_NEED_GRAD = 4,       /* also produce parameter gradients (solver.py:85-97) */
    DPB_FLAG_PROPAGATE_ONLY = 8   /* stop after the rollout: only x_smp/dt/coef are written (equation.py:46-106) */
};

typedef struct dpb_config {
    int32_t dtype;                 /* DPB_F32 | DPB_F64            (net_config.dtype) */
    int32_t eqn;                   /* DPB_EQN_*                    (eqn_config.eqn_name) */
    int32_t dim, control_dim;      /* eqn_config.dim, control_dim */
    int32_t scheme;                /* DPB_SCHEME_*                 (train_config.scheme) */
    int32_t td_type;               /* DPB_TD1 | DPB_TD2            (train_config.TD_type) */
    int32_t ekn_sigma_fix;         /* 0: sigma=sqrt(2) as equation.py:268; 1: sqrt(2*epsl) (consistent dynamics) */
    int32_t n_hidden_actor, n_hidden_critic;
    int32_t hidden_actor[DPB_MAX_HIDDEN];
    int32_t hidden_critic[DPB_MAX_HIDDEN];
    int32_t impl;                  /* DPB_IMPL_EXACT: FP32/FP64 CUDA-core FMA; DPB_IMPL_TENSOR: tcgen05 MLP layers */
    int32_t reserved[3];           /* reserved[0] bit 0: 1 = do NOT sort the paths by lifetime before tiling them (tensor path, naive
                                      scheme; the sort is a pure relabelling of the paths -- see dpb_api.cu -- and is on by default) */
    double R, discount;            /* eqn_config.R, discount */
    double p, q, beta;             /* LQR / LQR_var */
    double a, epsilon;             /* VDP (a, epsilon) / LQR_var (epsilon) */
    double a2, a3;                 /* ekn */
} dpb_config;

/* One rollout's inputs: the tuple (x0, dw, x_bdry) of solver.py:160,208 restricted to the
 * local shard.  dw may be NULL when dw_mode is a PHILOX mode: increments are then generated
 * in-kernel by Philox4x32-10 keyed by (seed, stream, GLOBAL path index, step, component/4). */
typedef struct dpb_inputs {
    const void* x0;                /* [B_local][dim] */
    const void* dw;                /* [B_local][dim][N] (N innermost, equation.py:19) or NULL */
    const void* x_bdry;            /* [B_local][dim] (critic only) */
    int32_t dw_mode;               /* DPB_DW_* */
    int32_t reserved;
    uint64_t seed;                 /* Philox key */
    uint64_t stream;               /* Philox stream id: (iteration << 1) | phase */
    const uint64_t* stream_base;   /* optional DEVICE word added to `stream` when the kernel runs (NULL: none).  Lets a captured
                                      CUDA graph of one training iteration be replayed: the host writes iteration << 1 there. */
} dpb_inputs;

/* Optional per-path outputs (any pointer may be NULL). */
typedef struct dpb_path_outputs {
    void* x_smp;                   /* [B_local][dim][N+1]  (equation.py:71,106) */
    void* dt;                      /* [B_local][N] */
    void* coef;                    /* [B_local][N] */
    void* delta;                   /* [B_local]  critic: TD residual (solver.py:189); actor: y (solver.py:224) */
    void* delta_bdry;              /* [B_local]  critic only (solver.py:190) */
    int32_t* exit_index;           /* [B_local]  number of accepted steps' last index+1 (sum of coef) */
} dpb_path_outputs;

typedef struct dpb_handle dpb_handle;

/* ActorCriticSolver.__init__ (solver.py:9-34): validates cfg, fixes shapes; owns no device memory. */
int dpb_create(dpb_handle** out, const dpb_config* cfg);
int dpb_destroy(dpb_handle* h);
const char* dpb_last_error(const dpb_handle* h);   /* h may be NULL: last global error */
const char* dpb_version(void);

/* len(flat parameters) of DeepNN(config, AC) (solver.py:227-258). */
int64_t dpb_param_count(const dpb_handle* h, int which_net);

/* Bytes of device workspace the step calls need for shards up to B_local paths of N steps. */
int64_t dpb_workspace_bytes(const dpb_handle* h, int64_t B_local, int32_t N);

/* Extra bytes the *_host entry points need after the first dpb_workspace_bytes() bytes of `workspace`
 * to stage x0, x_bdry (and dw when dw_mode is EXTERNAL) on the device. */
int64_t dpb_staging_bytes(const dpb_handle* h, int64_t B_local, int32_t N, int32_t dw_mode);

/* CriticModel.call + loss_critic + grad_critic (solver.py:73-78,85-90,159-191).
 *   out_loss[2]: { 100/B_global * sum rho(delta), 100/B_global * sum rho(delta_bdry) } over the shard;
 *   grad_V, grad_G: gradient of loss_critic w.r.t. NN_value / NN_value_grad parameters restricted
 *   to the shard and scaled by 1/B_global, so a SUM all-reduce over shards gives the reference's
 *   batch-mean gradient.  grad_G is zero-filled under TD2 (Keras skips the None gradients). */
int dpb_critic_step(dpb_handle* h, const void* theta_actor, const void* theta_V, const void* theta_G,
                    const dpb_inputs* in, int64_t B_local, int64_t path_offset, int64_t B_global,
                    int32_t N, double T, uint32_t flags,
                    void* out_loss, void* grad_V, void* grad_G, const dpb_path_outputs* outs,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* ActorModel.call + loss_actor + grad_actor (solver.py:80-83,92-97,207-224): back-propagation
 * through the whole trajectory by an explicit reverse sweep.  out_loss[1] = 1/B_global * sum y. */
int dpb_actor_step(dpb_handle* h, const void* theta_actor, const void* theta_V,
                   const dpb_inputs* in, int64_t B_local, int64_t path_offset, int64_t B_global,
                   int32_t N, double T, uint32_t flags,
                   void* out_loss, void* grad_actor, const dpb_path_outputs* outs,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* DeepNN.call (solver.py:260-278) on n points: x[n][in] -> out[n][out_dim] (ekn actor: control_dim). */
int dpb_mlp_forward(dpb_handle* h, int which_net, const void* theta, const void* x, int64_t n,
                    void* out, void* workspace, int64_t workspace_bytes, void* stream);

/* Closed forms of the equation on n points (equation.py:157-167,201-227,252-265,292-302):
 *   which = DPB_CF_V_TRUE (out[n]), DPB_CF_U_TRUE (out[n][control_dim]), DPB_CF_V_GRAD_TRUE (out[n][dim]),
 *   DPB_CF_Z (Z_tf, out[n]), DPB_CF_W (w_tf(x,u), out[n]; u[n][control_dim] required, else NULL),
 *   DPB_CF_DRIFT (Equation.drift(x,u), equation.py:172,232-235,270-273,307: out[n][dim], u required),
 *   DPB_CF_SIGMA (Equation.sigma(x,u,n), equation.py:170,230,268,305: out[n][dim][dim], diagonal; u required for LQR_var). */
int dpb_closed_form(dpb_handle* h, int which, const void* x, const void* u, int64_t n, void* out, void* stream);

/* Equation.diffusion(x, u, dw, n) = sigma(x,u) . dw (equation.py:175-176,237-238,275-276,310-311): x[n][dim],
 * u[n][control_dim] (NULL allowed unless the equation is LQR_var), dw[n][dim] -> out[n][dim]. */
int dpb_diffusion(dpb_handle* h, const void* x, const void* u, const void* dw, int64_t n, void* out, void* stream);

/* The reductions of err_value / err_control / err_value_grad / err_value_infty (solver.py:109-130) on n
 * values: out3 = { sum (truth-approx)^2, sum truth^2, max |truth-approx| } (device, deterministic). */
int dpb_err_metrics(dpb_handle* h, const void* truth, const void* approx, int64_t n, void* out3, void* stream);

/* tf.keras Adam step as used at solver.py:16-21,99-107 on a flat vector:
 *   m += (g-m)(1-b1); v += (g*g-v)(1-b2); theta -= lr_t * m / (sqrt(v)+eps),  lr_t given by the host. */
int dpb_adam_step(dpb_handle* h, void* theta, const void* grad, void* m, void* v, int64_t n,
                  double lr_t, const double* lr_t_dev, double beta1, double beta2, double eps, void* stream);
/* (lr_t_dev: optional DEVICE double that overrides lr_t when the kernel runs -- CUDA-graph replays, as dpb_inputs.stream_base) */

/* The Brownian increments the PHILOX modes use, materialised as dw[B_local][dim][N] (for parity
 * tests: feed the same tensor to the oracle). */
int dpb_philox_dw(dpb_handle* h, int32_t dw_mode, uint64_t seed, uint64_t stream_id, int64_t path_offset,
                  int64_t B_local, int32_t N, void* dw_out, void* stream);

/* Device-side version of the x0 / x_bdry part of Equation.sample_normal (equation.py:14-22): x0 uniform
 * in the ball |x|<R, x_bdry uniform on the sphere, Philox-keyed by the GLOBAL path index (so that a run
 * sees the same paths however it is sharded).  Either output may be NULL. */
int dpb_sample_x(dpb_handle* h, uint64_t seed, uint64_t stream_id, const uint64_t* stream_base, int64_t path_offset, int64_t B_local,
                 void* x0_out, void* xb_out, void* stream);
/* (stream_base: optional DEVICE word added to stream_id, see dpb_inputs.stream_base) */

/* Host-buffer convenience used for end-to-end timing: same as dpb_critic_step / dpb_actor_step but
 * x0/dw/x_bdry are HOST pointers (pinned or pageable); they are copied to device staging inside
 * `workspace` (after its first dpb_workspace_bytes() bytes; see dpb_staging_bytes) on `stream`, and out_loss_host receives the losses after a stream synchronise. */
int dpb_critic_step_host(dpb_handle* h, const void* theta_actor, const void* theta_V, const void* theta_G,
                         const dpb_inputs* in_host, int64_t B_local, int64_t path_offset, int64_t B_global,
                         int32_t N, double T, uint32_t flags,
                         void* out_loss_host, void* grad_V, void* grad_G,
                         void* workspace, int64_t workspace_bytes, void* stream);
int dpb_actor_step_host(dpb_handle* h, const void* theta_actor, const void* theta_V,
                        const dpb_inputs* in_host, int64_t B_local, int64_t path_offset, int64_t B_global,
                        int32_t N, double T, uint32_t flags,
                        void* out_loss_host, void* grad_actor,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* Number of kernel launches issued by this handle since creation (for bench.py's gpu_launches). */
int64_t dpb_launch_count(const dpb_handle* h);

/* Switch the CUDA-event timing of the rollout kernels off / on (off while an iteration is being captured into a CUDA graph:
 * events recorded during capture cannot be queried). */
int dpb_set_timing(dpb_handle* h, int enable);

/* Device time in ms of the most recent critic/actor rollout kernel of this handle, from CUDA events the
 * library records on the caller's stream around that launch (synchronises on the stop event; < 0 if none). */
double dpb_last_kernel_ms(dpb_handle* h);

/* Diagnostic: one CTA runs the three tcgen05 product forms of the tensor path on A[128][K], B[208][K]
 * (device, float, bf16-representable) and writes D[3][128][208]: D0 = A B^T (operands in shared memory),
 * D1 = A B^T (A read from tensor memory), D2[f][g] = sum_p A[p][f] B[p][g] (both operands MN-major). */
int dpb_tc_selftest(const float* A, const float* B, float* D, int K, void* stream);

/* Diagnostic (tensor path): cycles of one empty control-thread <-> path-thread round trip (mbarrier hand-offs, one
 * minimal tcgen05.mma + commit, one TMEM load/store): out_host[0] = cycles per round trip, out_host[1] = rounds. */
int dpb_tc_handshake_cycles(int64_t* out_host, int rounds);

/* Diagnostic (tensor path): cycles of one 200-wide hidden-layer epilogue (13 chunks: TMEM load, affine + activation,
 * bf16 hi/lo split, TMEM stores in place) with `ngroups` (1..4) groups of 4 warps sharing the chunks; with_mma = 1: while
 * another warp keeps the tensor pipe busy with N=208 products; publish = 1: with the per-chunk publish sequence of the
 * kernels.  out_host[0] = cycles. */
int dpb_tc_epilogue_cycles(int64_t* out_host, int rounds, int ngroups, int with_mma, int publish);

/* Diagnostic (tensor path): cycles per tcgen05.mma (M=128, K=16, bf16, FP32 accumulation) with N = n output columns, issued
 * back to back by one thread, `per_commit` MMAs per tcgen05.commit; ts = 1: A operand in tensor memory, 0: shared memory.
 * out_host[0] = cycles per MMA until the last one completed, out_host[1] = cycles per MMA spent issuing. */
int dpb_tc_mma_cycles(int64_t* out_host, int n, int rounds, int ts, int per_commit);

/* Diagnostic (tensor path): cycle counters of CTA 0 of the last critic/actor launch that used `workspace`
 * (synchronous copy): [0] kernel cycles, [1] control thread waiting for the path threads, [3] tensor-pipe ops,
 * [4] path thread 0 waiting for the tensor pipe, [5] its epilogue cycles, [6] of which hidden-layer epilogues,
 * [7] control thread inside MMA issue loops. */
int dpb_tc_stats(dpb_handle* h, const void* workspace, int64_t B_local, int32_t N, int64_t* out_host);

/* Diagnostic (tensor path, library built with DPB_TC_STATS; otherwise an error): event trace of CTA 0 of the last
 * critic/actor launch -- out_host[3][4096] words (clock64 << 8) | event id for owner thread 0, helper thread 128 and the
 * control warp; unused entries keep what an earlier launch left.  Event ids: deeppde_actorcritic_b200/csrc/dpb_tc_nets.cuh. */
int dpb_tc_trace(dpb_handle* h, const void* workspace, int64_t B_local, int32_t N, uint64_t* out_host);

#ifdef __cplusplus
}
#endif
#endif /* DEEPPDE_B200_H */
