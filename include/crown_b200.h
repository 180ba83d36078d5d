/*
 * crown_b200.h — C-ABI of libcrown_b200.so: B200-native (sm_100a) batched backward linear bound
 * propagation (CROWN / alpha-CROWN / beta-CROWN) over a batch of BaB sub-domains.
 *
 * This is the drop-in boundary for NeuralSAT's theory-solver hot path.  The reference has no FFI
 * for this path (it is Python calling PyTorch ops); the entry points below are what a binding
 * for it would bind, one per reference function (paths under /root/reference/neuralsat-pt201):
 *
 *   cb_plan_create      <- BoundedModule.__init__ graph construction
 *                          (auto_LiRPA/bound_general.py:557-648) for the operator set of
 *                          auto_LiRPA/operators/{linear,convolution,normalization,add_sub,relu,shape}.py
 *   cb_crown_pass       <- BoundedModule.compute_bounds(method='backward', reuse_alpha=True,
 *                          interm_bounds=...) -> backward_general
 *                          (auto_LiRPA/bound_general.py:921-1179, auto_LiRPA/backward_bound.py:102-311)
 *                          as called by NetworkAbstractor._forward_hidden(simplify=True)
 *                          (abstractor/abstractor.py:274-287)                      [form F1]
 *   cb_crown_grad       <- loss.backward() of one optimiser iteration
 *                          (auto_LiRPA/optimized_bounds.py:554, operators/clampmult.py:49-95)
 *   cb_optimize         <- BoundedModule.compute_bounds(method='crown-optimized', interm_bounds=...)
 *                          -> _get_optimized_bounds (auto_LiRPA/optimized_bounds.py:255-629)
 *                          as called by NetworkAbstractor._forward_hidden(simplify=False)
 *                          (abstractor/abstractor.py:289-311)                      [form F2]
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless named h_*;
 *     pointer TABLES (arrays of device pointers, `const float* const*`) live in HOST memory.
 *   - no allocation on the per-call path: the caller provides `workspace` of at least
 *     cb_workspace_bytes(); cb_plan_create may allocate device memory for pre-split weights.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls return after
 *     enqueueing unless documented otherwise (cb_optimize syncs only when early-stop is enabled).
 *   - return value: 0 = ok; CB_ERR_* otherwise; cb_last_error() gives a message for the calling
 *     thread.  Allocation failure is CB_ERR_OOM — the Python shim maps it to
 *     RuntimeError("CUDA out of memory. ...") as the reference's batch-halving logic expects
 *     (util/misc/torch_cuda_memory.py:58-61).
 *   - layouts are the reference's: C [Bd,S,n_out]; x_L,x_U [Bd,n_in]; lower/upper[k] [Bd,n_k];
 *     lb [Bd,S]; internal A and the lA outputs are [S,Bd,n_k] (auto_LiRPA/backward_bound.py:590-594);
 *     alpha[k] points at plane 0 of the reference's [2,S1,Bd,n_alpha_k] tensor, S1 in {1,S};
 *     beta_{val,sign,bias}[k] [Bd,J_k] fp32, beta_loc[k] [Bd,J_k] int64 (auto_LiRPA/beta_crown.py:11-42).
 *   - fp32 arithmetic throughout; int64 indices as in the reference.
 */
#ifndef CROWN_B200_H
#define CROWN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_OK            0
#define CB_ERR_ARG       1   /* bad argument / unsupported graph */
#define CB_ERR_CUDA      2   /* CUDA runtime error (message in cb_last_error) */
#define CB_ERR_OOM       3   /* device allocation failed */
#define CB_ERR_WORKSPACE 4   /* workspace too small */

enum cb_op {
    CB_OP_INPUT = 0,
    CB_OP_LINEAR = 1,       /* y = x W^T + b,  weight [out,in]                       */
    CB_OP_CONV2D = 2,       /* weight [Cout,Cin,kh,kw], groups == 1                  */
    CB_OP_BATCHNORM2D = 3,  /* eval mode; weight = gamma/sqrt(var+eps), bias = beta-mean*weight */
    CB_OP_ADD = 4,
    CB_OP_SUB = 5,
    CB_OP_FLATTEN = 6,      /* any reshape: a view                                   */
    CB_OP_RELU = 7,
    CB_OP_SIGMOID = 8,      /* weight = d_lower, bias = d_upper: the tangent-point tables of              */
    CB_OP_TANH = 9,         /* auto_LiRPA/operators/tanh.py:65-130 (device, kh = entries per table)      */
    CB_OP_ADDCONST = 10     /* y = x + bias, bias [numel] an unperturbed operand (Add/Sub with a constant,
                               auto_LiRPA/backward_bound.py:712-721, operators/base.py:320-341)           */
};

/* One graph node, topological order, node 0 = input, last node = output. */
typedef struct cb_node {
    int32_t op;             /* enum cb_op */
    int32_t in0, in1;       /* input node indices, -1 if unused */
    int32_t c, h, w;        /* output shape without batch: (c,h,w), or (n,1,1) for vectors */
    const float* weight;    /* device, see cb_op */
    const float* bias;      /* device or NULL */
    int32_t kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, groups;
} cb_node_t;

typedef struct cb_plan cb_plan_t;

/* Per-call description of one batch of sub-domains.  n_act = number of activation nodes of the
 * plan (cb_plan_num_activations), tables are indexed by activation order (== the reference's
 * net.relus / net.split_nodes order). */
typedef struct cb_problem {
    int32_t Bd;                         /* sub-domains in the batch                        */
    int32_t S;                          /* spec rows per sub-domain (C.shape[1])           */
    const float* C;                     /* [Bd,S,n_out]                                    */
    const float* x_L;                   /* [Bd,n_in]                                       */
    const float* x_U;                   /* [Bd,n_in]                                       */
    const float* const* lower;          /* h_table[n_act] -> [Bd,n_k]                      */
    const float* const* upper;          /* h_table[n_act] -> [Bd,n_k]                      */
    float* const* alpha;                /* h_table[n_act] -> [S1,Bd,n_alpha_k] or NULL entry
                                           (NULL entry / NULL table => CROWN-adaptive slope,
                                           auto_LiRPA/operators/relu.py:347-371)            */
    const int32_t* const* alpha_pos;    /* h_table[n_act] -> [n_k] neuron -> column of alpha[k],
                                           -1 = not stored (sparse-feature alpha,
                                           operators/relu.py:208-221); NULL entry => dense   */
    const int32_t* n_alpha;             /* h[n_act] columns of alpha[k]                     */
    int32_t alpha_S1;                   /* spec dim of alpha: 1 (shared) or S               */
    float* const* beta_val;             /* h_table[n_act] -> [Bd,J_k] or NULL table = no beta */
    const int64_t* const* beta_loc;     /* h_table[n_act] -> [Bd,J_k]                       */
    const float* const* beta_sign;      /* h_table[n_act] -> [Bd,J_k]                       */
    const float* const* beta_bias;      /* h_table[n_act] -> [Bd,J_k] or NULL entries       */
    const int32_t* beta_J;              /* h[n_act] J_k (0 => no beta on that layer)        */
    float* lb;                          /* out [Bd,S]                                      */
    float* const* lA;                   /* h_table[n_act] -> out [S,Bd,n_k]; NULL table or NULL
                                           entries => not returned                          */
} cb_problem_t;

/* Options of the alpha/beta optimisation loop (auto_LiRPA/optimized_bounds.py:16-58 and
 * abstractor/params.py:51-65). */
typedef struct cb_opt {
    int32_t iteration;          /* 20                                                       */
    float lr_alpha, lr_beta;    /* 0.1, 0.1                                                 */
    float lr_decay;             /* 0.98 (ExponentialLR, stepped every iteration)            */
    int32_t early_stop_patience;/* 10                                                       */
    float start_save_best;      /* 0.5                                                      */
    int32_t enable_beta;        /* optimise beta_val as well                                */
    int32_t early_stop;         /* 1: reproduce the reference's data-dependent early exits
                                   (all-verified / patience); costs one 16-byte D2H + stream
                                   sync per iteration.  0: always run `iteration` passes    */
    const float* rhs;           /* [Bd,S] decision threshold; NULL => nothing is ever "verified" */
} cb_opt_t;

const char* cb_last_error(void);
int cb_version(void);

int cb_plan_create(const cb_node_t* h_nodes, int32_t n_nodes, cb_plan_t** out_plan);
void cb_plan_destroy(cb_plan_t* plan);
int32_t cb_plan_num_activations(const cb_plan_t* plan);
/* node index of the k-th activation / of its pre-activation node; -1 if out of range */
int32_t cb_plan_activation_node(const cb_plan_t* plan, int32_t k);
int32_t cb_plan_preact_node(const cb_plan_t* plan, int32_t k);

/* mode: 0 = cb_crown_pass only, 1 = + cb_crown_grad, 2 = cb_optimize */
size_t cb_workspace_bytes(const cb_plan_t* plan, int32_t Bd, int32_t S, int32_t mode,
                          const cb_problem_t* problem /* may be NULL: worst case sizes */);

/* F1: one backward pass; writes problem->lb (and lA when requested). */
int cb_crown_pass(const cb_plan_t* plan, const cb_problem_t* problem,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Gradient of sum_{b,s} lb[b,s] w.r.t. alpha[k] (-> grad_alpha[k], same shape as alpha[k]) and
 * beta_val[k] (-> grad_beta[k] [Bd,J_k]); runs the pass first.  Test/diagnostic entry point. */
int cb_crown_grad(const cb_plan_t* plan, const cb_problem_t* problem,
                  float* const* h_grad_alpha, float* const* h_grad_beta,
                  void* workspace, size_t workspace_bytes, void* stream);

/* F2: the optimisation loop.  On return (after enqueue) problem->lb holds the per-domain best
 * bounds, alpha[k] / beta_val[k] hold the reference's best snapshots, lA[k] the coefficients of
 * the LAST executed pass (operators/relu.py:244).  h_n_iter (host, may be NULL) receives the
 * number of passes executed (only meaningful with early_stop=1, otherwise == iteration). */
int cb_optimize(const cb_plan_t* plan, const cb_problem_t* problem, const cb_opt_t* opt,
                void* workspace, size_t workspace_bytes, void* stream, int32_t* h_n_iter);

/* Number of Linear contractions of the plan that run on the tcgen05 tensor cores (pass + gradient
 * direction); 0 when the environment sets CROWN_B200_DISABLE_TC=1 at plan creation. */
int32_t cb_plan_uses_tensor_cores(const cb_plan_t* plan);

/* 1 when the plan is a Linear/ReLU chain whose whole pass runs in one kernel (crown_chain.cu), 2 when
 * its alpha/beta gradient does too (crown_chain_grad.cu, used for S == 1 batches);
 * 0 otherwise or when the environment sets CROWN_B200_DISABLE_CHAIN=1 at plan creation
 * (CROWN_B200_DISABLE_CHAIN_GRAD=1 keeps the pass kernel only). */
int32_t cb_plan_uses_chain(const cb_plan_t* plan);

/* Number of (Conv2d node, direction) pairs - transpose-convolution in the pass, convolution in the gradient - that
 * run as tcgen05 implicit GEMMs (crown_conv_tc.cu); 0 when CROWN_B200_DISABLE_CONV_TC=1 or CROWN_B200_DISABLE_TC=1 at
 * plan creation.  Per pair the plan keeps the faster of the tensor-core and the register-tiled SIMT kernel, timed once
 * on this device (CROWN_B200_CONV_AUTOTUNE=0: the tensor-core kernel wherever its geometry applies). */
int32_t cb_plan_uses_conv_tc(const cb_plan_t* plan);

/* The per-(Conv2d node, direction) decisions as a string, two characters per node in graph order (pass, gradient):
 * 'T' tensor-core, 'S' register-tiled SIMT.  Returns its length, -1 if `cap` is too small.  Exported in the environment
 * variable CROWN_B200_CONV_CHOICES at plan creation, the string is replayed instead of timing the kernels: profiling
 * runs use it, because event timings taken under a profiler do not rank the kernels the way an ordinary run does. */
int32_t cb_plan_conv_choices(const cb_plan_t* plan, char* out, int32_t cap);

/* Self-test of the tensor-core convolution alone (square stride / padding, dilation 1), all device pointers:
 * dir 0: Y[rows,Cin,Hin,Win] (+)= conv_transpose2d(X[rows,Cout,Hout,Wout], W), bias_rows[r] += sum X[r,co,:,:] bias[co]
 * dir 1: Y[rows,Cout,Hout,Wout] = conv2d(X[rows,Cin,Hin,Win], W) + bias.  Synchronises the stream. */
int cb_debug_conv_tc(const float* X, const float* W, const float* bias, float* Y, float* bias_rows, int32_t rows,
                     int32_t Cin, int32_t Hin, int32_t Win, int32_t Cout, int32_t KH, int32_t KW, int32_t stride,
                     int32_t pad, int32_t dir, int32_t accumulate, void* stream);

/* Self-test of the register-tiled fp32 convolution (crown_conv.cu: the thin first-layer kernels when the image side
 * has at most four channels, the tiled kernels otherwise); same conventions as cb_debug_conv_tc, no bias_rows. */
int cb_debug_conv_simt(const float* X, const float* W, const float* bias, float* Y, int32_t rows, int32_t Cin,
                       int32_t Hin, int32_t Win, int32_t Cout, int32_t KH, int32_t KW, int32_t stride, int32_t pad,
                       int32_t dir, int32_t accumulate, void* stream);

/* Self-test of the tcgen05 3xTF32 contraction alone: Y[rows,N] = X[rows,K] . W[N,K]^T (+ col_bias[N]),
 * all device pointers, row-major fp32.  bn = column tile (0 = automatic).  Allocates scratch and
 * synchronises the stream; not part of the hot path. */
int cb_debug_tc_gemm(const float* X, const float* W, const float* col_bias, float* Y, int32_t rows,
                     int32_t N, int32_t K, int32_t bn, int32_t dbg, void* stream);

/* Self-test: when set to a device buffer of 8 int64 per CTA, every tensor-core launch records
 * clock64 stamps (start, setup done, first operand stage landed, last MMA issued, epilogue operands
 * landed, accumulator ready, epilogue done).  NULL switches it off. */
void cb_debug_tc_times(void* device_buffer);

/* ---- device-resident domain store and branching (crown_store.cu) ---------------------------------------------
 * Next to the bounding path: what DomainsList / TensorStorage (heuristic/domains_list.py:153-311,
 * util/misc/tensor_storage.py:4-97), the child construction of NetworkAbstractor._forward_hidden
 * (abstractor/utils.py:159-250) and the BaBSR + look-ahead branching (heuristic/util.py:31-72,
 * heuristic/decision_heuristics.py:78-251) do on the host, as kernels over records that never leave HBM. */
enum cb_copy_mode { CB_COPY_F32 = 0, CB_COPY_F16_TO_F32 = 1, CB_COPY_F32_TO_F16 = 2, CB_COPY_I32 = 3,
                    CB_COPY_I32_TO_I64 = 4, CB_COPY_I64_TO_I32 = 5 };
typedef struct cb_copy_desc {
    const void* src; void* dst;
    int32_t width;                 /* elements copied per row                                              */
    int32_t dst_width;             /* > width: the rest of the destination row is zero-filled (F32 / I32..) */
    int32_t mode;                  /* enum cb_copy_mode                                                    */
    int32_t src_S, src_Bd;         /* src_S > 0: the source is [S,Bd,width] (lA of a pass) and row b of the
                                      destination is [S,width] (the store keeps lAs as [Bd,S,n])           */
    int64_t src_stride, dst_stride;/* elements between rows                                                */
} cb_copy_desc_t;
/* every descriptor: dst[dst_map[r]] = convert(src[src_map[r]]) for r < R; NULL map = identity, negative = skip.
 * d_descs is a DEVICE array. */
int cb_store_multi_copy(const cb_copy_desc_t* d_descs, int32_t n_descs, const int32_t* src_map, const int32_t* dst_map,
                        int32_t R, void* stream);
typedef struct cb_split_layer {
    float* lower; float* upper; int32_t n;          /* [R,n] children                                     */
    int32_t J;                                      /* width of the history arrays                        */
    int32_t* hist_cnt;                              /* [R] or NULL                                        */
    int64_t* hist_loc; float* hist_sign; float* beta_val; float* hist_point;   /* [R,J]                   */
} cb_split_layer_t;
/* child r: lower[layer][r, neuron] = point if side > 0 else upper[...] = point, plus the history entry
 * (loc, sign, beta 0); d_layers is a DEVICE array (abstractor/utils.py:159-178, :214-250). */
int cb_store_apply_split(cb_split_layer_t* d_layers, int32_t n_layers, const int32_t* dec_layer, const int32_t* dec_neuron,
                         const float* dec_side, const float* dec_point, int32_t R, void* stream);
/* keep = all_s(lb <= rhs) (domains_list.py:246); rank[r] = base + #kept before r, or -1; out[0] = #kept,
 * out[1+k] = max(out[1+k], hist_cnt[k][r] over kept rows).  d_hist_cnt: DEVICE table of n_layers pointers. */
int cb_store_keep_rank(const float* lb, const float* rhs, int32_t R, int32_t S, int32_t base, int32_t* rank, int32_t* out,
                       const int32_t* const* d_hist_cnt, int32_t n_layers, void* stream);
/* BaBSR score / intercept score (and the unstable mask l < 0 < u, may be NULL) of one layer into columns
 * [col0, col0+n) of [B,ld] matrices (heuristic/util.py:17-72) */
int cb_babsr_scores(const float* lA, const float* lower, const float* upper, const float* bias, int32_t B, int32_t S,
                    int32_t n, float* score, float* backup, float* mask_out, int32_t ld, int32_t col0, void* stream);
/* k largest (largest != 0) or smallest entries of every row of x [B,n]: vals / idx [B,k], ties to the lowest index */
int cb_topk_rows(const float* x, int32_t B, int32_t n, int32_t k, int32_t largest, float* vals, int32_t* idx, void* stream);
/* arg-max over the K look-ahead passes and score-vs-backup choice (decision_heuristics.py:159-251):
 * lb_k [K,4B] = (lb - rhs).max(-1) of pass k, row = half*2B + slot; dec_flat [B] = flat neuron index */
int cb_pick_decision(const float* lb_k, const float* score_val, const int32_t* score_idx, const float* backup_val,
                     const int32_t* backup_idx, const float* mask_cat, int32_t n_total, int32_t B, int32_t K,
                     int32_t* dec_flat, void* stream);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------
 * cb_launch_count: kernels launched by this library in this process so far.
 * cb_profile_enable(1): every launch is bracketed by CUDA events on its own stream;
 * cb_profile_collect synchronises those events and returns, per kernel class, the summed device
 * time [ms] and the number of launches since the last collect. */
void cb_profile_enable(int32_t on);
int64_t cb_launch_count(void);
int32_t cb_profile_num_kernels(void);
const char* cb_profile_kernel_name(int32_t id);
int32_t cb_profile_collect(double* h_ms, int64_t* h_launches, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* CROWN_B200_H */
