// C-ABI of libcrown_b200.so (see include/crown_b200.h): plan construction, workspace carving and
// the host-side schedules of the backward pass (F1), its gradient, and the alpha/beta loop (F2).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/crown_b200.h"
#include "crown_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CB_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            const int _code = (_e == cudaErrorMemoryAllocation) ? CB_ERR_OOM : CB_ERR_CUDA; \
            return fail(_code, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
        }                                                                               \
    } while (0)

struct Node {
    cb_node_t d;
    int64_t numel = 0;
    std::vector<int> consumers;
    int act_index = -1;      // k if this node is the k-th activation
    int preact_index = -1;   // k if this node feeds the k-th activation
    bool on_path = false;    // reaches the output node
    bool need_g = false;     // gradient w.r.t. its A is needed by some alpha/beta
    int a_alias = -1;        // node that owns this node's A storage (flatten of a single-consumer input)
    float* wt = nullptr;     // conv: weight transposed to [Cin,KH,KW,Cout] (owned)
    float *wk_b = nullptr, *wk_f = nullptr;   // conv: padded layouts of the register-tiled kernels (owned)
    // conv on the tensor cores (crown_conv_tc.cu): geometry + packed bf16x3 weights per direction (owned)
    bool ct_ok = false;
    bool ct_use_pass = false, ct_use_grad = false;   // per direction: the tensor-core kernel is the faster one (autotuned)
    cb::ConvTcGeom ct_pass, ct_grad;
    uint16_t *wct_pass = nullptr, *wct_grad = nullptr;
    // tcgen05 path (crown_tc.cu); all graph-static
    int tc_pass = 0;         // linear: 0 = SIMT, 1 = fused Linear+ReLU-below, 2 = fused Linear+concretize
    int tc_relu = -1;        // tc_pass == 1: the ReLU node below
    int tc_dst = -1;         // node whose A the fused launch produces (pre-activation node, or 0)
    int tc_grad = 0;         // linear: forward GEMM fused with the gradient of the ReLU above
    int tc_grad_relu = -1;
    int tc_src = -1;         // linear: input node with single-consumer flattens skipped
    int bn_pass = 0, bn_grad = 0;
    uint16_t *wp_pass = nullptr, *wp_grad = nullptr;   // packed bf16x3 weights (owned)
    uint16_t* wp_chain = nullptr;                      // packed for the whole-network kernel (owned)
    uint16_t* wp_chain_g = nullptr;                    // ... gradient direction (owned)
    bool a_packed = false;   // A of this node is consumed in packed form (it is a tc_pass linear)
    bool a_plain = true;     // A of this node is consumed by a SIMT kernel (plain fp32 rows)
    bool g_packed = false;   // dlb/dA of this node is consumed by a tc_grad linear
    bool g_plain = true;     // ... by a SIMT kernel
};

bool is_act(int op) { return op == CB_OP_RELU || op == CB_OP_SIGMOID || op == CB_OP_TANH; }

}  // namespace

struct cb_plan {
    std::vector<Node> nodes;
    std::vector<int> acts;      // node index of k-th activation
    int64_t sum_numel = 0;      // floats per row over all A buffers
    int n_in = 0, n_out = 0;
    bool use_tc = false;
    bool chain = false;                 // Linear/ReLU chain: the whole pass runs in one kernel (crown_chain.cu)
    bool chain_grad = false;            // ... and so does the gradient (crown_chain_grad.cu)
    std::vector<int> chain_lin;         // Linear nodes, output first
    std::vector<int> chain_relu;        // chain_relu[j] = ReLU below chain_lin[j] (-1 for the first Linear)
    ~cb_plan() {
        for (auto& n : nodes) {
            if (n.wt) cudaFree(n.wt);
            if (n.wk_b) cudaFree(n.wk_b);
            if (n.wk_f) cudaFree(n.wk_f);
            if (n.wct_pass) cudaFree(n.wct_pass);
            if (n.wct_grad) cudaFree(n.wct_grad);
            if (n.wp_pass) cudaFree(n.wp_pass);
            if (n.wp_grad) cudaFree(n.wp_grad);
            if (n.wp_chain) cudaFree(n.wp_chain);
            if (n.wp_chain_g) cudaFree(n.wp_chain_g);
        }
    }
};

namespace {

cb::ConvGeom conv_geom(const cb_plan* p, const Node& n) {
    const Node& src = p->nodes[n.d.in0];
    cb::ConvGeom g;
    g.Cin = src.d.c; g.Hin = src.d.h; g.Win = src.d.w;
    g.Cout = n.d.c; g.Hout = n.d.h; g.Wout = n.d.w;
    g.KH = n.d.kh; g.KW = n.d.kw; g.sh = n.d.stride_h; g.sw = n.d.stride_w;
    g.ph = n.d.pad_h; g.pw = n.d.pad_w; g.dh = n.d.dil_h; g.dw = n.d.dil_w;
    return g;
}

__global__ void k_transpose_conv_w(const float* __restrict__ W, float* __restrict__ Wt, int Cout,
                                   int Cin, int KHW) {
    // W [Cout,Cin,KHW] -> Wt [Cin,KHW,Cout]
    const size_t total = (size_t)Cout * Cin * KHW;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % KHW);
        const int ci = (int)((i / KHW) % Cin);
        const int co = (int)(i / ((size_t)KHW * Cin));
        Wt[((size_t)ci * KHW + k) * Cout + co] = W[i];
    }
}

size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Carves the caller's workspace.  Layout must match cb_workspace_bytes().
struct Carver {
    char* base;
    size_t off = 0, cap;
    bool dry;
    Carver(void* b, size_t c, bool d) : base((char*)b), cap(c), dry(d) {}
    template <typename T>
    T* take(size_t count) {
        const size_t bytes = align_up(count * sizeof(T));
        T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
        off += bytes;
        return p;
    }
    bool ok() const { return dry || off <= cap; }
};

struct Buffers {
    std::vector<float*> A;      // per node
    std::vector<float*> G;      // per node (mode >= 1)
    std::vector<uint16_t*> Ap;          // packed A per node (tc_pass linears)
    std::vector<uint16_t*> Gp;          // packed dlb/dA per node (inputs of tc_grad linears), mode >= 1
    float* bias_rows = nullptr; // [S*Bd]
    // mode 1/2
    std::vector<float*> grad_alpha, grad_beta;
    std::vector<int> alpha_tab;         // per activation: index of its slope tensor in h_tables, -1 if none
    bool g0_in_pass = false;            // the last run_pass wrote the gradient seed G[0] in its concretize launch
    // mode 2
    float* lb_cur = nullptr;
    float *best_l = nullptr, *best_ret = nullptr, *ret0 = nullptr;
    uint8_t *stopped = nullptr, *mask0 = nullptr, *snap = nullptr;
    cb::OptState* state = nullptr;          // [2]
    cb::RowTable* d_tables = nullptr;
    std::vector<cb::RowTable> h_tables;
    int max_rows = 0, max_cols = 0;
};

int alpha_cols(const cb_plan* p, const cb_problem_t* pr, int k) {
    if (pr && pr->alpha && pr->alpha[k] && pr->n_alpha) return pr->n_alpha[k];
    return (int)p->nodes[p->acts[k]].numel;
}

int beta_cols(const cb_problem_t* pr, int k) {
    if (pr && pr->beta_val && pr->beta_J) return pr->beta_J[k];
    return 0;
}

// mode 0: pass; 1: + gradient; 2: optimisation loop
void carve(const cb_plan* p, int Bd, int S, int mode, const cb_problem_t* pr, Carver& cv,
           Buffers& bf) {
    const size_t rows = (size_t)Bd * S;
    const int nn = (int)p->nodes.size();
    bf.A.assign(nn, nullptr);
    bf.G.assign(nn, nullptr);
    for (int i = 0; i < nn; ++i) {
        const Node& n = p->nodes[i];
        if (!n.on_path) continue;
        if (n.a_alias >= 0) continue;
        bool external = n.act_index >= 0 && pr && pr->lA && pr->lA[n.act_index];
        if (external) bf.A[i] = pr->lA[n.act_index];
        else bf.A[i] = cv.take<float>(rows * n.numel);
    }
    for (int i = 0; i < nn; ++i)
        if (p->nodes[i].on_path && p->nodes[i].a_alias >= 0) bf.A[i] = bf.A[p->nodes[i].a_alias];
    bf.Ap.assign(nn, nullptr);
    bf.Gp.assign(nn, nullptr);
    for (int i = 0; i < nn; ++i) {
        const Node& n = p->nodes[i];
        if (!n.on_path) continue;
        if (n.a_packed) bf.Ap[i] = cv.take<uint16_t>(cb::tc_x_elems((int)rows, (int)n.numel));
        if (mode >= 1 && n.g_packed) bf.Gp[i] = cv.take<uint16_t>(cb::tc_x_elems((int)rows, (int)n.numel));
    }
    bf.bias_rows = cv.take<float>(rows);
    if (mode >= 1) {
        for (int i = 0; i < nn; ++i) {
            const Node& n = p->nodes[i];
            if (!n.need_g) continue;
            if (n.d.op == CB_OP_FLATTEN) continue;
            bf.G[i] = cv.take<float>(rows * n.numel);
        }
        for (int i = 0; i < nn; ++i)
            if (p->nodes[i].need_g && p->nodes[i].d.op == CB_OP_FLATTEN)
                bf.G[i] = bf.G[p->nodes[i].d.in0];
    }
    if (mode >= 2) {
        const int na = (int)p->acts.size();
        const int S1 = pr ? pr->alpha_S1 : S;
        bf.grad_alpha.assign(na, nullptr);
        bf.alpha_tab.assign(na, -1);
        bf.grad_beta.assign(na, nullptr);
        bf.h_tables.clear();
        for (int k = 0; k < na; ++k) {
            const bool has_a = !pr || (pr->alpha && pr->alpha[k]);
            if (has_a) {
                cb::RowTable t;
                const bool ss = p->nodes[p->acts[k]].d.op != CB_OP_RELU;    // [8,S1,Bd,n] tangent points
                t.rows = (ss ? 8 : 1) * S1 * Bd;
                t.cols = alpha_cols(p, pr, k);
                const size_t cnt = (size_t)t.rows * t.cols;
                t.p = pr ? pr->alpha[k] : nullptr;
                t.g = cv.take<float>(cnt);
                t.m = cv.take<float>(cnt);
                t.v = cv.take<float>(cnt);
                t.best = cv.take<float>(cnt);
                t.group = ss ? 2 : 0;
                t.fused = 0;
                bf.alpha_tab[k] = cnt ? (int)bf.h_tables.size() : -1;
                bf.grad_alpha[k] = t.g;
                if (cnt) bf.h_tables.push_back(t);
            }
        }
        for (int k = 0; k < na; ++k) {
            const int J = pr ? beta_cols(pr, k) : 64;
            if (J <= 0 || (pr && !pr->beta_val[k])) continue;
            cb::RowTable t;
            t.rows = Bd;
            t.cols = J;
            const size_t cnt = (size_t)Bd * J;
            t.p = pr ? pr->beta_val[k] : nullptr;
            t.g = cv.take<float>(cnt);
            t.m = cv.take<float>(cnt);
            t.v = cv.take<float>(cnt);
            t.best = cv.take<float>(cnt);
            t.group = 1;
            t.fused = 0;
            bf.grad_beta[k] = t.g;
            bf.h_tables.push_back(t);
        }
        for (auto& t : bf.h_tables) {
            if (t.rows > bf.max_rows) bf.max_rows = t.rows;
            if (t.cols > bf.max_cols) bf.max_cols = t.cols;
        }
        bf.lb_cur = cv.take<float>(rows);
        bf.best_l = cv.take<float>(rows);
        bf.best_ret = cv.take<float>(rows);
        bf.ret0 = cv.take<float>(rows);
        bf.stopped = cv.take<uint8_t>(Bd);
        bf.mask0 = cv.take<uint8_t>(Bd);
        bf.snap = cv.take<uint8_t>(Bd);
        bf.state = cv.take<cb::OptState>(2);
        bf.d_tables = cv.take<cb::RowTable>(2 * p->acts.size() + 1);
    }
}

int check_problem(const cb_plan* p, const cb_problem_t* pr) {
    if (!p || !pr) return fail(CB_ERR_ARG, "null plan/problem");
    if (pr->Bd <= 0 || pr->S <= 0) return fail(CB_ERR_ARG, "Bd and S must be positive");
    if (!pr->C || !pr->x_L || !pr->x_U || !pr->lb) return fail(CB_ERR_ARG, "C/x_L/x_U/lb must be set");
    if (!p->acts.empty() && (!pr->lower || !pr->upper))
        return fail(CB_ERR_ARG, "lower/upper tables must be set");
    for (size_t k = 0; k < p->acts.size(); ++k)
        if (p->nodes[p->acts[k]].on_path && (!pr->lower[k] || !pr->upper[k]))
            return fail(CB_ERR_ARG, "missing intermediate bounds for an activation");
    if (pr->alpha && !(pr->alpha_S1 == 1 || pr->alpha_S1 == pr->S))
        return fail(CB_ERR_ARG, "alpha_S1 must be 1 or S");
    if (pr->alpha && !pr->n_alpha) return fail(CB_ERR_ARG, "n_alpha table must be set with alpha");
    if (pr->beta_val && (!pr->beta_loc || !pr->beta_sign || !pr->beta_J))
        return fail(CB_ERR_ARG, "beta tables incomplete");
    // what can be checked without looking at device memory (the index RANGE of beta_loc is the binding's job)
    for (size_t k = 0; k < p->acts.size(); ++k) {
        const long long n_k = (long long)p->nodes[p->acts[k]].numel;
        if (pr->alpha && pr->alpha[k]) {
            const int na = pr->n_alpha[k];
            if (na < 0 || na > n_k) return fail(CB_ERR_ARG, "n_alpha of a layer must be in [0, neurons of the layer]");
            if (na != n_k && p->nodes[p->acts[k]].d.op == CB_OP_RELU && !(pr->alpha_pos && pr->alpha_pos[k]))
                return fail(CB_ERR_ARG, "sparse alpha (n_alpha < neurons) needs its alpha_pos map");
        }
        if (pr->beta_val) {
            const int J = pr->beta_J[k];
            if (J < 0) return fail(CB_ERR_ARG, "beta_J must not be negative");
            if (J > 0 && pr->beta_val[k] && (!pr->beta_loc[k] || !pr->beta_sign[k]))
                return fail(CB_ERR_ARG, "beta loc / sign missing for a layer with beta values");
        }
    }
    return CB_OK;
}

cb::ReluArgs relu_args(const cb_plan* p, const cb_problem_t* pr, int k) {
    cb::ReluArgs ra;
    ra.lower = pr->lower[k];
    ra.upper = pr->upper[k];
    ra.alpha = (pr->alpha && pr->alpha[k]) ? pr->alpha[k] : nullptr;
    ra.alpha_pos = (ra.alpha && pr->alpha_pos) ? pr->alpha_pos[k] : nullptr;
    ra.n_alpha = ra.alpha ? pr->n_alpha[k] : 0;
    ra.S1 = pr->alpha_S1 > 0 ? pr->alpha_S1 : 1;
    return ra;
}

cb::SshapeArgs sshape_args(const cb_plan* p, const cb_problem_t* pr, int k) {
    const Node& n = p->nodes[p->acts[k]];
    cb::SshapeArgs sa;
    sa.lower = pr->lower[k];
    sa.upper = pr->upper[k];
    sa.alpha = (pr->alpha && pr->alpha[k]) ? pr->alpha[k] : nullptr;
    sa.S1 = pr->alpha_S1 > 0 ? pr->alpha_S1 : 1;
    sa.is_tanh = n.d.op == CB_OP_TANH;
    sa.d_lower_t = n.d.weight;
    sa.d_upper_t = n.d.bias;
    sa.table_n = n.d.kh;
    return sa;
}

// Fills the operand / shape part of a TcArgs for the rows of this call.
cb::TcArgs tc_base(const cb_problem_t* pr, const int* done) {
    cb::TcArgs a;
    memset(&a, 0, sizeof(a));
    a.rows = pr->Bd * pr->S;
    a.Bd = pr->Bd;
    a.S = pr->S;
    a.S1 = 1;
    a.done = done;
    return a;
}

void tc_set_relu(cb::TcArgs& a, const cb::ReluArgs& ra) {
    a.lower = ra.lower;
    a.upper = ra.upper;
    a.alpha = ra.alpha;
    a.alpha_pos = ra.alpha_pos;
    a.n_alpha = ra.n_alpha;
    a.S1 = ra.S1;
}

void tc_set_beta(cb::TcArgs& a, const cb_problem_t* pr, int k, bool use_beta) {
    if (!use_beta || !pr->beta_val || k < 0) return;
    const int J = pr->beta_J[k];
    if (J <= 0 || !pr->beta_val[k]) return;
    a.beta_val = pr->beta_val[k];
    a.beta_loc = pr->beta_loc[k];
    a.beta_sign = pr->beta_sign[k];
    a.beta_bias = pr->beta_bias ? pr->beta_bias[k] : nullptr;
    a.J = J;
}

// One backward pass (auto_LiRPA/backward_bound.py:183-298): reverse topological order, the
// first contribution to a node's A writes, later ones accumulate (add_bound, :691-709).
// Linear nodes marked tc_pass run on the tensor cores fused with the node below (crown_tc.cu);
// everything else takes the SIMT kernels.
bool chain_applies(const cb_plan* p, const cb_problem_t* pr, bool use_beta) {
    if (!p->chain) return false;
    if (use_beta && pr->beta_val && pr->beta_J)
        for (size_t k = 0; k < p->acts.size(); ++k)
            if (pr->beta_val[k] && pr->beta_J[k] > cb::CHAIN_JMAX) return false;
    return true;
}

// The whole pass of a Linear/ReLU chain in one launch (crown_chain.cu).
// keep-best bookkeeping (k_keepbest_a) a pass may do in its own tail: set by cb_optimize, consumed (fused = true) by
// run_pass_chain when the whole-network kernel runs with S == 1
struct KeepBest {
    int iter = 0;
    const float* rhs = nullptr;
    cb::OptState* state = nullptr;
    bool fused = false;
};

int run_pass_chain(const cb_plan* p, const cb_problem_t* pr, Buffers& bf, float* lb_out, bool use_beta,
                   bool keep_lA, const int* done, cudaStream_t st, KeepBest* kb = nullptr) {
    cb::ChainArgs a;
    memset(&a, 0, sizeof(a));
    if (kb != nullptr && kb->state != nullptr && pr->S == 1) {
        a.kb_iter = kb->iter; a.kb_rhs = kb->rhs; a.kb_state = kb->state;
        a.kb_best_l = bf.best_l; a.kb_best_ret = bf.best_ret; a.kb_ret0 = bf.ret0;
        a.kb_stopped = bf.stopped; a.kb_mask0 = bf.mask0;
        kb->fused = true;
    }
    a.rows = pr->Bd * pr->S; a.Bd = pr->Bd; a.S = pr->S;
    a.S1 = pr->alpha_S1 > 0 ? pr->alpha_S1 : 1;
    a.n_steps = (int)p->chain_lin.size();
    for (int j = 0; j < a.n_steps; ++j) {
        const Node& lin = p->nodes[p->chain_lin[j]];
        cb::ChainStep& s = a.step[j];
        s.wp = lin.wp_chain;
        s.M = (int)p->nodes[lin.d.in0].numel;
        s.Kp = cb::tc_kp((int)lin.numel);
        const int R = p->chain_relu[j];
        if (R < 0) continue;
        const Node& r = p->nodes[R];
        const int k = r.act_index;
        const cb::ReluArgs ra = relu_args(p, pr, k);
        s.lower = ra.lower; s.upper = ra.upper;
        s.alpha = ra.alpha; s.alpha_pos = ra.alpha_pos; s.n_alpha = ra.n_alpha;
        s.bias_below = p->nodes[r.d.in0].d.bias;
        s.lA = (keep_lA || (pr->lA && pr->lA[k])) ? bf.A[R] : nullptr;
        if (use_beta && pr->beta_val && pr->beta_J[k] > 0 && pr->beta_val[k]) {
            s.beta_val = pr->beta_val[k]; s.beta_loc = pr->beta_loc[k]; s.beta_sign = pr->beta_sign[k];
            s.beta_bias = pr->beta_bias ? pr->beta_bias[k] : nullptr;
            s.J = pr->beta_J[k];
        }
    }
    a.C = pr->C; a.n_out = p->n_out;
    a.b_out = p->nodes[p->chain_lin[0]].d.bias;
    a.x_L = pr->x_L; a.x_U = pr->x_U;
    a.lb = lb_out;
    a.g0_plain = bf.G[0];          // null in pass-only mode
    a.done = done;
    CB_CUDA(cb::chain_pass(a, st));
    return CB_OK;
}

// The whole gradient of a Linear/ReLU chain in one launch (crown_chain_grad.cu); needs the lA stash and the
// worst-case input point G[0] written by run_pass_chain.
bool chain_grad_applies(const cb_plan* p, const cb_problem_t* pr, bool use_beta) {
    return p->chain_grad && pr->S == 1 && chain_applies(p, pr, use_beta);
}

// second half of the keep-best bookkeeping (k_keepbest_b) a gradient launch may do in its own head: set by
// cb_optimize, consumed (fused = true) by run_grad_chain
struct KeepBestB {
    int iter = 0, iteration = 0, save_from = 0, patience_limit = 0;
    const cb::OptState* cur = nullptr;
    cb::OptState* next = nullptr;
    bool fused = false;
};

int run_grad_chain(const cb_plan* p, const cb_problem_t* pr, Buffers& bf, float* const* grad_alpha,
                   float* const* grad_beta, bool use_beta, const int* done, cudaStream_t st, KeepBestB* kb = nullptr) {
    cb::ChainGradArgs a;
    memset(&a, 0, sizeof(a));
    const int nl = (int)p->chain_lin.size();
    a.rows = pr->Bd;
    a.n_steps = nl - 1;
    for (int s = 0; s < a.n_steps; ++s) {
        const Node& lin = p->nodes[p->chain_lin[nl - 1 - s]];
        const int R = p->chain_relu[nl - 2 - s];              // the ReLU that consumes this Linear
        const Node& r = p->nodes[R];
        const int k = r.act_index;
        cb::GradStep& g = a.step[s];
        g.wp = lin.wp_chain_g;
        g.M = (int)lin.numel;
        g.Kp = cb::tc_kp((int)p->nodes[lin.d.in0].numel);
        g.bias = lin.d.bias;
        const cb::ReluArgs ra = relu_args(p, pr, k);
        g.lower = ra.lower; g.upper = ra.upper;
        g.alpha = ra.alpha; g.alpha_pos = ra.alpha_pos; g.n_alpha = ra.n_alpha;
        g.a_post = bf.A[R];
        g.grad_alpha = (grad_alpha && ra.alpha) ? grad_alpha[k] : nullptr;
        if (use_beta && grad_beta && grad_beta[k] && pr->beta_val && pr->beta_J[k] > 0 && pr->beta_val[k]) {
            g.beta_loc = pr->beta_loc[k]; g.beta_sign = pr->beta_sign[k];
            g.beta_bias = pr->beta_bias ? pr->beta_bias[k] : nullptr;
            g.grad_beta = grad_beta[k];
            g.J = pr->beta_J[k];
        }
        g.need_y = s + 1 < a.n_steps;
    }
    a.g0 = bf.G[0];
    a.n_in = p->n_in;
    a.done = done;
    if (kb != nullptr && kb->cur != nullptr) {
        a.kb_iter = kb->iter; a.kb_iteration = kb->iteration; a.kb_save_from = kb->save_from;
        a.kb_patience_limit = kb->patience_limit;
        a.kb_lb_cur = bf.lb_cur; a.kb_ret0 = bf.ret0; a.kb_mask0 = bf.mask0; a.kb_snap = bf.snap;
        a.kb_cur = kb->cur; a.kb_next = kb->next;
        kb->fused = true;
    }
    CB_CUDA(cb::chain_grad(a, st));
    return CB_OK;
}

int run_pass(const cb_plan* p, const cb_problem_t* pr, Buffers& bf, float* lb_out, bool use_beta,
             bool keep_lA, const int* done, cudaStream_t st, KeepBest* kb = nullptr) {
    if (chain_applies(p, pr, use_beta)) return run_pass_chain(p, pr, bf, lb_out, use_beta, keep_lA, done, st, kb);
    const int nn = (int)p->nodes.size();
    const int Bd = pr->Bd, S = pr->S;
    const int rows = Bd * S;
    std::vector<char> written(nn, 0), packed(nn, 0), beta_done(nn, 0);
    bf.g0_in_pass = false;
    bool concretized = false;
    cb::fill_zero(bf.bias_rows, rows, done, st);
    if (p->nodes[nn - 1].a_packed) {
        // the output node is a tc_pass linear: C goes straight into the packed operand format
        const Node& o = p->nodes[nn - 1];
        cb::tc_pack_rows(pr->C, true, rows, Bd, S, p->n_out, cb::tc_kp(p->n_out), bf.Ap[nn - 1], o.d.bias,
                         bf.bias_rows, done, st);
        packed[nn - 1] = 1;
    } else {
        cb::spec_to_rows(pr->C, bf.A[nn - 1], Bd, S, p->n_out, done, st);
    }
    written[nn - 1] = 1;
    for (int idx = nn - 1; idx >= 1; --idx) {
        const Node& n = p->nodes[idx];
        if (!n.on_path || !written[idx]) continue;
        float* a = bf.A[idx];
        if (use_beta && n.preact_index >= 0 && idx != nn - 1 && pr->beta_val && !beta_done[idx]) {
            const int k = n.preact_index;
            const int J = pr->beta_J[k];
            if (J > 0 && pr->beta_val[k])
                cb::beta_scatter(a, bf.bias_rows, pr->beta_val[k], pr->beta_loc[k], pr->beta_sign[k],
                                 pr->beta_bias ? pr->beta_bias[k] : nullptr, J, Bd, S, (int)n.numel,
                                 done, st);
        }
        const int i0 = n.d.in0, i1 = n.d.in1;
        if (n.d.op == CB_OP_LINEAR && n.tc_pass) {
            if (!packed[idx]) {
                // A of this node was produced by SIMT kernels: split it into the packed planes
                cb::tc_pack_rows(a, false, rows, Bd, S, (int)n.numel, cb::tc_kp((int)n.numel), bf.Ap[idx],
                                 n.d.bias, bf.bias_rows, done, st);
                packed[idx] = 1;
            }
            const Node& dst = p->nodes[n.tc_dst];
            cb::TcArgs ta = tc_base(pr, done);
            ta.xp = bf.Ap[idx];
            ta.wp = n.wp_pass;
            ta.N = (int)dst.numel; ta.Kp = cb::tc_kp((int)n.numel); ta.BN = n.bn_pass;
            ta.bias_rows = bf.bias_rows;
            if (n.tc_pass == 1) {
                const Node& r = p->nodes[n.tc_relu];
                tc_set_relu(ta, relu_args(p, pr, r.act_index));
                tc_set_beta(ta, pr, dst.preact_index, use_beta);
                ta.lA = (keep_lA || (pr->lA && pr->lA[r.act_index])) ? bf.A[n.tc_relu] : nullptr;
                if (dst.a_packed) {
                    ta.yp = bf.Ap[n.tc_dst];
                    ta.y_Kp = cb::tc_kp((int)dst.numel);
                    ta.blin = dst.d.bias;          // the consumer GEMM does not do its own bias dot
                    packed[n.tc_dst] = 1;
                }
                if (dst.a_plain) ta.y_plain = bf.A[n.tc_dst];
                CB_CUDA(cb::tc_linear(cb::TC_MODE_RELAX, ta, st));
                beta_done[n.tc_dst] = 1;
                written[n.tc_dst] = 1;
            } else {
                ta.x_L = pr->x_L; ta.x_U = pr->x_U;
                const Node& in = p->nodes[0];
                if (bf.Gp[0] && in.g_packed) {
                    ta.yp = bf.Gp[0];
                    ta.y_Kp = cb::tc_kp((int)in.numel);
                }
                if (bf.G[0] && in.g_plain) ta.y_plain = bf.G[0];
                CB_CUDA(cb::tc_linear(cb::TC_MODE_CONCRETIZE, ta, st));
                written[0] = 1;
                concretized = true;
            }
            continue;
        }
        switch (n.d.op) {
            case CB_OP_LINEAR: {
                const int K = (int)n.numel, N = (int)p->nodes[i0].numel;
                cb::sgemm(false, a, n.d.weight, bf.A[i0], rows, N, K, written[i0], n.d.bias,
                          bf.bias_rows, nullptr, done, st);
                written[i0] = 1;
                break;
            }
            case CB_OP_CONV2D: {
                const cb::ConvGeom g = conv_geom(p, n);
                if (n.ct_use_pass) {
                    // tcgen05 implicit GEMM; the bias dot product rides along in its loader
                    CB_CUDA(cb::conv_tc(n.ct_pass, a, bf.A[i0], n.wct_pass, n.d.bias, bf.bias_rows, rows, written[i0], done, st));
                    written[i0] = 1;
                    break;
                }
                if (!cb::conv_bwd_tiled(a, n.wk_b, bf.A[i0], g, rows, written[i0], done, st))
                    cb::conv_bwd(a, n.wt, bf.A[i0], g, rows, written[i0], done, st);
                if (n.d.bias) cb::chan_rowdot(a, n.d.bias, bf.bias_rows, rows, n.d.c, n.d.h * n.d.w, done, st);
                written[i0] = 1;
                break;
            }
            case CB_OP_BATCHNORM2D: {
                cb::chan_affine(a, bf.A[i0], n.d.weight, nullptr, rows, n.d.c, n.d.h * n.d.w,
                                written[i0], done, st);
                cb::chan_rowdot(a, n.d.bias, bf.bias_rows, rows, n.d.c, n.d.h * n.d.w, done, st);
                written[i0] = 1;
                break;
            }
            case CB_OP_ADD:
            case CB_OP_SUB: {
                const size_t cnt = (size_t)rows * n.numel;
                if (bf.A[i0] != a) cb::axpy(a, bf.A[i0], 1.f, cnt, written[i0], done, st);       // else: shared storage
                written[i0] = 1;
                if (bf.A[i1] != a) cb::axpy(a, bf.A[i1], n.d.op == CB_OP_ADD ? 1.f : -1.f, cnt, written[i1], done, st);
                written[i1] = 1;
                break;
            }
            case CB_OP_ADDCONST:
                // unperturbed operand: lb += A . value (backward_bound.py:712-721); A passes through
                cb::chan_rowdot(a, n.d.bias, bf.bias_rows, rows, (int)n.numel, 1, done, st);
                // fall through
            case CB_OP_FLATTEN: {
                if (bf.A[i0] != a)
                    cb::axpy(a, bf.A[i0], 1.f, (size_t)rows * n.numel, written[i0], done, st);
                written[i0] = 1;
                break;
            }
            case CB_OP_RELU: {
                const cb::ReluArgs ra = relu_args(p, pr, n.act_index);
                // the split constraints of the pre-activation node ride in the same launch (the stand-alone
                // beta_scatter at that node's visit is then skipped)
                cb::BetaScatter bs;
                const bool beta_in_relu = getenv("CROWN_B200_DISABLE_BETA_IN_RELU") == nullptr;      // read per call: tests toggle it
                const Node& pre = p->nodes[i0];
                if (beta_in_relu && use_beta && pre.preact_index >= 0 && i0 != nn - 1 && pr->beta_val && !beta_done[i0]) {
                    const int k = pre.preact_index;
                    if (pr->beta_J[k] > 0 && pr->beta_val[k]) {
                        bs.val = pr->beta_val[k]; bs.loc = pr->beta_loc[k]; bs.sign = pr->beta_sign[k];
                        bs.bias = pr->beta_bias ? pr->beta_bias[k] : nullptr;
                        bs.J = pr->beta_J[k];
                        beta_done[i0] = 1;
                    }
                }
                cb::relu_bwd(a, bf.A[i0], written[i0], bf.bias_rows, ra, Bd, S, (int)n.numel, done, st,
                             bs.J > 0 ? &bs : nullptr);
                written[i0] = 1;
                break;
            }
            case CB_OP_SIGMOID:
            case CB_OP_TANH: {
                const cb::SshapeArgs sa = sshape_args(p, pr, n.act_index);
                cb::sshape_clip(sa, Bd, (int)n.numel, done, st);
                cb::sshape_bwd(a, bf.A[i0], written[i0], bf.bias_rows, sa, Bd, S, (int)n.numel, done, st);
                written[i0] = 1;
                break;
            }
            default:
                return fail(CB_ERR_ARG, "operator not supported by the CUDA path yet");
        }
    }
    if (!written[0]) return fail(CB_ERR_ARG, "the input node is not reachable from the output");
    if (concretized)
        cb::rows_to_lb(bf.bias_rows, lb_out, Bd, S, done, st);
    else {
        // when a gradient sweep can follow (its buffers exist), the same launch writes the sweep's seed G[0]
        float* g0 = (getenv("CROWN_B200_DISABLE_SEED_IN_CONCRETIZE") == nullptr && !bf.G.empty()) ? bf.G[0] : nullptr;
        cb::concretize(bf.A[0], pr->x_L, pr->x_U, bf.bias_rows, lb_out, Bd, S, p->n_in, done, st, g0);
        bf.g0_in_pass = g0 != nullptr;
    }
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// Gradient of sum lb w.r.t. alpha / beta_val.  d lb / d A flows input -> output through the same
// operators transposed; it equals evaluating the network at the worst-case input with each ReLU
// replaced by the line the sign of A selected (operators/clampmult.py:49-95).
// The Adam step of the ReLU slopes whose gradient comes from a stand-alone relu_grad launch runs inside that launch.
struct AdamInGrad {
    std::vector<char> on;                 // per activation
    const uint8_t* stopped = nullptr;
    const uint8_t* snap = nullptr;
    float step = 0.f, bc2_sqrt = 1.f;
};

// Does the gradient of activation k's slopes come from cb::relu_grad (not from a fused linear+ReLU tensor-core
// launch, not from the whole-network chain kernel)?  Mirrors the dispatch of run_grad below.
bool relu_grad_standalone(const cb_plan* p, const cb_problem_t* pr, bool use_beta, int k) {
    if (chain_grad_applies(p, pr, use_beta)) return false;
    const int R = p->acts[k];
    if (p->nodes[R].d.op != CB_OP_RELU || !p->nodes[R].on_path) return false;
    for (const Node& n : p->nodes)
        if (n.on_path && n.d.op == CB_OP_LINEAR && n.tc_grad && n.need_g && n.tc_grad_relu == R) return false;
    return true;
}

int run_grad(const cb_plan* p, const cb_problem_t* pr, Buffers& bf, float* const* grad_alpha,
             float* const* grad_beta, bool use_beta, const int* done, cudaStream_t st, KeepBestB* kb = nullptr,
             const AdamInGrad* aig = nullptr) {
    const int nn = (int)p->nodes.size();
    const int Bd = pr->Bd, S = pr->S;
    const int rows = Bd * S;
    std::vector<char> handled(nn, 0), gpacked(nn, 0), beta_deferred(nn, 0);
    const bool beta_in_relu = getenv("CROWN_B200_DISABLE_BETA_IN_RELU") == nullptr;      // read per call: tests toggle it
    const Node& first = p->nodes[0];
    bool g0_from_pass = false;
    for (const Node& n : p->nodes)
        if (n.on_path && n.d.op == CB_OP_LINEAR && n.tc_pass == 2) g0_from_pass = true;
    if (chain_grad_applies(p, pr, use_beta))
        return run_grad_chain(p, pr, bf, grad_alpha, grad_beta, use_beta, done, st, kb);
    if (chain_applies(p, pr, use_beta)) {
        gpacked[0] = 0;                           // the chain pass wrote the plain seed into G[0]
    } else if (g0_from_pass) {
        gpacked[0] = first.g_packed ? 1 : 0;      // written by the concretize epilogue of the pass
    } else if (bf.g0_in_pass) {
        gpacked[0] = 0;                           // plain seed written by the concretize kernel of the pass
    } else {
        cb::grad_init(bf.A[0], pr->x_L, pr->x_U, bf.G[0], Bd, S, p->n_in, done, st);
    }
    for (int idx = 1; idx < nn; ++idx) {
        const Node& n = p->nodes[idx];
        if (!n.on_path) continue;
        const int i0 = n.d.in0, i1 = n.d.in1;
        if (n.d.op == CB_OP_LINEAR && n.tc_grad && n.need_g) {
            const int R = n.tc_grad_relu;
            const Node& r = p->nodes[R];
            const int k = r.act_index;
            const cb::ReluArgs ra = relu_args(p, pr, k);
            float* ga = (grad_alpha && ra.alpha) ? grad_alpha[k] : nullptr;
            float* gb = nullptr;
            cb::TcArgs ta = tc_base(pr, done);
            tc_set_relu(ta, ra);
            if (use_beta && grad_beta && grad_beta[k]) {
                tc_set_beta(ta, pr, k, true);
                if (ta.J > 0) gb = grad_beta[k];
            }
            handled[R] = 1;
            if (!r.need_g && !ga && !gb) continue;          // nothing upstream needs this layer
            const int src = n.tc_src;
            const Node& sn = p->nodes[src];
            if (!gpacked[src]) {
                cb::tc_pack_rows(bf.G[src], false, rows, Bd, S, (int)sn.numel, cb::tc_kp((int)sn.numel), bf.Gp[src],
                                 nullptr, nullptr, done, st);
                gpacked[src] = 1;
            }
            ta.xp = bf.Gp[src];
            ta.wp = n.wp_grad;
            ta.N = (int)n.numel; ta.Kp = cb::tc_kp((int)sn.numel); ta.BN = n.bn_grad;
            ta.col_bias = n.d.bias;
            ta.a_post = bf.A[R];
            ta.grad_alpha = ga;
            ta.grad_beta = gb;
            if (S > 1) {   // rows of different spec index accumulate into the same entry
                if (ga && ra.S1 == 1) cb::fill_zero(ga, (size_t)Bd * ra.n_alpha, done, st);
                if (gb) cb::fill_zero(gb, (size_t)Bd * ta.J, done, st);
            }
            if (r.need_g) {
                if (r.g_packed) {
                    ta.yp = bf.Gp[R];
                    ta.y_Kp = cb::tc_kp((int)r.numel);
                    gpacked[R] = 1;
                }
                if (r.g_plain) ta.y_plain = bf.G[R];
            }
            CB_CUDA(cb::tc_linear(cb::TC_MODE_GRAD, ta, st));
            continue;
        }
        if (is_act(n.d.op)) {
            if (handled[idx]) continue;
            const int k = n.act_index;
            float* ga = grad_alpha ? grad_alpha[k] : nullptr;
            if (n.d.op != CB_OP_RELU) {
                const cb::SshapeArgs sa = sshape_args(p, pr, k);
                if (n.need_g || (ga && sa.alpha))
                    cb::sshape_grad(bf.A[idx], bf.G[i0], n.need_g ? bf.G[idx] : nullptr, ga, sa, Bd, S,
                                    (int)n.numel, done, st);
                continue;
            }
            const cb::ReluArgs ra = relu_args(p, pr, k);
            cb::AdamFuse af;
            const bool fuse = aig != nullptr && aig->on[k] && ga && ra.alpha && bf.alpha_tab[k] >= 0;
            if (fuse) {
                const cb::RowTable& t = bf.h_tables[bf.alpha_tab[k]];
                af.p = t.p; af.m = t.m; af.v = t.v; af.best = t.best;
                af.stopped = aig->stopped; af.snap = aig->snap; af.step = aig->step; af.bc2_sqrt = aig->bc2_sqrt;
            }
            cb::BetaGrad bg;
            if (beta_deferred[idx]) {
                const int kb = p->nodes[i0].preact_index;
                bg.grad_val = grad_beta[kb]; bg.loc = pr->beta_loc[kb]; bg.sign = pr->beta_sign[kb];
                bg.bias = pr->beta_bias ? pr->beta_bias[kb] : nullptr;
                bg.J = pr->beta_J[kb];
            }
            if (n.need_g || (ga && ra.alpha))
                cb::relu_grad(bf.A[idx], bf.G[i0], n.need_g ? bf.G[idx] : nullptr, ga, ra, Bd, S,
                              (int)n.numel, done, st, fuse ? &af : nullptr, bg.J > 0 ? &bg : nullptr);
            else if (bg.J > 0)
                return fail(CB_ERR_ARG, "deferred beta gradient without its relu_grad launch");
            continue;
        }
        if (!n.need_g) continue;
        switch (n.d.op) {
            case CB_OP_LINEAR: {
                const int K = (int)p->nodes[i0].numel, N = (int)n.numel;
                cb::sgemm(true, bf.G[i0], n.d.weight, bf.G[idx], rows, N, K, false, nullptr, nullptr,
                          n.d.bias, done, st);
                break;
            }
            case CB_OP_CONV2D:
                if (n.ct_use_grad) {
                    CB_CUDA(cb::conv_tc(n.ct_grad, bf.G[i0], bf.G[idx], n.wct_grad, n.d.bias, nullptr, rows, false, done, st));
                    break;
                }
                if (!cb::conv_fwd_tiled(bf.G[i0], n.wk_f, n.d.bias, bf.G[idx], conv_geom(p, n), rows, done, st))
                    cb::conv_fwd(bf.G[i0], n.d.weight, n.d.bias, bf.G[idx], conv_geom(p, n), rows, done, st);
                break;
            case CB_OP_BATCHNORM2D:
                cb::chan_affine(bf.G[i0], bf.G[idx], n.d.weight, n.d.bias, rows, n.d.c, n.d.h * n.d.w,
                                false, done, st);
                break;
            case CB_OP_ADD:
            case CB_OP_SUB:
                cb::add2(bf.G[i0], bf.G[i1], bf.G[idx], n.d.op == CB_OP_ADD ? 1.f : -1.f,
                         (size_t)rows * n.numel, done, st);
                break;
            case CB_OP_ADDCONST:
                cb::add_rowvec(bf.G[i0], n.d.bias, bf.G[idx], rows, (int)n.numel, done, st);
                break;
            case CB_OP_FLATTEN:
                break;  // alias
            default:
                return fail(CB_ERR_ARG, "operator not supported by the CUDA path yet");
        }
        if (use_beta && n.preact_index >= 0 && grad_beta && pr->beta_val) {
            const int k = n.preact_index;
            const int J = pr->beta_J[k];
            if (J > 0 && grad_beta[k]) {
                // rides in the relu_grad launch of the activation this node feeds whenever that launch is certain
                // (same condition as at the activation's visit below)
                const int R = p->acts[k];
                const Node& rn = p->nodes[R];
                const bool ga_k = grad_alpha && grad_alpha[k] && pr->alpha && pr->alpha[k];
                const bool in_relu = beta_in_relu && rn.d.op == CB_OP_RELU && rn.d.in0 == idx && R > idx && rn.on_path &&
                                     !handled[R] && (rn.need_g || ga_k);
                if (in_relu)
                    beta_deferred[R] = 1;
                else
                    cb::beta_grad(bf.G[idx], grad_beta[k], pr->beta_loc[k], pr->beta_sign[k],
                                  pr->beta_bias ? pr->beta_bias[k] : nullptr, J, Bd, S, (int)n.numel,
                                  done, st);
            }
        }
    }
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

}  // namespace

extern "C" {

const char* cb_last_error(void) { return g_err.c_str(); }
int cb_version(void) { return 1; }

void cb_profile_enable(int32_t on) { cb::profile_enable(on != 0); }
int64_t cb_launch_count(void) { return (int64_t)cb::launch_count(); }
int32_t cb_profile_num_kernels(void) { return (int32_t)cb::K_COUNT; }
const char* cb_profile_kernel_name(int32_t id) { return cb::kernel_name(id); }
int32_t cb_profile_collect(double* h_ms, int64_t* h_launches, int32_t n) {
    return cb::profile_collect(h_ms, reinterpret_cast<long long*>(h_launches), n);
}

int cb_plan_create(const cb_node_t* h_nodes, int32_t n_nodes, cb_plan_t** out_plan) {
    if (!h_nodes || n_nodes < 2 || !out_plan) return fail(CB_ERR_ARG, "bad arguments");
    if (h_nodes[0].op != CB_OP_INPUT) return fail(CB_ERR_ARG, "node 0 must be the input");
    cb_plan* p = new cb_plan();
    p->nodes.resize(n_nodes);
    for (int i = 0; i < n_nodes; ++i) {
        Node& n = p->nodes[i];
        n.d = h_nodes[i];
        n.numel = (int64_t)n.d.c * n.d.h * n.d.w;
        if (n.numel <= 0) { delete p; return fail(CB_ERR_ARG, "node with empty shape"); }
        if (i > 0) {
            if (n.d.in0 < 0 || n.d.in0 >= i) { delete p; return fail(CB_ERR_ARG, "inputs must precede the node"); }
            p->nodes[n.d.in0].consumers.push_back(i);
            const bool two = n.d.op == CB_OP_ADD || n.d.op == CB_OP_SUB;
            if (two) {
                if (n.d.in1 < 0 || n.d.in1 >= i) { delete p; return fail(CB_ERR_ARG, "inputs must precede the node"); }
                p->nodes[n.d.in1].consumers.push_back(i);
                if (p->nodes[n.d.in0].numel != n.numel || p->nodes[n.d.in1].numel != n.numel) {
                    delete p; return fail(CB_ERR_ARG, "add/sub with broadcasting is not supported");
                }
            }
        }
        switch (n.d.op) {
            case CB_OP_INPUT: break;
            case CB_OP_LINEAR:
                if (!n.d.weight) { delete p; return fail(CB_ERR_ARG, "linear without weight"); }
                break;
            case CB_OP_CONV2D:
                if (!n.d.weight || n.d.groups != 1) { delete p; return fail(CB_ERR_ARG, "conv2d needs weight and groups==1"); }
                break;
            case CB_OP_BATCHNORM2D:
                if (!n.d.weight || !n.d.bias) { delete p; return fail(CB_ERR_ARG, "batchnorm needs folded scale and shift"); }
                break;
            case CB_OP_ADD: case CB_OP_SUB: case CB_OP_FLATTEN: break;
            case CB_OP_ADDCONST:
                if (!n.d.bias || p->nodes[n.d.in0].numel != n.numel) { delete p; return fail(CB_ERR_ARG, "addconst needs a value of the node's shape"); }
                break;
            case CB_OP_RELU: case CB_OP_SIGMOID: case CB_OP_TANH:
                if (n.d.op != CB_OP_RELU && (!n.d.weight || !n.d.bias || n.d.kh < 2)) {
                    delete p;
                    return fail(CB_ERR_ARG, "sigmoid/tanh need the tangent tables (weight = d_lower, bias = d_upper, kh = length)");
                }
                n.act_index = (int)p->acts.size();
                p->acts.push_back(i);
                p->nodes[n.d.in0].preact_index = n.act_index;
                break;
            default: delete p; return fail(CB_ERR_ARG, "unknown operator");
        }
    }
    // reachability from the output
    p->nodes[n_nodes - 1].on_path = true;
    for (int i = n_nodes - 1; i >= 1; --i) {
        Node& n = p->nodes[i];
        if (!n.on_path) continue;
        p->nodes[n.d.in0].on_path = true;
        if (n.d.op == CB_OP_ADD || n.d.op == CB_OP_SUB) p->nodes[n.d.in1].on_path = true;
    }
    // which dlb/dA are needed: pre-activation nodes and everything upstream of them
    for (int k = 0; k < (int)p->acts.size(); ++k)
        if (p->nodes[p->acts[k]].on_path) p->nodes[p->nodes[p->acts[k]].d.in0].need_g = true;
    for (int i = n_nodes - 1; i >= 1; --i) {
        Node& n = p->nodes[i];
        if (!n.need_g) continue;
        p->nodes[n.d.in0].need_g = true;
        if (n.d.op == CB_OP_ADD || n.d.op == CB_OP_SUB) p->nodes[n.d.in1].need_g = true;
    }
    // flatten shares its input's A buffer when it is the input's only consumer (the input owns
    // the storage, which may be a caller-provided lA tensor when the input is an activation)
    for (int i = 1; i < n_nodes; ++i) {
        Node& n = p->nodes[i];
        if ((n.d.op == CB_OP_FLATTEN || n.d.op == CB_OP_ADDCONST) && p->nodes[n.d.in0].consumers.size() == 1) {
            int a = n.d.in0;
            while (p->nodes[a].a_alias >= 0) a = p->nodes[a].a_alias;
            n.a_alias = a;
        }
    }
    // Add / Sub hand their A on unchanged (operators/add_sub.py:19-29): an input whose only consumer is the Add reads
    // the Add's own buffer instead of a copy (the second operand of a Sub needs the negated copy)
    for (int i = n_nodes - 1; i >= 1; --i) {
        Node& n = p->nodes[i];
        if (n.d.op != CB_OP_ADD && n.d.op != CB_OP_SUB) continue;
        int owner = i;
        while (p->nodes[owner].a_alias >= 0) owner = p->nodes[owner].a_alias;
        const int ins[2] = {n.d.in0, n.d.op == CB_OP_ADD ? n.d.in1 : -1};
        for (int j : ins) {
            if (j <= 0) continue;
            Node& src = p->nodes[j];
            if (src.consumers.size() != 1 || src.a_alias >= 0 || src.act_index >= 0) continue;
            bool aliased_by_other = false;            // a flatten / addconst that already shares src's storage
            for (const Node& m : p->nodes) aliased_by_other |= (m.a_alias == j);
            if (aliased_by_other) continue;
            src.a_alias = owner;
        }
    }
    // ---- tensor-core eligibility (crown_tc.cu) --------------------------------------------------
    {
        const char* env = getenv("CROWN_B200_DISABLE_TC");
        p->use_tc = !(env && env[0] == '1');
        auto single = [&](int i) { return p->nodes[i].consumers.size() == 1; };
        auto root = [&](int i) {          // skip flattens: they alias their input's storage
            while (p->nodes[i].d.op == CB_OP_FLATTEN) i = p->nodes[i].d.in0;
            return i;
        };
        std::vector<char> fused_relu(n_nodes, 0);
        for (int idx = 1; idx < n_nodes && p->use_tc; ++idx) {
            Node& n = p->nodes[idx];
            if (!n.on_path || n.d.op != CB_OP_LINEAR) continue;
            n.tc_src = root(n.d.in0);
            // pass direction: the chain linear -> (flattens) -> relu -> pre-activation node must be private
            int i = n.d.in0;
            bool priv = true;
            while (p->nodes[i].d.op == CB_OP_FLATTEN) {
                if (!single(i)) { priv = false; break; }
                i = p->nodes[i].d.in0;
            }
            if (priv && single(i)) {
                if (i == 0) {
                    n.tc_pass = 2;
                    n.tc_dst = 0;
                } else if (p->nodes[i].d.op == CB_OP_RELU && single(p->nodes[i].d.in0)) {
                    n.tc_pass = 1;
                    n.tc_relu = i;
                    n.tc_dst = p->nodes[i].d.in0;
                }
            }
            // gradient direction: linear -> relu
            if (n.need_g && single(idx)) {
                const int c = n.consumers[0];
                if (p->nodes[c].d.op == CB_OP_RELU && p->nodes[c].on_path) {
                    n.tc_grad = 1;
                    n.tc_grad_relu = c;
                    fused_relu[c] = 1;
                }
            }
        }
        for (int idx = 0; idx < n_nodes; ++idx) {
            Node& n = p->nodes[idx];
            n.a_packed = n.tc_pass != 0;
            n.a_plain = !n.a_packed;
            n.g_packed = false;
            n.g_plain = false;
        }
        for (int c = 1; c < n_nodes; ++c) {
            Node& n = p->nodes[c];
            if (!n.on_path) continue;
            if (n.d.op == CB_OP_LINEAR && n.tc_grad) { p->nodes[n.tc_src].g_packed = true; continue; }
            if (n.d.op == CB_OP_FLATTEN) continue;
            const bool reads_g = is_act(n.d.op) ? !fused_relu[c] : n.need_g;
            if (!reads_g) continue;
            p->nodes[root(n.d.in0)].g_plain = true;
            if (n.d.op == CB_OP_ADD || n.d.op == CB_OP_SUB) p->nodes[root(n.d.in1)].g_plain = true;
        }
        for (int idx = 1; idx < n_nodes; ++idx) {
            Node& n = p->nodes[idx];
            if (n.d.op != CB_OP_LINEAR || (!n.tc_pass && !n.tc_grad)) continue;
            const int out_f = (int)n.numel, in_f = (int)p->nodes[n.d.in0].numel;
            cudaError_t e = cudaSuccess;
            if (n.tc_pass) {          // D[rows,in] = A[rows,out] . W  :  B(n=in,k=out) = W[k*in + n]
                n.bn_pass = cb::tc_pick_bn(in_f, n.tc_pass == 2 ? 128 : 64);
                e = cudaMalloc(&n.wp_pass, cb::tc_w_elems(in_f, out_f, n.bn_pass) * sizeof(uint16_t));
                if (e == cudaSuccess)
                    cb::tc_pack_weight(n.d.weight, 1, in_f, in_f, out_f, cb::tc_kp(out_f), n.bn_pass, n.wp_pass, 0);
            }
            if (e == cudaSuccess && n.tc_grad) {   // G[rows,out] = G[rows,in] . W^T : B(n=out,k=in) = W[n*in + k]
                n.bn_grad = cb::tc_pick_bn(out_f, 64);
                e = cudaMalloc(&n.wp_grad, cb::tc_w_elems(out_f, in_f, n.bn_grad) * sizeof(uint16_t));
                if (e == cudaSuccess)
                    cb::tc_pack_weight(n.d.weight, in_f, 1, out_f, in_f, cb::tc_kp(in_f), n.bn_grad, n.wp_grad, 0);
            }
            if (e != cudaSuccess) {
                delete p;
                return fail(e == cudaErrorMemoryAllocation ? CB_ERR_OOM : CB_ERR_CUDA,
                            std::string("cudaMalloc(packed weight): ") + cudaGetErrorString(e));
            }
        }
    }
    p->n_in = (int)p->nodes[0].numel;
    p->n_out = (int)p->nodes[n_nodes - 1].numel;
    for (auto& n : p->nodes) p->sum_numel += n.numel;
    // ---- Linear/ReLU chain: input -> (flatten)* -> Linear -> ReLU -> ... -> Linear ----------------
    {
        const char* env = getenv("CROWN_B200_DISABLE_CHAIN");
        bool ok = p->use_tc && !(env && env[0] == '1');
        std::vector<int> lins, relus;
        int visited = 1, idx = n_nodes - 1;
        while (ok) {
            const Node& lin = p->nodes[idx];
            if (lin.d.op != CB_OP_LINEAR || lin.numel > cb::CHAIN_KMAX || (idx != n_nodes - 1 && lin.consumers.size() != 1)) { ok = false; break; }
            ++visited;
            int i = lin.d.in0;
            while (i > 0 && p->nodes[i].d.op == CB_OP_FLATTEN && p->nodes[i].consumers.size() == 1) { i = p->nodes[i].d.in0; ++visited; }
            lins.push_back(idx);
            if (i == 0) { relus.push_back(-1); break; }
            const Node& r = p->nodes[i];
            if (r.d.op != CB_OP_RELU || r.consumers.size() != 1 || r.numel > cb::CHAIN_KMAX) { ok = false; break; }
            ++visited;
            relus.push_back(i);
            idx = r.d.in0;
        }
        ok = ok && visited == n_nodes && lins.size() >= 2 && (int)lins.size() <= cb::CHAIN_MAX_STEPS &&
             p->nodes[0].consumers.size() == 1;
        if (ok) {
            for (int li : lins) {
                Node& n = p->nodes[li];
                const int out_f = (int)n.numel, in_f = (int)p->nodes[n.d.in0].numel;
                cudaError_t e = cudaMalloc(&n.wp_chain, cb::tc_w_elems(in_f, out_f, 128) * sizeof(uint16_t));
                if (e != cudaSuccess) {
                    delete p;
                    return fail(e == cudaErrorMemoryAllocation ? CB_ERR_OOM : CB_ERR_CUDA,
                                std::string("cudaMalloc(chain weight): ") + cudaGetErrorString(e));
                }
                cb::tc_pack_weight(n.d.weight, 1, in_f, in_f, out_f, cb::tc_kp(out_f), 128, n.wp_chain, 0);
                if (li == lins[0]) continue;          // nothing upstream of the output layer needs its gradient
                e = cudaMalloc(&n.wp_chain_g, cb::tc_w_elems(out_f, in_f, 128) * sizeof(uint16_t));
                if (e != cudaSuccess) {
                    delete p;
                    return fail(e == cudaErrorMemoryAllocation ? CB_ERR_OOM : CB_ERR_CUDA,
                                std::string("cudaMalloc(chain weight): ") + cudaGetErrorString(e));
                }
                cb::tc_pack_weight(n.d.weight, in_f, 1, out_f, in_f, cb::tc_kp(in_f), 128, n.wp_chain_g, 0);
            }
            p->chain = true;
            const char* eg = getenv("CROWN_B200_DISABLE_CHAIN_GRAD");
            p->chain_grad = !(eg && eg[0] == '1');
            p->chain_lin = lins;
            p->chain_relu = relus;
        }
    }
    // conv weights transposed for the backward (transpose-conv) kernel
    for (auto& n : p->nodes) {
        if (n.d.op != CB_OP_CONV2D) continue;
        const Node& src = p->nodes[n.d.in0];
        const size_t cnt = (size_t)n.d.c * src.d.c * n.d.kh * n.d.kw;
        cudaError_t e = cudaMalloc(&n.wt, cnt * sizeof(float));
        if (e != cudaSuccess) {
            n.wt = nullptr;
            delete p;
            return fail(e == cudaErrorMemoryAllocation ? CB_ERR_OOM : CB_ERR_CUDA,
                        std::string("cudaMalloc(conv weight): ") + cudaGetErrorString(e));
        }
        k_transpose_conv_w<<<(unsigned)((cnt + 255) / 256), 256>>>(n.d.weight, n.wt, n.d.c, src.d.c,
                                                                   n.d.kh * n.d.kw);
        const int khw = n.d.kh * n.d.kw;
        e = cudaMalloc(&n.wk_b, (size_t)khw * n.d.c * cb::conv_pad(src.d.c) * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&n.wk_f, (size_t)src.d.c * khw * cb::conv_pad(n.d.c) * sizeof(float));
        if (e != cudaSuccess) {
            delete p;
            return fail(e == cudaErrorMemoryAllocation ? CB_ERR_OOM : CB_ERR_CUDA,
                        std::string("cudaMalloc(conv weight): ") + cudaGetErrorString(e));
        }
        cb::conv_relayout(n.d.weight, n.wk_b, n.d.c, src.d.c, khw, false, 0);
        cb::conv_relayout(n.d.weight, n.wk_f, n.d.c, src.d.c, khw, true, 0);
        // tensor-core form (both directions), unless switched off
        const char* ec = getenv("CROWN_B200_DISABLE_CONV_TC");
        if (p->use_tc && !(ec && ec[0] == '1')) {
            const cb::ConvGeom cg = conv_geom(p, n);
            if (cb::conv_tc_setup(cg, 0, n.ct_pass) && cb::conv_tc_setup(cg, 1, n.ct_grad)) {
                e = cudaMalloc(&n.wct_pass, cb::conv_tc_w_elems(n.ct_pass) * sizeof(uint16_t));
                if (e == cudaSuccess) e = cudaMalloc(&n.wct_grad, cb::conv_tc_w_elems(n.ct_grad) * sizeof(uint16_t));
                if (e != cudaSuccess) {
                    delete p;
                    return fail(e == cudaErrorMemoryAllocation ? CB_ERR_OOM : CB_ERR_CUDA,
                                std::string("cudaMalloc(conv tensor-core weight): ") + cudaGetErrorString(e));
                }
                cb::conv_tc_pack_weight(n.d.weight, n.ct_pass, n.d.c, src.d.c, n.wct_pass, 0);
                cb::conv_tc_pack_weight(n.d.weight, n.ct_grad, n.d.c, src.d.c, n.wct_grad, 0);
                n.ct_ok = n.ct_use_pass = n.ct_use_grad = true;
            }
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { delete p; return fail(CB_ERR_CUDA, cudaGetErrorString(e)); }
    // Per convolution and direction, keep whichever of the tensor-core and the register-tiled SIMT kernel is faster
    // on this device: the implicit GEMM wins from ~16 channels up, the SIMT kernel on the 3- and 8-channel layers of
    // the small CNNs.  CROWN_B200_CONV_AUTOTUNE=0 keeps the tensor-core kernel wherever it applies.
    {
        const char* ea = getenv("CROWN_B200_CONV_AUTOTUNE");
        // CROWN_B200_CONV_CHOICES=<string from cb_plan_conv_choices>: replay recorded decisions instead of timing -
        // under a profiler the event timings of the trial launches are meaningless (every profiled launch is
        // serialised and padded), and a profile must show the kernels the un-profiled run uses
        const char* ec = getenv("CROWN_B200_CONV_CHOICES");
        bool replayed = false;
        if (ec && ec[0]) {
            size_t n_ok = 0;
            for (auto& n : p->nodes) n_ok += n.ct_ok ? 2 : 0;
            if (strlen(ec) == n_ok) {
                size_t i = 0;
                for (auto& n : p->nodes)
                    if (n.ct_ok) {
                        n.ct_use_pass = ec[i++] == 'T';
                        n.ct_use_grad = ec[i++] == 'T';
                    }
                replayed = true;
            }
        }
        const bool tune = !(ea && ea[0] == '0') && !replayed;
        const int R = 2048;                              // rows of the trial batch: enough position tiles for every persistent CTA
        size_t need = 0;
        for (auto& n : p->nodes)
            if (n.ct_ok) {
                const size_t m = (size_t)R * (n.numel > p->nodes[n.d.in0].numel ? n.numel : p->nodes[n.d.in0].numel);
                if (m > need) need = m;
            }
        float *ta = nullptr, *tb = nullptr;
        if (tune && need && cudaMalloc(&ta, need * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&tb, need * sizeof(float)) == cudaSuccess) {
            cudaMemset(ta, 0, need * sizeof(float));
            cudaMemset(tb, 0, need * sizeof(float));
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            auto time_it = [&](auto&& fn) {
                float best = 1e30f;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaEventRecord(e0, 0);
                    fn();
                    cudaEventRecord(e1, 0);
                    cudaEventSynchronize(e1);
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (rep > 0 && ms < best) best = ms;
                }
                return best;
            };
            for (auto& n : p->nodes) {
                if (!n.ct_ok) continue;
                const cb::ConvGeom cg = conv_geom(p, n);
                const float t_tc_p = time_it([&] { cb::conv_tc(n.ct_pass, ta, tb, n.wct_pass, n.d.bias, nullptr, R, false, nullptr, 0); });
                bool ok = true;
                const float t_si_p = time_it([&] { ok = cb::conv_bwd_tiled(ta, n.wk_b, tb, cg, R, false, nullptr, 0) && ok; });
                n.ct_use_pass = !ok || t_tc_p <= t_si_p;
                const float t_tc_g = time_it([&] { cb::conv_tc(n.ct_grad, ta, tb, n.wct_grad, n.d.bias, nullptr, R, false, nullptr, 0); });
                ok = true;
                const float t_si_g = time_it([&] { ok = cb::conv_fwd_tiled(ta, n.wk_f, n.d.bias, tb, cg, R, nullptr, 0) && ok; });
                n.ct_use_grad = !ok || t_tc_g <= t_si_g;
            }
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
        if (ta) cudaFree(ta);
        if (tb) cudaFree(tb);
        cudaGetLastError();
    }
    *out_plan = p;
    return CB_OK;
}

void cb_plan_destroy(cb_plan_t* plan) { delete plan; }

int32_t cb_plan_num_activations(const cb_plan_t* plan) { return plan ? (int32_t)plan->acts.size() : 0; }

int32_t cb_plan_activation_node(const cb_plan_t* plan, int32_t k) {
    if (!plan || k < 0 || k >= (int32_t)plan->acts.size()) return -1;
    return plan->acts[k];
}

int32_t cb_plan_preact_node(const cb_plan_t* plan, int32_t k) {
    if (!plan || k < 0 || k >= (int32_t)plan->acts.size()) return -1;
    return plan->nodes[plan->acts[k]].d.in0;
}

size_t cb_workspace_bytes(const cb_plan_t* plan, int32_t Bd, int32_t S, int32_t mode,
                          const cb_problem_t* problem) {
    if (!plan || Bd <= 0 || S <= 0) return 0;
    Carver cv(nullptr, 0, true);
    Buffers bf;
    carve(plan, Bd, S, mode, problem, cv, bf);
    return cv.off + 256;
}

int cb_crown_pass(const cb_plan_t* plan, const cb_problem_t* problem, void* workspace,
                  size_t workspace_bytes, void* stream) {
    int rc = check_problem(plan, problem);
    if (rc) return rc;
    Carver cv(workspace, workspace_bytes, false);
    Buffers bf;
    carve(plan, problem->Bd, problem->S, 0, problem, cv, bf);
    if (!workspace || !cv.ok()) return fail(CB_ERR_WORKSPACE, "workspace too small");
    return run_pass(plan, problem, bf, problem->lb, problem->beta_val != nullptr, false, nullptr,
                    (cudaStream_t)stream);
}

int cb_crown_grad(const cb_plan_t* plan, const cb_problem_t* problem, float* const* h_grad_alpha,
                  float* const* h_grad_beta, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_problem(plan, problem);
    if (rc) return rc;
    Carver cv(workspace, workspace_bytes, false);
    Buffers bf;
    carve(plan, problem->Bd, problem->S, 1, problem, cv, bf);
    if (!workspace || !cv.ok()) return fail(CB_ERR_WORKSPACE, "workspace too small");
    const bool use_beta = problem->beta_val != nullptr;
    rc = run_pass(plan, problem, bf, problem->lb, use_beta, true, nullptr, (cudaStream_t)stream);
    if (rc) return rc;
    return run_grad(plan, problem, bf, h_grad_alpha, h_grad_beta, use_beta, nullptr,
                    (cudaStream_t)stream);
}

int cb_optimize(const cb_plan_t* plan, const cb_problem_t* problem, const cb_opt_t* opt,
                void* workspace, size_t workspace_bytes, void* stream, int32_t* h_n_iter) {
    int rc = check_problem(plan, problem);
    if (rc) return rc;
    if (!opt || opt->iteration <= 0) return fail(CB_ERR_ARG, "bad optimiser options");
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(workspace, workspace_bytes, false);
    Buffers bf;
    carve(plan, problem->Bd, problem->S, 2, problem, cv, bf);
    if (!workspace || !cv.ok()) return fail(CB_ERR_WORKSPACE, "workspace too small");
    const int Bd = problem->Bd, S = problem->S;
    const bool use_beta = opt->enable_beta && problem->beta_val != nullptr;

    // optimisable tensors: drop beta tables when beta is disabled
    // slopes stepped inside relu_grad: S == 1 or one slope row per spec row (every slope written exactly once)
    AdamInGrad aig;
    aig.on.assign(plan->acts.size(), 0);
    const bool fuse_adam = getenv("CROWN_B200_DISABLE_ADAM_IN_GRAD") == nullptr;
    for (size_t k = 0; k < plan->acts.size(); ++k)
        if (fuse_adam && bf.alpha_tab[k] >= 0 && (S == 1 || problem->alpha_S1 == S) &&
            relu_grad_standalone(plan, problem, use_beta, (int)k)) {
            aig.on[k] = 1;
            bf.h_tables[bf.alpha_tab[k]].fused = 1;
        }
    std::vector<cb::RowTable> tabs;
    for (auto& t : bf.h_tables)
        if (t.group != 1 || use_beta) tabs.push_back(t);
    const int nt = (int)tabs.size();
    bool vec_ok = true;
    for (auto& t : tabs)
        for (const float* q : {t.p, t.g, t.m, t.v, t.best})
            if ((reinterpret_cast<uintptr_t>(q) & 15u) != 0) vec_ok = false;
    if (nt) CB_CUDA(cudaMemcpyAsync(bf.d_tables, tabs.data(), nt * sizeof(cb::RowTable),
                                    cudaMemcpyHostToDevice, st));
    // Adam state and snapshots: g = m = v = 0, best = initial parameters (optimized_bounds.py:71-90), one launch
    cb::opt_init(bf.d_tables, nt, bf.max_rows, bf.max_cols, st);
    CB_CUDA(cudaMemsetAsync(bf.state, 0, 2 * sizeof(cb::OptState), st));
    CB_CUDA(cudaMemsetAsync(bf.snap, 0, Bd, st));

    const int iteration = opt->iteration;
    const int save_from = (int)(iteration * opt->start_save_best);
    double lr_a = opt->lr_alpha, lr_b = opt->lr_beta;
    int executed = 0;
    bool snap_in_finalize = false;

    for (int i = 0; i < iteration; ++i) {
        cb::OptState* st_cur = bf.state + (i & 1);
        cb::OptState* st_next = bf.state + ((i + 1) & 1);
        const int* done = &st_cur->done;
        KeepBest kb;
        kb.iter = i; kb.rhs = opt->rhs; kb.state = st_cur;
        rc = run_pass(plan, problem, bf, bf.lb_cur, use_beta, true, done, st, &kb);
        if (rc) return rc;
        if (!kb.fused)
            cb::keepbest_a(i, bf.lb_cur, opt->rhs, bf.best_l, bf.best_ret, bf.ret0, bf.stopped,
                           bf.mask0, st_cur, Bd, S, st);
        // the second half runs in the head of the whole-network gradient kernel whenever one follows unconditionally
        const bool fuse_b = !opt->early_stop && i != iteration - 1 && chain_grad_applies(plan, problem, use_beta);
        if (!fuse_b)
            cb::keepbest_b(i, iteration, save_from, opt->early_stop_patience, bf.lb_cur, bf.ret0,
                           bf.mask0, bf.snap, st_cur, st_next, Bd, S, st);
        // the snapshot is fused into the Adam step below whenever a step follows unconditionally
        const bool fuse_snap = !opt->early_stop && i != iteration - 1;
        // ... and into the final copy-back after the last iteration of a loop that runs to the end on the device
        snap_in_finalize = !opt->early_stop && i == iteration - 1;
        if (!fuse_snap && !snap_in_finalize) cb::snapshot(bf.d_tables, nt, bf.max_rows, bf.max_cols, bf.snap, Bd, st);
        executed = i + 1;
        if (opt->early_stop) {
            cb::OptState h;
            CB_CUDA(cudaMemcpyAsync(&h, st_next, sizeof(h), cudaMemcpyDeviceToHost, st));
            CB_CUDA(cudaStreamSynchronize(st));
            executed = h.n_iter;
            if (h.done) break;
        }
        if (i != iteration - 1) {
            const int* done_next = &st_next->done;
            const double bc1 = 1.0 - pow(0.9, i + 1);
            const double bc2 = 1.0 - pow(0.999, i + 1);
            KeepBestB kbb;
            if (fuse_b) {
                kbb.iter = i; kbb.iteration = iteration; kbb.save_from = save_from;
                kbb.patience_limit = opt->early_stop_patience;
                kbb.cur = st_cur; kbb.next = st_next;
            }
            aig.stopped = bf.stopped;
            aig.snap = fuse_snap ? bf.snap : nullptr;
            aig.step = (float)lr_a / (float)bc1;
            aig.bc2_sqrt = (float)sqrt(bc2);
            rc = run_grad(plan, problem, bf, bf.grad_alpha.data(), bf.grad_beta.data(), use_beta,
                          done_next, st, fuse_b ? &kbb : nullptr, &aig);
            if (rc) return rc;
            if (fuse_b && !kbb.fused) return fail(CB_ERR_ARG, "keep-best bookkeeping was not run");
            // (the step as a tail of the gradient kernel - every CTA updating its own rows while g and p are still in
            // the L2 - was measured: 187 us against 100 + 59 us for the two launches; 512 threads per SM do not cover it)
            cb::adam_step(bf.d_tables, nt, bf.max_rows, bf.max_cols, bf.stopped, fuse_snap ? bf.snap : nullptr, Bd,
                          (float)lr_a, (float)lr_b, (float)bc1, (float)sqrt(bc2), vec_ok, done_next, st);
            lr_a *= opt->lr_decay;
            lr_b *= opt->lr_decay;
        }
    }
    cb::finalize(bf.d_tables, nt, bf.max_rows, bf.max_cols, bf.best_ret, problem->lb, Bd * S,
                 snap_in_finalize ? bf.snap : nullptr, Bd, st);
    CB_CUDA(cudaGetLastError());
    if (h_n_iter) *h_n_iter = executed;
    return CB_OK;
}

int cb_debug_tc_gemm(const float* X, const float* W, const float* col_bias, float* Y, int32_t rows,
                     int32_t N, int32_t K, int32_t bn, int32_t dbg, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!X || !W || !Y || rows <= 0 || N <= 0 || K <= 0) return fail(CB_ERR_ARG, "bad arguments");
    const int BN = bn > 0 ? bn : cb::tc_pick_bn(N, 128);
    if (BN % 32 != 0 || BN < 32 || BN > 128) return fail(CB_ERR_ARG, "BN must be a multiple of 32 in [32,128]");
    const int Kp = cb::tc_kp(K);
    uint16_t *xp = nullptr, *wp = nullptr;
    CB_CUDA(cudaMalloc(&xp, cb::tc_x_elems(rows, K) * sizeof(uint16_t)));
    CB_CUDA(cudaMalloc(&wp, cb::tc_w_elems(N, K, BN) * sizeof(uint16_t)));
    cb::tc_pack_rows(X, false, rows, rows, 1, K, Kp, xp, nullptr, nullptr, nullptr, st);
    cb::tc_pack_weight(W, K, 1, N, K, Kp, BN, wp, st);
    cb::TcArgs a;
    memset(&a, 0, sizeof(a));
    a.xp = xp; a.wp = wp;
    a.rows = rows; a.Bd = rows; a.S = 1; a.S1 = 1;
    a.N = N; a.Kp = Kp; a.BN = BN;
    a.y_plain = Y;
    a.col_bias = col_bias;
    a.dbg = dbg;
    cudaError_t e = cb::tc_linear(cb::TC_MODE_STORE, a, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(xp); cudaFree(wp);
    if (e != cudaSuccess) return fail(CB_ERR_CUDA, std::string("tc gemm: ") + cudaGetErrorString(e));
    return CB_OK;
}

void cb_debug_tc_times(void* device_buffer) { cb::tc_debug_set_times(static_cast<long long*>(device_buffer)); }

int32_t cb_plan_conv_choices(const cb_plan_t* plan, char* out, int32_t cap) {
    if (!plan || !out || cap <= 0) return -1;
    int32_t n = 0;
    for (const auto& nd : plan->nodes)
        if (nd.ct_ok) {
            if (n + 2 >= cap) return -1;
            out[n++] = nd.ct_use_pass ? 'T' : 'S';
            out[n++] = nd.ct_use_grad ? 'T' : 'S';
        }
    out[n] = 0;
    return n;
}

int32_t cb_plan_uses_conv_tc(const cb_plan_t* plan) {
    if (!plan) return 0;
    int n = 0;
    for (const auto& nd : plan->nodes) n += (nd.ct_use_pass ? 1 : 0) + (nd.ct_use_grad ? 1 : 0);
    return n;
}

int cb_debug_conv_tc(const float* X, const float* W, const float* bias, float* Y, float* bias_rows, int32_t rows,
                     int32_t Cin, int32_t Hin, int32_t Win, int32_t Cout, int32_t KH, int32_t KW, int32_t stride,
                     int32_t pad, int32_t dir, int32_t accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!X || !W || !Y || rows <= 0) return fail(CB_ERR_ARG, "bad arguments");
    cb::ConvGeom g;
    g.Cin = Cin; g.Hin = Hin; g.Win = Win; g.Cout = Cout; g.KH = KH; g.KW = KW;
    g.sh = g.sw = stride; g.ph = g.pw = pad; g.dh = g.dw = 1;
    g.Hout = (Hin + 2 * pad - KH) / stride + 1;
    g.Wout = (Win + 2 * pad - KW) / stride + 1;
    cb::ConvTcGeom ct;
    if (!cb::conv_tc_setup(g, dir, ct)) return fail(CB_ERR_ARG, "geometry not supported by the tensor-core convolution");
    uint16_t* wp = nullptr;
    CB_CUDA(cudaMalloc(&wp, cb::conv_tc_w_elems(ct) * sizeof(uint16_t)));
    cb::conv_tc_pack_weight(W, ct, Cout, Cin, wp, st);
    cudaError_t e = cb::conv_tc(ct, X, Y, wp, bias, bias_rows, rows, accumulate != 0, nullptr, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(wp);
    if (e != cudaSuccess) return fail(CB_ERR_CUDA, std::string("conv tc: ") + cudaGetErrorString(e));
    return CB_OK;
}

int cb_debug_conv_simt(const float* X, const float* W, const float* bias, float* Y, int32_t rows, int32_t Cin,
                       int32_t Hin, int32_t Win, int32_t Cout, int32_t KH, int32_t KW, int32_t stride, int32_t pad,
                       int32_t dir, int32_t accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!X || !W || !Y || rows <= 0) return fail(CB_ERR_ARG, "bad arguments");
    cb::ConvGeom g;
    g.Cin = Cin; g.Hin = Hin; g.Win = Win; g.Cout = Cout; g.KH = KH; g.KW = KW;
    g.sh = g.sw = stride; g.ph = g.pw = pad; g.dh = g.dw = 1;
    g.Hout = (Hin + 2 * pad - KH) / stride + 1;
    g.Wout = (Win + 2 * pad - KW) / stride + 1;
    const int P = cb::conv_pad(dir == 0 ? Cin : Cout);
    const size_t n = dir == 0 ? (size_t)KH * KW * Cout * P : (size_t)Cin * KH * KW * P;
    float* wk = nullptr;
    CB_CUDA(cudaMalloc(&wk, n * sizeof(float)));
    cb::conv_relayout(W, wk, Cout, Cin, KH * KW, dir != 0, st);
    bool ok = dir == 0 ? cb::conv_bwd_tiled(X, wk, Y, g, rows, accumulate != 0, nullptr, st)
                       : cb::conv_fwd_tiled(X, wk, bias, Y, g, rows, nullptr, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(wk);
    if (!ok) return fail(CB_ERR_ARG, "geometry not supported by the register-tiled convolution");
    if (e != cudaSuccess) return fail(CB_ERR_CUDA, std::string("conv simt: ") + cudaGetErrorString(e));
    return CB_OK;
}

int32_t cb_plan_uses_chain(const cb_plan_t* plan) { return (plan && plan->chain) ? (plan->chain_grad ? 2 : 1) : 0; }

int32_t cb_plan_uses_tensor_cores(const cb_plan_t* plan) {
    if (!plan) return 0;
    int n = 0;
    for (const auto& nd : plan->nodes) n += (nd.tc_pass ? 1 : 0) + (nd.tc_grad ? 1 : 0);
    return n;
}

}  // extern "C"
