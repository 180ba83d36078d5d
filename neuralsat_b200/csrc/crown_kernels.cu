// CROWN / alpha-beta-CROWN kernels for sm_100a — generic (any-topology) path.
//
// Everything here is fp32 SIMT; the dense contractions of wide FC layers are taken over by the
// tcgen05 kernels in crown_tc.cu when the plan enables them.  Layouts follow the reference:
// coefficient matrices A are [S,Bd,n] (row r = s*Bd + b), per-domain data (l,u,x,alpha,beta) are
// [Bd,...].  Every kernel takes `done`: a device flag set by the optimisation loop once the
// reference would have left its loop (auto_LiRPA/optimized_bounds.py:522-530); kernels then
// return immediately, which reproduces the data-dependent early exit without a host sync.
#include "crown_kernels.cuh"

#include <mutex>
#include <vector>

namespace cb {

// ---------------------------------------------------------------------------------------------
// built-in profiler: launch counter + optional per-launch CUDA events on the launching stream
// ---------------------------------------------------------------------------------------------
namespace {
struct EvRec { int id; cudaEvent_t a, b; };
std::mutex g_prof_mu;
bool g_prof_on = false;
long long g_launches = 0;
std::vector<EvRec*> g_recs;
}  // namespace

const char* kernel_name(int id) {
    static const char* names[K_COUNT] = {
        "sgemm_nn", "sgemm_nt", "relu_bwd", "relu_grad", "beta_scatter", "beta_grad", "concretize",
        "grad_init", "conv_bwd", "conv_fwd", "chan", "elementwise", "keepbest", "snapshot", "adam",
        "tc_linear", "tc_pack", "chain_pass", "chain_grad", "sshape", "conv_tc_bwd", "conv_tc_fwd", "store", "branch"};
    return (id >= 0 && id < K_COUNT) ? names[id] : "?";
}

Launch::Launch(int id_, cudaStream_t st_) : id(id_), st(st_), rec(nullptr) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ++g_launches;
    if (g_prof_on) {
        EvRec* r = new EvRec();
        r->id = id;
        cudaEventCreate(&r->a);
        cudaEventCreate(&r->b);
        cudaEventRecord(r->a, st);
        g_recs.push_back(r);
        rec = r;
    }
}

Launch::~Launch() {
    if (rec) cudaEventRecord(static_cast<EvRec*>(rec)->b, st);
}

void profile_enable(bool on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on;
}

long long launch_count() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    return g_launches;
}

int profile_collect(double* ms, long long* launches, int n) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < n; ++i) { ms[i] = 0.0; launches[i] = 0; }
    for (EvRec* r : g_recs) {
        cudaEventSynchronize(r->b);
        float t = 0.f;
        cudaEventElapsedTime(&t, r->a, r->b);
        if (r->id < n) { ms[r->id] += t; launches[r->id] += 1; }
        cudaEventDestroy(r->a);
        cudaEventDestroy(r->b);
        delete r;
    }
    g_recs.clear();
    return K_COUNT;
}

#define CB_DONE_CHECK(done) do { if ((done) != nullptr && *(done) != 0) return; } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the TPR threads that share one row. TPR in {32, 256}; 256 => whole block (8 warps).
template <int TPR>
__device__ __forceinline__ float row_sum(float v, float* red /* [8] shared, TPR==256 only */) {
    v = warp_sum(v);
    if (TPR == 32) return v;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

// ---------------------------------------------------------------------------------------------
// ReLU relaxation (auto_LiRPA/operators/relu.py:456-494)
// ---------------------------------------------------------------------------------------------
struct Relax {
    float d_u, b_u, d_l;
    bool alpha_live;   // gradient reaches alpha: unstable neuron and alpha inside the clamp
};

__device__ __forceinline__ Relax relu_relax(float l, float u, bool has_alpha, float a) {
    Relax r;
    const float lb_r = fminf(l, 0.f);
    float ub_r = fmaxf(u, 0.f);
    ub_r = fmaxf(ub_r, lb_r + 1e-8f);
    r.d_u = slope_div(ub_r, ub_r - lb_r);
    r.b_u = -lb_r * r.d_u;
    if (has_alpha) {
        const float lower_mask = (l >= 0.f) ? 1.f : 0.f;
        const float upper_mask = (u <= 0.f) ? 1.f : 0.f;
        const float no_mask = (1.f - lower_mask) * (1.f - upper_mask);
        r.d_l = fminf(fmaxf(a, 0.f), 1.f) * no_mask + lower_mask;
        r.alpha_live = (no_mask != 0.f) && (a >= 0.f) && (a <= 1.f);
    } else {
        r.d_l = (r.d_u > 0.5f) ? 1.f : 0.f;
        r.alpha_live = false;
    }
    return r;
}

__global__ void k_spec_to_rows(const float* __restrict__ C, float* __restrict__ A, int Bd, int S,
                               int n, const int* done) {
    CB_DONE_CHECK(done);
    const size_t total = (size_t)Bd * S * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % n);
        const size_t r = i / n;              // r = s*Bd + b
        const int b = (int)(r % Bd), s = (int)(r / Bd);
        A[i] = C[((size_t)b * S + s) * n + k];
    }
}

void spec_to_rows(const float* C, float* A, int Bd, int S, int n, const int* done, cudaStream_t st) {
    Launch _l(K_ELEMWISE, st);
    const size_t total = (size_t)Bd * S * n;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    k_spec_to_rows<<<blocks, 256, 0, st>>>(C, A, Bd, S, n, done);
}

// ---------------------------------------------------------------------------------------------
// SGEMM  C[M,N] (+)= A[M,K] * op(B)
// ---------------------------------------------------------------------------------------------
template <bool TB>
__global__ void __launch_bounds__(256)
k_sgemm(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M,
        int N, int K, int accumulate, const float* __restrict__ rvec, float* __restrict__ rout,
        const float* __restrict__ cbias, const int* done) {
    CB_DONE_CHECK(done);
    constexpr int BM = 128, BN = 64, BK = 16;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float rdot = 0.f;
    const bool do_rdot = (rvec != nullptr) && (blockIdx.x == 0) && (tid < BM);

    float ra[8], rb[4];
    auto gload = [&](int k0) {
        const int ka = k0 + (tid & 15);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + (tid >> 4) + 16 * i;
            ra[i] = (m < M && ka < K) ? __ldg(A + (size_t)m * K + ka) : 0.f;
        }
        if (!TB) {
            const int n = n0 + (tid & 63);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + (tid >> 6) + 4 * i;
                rb[i] = (n < N && k < K) ? __ldg(B + (size_t)k * N + n) : 0.f;
            }
        } else {
            const int k = k0 + (tid & 15);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = n0 + (tid >> 4) + 16 * i;
                rb[i] = (n < N && k < K) ? __ldg(B + (size_t)n * K + k) : 0.f;
            }
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[tid & 15][(tid >> 4) + 16 * i] = ra[i];
        if (!TB) {
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[(tid >> 6) + 4 * i][tid & 63] = rb[i];
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[tid & 15][(tid >> 4) + 16 * i] = rb[i];
        }
    };

    gload(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
        sstore();
        __syncthreads();
        if (k0 + BK < K) gload(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (do_rdot) {
#pragma unroll
            for (int kk = 0; kk < BK; ++kk)
                if (k0 + kk < K) rdot = fmaf(As[kk][tid], __ldg(rvec + k0 + kk), rdot);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (cbias) v += __ldg(cbias + n);
            float* p = C + (size_t)m * N + n;
            *p = accumulate ? (*p + v) : v;
        }
    }
    if (do_rdot && m0 + tid < M) rout[m0 + tid] += rdot;
}

void sgemm(bool trans_b, const float* A, const float* B, float* C, int M, int N, int K,
           bool accumulate, const float* rowdot_vec, float* rowdot_out, const float* col_bias,
           const int* done, cudaStream_t st) {
    Launch _l(trans_b ? K_SGEMM_NT : K_SGEMM_NN, st);
    dim3 grid((N + 63) / 64, (M + 127) / 128);
    if (trans_b)
        k_sgemm<true><<<grid, 256, 0, st>>>(A, B, C, M, N, K, accumulate ? 1 : 0, rowdot_vec,
                                            rowdot_out, col_bias, done);
    else
        k_sgemm<false><<<grid, 256, 0, st>>>(A, B, C, M, N, K, accumulate ? 1 : 0, rowdot_vec,
                                             rowdot_out, col_bias, done);
}

// ---------------------------------------------------------------------------------------------
// ReLU backward relaxation + sign-split multiply (operators/clampmult.py:17-43)
// ---------------------------------------------------------------------------------------------
template <int TPR>
__global__ void __launch_bounds__(256)
k_relu_bwd(const float* __restrict__ A_post, float* __restrict__ A_pre, int accumulate,
           float* __restrict__ bias_rows, ReluArgs ra, BetaScatter bs, int Bd, int S, int n, const int* done) {
    CB_DONE_CHECK(done);
    __shared__ float red[8];
    const int b = blockIdx.x * (256 / TPR) + threadIdx.x / TPR;
    const int lane = threadIdx.x % TPR;
    const bool active = b < Bd;
    const bool has_alpha = ra.alpha != nullptr;
    const bool vec = (n & 3) == 0 && ra.alpha_pos == nullptr && (!has_alpha || ra.n_alpha == n) &&
                     ((reinterpret_cast<uintptr_t>(A_post) | reinterpret_cast<uintptr_t>(A_pre) |
                       reinterpret_cast<uintptr_t>(ra.lower) | reinterpret_cast<uintptr_t>(ra.upper) |
                       reinterpret_cast<uintptr_t>(ra.alpha)) & 15u) == 0;
    for (int s = 0; s < S; ++s) {
        float part = 0.f;
        if (active) {
            const size_t r = (size_t)s * Bd + b;
            const float* ap = A_post + r * n;
            float* op = A_pre + r * n;
            const float* lp = ra.lower + (size_t)b * n;
            const float* up = ra.upper + (size_t)b * n;
            const float* al = has_alpha
                ? ra.alpha + ((size_t)(ra.S1 == 1 ? 0 : s) * Bd + b) * ra.n_alpha : nullptr;
            if (vec) {
                // dense slopes, n % 4 == 0: 16-byte accesses (four neurons per thread and trip)
                const float4* ap4 = reinterpret_cast<const float4*>(ap);
                float4* op4 = reinterpret_cast<float4*>(op);
                const float4* lp4 = reinterpret_cast<const float4*>(lp);
                const float4* up4 = reinterpret_cast<const float4*>(up);
                const float4* al4 = reinterpret_cast<const float4*>(al);
                for (int i = lane; i < (n >> 2); i += TPR) {
                    const float4 l4 = __ldg(lp4 + i), u4 = __ldg(up4 + i), a4 = ap4[i];
                    const float4 v4 = has_alpha ? __ldg(al4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 o = accumulate ? op4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, uv[4] = {u4.x, u4.y, u4.z, u4.w};
                    const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
                    float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const Relax rx = relu_relax(lv[e], uv[e], has_alpha, vv[e]);
                        const float a_pos = fmaxf(aa[e], 0.f), a_neg = fminf(aa[e], 0.f);
                        const float v = rx.d_l * a_pos + rx.d_u * a_neg;
                        ov[e] = accumulate ? (ov[e] + v) : v;
                        part = fmaf(a_neg, rx.b_u, part);
                    }
                    op4[i] = make_float4(ov[0], ov[1], ov[2], ov[3]);
                }
            } else
            for (int i = lane; i < n; i += TPR) {
                float av = 0.f;
                if (has_alpha) {
                    const int pos = ra.alpha_pos ? __ldg(ra.alpha_pos + i) : i;
                    av = pos >= 0 ? __ldg(al + pos) : 0.f;
                }
                const Relax rx = relu_relax(__ldg(lp + i), __ldg(up + i), has_alpha, av);
                const float a = ap[i];
                const float a_pos = fmaxf(a, 0.f), a_neg = fminf(a, 0.f);
                const float v = rx.d_l * a_pos + rx.d_u * a_neg;
                op[i] = accumulate ? (op[i] + v) : v;
                part = fmaf(a_neg, rx.b_u, part);
            }
        }
        // split constraints of the layer: their bias term joins the row sum, their coefficients are added to the row
        // once every thread of the row has stored its part (row_sum synchronises the row's threads)
        if (active && bs.J > 0 && bs.bias != nullptr)
            for (int j = lane; j < bs.J; j += TPR) {
                const size_t q = (size_t)b * bs.J + j;
                part = fmaf(bs.val[q] * bs.sign[q], bs.bias[q], part);
            }
        const float tot = row_sum<TPR>(part, red);
        if (TPR == 32) __syncwarp();
        if (active && bs.J > 0) {
            float* const op = A_pre + ((size_t)s * Bd + b) * n;
            for (int j = lane; j < bs.J; j += TPR) {
                const size_t q = (size_t)b * bs.J + j;
                const float vs = bs.val[q] * bs.sign[q];
                if (vs != 0.f) atomicAdd(op + bs.loc[q], -vs);
            }
        }
        if (active && lane == 0) bias_rows[(size_t)s * Bd + b] += tot;
    }
}

void relu_bwd(const float* A_post, float* A_pre, bool accumulate, float* bias_rows,
              const ReluArgs& ra, int Bd, int S, int n, const int* done, cudaStream_t st, const BetaScatter* beta) {
    Launch _l(K_RELU_BWD, st);
    const BetaScatter bs = beta ? *beta : BetaScatter();
    if (n <= 1024) {
        k_relu_bwd<32><<<(Bd + 7) / 8, 256, 0, st>>>(A_post, A_pre, accumulate, bias_rows, ra, bs, Bd,
                                                     S, n, done);
    } else {
        k_relu_bwd<256><<<Bd, 256, 0, st>>>(A_post, A_pre, accumulate, bias_rows, ra, bs, Bd, S, n, done);
    }
}

template <int TPR>
__global__ void __launch_bounds__(256)
k_relu_grad(const float* __restrict__ A_post, const float* __restrict__ g_pre,
            float* __restrict__ g_post, float* __restrict__ grad_alpha, ReluArgs ra, AdamFuse af, BetaGrad bg, int Bd,
            int S, int n, const int* done) {
    const bool dn = done != nullptr && *done != 0;
    const bool fuse = af.p != nullptr;
    // `done` (the reference left its loop in this iteration) cancels everything but the keep-best snapshot, which the
    // reference takes before breaking (optimized_bounds.py:483-530; same rule as k_adam)
    if (dn && !(fuse && af.snap)) return;
    const int b = blockIdx.x * (256 / TPR) + threadIdx.x / TPR;
    const int lane = threadIdx.x % TPR;
    if (b >= Bd) return;
    const bool has_alpha = ra.alpha != nullptr;
    const float* lp = ra.lower + (size_t)b * n;
    const float* up = ra.upper + (size_t)b * n;
    const bool vec = (n & 3) == 0 && ra.alpha_pos == nullptr && (!has_alpha || ra.n_alpha == n) &&
                     ((reinterpret_cast<uintptr_t>(A_post) | reinterpret_cast<uintptr_t>(g_pre) |
                       reinterpret_cast<uintptr_t>(g_post) | reinterpret_cast<uintptr_t>(grad_alpha) |
                       reinterpret_cast<uintptr_t>(ra.lower) | reinterpret_cast<uintptr_t>(ra.upper) |
                       reinterpret_cast<uintptr_t>(ra.alpha) | reinterpret_cast<uintptr_t>(af.m) |
                       reinterpret_cast<uintptr_t>(af.v) | reinterpret_cast<uintptr_t>(af.best)) & 15u) == 0;
    const bool snap_b = fuse && af.snap != nullptr && af.snap[b] != 0;
    const bool stop_b = fuse && af.stopped[b] != 0;
    if (dn) {
        // snapshot only: best <- p for a flagged sub-domain
        if (!snap_b) return;
        for (int s = 0; s < ra.S1; ++s) {
            const size_t arow = ((size_t)s * Bd + b) * ra.n_alpha;
            for (int i = lane; i < ra.n_alpha; i += TPR) af.best[arow + i] = af.p[arow + i];
        }
        return;
    }
    // gradient of the split multipliers of this layer (same arithmetic as k_beta_grad)
    for (int j = lane; j < bg.J; j += TPR) {
        const size_t q = (size_t)b * bg.J + j;
        const float sg = bg.sign[q];
        const int64_t lc = bg.loc[q];
        float acc = 0.f;
        for (int s = 0; s < S; ++s) {
            acc -= sg * g_pre[((size_t)s * Bd + b) * n + lc];
            if (bg.bias) acc = fmaf(sg, bg.bias[q], acc);
        }
        bg.grad_val[q] = acc;
    }
    for (int s = 0; s < S; ++s) {
        const size_t r = (size_t)s * Bd + b;
        const size_t arow = ((size_t)(ra.S1 == 1 ? 0 : s) * Bd + b) * ra.n_alpha;
        const float* al = has_alpha ? ra.alpha + arow : nullptr;
        float* ga = (grad_alpha && has_alpha && !fuse) ? grad_alpha + arow : nullptr;
        if (vec) {
            const float4* ap4 = reinterpret_cast<const float4*>(A_post + r * n);
            const float4* gp4 = reinterpret_cast<const float4*>(g_pre + r * n);
            float4* go4 = g_post ? reinterpret_cast<float4*>(g_post + r * n) : nullptr;
            const float4* lp4 = reinterpret_cast<const float4*>(lp);
            const float4* up4 = reinterpret_cast<const float4*>(up);
            const float4* al4 = reinterpret_cast<const float4*>(al);
            float4* ga4 = reinterpret_cast<float4*>(ga);
            const bool add = ra.S1 == 1 && s > 0;
            for (int i = lane; i < (n >> 2); i += TPR) {
                const float4 l4 = __ldg(lp4 + i), u4 = __ldg(up4 + i), a4 = ap4[i], g4 = gp4[i];
                // fused: the slopes are rewritten below, so no read-only path for them
                const float4 v4 = !has_alpha ? make_float4(0.f, 0.f, 0.f, 0.f) : (fuse ? al4[i] : __ldg(al4 + i));
                const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, uv[4] = {u4.x, u4.y, u4.z, u4.w};
                const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
                const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
                float go[4], gc[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const Relax rx = relu_relax(lv[e], uv[e], has_alpha, vv[e]);
                    go[e] = gg[e] * (aa[e] >= 0.f ? rx.d_l : rx.d_u) + (aa[e] < 0.f ? rx.b_u : 0.f);
                    gc[e] = (rx.alpha_live && aa[e] >= 0.f) ? gg[e] * aa[e] : 0.f;
                }
                if (go4) go4[i] = make_float4(go[0], go[1], go[2], go[3]);
                if (fuse) {
                    const size_t q = (arow >> 2) + i;
                    if (snap_b) reinterpret_cast<float4*>(af.best)[q] = v4;
                    float4 m4 = reinterpret_cast<float4*>(af.m)[q], w4 = reinterpret_cast<float4*>(af.v)[q];
                    float4 p4;
                    p4.x = adam_one(vv[0], gc[0], m4.x, w4.x, stop_b, af.step, af.bc2_sqrt, 0);
                    p4.y = adam_one(vv[1], gc[1], m4.y, w4.y, stop_b, af.step, af.bc2_sqrt, 0);
                    p4.z = adam_one(vv[2], gc[2], m4.z, w4.z, stop_b, af.step, af.bc2_sqrt, 0);
                    p4.w = adam_one(vv[3], gc[3], m4.w, w4.w, stop_b, af.step, af.bc2_sqrt, 0);
                    reinterpret_cast<float4*>(af.m)[q] = m4;
                    reinterpret_cast<float4*>(af.v)[q] = w4;
                    reinterpret_cast<float4*>(af.p)[q] = p4;
                } else if (ga4) {
                    if (add) { const float4 o = ga4[i]; gc[0] += o.x; gc[1] += o.y; gc[2] += o.z; gc[3] += o.w; }
                    ga4[i] = make_float4(gc[0], gc[1], gc[2], gc[3]);
                }
            }
            continue;
        }
        for (int i = lane; i < n; i += TPR) {
            int pos = i;
            float av = 0.f;
            if (has_alpha) {
                pos = ra.alpha_pos ? __ldg(ra.alpha_pos + i) : i;
                av = pos >= 0 ? (fuse ? al[pos] : __ldg(al + pos)) : 0.f;
            }
            const Relax rx = relu_relax(__ldg(lp + i), __ldg(up + i), has_alpha, av);
            const float a = A_post[r * n + i];
            const float gp = g_pre[r * n + i];
            if (g_post) g_post[r * n + i] = gp * (a >= 0.f ? rx.d_l : rx.d_u) + (a < 0.f ? rx.b_u : 0.f);
            const float c = (rx.alpha_live && a >= 0.f) ? gp * a : 0.f;
            if (fuse) {
                if (pos >= 0) {
                    const size_t q = arow + pos;
                    if (snap_b) af.best[q] = av;
                    float m = af.m[q], v = af.v[q];
                    af.p[q] = adam_one(av, c, m, v, stop_b, af.step, af.bc2_sqrt, 0);
                    af.m[q] = m;
                    af.v[q] = v;
                }
            } else if (ga && pos >= 0) {
                if (ra.S1 == 1 && s > 0) ga[pos] += c; else ga[pos] = c;
            }
        }
    }
}

void relu_grad(const float* A_post, const float* g_pre, float* g_post, float* grad_alpha,
               const ReluArgs& ra, int Bd, int S, int n, const int* done, cudaStream_t st, const AdamFuse* adam,
               const BetaGrad* beta) {
    Launch _l(K_RELU_GRAD, st);
    AdamFuse af;
    if (adam != nullptr && ra.alpha != nullptr && (S == 1 || ra.S1 == S)) af = *adam;
    const BetaGrad bg = beta ? *beta : BetaGrad();
    if (n <= 1024)
        k_relu_grad<32><<<(Bd + 7) / 8, 256, 0, st>>>(A_post, g_pre, g_post, grad_alpha, ra, af, bg, Bd, S,
                                                      n, done);
    else
        k_relu_grad<256><<<Bd, 256, 0, st>>>(A_post, g_pre, g_post, grad_alpha, ra, af, bg, Bd, S, n, done);
}

// ---------------------------------------------------------------------------------------------
// beta injection (auto_LiRPA/beta_crown.py:163-204) and its gradient
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_beta_scatter(float* __restrict__ A, float* __restrict__ bias_rows, const float* __restrict__ val,
               const int64_t* __restrict__ loc, const float* __restrict__ sign,
               const float* __restrict__ bbias, int J, int Bd, int S, int n, const int* done) {
    CB_DONE_CHECK(done);
    const size_t r = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= (size_t)Bd * S) return;
    const int b = (int)(r % Bd);
    float part = 0.f;
    for (int j = lane; j < J; j += 32) {
        const float vs = val[(size_t)b * J + j] * sign[(size_t)b * J + j];
        if (vs != 0.f) atomicAdd(A + r * n + loc[(size_t)b * J + j], -vs);
        if (bbias) part = fmaf(vs, bbias[(size_t)b * J + j], part);
    }
    if (bbias) {
        part = warp_sum(part);
        if (lane == 0) bias_rows[r] += part;
    }
}

void beta_scatter(float* A, float* bias_rows, const float* val, const int64_t* loc,
                  const float* sign, const float* bbias, int J, int Bd, int S, int n,
                  const int* done, cudaStream_t st) {
    Launch _l(K_BETA_SCATTER, st);
    const size_t rows = (size_t)Bd * S;
    k_beta_scatter<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(A, bias_rows, val, loc, sign, bbias,
                                                              J, Bd, S, n, done);
}

__global__ void k_beta_grad(const float* __restrict__ g, float* __restrict__ grad_val,
                            const int64_t* __restrict__ loc, const float* __restrict__ sign,
                            const float* __restrict__ bbias, int J, int Bd, int S, int n,
                            const int* done) {
    CB_DONE_CHECK(done);
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)Bd * J) return;
    const int b = (int)(idx / J);
    const float sg = sign[idx];
    const int64_t lc = loc[idx];
    float acc = 0.f;
    for (int s = 0; s < S; ++s) {
        acc -= sg * g[((size_t)s * Bd + b) * n + lc];
        if (bbias) acc = fmaf(sg, bbias[idx], acc);
    }
    grad_val[idx] = acc;
}

void beta_grad(const float* g, float* grad_val, const int64_t* loc, const float* sign,
               const float* bbias, int J, int Bd, int S, int n, const int* done, cudaStream_t st) {
    Launch _l(K_BETA_GRAD, st);
    const size_t total = (size_t)Bd * J;
    if (total == 0) return;
    k_beta_grad<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, grad_val, loc, sign, bbias, J,
                                                                Bd, S, n, done);
}

// ---------------------------------------------------------------------------------------------
// concretisation (auto_LiRPA/perturbations.py:154-183, sign=-1) and the gradient seed
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float grad_seed(float a, float lo, float hi) {
    const float c = (hi + lo) / 2.0f, d = (hi - lo) / 2.0f;
    const float sg = (a > 0.f) ? 1.f : ((a < 0.f) ? -1.f : 0.f);
    return c - sg * d;
}

template <int TPR>
__global__ void __launch_bounds__(256)
k_concretize(const float* __restrict__ A0, const float* __restrict__ x_L,
             const float* __restrict__ x_U, const float* __restrict__ bias_rows,
             float* __restrict__ lb, float* __restrict__ g0, int Bd, int S, int n, const int* done) {
    CB_DONE_CHECK(done);
    __shared__ float red[8];
    const size_t rows = (size_t)Bd * S;
    const size_t r = (size_t)blockIdx.x * (256 / TPR) + threadIdx.x / TPR;
    const int lane = threadIdx.x % TPR;
    const bool active = r < rows;
    float part = 0.f;
    int b = 0, s = 0;
    if (active) {
        b = (int)(r % Bd);
        s = (int)(r / Bd);
        const float* a = A0 + r * n;
        const float* xl = x_L + (size_t)b * n;
        const float* xu = x_U + (size_t)b * n;
        for (int i = lane; i < n; i += TPR) {
            const float lo = __ldg(xl + i), hi = __ldg(xu + i);
            const float c = (hi + lo) / 2.0f, d = (hi - lo) / 2.0f;
            const float av = a[i];
            part += av * c - fabsf(av) * d;
            if (g0) g0[r * n + i] = grad_seed(av, lo, hi);
        }
    }
    const float tot = row_sum<TPR>(part, red);
    if (active && lane == 0) lb[(size_t)b * S + s] = bias_rows[r] + tot;
}

void concretize(const float* A0, const float* x_L, const float* x_U, const float* bias_rows,
                float* lb, int Bd, int S, int n_in, const int* done, cudaStream_t st, float* g0) {
    Launch _l(K_CONCRETIZE, st);
    const size_t rows = (size_t)Bd * S;
    if (n_in <= 2048)
        k_concretize<32><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(A0, x_L, x_U, bias_rows, lb, g0,
                                                                    Bd, S, n_in, done);
    else
        k_concretize<256><<<(unsigned)rows, 256, 0, st>>>(A0, x_L, x_U, bias_rows, lb, g0, Bd, S, n_in,
                                                         done);
}

template <bool VEC>
__global__ void k_grad_init(const float* __restrict__ A0, const float* __restrict__ x_L,
                            const float* __restrict__ x_U, float* __restrict__ g0, int Bd, int S,
                            int n, const int* done) {
    CB_DONE_CHECK(done);
    if (VEC) {
        // n % 4 == 0, 16-byte aligned rows: four inputs per thread and trip
        const int n4 = n >> 2;
        const size_t total = (size_t)Bd * S * n4;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
             i += (size_t)gridDim.x * blockDim.x) {
            const int k = (int)(i % n4);
            const int b = (int)((i / n4) % Bd);
            const float4 lo = __ldg(reinterpret_cast<const float4*>(x_L) + (size_t)b * n4 + k);
            const float4 hi = __ldg(reinterpret_cast<const float4*>(x_U) + (size_t)b * n4 + k);
            const float4 a = reinterpret_cast<const float4*>(A0)[i];
            reinterpret_cast<float4*>(g0)[i] = make_float4(grad_seed(a.x, lo.x, hi.x), grad_seed(a.y, lo.y, hi.y),
                                                           grad_seed(a.z, lo.z, hi.z), grad_seed(a.w, lo.w, hi.w));
        }
        return;
    }
    const size_t total = (size_t)Bd * S * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % n);
        const int b = (int)((i / n) % Bd);
        g0[i] = grad_seed(A0[i], __ldg(x_L + (size_t)b * n + k), __ldg(x_U + (size_t)b * n + k));
    }
}

void grad_init(const float* A0, const float* x_L, const float* x_U, float* g0, int Bd, int S,
               int n_in, const int* done, cudaStream_t st) {
    Launch _l(K_GRAD_INIT, st);
    const bool vec = (n_in & 3) == 0 && ((reinterpret_cast<uintptr_t>(A0) | reinterpret_cast<uintptr_t>(x_L) |
                                          reinterpret_cast<uintptr_t>(x_U) | reinterpret_cast<uintptr_t>(g0)) & 15u) == 0;
    const size_t total = (size_t)Bd * S * n_in / (vec ? 4 : 1);
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (vec) k_grad_init<true><<<blocks, 256, 0, st>>>(A0, x_L, x_U, g0, Bd, S, n_in, done);
    else k_grad_init<false><<<blocks, 256, 0, st>>>(A0, x_L, x_U, g0, Bd, S, n_in, done);
}

// ---------------------------------------------------------------------------------------------
// convolution (operators/convolution.py:51-96): direct SIMT kernels, one block per row
// ---------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(256)
k_conv_bwd(const float* __restrict__ A_out, const float* __restrict__ Wt, float* __restrict__ A_in,
           ConvGeom g, int accumulate, const int* done) {
    CB_DONE_CHECK(done);
    extern __shared__ float srow[];
    const size_t r = blockIdx.x;
    const int n_out = g.Cout * g.Hout * g.Wout, n_in = g.Cin * g.Hin * g.Win;
    const float* src = A_out + r * n_out;
    if (SMEM) {
        for (int i = threadIdx.x; i < n_out; i += blockDim.x) srow[i] = src[i];
        __syncthreads();
        src = srow;
    }
    const int HWo = g.Hout * g.Wout;
    for (int o = threadIdx.x; o < n_in; o += blockDim.x) {
        const int wi = o % g.Win, hi = (o / g.Win) % g.Hin, ci = o / (g.Win * g.Hin);
        float acc = 0.f;
        for (int kh = 0; kh < g.KH; ++kh) {
            const int hn = hi + g.ph - kh * g.dh;
            if (hn < 0 || hn % g.sh != 0) continue;
            const int ho = hn / g.sh;
            if (ho >= g.Hout) continue;
            for (int kw = 0; kw < g.KW; ++kw) {
                const int wn = wi + g.pw - kw * g.dw;
                if (wn < 0 || wn % g.sw != 0) continue;
                const int wo = wn / g.sw;
                if (wo >= g.Wout) continue;
                const float* wp = Wt + ((size_t)(ci * g.KH + kh) * g.KW + kw) * g.Cout;
                const float* ap = src + ho * g.Wout + wo;
                for (int co = 0; co < g.Cout; ++co) acc = fmaf(ap[co * HWo], __ldg(wp + co), acc);
            }
        }
        float* p = A_in + r * n_in + o;
        *p = accumulate ? (*p + acc) : acc;
    }
}

void conv_bwd(const float* A_out, const float* Wt, float* A_in, const ConvGeom& g, int rows,
              bool accumulate, const int* done, cudaStream_t st) {
    Launch _l(K_CONV_BWD, st);
    const size_t smem = (size_t)g.Cout * g.Hout * g.Wout * sizeof(float);
    if (smem <= 48 * 1024)
        k_conv_bwd<true><<<rows, 256, smem, st>>>(A_out, Wt, A_in, g, accumulate, done);
    else
        k_conv_bwd<false><<<rows, 256, 0, st>>>(A_out, Wt, A_in, g, accumulate, done);
}

template <bool SMEM>
__global__ void __launch_bounds__(256)
k_conv_fwd(const float* __restrict__ g_in, const float* __restrict__ W, const float* __restrict__ bias,
           float* __restrict__ g_out, ConvGeom g, const int* done) {
    CB_DONE_CHECK(done);
    extern __shared__ float srow[];
    const size_t r = blockIdx.x;
    const int n_out = g.Cout * g.Hout * g.Wout, n_in = g.Cin * g.Hin * g.Win;
    const float* src = g_in + r * n_in;
    if (SMEM) {
        for (int i = threadIdx.x; i < n_in; i += blockDim.x) srow[i] = src[i];
        __syncthreads();
        src = srow;
    }
    const int HWi = g.Hin * g.Win;
    for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
        const int wo = o % g.Wout, ho = (o / g.Wout) % g.Hout, co = o / (g.Wout * g.Hout);
        float acc = bias ? __ldg(bias + co) : 0.f;
        for (int kh = 0; kh < g.KH; ++kh) {
            const int hi = ho * g.sh - g.ph + kh * g.dh;
            if (hi < 0 || hi >= g.Hin) continue;
            for (int kw = 0; kw < g.KW; ++kw) {
                const int wi = wo * g.sw - g.pw + kw * g.dw;
                if (wi < 0 || wi >= g.Win) continue;
                const float* wp = W + ((size_t)co * g.Cin * g.KH + kh) * g.KW + kw;
                const float* ip = src + hi * g.Win + wi;
                for (int ci = 0; ci < g.Cin; ++ci)
                    acc = fmaf(ip[ci * HWi], __ldg(wp + (size_t)ci * g.KH * g.KW), acc);
            }
        }
        g_out[r * n_out + o] = acc;
    }
}

void conv_fwd(const float* g_in, const float* W, const float* b, float* g_out, const ConvGeom& g,
              int rows, const int* done, cudaStream_t st) {
    Launch _l(K_CONV_FWD, st);
    const size_t smem = (size_t)g.Cin * g.Hin * g.Win * sizeof(float);
    if (smem <= 48 * 1024)
        k_conv_fwd<true><<<rows, 256, smem, st>>>(g_in, W, b, g_out, g, done);
    else
        k_conv_fwd<false><<<rows, 256, 0, st>>>(g_in, W, b, g_out, g, done);
}

template <int TPR>
__global__ void __launch_bounds__(256)
k_chan_rowdot(const float* __restrict__ A, const float* __restrict__ vec,
              float* __restrict__ bias_rows, int rows, int C, int HW, const int* done) {
    CB_DONE_CHECK(done);
    __shared__ float red[8];
    const size_t r = (size_t)blockIdx.x * (256 / TPR) + threadIdx.x / TPR;
    const int lane = threadIdx.x % TPR;
    const bool active = r < (size_t)rows;
    const int n = C * HW;
    float part = 0.f;
    if (active) {
        const float* a = A + r * n;
        for (int i = lane; i < n; i += TPR) part = fmaf(a[i], __ldg(vec + i / HW), part);
    }
    const float tot = row_sum<TPR>(part, red);
    if (active && lane == 0) bias_rows[r] += tot;
}

void chan_rowdot(const float* A, const float* vec, float* bias_rows, int rows, int C, int HW,
                 const int* done, cudaStream_t st) {
    Launch _l(K_CHAN, st);
    if (C * HW <= 2048)
        k_chan_rowdot<32><<<(rows + 7) / 8, 256, 0, st>>>(A, vec, bias_rows, rows, C, HW, done);
    else
        k_chan_rowdot<256><<<rows, 256, 0, st>>>(A, vec, bias_rows, rows, C, HW, done);
}

__global__ void k_chan_affine(const float* __restrict__ in, float* __restrict__ out,
                              const float* __restrict__ scale, const float* __restrict__ shift,
                              size_t total, int C, int HW, int accumulate, const int* done) {
    CB_DONE_CHECK(done);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)((i / HW) % C);
        float v = in[i] * __ldg(scale + c);
        if (shift) v += __ldg(shift + c);
        out[i] = accumulate ? (out[i] + v) : v;
    }
}

static inline unsigned ew_blocks(size_t total) {
    const size_t b = (total + 255) / 256;
    return (unsigned)(b < (size_t)148 * 32 ? (b ? b : 1) : (size_t)148 * 32);
}

void chan_affine(const float* in, float* out, const float* scale, const float* shift, int rows,
                 int C, int HW, bool accumulate, const int* done, cudaStream_t st) {
    Launch _l(K_CHAN, st);
    const size_t total = (size_t)rows * C * HW;
    k_chan_affine<<<ew_blocks(total), 256, 0, st>>>(in, out, scale, shift, total, C, HW,
                                                   accumulate ? 1 : 0, done);
}

__global__ void k_axpy(const float* __restrict__ in, float* __restrict__ out, float sgn,
                       size_t n, int accumulate, const int* done) {
    CB_DONE_CHECK(done);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = accumulate ? (out[i] + sgn * in[i]) : sgn * in[i];
}

void axpy(const float* in, float* out, float sgn, size_t n, bool accumulate, const int* done,
          cudaStream_t st) {
    Launch _l(K_ELEMWISE, st);
    k_axpy<<<ew_blocks(n), 256, 0, st>>>(in, out, sgn, n, accumulate ? 1 : 0, done);
}

__global__ void k_add2(const float* __restrict__ a, const float* __restrict__ b,
                       float* __restrict__ out, float sgn, size_t n, const int* done) {
    CB_DONE_CHECK(done);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = a[i] + sgn * b[i];
}

void add2(const float* a, const float* b, float* out, float sgn, size_t n, const int* done,
          cudaStream_t st) {
    Launch _l(K_ELEMWISE, st);
    k_add2<<<ew_blocks(n), 256, 0, st>>>(a, b, out, sgn, n, done);
}

__global__ void k_add_rowvec(const float* __restrict__ in, const float* __restrict__ vec,
                             float* __restrict__ out, size_t total, int n, const int* done) {
    CB_DONE_CHECK(done);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = in[i] + __ldg(vec + (i % n));
}

void add_rowvec(const float* in, const float* vec, float* out, int rows, int n, const int* done,
                cudaStream_t st) {
    Launch _l(K_ELEMWISE, st);
    k_add_rowvec<<<ew_blocks((size_t)rows * n), 256, 0, st>>>(in, vec, out, (size_t)rows * n, n, done);
}

__global__ void k_fill_zero(float* __restrict__ p, size_t n, const int* done) {
    CB_DONE_CHECK(done);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        p[i] = 0.f;
}

void fill_zero(float* p, size_t n, const int* done, cudaStream_t st) {
    Launch _l(K_ELEMWISE, st);
    if (n == 0) return;
    k_fill_zero<<<ew_blocks(n), 256, 0, st>>>(p, n, done);
}

// ---------------------------------------------------------------------------------------------
// keep-best bookkeeping of the optimisation loop (auto_LiRPA/optimized_bounds.py:420-514)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float red_specs(const float* p, int S) {
    // loss_reduction_func = sum over specs, applied only when S != 1 (optimized_bounds.py:407-413)
    if (S == 1) return p[0];
    float t = 0.f;
    for (int s = 0; s < S; ++s) t += p[s];
    return t;
}

__global__ void k_keepbest_a(int iter, const float* __restrict__ lb_cur,
                             const float* __restrict__ rhs, float* __restrict__ best_l,
                             float* __restrict__ best_ret, float* __restrict__ ret0,
                             uint8_t* __restrict__ stopped, uint8_t* __restrict__ mask0,
                             OptState* st_cur, int Bd, int S) {
    if (st_cur->done) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Bd) return;
    const float* full = lb_cur + (size_t)b * S;
    float* bl = best_l + (size_t)b * S;
    float* br = best_ret + (size_t)b * S;
    float* r0 = ret0 + (size_t)b * S;
    if (iter == 0) {
        for (int s = 0; s < S; ++s) {
            bl[s] = -INFINITY;
            br[s] = full[s];
            r0[s] = full[s];
        }
    }
    bool stop = false;
    if (rhs) for (int s = 0; s < S; ++s) stop = stop || (full[s] > rhs[(size_t)b * S + s]);
    stopped[b] = stop ? 1 : 0;
    const float fr = red_specs(full, S);
    if (fr > red_specs(bl, S)) {
        for (int s = 0; s < S; ++s) {
            bl[s] = fmaxf(full[s], bl[s]);
            br[s] = fmaxf(full[s], br[s]);
        }
        atomicOr(&st_cur->any_improved, 1);
    }
    if (!stop) atomicAdd(&st_cur->n_not_stopped, 1);
    const bool m0 = fr > red_specs(r0, S);
    mask0[b] = m0 ? 1 : 0;
    if (m0) atomicOr(&st_cur->any_mask0, 1);
}

void keepbest_a(int iter, const float* lb_cur, const float* rhs, float* best_l, float* best_ret,
                float* ret0, uint8_t* stopped, uint8_t* mask0, OptState* st_cur, int Bd, int S,
                cudaStream_t st) {
    Launch _l(K_KEEPBEST, st);
    k_keepbest_a<<<(Bd + 255) / 256, 256, 0, st>>>(iter, lb_cur, rhs, best_l, best_ret, ret0,
                                                   stopped, mask0, st_cur, Bd, S);
}

__global__ void k_keepbest_b(int iter, int iteration, int save_from, int patience_limit,
                             const float* __restrict__ lb_cur, float* __restrict__ ret0,
                             const uint8_t* __restrict__ mask0, uint8_t* __restrict__ snap,
                             const OptState* st_cur, OptState* st_next, int Bd, int S) {
    if (st_cur->done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *st_next = *st_cur;
        return;
    }
    const int patience = st_cur->any_improved ? 0 : st_cur->patience + 1;
    const bool stop_final = st_cur->n_not_stopped == 0;
    // save window: first iteration, second half, or just before leaving (optimized_bounds.py:483-484)
    const bool window = (iter < 1) || (iter > save_from) || stop_final || (patience == patience_limit);
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < Bd) {
        uint8_t sn = 0;
        if (window) {
            if (st_cur->any_mask0) {
                if (mask0[b]) {
                    for (int s = 0; s < S; ++s) ret0[(size_t)b * S + s] = lb_cur[(size_t)b * S + s];
                    sn = 1;
                }
            } else {
                // reference quirk: `ret_0[None] = full_ret_l[None]` overwrites every domain
                for (int s = 0; s < S; ++s) ret0[(size_t)b * S + s] = lb_cur[(size_t)b * S + s];
            }
        }
        snap[b] = sn;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        OptState n;
        n.patience = patience;
        n.any_improved = 0;
        n.n_not_stopped = 0;
        n.any_mask0 = 0;
        n.done = (stop_final || patience > patience_limit || iter == iteration - 1) ? 1 : 0;
        n.n_iter = iter + 1;
        n.pad[0] = n.pad[1] = 0;
        *st_next = n;
    }
}

void keepbest_b(int iter, int iteration, int save_from, int patience_limit, const float* lb_cur,
                float* ret0, const uint8_t* mask0, uint8_t* snap, const OptState* st_cur,
                OptState* st_next, int Bd, int S, cudaStream_t st) {
    Launch _l(K_KEEPBEST, st);
    k_keepbest_b<<<(Bd + 255) / 256, 256, 0, st>>>(iter, iteration, save_from, patience_limit,
                                                   lb_cur, ret0, mask0, snap, st_cur, st_next, Bd, S);
}

// Start of an optimisation: g = m = v = 0, best = p for every optimisable tensor (optimized_bounds.py:71-90) in ONE
// launch (blockIdx.y = tensor) instead of three memsets and a copy per tensor.
__global__ void k_opt_init(const RowTable* __restrict__ tabs) {
    const RowTable t = tabs[blockIdx.y];
    const size_t total = (size_t)t.rows * t.cols;
    const bool vec = (total & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                       reinterpret_cast<uintptr_t>(t.v) | reinterpret_cast<uintptr_t>(t.best)) & 15u) == 0;
    if (vec) {
        const size_t n4 = total >> 2;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            reinterpret_cast<float4*>(t.g)[i] = z;
            reinterpret_cast<float4*>(t.m)[i] = z;
            reinterpret_cast<float4*>(t.v)[i] = z;
            reinterpret_cast<float4*>(t.best)[i] = reinterpret_cast<const float4*>(t.p)[i];
        }
        return;
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        t.g[i] = 0.f;
        t.m[i] = 0.f;
        t.v[i] = 0.f;
        t.best[i] = t.p[i];
    }
}

void opt_init(const RowTable* d_tables, int n_tables, int max_rows, int max_cols, cudaStream_t st) {
    Launch _l(K_SNAPSHOT, st);
    if (n_tables == 0) return;
    dim3 grid(ew_blocks((size_t)max_rows * max_cols / 4), n_tables);
    k_opt_init<<<grid, 256, 0, st>>>(d_tables);
}

// One launch over all optimisable tensors: blockIdx.y = tensor, grid-stride over its elements.
__global__ void k_snapshot(const RowTable* __restrict__ tabs, const uint8_t* __restrict__ snap, int Bd) {
    const RowTable t = tabs[blockIdx.y];
    const size_t total = (size_t)t.rows * t.cols;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)((i / t.cols) % Bd);
        if (snap[b]) t.best[i] = t.p[i];
    }
}

void snapshot(const RowTable* d_tables, int n_tables, int max_rows, int max_cols,
              const uint8_t* snap, int Bd, cudaStream_t st) {
    Launch _l(K_SNAPSHOT, st);
    if (n_tables == 0) return;
    dim3 grid(ew_blocks((size_t)max_rows * max_cols), n_tables);
    k_snapshot<<<grid, 256, 0, st>>>(d_tables, snap, Bd);
}

// torch.optim.Adam (betas=(0.9,0.999), eps=1e-8, single-tensor path) on loss = -sum lb over the
// domains that are not yet verified, then the reference's clamps (optimized_bounds.py:565-575).
// The keep-best snapshot of the SAME iteration (optimized_bounds.py:483-514: best <- p for the domains
// flagged in `snap`, taken before the step) is fused in: one read of p serves both.
template <bool VEC>
__global__ void __launch_bounds__(256)
k_adam(const RowTable* __restrict__ tabs, const uint8_t* __restrict__ stopped,
       const uint8_t* __restrict__ snap, int Bd, float step_a, float step_b, float bc2_sqrt,
       const int* done) {
    // `done` (the reference left its loop in this iteration) cancels the step but not the snapshot,
    // which the reference takes before breaking (optimized_bounds.py:483-530)
    const bool dn = done != nullptr && *done != 0;
    if (dn && snap == nullptr) return;
    const RowTable t = tabs[blockIdx.y];
    if (t.fused) return;                                  // stepped (and snapshotted) inside relu_grad
    const size_t total = (size_t)t.rows * t.cols;
    const float step = t.group == 1 ? step_b : step_a;
    if (VEC && (t.cols & 3) == 0 && total < (1ull << 32)) {
        // 32-bit indices: the row -> sub-domain division is the only integer work per 112 bytes moved
        const uint32_t n4 = (uint32_t)(total >> 2);
        const uint32_t c4 = (uint32_t)t.cols >> 2;
        float4* P = reinterpret_cast<float4*>(t.p);
        float4* Mv = reinterpret_cast<float4*>(t.m);
        float4* Vv = reinterpret_cast<float4*>(t.v);
        float4* Bv = reinterpret_cast<float4*>(t.best);
        const float4* G = reinterpret_cast<const float4*>(t.g);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
            const uint32_t b = (i / c4) % (uint32_t)Bd;
            float4 p = P[i];
            if (snap && snap[b]) Bv[i] = p;
            if (dn) continue;
            const float4 g = G[i];
            float4 m = Mv[i], v = Vv[i];
            const bool stop = stopped[b] != 0;
            p.x = adam_one(p.x, g.x, m.x, v.x, stop, step, bc2_sqrt, t.group);
            p.y = adam_one(p.y, g.y, m.y, v.y, stop, step, bc2_sqrt, t.group);
            p.z = adam_one(p.z, g.z, m.z, v.z, stop, step, bc2_sqrt, t.group);
            p.w = adam_one(p.w, g.w, m.w, v.w, stop, step, bc2_sqrt, t.group);
            Mv[i] = m; Vv[i] = v; P[i] = p;
        }
        return;
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)((i / t.cols) % Bd);
        float p = t.p[i];
        if (snap && snap[b]) t.best[i] = p;
        if (dn) continue;
        float m = t.m[i], v = t.v[i];
        p = adam_one(p, t.g[i], m, v, stopped[b] != 0, step, bc2_sqrt, t.group);
        t.m[i] = m;
        t.v[i] = v;
        t.p[i] = p;
    }
}

void adam_step(const RowTable* d_tables, int n_tables, int max_rows, int max_cols,
               const uint8_t* stopped, const uint8_t* snap, int Bd, float lr_alpha, float lr_beta, float bc1,
               float bc2_sqrt, bool vec_ok, const int* done, cudaStream_t st) {
    Launch _l(K_ADAM, st);
    if (n_tables == 0) return;
    const size_t work = (size_t)max_rows * max_cols / (vec_ok ? 4 : 1);
    dim3 grid(ew_blocks(work), n_tables);
    if (vec_ok)
        k_adam<true><<<grid, 256, 0, st>>>(d_tables, stopped, snap, Bd, lr_alpha / bc1, lr_beta / bc1, bc2_sqrt, done);
    else
        k_adam<false><<<grid, 256, 0, st>>>(d_tables, stopped, snap, Bd, lr_alpha / bc1, lr_beta / bc1, bc2_sqrt, done);
}

// p <- best for every optimisable tensor, lb <- best bounds.  With `snap` the keep-best snapshot of the last iteration
// (k_snapshot: best <- p for the flagged sub-domains) is folded in: a flagged row keeps its p - it is what the snapshot
// would have stored and this kernel copied back - and only the other rows are copied from best.
__global__ void k_finalize(const RowTable* __restrict__ tabs, int n_tables,
                           const float* __restrict__ best_ret, float* __restrict__ lb_out, int nlb,
                           const uint8_t* __restrict__ snap, int Bd) {
    if ((int)blockIdx.y == n_tables) {
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)nlb;
             i += (size_t)gridDim.x * blockDim.x)
            lb_out[i] = best_ret[i];
        return;
    }
    const RowTable t = tabs[blockIdx.y];
    const size_t total = (size_t)t.rows * t.cols;
    if ((t.cols & 3) == 0 && total < (1ull << 32) && ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.best)) & 15u) == 0) {
        const uint32_t n4 = (uint32_t)(total >> 2), c4 = (uint32_t)t.cols >> 2;
        float4* P = reinterpret_cast<float4*>(t.p);
        const float4* B = reinterpret_cast<const float4*>(t.best);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
            if (snap != nullptr && snap[(i / c4) % (uint32_t)Bd]) continue;
            P[i] = B[i];
        }
        return;
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        if (snap != nullptr && snap[(i / t.cols) % Bd]) continue;
        t.p[i] = t.best[i];
    }
}

void finalize(const RowTable* d_tables, int n_tables, int max_rows, int max_cols,
              const float* best_ret, float* lb_out, int nlb, const uint8_t* snap, int Bd, cudaStream_t st) {
    Launch _l(K_SNAPSHOT, st);
    size_t mx = (size_t)max_rows * max_cols;
    if ((size_t)nlb > mx) mx = nlb;
    dim3 grid(ew_blocks(mx), n_tables + 1);
    k_finalize<<<grid, 256, 0, st>>>(d_tables, n_tables, best_ret, lb_out, nlb, snap, Bd);
}

}  // namespace cb
