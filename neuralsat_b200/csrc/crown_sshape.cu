// S-shaped activations (sigmoid / tanh) of the CROWN pass for sm_100a: relaxation lines, the sign-split
// multiply and its gradient w.r.t. the tangent points (auto_LiRPA/operators/tanh.py:135-290,
// operators/activation_base.py:247-304, operators/clampmult.py:17-95).
//
// Only the lower-bound side is computed (the BaB loop never asks for upper bounds), i.e. planes 0, 2,
// 4, 6 of the reference's alpha tensor [8, S1, Bd, n] = (tp_pos, tp_neg, tp_both_lower, tp_both_upper)
// feed the lines; all 8 planes are clipped in place before a pass, as the reference does
// (operators/tanh.py:191-198).  HBM-bound elementwise work: one thread per (domain, neuron), looping
// over the spec rows, every access coalesced along the neuron dimension.
#include "crown_kernels.cuh"

namespace cb {

namespace {

#define CB_DONE_CHECK(done) do { if ((done) != nullptr && *(done) != 0) return; } while (0)

template <bool TANH>
__device__ __forceinline__ float f_act(float x) {
    if (TANH) return tanhf(x);
    return __fdiv_rn(1.f, 1.f + expf(-x));
}

// operators/tanh.py:8-16
template <bool TANH>
__device__ __forceinline__ float f_d1(float x) {
    if (TANH) {
        const float m = fabsf(x) < 25.f ? 1.f : 0.f;
        const float c = coshf(m * x + 1.f - m);
        return m * __fdiv_rn(1.f, c * c);
    }
    const float s = f_act<false>(x);
    return s * (1.f - s);
}

// derivative of f_d1 as autograd evaluates it
template <bool TANH>
__device__ __forceinline__ float f_d2(float x) {
    if (TANH) {
        if (!(fabsf(x) < 25.f)) return 0.f;
        const float c = coshf(x);
        return __fdiv_rn(-2.f * sinhf(x), c * c * c);
    }
    const float s = f_act<false>(x);
    const float d = s * (1.f - s);
    return d * (1.f - s) - s * d;
}

// operators/tanh.py:150-187: table tangent points valid on [l, u]
__device__ __forceinline__ void table_points(const SshapeArgs& a, float l, float u, float& dl, float& du) {
    long long iu = (long long)__fdiv_rn(u, 0.01f);
    if (iu < 0) iu = 0;
    iu += 1;
    dl = iu < a.table_n ? __ldg(a.d_lower_t + iu) : l;
    long long il = (long long)__fdiv_rn(l, -0.01f);
    if (il < 0) il = 0;
    il += 1;
    du = il < a.table_n ? __ldg(a.d_upper_t + il) : u;
}

struct Lines {
    float lw, lb, uw, ub;       // lower line (used where A >= 0) and upper line (A < 0)
    int lp, up;                 // alpha plane whose tangent gives the line, -1 = parameter-free line
    float ltp, utp;             // the tangent points
};

// tp[0..3] = alpha planes 0, 2, 4, 6 (already clipped) when has_alpha
template <bool TANH>
__device__ __forceinline__ Lines sshape_lines(const SshapeArgs& a, float l, float u, bool has_alpha, const float (&tp)[4]) {
    Lines r;
    const bool pos = l >= 0.f, neg = u <= 0.f, both = !(pos || neg);
    const float y_l = f_act<TANH>(l), y_u = f_act<TANH>(u);
    const float k_direct = __fdiv_rn(y_u - y_l, fmaxf(u - l, 1e-8f));
    const float b_direct = -l * k_direct + y_l;
    r.lw = r.lb = r.uw = r.ub = 0.f;
    r.lp = r.up = -1;
    r.ltp = r.utp = 0.f;
    if (neg) { r.uw = k_direct; r.ub = b_direct; }
    if (pos) { r.lw = k_direct; r.lb = b_direct; }
    float t_bl, t_bu, t_neg, t_pos;
    if (has_alpha) {
        t_pos = tp[0]; t_neg = tp[1]; t_bl = tp[2]; t_bu = tp[3];
    } else {
        table_points(a, l, u, t_bl, t_bu);
        t_pos = t_neg = (l + u) / 2.f;
    }
    if (both) {
        if (k_direct < f_d1<TANH>(l)) { r.lw = k_direct; r.lb = b_direct; }
        else { const float k = f_d1<TANH>(t_bl); r.lw = k; r.lb = -t_bl * k + f_act<TANH>(t_bl); r.lp = 4; r.ltp = t_bl; }
        if (k_direct < f_d1<TANH>(u)) { r.uw = k_direct; r.ub = b_direct; }
        else { const float k = f_d1<TANH>(t_bu); r.uw = k; r.ub = -t_bu * k + f_act<TANH>(t_bu); r.up = 6; r.utp = t_bu; }
    }
    if (neg) { const float k = f_d1<TANH>(t_neg); r.lw = k; r.lb = -t_neg * k + f_act<TANH>(t_neg); r.lp = 2; r.ltp = t_neg; }
    if (pos) { const float k = f_d1<TANH>(t_pos); r.uw = k; r.ub = -t_pos * k + f_act<TANH>(t_pos); r.up = 0; r.utp = t_pos; }
    return r;
}

// in-place clip of all 8 planes (operators/tanh.py:191-198); one thread per (s1, b, i)
__global__ void __launch_bounds__(256) k_sshape_clip(SshapeArgs a, int Bd, int n, const int* done) {
    CB_DONE_CHECK(done);
    const size_t per = (size_t)a.S1 * Bd * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
        const size_t bn = i % ((size_t)Bd * n);
        const float l = __ldg(a.lower + bn), u = __ldg(a.upper + bn);
        float dl, du;
        table_points(a, l, u, dl, du);
#pragma unroll
        for (int p = 0; p < 4; ++p) a.alpha[p * per + i] = fmaxf(fminf(a.alpha[p * per + i], u), l);
#pragma unroll
        for (int p = 4; p < 6; ++p) a.alpha[p * per + i] = fminf(a.alpha[p * per + i], dl);
#pragma unroll
        for (int p = 6; p < 8; ++p) a.alpha[p * per + i] = fmaxf(a.alpha[p * per + i], du);
    }
}

template <bool TANH>
__global__ void __launch_bounds__(256)
k_sshape_bwd(const float* __restrict__ A_post, float* __restrict__ A_pre, int accumulate,
             float* __restrict__ bias_rows, SshapeArgs a, int Bd, int S, int n, const int* done) {
    CB_DONE_CHECK(done);
    const int b = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (b >= Bd) return;
    const bool has_alpha = a.alpha != nullptr;
    const size_t per = (size_t)a.S1 * Bd * n;
    for (int s = 0; s < S; ++s) {
        const size_t r = (size_t)s * Bd + b;
        const size_t arow = ((size_t)(a.S1 == 1 ? 0 : s) * Bd + b) * n;
        float part = 0.f;
        for (int i = lane; i < n; i += 32) {
            float tp[4] = {0.f, 0.f, 0.f, 0.f};
            if (has_alpha) {
#pragma unroll
                for (int p = 0; p < 4; ++p) tp[p] = a.alpha[(size_t)(2 * p) * per + arow + i];
            }
            const Lines ln = sshape_lines<TANH>(a, __ldg(a.lower + (size_t)b * n + i), __ldg(a.upper + (size_t)b * n + i), has_alpha, tp);
            const float av = A_post[r * n + i];
            const float a_pos = fmaxf(av, 0.f), a_neg = fminf(av, 0.f);
            const float v = ln.lw * a_pos + ln.uw * a_neg;
            A_pre[r * n + i] = accumulate ? (A_pre[r * n + i] + v) : v;
            part += a_pos * ln.lb + a_neg * ln.ub;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) bias_rows[r] += part;
    }
}

// g_post = g_pre * w_sel + b_sel;  grad of sum lb w.r.t. the tangent point that produced the selected line:
// d/dtp [ A (k(tp) g + b(tp)) ], k = f', b = f(tp) - tp k  =>  A (g - tp) f''(tp) (+ A (f'_autograd - k) for tanh)
template <bool TANH>
__global__ void __launch_bounds__(256)
k_sshape_grad(const float* __restrict__ A_post, const float* __restrict__ g_pre, float* __restrict__ g_post,
              float* __restrict__ grad_alpha, SshapeArgs a, int Bd, int S, int n, const int* done) {
    CB_DONE_CHECK(done);
    const int b = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (b >= Bd) return;
    const bool has_alpha = a.alpha != nullptr;
    const size_t per = (size_t)a.S1 * Bd * n;
    for (int i = lane; i < n; i += 32) {
        const float l = __ldg(a.lower + (size_t)b * n + i), u = __ldg(a.upper + (size_t)b * n + i);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s = 0; s < S; ++s) {
            const size_t r = (size_t)s * Bd + b;
            const size_t arow = ((size_t)(a.S1 == 1 ? 0 : s) * Bd + b) * n;
            float tp[4] = {0.f, 0.f, 0.f, 0.f};
            if (has_alpha) {
#pragma unroll
                for (int p = 0; p < 4; ++p) tp[p] = a.alpha[(size_t)(2 * p) * per + arow + i];
            }
            const Lines ln = sshape_lines<TANH>(a, l, u, has_alpha, tp);
            const float av = A_post[r * n + i];
            const float gp = g_pre[r * n + i];
            const bool ps = av >= 0.f;                       // the A >= 0 tie rule of clampmult's backward
            if (g_post) g_post[r * n + i] = gp * (ps ? ln.lw : ln.uw) + (ps ? ln.lb : ln.ub);
            if (grad_alpha && has_alpha) {
                const int plane = ps ? ln.lp : ln.up;
                float c = 0.f;
                if (plane >= 0) {
                    const float t = ps ? ln.ltp : ln.utp;
                    c = av * (gp - t) * f_d2<TANH>(t);
                    if (TANH) { const float th = tanhf(t); c += av * ((1.f - th * th) - f_d1<true>(t)); }
                }
                if (a.S1 == 1) {
                    if (plane >= 0) acc[plane >> 1] += c;
                } else {
#pragma unroll
                    for (int p = 0; p < 4; ++p) grad_alpha[(size_t)(2 * p) * per + arow + i] = (plane == 2 * p) ? c : 0.f;
                }
            }
        }
        if (grad_alpha && has_alpha && a.S1 == 1) {
#pragma unroll
            for (int p = 0; p < 4; ++p) grad_alpha[(size_t)(2 * p) * per + (size_t)b * n + i] = acc[p];
        }
    }
}

unsigned blocks_for(size_t total) {
    size_t b = (total + 255) / 256;
    const size_t cap = 148 * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

void sshape_clip(const SshapeArgs& a, int Bd, int n, const int* done, cudaStream_t st) {
    if (!a.alpha) return;
    Launch _l(K_SSHAPE, st);
    k_sshape_clip<<<blocks_for((size_t)a.S1 * Bd * n), 256, 0, st>>>(a, Bd, n, done);
}

void sshape_bwd(const float* A_post, float* A_pre, bool accumulate, float* bias_rows, const SshapeArgs& a,
                int Bd, int S, int n, const int* done, cudaStream_t st) {
    Launch _l(K_SSHAPE, st);
    if (a.is_tanh) k_sshape_bwd<true><<<(Bd + 7) / 8, 256, 0, st>>>(A_post, A_pre, accumulate, bias_rows, a, Bd, S, n, done);
    else k_sshape_bwd<false><<<(Bd + 7) / 8, 256, 0, st>>>(A_post, A_pre, accumulate, bias_rows, a, Bd, S, n, done);
}

void sshape_grad(const float* A_post, const float* g_pre, float* g_post, float* grad_alpha, const SshapeArgs& a,
                 int Bd, int S, int n, const int* done, cudaStream_t st) {
    Launch _l(K_SSHAPE, st);
    if (a.is_tanh) k_sshape_grad<true><<<(Bd + 7) / 8, 256, 0, st>>>(A_post, g_pre, g_post, grad_alpha, a, Bd, S, n, done);
    else k_sshape_grad<false><<<(Bd + 7) / 8, 256, 0, st>>>(A_post, g_pre, g_post, grad_alpha, a, Bd, S, n, done);
}

}  // namespace cb
