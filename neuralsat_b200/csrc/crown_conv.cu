// Register-tiled fp32 convolution kernels of the CROWN pass for sm_100a (operators/convolution.py:51-96).
//
//   conv_bwd_tiled : A_in[r,ci,hi,wi] (+)= sum_{co,kh,kw} A_out[r,co,ho,wo] W[co,ci,kh,kw],  hi = ho*s - p + kh*d
//                    (conv_transpose2d of the coefficient matrix; the output_padding of the reference is implied
//                    by writing exactly the Hin x Win positions)
//   conv_fwd_tiled : g_out[r,co,ho,wo] = b[co] + sum_{ci,kh,kw} g_in[r,ci,hi,wi] W[co,ci,kh,kw]   (gradient direction)
//
// One CTA per sub-domain row: the row's source map is staged once in shared memory (coalesced float4), the
// layer's weights too when they fit.  A thread owns one output position and CT consecutive channels held in
// registers; lanes of a warp are consecutive positions of the SAME channel tile, so the weight vector of a
// (tap, source channel) is one 128-bit shared-memory broadcast per 4 FMAs and the source value one 32-bit load.
// For strided transposed convolutions the output positions are walked by residue class (hi mod s, wi mod s):
// within a class every position has the same valid taps, which keeps the tap loop free of divergence.
// Weights are re-laid out once per plan: bwd [KH,KW,Cout,CinP], fwd [Cin,KH,KW,CoutP] (P: padded to CT).
#include "crown_kernels.cuh"

namespace cb {

namespace {

#define CB_DONE_CHECK(done) do { if ((done) != nullptr && *(done) != 0) return; } while (0)

__device__ __forceinline__ void stage_row(float* dst, const float* __restrict__ src, int n) {
    if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) d4[i] = s4[i];
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
}

constexpr int CONV_KMAX = 8;          // kernel extent the tiled transposed convolution supports

template <int CT>
__device__ __forceinline__ void fma_tile(float (&acc)[CT], float av, const float* __restrict__ w) {
#pragma unroll
    for (int j = 0; j < CT; j += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + j);
        acc[j] = fmaf(av, w4.x, acc[j]);
        acc[j + 1] = fmaf(av, w4.y, acc[j + 1]);
        acc[j + 2] = fmaf(av, w4.z, acc[j + 2]);
        acc[j + 3] = fmaf(av, w4.w, acc[j + 3]);
    }
}

template <int CT>
__global__ void __launch_bounds__(256)
k_conv_bwd_tiled(const float* __restrict__ A_out, const float* __restrict__ Wk, float* __restrict__ A_in,
                 ConvGeom g, int CinP, int accumulate, int w_smem, const int* done) {
    CB_DONE_CHECK(done);
    extern __shared__ __align__(16) float sm[];
    const size_t r = blockIdx.x;
    const int HWo = g.Hout * g.Wout;
    const int n_out = g.Cout * HWo, n_in = g.Cin * g.Hin * g.Win;
    float* const sA = sm;
    float* const sW = sm + ((n_out + 3) & ~3);
    stage_row(sA, A_out + r * n_out, n_out);
    const int w_elems = g.KH * g.KW * g.Cout * CinP;
    if (w_smem) stage_row(sW, Wk, w_elems);
    __syncthreads();
    const float* __restrict__ Wp = w_smem ? sW : Wk;
    const int n_tiles = CinP / CT;
    const int w_tap = g.Cout * CinP;
    float* const orow = A_in + r * n_in;
    // all residue classes are walked as ONE item space (class-major, each class padded to whole warps), so that
    // small maps still fill the CTA: item -> (class, channel tile, position in class)
    const int n_cls = g.sh * g.sw;
    for (int base = 0, cls = 0; cls < n_cls; ++cls) {
        const int ph = cls / g.sw, pw = cls - ph * g.sw;
        const int na = ph < g.Hin ? (g.Hin - ph + g.sh - 1) / g.sh : 0;
        const int nb = pw < g.Win ? (g.Win - pw + g.sw - 1) / g.sw : 0;
        const int npos = na * nb;
        const int npos32 = (npos + 31) & ~31;
        const int cnt = n_tiles * npos32;
        // taps that reach this class: ho = a + oh, wo = b + ow with class constants oh / ow (exact divisions)
        int th[CONV_KMAX], oh[CONV_KMAX], tw[CONV_KMAX], ow[CONV_KMAX];
        int nth = 0, ntw = 0;
        for (int kh = 0; kh < g.KH; ++kh) {
            const int v = ph + g.ph - kh * g.dh;
            if (v % g.sh == 0) { th[nth] = kh; oh[nth] = v / g.sh; ++nth; }
        }
        for (int kw = 0; kw < g.KW; ++kw) {
            const int v = pw + g.pw - kw * g.dw;
            if (v % g.sw == 0) { tw[ntw] = kw; ow[ntw] = v / g.sw; ++ntw; }
        }
        // first item of this class owned by this thread: smallest item >= base with item % blockDim == threadIdx
        int first = base + (((int)threadIdx.x - base) % (int)blockDim.x + (int)blockDim.x) % (int)blockDim.x;
        for (int item = first - base; item < cnt; item += blockDim.x) {
            const int t = item / npos32;
            const int pidx = item - t * npos32;
            if (pidx >= npos) continue;
            const int a = pidx / nb, b = pidx - a * nb;
            const int hi = ph + a * g.sh, wi = pw + b * g.sw;
            float acc[CT];
#pragma unroll
            for (int j = 0; j < CT; ++j) acc[j] = 0.f;
            const float* const wt = Wp + t * CT;
            for (int ih = 0; ih < nth; ++ih) {
                const int ho = a + oh[ih];
                if (ho < 0 || ho >= g.Hout) continue;
                for (int iw = 0; iw < ntw; ++iw) {
                    const int wo = b + ow[iw];
                    if (wo < 0 || wo >= g.Wout) continue;
                    const float* ap = sA + ho * g.Wout + wo;
                    const float* wp = wt + (th[ih] * g.KW + tw[iw]) * w_tap;
#pragma unroll 4
                    for (int co = 0; co < g.Cout; ++co, ap += HWo, wp += CinP) fma_tile<CT>(acc, *ap, wp);
                }
            }
#pragma unroll
            for (int j = 0; j < CT; ++j) {
                const int ci = t * CT + j;
                if (ci < g.Cin) {
                    float* p = orow + ((size_t)ci * g.Hin + hi) * g.Win + wi;
                    *p = accumulate ? (*p + acc[j]) : acc[j];
                }
            }
        }
        base += cnt;
    }
}

template <int CT>
__global__ void __launch_bounds__(256)
k_conv_fwd_tiled(const float* __restrict__ g_in, const float* __restrict__ Wk, const float* __restrict__ bias,
                 float* __restrict__ g_out, ConvGeom g, int CoutP, int w_smem, const int* done) {
    CB_DONE_CHECK(done);
    extern __shared__ __align__(16) float sm[];
    const size_t r = blockIdx.x;
    const int HWi = g.Hin * g.Win, HWo = g.Hout * g.Wout;
    const int n_in = g.Cin * HWi, n_out = g.Cout * HWo;
    float* const sI = sm;
    float* const sW = sm + ((n_in + 3) & ~3);
    stage_row(sI, g_in + r * n_in, n_in);
    const int w_elems = g.Cin * g.KH * g.KW * CoutP;
    if (w_smem) stage_row(sW, Wk, w_elems);
    __syncthreads();
    const float* __restrict__ Wp = w_smem ? sW : Wk;
    const int n_tiles = CoutP / CT;
    const int npos32 = (HWo + 31) & ~31;
    const int w_ci = g.KH * g.KW * CoutP;
    float* const orow = g_out + r * n_out;
    for (int item = threadIdx.x; item < n_tiles * npos32; item += blockDim.x) {
        const int t = item / npos32;
        const int pidx = item - t * npos32;
        if (pidx >= HWo) continue;
        const int ho = pidx / g.Wout, wo = pidx - ho * g.Wout;
        float acc[CT];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            const int co = t * CT + j;
            acc[j] = (bias && co < g.Cout) ? __ldg(bias + co) : 0.f;
        }
        for (int kh = 0; kh < g.KH; ++kh) {
            const int hi = ho * g.sh - g.ph + kh * g.dh;
            if (hi < 0 || hi >= g.Hin) continue;
            for (int kw = 0; kw < g.KW; ++kw) {
                const int wi = wo * g.sw - g.pw + kw * g.dw;
                if (wi < 0 || wi >= g.Win) continue;
                const float* ip = sI + hi * g.Win + wi;
                const float* wp = Wp + (size_t)(kh * g.KW + kw) * CoutP + t * CT;
#pragma unroll 4
                for (int ci = 0; ci < g.Cin; ++ci, ip += HWi, wp += w_ci) fma_tile<CT>(acc, *ip, wp);
            }
        }
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            const int co = t * CT + j;
            if (co < g.Cout) orow[(size_t)co * HWo + pidx] = acc[j];
        }
    }
}

// W [Cout,Cin,KH,KW] -> bwd layout [KH,KW,Cout,CinP] / fwd layout [Cin,KH,KW,CoutP], zero padded
__global__ void k_conv_relayout(const float* __restrict__ W, float* __restrict__ out, int Cout, int Cin, int KHW,
                                int P, int fwd) {
    const size_t total = fwd ? (size_t)Cin * KHW * P : (size_t)KHW * Cout * P;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int co, ci, k;
        if (fwd) { co = (int)(i % P); k = (int)((i / P) % KHW); ci = (int)(i / ((size_t)P * KHW)); }
        else { ci = (int)(i % P); co = (int)((i / P) % Cout); k = (int)(i / ((size_t)P * Cout)); }
        out[i] = (co < Cout && ci < Cin) ? W[((size_t)co * Cin + ci) * KHW + k] : 0.f;
    }
}

constexpr size_t CONV_SMEM_MAX = 200 * 1024;

template <typename K>
bool opt_in(K kernel, size_t smem) {
    if (smem <= 48 * 1024) return true;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV_SMEM_MAX) == cudaSuccess;
}

}  // namespace

int conv_pad(int c) { return c <= 4 ? 4 : (c + 7) / 8 * 8; }

void conv_relayout(const float* W, float* out, int Cout, int Cin, int KHW, bool fwd, cudaStream_t st) {
    const int P = conv_pad(fwd ? Cout : Cin);
    const size_t total = fwd ? (size_t)Cin * KHW * P : (size_t)KHW * Cout * P;
    k_conv_relayout<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, out, Cout, Cin, KHW, P, fwd ? 1 : 0);
}

bool conv_bwd_tiled(const float* A_out, const float* Wk, float* A_in, const ConvGeom& g, int rows, bool accumulate,
                    const int* done, cudaStream_t st) {
    const int CinP = conv_pad(g.Cin);
    const size_t map = (size_t)((g.Cout * g.Hout * g.Wout + 3) & ~3) * sizeof(float);
    const size_t wbytes = (size_t)g.KH * g.KW * g.Cout * CinP * sizeof(float);
    if (map > CONV_SMEM_MAX || g.KH > CONV_KMAX || g.KW > CONV_KMAX) return false;
    const int w_smem = map + wbytes <= CONV_SMEM_MAX / 2 ? 1 : 0;      // keep >= 2 CTAs per SM when staging weights
    const size_t smem = map + (w_smem ? wbytes : 0);
    if (!(CinP == 4 ? opt_in(k_conv_bwd_tiled<4>, smem) : opt_in(k_conv_bwd_tiled<8>, smem))) return false;
    Launch _l(K_CONV_BWD, st);
    if (CinP == 4)
        k_conv_bwd_tiled<4><<<rows, 256, smem, st>>>(A_out, Wk, A_in, g, CinP, accumulate, w_smem, done);
    else
        k_conv_bwd_tiled<8><<<rows, 256, smem, st>>>(A_out, Wk, A_in, g, CinP, accumulate, w_smem, done);
    return true;
}

bool conv_fwd_tiled(const float* g_in, const float* Wk, const float* b, float* g_out, const ConvGeom& g, int rows,
                    const int* done, cudaStream_t st) {
    const int CoutP = conv_pad(g.Cout);
    const size_t map = (size_t)((g.Cin * g.Hin * g.Win + 3) & ~3) * sizeof(float);
    const size_t wbytes = (size_t)g.Cin * g.KH * g.KW * CoutP * sizeof(float);
    if (map > CONV_SMEM_MAX) return false;
    const int w_smem = map + wbytes <= CONV_SMEM_MAX / 2 ? 1 : 0;
    const size_t smem = map + (w_smem ? wbytes : 0);
    if (!(CoutP == 4 ? opt_in(k_conv_fwd_tiled<4>, smem) : opt_in(k_conv_fwd_tiled<8>, smem))) return false;
    Launch _l(K_CONV_FWD, st);
    if (CoutP == 4)
        k_conv_fwd_tiled<4><<<rows, 256, smem, st>>>(g_in, Wk, b, g_out, g, CoutP, w_smem, done);
    else
        k_conv_fwd_tiled<8><<<rows, 256, smem, st>>>(g_in, Wk, b, g_out, g, CoutP, w_smem, done);
    return true;
}

}  // namespace cb
