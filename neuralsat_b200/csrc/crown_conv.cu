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
#include <cstdlib>

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


// ---- thin convolutions: at most four channels on the image side (the first layer of a convnet) ---------------------
// The tiled kernels above keep one position and a channel tile per thread: with 3 image channels a weight vector
// feeds 4 FMAs per shared-memory read and the kernel is bound by those reads.  Here a thread owns FOUR neighbouring
// positions (bwd: four coarse positions x all S*S residue classes x the image channels; fwd: four output positions
// x eight output channels), so a weight vector read once feeds 12-32 FMAs and the source values sit in registers.
// Kernel extent, stride and padding are template parameters: every tap's residue class and register index is a
// compile-time constant.  Same arithmetic order per output as the tiled kernels is NOT required (fp32 sums, parity
// is checked against torch with the usual tolerance).
__host__ __device__ constexpr int cx_mod(int a, int m) { return ((a % m) + m) % m; }
__host__ __device__ constexpr int cx_cls(int k, int S, int P) { return cx_mod(k - P, S); }              // residue class a tap writes to
__host__ __device__ constexpr int cx_off(int k, int S, int P) { return (cx_cls(k, S, P) + P - k) / S; }   // source offset: ho = a + off

template <int K, int S, int P, int CI>
__global__ void __launch_bounds__(128)
k_conv_bwd_thin(const float* __restrict__ A_out, const float* __restrict__ Wk, float* __restrict__ A_in, ConvGeom g,
                int rows, int accumulate, const int* done) {
    CB_DONE_CHECK(done);
    constexpr int PT = 4;
    constexpr int OMAX = cx_off(0, S, P), OMIN = cx_off(K - 1, S, P);
    constexpr int NR = OMAX - OMIN + 1, NC = PT + NR - 1;
    extern __shared__ __align__(16) float sm[];
    float4* const sW = reinterpret_cast<float4*>(sm);                   // [K*K][Cout] x (4 image channels)
    for (int i = threadIdx.x; i < K * K * g.Cout; i += blockDim.x) sW[i] = reinterpret_cast<const float4*>(Wk)[i];
    __syncthreads();
    const int Hg = (g.Hin + S - 1) / S, Wg = (g.Win + S - 1) / S, WQ = (Wg + PT - 1) / PT;
    const long long n_items = (long long)rows * Hg * WQ;
    const int HWo = g.Hout * g.Wout;
    const bool vec = (g.Win % (PT * S)) == 0 && (PT * S) % 4 == 0;
    for (long long item = blockIdx.x * (long long)blockDim.x + threadIdx.x; item < n_items;
         item += (long long)gridDim.x * blockDim.x) {
        const int bq = (int)(item % WQ);
        const int a = (int)((item / WQ) % Hg);
        const size_t r = (size_t)(item / ((long long)WQ * Hg));
        const int b0 = bq * PT;
        float acc[S * S][PT][CI];
#pragma unroll
        for (int c = 0; c < S * S; ++c)
#pragma unroll
            for (int j = 0; j < PT; ++j)
#pragma unroll
                for (int i = 0; i < CI; ++i) acc[c][j][i] = 0.f;
        const float* const src = A_out + r * (size_t)g.Cout * HWo;
        int roff[NR];
        bool rok[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
            const int ho = a + OMIN + rr;
            rok[rr] = ho >= 0 && ho < g.Hout;
            roff[rr] = ho * g.Wout;
        }
        bool cok[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
            const int wo = b0 + OMIN + cc;
            cok[cc] = wo >= 0 && wo < g.Wout;
        }
#pragma unroll 2
        for (int co = 0; co < g.Cout; ++co) {
            const float* const sp = src + (size_t)co * HWo + (b0 + OMIN);
            float v[NR][NC];
#pragma unroll
            for (int rr = 0; rr < NR; ++rr)
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) v[rr][cc] = (rok[rr] && cok[cc]) ? __ldg(sp + roff[rr] + cc) : 0.f;
#pragma unroll
            for (int kh = 0; kh < K; ++kh)
#pragma unroll
                for (int kw = 0; kw < K; ++kw) {
                    const int cls = cx_cls(kh, S, P) * S + cx_cls(kw, S, P);
                    const int rr = cx_off(kh, S, P) - OMIN, c0 = cx_off(kw, S, P) - OMIN;
                    const float4 w = sW[(kh * K + kw) * g.Cout + co];
#pragma unroll
                    for (int j = 0; j < PT; ++j) {
                        const float x = v[rr][c0 + j];
                        acc[cls][j][0] = fmaf(x, w.x, acc[cls][j][0]);
                        if (CI > 1) acc[cls][j][1 % CI] = fmaf(x, w.y, acc[cls][j][1 % CI]);
                        if (CI > 2) acc[cls][j][2 % CI] = fmaf(x, w.z, acc[cls][j][2 % CI]);
                        if (CI > 3) acc[cls][j][3 % CI] = fmaf(x, w.w, acc[cls][j][3 % CI]);
                    }
                }
        }
        float* const orow = A_in + r * (size_t)g.Cin * g.Hin * g.Win;
#pragma unroll
        for (int ph = 0; ph < S; ++ph) {
            const int hi = a * S + ph;
            if (hi >= g.Hin) continue;
#pragma unroll
            for (int i = 0; i < CI; ++i) {
                if (i >= g.Cin) continue;
                float* const p = orow + ((size_t)i * g.Hin + hi) * g.Win + (size_t)b0 * S;
                if (vec) {
                    // the PT*S values of this thread are consecutive along wi: index j*S + pw
#pragma unroll
                    for (int q4 = 0; q4 < PT * S / 4; ++q4) {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int idx = q4 * 4 + e;
                            o[e] = acc[ph * S + idx % S][idx / S][i];
                        }
                        float4* const p4 = reinterpret_cast<float4*>(p) + q4;
                        float4 val = make_float4(o[0], o[1], o[2], o[3]);
                        if (accumulate) { const float4 old = *p4; val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w; }
                        *p4 = val;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < PT; ++j)
#pragma unroll
                        for (int pw = 0; pw < S; ++pw) {
                            const int wi = (b0 + j) * S + pw;
                            if (wi < g.Win) {
                                float* const q = p + j * S + pw;
                                *q = accumulate ? (*q + acc[ph * S + pw][j][i]) : acc[ph * S + pw][j][i];
                            }
                        }
                }
            }
        }
    }
}

template <int K, int S, int P>
__global__ void __launch_bounds__(128)
k_conv_fwd_thin(const float* __restrict__ g_in, const float* __restrict__ Wk, const float* __restrict__ bias,
                float* __restrict__ g_out, ConvGeom g, int CoutP, int rows, const int* done) {
    CB_DONE_CHECK(done);
    constexpr int PT = 4, CT = 8;
    constexpr int NC = (PT - 1) * S + K;
    extern __shared__ __align__(16) float sm[];
    const int CI = g.Cin;
    const int w_elems = CI * K * K * CoutP;                             // [Cin][K][K][CoutP]
    for (int i = threadIdx.x; i < (w_elems >> 2); i += blockDim.x)
        reinterpret_cast<float4*>(sm)[i] = reinterpret_cast<const float4*>(Wk)[i];
    __syncthreads();
    const int WQ = (g.Wout + PT - 1) / PT;
    const long long n_items = (long long)rows * g.Hout * WQ;
    const int HWi = g.Hin * g.Win, HWo = g.Hout * g.Wout;
    const bool vec = (g.Wout % PT) == 0;
    for (long long item = blockIdx.x * (long long)blockDim.x + threadIdx.x; item < n_items;
         item += (long long)gridDim.x * blockDim.x) {
        const int wq = (int)(item % WQ);
        const int ho = (int)((item / WQ) % g.Hout);
        const size_t r = (size_t)(item / ((long long)WQ * g.Hout));
        const int wo0 = wq * PT;
        const int wi0 = wo0 * S - P, hi0 = ho * S - P;
        const float* const src = g_in + r * (size_t)g.Cin * HWi;
        float* const orow = g_out + r * (size_t)g.Cout * HWo + (size_t)ho * g.Wout + wo0;
        bool cok[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) cok[cc] = wi0 + cc >= 0 && wi0 + cc < g.Win;
        for (int co0 = 0; co0 < CoutP; co0 += CT) {
            float acc[PT][CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float b = (bias != nullptr && co0 + c < g.Cout) ? __ldg(bias + co0 + c) : 0.f;
#pragma unroll
                for (int j = 0; j < PT; ++j) acc[j][c] = b;
            }
#pragma unroll 1
            for (int ci = 0; ci < CI; ++ci) {
                float v[K][NC];
#pragma unroll
                for (int kh = 0; kh < K; ++kh) {
                    const int hi = hi0 + kh;
                    const bool rok = hi >= 0 && hi < g.Hin;
                    const float* const sp = src + (size_t)ci * HWi + hi * g.Win + wi0;
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) v[kh][cc] = (rok && cok[cc]) ? __ldg(sp + cc) : 0.f;
                }
#pragma unroll
                for (int kh = 0; kh < K; ++kh)
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) {
                        const float* const wp = sm + ((ci * K + kh) * K + kw) * CoutP + co0;
                        const float4 w0 = *reinterpret_cast<const float4*>(wp);
                        const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
                        for (int j = 0; j < PT; ++j) {
                            const float x = v[kh][j * S + kw];
                            acc[j][0] = fmaf(x, w0.x, acc[j][0]); acc[j][1] = fmaf(x, w0.y, acc[j][1]);
                            acc[j][2] = fmaf(x, w0.z, acc[j][2]); acc[j][3] = fmaf(x, w0.w, acc[j][3]);
                            acc[j][4] = fmaf(x, w1.x, acc[j][4]); acc[j][5] = fmaf(x, w1.y, acc[j][5]);
                            acc[j][6] = fmaf(x, w1.z, acc[j][6]); acc[j][7] = fmaf(x, w1.w, acc[j][7]);
                        }
                    }
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                if (co0 + c >= g.Cout) continue;
                float* const p = orow + (size_t)(co0 + c) * HWo;
                if (vec) {
                    *reinterpret_cast<float4*>(p) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
                } else {
#pragma unroll
                    for (int j = 0; j < PT; ++j)
                        if (wo0 + j < g.Wout) p[j] = acc[j][c];
                }
            }
        }
    }
}

inline unsigned thin_grid(long long items) {
    const long long blocks = (items + 127) / 128;
    const long long cap = 148ll * 16 * 8;
    return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

template <int K, int S, int P>
bool launch_bwd_thin(const float* A_out, const float* Wk, float* A_in, const ConvGeom& g, int rows, bool accumulate,
                     const int* done, cudaStream_t st) {
    constexpr int PT = 4;
    const int Hg = (g.Hin + S - 1) / S, Wg = (g.Win + S - 1) / S, WQ = (Wg + PT - 1) / PT;
    const size_t smem = (size_t)K * K * g.Cout * 4 * sizeof(float);
    if (smem > 48 * 1024) return false;
    const unsigned grid = thin_grid((long long)rows * Hg * WQ);
    Launch _l(K_CONV_BWD, st);
    if (g.Cin == 1) k_conv_bwd_thin<K, S, P, 1><<<grid, 128, smem, st>>>(A_out, Wk, A_in, g, rows, accumulate, done);
    else if (g.Cin == 3) k_conv_bwd_thin<K, S, P, 3><<<grid, 128, smem, st>>>(A_out, Wk, A_in, g, rows, accumulate, done);
    else k_conv_bwd_thin<K, S, P, 4><<<grid, 128, smem, st>>>(A_out, Wk, A_in, g, rows, accumulate, done);
    return true;
}

template <int K, int S, int P>
bool launch_fwd_thin(const float* g_in, const float* Wk, const float* b, float* g_out, const ConvGeom& g, int CoutP,
                     int rows, const int* done, cudaStream_t st) {
    constexpr int PT = 4;
    const int WQ = (g.Wout + PT - 1) / PT;
    const size_t smem = (size_t)g.Cin * K * K * CoutP * sizeof(float);
    if (smem > 48 * 1024 || (CoutP & 7) != 0) return false;
    const unsigned grid = thin_grid((long long)rows * g.Hout * WQ);
    Launch _l(K_CONV_FWD, st);
    k_conv_fwd_thin<K, S, P><<<grid, 128, smem, st>>>(g_in, Wk, b, g_out, g, CoutP, rows, done);
    return true;
}

// (extent, stride, padding) combinations compiled in; anything else stays on the tiled kernels
#define CB_THIN_CASES(X) X(3, 1, 1) X(3, 2, 1) X(3, 2, 0) X(4, 2, 1) X(4, 2, 0) X(5, 1, 2) X(5, 2, 2)

// gradient direction: the channel loop is a run-time loop (one source patch in registers at a time), so the kernel also
// serves the second layers of the small CNNs (8 -> 16 channels) where an implicit GEMM has too little to chew on
bool fwd_thin_geometry(const ConvGeom& g) {
    return g.Cin <= 16 && g.Cout <= 64 && g.KH == g.KW && g.sh == g.sw && g.ph == g.pw && g.dh == 1 && g.dw == 1 &&
           getenv("CROWN_B200_DISABLE_CONV_THIN") == nullptr;
}

bool thin_geometry(const ConvGeom& g) {
    return g.Cin <= 4 && g.KH == g.KW && g.sh == g.sw && g.ph == g.pw && g.dh == 1 && g.dw == 1 &&
           getenv("CROWN_B200_DISABLE_CONV_THIN") == nullptr;
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
    if (thin_geometry(g) && CinP == 4) {
#define X(K, S, P) if (g.KH == K && g.sh == S && g.ph == P) { if (launch_bwd_thin<K, S, P>(A_out, Wk, A_in, g, rows, accumulate, done, st)) return true; }
        CB_THIN_CASES(X)
#undef X
    }
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
    if (fwd_thin_geometry(g)) {
#define X(K, S, P) if (g.KH == K && g.sh == S && g.ph == P) { if (launch_fwd_thin<K, S, P>(g_in, Wk, b, g_out, g, CoutP, rows, done, st)) return true; }
        CB_THIN_CASES(X)
#undef X
    }
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
