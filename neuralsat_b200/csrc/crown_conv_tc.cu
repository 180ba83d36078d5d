// Convolutions of the CROWN pass and of its gradient as implicit GEMMs on the 5th-generation tensor cores (sm_100a).
//
//   pass      A_in[r,ci,hi,wi]  (+)= sum_{co,kh,kw} A_out[r,co,ho,wo] W[co,ci,kh,kw],   hi = ho*s - p + kh
//             = F.conv_transpose2d(A_out, W, output_padding = what makes the map Hin x Win)  (operators/convolution.py:66-96)
//             and bias_rows[r] += sum A_out[r,co,:,:] b[co]                                   (convolution.py:93)
//   gradient  g_out[r,co,ho,wo] = b[co] + sum_{ci,kh,kw} g_in[r,ci,hi,wi] W[co,ci,kh,kw]      (the same operator forward)
//
// Formulation.  The conv OUTPUT map (Ho x Wo, "coarse") and the residue classes (hi mod s, wi mod s) of the conv INPUT
// map ("fine", each class a grid of about Hin/s x Win/s) are all laid on ONE zero-padded grid of Hp x Wp positions per
// sub-domain row, rows stacked: q = r*Hp*Wp + y*Wp + x.  A tap (kh,kw) with kh - p = s*qh + ch connects fine class ch at
// class coordinate y + qh with coarse position y, so on that grid a tap is a CONSTANT shift of the linear index:
//   pass      D_class[q, ci] = sum_{taps of the class} sum_co Src[q - (qh*Wp + qw), co] * W_t[co, ci]
//   gradient  D[q, co]       = sum_{taps}              sum_ci Src_class(t)[q + (qh*Wp + qw), ci] * W_t[ci, co]
// The pad columns / rows are shared between neighbouring map rows / sub-domain rows (Wp = W + max|qw|): a read that
// leaves the map lands on a position that is invalid for the source grid and therefore holds zero.
//
// Kernel.  One CTA owns n_mt x 128 consecutive positions q (M tiles of the MMA: TMEM lane = position) and an N tile of
// <= 128 destination channels.  The 16 loader warps read the fp32 source (coalesced along x), split it into the three
// bf16 planes of the bf16x3 scheme (crown_tc.cu) and store it K-major, [plane][channel/8][position][channel%8], into
// shared memory - ONCE per K chunk: a tap's operand is the same buffer with the descriptor start address moved by
// shift*16 bytes (no im2col copy).  Weights [tap][channel/16][plane][2][N][8] arrive by bulk copies (resident when
// they fit, else a ring of (chunk, tap) blocks).  One thread issues six MMAs per (tap, M tile, 16 channels) into a
// main and a small-terms accumulator per accumulator set (TMEM).  The loader warps then turn into the epilogue:
// tcgen05.ld, main + small (+ bias), strided NCHW stores (lanes = consecutive x).
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "crown_kernels.cuh"
#include "crown_tc_common.cuh"

namespace cb {

namespace {

using namespace tcc;

constexpr int CT_LOAD_WARPS = 8;          // loader / epilogue warps: 4 TMEM lane quarters x 2 column groups
constexpr int CT_LOADERS = CT_LOAD_WARPS * 32;
constexpr int CT_THREADS = 64 + CT_LOADERS;      // warp 0: weight producer, warp 1: MMA issuer + TMEM allocation
constexpr int CT_SRC_STAGES = 2;
constexpr int CT_W_STAGES = 8;
constexpr int CT_SMEM_MAX = 227 * 1024 - 1024;
constexpr int CT_SMEM_SHARED = 110 * 1024;       // two CTAs per SM below this

__device__ __forceinline__ void ct_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void ct_tmem_ld8_raw(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

__global__ void __launch_bounds__(CT_THREADS, 2) k_conv_tc(const __grid_constant__ ConvTcArgs a) {
    if (a.done != nullptr && *a.done != 0) return;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t src_full[CT_SRC_STAGES];
    __shared__ __align__(8) uint64_t src_empty[CT_SRC_STAGES];
    __shared__ __align__(8) uint64_t w_full[CT_W_STAGES];
    __shared__ __align__(8) uint64_t w_empty[CT_W_STAGES];
    __shared__ __align__(8) uint64_t acc_full;
    __shared__ uint32_t tmem_base_s;

    const ConvTcGeom& g = a.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.y;
    const int Mcta = g.n_mt * 128;
    const int m0 = (int)blockIdx.x * Mcta;                 // rows * G < 2^31 is checked by the launcher
    const int KG = g.KC >> 3;                              // 8-channel groups per chunk
    const int n_chunks = g.Kp / g.KC;
    const int n_buf = g.dir == 0 ? 1 : g.n_cls;
    const int P = g.P;
    const int plane_bytes = KG * P * 16;                   // one bf16 plane of one buffer
    const int buf_bytes = 3 * plane_bytes;
    const int stage_bytes = n_buf * buf_bytes;
    const int w_kstep = 3 * g.N16 * 32;                    // bytes of one (tap, 16 channels) weight block (three planes)
    const int w_block = (g.KC >> 4) * w_kstep;             // one (chunk, tap)
    const int acc_w = (g.mma3 ? 3 : 2) * g.N16;            // TMEM columns of one (accumulator set, M tile)
    uint8_t* const s_src = smem;
    uint8_t* const s_w = smem + (size_t)g.src_stages * stage_bytes;
    const int w_total = g.n_taps * (g.Kp >> 4) * w_kstep;  // resident form: every tap, every channel of this N tile
    int* const s_off = reinterpret_cast<int*>(s_w + (g.w_resident ? w_total : g.w_stages * w_block));   // [n_buf][P]
    int* const s_prow = s_off + n_buf * P;                 // [P] sub-domain row of a source position (pass only)
    float* const s_bias = reinterpret_cast<float*>(s_prow + P);                                          // [Mcta + 8]

    if (threadIdx.x == 0) {
        for (int s = 0; s < CT_SRC_STAGES; ++s) { mbar_init(&src_full[s], CT_LOAD_WARPS); mbar_init(&src_empty[s], 1); }
        for (int s = 0; s < CT_W_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(&acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(g.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // position tables: element offset of (row, y, x), channel 0, in the source tensor; -1 = holds zero
    const int HWs = g.Hsrc * g.Wsrc;
    for (int i = threadIdx.x; i < n_buf * P; i += CT_THREADS) {
        const int buf = i / P, pl = i - buf * P;
        const int q = m0 + g.dmin + pl;
        int off = -1, row = -1;
        if (q >= 0) {
            const int r = q / g.G;
            if (r < a.rows) {
                const int rem = q - r * g.G;
                const int y = rem / g.Wp, x = rem - y * g.Wp;
                int hv, wv, ys, xs;
                if (g.dir == 0) { hv = g.Hsrc; wv = g.Wsrc; ys = y; xs = x; }
                else { hv = g.cls_h[buf]; wv = g.cls_w[buf]; ys = g.sh * y + g.cls_oh[buf]; xs = g.sw * x + g.cls_ow[buf]; }
                if (y < hv && x < wv) { off = r * g.Csrc * HWs + ys * g.Wsrc + xs; row = r; }
            }
        }
        s_off[i] = off;
        if (buf == 0) s_prow[pl] = row;
    }
    for (int i = threadIdx.x; i < Mcta + 8; i += CT_THREADS) s_bias[i] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== weight producer =====
        if (lane == 0) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wp) + (size_t)n_tile * w_total;
            if (g.w_resident) {
                // the mbarrier transaction count is 20 bits: hand the tensor over in slices
                int off = 0, s = 0;
                while (off < w_total) {
                    const int n = min(w_total - off, 64 * 1024);
                    mbar_expect_tx(&w_full[s], (uint32_t)n);
                    bulk_g2s(s_w + off, wsrc + off, (uint32_t)n, &w_full[s]);
                    off += n;
                    ++s;
                }
            } else {
                const int n_blocks = n_chunks * g.n_taps;
                for (int b = 0; b < n_blocks; ++b) {
                    const int s = b % g.w_stages;
                    if (b >= g.w_stages) mbar_wait(&w_empty[s], (uint32_t)((b / g.w_stages) - 1) & 1u);
                    const int chunk = b / g.n_taps, tap = b - chunk * g.n_taps;
                    mbar_expect_tx(&w_full[s], (uint32_t)w_block);
                    bulk_g2s(s_w + (size_t)s * w_block,
                             wsrc + ((size_t)tap * (g.Kp >> 4) + (size_t)chunk * (g.KC >> 4)) * w_kstep, (uint32_t)w_block,
                             &w_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t a_lbo = (uint32_t)P * 16;
            if (g.w_resident) {
                const int n_slices = (w_total + 64 * 1024 - 1) / (64 * 1024);
                for (int s = 0; s < n_slices; ++s) mbar_wait(&w_full[s], 0);
            }
            for (int chunk = 0; chunk < n_chunks; ++chunk) {
                const int cs = chunk % g.src_stages;
                mbar_wait(&src_full[cs], (uint32_t)(chunk / g.src_stages) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t s_stage = smem_u32(s_src + (size_t)cs * stage_bytes);
                for (int tap = 0; tap < g.n_taps; ++tap) {
                    const int b = chunk * g.n_taps + tap;
                    uint32_t wb;
                    if (g.w_resident) {
                        wb = smem_u32(s_w) + (uint32_t)((tap * (g.Kp >> 4) + chunk * (g.KC >> 4)) * w_kstep);
                    } else {
                        const int s = b % g.w_stages;
                        mbar_wait(&w_full[s], (uint32_t)(b / g.w_stages) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        wb = smem_u32(s_w + (size_t)s * w_block);
                    }
                    const ConvTcTap tp = g.taps[tap];
                    const int slot = g.dir == 0 ? g.acc_slot[tp.acc] : 0;
                    const uint32_t sa0 = s_stage + (uint32_t)tp.buf * buf_bytes + (uint32_t)tp.shift * 16;
                    for (int mt = 0; mt < g.n_mt; ++mt) {
                        const uint32_t d0 = tmem_base + (uint32_t)((slot * g.n_mt + mt) * acc_w);
                        for (int ks = 0; ks < (g.KC >> 4); ++ks) {
                            const uint32_t first = (chunk == 0 && tp.first && ks == 0) ? 0u : 1u;
                            const uint32_t sa = sa0 + (uint32_t)(mt * 128) * 16 + (uint32_t)(ks * 2) * a_lbo;
                            const uint32_t sb = wb + (uint32_t)ks * w_kstep;
                            uint64_t ad[3];
#pragma unroll
                            for (int pl = 0; pl < 3; ++pl) ad[pl] = umma_desc(sa + pl * plane_bytes, a_lbo, 128);
                            if (g.mma3) {
                                // the weight planes sit side by side along N: [w1 | w2 | w3] on ONE descriptor, so
                                //   x1.[w1|w2|w3] -> [main | s1 | s2],  x2.[w1|w2] -> [s1 | s2],  x3.[w1] -> [s2]
                                // gives the six products of the bf16x3 split in three MMAs (as crown_chain.cu does)
                                const uint64_t bd = umma_desc(sb, (uint32_t)(3 * g.N16) * 16, 128);
                                umma_bf16(d0, ad[0], bd, umma_idesc_bf16(3 * g.N16), first);
                                umma_bf16(d0 + (uint32_t)g.N16, ad[1], bd, umma_idesc_bf16(2 * g.N16), 1u);
                                umma_bf16(d0 + (uint32_t)(2 * g.N16), ad[2], bd, umma_idesc_bf16(g.N16), 1u);
                            } else {
                                const uint32_t idesc = umma_idesc_bf16(g.N16);
                                const uint32_t b_lbo = (uint32_t)g.N16 * 16;
                                const int b_plane = g.N16 * 32;
                                uint64_t bd[3];
#pragma unroll
                                for (int pl = 0; pl < 3; ++pl) bd[pl] = umma_desc(sb + pl * b_plane, b_lbo, 128);
                                const uint32_t d_small = d0 + (uint32_t)g.N16;
                                // five correction terms into their own accumulator (see crown_tc.cu)
                                umma_bf16(d_small, ad[2], bd[0], idesc, first);
                                umma_bf16(d_small, ad[1], bd[1], idesc, 1u);
                                umma_bf16(d_small, ad[0], bd[2], idesc, 1u);
                                umma_bf16(d_small, ad[1], bd[0], idesc, 1u);
                                umma_bf16(d_small, ad[0], bd[1], idesc, 1u);
                                umma_bf16(d0, ad[0], bd[0], idesc, first);
                            }
                        }
                    }
                    if (!g.w_resident) umma_commit(&w_empty[b % g.w_stages]);
                }
                umma_commit(&src_empty[cs]);
            }
            umma_commit(&acc_full);
        }
    } else {
        // ===== loaders (source -> bf16x3 planes in shared memory), then epilogue =====
        const int te = threadIdx.x - 64;
        const bool do_bias = g.dir == 0 && a.bias != nullptr && a.bias_rows != nullptr && n_tile == 0;
        const int r_first = m0 / g.G;
        int cur_r = -1;
        float cur_acc = 0.f;
        const int per_buf = P * KG;
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int cs = chunk % g.src_stages;
            if (chunk >= g.src_stages) mbar_wait(&src_empty[cs], (uint32_t)((chunk / g.src_stages) - 1) & 1u);
            uint8_t* const stage = s_src + (size_t)cs * stage_bytes;
            for (int buf = 0; buf < n_buf; ++buf) {
                const int* const tab = s_off + buf * P;
                uint8_t* const bstage = stage + (size_t)buf * buf_bytes;
                for (int item = te; item < per_buf; item += CT_LOADERS) {
                    const int g8 = item / P, pl = item - g8 * P;
                    const int off = tab[pl];
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.f;
                    if (off >= 0) {
                        const int c0 = chunk * g.KC + g8 * 8;
                        const float* sp = a.src + (size_t)off + (size_t)c0 * HWs;
                        if (c0 + 8 <= g.Csrc) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = __ldg(sp + (size_t)i * HWs);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (c0 + i < g.Csrc) v[i] = __ldg(sp + (size_t)i * HWs);
                        }
                        if (do_bias && pl >= -g.dmin && pl < -g.dmin + Mcta) {
                            float t = 0.f;
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (c0 + i < g.Csrc) t = fmaf(v[i], __ldg(a.bias + c0 + i), t);
                            const int r = s_prow[pl];
                            if (r != cur_r) {
                                if (cur_r >= 0) atomicAdd(s_bias + (cur_r - r_first), cur_acc);
                                cur_r = r;
                                cur_acc = 0.f;
                            }
                            cur_acc += t;
                        }
                    }
                    uint4 p1, p2, p3;
                    pack8(v, p1, p2, p3);
                    uint8_t* d = bstage + (size_t)item * 16;       // [kgroup][position]: item = g8 * P + pl
                    *reinterpret_cast<uint4*>(d) = p1;
                    *reinterpret_cast<uint4*>(d + plane_bytes) = p2;
                    *reinterpret_cast<uint4*>(d + 2 * plane_bytes) = p3;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) ct_mbar_arrive(&src_full[cs]);
        }
        if (do_bias) {
            if (cur_r >= 0) atomicAdd(s_bias + (cur_r - r_first), cur_acc);
            asm volatile("bar.sync 1, %0;" ::"n"(CT_LOADERS) : "memory");
            for (int i = te; i < Mcta + 8; i += CT_LOADERS) {
                const float t = s_bias[i];
                if (t != 0.f && r_first + i < a.rows) atomicAdd(a.bias_rows + r_first + i, t);
            }
        }

        // ---- epilogue: TMEM lane quarter = warp % 4, column group = (warp - 2) / 4 ----
        const int quarter = warp & 3;
        const int cgrp = (warp - 2) >> 2;
        const int HWd = g.Hdst * g.Wdst;
        mbar_wait(&acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int n_acc = g.dir == 0 ? g.n_cls : 1;
        for (int ai = 0; ai < n_acc; ++ai) {
            const int slot = g.dir == 0 ? g.acc_slot[ai] : 0;
            if (slot < 0 && a.accumulate) continue;          // nothing reaches this class
            for (int mt = 0; mt < g.n_mt; ++mt) {
                const int q = m0 + mt * 128 + quarter * 32 + lane;
                const int r = q / g.G;
                const int rem = q - r * g.G;
                const int y = rem / g.Wp, x = rem - y * g.Wp;
                int hv, wv, ys, xs;
                if (g.dir == 0) { hv = g.cls_h[ai]; wv = g.cls_w[ai]; ys = g.sh * y + g.cls_oh[ai]; xs = g.sw * x + g.cls_ow[ai]; }
                else { hv = g.Hdst; wv = g.Wdst; ys = y; xs = x; }
                const bool valid = r < a.rows && y < hv && x < wv;
                float* const dp = a.dst + (size_t)r * g.Cdst * HWd + (size_t)ys * g.Wdst + xs;
                const uint32_t tcol = trow + (uint32_t)((max(slot, 0) * g.n_mt + mt) * acc_w);
                for (int c0 = cgrp * 8; c0 < g.N16; c0 += 8 * (CT_LOAD_WARPS / 4)) {
                    float d[8];
                    if (slot >= 0) {
                        uint32_t r0[8], r1[8], r2[8];
                        ct_tmem_ld8_raw(tcol + (uint32_t)c0, r0);
                        ct_tmem_ld8_raw(tcol + (uint32_t)(g.N16 + c0), r1);
                        if (g.mma3) ct_tmem_ld8_raw(tcol + (uint32_t)(2 * g.N16 + c0), r2);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float sm = __uint_as_float(r1[i]);
                            if (g.mma3) sm += __uint_as_float(r2[i]);
                            d[i] = __uint_as_float(r0[i]) + sm;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) d[i] = 0.f;
                    }
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int c = n_tile * g.N16 + c0 + i;
                            if (c < g.Cdst) {
                                float val = d[i];
                                if (g.dir == 1 && a.bias != nullptr) val += __ldg(a.bias + c);
                                float* p = dp + (size_t)c * HWd;
                                if (a.accumulate) val += *p;
                                *p = val;
                            }
                        }
                    }
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols) : "memory");
    }
}

// W [Cout,Cin,T] -> [n_tile][tap][Kp/16][plane][2][N16][8]; pass: (n, k) = (ci, co), gradient: (n, k) = (co, ci)
// (mma3: [n_tile][tap][Kp/16][2][3*N16][8], the planes side by side along N)
__global__ void k_conv_tc_pack_w(const float* __restrict__ W, int Cout, int Cin, int T, int dir, int Kp, int N16,
                                 int n_ntiles, int mma3, uint16_t* __restrict__ out) {
    const int Ns = dir == 0 ? Cin : Cout, Ks = dir == 0 ? Cout : Cin;
    const size_t total = (size_t)n_ntiles * T * (Kp >> 3) * N16;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int nl = (int)(i % N16);
        const int kg = (int)((i / N16) % (Kp >> 3));
        const int t = (int)((i / ((size_t)N16 * (Kp >> 3))) % T);
        const int nt = (int)(i / ((size_t)N16 * (Kp >> 3) * T));
        const int n = nt * N16 + nl;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kg * 8 + j;
            float w = 0.f;
            if (n < Ns && k < Ks) {
                const int co = dir == 0 ? k : n, ci = dir == 0 ? n : k;
                w = W[((size_t)co * Cin + ci) * T + t];
            }
            v[j] = w;
        }
        uint4 p1, p2, p3;
        pack8(v, p1, p2, p3);
        const size_t kstep = ((size_t)(nt * T + t) * (Kp >> 4) + (kg >> 1)) * ((size_t)3 * N16 * 16);
        if (mma3) {
            const size_t base = kstep + (size_t)(kg & 1) * 3 * N16 * 8 + (size_t)nl * 8;
            *reinterpret_cast<uint4*>(out + base) = p1;
            *reinterpret_cast<uint4*>(out + base + (size_t)N16 * 8) = p2;
            *reinterpret_cast<uint4*>(out + base + (size_t)2 * N16 * 8) = p3;
        } else {
            const size_t base = kstep + (size_t)(kg & 1) * N16 * 8 + (size_t)nl * 8;
            *reinterpret_cast<uint4*>(out + base) = p1;
            *reinterpret_cast<uint4*>(out + base + (size_t)2 * N16 * 8) = p2;
            *reinterpret_cast<uint4*>(out + base + (size_t)4 * N16 * 8) = p3;
        }
    }
}

int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// Launch configuration for `rows` sub-domain rows: M tiles per CTA, channel chunk, stage counts, shared memory.
// Layers whose weights stay resident are latency-bound (load -> MMA -> epilogue inside one CTA): they are sized so that
// two CTAs share an SM and cover each other's phases.  Layers that stream their weights are MMA-bound: they take the
// largest M per CTA instead (every weight block is then reused by more positions).
bool conv_tc_try(ConvTcGeom& g, int n_mt, bool shared_sm) {
    const int n_slots = g.dir == 0 ? g.n_slots : 1;
    const int n_buf = g.dir == 0 ? 1 : g.n_cls;
    const int acc_w = (g.mma3 ? 3 : 2) * g.N16;
    const int Mcta = n_mt * 128;
    const int P = (Mcta + g.span + 7) & ~7;
    const int w_kstep = 3 * g.N16 * 32;
    const int w_total = g.n_taps * (g.Kp >> 4) * w_kstep;
    const int cols = n_slots * n_mt * acc_w;
    int tm = 32;
    while (tm < cols) tm <<= 1;
    if (tm > 512 || (shared_sm && tm > 256)) return false;
    const long long cap = shared_sm ? CT_SMEM_SHARED : CT_SMEM_MAX;
    for (int KC = 64; KC >= 16; KC >>= 1) {
        if (g.Kp % KC != 0) continue;
        const int n_chunks = g.Kp / KC;
        const int stages = n_chunks > 1 ? CT_SRC_STAGES : 1;
        const long long src_bytes = (long long)stages * n_buf * 3 * (KC >> 3) * P * 16;
        const long long extra = (long long)(n_buf + 1) * P * 4 + (long long)(Mcta + 8) * 4 + 1024;
        const long long left = cap - src_bytes - extra;
        if (left <= 0) continue;
        const int w_block = (KC >> 4) * w_kstep;
        int resident = 0, w_stages = 0;
        if (w_total <= left && w_total <= CT_W_STAGES * 64 * 1024) resident = 1;
        else {
            if (shared_sm) continue;
            w_stages = (int)(left / w_block);
            if (w_stages > CT_W_STAGES) w_stages = CT_W_STAGES;
            if (w_stages < 2) continue;
        }
        g.KC = KC; g.n_mt = n_mt; g.P = P; g.src_stages = stages; g.w_stages = w_stages; g.w_resident = resident;
        g.tmem_cols = tm;
        g.smem_bytes = (int)(src_bytes + (resident ? w_total : (long long)w_stages * w_block) + extra);
        return true;
    }
    return false;
}

bool conv_tc_configure(ConvTcGeom& g, int rows) {
    const int n_slots = g.dir == 0 ? g.n_slots : 1;
    const int acc_w = (g.mma3 ? 3 : 2) * g.N16;
    int n_max = 512 / (n_slots * acc_w);
    if (n_max > 4) n_max = 4;
    if (n_max < 1) return false;
    const long long total = (long long)rows * g.G;
    while (n_max > 1 && (total + n_max * 128 - 1) / (n_max * 128) < 2 * 148) --n_max;      // keep every SM busy on small batches
    for (int n_mt = n_max; n_mt >= 1; --n_mt)
        if (conv_tc_try(g, n_mt, true)) return true;
    for (int n_mt = n_max; n_mt >= 1; --n_mt)
        if (conv_tc_try(g, n_mt, false)) return true;
    return false;
}

}  // namespace

bool conv_tc_setup(const ConvGeom& c, int dir, ConvTcGeom& g) {
    memset(&g, 0, sizeof(g));
    if (c.dh != 1 || c.dw != 1) return false;
    if (c.sh < 1 || c.sw < 1 || c.sh * c.sw > CT_MAX_CLS) return false;
    if (c.KH * c.KW > CT_MAX_TAPS) return false;
    g.dir = dir;
    g.sh = c.sh; g.sw = c.sw;
    g.n_cls = c.sh * c.sw;
    const int Hg = (c.Hin + c.sh - 1) / c.sh, Wg = (c.Win + c.sw - 1) / c.sw;     // class grids of the fine map
    int hh = 0, hw = 0;                            // largest |qh|, |qw|: the shared pad rows / columns
    for (int kh = 0; kh < c.KH; ++kh) { const int q = abs(floor_div(kh - c.ph, c.sh)); if (q > hh) hh = q; }
    for (int kw = 0; kw < c.KW; ++kw) { const int q = abs(floor_div(kw - c.pw, c.sw)); if (q > hw) hw = q; }
    g.Hp = (c.Hout > Hg ? c.Hout : Hg) + hh;
    g.Wp = (c.Wout > Wg ? c.Wout : Wg) + hw;
    g.G = g.Hp * g.Wp;
    if (dir == 0) { g.Csrc = c.Cout; g.Cdst = c.Cin; g.Hsrc = c.Hout; g.Wsrc = c.Wout; g.Hdst = c.Hin; g.Wdst = c.Win; }
    else { g.Csrc = c.Cin; g.Cdst = c.Cout; g.Hsrc = c.Hin; g.Wsrc = c.Win; g.Hdst = c.Hout; g.Wdst = c.Wout; }
    g.Kp = (g.Csrc + 15) / 16 * 16;
    for (int cls = 0; cls < g.n_cls; ++cls) {
        const int ch = cls / c.sw, cw = cls - ch * c.sw;
        g.cls_oh[cls] = ch; g.cls_ow[cls] = cw;
        g.cls_h[cls] = ch < c.Hin ? (c.Hin - ch + c.sh - 1) / c.sh : 0;
        g.cls_w[cls] = cw < c.Win ? (c.Win - cw + c.sw - 1) / c.sw : 0;
        g.acc_slot[cls] = -1;
    }
    // taps: kh - p = sh*qh + ch
    int dmin = 0, dmax = 0;
    int deltas[CT_MAX_TAPS];
    g.n_taps = c.KH * c.KW;
    for (int kh = 0; kh < c.KH; ++kh)
        for (int kw = 0; kw < c.KW; ++kw) {
            const int t = kh * c.KW + kw;
            const int qh = floor_div(kh - c.ph, c.sh), qw = floor_div(kw - c.pw, c.sw);
            const int ch = kh - c.ph - qh * c.sh, cw = kw - c.pw - qw * c.sw;
            const int cls = ch * c.sw + cw;
            const int d = (qh * g.Wp + qw) * (dir == 0 ? -1 : 1);
            deltas[t] = d;
            if (d < dmin) dmin = d;
            if (d > dmax) dmax = d;
            g.taps[t].acc = (short)(dir == 0 ? cls : 0);
            g.taps[t].buf = (short)(dir == 0 ? 0 : cls);
        }
    g.dmin = dmin;
    g.span = dmax - dmin;
    bool seen[CT_MAX_CLS] = {false, false, false, false};
    g.n_slots = 0;
    for (int t = 0; t < g.n_taps; ++t) {
        g.taps[t].shift = deltas[t] - dmin;
        const int acc = g.taps[t].acc;
        g.taps[t].first = seen[acc] ? 0 : 1;
        if (!seen[acc]) {
            seen[acc] = true;
            if (dir == 0) g.acc_slot[acc] = g.n_slots++;
        }
    }
    if (dir == 1) g.n_slots = 1;
    // N tile: TMEM holds n_slots x n_mt x (main + small) accumulators of N16 columns
    int N16 = (g.Cdst + 15) / 16 * 16;
    if (N16 > 128) N16 = 128;
    while (g.n_slots * 2 * N16 > 512 && N16 > 16) N16 >>= 1;
    if (g.n_slots * 2 * N16 > 512) return false;
    // N16 must keep 16-byte aligned planes and a valid MMA shape (multiple of 16)
    N16 = (N16 + 15) / 16 * 16;
    g.N16 = N16;
    g.mma3 = 3 * N16 <= 256 ? 1 : 0;
    if (g.mma3 && g.n_slots * 3 * N16 > 512) g.mma3 = 0;
    g.n_ntiles = (g.Cdst + N16 - 1) / N16;
    ConvTcGeom probe = g;
    return conv_tc_configure(probe, 148 * 4);
}

size_t conv_tc_w_elems(const ConvTcGeom& g) {
    return (size_t)g.n_ntiles * g.n_taps * (g.Kp >> 4) * 3 * g.N16 * 16;
}

void conv_tc_pack_weight(const float* W, const ConvTcGeom& g, int Cout, int Cin, uint16_t* out, cudaStream_t st) {
    const size_t total = (size_t)g.n_ntiles * g.n_taps * (g.Kp >> 3) * g.N16;
    const unsigned blocks = (unsigned)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    k_conv_tc_pack_w<<<blocks, 256, 0, st>>>(W, Cout, Cin, g.n_taps, g.dir, g.Kp, g.N16, g.n_ntiles, g.mma3, out);
}

cudaError_t conv_tc(const ConvTcGeom& g_in, const float* src, float* dst, const uint16_t* wp, const float* bias,
                    float* bias_rows, int rows, bool accumulate, const int* done, cudaStream_t st) {
    ConvTcArgs a;
    a.g = g_in;
    if (!conv_tc_configure(a.g, rows)) return cudaErrorInvalidConfiguration;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_MAX);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    a.src = src; a.dst = dst; a.wp = wp; a.bias = bias; a.bias_rows = bias_rows;
    a.rows = rows; a.accumulate = accumulate ? 1 : 0; a.done = done;
    Launch _l(g_in.dir == 0 ? K_CONV_TC_BWD : K_CONV_TC_FWD, st);
    const long long total = (long long)rows * a.g.G;
    const int Mcta = a.g.n_mt * 128;
    if (total + Mcta + a.g.span >= (1ll << 31)) return cudaErrorInvalidValue;
    if ((long long)rows * a.g.Csrc * a.g.Hsrc * a.g.Wsrc >= (1ll << 31)) return cudaErrorInvalidValue;      // 32-bit position tables
    dim3 grid((unsigned)((total + Mcta - 1) / Mcta), (unsigned)a.g.n_ntiles);
    k_conv_tc<<<grid, CT_THREADS, a.g.smem_bytes, st>>>(a);
    return cudaGetLastError();
}

}  // namespace cb
