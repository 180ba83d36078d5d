// Whole-network CROWN pass for fully-connected ReLU chains in ONE kernel (sm_100a, tcgen05 + TMA).
//
// A CTA owns 64 sub-domain rows (row r = s*Bd + b) and walks the network backwards,
//   A_{k-1}' = A_k . W_k            (tensor cores; operators/linear.py:167-175)
//   A_{k-1}  = relax(A_{k-1}')      (epilogue; operators/relu.py:456-494, clampmult.py:17-43, beta_crown.py:163-204)
// without the coefficient matrix ever leaving the SM: the contraction is issued TRANSPOSED,
//   D^T[neurons(128 per MMA) x rows(64)] = W_k^T[neurons x K] . A_k^T[K x rows],
// so that (i) the weights are the streamed M-side operand (bulk TMA from L2, 8-stage ring), (ii) the
// sub-domain tile is the N-side operand, small enough (64 rows x 256 k x 3 bf16 planes = 96 KB) to
// stay resident in shared memory, where the epilogue of one layer writes it in UMMA layout for the
// MMAs of the next, and (iii) a TMEM lane is a NEURON: the 32 lanes of an epilogue warp read 32
// consecutive neurons of one sub-domain row of l / u / alpha / x_L / x_U and write lA the same way,
// i.e. every global access of the epilogue is a coalesced 128-byte line with no staging.
// TMEM holds two layers (2 x [2 M-tiles x (main + small-terms accumulator) x 64 columns] = 512
// columns), so the MMAs of layer k-1 run under the epilogue of layer k, K-chunk by K-chunk.
//
// fp32 fidelity: the bf16x3 split with separate accumulators of crown_tc.cu (see its header).
//
// Operand layouts (bf16):
//   W_k^T packed by tc_pack_weight(TR = 128): [m / 128][k / 16][plane][(k / 8) % 2][m % 128][k % 8]
//          one (M-tile, k-step) = 12 KB contiguous = one bulk copy; LBO = 2048 B, SBO = 128 B
//   X (shared memory only): block (plane, k / 8) at (plane * 32 + k / 8) * 1040 B holds [row 0..63][k % 8];
//          LBO = 1040 B (the 16 B pad makes the epilogue's 2-byte stores bank-conflict free), SBO = 128 B
#include "crown_kernels.cuh"
#include "crown_tc_common.cuh"

namespace cb {

namespace {

using namespace tcc;

constexpr int CH_TR = 64;                       // sub-domain rows per CTA = MMA N
constexpr int CH_WSTAGES = 8;
constexpr int CH_WSTAGE = 3 * 128 * 16 * 2;     // 12288 B: three planes of a [128 x 16] weight tile
constexpr int CH_WPLANE = 128 * 16 * 2;
constexpr int CH_XCG = 64 * 16 + 16;            // 1040 B per (plane, 8 k-values) block
constexpr int CH_XBYTES = 3 * (CHAIN_KMAX / 8) * CH_XCG;
constexpr int CH_EPI_WARPS = 8;
constexpr int CH_EPI_THREADS = CH_EPI_WARPS * 32;
constexpr int CH_THREADS = 64 + CH_EPI_THREADS;
constexpr int CH_SMEM = CH_WSTAGES * CH_WSTAGE + CH_XBYTES + CH_TR * 8 * 4 + 2 * CH_TR * CHAIN_JMAX * 4 +
                        8 * CH_TR * 4 + CH_TR * 4;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void epi_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI_THREADS) : "memory");
}

// main + small-terms accumulator of 8 columns, one wait
__device__ __forceinline__ void tmem_ld8x2(uint32_t t_main, uint32_t t_small, float (&d)[8]) {
    uint32_t a[8], b[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%17];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
          "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7])
        : "r"(t_main), "r"(t_small)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = __uint_as_float(a[i]) + __uint_as_float(b[i]);
}

__device__ __forceinline__ float warp_sum32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y -> three bf16 planes at (row, k) of the resident X operand
__device__ __forceinline__ void x_store(uint8_t* X, int k, int row, float y) {
    const __nv_bfloat16 h1 = __float2bfloat16_rn(y);
    const float r1 = y - __bfloat162float(h1);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 h3 = __float2bfloat16_rn(r1 - __bfloat162float(h2));
    uint8_t* p = X + (size_t)(k >> 3) * CH_XCG + row * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(p) = h1;
    *reinterpret_cast<__nv_bfloat16*>(p + 32 * CH_XCG) = h2;
    *reinterpret_cast<__nv_bfloat16*>(p + 64 * CH_XCG) = h3;
}

__global__ void __launch_bounds__(CH_THREADS, 1) k_chain_pass(const __grid_constant__ ChainArgs a) {
    if (a.done != nullptr && *a.done != 0) return;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t w_full[CH_WSTAGES];
    __shared__ __align__(8) uint64_t w_empty[CH_WSTAGES];
    __shared__ __align__(8) uint64_t x_full[2];
    __shared__ __align__(8) uint64_t acc_full[2];
    __shared__ __align__(8) uint64_t acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    uint8_t* const wring = smem;
    uint8_t* const X = smem + CH_WSTAGES * CH_WSTAGE;
    uint32_t* const s_bmask = reinterpret_cast<uint32_t*>(X + CH_XBYTES);             // [64][8]
    int* const s_bloc = reinterpret_cast<int*>(s_bmask + CH_TR * 8);                   // [64][JMAX]
    float* const s_bvs = reinterpret_cast<float*>(s_bloc + CH_TR * CHAIN_JMAX);        // [64][JMAX]
    float* const s_part = s_bvs + CH_TR * CHAIN_JMAX;                                  // [2][4][64]
    float* const s_extra = s_part + 8 * CH_TR;                                         // [64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * CH_TR;

    if (threadIdx.x == 0) {
        for (int s = 0; s < CH_WSTAGES; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&x_full[i], CH_EPI_WARPS);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], CH_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== weight producer: one (k-step, M-tile) block per ring slot, in MMA order =====
        if (lane == 0) {
            uint32_t wst = 0;
            for (int j = 0; j < a.n_steps; ++j) {
                const ChainStep& st = a.step[j];
                const int nks = st.Kp >> 4;
                const int n_mt = (st.M + 127) >> 7;
                for (int mt0 = 0; mt0 < n_mt; mt0 += 2) {
                    const int nmt = min(2, n_mt - mt0);
                    for (int ks = 0; ks < nks; ++ks)
                        for (int mi = 0; mi < nmt; ++mi, ++wst) {
                            const int s = wst % CH_WSTAGES;
                            mbar_wait(&w_empty[s], ((wst / CH_WSTAGES) & 1u) ^ 1u);
                            mbar_expect_tx(&w_full[s], CH_WSTAGE);
                            bulk_g2s(wring + (size_t)s * CH_WSTAGE,
                                     st.wp + ((size_t)(mt0 + mi) * nks + ks) * (CH_WSTAGE / 2), CH_WSTAGE, &w_full[s]);
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(CH_TR);
            uint32_t wst = 0, jc = 0, xph0 = 0, xph1 = 0;
            const uint32_t xbase = smem_u32(X);
            for (int j = 0; j < a.n_steps; ++j) {
                const ChainStep& st = a.step[j];
                const int nks = st.Kp >> 4;
                const int n_mt = (st.M + 127) >> 7;
                for (int mt0 = 0; mt0 < n_mt; mt0 += 2, ++jc) {
                    const int nmt = min(2, n_mt - mt0);
                    const uint32_t p = jc & 1u;
                    mbar_wait(&acc_empty[p], ((jc >> 1) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int ks = 0; ks < nks; ++ks) {
                        if (mt0 == 0 && (ks & 7) == 0) {       // X chunk ks/8 of this step: written by the epilogue above
                            if (ks == 0) { mbar_wait(&x_full[0], xph0); xph0 ^= 1u; }
                            else { mbar_wait(&x_full[1], xph1); xph1 ^= 1u; }
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        uint64_t bd[3];
#pragma unroll
                        for (int pl = 0; pl < 3; ++pl)
                            bd[pl] = umma_desc(xbase + (uint32_t)(pl * 32 + ks * 2) * CH_XCG, CH_XCG, 128);
                        for (int mi = 0; mi < nmt; ++mi, ++wst) {
                            const int s = wst % CH_WSTAGES;
                            mbar_wait(&w_full[s], (wst / CH_WSTAGES) & 1u);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint32_t sa = smem_u32(wring + (size_t)s * CH_WSTAGE);
                            uint64_t ad[3];
#pragma unroll
                            for (int pl = 0; pl < 3; ++pl) ad[pl] = umma_desc(sa + pl * CH_WPLANE, 2048, 128);
                            const uint32_t d_main = tmem_base + p * 256 + mi * 128;
                            const uint32_t d_small = d_main + 64;
                            const uint32_t acc = ks ? 1u : 0u;
                            umma_bf16(d_small, ad[2], bd[0], idesc, acc);
                            umma_bf16(d_small, ad[1], bd[1], idesc, 1u);
                            umma_bf16(d_small, ad[0], bd[2], idesc, 1u);
                            umma_bf16(d_small, ad[1], bd[0], idesc, 1u);
                            umma_bf16(d_small, ad[0], bd[1], idesc, 1u);
                            umma_bf16(d_main, ad[0], bd[0], idesc, acc);
                            umma_commit(&w_empty[s]);
                        }
                    }
                    umma_commit(&acc_full[p]);
                }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane = neuron, column = sub-domain row =====
        const int te = threadIdx.x - 64;
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int h = (warp - 2) >> 2;           // column half: rows h*32 .. h*32+31 of the tile
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const int Bd = a.Bd, S = a.S, rows = a.rows;

        // ---- 0. zero the bias slots, pack C into X (the operand of the output layer) ----
        for (int i = te; i < 8 * CH_TR; i += CH_EPI_THREADS) s_part[i] = 0.f;
        {
            const int row = te & 63, kg0 = te >> 6;
            const int r = row0 + row;
            const bool vr = r < rows;
            const int b = vr ? r % Bd : 0, s = vr ? r / Bd : 0;
            const float* crow = a.C + ((size_t)b * S + s) * a.n_out;
            const int Kp0 = a.step[0].Kp;
            for (int kg = kg0; kg < (Kp0 >> 3); kg += 4) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (vr && kg * 8 + i < a.n_out) ? __ldg(crow + kg * 8 + i) : 0.f;
                uint4 p1, p2, p3;
                pack8(v, p1, p2, p3);
                uint8_t* dst = X + (size_t)kg * CH_XCG + row * 16;
                *reinterpret_cast<uint4*>(dst) = p1;
                *reinterpret_cast<uint4*>(dst + 32 * CH_XCG) = p2;
                *reinterpret_cast<uint4*>(dst + 64 * CH_XCG) = p3;
            }
            if (te < CH_TR) {
                float t = 0.f;
                if (vr && a.b_out)
                    for (int k = 0; k < a.n_out; ++k) t = fmaf(__ldg(crow + k), __ldg(a.b_out + k), t);
                s_extra[te] = t;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&x_full[0]);
                if (Kp0 > 128) mbar_arrive(&x_full[1]);
            }
        }
        epi_sync();                                      // bias slots are zeroed before anybody adds to them

        uint32_t jc = 0;
        for (int j = 0; j < a.n_steps; ++j) {
            const ChainStep& st = a.step[j];
            const int M = st.M;
            const int n_mt = (M + 127) >> 7;
            const bool last = (j == a.n_steps - 1);          // the concretize step
            const bool has_alpha = st.alpha != nullptr;
            const int J = last ? 0 : st.J;
            if (!last) {
                // ---- beta records of the pre-activation node, per row (beta_crown.py:163-204) ----
                epi_sync();                                  // everybody is done with the previous lists
                for (int i = te; i < CH_TR * 8; i += CH_EPI_THREADS) s_bmask[i] = 0u;
                epi_sync();
                if (J > 0) {
                    const int row = te >> 2;
                    const int r = row0 + row;
                    if (r < rows) {
                        const size_t jb = (size_t)(r % Bd) * J;
                        for (int jj = te & 3; jj < J; jj += 4) {
                            const float vs = __ldg(st.beta_val + jb + jj) * __ldg(st.beta_sign + jb + jj);
                            const int lc = (int)__ldg(st.beta_loc + jb + jj);
                            const bool on = vs != 0.f && lc >= 0 && lc < M;
                            s_bloc[row * CHAIN_JMAX + jj] = on ? lc : -1;
                            s_bvs[row * CHAIN_JMAX + jj] = vs;
                            if (on) atomicOr(&s_bmask[row * 8 + (lc >> 5)], 1u << (lc & 31));
                        }
                    }
                    epi_sync();
                    if (te < CH_TR && st.beta_bias != nullptr && row0 + te < rows) {
                        const size_t jb = (size_t)((row0 + te) % Bd) * J;
                        float t = s_extra[te];
                        for (int jj = 0; jj < J; ++jj) t = fmaf(s_bvs[te * CHAIN_JMAX + jj], __ldg(st.beta_bias + jb + jj), t);
                        s_extra[te] = t;
                    }
                }
            }
            for (int mt0 = 0; mt0 < n_mt; mt0 += 2, ++jc) {
                const int nmt = min(2, n_mt - mt0);
                const uint32_t p = jc & 1u;
                mbar_wait(&acc_full[p], (jc >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int mi = 0; mi < nmt; ++mi) {
                    const int mt = mt0 + mi;
                    const int m = mt * 128 + q * 32 + lane;
                    const bool vm = m < M;
                    const uint32_t tcol = trow + p * 256 + mi * 128;
                    float* const slot = s_part + ((mt & 1) * 4 + q) * CH_TR;
                    if (!last) {
                        const float bbelow = (vm && st.bias_below) ? __ldg(st.bias_below + m) : 0.f;
                        int apos = -1;
                        if (has_alpha && vm) apos = st.alpha_pos ? __ldg(st.alpha_pos + m) : m;
                        const bool wx = m < ((M + 15) & ~15);          // K range of the next layer (zero padded)
#pragma unroll 1
                        for (int cc = 0; cc < 32; cc += 8) {
                            const int c0 = h * 32 + cc;
                            float d[8], l[8], u[8], al[8], part[8];
                            tmem_ld8x2(tcol + c0, tcol + 64 + c0, d);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int r = row0 + c0 + i;
                                l[i] = 0.f; u[i] = 0.f; al[i] = 0.f;
                                if (vm && r < rows) {
                                    const int b = (S == 1) ? r : r % Bd;
                                    l[i] = __ldg(st.lower + (size_t)b * M + m);
                                    u[i] = __ldg(st.upper + (size_t)b * M + m);
                                    if (apos >= 0) {
                                        const size_t arow = (a.S1 == 1) ? (size_t)b : (size_t)r;
                                        al[i] = __ldg(st.alpha + arow * st.n_alpha + apos);
                                    }
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int r = row0 + c0 + i;
                                const bool ok = vm && r < rows;
                                float y = 0.f;
                                part[i] = 0.f;
                                if (ok) {
                                    if (st.lA) st.lA[(size_t)r * M + m] = d[i];
                                    const Relax8 rx = relax1(l[i], u[i], has_alpha, al[i]);
                                    const float a_pos = fmaxf(d[i], 0.f), a_neg = fminf(d[i], 0.f);
                                    y = rx.d_l * a_pos + rx.d_u * a_neg;
                                    float acc = a_neg * rx.b_u;
                                    if (J > 0 && ((s_bmask[(c0 + i) * 8 + (m >> 5)] >> (m & 31)) & 1u)) {
                                        for (int jj = 0; jj < J; ++jj)
                                            if (s_bloc[(c0 + i) * CHAIN_JMAX + jj] == m) y -= s_bvs[(c0 + i) * CHAIN_JMAX + jj];
                                    }
                                    acc = fmaf(y, bbelow, acc);
                                    part[i] = acc;
                                }
                                if (wx) x_store(X, m, c0 + i, y);
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float t = warp_sum32(part[i]);
                                if (lane == 0) slot[c0 + i] += t;
                            }
                        }
                        // chunk mt of the next layer's operand is complete
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&x_full[mt]);
                    } else {
                        // ---- concretise against the input box (perturbations.py:154-183) ----
                        const int w32 = (M + 31) >> 5;
#pragma unroll 1
                        for (int cc = 0; cc < 32; cc += 8) {
                            const int c0 = h * 32 + cc;
                            float d[8], xl[8], xu[8], part[8];
                            tmem_ld8x2(tcol + c0, tcol + 64 + c0, d);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int r = row0 + c0 + i;
                                xl[i] = 0.f; xu[i] = 0.f;
                                if (vm && r < rows) {
                                    const int b = (S == 1) ? r : r % Bd;
                                    xl[i] = __ldg(a.x_L + (size_t)b * M + m);
                                    xu[i] = __ldg(a.x_U + (size_t)b * M + m);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int r = row0 + c0 + i;
                                const bool ok = vm && r < rows;
                                const float av = ok ? d[i] : 0.f;
                                const float cen = (xu[i] + xl[i]) / 2.0f, dif = (xu[i] - xl[i]) / 2.0f;
                                part[i] = av * cen - fabsf(av) * dif;
                                const unsigned pm = __ballot_sync(0xffffffffu, av > 0.f);
                                const unsigned nm = __ballot_sync(0xffffffffu, av < 0.f);
                                if (a.sign_pos && lane == 0 && r < rows && (m >> 5) < w32) {
                                    a.sign_pos[(size_t)r * w32 + (m >> 5)] = pm;
                                    a.sign_neg[(size_t)r * w32 + (m >> 5)] = nm;
                                }
                                if (a.g0_plain && ok) {
                                    const float sg = (av > 0.f) ? 1.f : ((av < 0.f) ? -1.f : 0.f);
                                    a.g0_plain[(size_t)r * M + m] = cen - sg * dif;
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float t = warp_sum32(part[i]);
                                if (lane == 0) slot[c0 + i] += t;
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[p]);
            }
        }
        // ---- lower bounds: fixed summation order over the per-warp slots ----
        epi_sync();
        if (te < CH_TR && row0 + te < rows) {
            const int r = row0 + te;
            float t = s_extra[te];
#pragma unroll
            for (int i = 0; i < 8; ++i) t += s_part[i * CH_TR + te];
            a.lb[(size_t)(r % Bd) * S + r / Bd] = t;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace

size_t chain_smem_bytes() { return CH_SMEM; }

cudaError_t chain_pass(const ChainArgs& a, cudaStream_t st) {
    Launch _l(K_CHAIN_PASS, st);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_chain_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int tiles = (a.rows + CH_TR - 1) / CH_TR;
    k_chain_pass<<<tiles, CH_THREADS, CH_SMEM, st>>>(a);
    return cudaGetLastError();
}

}  // namespace cb
