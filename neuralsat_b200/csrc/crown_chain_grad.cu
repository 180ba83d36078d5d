// Whole-network alpha/beta GRADIENT for fully-connected ReLU chains in ONE kernel (sm_100a, tcgen05 + TMA).
//
// d(sum lb)/dA flows input -> output through the same operators transposed (operators/clampmult.py:49-95 is
// the hand-written backward it reproduces): with g_0 = c - sign(A_0) d (the worst-case input, written by the
// pass kernel),
//   z_k     = g_{k-1} . W_k^T + b_k                      (tensor cores)
//   g_k     = z_k * (A_k >= 0 ? d_l : d_u) + (A_k < 0 ? b_u : 0)      A_k = lA stash of the pass
//   dalpha  = (unstable, alpha inside its clamp, A_k >= 0) ? z_k * A_k : 0
//   dbeta_j = -sign_j * z_k[loc_j] (+ sign_j * bias_j)   (beta_crown.py:163-204 transposed)
// A CTA owns 64 sub-domain rows and walks the layers forwards with the same structure as the pass kernel
// (crown_chain.cu): contraction issued transposed, D^T[neurons x rows] = W_k[neurons x K] . g_{k-1}^T[K x rows],
// weights streamed by bulk TMA through an 4-stage ring, the row tile resident in shared memory in UMMA
// MN-major layout and rewritten by the epilogue of one layer for the MMAs of the next, TMEM lane = neuron so
// that every global access of the epilogue (l, u, alpha, lA in; dalpha out) is a coalesced 128-byte line, two
// layers of accumulators in TMEM.  The first layer's K (= n_in, 784 for MNIST) does not fit the resident
// operand: the epilogue warps stream g_0 through the two 128-k halves of the operand buffer (x_full / x_empty
// barriers) before they have any epilogue work.
//
// Restrictions (the host falls back to the per-layer kernels otherwise): S == 1, layer widths <= CHAIN_KMAX,
// <= CHAIN_JMAX beta records per row and layer.
#include "crown_chain_common.cuh"

namespace cb {

namespace {

using namespace tcc;
using namespace chn;

__global__ void __launch_bounds__(CH_THREADS, CH_CTAS_PER_SM) k_chain_grad(const __grid_constant__ ChainGradArgs a) {
    if (a.done != nullptr && *a.done != 0) return;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t w_full[CH_WSTAGES];
    __shared__ __align__(8) uint64_t w_empty[CH_WSTAGES];
    __shared__ __align__(8) uint64_t x_full[2];
    __shared__ __align__(8) uint64_t x_empty[2];
    __shared__ __align__(8) uint64_t acc_full[2][2];         // [TMEM buffer][M-tile]
    __shared__ __align__(8) uint64_t acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    uint8_t* const wring = smem;
    uint8_t* const X = smem + CH_WSTAGES * CH_WSTAGE;
    uint32_t* const s_bmask = reinterpret_cast<uint32_t*>(X + CH_XBYTES);             // [64][8]
    int* const s_bloc = reinterpret_cast<int*>(s_bmask + CH_TR * 8);                   // [64][JMAX]
    float* const s_bsg = reinterpret_cast<float*>(s_bloc + CH_TR * CHAIN_JMAX);        // [64][JMAX]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * CH_TR;
    const int n_steps = a.n_steps;
    const int Kp0 = a.step[0].Kp;
    const int n_chunks0 = (Kp0 + 127) >> 7;

    if (threadIdx.x == 0) {
        for (int s = 0; s < CH_WSTAGES; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&x_full[i], CH_EPI_WARPS);
            mbar_init(&x_empty[i], 1);
            mbar_init(&acc_full[i][0], 1);
            mbar_init(&acc_full[i][1], 1);
            mbar_init(&acc_empty[i], CH_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(CH_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== weight producer: one (k-step, M-tile) block per ring slot, in MMA order =====
        uint32_t wst = 0;
        for (int j = 0; j < n_steps; ++j) {
            const GradStep& st = a.step[j];
            const int nks = st.Kp >> 4;
            const int n_mt = (st.M + 127) >> 7;
            // MMA order: K chunk (8 k-steps) major, then M-tile, then k-step
            for (int kc = 0; kc < nks; kc += 8)
                for (int mi = 0; mi < n_mt; ++mi)
                    for (int ks = kc; ks < min(kc + 8, nks); ++ks, ++wst) {
                        const int s = wst % CH_WSTAGES;
                        mbar_wait(&w_empty[s], ((wst / CH_WSTAGES) & 1u) ^ 1u);
                        if (elect_one()) {
                            mbar_expect_tx(&w_full[s], CH_WSTAGE);
                            bulk_g2s(wring + (size_t)s * CH_WSTAGE, st.wp + ((size_t)mi * nks + ks) * (CH_WSTAGE / 2), CH_WSTAGE,
                                     &w_full[s]);
                        }
                        __syncwarp();
                    }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma_idesc_bf16(CH_TR) | (1u << 16);      // B (the row tile) is MN-major
        uint32_t wst = 0, xph0 = 0, xph1 = 0;
        const uint64_t a_base = umma_desc(smem_u32(wring), 2048, 128);
        const uint64_t b_base = umma_desc(smem_u32(X), CH_XKG, 128);
        for (int j = 0; j < n_steps; ++j) {
            const int nks = a.step[j].Kp >> 4;
            const int n_mt = (a.step[j].M + 127) >> 7;
            const uint32_t p = (uint32_t)j & 1u;
            mbar_wait(&acc_empty[p], (((uint32_t)j >> 1) & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // K chunk major, then M-tile: M-tile 0 completes (and its epilogue starts) while the last chunk's MMAs of
            // M-tile 1 are still running
            for (int kc = 0; kc < nks; kc += 8) {
                const int slot = (kc >> 3) & 1;            // 128-k half of the operand buffer
                if (slot == 0) { mbar_wait(&x_full[0], xph0); xph0 ^= 1u; }
                else { mbar_wait(&x_full[1], xph1); xph1 ^= 1u; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int ke = min(kc + 8, nks);
                for (int mi = 0; mi < 2; ++mi) {
                    if (mi < n_mt) {
                        for (int ks = kc; ks < ke; ++ks, ++wst) {
                            const uint64_t b0 = b_base + (uint64_t)((slot * 8 + (ks & 7)) * (2 * CH_XKG >> 4));
                            const uint64_t b1 = b0 + (CH_XPLANE >> 4), b2 = b0 + 2 * (CH_XPLANE >> 4);
                            const uint32_t s = wst % CH_WSTAGES;
                            mbar_wait(&w_full[s], (wst / CH_WSTAGES) & 1u);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint64_t a0 = a_base + (uint64_t)(s * (CH_WSTAGE >> 4));
                            const uint64_t a1 = a0 + (CH_WPLANE >> 4), a2 = a0 + 2 * (CH_WPLANE >> 4);
                            const uint32_t d_main = tmem_base + p * CH_TBUF + mi * CH_TMT;
                            const uint32_t d_small = d_main + CH_TR;
                            const uint32_t acc = ks ? 1u : 0u;
                            if (elect_one()) {
                                umma_bf16(d_small, a2, b0, idesc, acc);
                                umma_bf16(d_small, a1, b1, idesc, 1u);
                                umma_bf16(d_small, a0, b2, idesc, 1u);
                                umma_bf16(d_small, a1, b0, idesc, 1u);
                                umma_bf16(d_small, a0, b1, idesc, 1u);
                                umma_bf16(d_main, a0, b0, idesc, acc);
                                umma_commit(&w_empty[s]);
                            }
                            __syncwarp();
                        }
                    }
                    if (ke == nks) {                       // both barriers advance once per step: phase = j / 2
                        if (elect_one()) umma_commit(&acc_full[p][mi]);
                        __syncwarp();
                    }
                }
                if (elect_one()) umma_commit(&x_empty[slot]);       // the MMAs reading this half have been issued
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane = neuron, column = sub-domain row =====
        const int te = threadIdx.x - 64;
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int h = (warp - 2) >> 2;           // rows h*CH_RPW .. (h+1)*CH_RPW-1
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const int rows = a.rows;

        // ---- A. stream g_0 (worst-case input point, [rows, n_in] fp32) into the operand halves ----
        {
            const int k_local = te & 127, rg0 = te >> 7;          // 512 threads: 128 k x 4 row groups, twice
            for (int c = 0; c < n_chunks0; ++c) {
                const int slot = c & 1;
                if (c >= 2) {
                    mbar_wait(&x_empty[slot], (uint32_t)((c - 2) >> 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const int k = c * 128 + k_local;
                if (k < Kp0) {
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int rg = rg0 + (CH_EPI_THREADS / 128) * it;
                        float y[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = row0 + rg * 8 + i;
                            y[i] = (k < a.n_in && r < rows) ? __ldg(a.g0 + (size_t)r * a.n_in + k) : 0.f;
                        }
                        x_store8(X, slot * 128 + k_local, rg * 8, y);
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&x_full[slot]);
            }
        }

        // ---- B. per-layer epilogues ----
        struct Cur { int j, mt, cc; };
        auto advance = [&](Cur& c) {
            c.cc += 8;
            if (c.cc == CH_RPW) {
                c.cc = 0;
                if (++c.mt >= ((a.step[c.j].M + 127) >> 7)) { c.mt = 0; ++c.j; }
            }
        };
        const bool fast = row0 + CH_TR <= rows;
        // operands of an item: v[0..7] = l, v[8..15] = u, v[16..23] = alpha, v[24..31] = lA stash, v[32] = bias
        auto issue = [&](const Cur& c, float (&v)[33]) {
            const GradStep& st = a.step[c.j];
            const int M = st.M;
            const int m = c.mt * 128 + q * 32 + lane;
            const int c0 = h * CH_RPW + c.cc;
#pragma unroll
            for (int i = 0; i < 33; ++i) v[i] = 0.f;
            if (m >= M) return;
            int apos = -1;
            if (st.alpha != nullptr) apos = st.alpha_pos ? __ldg(st.alpha_pos + m) : m;
            if (st.bias) v[32] = __ldg(st.bias + m);
            const size_t o = (size_t)(row0 + c0) * M + m;
            const float* qa = (apos >= 0) ? st.alpha + (size_t)(row0 + c0) * st.n_alpha + apos : nullptr;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (!fast && row0 + c0 + i >= rows) continue;
                v[i] = __ldg(st.lower + o + (size_t)i * M);
                v[8 + i] = __ldg(st.upper + o + (size_t)i * M);
                if (qa) v[16 + i] = __ldg(qa + (size_t)i * st.n_alpha);
                v[24 + i] = __ldg(st.a_post + o + (size_t)i * M);
            }
        };

        auto process = [&](const Cur& cur, const float (&pre)[33], float (&nx)[33], Cur& nxt) {
            const GradStep& st = a.step[cur.j];
            const int M = st.M;
            const int n_mt = (M + 127) >> 7;
            const bool has_alpha = st.alpha != nullptr;
            const int J = st.grad_beta ? st.J : 0;
            const uint32_t p = (uint32_t)cur.j & 1u;
            if (cur.cc == 0 && cur.mt == 0) {
                // ---- beta records of this pre-activation node, per row ----
                epi_sync();                                  // everybody is done with the previous lists
                for (int i = te; i < CH_TR * 8; i += CH_EPI_THREADS) s_bmask[i] = 0u;
                epi_sync();
                if (J > 0) {
                    constexpr int TPR = CH_EPI_THREADS / CH_TR;
                    const int row = te / TPR;
                    const int r = row0 + row;
                    if (r < rows) {
                        const size_t jb = (size_t)r * J;
                        for (int jj = te % TPR; jj < J; jj += TPR) {
                            const float sg = __ldg(st.beta_sign + jb + jj);
                            const int lc = (int)__ldg(st.beta_loc + jb + jj);
                            const bool on = sg != 0.f && lc >= 0 && lc < M;
                            s_bloc[row * CHAIN_JMAX + jj] = on ? lc : -1;
                            s_bsg[row * CHAIN_JMAX + jj] = sg;
                            if (on) atomicOr(&s_bmask[row * 8 + (lc >> 5)], 1u << (lc & 31));
                        }
                    }
                    epi_sync();
                }
            }
            if (cur.cc == 0) {                               // first item of an M-tile: its accumulator must be complete
                mbar_wait(&acc_full[p][cur.mt & 1], ((uint32_t)cur.j >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            nxt = cur;
            advance(nxt);
            if (nxt.j < n_steps) issue(nxt, nx);

            const int mt = cur.mt;
            const int m = mt * 128 + q * 32 + lane;
            const bool vm = m < M;
            const int c0 = h * CH_RPW + cur.cc;
            const uint32_t tcol = trow + p * CH_TBUF + (uint32_t)mt * CH_TMT;
            float d[8], y[8];
            tmem_ld8x2(tcol + c0, tcol + CH_TR + c0, d);
            int apos = -1;
            if (has_alpha && vm) apos = st.alpha_pos ? __ldg(st.alpha_pos + m) : m;
            float* const gap = (st.grad_alpha && apos >= 0) ? st.grad_alpha + (size_t)(row0 + c0) * st.n_alpha + apos : nullptr;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                y[i] = 0.f;
                const bool ok = vm && (fast || row0 + c0 + i < rows);
                if (!ok) continue;
                const float z = d[i] + pre[32];
                const float ap = pre[24 + i];
                const Relax8 rx = relax1(pre[i], pre[8 + i], has_alpha, pre[16 + i]);
                y[i] = z * (ap >= 0.f ? rx.d_l : rx.d_u) + (ap < 0.f ? rx.b_u : 0.f);
                if (gap) gap[(size_t)i * st.n_alpha] = (rx.live && ap >= 0.f) ? z * ap : 0.f;
                if (J > 0 && ((s_bmask[(c0 + i) * 8 + (m >> 5)] >> (m & 31)) & 1u)) {
                    const size_t jb = (size_t)(row0 + c0 + i) * J;
                    for (int jj = 0; jj < J; ++jj)
                        if (s_bloc[(c0 + i) * CHAIN_JMAX + jj] == m) {
                            const float sg = s_bsg[(c0 + i) * CHAIN_JMAX + jj];
                            float gb = -sg * z;
                            if (st.beta_bias) gb = fmaf(sg, __ldg(st.beta_bias + jb + jj), gb);
                            st.grad_beta[jb + jj] = gb;
                        }
                }
            }
            if (st.need_y) {
                if (m < ((M + 15) & ~15)) x_store8(X, m, c0, y);       // K range of the next layer (zero padded)
                if (cur.cc == CH_RPW - 8) {                  // half mt of the next layer's operand is complete
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&x_full[mt]);
                }
            }
            if (cur.cc == CH_RPW - 8 && mt == n_mt - 1) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[p]);
            }
        };
        {
            Cur c0{0, 0, 0}, c1;
            float va[33], vb[33];
            issue(c0, va);
            while (c0.j < n_steps) {
                process(c0, va, vb, c1);
                if (c1.j >= n_steps) break;
                process(c1, vb, va, c0);
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CH_TMEM_COLS) : "memory");
    }
}

}  // namespace

cudaError_t chain_grad(const ChainGradArgs& a, cudaStream_t st) {
    Launch _l(K_CHAIN_GRAD, st);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_chain_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM);
        if (e != cudaSuccess) return e;
        // leave what shared memory does not need to the L1: the epilogue's per-row loads allocate L1 lines
        e = cudaFuncSetAttribute(k_chain_grad, cudaFuncAttributePreferredSharedMemoryCarveout, CH_CTAS_PER_SM > 1 ? 100 : (CH_SMEM + 2048) * 100 / (228 * 1024) + 1);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int tiles = (a.rows + CH_TR - 1) / CH_TR;
    k_chain_grad<<<tiles, CH_THREADS, CH_SMEM, st>>>(a);
    return cudaGetLastError();
}

}  // namespace cb
