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
// weights streamed by bulk copies (24 KB blocks, three producer warps, 3-stage ring), the row tile resident in shared
// memory in UMMA MN-major layout (three bf16 planes side by side along N: three MMAs per k-step) and rewritten by the
// epilogue of one layer for the MMAs of the next, TMEM lane = neuron so that every global access of the epilogue
// (l, u, alpha, lA in; dalpha out) is a coalesced 128-byte line, one TMEM slot [main | small-1 | small-2] per M-tile.
// The first layer's K (= n_in, 784 for MNIST) does not fit the resident operand: the epilogue warps stream g_0 through the two 128-k halves of the operand buffer (x_full / x_empty
// barriers) before they have any epilogue work.
//
// Restrictions (the host falls back to the per-layer kernels otherwise): S == 1, layer widths <= CHAIN_KMAX,
// <= CHAIN_JMAX beta records per row and layer.
#include <type_traits>

#include "crown_chain_common.cuh"

namespace cb {

namespace {

using namespace tcc;
using namespace chn;

__global__ void __launch_bounds__(CH_THREADS, CH_CTAS_PER_SM) k_chain_grad(const __grid_constant__ ChainGradArgs a) {
    if (a.kb_cur != nullptr) {
        // ---- k_keepbest_b of this iteration (optimized_bounds.py:473-530), S == 1: sub-domain b = row ----
        const int c_done = a.kb_cur->done, c_improved = a.kb_cur->any_improved, c_patience = a.kb_cur->patience;
        const int c_not_stopped = a.kb_cur->n_not_stopped, c_mask0 = a.kb_cur->any_mask0;
        if (c_done) {
            if (blockIdx.x == 0 && threadIdx.x == 0) *a.kb_next = *a.kb_cur;
            return;
        }
        const int patience = c_improved ? 0 : c_patience + 1;
        const bool stop_final = c_not_stopped == 0;
        // save window: first iteration, second half, or just before leaving (optimized_bounds.py:483-484)
        const bool window = (a.kb_iter < 1) || (a.kb_iter > a.kb_save_from) || stop_final || (patience == a.kb_patience_limit);
        const int b = blockIdx.x * CH_TR + (int)threadIdx.x;
        if (threadIdx.x < CH_TR && b < a.rows) {
            uint8_t sn = 0;
            if (window) {
                if (c_mask0) {
                    if (a.kb_mask0[b]) { a.kb_ret0[b] = a.kb_lb_cur[b]; sn = 1; }
                } else {
                    a.kb_ret0[b] = a.kb_lb_cur[b];          // reference quirk: `ret_0[None] = full_ret_l[None]`
                }
            }
            a.kb_snap[b] = sn;
        }
        const bool done_next = stop_final || patience > a.kb_patience_limit || a.kb_iter == a.kb_iteration - 1;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            OptState n;
            n.patience = patience;
            n.any_improved = 0;
            n.n_not_stopped = 0;
            n.any_mask0 = 0;
            n.done = done_next ? 1 : 0;
            n.n_iter = a.kb_iter + 1;
            n.pad[0] = n.pad[1] = 0;
            *a.kb_next = n;
        }
        if (done_next) return;                               // the same decision in every CTA
    } else if (a.done != nullptr && *a.done != 0) {
        return;
    }
    extern __shared__ __align__(128) uint8_t smem[];      // no-swizzle operands and bulk copies need 16-byte alignment only
    __shared__ __align__(8) uint64_t w_full[CH_WSTAGES];
    __shared__ __align__(8) uint64_t w_empty[CH_WSTAGES];
    __shared__ __align__(8) uint64_t x_full[2];
    __shared__ __align__(8) uint64_t x_empty[2];
    __shared__ __align__(8) uint64_t acc_full[2];            // [TMEM slot = M-tile]
    __shared__ __align__(8) uint64_t acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    uint8_t* const wring = smem;
    uint8_t* const X = smem + CH_WSTAGES * CH_WSTAGE;
    uint8_t* const s_bidx = X + CH_XBYTES;                                             // [rows][CHAIN_KMAX] bytes
    float* const s_bsg = reinterpret_cast<float*>(s_bidx + CH_BIDX_BYTES);             // [rows][JMAX]
    uint8_t* const s_bnext = reinterpret_cast<uint8_t*>(s_bsg) + CH_BVAL_BYTES;        // [rows][JMAX] bytes
    uint8_t* const s_bloc8 = s_bnext + CH_BLOC8_BYTES;                                 // [rows][JMAX] bytes (set-up only)

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
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], CH_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == CH_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(CH_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp >= CH_EPI_WARPS) chain_set_regs(false);         // whole warpgroup; no code path joins the epilogue's before the end
    // M-tile mi of a step accumulates in TMEM slot mi (<= 2 M-tiles per step); each side keeps the parity of a slot's
    // use count for the barrier phases.  MMA order: K chunk (8 k-steps) major, then M-tile - step 0 streams its K
    // through the two operand halves, and in the later steps chunk 0 of BOTH M-tiles has been read before M-tile 0
    // completes and its epilogue overwrites chunk 0 with the next layer's operand.
    if (warp >= CH_EPI_WARPS && warp != CH_WARP_MMA) {
        // ===== weight producers: block wb (<= CH_WBLOCK_KS k-steps of one M-tile, MMA order) goes to ring stage
        // wb % CH_WSTAGES and is fetched by that stage's warp (see crown_chain_common.cuh) =====
        const int my_stage = chain_producer_stage(warp);
        uint32_t wb = 0;
        if (my_stage >= 0)
            for (int j = 0; j < n_steps; ++j) {
                const GradStep& st = a.step[j];
                const int nks = st.Kp >> 4;
                const int n_mt = (st.M + 127) >> 7;
                for (int kc = 0; kc < nks; kc += 8) {
                    const int ke = min(kc + 8, nks);
                    for (int mi = 0; mi < n_mt; ++mi)
                        for (int ks = kc; ks < ke; ks += CH_WBLOCK_KS, ++wb) {
                            const int s = wb % CH_WSTAGES;
                            if (s != my_stage) continue;
                            const uint32_t bytes = (uint32_t)min(CH_WBLOCK_KS, ke - ks) * CH_WKSTEP;
                            mbar_wait(&w_empty[s], ((wb / CH_WSTAGES) & 1u) ^ 1u);
                            if (elect_one()) {
                                mbar_expect_tx(&w_full[s], bytes);
                                bulk_g2s(wring + (size_t)s * CH_WSTAGE, st.wp + ((size_t)mi * nks + ks) * (CH_WKSTEP / 2), bytes,
                                         &w_full[s]);
                            }
                            __syncwarp();
                        }
                }
            }
    } else if (warp == CH_WARP_MMA) {
        // ===== MMA issuer =====
        const uint32_t idesc1 = umma_idesc_bf16(CH_TR) | (1u << 16);     // B (the row tile) is MN-major
        uint32_t wb = 0, xph0 = 0, xph1 = 0, useb = 0;
        const uint64_t a_base = umma_desc(smem_u32(wring), 2048, 128);
        const uint64_t b_base = umma_desc(smem_u32(X), CH_XKG3, 128);
        const uint32_t a_lo0 = (uint32_t)a_base, a_hi = (uint32_t)(a_base >> 32);
        const uint32_t b_lo0 = (uint32_t)b_base, b_hi = (uint32_t)(b_base >> 32);
        for (int j = 0; j < n_steps; ++j) {
            const int nks = a.step[j].Kp >> 4;
            const int n_mt = (a.step[j].M + 127) >> 7;
            for (int kc = 0; kc < nks; kc += 8) {
                const int slot_x = (kc >> 3) & 1;          // 128-k half of the operand buffer
                if (slot_x == 0) { mbar_wait(&x_full[0], xph0); xph0 ^= 1u; }
                else { mbar_wait(&x_full[1], xph1); xph1 ^= 1u; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int ke = min(kc + 8, nks);
                for (int mi = 0; mi < n_mt; ++mi) {
                    const uint32_t slot = (uint32_t)mi & 1u;
                    if (kc == 0) {                           // the epilogue has drained this slot's previous M-tile
                        mbar_wait(&acc_empty[slot], ((useb >> slot) & 1u) ^ 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const uint32_t d0 = tmem_base + slot * CH_TSLOT;
                    for (int ks = kc; ks < ke; ks += CH_WBLOCK_KS, ++wb) {
                        const uint32_t s = wb % CH_WSTAGES;
                        mbar_wait(&w_full[s], (wb / CH_WSTAGES) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const int nk = min(CH_WBLOCK_KS, ke - ks);
                        const uint32_t a_lo = a_lo0 + s * (CH_WSTAGE >> 4);
                        const uint32_t b_lo = b_lo0 + (uint32_t)(slot_x * 8 + (ks & 7)) * (2 * CH_XKG3 >> 4);       // ks even: ks + 1 stays in the half
                        if (elect_one()) {
                            umma_split3(d0, a_lo, a_hi, b_lo, b_hi, idesc1, ks ? 1u : 0u);
                            if (nk > 1) umma_split3(d0, a_lo + (CH_WKSTEP >> 4), a_hi, b_lo + (2 * CH_XKG3 >> 4), b_hi, idesc1, 1u);
                            umma_commit(&w_empty[s]);
                        }
                        __syncwarp();
                    }
                    // This M-tile's accumulators are complete.  M-tile 0's epilogue overwrites operand half 0: it may start
                    // early only if what M-tile 1 still has to read lies in half 1 (two K chunks exactly); step 0
                    // (streamed K, last chunk in either half) and single-chunk steps release both M-tiles at the end.
                    const bool early = nks > 8 && nks <= 16;
                    if (ke == nks && (early || mi == n_mt - 1)) {
                        for (int m2 = early ? mi : 0; m2 <= mi; ++m2) {
                            const uint32_t s2 = (uint32_t)m2 & 1u;
                            if (elect_one()) umma_commit(&acc_full[s2]);
                            __syncwarp();
                            useb ^= 1u << s2;
                        }
                    }
                }
                if (elect_one()) umma_commit(&x_empty[slot_x]);       // the MMAs reading this half have been issued
                __syncwarp();
            }
        }
    } else if (warp < CH_EPI_WARPS) {
        // ===== epilogue warps: TMEM lane = neuron, column = sub-domain row =====
        chain_set_regs(true);
        const int te = threadIdx.x;
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int h = warp >> 2;                 // rows h*CH_RPW .. (h+1)*CH_RPW-1
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
        // Same loop as the pass kernel (crown_chain.cu): per layer and M-tile two items of 8 rows, the operands
        // of the next item (l, u, alpha, lA stash) requested one item ahead in a second register set, the per-neuron
        // constants of the next M-tile (alpha column, bias) one M-tile ahead.
        static_assert(CH_RPW == 16, "an epilogue warp owns two 8-row items per M-tile");
        bool fast = row0 + CH_TR <= rows;                    // all 64 rows valid, element offsets fit 32 bits
        {
            int mx = 1;
            for (int j = 0; j < n_steps; ++j) mx = max(mx, max(a.step[j].M, a.step[j].alpha ? a.step[j].n_alpha : 0));
            if ((unsigned long long)rows * (unsigned long long)mx >= (1ull << 32)) fast = false;
        }
        struct Ops { float l[8], u[8], al[8], ap[8]; };

        auto tile_consts = [&](int j, int mt, int& apos, float& bz) {
            const GradStep& st = a.step[j];
            const int mc = min(mt * 128 + q * 32 + lane, st.M - 1);
            apos = -1;
            bz = 0.f;
            if (st.alpha != nullptr) apos = st.alpha_pos ? __ldg(st.alpha_pos + mc) : mc;
            if (st.bias) bz = __ldg(st.bias + mc);
        };
        auto issue = [&](auto tag, int j, int mt, int cc, int apos, Ops& v) {
            constexpr bool F = decltype(tag)::value;
            const GradStep& st = a.step[j];
            const int M = st.M;
            const int mc = min(mt * 128 + q * 32 + lane, M - 1);
            const int c0 = h * CH_RPW + cc;
            const float* __restrict__ p0 = st.lower;
            const float* __restrict__ p1 = st.upper;
            const float* __restrict__ p2 = st.a_post;
            if constexpr (F) {
                uint32_t o = (uint32_t)(row0 + c0) * (uint32_t)M + (uint32_t)mc;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v.l[i] = __ldg(p0 + o);
                    v.u[i] = __ldg(p1 + o);
                    v.ap[i] = __ldg(p2 + o);
                    o += (uint32_t)M;
                }
                if (apos >= 0) {
                    const float* __restrict__ pa = st.alpha;
                    const uint32_t na = (uint32_t)st.n_alpha;
                    uint32_t oa = (uint32_t)(row0 + c0) * na + (uint32_t)apos;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v.al[i] = __ldg(pa + oa);
                        oa += na;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.al[i] = 0.f;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = row0 + c0 + i;
                    v.l[i] = 0.f; v.u[i] = 0.f; v.al[i] = 0.f; v.ap[i] = 0.f;
                    if (r < rows) {
                        const size_t o = (size_t)r * M + mc;
                        v.l[i] = __ldg(p0 + o);
                        v.u[i] = __ldg(p1 + o);
                        v.ap[i] = __ldg(p2 + o);
                        if (apos >= 0) v.al[i] = __ldg(st.alpha + (size_t)r * st.n_alpha + apos);
                    }
                }
            }
        };
        auto item = [&](auto tag, const GradStep& st, int j, int mt, int cc, const Ops& pre, int apos, float bz) {
            constexpr bool F = decltype(tag)::value;
            const int M = st.M;
            const int m = mt * 128 + q * 32 + lane;
            const bool vm = m < M;
            const bool has_alpha = st.alpha != nullptr;
            const int J = st.grad_beta ? st.J : 0;
            const int c0 = h * CH_RPW + cc;
            float d[8], y[8];
            tmem_ld8x3(trow + (uint32_t)(mt & 1) * CH_TSLOT + c0, d);
            unsigned okm = 0xffu;                            // rows of this item this lane may store to
            if constexpr (!F) {
                okm = 0u;
#pragma unroll
                for (int i = 0; i < 8; ++i) okm |= (row0 + c0 + i < rows) ? (1u << i) : 0u;
            }
            if (!vm) okm = 0u;
            const bool want_ga = st.grad_alpha != nullptr && apos >= 0 && vm;
            float ga[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float z = d[i] + bz;
                const float ap = pre.ap[i];
                const Relax8 rx = relax1(pre.l[i], pre.u[i], has_alpha, pre.al[i]);
                y[i] = z * (ap >= 0.f ? rx.d_l : rx.d_u) + (ap < 0.f ? rx.b_u : 0.f);
                ga[i] = (rx.live && ap >= 0.f) ? z * ap : 0.f;
                d[i] = z;
            }
            if (want_ga) {
                float* __restrict__ gp = st.grad_alpha;
                if constexpr (F) {
                    const uint32_t na = (uint32_t)st.n_alpha;
                    uint32_t oa = (uint32_t)(row0 + c0) * na + (uint32_t)apos;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        gp[oa] = ga[i];
                        oa += na;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if ((okm >> i) & 1u) gp[(size_t)(row0 + c0 + i) * st.n_alpha + apos] = ga[i];
                }
            }
            if (J > 0) {                                     // d(beta) of the records at this neuron (beta_crown.py:163-204 transposed)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    unsigned bi = s_bidx[(c0 + i) * CHAIN_KMAX + m];
                    while (bi) {
                        const int jj = (int)bi - 1;
                        const size_t jb = (size_t)(row0 + c0 + i) * J;
                        const float sg = s_bsg[(c0 + i) * CHAIN_JMAX + jj];
                        float gb = -sg * d[i];
                        if (st.beta_bias) gb = fmaf(sg, __ldg(st.beta_bias + jb + jj), gb);
                        st.grad_beta[jb + jj] = gb;
                        bi = s_bnext[(c0 + i) * CHAIN_JMAX + jj];
                    }
                }
            }
            if (st.need_y) {
                if (mt * 128 + q * 32 + 32 > M) {            // ragged warp: neurons >= M feed zeros to the next layer
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = vm ? y[i] : 0.f;
                }
                if (m < ((M + 15) & ~15)) x_store8(X, m, c0, y);       // K range of the next layer (zero padded)
            }
        };
        auto run = [&](auto tag) {
            Ops va, vb;
            int apos, apos_n = -1;
            float bz, bz_n = 0.f;
            uint32_t useb = 0;                               // bit s = parity of slot s' use count
            bool map_dirty = true;                           // s_bidx holds non-zero entries (or was never cleared)
            tile_consts(0, 0, apos, bz);
            issue(tag, 0, 0, 0, apos, va);
            for (int j = 0; j < n_steps; ++j) {
                const GradStep& st = a.step[j];
                const int M = st.M;
                const int n_mt = (M + 127) >> 7;
                {
                    // ---- beta records of this pre-activation node, per row ----
                    const int J = st.grad_beta ? st.J : 0;
                    if (J > 0 || map_dirty) {
                        epi_sync();                          // everybody is done with the previous map
                        for (int i = te; i < CH_BIDX_BYTES / 16; i += CH_EPI_THREADS)
                            reinterpret_cast<uint4*>(s_bidx)[i] = make_uint4(0u, 0u, 0u, 0u);
                        map_dirty = J > 0;
                    }
                    if (J > 0) {
                        constexpr int TPR = CH_EPI_THREADS / CH_TR;
                        {
                            const int row = te / TPR;
                            const int r = row0 + row;
                            const size_t jb = (size_t)(r < rows ? r : 0) * J;
                            for (int jj = te % TPR; jj < J; jj += TPR) {
                                float sg = 0.f;
                                int lc = 0;
                                if (r < rows) {
                                    sg = __ldg(st.beta_sign + jb + jj);
                                    lc = (int)__ldg(st.beta_loc + jb + jj);
                                }
                                const bool on = sg != 0.f && lc >= 0 && lc < M;
                                s_bsg[row * CHAIN_JMAX + jj] = on ? sg : 0.f;
                                s_bloc8[row * CHAIN_JMAX + jj] = (uint8_t)(on ? lc : 0);
                            }
                        }
                        epi_sync();
                        if (te < CH_TR && row0 + te < rows) {            // one thread per row: no races on the map
                            for (int jj = 0; jj < J; ++jj) {
                                if (s_bsg[te * CHAIN_JMAX + jj] == 0.f) continue;
                                const int lc = s_bloc8[te * CHAIN_JMAX + jj];
                                s_bnext[te * CHAIN_JMAX + jj] = s_bidx[te * CHAIN_KMAX + lc];      // chain records at one neuron
                                s_bidx[te * CHAIN_KMAX + lc] = (uint8_t)(jj + 1);
                            }
                        }
                        epi_sync();
                    }
                }
                for (int mt = 0; mt < n_mt; ++mt) {
                    int j2 = j, mt2 = mt + 1;                // the M-tile after this one
                    if (mt2 == n_mt) { mt2 = 0; j2 = j + 1; }
                    const uint32_t slot = (uint32_t)mt & 1u;
                    mbar_wait(&acc_full[slot], (useb >> slot) & 1u);       // this M-tile's accumulators are complete
                    useb ^= 1u << slot;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    issue(tag, j, mt, 8, apos, vb);
                    if (j2 < n_steps) tile_consts(j2, mt2, apos_n, bz_n);
                    item(tag, st, j, mt, 0, va, apos, bz);
                    if (j2 < n_steps) issue(tag, j2, mt2, 0, apos_n, va);
                    item(tag, st, j, mt, 8, vb, apos, bz);
                    if (st.need_y) {                         // half mt of the next layer's operand is complete
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&x_full[mt]);
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      // the slot is drained
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[slot]);
                    apos = apos_n;
                    bz = bz_n;
                }
            }
        };
        if (fast) run(std::true_type{});
        else run(std::false_type{});
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == CH_WARP_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CH_TMEM_COLS) : "memory");
    }
}

}  // namespace

cudaError_t chain_grad(const ChainGradArgs& a, cudaStream_t st) {
    Launch _l(K_CHAIN_GRAD, st);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_chain_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_GRAD);
        if (e != cudaSuccess) return e;
        // leave what shared memory does not need to the L1: the epilogue's per-row loads allocate L1 lines
        e = cudaFuncSetAttribute(k_chain_grad, cudaFuncAttributePreferredSharedMemoryCarveout, CH_CTAS_PER_SM > 1 ? 100 : (CH_SMEM_GRAD + 2048) * 100 / (228 * 1024) + 1);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int tiles = (a.rows + CH_TR - 1) / CH_TR;
    k_chain_grad<<<tiles, CH_THREADS, CH_SMEM_GRAD, st>>>(a);
    return cudaGetLastError();
}

}  // namespace cb
