// Whole-network CROWN pass for fully-connected ReLU chains in ONE kernel (sm_100a, tcgen05 + TMA).
//
// A CTA owns 64 sub-domain rows (row r = s*Bd + b) and walks the network backwards,
//   A_{k-1}' = A_k . W_k            (tensor cores; operators/linear.py:167-175)
//   A_{k-1}  = relax(A_{k-1}')      (epilogue; operators/relu.py:456-494, clampmult.py:17-43, beta_crown.py:163-204)
// without the coefficient matrix ever leaving the SM: the contraction is issued TRANSPOSED,
//   D^T[neurons(128 per MMA) x rows(64)] = W_k^T[neurons x K] . A_k^T[K x rows],
// so that (i) the weights are the streamed M-side operand (bulk copies from the L2 in 24 KB blocks of two k-steps,
// three producer warps feeding a 3-stage ring), (ii) the sub-domain tile is the N-side operand, small enough
// (64 rows x 256 k x 3 bf16 planes = 96 KB) to stay resident in shared memory, where the epilogue of one layer writes
// it in UMMA layout for the MMAs of the next, and (iii) a TMEM lane is a NEURON: the 32 lanes of an epilogue warp
// read 32 consecutive neurons of one sub-domain row of l / u / alpha / x_L / x_U and write lA the same way, i.e.
// every global access of the epilogue is a coalesced 128-byte line with no staging.
//
// fp32 fidelity: the bf16x3 split of crown_tc.cu (see its header), here with the three planes of the row tile side
// by side along N so that the six products are THREE MMAs per k-step (N = 192 / 128 / 64 on one descriptor) into three
// accumulators [main | small-1 | small-2] (crown_chain_common.cuh).  TMEM holds two such M-tile slots (2 x 192
// columns): M-tile mt+1 accumulates while the epilogue drains M-tile mt.
//
// Warp roles (640 threads): warps 0-15 epilogue (four warpgroups, grown to 112 registers by setmaxnreg), warp 17 MMA
// issuer, warps 16 / 18 / 19 weight producers (32 registers).
//
// Operand layouts (bf16):
//   W_k^T packed by tc_pack_weight(TR = 128): [m / 128][k / 16][plane][(k / 8) % 2][m % 128][k % 8]
//          one (M-tile, k-step) = 12 KB contiguous; LBO = 2048 B, SBO = 128 B
//   X (shared memory only), MN-major (rows contiguous) so that an epilogue thread (one neuron k, 8 rows)
//          writes ONE 16-byte word per plane: element (row n, k) of plane p at
//          (k / 8) * 3072 + p * 1024 + (n / 8) * 128 + (k % 8) * 16 + (n % 8) * 2;  LBO = 3072 B, SBO = 128 B
#include <type_traits>

#include "crown_chain_common.cuh"

namespace cb {

namespace {

using namespace tcc;
using namespace chn;

__global__ void __launch_bounds__(CH_THREADS, CH_CTAS_PER_SM) k_chain_pass(const __grid_constant__ ChainArgs a) {
    if (a.done != nullptr && *a.done != 0) return;
    extern __shared__ __align__(128) uint8_t smem[];      // no-swizzle operands and bulk copies need 16-byte alignment only
    __shared__ __align__(8) uint64_t w_full[CH_WSTAGES];
    __shared__ __align__(8) uint64_t w_empty[CH_WSTAGES];
    __shared__ __align__(8) uint64_t x_full[2];
    __shared__ __align__(8) uint64_t acc_full[2];            // [TMEM slot = M-tile parity]
    __shared__ __align__(8) uint64_t acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    uint8_t* const wring = smem;
    uint8_t* const X = smem + CH_WSTAGES * CH_WSTAGE;
    uint8_t* const s_bidx = X + CH_XBYTES;                                             // [rows][CHAIN_KMAX] bytes
    float* const s_bvs = reinterpret_cast<float*>(s_bidx + CH_BIDX_BYTES);             // [rows][JMAX]
    float* const s_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_bvs) + CH_BVAL_BYTES);       // [4][rows]
    uint8_t* const s_bloc8 = reinterpret_cast<uint8_t*>(s_part);                       // [rows][JMAX] bytes (set-up only)
    float* const s_extra = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_part) + CH_PART_BYTES);    // [rows]
    float4* const s_acc = reinterpret_cast<float4*>(s_extra + CH_TR);                  // [4][epilogue threads]: per-thread bias sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * CH_TR;
    long long* const dbg = a.dbg ? a.dbg + 64 * (size_t)blockIdx.x : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();

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
    // M-tiles are enumerated step by step, mt ascending; M-tile mt accumulates in TMEM slot mt & 1, and each side keeps
    // the parity of every slot's use count for the barrier phases.  MMA order inside a step: hidden steps (the epilogue rewrites X in
    // place for the next layer) take their M-tiles in pairs, K chunk (8 k-steps) major - chunk 0 of BOTH M-tiles has
    // been read before M-tile 0 completes and its epilogue overwrites chunk 0; the last step (no X rewrite) goes
    // M-tile by M-tile, so that the MMAs of M-tile mt+1 run under the epilogue of M-tile mt.
    if (warp >= CH_EPI_WARPS && warp != CH_WARP_MMA) {
        // ===== weight producers: block wb (<= CH_WBLOCK_KS k-steps of one M-tile, MMA order) goes to ring stage
        // wb % CH_WSTAGES and is fetched by that stage's warp =====
#ifdef CB_CHAIN_NORING
        const int my_stage = -1;
#else
        const int my_stage = chain_producer_stage(warp);
#endif
        uint32_t wb = 0;
        if (my_stage >= 0)
            for (int j = 0; j < a.n_steps; ++j) {
                const ChainStep& st = a.step[j];
                const int nks = st.Kp >> 4;
                const int n_mt = (st.M + 127) >> 7;
                const int gs = (j == a.n_steps - 1) ? 1 : 2;
                for (int mt0 = 0; mt0 < n_mt; mt0 += gs) {
                    const int nmt = min(gs, n_mt - mt0);
                    for (int kc = 0; kc < nks; kc += 8) {
                        const int ke = min(kc + 8, nks);
                        for (int mi = 0; mi < nmt; ++mi)
                            for (int ks = kc; ks < ke; ks += CH_WBLOCK_KS, ++wb) {
                                const int s = wb % CH_WSTAGES;
                                if (s != my_stage) continue;
                                const uint32_t bytes = (uint32_t)min(CH_WBLOCK_KS, ke - ks) * CH_WKSTEP;
                                mbar_wait(&w_empty[s], ((wb / CH_WSTAGES) & 1u) ^ 1u);
                                if (elect_one()) {
#ifdef CB_CHAIN_NOSTREAM                          // timing experiment: garbage weights, no copy
                                    mbar_arrive(&w_full[s]);
#else
                                    mbar_expect_tx(&w_full[s], bytes);
                                    bulk_g2s(wring + (size_t)s * CH_WSTAGE, st.wp + ((size_t)(mt0 + mi) * nks + ks) * (CH_WKSTEP / 2),
                                             bytes, &w_full[s]);
#endif
                                }
                                __syncwarp();
                            }
                    }
                }
            }
    } else if (warp == CH_WARP_MMA) {
        // ===== MMA issuer: the warp runs the loop converged, one elected lane issues =====
        const uint32_t idesc1 = umma_idesc_bf16(CH_TR) | (1u << 16);     // B (the row tile) is MN-major
        uint32_t wb = 0, gc = 0, xph0 = 0, xph1 = 0, useb = 0;     // useb: bit s = parity of slot s' use count
        // descriptors differ from these bases only in the start-address field (16-byte units) of the low word
        const uint64_t a_base = umma_desc(smem_u32(wring), 2048, 128);
        const uint64_t b_base = umma_desc(smem_u32(X), CH_XKG3, 128);
        const uint32_t a_lo0 = (uint32_t)a_base, a_hi = (uint32_t)(a_base >> 32);
        const uint32_t b_lo0 = (uint32_t)b_base, b_hi = (uint32_t)(b_base >> 32);
        for (int j = 0; j < a.n_steps; ++j) {
            const int nks = a.step[j].Kp >> 4;
            const int n_mt = (a.step[j].M + 127) >> 7;
            const int gs = (j == a.n_steps - 1) ? 1 : 2;
            for (int mt0 = 0; mt0 < n_mt; mt0 += gs, ++gc) {
                const int nmt = min(gs, n_mt - mt0);
                if (dbg && gc < 10 && lane == 0) dbg[1 + 3 * gc] = clock64();
                for (int kc = 0; kc < nks; kc += 8) {
                    if (mt0 == 0) {                            // X chunk kc/8 of this step: written by the epilogue above
                        if (kc == 0) { mbar_wait(&x_full[0], xph0); xph0 ^= 1u; }
                        else { mbar_wait(&x_full[1], xph1); xph1 ^= 1u; }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (dbg && gc < 10 && kc == 0 && lane == 0) dbg[2 + 3 * gc] = clock64();
                    }
                    const int ke = min(kc + 8, nks);
                    for (int mi = 0; mi < nmt; ++mi) {
                        const uint32_t slot = (uint32_t)(mt0 + mi) & 1u;
                        if (kc == 0) {                         // the epilogue has drained this slot's previous M-tile
                            mbar_wait(&acc_empty[slot], ((useb >> slot) & 1u) ^ 1u);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        const uint32_t d0 = tmem_base + slot * CH_TSLOT;
                        for (int ks = kc; ks < ke; ks += CH_WBLOCK_KS, ++wb) {
                            const uint32_t s = wb % CH_WSTAGES;
#ifndef CB_CHAIN_NORING                             // timing experiment: the MMA warp never waits for weights
                            mbar_wait(&w_full[s], (wb / CH_WSTAGES) & 1u);
#endif
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const int nk = min(CH_WBLOCK_KS, ke - ks);
                            const uint32_t a_lo = a_lo0 + s * (CH_WSTAGE >> 4);
                            const uint32_t b_lo = b_lo0 + (uint32_t)ks * (2 * CH_XKG3 >> 4);
                            if (elect_one()) {
#ifndef CB_CHAIN_NOMMA                              // timing experiment: no MMAs, the ring is released at once
                                umma_split3(d0, a_lo, a_hi, b_lo, b_hi, idesc1, ks ? 1u : 0u);
                                if (nk > 1) umma_split3(d0, a_lo + (CH_WKSTEP >> 4), a_hi, b_lo + (2 * CH_XKG3 >> 4), b_hi, idesc1, 1u);
#endif
#ifndef CB_CHAIN_NORING
                                umma_commit(&w_empty[s]);
#endif
                            }
                            __syncwarp();
                        }
                        // This M-tile's accumulators are complete.  With a single K chunk the epilogue of M-tile 0 must
                        // not start (and overwrite X) before M-tile 1's MMAs have read it: both are released at the end.
                        if (ke == nks && (nks > 8 || mi == nmt - 1)) {
                            const int first = nks > 8 ? mi : 0;
                            for (int m2 = first; m2 <= mi; ++m2) {
                                const uint32_t s2 = (uint32_t)(mt0 + m2) & 1u;
                                if (elect_one()) umma_commit(&acc_full[s2]);
                                __syncwarp();
                                useb ^= 1u << s2;
                            }
                        }
                    }
                }
                if (dbg && gc < 10 && lane == 0) dbg[3 + 3 * gc] = clock64();
            }
        }
        if (dbg && lane == 0) dbg[62] = clock64();
    } else if (warp < CH_EPI_WARPS) {
        // ===== epilogue warps: TMEM lane = neuron, column = sub-domain row =====
        chain_set_regs(true);
        // A warp owns TMEM lane quarter q (32 neurons of every M-tile) and CH_RPW consecutive rows of the tile.
        const int te = threadIdx.x;
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int h = warp >> 2;                 // rows h*CH_RPW .. (h+1)*CH_RPW-1
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const int Bd = a.Bd, S = a.S, rows = a.rows;
        const int n_steps = a.n_steps;

        // ---- 0. pack C into X (the operand of the output layer), C . b_out into the row sums ----
        // (called from run() AFTER the first item's operands have been requested: the cold misses overlap)
        auto pack_C = [&]() {
            const int row = te % CH_TR, kg0 = te / CH_TR;
            const int r = row0 + row;
            const bool vr = r < rows;
            const int b = vr ? r % Bd : 0, s = vr ? r / Bd : 0;
            const float* crow = a.C + ((size_t)b * S + s) * a.n_out;
            const int Kp0 = a.step[0].Kp;
            for (int kg = kg0; kg < (Kp0 >> 3); kg += CH_EPI_THREADS / CH_TR) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (vr && kg * 8 + i < a.n_out) ? __ldg(crow + kg * 8 + i) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) x_store(X, kg * 8 + i, row, v[i]);
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
        };

        // Work of a warp = for every layer and M-tile two ITEMS of 8 rows (A: rows c0..c0+7, B: c0+8..c0+15) of the
        // 32 neurons of its lane quarter.  The loop is software pipelined by one item: the l / u / alpha values of
        // the next item are requested before the current one is processed (two register sets that swap roles),
        // and the per-neuron constants of the next M-tile (alpha column, bias below) one M-tile ahead, so that no
        // global-load latency sits on the epilogue -> MMA -> epilogue chain.
        // The bias terms A^- . b_u + A . b_below are summed per thread over all layers (s_acc: four float4 per thread = its
        // 16 rows; registers are needed for the operand sets) and reduced across lanes ONCE at the end; neurons >= M of a ragged M-tile and rows >= rows of a ragged tile
        // need no arithmetic masks because their accumulator values are exact zeros (zero-padded weights / C rows).
        static_assert(CH_RPW == 16, "an epilogue warp owns two 8-row items per M-tile");
        // fast tiles: all 64 rows valid and of one spec row s (b = boff + local row), element offsets fit 32 bits
        const int s_first = row0 / Bd;
        bool fast = (row0 + CH_TR <= rows) && (S == 1 || s_first == (row0 + CH_TR - 1) / Bd);
        {
            int mx = 1;
            for (int j = 0; j < n_steps; ++j) mx = max(mx, max(a.step[j].M, a.step[j].alpha ? a.step[j].n_alpha : 0));
            if ((unsigned long long)rows * (unsigned long long)mx >= (1ull << 32)) fast = false;
        }
        const int boff = row0 - s_first * Bd;
        struct Ops { float l[8], u[8], al[8]; };      // l (x_L), u (x_U), alpha of 8 rows x this lane's neuron

        // per-neuron constants of M-tile (j, mt): alpha column (or -1) and the bias of the Linear below
        auto tile_consts = [&](int j, int mt, int& apos, float& bb) {
            apos = -1;
            bb = 0.f;
            if (j >= n_steps - 1) return;                    // the concretize step has neither
            const ChainStep& st = a.step[j];
            const int mc = min(mt * 128 + q * 32 + lane, st.M - 1);
            if (st.alpha != nullptr) apos = st.alpha_pos ? __ldg(st.alpha_pos + mc) : mc;
            if (st.bias_below) bb = __ldg(st.bias_below + mc);
        };
        // request the per-row operands of item (j, mt, cc)
        auto issue = [&](auto tag, int j, int mt, int cc, int apos, Ops& v) {
            constexpr bool F = decltype(tag)::value;
            const ChainStep& st = a.step[j];
            const bool lastst = j == n_steps - 1;
            const int M = st.M;
            const int mc = min(mt * 128 + q * 32 + lane, M - 1);
            const float* __restrict__ p0 = lastst ? a.x_L : st.lower;
            const float* __restrict__ p1 = lastst ? a.x_U : st.upper;
            const int c0 = h * CH_RPW + cc;
            if constexpr (F) {
                uint32_t o = (uint32_t)(boff + c0) * (uint32_t)M + (uint32_t)mc;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v.l[i] = __ldg(p0 + o);
                    v.u[i] = __ldg(p1 + o);
                    o += (uint32_t)M;
                }
                if (apos >= 0) {
                    const float* __restrict__ pa = st.alpha;
                    const uint32_t na = (uint32_t)st.n_alpha;
                    uint32_t oa = (uint32_t)((a.S1 == 1 ? boff : row0) + c0) * na + (uint32_t)apos;
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
                    v.l[i] = 0.f; v.u[i] = 0.f; v.al[i] = 0.f;
                    if (r < rows) {
                        const int b = r % Bd;
                        v.l[i] = __ldg(p0 + (size_t)b * M + mc);
                        v.u[i] = __ldg(p1 + (size_t)b * M + mc);
                        if (apos >= 0) v.al[i] = __ldg(st.alpha + ((a.S1 == 1) ? (size_t)b : (size_t)r) * st.n_alpha + apos);
                    }
                }
            }
        };
        // sum part[i] over the 32 lanes (fixed order) into slot[c0 + i]: 9 shuffles
        auto reduce8 = [&](float (&part)[8], float* slot, int c0) {
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
            float w[4], z[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float send = b4 ? part[i] : part[i + 4];
                const float keep = b4 ? part[i + 4] : part[i];
                w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float send = b3 ? w[i] : w[i + 2];
                const float keep = b3 ? w[i + 2] : w[i];
                z[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            float y = (b2 ? z[1] : z[0]) + __shfl_xor_sync(0xffffffffu, b2 ? z[0] : z[1], 4);
            y += __shfl_xor_sync(0xffffffffu, y, 2);
            y += __shfl_xor_sync(0xffffffffu, y, 1);
            if ((lane & 3) == 0) slot[c0 + (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0)] = y;      // every slot has one writer
        };
        // one item: accumulator -> relaxation (or concretisation) -> operand of the next layer; acc += bias terms
        auto item = [&](auto tag, const ChainStep& st, bool last, int mt, int cc, const Ops& pre, float4* sacc, float bb) {
            constexpr bool F = decltype(tag)::value;
            const int M = st.M;
            const int m = mt * 128 + q * 32 + lane;
            const bool vm = m < M;
            const int c0 = h * CH_RPW + cc;
            float d[8], acc[8];
            tmem_ld8x3(trow + (uint32_t)(mt & 1) * CH_TSLOT + c0, d);
#ifdef CB_CHAIN_NOEPI                                 // timing experiment: the epilogue only hands the barriers on
            if (d[0] != 123.456f) return;
#endif
            {
                const float4 t0 = sacc[0], t1 = sacc[CH_EPI_THREADS];
                acc[0] = t0.x; acc[1] = t0.y; acc[2] = t0.z; acc[3] = t0.w;
                acc[4] = t1.x; acc[5] = t1.y; acc[6] = t1.z; acc[7] = t1.w;
            }
            unsigned okm = 0xffu;                            // rows of this item this lane may store to
            if constexpr (!F) {
                okm = 0u;
#pragma unroll
                for (int i = 0; i < 8; ++i) okm |= (row0 + c0 + i < rows) ? (1u << i) : 0u;
            }
            if (!vm) okm = 0u;
            if (!last) {
                const bool has_alpha = st.alpha != nullptr;
                const int J = st.J;
                if (st.lA != nullptr) {
                    float* __restrict__ lap = st.lA;
                    if constexpr (F) {
                        if (vm) {
                            uint32_t o = (uint32_t)(row0 + c0) * (uint32_t)M + (uint32_t)m;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                lap[o] = d[i];
                                o += (uint32_t)M;
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if ((okm >> i) & 1u) lap[(size_t)(row0 + c0 + i) * M + m] = d[i];
                    }
                }
                // ---- operators/relu.py:456-494, same arithmetic as relax1() ----
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float dv = d[i];
                    const float l = pre.l[i], u = pre.u[i];
                    const float lb_r = fminf(l, 0.f);
                    const float ub_r = fmaxf(fmaxf(u, 0.f), lb_r + 1e-8f);
                    const float d_u = slope_div(ub_r, ub_r - lb_r);
                    // branch-free: alpha clipped to [0, 1] where unstable, 1 / 0 where stably active / inactive
                    float d_la = (u <= 0.f) ? 0.f : fminf(fmaxf(pre.al[i], 0.f), 1.f);
                    d_la = (l >= 0.f) ? 1.f : d_la;
                    const float d_l = has_alpha ? d_la : ((d_u > 0.5f) ? 1.f : 0.f);
                    const float a_pos = fmaxf(dv, 0.f), a_neg = fminf(dv, 0.f);
                    y[i] = d_l * a_pos + d_u * a_neg;
                    acc[i] += fmaf(y[i], bb, a_neg * (-lb_r * d_u));
                }
                if (J > 0) {                                 // beta_crown.py:163-204: A -= beta * sign at the split neurons
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const unsigned bi = s_bidx[(c0 + i) * CHAIN_KMAX + m];
                        if (bi) {
                            const float bv = s_bvs[(c0 + i) * CHAIN_JMAX + bi - 1];
                            y[i] -= bv;
                            acc[i] = fmaf(-bv, bb, acc[i]);
                        }
                    }
                }
                if (m < ((M + 15) & ~15)) x_store8(X, m, c0, y);       // K range of the next layer (zero padded)
            } else {
                // ---- concretise against the input box (perturbations.py:154-183) ----
                const int w32 = (M + 31) >> 5;
                float* const g0p = a.g0_plain ? a.g0_plain + (size_t)(row0 + c0) * M + m : nullptr;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float av = d[i];
                    const float cen = (pre.u[i] + pre.l[i]) / 2.0f, dif = (pre.u[i] - pre.l[i]) / 2.0f;
                    acc[i] += av * cen - fabsf(av) * dif;
                    if (a.sign_pos) {
                        const unsigned pm = __ballot_sync(0xffffffffu, av > 0.f);
                        const unsigned nm = __ballot_sync(0xffffffffu, av < 0.f);
                        if (lane == 0 && (F || row0 + c0 + i < rows) && (m >> 5) < w32) {
                            a.sign_pos[(size_t)(row0 + c0 + i) * w32 + (m >> 5)] = pm;
                            a.sign_neg[(size_t)(row0 + c0 + i) * w32 + (m >> 5)] = nm;
                        }
                    }
                    if (g0p && ((okm >> i) & 1u)) {
                        const float sg = (av > 0.f) ? 1.f : ((av < 0.f) ? -1.f : 0.f);
                        g0p[(size_t)i * M] = cen - sg * dif;
                    }
                }
            }
            sacc[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            sacc[CH_EPI_THREADS] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        };
        auto run = [&](auto tag) {
            Ops va, vb;
            float4* const saccA = s_acc + te;                // this thread's bias sums: rows c0..c0+7 (two float4 slots)
            float4* const saccB = saccA + 2 * CH_EPI_THREADS;                       // rows c0+8..c0+15
#pragma unroll
            for (int i = 0; i < 4; ++i) saccA[i * CH_EPI_THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
            int apos, apos_n = -1;
            float bb, bb_n = 0.f;
            uint32_t useb = 0, tcnt = 0;                // useb: bit s = parity of slot s' use count
            bool map_dirty = true;                      // s_bidx holds non-zero entries (or was never cleared)
            tile_consts(0, 0, apos, bb);
            issue(tag, 0, 0, 0, apos, va);
            pack_C();
            for (int j = 0; j < n_steps; ++j) {
                const ChainStep& st = a.step[j];
                const int M = st.M;
                const int n_mt = (M + 127) >> 7;
                const bool last = (j == n_steps - 1);        // the concretize step
                if (!last) {
                    // ---- beta records of the pre-activation node, per row (beta_crown.py:163-204) ----
                    const int J = st.J;
                    if (J > 0 || map_dirty) {
                        epi_sync();                          // everybody is done with the previous map
                        for (int i = te; i < CH_BIDX_BYTES / 16; i += CH_EPI_THREADS)
                            reinterpret_cast<uint4*>(s_bidx)[i] = make_uint4(0u, 0u, 0u, 0u);
                        map_dirty = J > 0;
                    }
                    if (J > 0) {
                        constexpr int TPR = CH_EPI_THREADS / CH_TR;      // threads per row
                        {
                            const int row = te / TPR;
                            const int r = row0 + row;
                            const size_t jb = (size_t)((r < rows ? r : 0) % Bd) * J;
                            float t = 0.f;                    // this thread's share of sum_j beta_j sign_j bias_j
                            for (int jj = te % TPR; jj < J; jj += TPR) {
                                float vs = 0.f;
                                int lc = 0;
                                if (r < rows) {
                                    vs = __ldg(st.beta_val + jb + jj) * __ldg(st.beta_sign + jb + jj);
                                    lc = (int)__ldg(st.beta_loc + jb + jj);
                                }
                                const bool on = vs != 0.f && lc >= 0 && lc < M;
                                if (on && st.beta_bias != nullptr) t = fmaf(vs, __ldg(st.beta_bias + jb + jj), t);
                                s_bvs[row * CHAIN_JMAX + jj] = on ? vs : 0.f;
                                s_bloc8[row * CHAIN_JMAX + jj] = (uint8_t)(on ? lc : 0);
                            }
                            // the TPR threads of a row are neighbours in one warp: fixed-order tree, then the row sum
                            static_assert(TPR == 8, "row reduction below");
                            t += __shfl_xor_sync(0xffffffffu, t, 1);
                            t += __shfl_xor_sync(0xffffffffu, t, 2);
                            t += __shfl_xor_sync(0xffffffffu, t, 4);
                            if ((te % TPR) == 0) s_extra[row] += t;
                        }
                        epi_sync();
                        if (te < CH_TR && row0 + te < rows) {            // one thread per row: no races on the map
                            for (int jj = 0; jj < J; ++jj) {
                                const float vs = s_bvs[te * CHAIN_JMAX + jj];
                                if (vs == 0.f) continue;
                                const int lc = s_bloc8[te * CHAIN_JMAX + jj];
                                const unsigned cur = s_bidx[te * CHAIN_KMAX + lc];
                                if (cur == 0) s_bidx[te * CHAIN_KMAX + lc] = (uint8_t)(jj + 1);
                                else s_bvs[te * CHAIN_JMAX + cur - 1] += vs;           // same neuron twice: one combined record
                            }
                        }
                        epi_sync();
                    }
                }
                for (int mt = 0; mt < n_mt; ++mt) {
                    const uint32_t slot = (uint32_t)mt & 1u;
                    int j2 = j, mt2 = mt + 1;                // the M-tile after this one
                    if (mt2 == n_mt) { mt2 = 0; j2 = j + 1; }
                    mbar_wait(&acc_full[slot], (useb >> slot) & 1u);       // this M-tile's accumulators are complete
                    useb ^= 1u << slot;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (dbg && te == 0 && tcnt < 14) dbg[32 + 2 * tcnt] = clock64();
                    issue(tag, j, mt, 8, apos, vb);
                    if (j2 < n_steps) tile_consts(j2, mt2, apos_n, bb_n);
                    item(tag, st, last, mt, 0, va, saccA, bb);
                    if (j2 < n_steps) issue(tag, j2, mt2, 0, apos_n, va);
                    item(tag, st, last, mt, 8, vb, saccB, bb);
                    if (!last) {                             // chunk mt of the next layer's operand is complete
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&x_full[mt]);
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      // the slot is drained
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[slot]);
                    if (dbg && te == 0 && tcnt < 14) dbg[33 + 2 * tcnt] = clock64();
                    ++tcnt;
                    apos = apos_n;
                    bb = bb_n;
                }
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const float4 t0 = saccA[(2 * hh) * CH_EPI_THREADS], t1 = saccA[(2 * hh + 1) * CH_EPI_THREADS];
                float part[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                reduce8(part, s_part + q * CH_TR, h * CH_RPW + 8 * hh);
            }
        };
        if (fast) run(std::true_type{});
        else run(std::false_type{});
        // ---- lower bounds: fixed summation order over the per-warp slots ----
        epi_sync();
        if (te < CH_TR) {                                // whole warps: the keep-best counters are warp-aggregated
            const int r = row0 + te;
            const bool vr = r < rows;
            float t = s_extra[te];
#pragma unroll
            for (int i = 0; i < 4; ++i) t += s_part[i * CH_TR + te];          // one slot per TMEM lane quarter
            if (vr) a.lb[(size_t)(r % Bd) * S + r / Bd] = t;
            if (a.kb_state != nullptr) {
                // ---- keep-best bookkeeping of sub-domain b = r (S == 1): same arithmetic as k_keepbest_a ----
                bool improved = false, not_stopped = false, m0 = false;
                if (vr) {
                    float bl = -INFINITY, br = t, r0 = t;
                    if (a.kb_iter != 0) { bl = a.kb_best_l[r]; br = a.kb_best_ret[r]; r0 = a.kb_ret0[r]; }
                    else a.kb_ret0[r] = t;
                    const bool stop = a.kb_rhs != nullptr && t > a.kb_rhs[r];
                    a.kb_stopped[r] = stop ? 1 : 0;
                    improved = t > bl;
                    if (improved) { bl = fmaxf(t, bl); br = fmaxf(t, br); }
                    if (improved || a.kb_iter == 0) { a.kb_best_l[r] = bl; a.kb_best_ret[r] = br; }
                    not_stopped = !stop;
                    m0 = t > r0;
                    a.kb_mask0[r] = m0 ? 1 : 0;
                }
                const unsigned n_imp = __ballot_sync(0xffffffffu, improved), n_ns = __ballot_sync(0xffffffffu, not_stopped),
                               n_m0 = __ballot_sync(0xffffffffu, m0);
                if (lane == 0) {
                    if (n_imp) atomicOr(&a.kb_state->any_improved, 1);
                    if (n_ns) atomicAdd(&a.kb_state->n_not_stopped, __popc(n_ns));
                    if (n_m0) atomicOr(&a.kb_state->any_mask0, 1);
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == CH_WARP_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CH_TMEM_COLS) : "memory");
    }
}

}  // namespace

size_t chain_smem_bytes() { return CH_SMEM_PASS; }

cudaError_t chain_pass(const ChainArgs& a, cudaStream_t st) {
    Launch _l(K_CHAIN_PASS, st);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_chain_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_PASS);
        if (e != cudaSuccess) return e;
        // leave what shared memory does not need to the L1: the epilogue's per-row loads allocate L1 lines
        e = cudaFuncSetAttribute(k_chain_pass, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int tiles = (a.rows + CH_TR - 1) / CH_TR;
    ChainArgs b = a;
    b.dbg = tc_debug_get_times();
    k_chain_pass<<<tiles, CH_THREADS, CH_SMEM_PASS, st>>>(b);
    return cudaGetLastError();
}

}  // namespace cb
