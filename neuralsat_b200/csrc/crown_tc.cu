// Dense contractions of the CROWN pass on the 5th-generation tensor cores (sm_100a).
//
//   D[128 x BN] (fp32, TMEM) = X[128 x K] . B[BN x K]^T      tcgen05.mma.kind::f16 on bf16 triples
//
// One CTA owns a 128-row tile of sub-domain rows (row r = s*Bd + b) and a BN-column tile of the
// layer; both operands arrive in shared memory by bulk TMA (cp.async.bulk, mbarrier tx-count)
// already split into three bf16 planes and already laid out as UMMA canonical K-major no-swizzle
// tiles, so the mainloop is: 1 producer thread (TMA), 1 MMA thread, 4-stage ring.
//
// fp32 fidelity (the north star asks for bounds within 1e-5 relative, and the 20-step Adam
// trajectory amplifies contraction noise ~100x): x = x1 + x2 + x3 with x1 = bf16(x),
// x2 = bf16(x - x1), x3 = bf16(x - x1 - x2) represents every fp32 value to 2^-27, and
//   x*w ~= x3*w1 + x2*w2 + x1*w3 + x2*w1 + x1*w2 + x1*w1      (dropped terms <= 2^-26 |x||w|)
// is six bf16 MMAs per 16 k-values = the tensor-pipe time of three tf32 MMAs per 8, at 6 instead
// of 8 operand bytes per value.  (3xTF32 was measured first: 5e-7 relative contraction error,
// which the optimiser loop amplified to 6e-5 on the bounds; the bf16 triple is at fp32 level.)
//
// The epilogue (8 warps, thread = one row, TMEM lane = row) is where the CROWN work happens, so
// the coefficient matrix A never makes an extra HBM round trip between a Linear and the node below:
//   MODE_RELAX       pass through "Linear then ReLU": stores lA (= D), applies the ReLU relaxation
//                    (operators/relu.py:456-494) + sign-split multiply (operators/clampmult.py:17-43)
//                    + beta injection of the pre-activation node (beta_crown.py:163-204) + the bias
//                    dot product of the Linear below (operators/linear.py:167-175), and writes the
//                    next layer's X operand (packed planes) and/or a plain fp32 matrix.
//   MODE_CONCRETIZE  first layer: lb += D.c - |D|.d (perturbations.py:154-183) and the gradient
//                    seed g0 = c - sign(D) d.
//   MODE_GRAD        forward direction of the alpha/beta gradient: g_pre = G.W^T + b, then the
//                    hand-written backward of the sign-split multiply (clampmult.py:49-95):
//                    grad_alpha, grad_beta, g_post.
//   MODE_STORE       plain GEMM (+ column bias), used by the self-test.
//
// Packed operand format, shared by activations (TR = 128) and weights (TR = BN): for M[R, K]
//   buf[r / TR][k / 16][plane 0..2][(k / 8) % 2][r % TR][k % 8]      (bf16)
// so the three planes of one [TR x 16] k-step are ONE contiguous block (one bulk copy per
// operand per stage) and each plane is a UMMA K-major SWIZZLE_NONE tile with
// LBO = TR*16 B (next 8 k-values) and SBO = 128 B (next 8 rows).
#include <cuda_bf16.h>

#include "crown_kernels.cuh"
#include "crown_tc_common.cuh"

namespace cb {

namespace {

using namespace tcc;

__device__ __forceinline__ void epi_bar() {      // the epilogue warps only
    asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_linear(const __grid_constant__ TcArgs a) {
    if (a.done != nullptr && *a.done != 0) return;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[TC_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x, m_tile = blockIdx.y;
    long long* const tdbg = a.dbg_times ? a.dbg_times + 8 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
    if (tdbg && threadIdx.x == 0) tdbg[0] = clock64();
    const int BN = a.BN;
    const int b_plane = BN * TC_BK * 2;
    const int stage_bytes = TC_A_STAGE + 3 * b_plane;
    const int num_kb = a.Kp / TC_BK;
    // Staging area of the epilogue: the [128 x BN] tiles of its per-row operands (l, u, alpha,
    // a_post / x_L, x_U) are fetched with coalesced cp.async while the MMA mainloop runs, and its
    // per-row results (lA, grad_alpha, plain output) leave through the same tiles with coalesced
    // 16-byte stores: thread = row in the arithmetic, warp = 2 contiguous rows on the memory side.
    const int n_stages = a.stages;
    uint8_t* const ew = smem + (size_t)n_stages * stage_bytes;
    const int ew_pitch = BN * 4 + 16;                  // +16 B: conflict-free 128-bit accesses, thread = row
    const int ew_arr = TC_BM * ew_pitch;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (tdbg && threadIdx.x == 0) tdbg[1] = clock64();

    if (warp == 0) {
        // ===== TMA producer (one thread) =====
        if (lane == 0) {
            // element offsets: one k-step of a tile = 3 planes x TR x 16 bf16, contiguous
            const uint16_t* a_src = a.xp + (size_t)m_tile * (a.Kp >> 4) * (3 * TC_BM * TC_BK);
            const uint16_t* b_src = a.wp + (size_t)n_tile * (a.Kp >> 4) * ((size_t)3 * BN * TC_BK);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % n_stages;
                const uint32_t ph = (uint32_t)(kb / n_stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                bulk_g2s(st, a_src + (size_t)kb * (3 * TC_BM * TC_BK), TC_A_STAGE, &full_bar[s]);
                bulk_g2s(st + TC_A_STAGE, b_src + (size_t)kb * ((size_t)3 * BN * TC_BK), (uint32_t)(3 * b_plane),
                         &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(BN);
            const uint32_t a_lbo = TC_BM * 16, b_lbo = (uint32_t)BN * 16;
            const bool swap = (a.dbg & 1) != 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % n_stages;
                const uint32_t ph = (uint32_t)(kb / n_stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                if (tdbg && kb == 0) tdbg[2] = clock64();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t sb = sa + TC_A_STAGE;
                uint64_t ad[3], bd[3];
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    ad[pl] = umma_desc(sa + pl * TC_A_PLANE, swap ? 128 : a_lbo, swap ? a_lbo : 128);
                    bd[pl] = umma_desc(sb + pl * b_plane, swap ? 128 : b_lbo, swap ? b_lbo : 128);
                }
                // The tensor core adds each MMA into the fp32 accumulator with truncation (measured:
                // error ~ #MMAs * 2^-25 |acc|), so the five correction terms (<= 2^-8 of the product)
                // get their own accumulator: the main one sees K/16 additions instead of 6K/16 and
                // the corrections are truncated relative to their own small magnitude.
                const uint32_t d_small = tmem_base + TC_SMALL_COL;
                umma_bf16(d_small, ad[2], bd[0], idesc, kb ? 1u : 0u);
                umma_bf16(d_small, ad[1], bd[1], idesc, 1u);
                umma_bf16(d_small, ad[0], bd[2], idesc, 1u);
                umma_bf16(d_small, ad[1], bd[0], idesc, 1u);
                umma_bf16(d_small, ad[0], bd[1], idesc, 1u);
                umma_bf16(tmem_base, ad[0], bd[0], idesc, kb ? 1u : 0u);
                umma_commit(&empty_bar[s]);       // frees the smem slot when these MMAs retire
            }
            umma_commit(&tmem_full_bar);          // accumulator complete
            if (tdbg) tdbg[3] = clock64();
        }
    } else {
        // ===== epilogue: TMEM lane quarter = warp % 4, column group = (warp - 2) / 4 =====
        const int te = threadIdx.x - 64;
        const int quarter = warp & 3;
        const int cgrp = (warp - 2) >> 2;
        const int row_local = quarter * 32 + lane;
        const int grow = m_tile * TC_BM + row_local;
        const bool valid = grow < a.rows;
        const int b = valid ? grow % a.Bd : 0;
        const int s_idx = valid ? grow / a.Bd : 0;
        const int half = BN / TC_CGROUPS;                    // columns per thread (a multiple of 8)
        const int col0 = n_tile * BN;
        const int cbeg = col0 + cgrp * half;                 // first global column of this thread
        const int N = a.N;
        const bool vecN = (N & 3) == 0;
        const bool has_alpha = a.alpha != nullptr;
        const bool staged = a.ew_stage != 0 && MODE != TC_MODE_STORE;
        const bool st_alpha = staged && has_alpha && a.alpha_pos == nullptr;
        const bool ga_add = (a.S1 == 1 && a.S > 1);          // several rows accumulate into one alpha row
        const int rows_valid = min(TC_BM, a.rows - m_tile * TC_BM);
        const int ncols = min(BN, N - col0);                 // valid columns of this tile
        const int bn4 = BN >> 2;

        // ---- 1. cooperative, coalesced prefetch of the per-row operands (overlaps the mainloop) ----
        if (staged) {
            for (int idx = te; idx < TC_BM * bn4; idx += TC_EPI_THREADS) {
                const int r = idx / bn4, c4 = idx - r * bn4;
                if (r >= rows_valid || c4 * 4 >= ncols) continue;
                const int gr = m_tile * TC_BM + r;
                const int rb = gr % a.Bd, rs = gr / a.Bd;
                uint8_t* dst = ew + (size_t)r * ew_pitch + c4 * 16;
                const size_t coff = (size_t)col0 + c4 * 4;
                if (MODE == TC_MODE_CONCRETIZE) {
                    cp_async16(dst, a.x_L + (size_t)rb * N + coff);
                    cp_async16(dst + ew_arr, a.x_U + (size_t)rb * N + coff);
                } else {
                    cp_async16(dst, a.lower + (size_t)rb * N + coff);
                    cp_async16(dst + ew_arr, a.upper + (size_t)rb * N + coff);
                    if (st_alpha)
                        cp_async16(dst + 2 * ew_arr,
                                   a.alpha + ((size_t)(a.S1 == 1 ? 0 : rs) * a.Bd + rb) * a.n_alpha + coff);
                    if (MODE == TC_MODE_GRAD) cp_async16(dst + 3 * ew_arr, a.a_post + (size_t)gr * N + coff);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }

        // per-row pointers of the direct (unstaged) path
        const float* lrow = nullptr; const float* urow = nullptr; const float* alrow = nullptr;
        const float* aprow = nullptr; float* garow = nullptr; float* larow = nullptr; float* yrow = nullptr;
        const float* xlrow = nullptr; const float* xurow = nullptr;
        if (MODE == TC_MODE_RELAX || MODE == TC_MODE_GRAD) {
            lrow = a.lower + (size_t)b * N;
            urow = a.upper + (size_t)b * N;
            const size_t arow = ((size_t)(a.S1 == 1 ? 0 : s_idx) * a.Bd + b) * a.n_alpha;
            if (has_alpha) alrow = a.alpha + arow;
            if (MODE == TC_MODE_GRAD) {
                aprow = a.a_post + (size_t)grow * N;
                if (a.grad_alpha && has_alpha) garow = a.grad_alpha + arow;
            } else if (a.lA) {
                larow = a.lA + (size_t)grow * N;
            }
        }
        if (MODE == TC_MODE_CONCRETIZE) {
            xlrow = a.x_L + (size_t)b * N;
            xurow = a.x_U + (size_t)b * N;
        }
        if (a.y_plain) yrow = a.y_plain + (size_t)grow * N;
        const bool vecA = vecN && (a.alpha_pos == nullptr) && ((a.n_alpha & 3) == 0);
        // which results leave through the staging tiles (slot): lA -> 0, plain y -> 1 (0 for
        // CONCRETIZE), grad_alpha -> 2
        const bool out_lA = staged && MODE == TC_MODE_RELAX && a.lA != nullptr;
        const bool out_y = staged && a.y_plain != nullptr;
        const bool out_ga = st_alpha && MODE == TC_MODE_GRAD && a.grad_alpha != nullptr && !ga_add;
        const int y_slot = (MODE == TC_MODE_CONCRETIZE) ? 0 : 1;

        // beta entries of this row that fall into this thread's columns (<= 32 columns)
        unsigned long long bmask = 0ull;
        const int J = (MODE == TC_MODE_RELAX || MODE == TC_MODE_GRAD) ? a.J : 0;
        float acc = 0.f;
        if (J > 0 && valid) {
            const size_t jb = (size_t)b * J;
            for (int j = 0; j < J; ++j) {
                const float sg = __ldg(a.beta_sign + jb + j);
                const float key = (MODE == TC_MODE_RELAX) ? sg * __ldg(a.beta_val + jb + j) : sg;
                if (key == 0.f) continue;
                const long long lc = __ldg(a.beta_loc + jb + j);
                if (lc >= cbeg && lc < cbeg + half) bmask |= 1ull << (int)(lc - cbeg);
                if (MODE == TC_MODE_RELAX && a.beta_bias && n_tile == 0 && cgrp == 0)
                    acc = fmaf(key, __ldg(a.beta_bias + jb + j), acc);
            }
        }

        uint8_t* const ew_row = ew + (size_t)row_local * ew_pitch;
        // 8 values of staged slot `arr` at tile column ccol (tile columns >= ncols were not loaded)
        auto ew8 = [&](int arr, int ccol, int gcol, float (&o)[8]) {
            const float4 p = *reinterpret_cast<const float4*>(ew_row + (size_t)arr * ew_arr + ccol * 4);
            const float4 q = *reinterpret_cast<const float4*>(ew_row + (size_t)arr * ew_arr + ccol * 4 + 16);
            o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.w; o[4] = q.x; o[5] = q.y; o[6] = q.z; o[7] = q.w;
            if (gcol + 8 > N) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (gcol + i >= N) o[i] = 0.f;
            }
        };
        auto ew_put8 = [&](int arr, int ccol, const float (&v)[8]) {
            *reinterpret_cast<float4*>(ew_row + (size_t)arr * ew_arr + ccol * 4) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(ew_row + (size_t)arr * ew_arr + ccol * 4 + 16) =
                make_float4(v[4], v[5], v[6], v[7]);
        };
        if (staged) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            epi_bar();
        }
        if (tdbg && te == 0) tdbg[4] = clock64();
        mbar_wait(&tmem_full_bar, 0);
        if (tdbg && te == 0) tdbg[5] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);

        // ---- 2. arithmetic: thread = one row, 8 columns at a time ----
        for (int cl = 0; cl < half; cl += 8) {
            const int ccol = cgrp * half + cl;               // column inside the tile
            const int gcol = col0 + ccol;                    // global column
            float d[8];
            {
                float ds[8];
                tmem_ld8(trow + (uint32_t)(TC_SMALL_COL + ccol), ds);
                tmem_ld8(trow + (uint32_t)ccol, d);
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] += ds[i];
            }
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = 0.f;
            if (valid && gcol < N) {
                if (MODE == TC_MODE_STORE) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        y[i] = d[i] + ((a.col_bias && gcol + i < N) ? __ldg(a.col_bias + gcol + i) : 0.f);
                } else if (MODE == TC_MODE_CONCRETIZE) {
                    float xl[8], xu[8];
                    if (staged) {
                        ew8(0, ccol, gcol, xl);
                        ew8(1, ccol, gcol, xu);
                    } else {
                        load8(xlrow, gcol, N, vecN, xl, 0.f);
                        load8(xurow, gcol, N, vecN, xu, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float cen = (xu[i] + xl[i]) / 2.0f, dif = (xu[i] - xl[i]) / 2.0f;
                        const float av = (gcol + i < N) ? d[i] : 0.f;
                        acc += av * cen - fabsf(av) * dif;
                        const float sg = (av > 0.f) ? 1.f : ((av < 0.f) ? -1.f : 0.f);
                        y[i] = (gcol + i < N) ? (cen - sg * dif) : 0.f;
                    }
                } else {
                    float l[8], u[8], al[8];
                    if (staged) {
                        ew8(0, ccol, gcol, l);
                        ew8(1, ccol, gcol, u);
                    } else {
                        load8(lrow, gcol, N, vecN, l, 0.f);
                        load8(urow, gcol, N, vecN, u, 0.f);
                    }
                    int pos[8];
                    if (has_alpha) {
                        if (a.alpha_pos == nullptr) {
                            if (st_alpha) ew8(2, ccol, gcol, al);
                            else load8(alrow, gcol, N, vecA, al, 0.f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) pos[i] = gcol + i;
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                pos[i] = (gcol + i < N) ? __ldg(a.alpha_pos + gcol + i) : -1;
                                al[i] = pos[i] >= 0 ? __ldg(alrow + pos[i]) : 0.f;
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { al[i] = 0.f; pos[i] = -1; }
                    }
                    if (MODE == TC_MODE_RELAX) {
                        if (out_lA) ew_put8(0, ccol, d);
                        else if (larow) store8(larow, gcol, N, vecN, d);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const Relax8 rx = relax1(l[i], u[i], has_alpha, al[i]);
                            const float a_pos = fmaxf(d[i], 0.f), a_neg = fminf(d[i], 0.f);
                            y[i] = (gcol + i < N) ? (rx.d_l * a_pos + rx.d_u * a_neg) : 0.f;
                            acc = fmaf(a_neg, rx.b_u, acc);
                        }
                        const unsigned hit = (unsigned)(bmask >> cl) & 0xffu;
                        if (hit) {
                            const size_t jb = (size_t)b * J;
                            for (int j = 0; j < J; ++j) {
                                const int rel = (int)(__ldg(a.beta_loc + jb + j) - gcol);
                                if (rel < 0 || rel >= 8) continue;
                                const float vs = __ldg(a.beta_val + jb + j) * __ldg(a.beta_sign + jb + j);
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (i == rel) y[i] -= vs;
                            }
                        }
                        if (a.blin) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (gcol + i < N) acc = fmaf(y[i], __ldg(a.blin + gcol + i), acc);
                        }
                    } else {   // TC_MODE_GRAD
                        float ap[8], g[8], ga[8];
                        if (staged) ew8(3, ccol, gcol, ap);
                        else load8(aprow, gcol, N, vecN, ap, 0.f);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            g[i] = d[i] + ((a.col_bias && gcol + i < N) ? __ldg(a.col_bias + gcol + i) : 0.f);
                            const Relax8 rx = relax1(l[i], u[i], has_alpha, al[i]);
                            y[i] = (gcol + i < N)
                                       ? (g[i] * (ap[i] >= 0.f ? rx.d_l : rx.d_u) + (ap[i] < 0.f ? rx.b_u : 0.f))
                                       : 0.f;
                            ga[i] = (rx.live && ap[i] >= 0.f) ? g[i] * ap[i] : 0.f;
                        }
                        if (out_ga) {
                            ew_put8(2, ccol, ga);
                        } else if (garow) {
                            if (a.alpha_pos == nullptr && !ga_add) {
                                store8(garow, gcol, N, vecA, ga);
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    if (pos[i] < 0 || gcol + i >= N) continue;
                                    if (ga_add) atomicAdd(garow + pos[i], ga[i]);
                                    else garow[pos[i]] = ga[i];
                                }
                            }
                        }
                        const unsigned hit = (unsigned)(bmask >> cl) & 0xffu;
                        if (hit && a.grad_beta) {
                            const size_t jb = (size_t)b * J;
                            for (int j = 0; j < J; ++j) {
                                const float sg = __ldg(a.beta_sign + jb + j);
                                if (sg == 0.f) continue;
                                const int rel = (int)(__ldg(a.beta_loc + jb + j) - gcol);
                                if (rel < 0 || rel >= 8) continue;
                                float gv = 0.f;
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (i == rel) gv = g[i];
                                float gb = -sg * gv;
                                if (a.beta_bias) gb = fmaf(sg, __ldg(a.beta_bias + jb + j), gb);
                                if (a.S > 1) atomicAdd(a.grad_beta + jb + j, gb);
                                else a.grad_beta[jb + j] = gb;
                            }
                        }
                    }
                }
                if (out_y) ew_put8(y_slot, ccol, y);
                else if (yrow) store8(yrow, gcol, N, vecN, y);
            }
            if (a.yp && gcol < a.y_Kp) store_packed8(a.yp, m_tile, row_local, gcol, a.y_Kp, y);
        }
        if ((MODE == TC_MODE_RELAX || MODE == TC_MODE_CONCRETIZE) && valid && a.bias_rows)
            atomicAdd(a.bias_rows + grow, acc);
        if (tdbg && te == 0) tdbg[6] = clock64();

        // ---- 3. cooperative, coalesced copy-out of the per-row results ----
        if (out_lA || out_y || out_ga) {
            epi_bar();
            for (int idx = te; idx < TC_BM * bn4; idx += TC_EPI_THREADS) {
                const int r = idx / bn4, c4 = idx - r * bn4;
                if (r >= rows_valid || c4 * 4 >= ncols) continue;
                const int gr = m_tile * TC_BM + r;
                const uint8_t* src = ew + (size_t)r * ew_pitch + c4 * 16;
                const size_t coff = (size_t)col0 + c4 * 4;
                if (out_lA)
                    *reinterpret_cast<float4*>(a.lA + (size_t)gr * N + coff) = *reinterpret_cast<const float4*>(src);
                if (out_y)
                    *reinterpret_cast<float4*>(a.y_plain + (size_t)gr * N + coff) =
                        *reinterpret_cast<const float4*>(src + (size_t)y_slot * ew_arr);
                if (out_ga) {
                    const int rb = gr % a.Bd, rs = gr / a.Bd;
                    *reinterpret_cast<float4*>(a.grad_alpha + ((size_t)(a.S1 == 1 ? 0 : rs) * a.Bd + rb) * a.n_alpha + coff) =
                        *reinterpret_cast<const float4*>(src + 2 * (size_t)ew_arr);
                }
            }
        }
        if (tdbg && te == 0) tdbg[7] = clock64();
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                     : "memory");
    }
}

// W (element (n,k) at W[n*sn + k*sk]) -> packed planes, TR = BN.  One thread per (n, 8 k-values).
__global__ void k_pack_weight(const float* __restrict__ W, long long sn, long long sk, int N, int K, int Kp,
                              int BN, int n_tiles, uint16_t* __restrict__ out) {
    const size_t total = (size_t)n_tiles * BN * (Kp >> 3);
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int nl = (int)(t % BN);
        const int kg = (int)((t / BN) % (Kp >> 3));
        const int nt = (int)(t / ((size_t)BN * (Kp >> 3)));
        const int n = nt * BN + nl;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = kg * 8 + i;
            v[i] = (n < N && k < K) ? W[(size_t)n * sn + (size_t)k * sk] : 0.f;
        }
        uint4 p1, p2, p3;
        pack8(v, p1, p2, p3);
        *reinterpret_cast<uint4*>(out + packed_off(nt, BN, Kp, kg * 8, 0, nl)) = p1;
        *reinterpret_cast<uint4*>(out + packed_off(nt, BN, Kp, kg * 8, 1, nl)) = p2;
        *reinterpret_cast<uint4*>(out + packed_off(nt, BN, Kp, kg * 8, 2, nl)) = p3;
    }
}

// Row-major X[rows,K] (or the spec matrix C[Bd,S,K], row = s*Bd + b) -> packed planes for Mp rows.
__global__ void k_pack_rows(const float* __restrict__ src, int spec_layout, int rows, int Bd, int S, int K, int Kp,
                            int m_tiles, uint16_t* __restrict__ out, const float* __restrict__ rowdot_vec,
                            float* __restrict__ bias_rows, const int* done) {
    if (done != nullptr && *done != 0) return;
    const size_t total = (size_t)m_tiles * TC_BM * (Kp >> 3);
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int rl = (int)(t % TC_BM);
        const int kg = (int)((t / TC_BM) % (Kp >> 3));
        const int mt = (int)(t / ((size_t)TC_BM * (Kp >> 3)));
        const int r = mt * TC_BM + rl;
        const float* row = nullptr;
        if (r < rows) {
            if (spec_layout) {
                const int b = r % Bd, s = r / Bd;
                row = src + ((size_t)b * S + s) * K;
            } else {
                row = src + (size_t)r * K;
            }
        }
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = kg * 8 + i;
            v[i] = (row && k < K) ? row[k] : 0.f;
        }
        uint4 p1, p2, p3;
        pack8(v, p1, p2, p3);
        *reinterpret_cast<uint4*>(out + packed_off(mt, TC_BM, Kp, kg * 8, 0, rl)) = p1;
        *reinterpret_cast<uint4*>(out + packed_off(mt, TC_BM, Kp, kg * 8, 1, rl)) = p2;
        *reinterpret_cast<uint4*>(out + packed_off(mt, TC_BM, Kp, kg * 8, 2, rl)) = p3;
        if (rowdot_vec && kg == 0 && row) {
            float t2 = 0.f;
            for (int k = 0; k < K; ++k) t2 = fmaf(row[k], __ldg(rowdot_vec + k), t2);
            bias_rows[r] += t2;
        }
    }
}

__global__ void k_rows_to_lb(const float* __restrict__ bias_rows, float* __restrict__ lb, int Bd, int S,
                             const int* done) {
    if (done != nullptr && *done != 0) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Bd * S) return;
    const int b = r % Bd, s = r / Bd;
    lb[(size_t)b * S + s] = bias_rows[r];
}

template <int MODE>
cudaError_t launch_tc(const TcArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_linear<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             TC_SMEM_MAX);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    k_tc_linear<MODE><<<grid, TC_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

int tc_pick_bn(int N, int cap) {
    if (N <= cap) return (N + 31) / 32 * 32;
    // operand bytes per k-step of one 128-row tile ~ n_tiles * (128 + bn): fewer, wider tiles re-read X less
    int best = cap;
    long best_cost = -1;
    for (int bn = cap; bn >= 32; bn -= 32) {
        const long cost = (long)((N + bn - 1) / bn) * (128 + bn);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

void tc_pack_weight(const float* W, long long sn, long long sk, int N, int K, int Kp, int BN, uint16_t* out,
                    cudaStream_t st) {
    const int n_tiles = (N + BN - 1) / BN;
    const size_t total = (size_t)n_tiles * BN * (Kp >> 3);
    const unsigned blocks = (unsigned)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    k_pack_weight<<<blocks, 256, 0, st>>>(W, sn, sk, N, K, Kp, BN, n_tiles, out);
}

void tc_pack_rows(const float* src, bool spec_layout, int rows, int Bd, int S, int K, int Kp, uint16_t* out,
                  const float* rowdot_vec, float* bias_rows, const int* done, cudaStream_t st) {
    Launch _l(K_TC_PACK, st);
    const int m_tiles = (rows + TC_BM - 1) / TC_BM;
    const size_t total = (size_t)m_tiles * TC_BM * (Kp >> 3);
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_pack_rows<<<blocks, 256, 0, st>>>(src, spec_layout ? 1 : 0, rows, Bd, S, K, Kp, m_tiles, out, rowdot_vec,
                                        bias_rows, done);
}

void rows_to_lb(const float* bias_rows, float* lb, int Bd, int S, const int* done, cudaStream_t st) {
    Launch _l(K_CONCRETIZE, st);
    k_rows_to_lb<<<(Bd * S + 255) / 256, 256, 0, st>>>(bias_rows, lb, Bd, S, done);
}

static long long* g_dbg_times = nullptr;
void tc_debug_set_times(long long* p) { g_dbg_times = p; }
long long* tc_debug_get_times() { return g_dbg_times; }

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

cudaError_t tc_linear(int mode, const TcArgs& a_in, cudaStream_t st) {
    Launch _l(K_TC_LINEAR, st);
    TcArgs a = a_in;
    a.dbg_times = g_dbg_times;
    const int m_tiles = (a.rows + TC_BM - 1) / TC_BM;
    const int n_tiles = (a.N + a.BN - 1) / a.BN;
    dim3 grid(n_tiles, m_tiles);
    const size_t stage_bytes = TC_A_STAGE + 3 * (size_t)a.BN * TC_BK * 2;
    a.stages = TC_STAGES;
    size_t smem = (size_t)a.stages * stage_bytes;
    // stage the per-row epilogue operands through shared memory when the rows are 16-byte
    // sliceable and everything fits; otherwise the epilogue reads them straight from global
    a.ew_stage = 0;
    if (mode != TC_MODE_STORE && (a.N & 3) == 0) {
        const bool dense_alpha = a.alpha != nullptr && a.alpha_pos == nullptr;
        int n_arr = 2;
        bool ok = true;
        if (mode == TC_MODE_CONCRETIZE) {
            ok = aligned16(a.x_L) && aligned16(a.x_U);
        } else {
            ok = aligned16(a.lower) && aligned16(a.upper);
            if (dense_alpha) {
                ok = ok && aligned16(a.alpha) && a.n_alpha == a.N;
                n_arr = 3;
            }
            if (mode == TC_MODE_GRAD) {
                ok = ok && aligned16(a.a_post);
                n_arr = 4;           // slot 3 is a_post even when alpha is not staged
            }
        }
        const size_t ew_bytes = (size_t)n_arr * TC_BM * ((size_t)a.BN * 4 + 16);
        int stages = TC_STAGES;
        while (stages > 2 && (size_t)stages * stage_bytes + ew_bytes > (size_t)TC_SMEM_MAX) --stages;
        if (ok && (size_t)stages * stage_bytes + ew_bytes <= (size_t)TC_SMEM_MAX) {
            a.ew_stage = 1;
            a.stages = stages;
            smem = (size_t)stages * stage_bytes + ew_bytes;
        }
    }
    switch (mode) {
        case TC_MODE_STORE: return launch_tc<TC_MODE_STORE>(a, grid, smem, st);
        case TC_MODE_RELAX: return launch_tc<TC_MODE_RELAX>(a, grid, smem, st);
        case TC_MODE_CONCRETIZE: return launch_tc<TC_MODE_CONCRETIZE>(a, grid, smem, st);
        case TC_MODE_GRAD: return launch_tc<TC_MODE_GRAD>(a, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cb
