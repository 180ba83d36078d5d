// Device-resident domain store and branching for the BaB loop: the per-domain records
// (intermediate bounds, fp16 slopes, lA, split histories, beta) stay in HBM between iterations; only counts cross
// PCIe.  What the reference does on the host with Python lists and pinned TensorStorage
// (heuristic/domains_list.py:153-311, abstractor/utils.py:159-250, heuristic/util.py:31-72,
// heuristic/decision_heuristics.py:78-251) is a handful of HBM-bound kernels here:
//
//   k_multi_copy     one launch moves EVERY tensor of a record set: row gather (pick the parents of the children),
//                    row scatter to ranked slots (append the surviving children), with the fp16 <-> fp32 conversion of
//                    the slopes (abstractor/utils.py:51-59), int32 -> int64 of the split locations and the
//                    [S,Bd,n] -> [Bd,S,n] transpose of lA fused into the copy
//   k_apply_split    child r of parent p: lower[layer][r, n] = point (active side) or upper[...] = point (inactive side)
//                    and the new (loc, sign, beta = 0) history entry (abstractor/utils.py:159-178, :214-250)
//   k_keep_rank      keep = all_s(lb <= rhs) (domains_list.py:246), exclusive scan -> slot of every survivor, count,
//                    per-layer maximum history length of the survivors
//   k_babsr          BaBSR score and intercept ("backup") score of every neuron of a layer (heuristic/util.py:31-72)
//   k_topk_rows      k largest / smallest entries of every row, ties to the lowest index
//   k_pick_decision  arg-max over the k look-ahead passes, score candidate vs backup candidate per parent, fallback
//                    to the first unstable neuron (decision_heuristics.py:159-251)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/crown_b200.h"
#include "crown_kernels.cuh"

namespace cb {

namespace {

__global__ void k_multi_copy(const cb_copy_desc_t* __restrict__ descs, const int32_t* __restrict__ src_map,
                             const int32_t* __restrict__ dst_map, int R) {
    const cb_copy_desc_t d = descs[blockIdx.y];
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        const long long sr = src_map ? src_map[r] : r;
        const long long dr = dst_map ? dst_map[r] : r;
        if (sr < 0 || dr < 0) continue;
        const int rep = d.src_S > 0 ? d.src_S : 1;               // lA: [S,Bd,n] -> row (b) of [Bd,S,n]
        for (int s = 0; s < rep; ++s) {
            const size_t so = d.src_S > 0 ? ((size_t)s * d.src_Bd + sr) * d.width : (size_t)sr * d.src_stride;
            const size_t dofs = (size_t)dr * d.dst_stride + (size_t)s * d.width;
            switch (d.mode) {
                case CB_COPY_F32: {
                    const float* sp = static_cast<const float*>(d.src) + so;
                    float* dp = static_cast<float*>(d.dst) + dofs;
                    if ((d.width & 3) == 0 && ((reinterpret_cast<uintptr_t>(sp) | reinterpret_cast<uintptr_t>(dp)) & 15u) == 0) {
                        for (int i = threadIdx.x; i < (d.width >> 2); i += blockDim.x)
                            reinterpret_cast<float4*>(dp)[i] = reinterpret_cast<const float4*>(sp)[i];
                    } else {
                        for (int i = threadIdx.x; i < d.width; i += blockDim.x) dp[i] = sp[i];
                    }
                    for (int i = d.width + threadIdx.x; i < d.dst_width; i += blockDim.x) dp[i] = 0.f;
                    break;
                }
                case CB_COPY_F16_TO_F32: {
                    const __half* sp = static_cast<const __half*>(d.src) + so;
                    float* dp = static_cast<float*>(d.dst) + dofs;
                    for (int i = threadIdx.x; i < d.width; i += blockDim.x) dp[i] = __half2float(sp[i]);
                    break;
                }
                case CB_COPY_F32_TO_F16: {
                    const float* sp = static_cast<const float*>(d.src) + so;
                    __half* dp = static_cast<__half*>(d.dst) + dofs;
                    for (int i = threadIdx.x; i < d.width; i += blockDim.x) dp[i] = __float2half_rn(sp[i]);
                    break;
                }
                case CB_COPY_I32: {
                    const int32_t* sp = static_cast<const int32_t*>(d.src) + so;
                    int32_t* dp = static_cast<int32_t*>(d.dst) + dofs;
                    for (int i = threadIdx.x; i < d.width; i += blockDim.x) dp[i] = sp[i];
                    for (int i = d.width + threadIdx.x; i < d.dst_width; i += blockDim.x) dp[i] = 0;
                    break;
                }
                case CB_COPY_I32_TO_I64: {
                    const int32_t* sp = static_cast<const int32_t*>(d.src) + so;
                    int64_t* dp = static_cast<int64_t*>(d.dst) + dofs;
                    for (int i = threadIdx.x; i < d.width; i += blockDim.x) dp[i] = sp[i];
                    for (int i = d.width + threadIdx.x; i < d.dst_width; i += blockDim.x) dp[i] = 0;
                    break;
                }
                case CB_COPY_I64_TO_I32: {
                    const int64_t* sp = static_cast<const int64_t*>(d.src) + so;
                    int32_t* dp = static_cast<int32_t*>(d.dst) + dofs;
                    for (int i = threadIdx.x; i < d.width; i += blockDim.x) dp[i] = (int32_t)sp[i];
                    for (int i = d.width + threadIdx.x; i < d.dst_width; i += blockDim.x) dp[i] = 0;
                    break;
                }
                default: break;
            }
        }
    }
}

// One thread per child row: the split and its history entry.
__global__ void k_apply_split(cb_split_layer_t* __restrict__ layers, int n_layers, const int32_t* __restrict__ dec_layer,
                              const int32_t* __restrict__ dec_neuron, const float* __restrict__ dec_side,
                              const float* __restrict__ dec_point, int R) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int k = dec_layer[r];
    if (k < 0 || k >= n_layers) return;
    const cb_split_layer_t L = layers[k];
    const int n = dec_neuron[r];
    const float side = dec_side[r];
    const float pt = dec_point ? dec_point[r] : 0.f;
    if (side > 0.f) L.lower[(size_t)r * L.n + n] = pt;           // first half: x >= point
    else L.upper[(size_t)r * L.n + n] = pt;                      // second half: x <= point
    if (L.hist_cnt != nullptr) {
        const int c = L.hist_cnt[r];
        if (c < L.J) {
            L.hist_loc[(size_t)r * L.J + c] = n;
            L.hist_sign[(size_t)r * L.J + c] = side;
            L.beta_val[(size_t)r * L.J + c] = 0.f;
            if (L.hist_point) L.hist_point[(size_t)r * L.J + c] = pt;
        }
        L.hist_cnt[r] = c + 1;
    }
}

// keep[r] = all_s(lb[r,s] <= rhs[r,s]); rank[r] = base + #kept before r (or -1); out[0] = #kept,
// out[1 + k] = max(out[1 + k], history length of layer k over the kept rows).  One CTA.
__global__ void k_keep_rank(const float* __restrict__ lb, const float* __restrict__ rhs, int R, int S, int base,
                            int32_t* __restrict__ rank, int32_t* __restrict__ out, const int32_t* const* __restrict__ hist_cnt,
                            int n_layers) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int r0 = 0; r0 < R; r0 += blockDim.x) {
        const int r = r0 + tid;
        int keep = 0;
        if (r < R) {
            keep = 1;
            for (int s = 0; s < S; ++s)
                if (!(lb[(size_t)r * S + s] <= rhs[(size_t)r * S + s])) keep = 0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int wpre = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (r < R) rank[r] = keep ? base + off + wpre : -1;
        if (keep)
            for (int k = 0; k < n_layers; ++k)
                if (hist_cnt[k]) atomicMax(out + 1 + k, hist_cnt[k][r]);
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < nw; ++w) t += s_warp[w];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) out[0] = s_base;
}

// heuristic/util.py:31-72 for one layer; lA [B,S,n], lower/upper [B,n], bias [n] (already broadcast) or null
__global__ void k_babsr(const float* __restrict__ lA, const float* __restrict__ lower, const float* __restrict__ upper,
                        const float* __restrict__ bias, int B, int S, int n, float* __restrict__ score,
                        float* __restrict__ backup, float* __restrict__ mask_out, int ld, int col0) {
    const size_t total = (size_t)B * n;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / n), i = (int)(t - (size_t)b * n);
        const float l = lower[t], u = upper[t];
        const float mask = (l < 0.f && u > 0.f) ? 1.f : 0.f;
        const float lt = fminf(l, 0.f), ut = fmaxf(u, 0.f);
        const float r0 = __fdiv_rn(ut, ut - lt);                 // _compute_ratio: no epsilon guard in the reference
        const float r1 = -1.f * lt * r0;
        const float bi = bias ? bias[i] : 0.f;
        float sc = 0.f, bk = 0.f;
        for (int s = 0; s < S; ++s) {
            const float a = lA[((size_t)b * S + s) * n + i];
            const float ic = fminf(a, 0.f) * r1;
            const float bt = bi * a;
            const float c1 = bt * (r0 - 1.f), c2 = bt * r0;
            sc += fabsf(fmaxf(c1, c2) + ic) * mask;
            bk += ic * mask;
        }
        score[(size_t)b * ld + col0 + i] = sc / (float)S;
        backup[(size_t)b * ld + col0 + i] = bk / (float)S;
        if (mask_out) mask_out[(size_t)b * ld + col0 + i] = mask;
    }
}

// one CTA per row: k rounds of arg-max (largest) / arg-min; ties -> lowest index; NaN never wins
__global__ void k_topk_rows(const float* __restrict__ x, int n, int k, int largest, float* __restrict__ vals,
                            int32_t* __restrict__ idx) {
    extern __shared__ unsigned char s_raw[];
    float* s_val = reinterpret_cast<float*>(s_raw);
    int* s_idx = reinterpret_cast<int*>(s_raw + 32 * sizeof(float));
    int* s_taken = reinterpret_cast<int*>(s_raw + 64 * sizeof(float));          // [k]
    const float* row = x + (size_t)blockIdx.x * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int j = 0; j < k; ++j) {
        float best = 0.f;
        int bi = -1;
        for (int i = tid; i < n; i += blockDim.x) {
            float v = row[i];
            if (v != v) continue;
            bool taken = false;
            for (int t = 0; t < j; ++t) taken |= (s_taken[t] == i);
            if (taken) continue;
            if (!largest) v = -v;
            if (bi < 0 || v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            float bv = 0.f;
            int bj = -1;
            for (int w = 0; w < nw; ++w)
                if (s_idx[w] >= 0 && (bj < 0 || s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bj))) { bv = s_val[w]; bj = s_idx[w]; }
            s_taken[j] = bj;
            vals[(size_t)blockIdx.x * k + j] = bj >= 0 ? row[bj] : (largest ? -FLT_MAX : FLT_MAX);
            idx[(size_t)blockIdx.x * k + j] = bj >= 0 ? bj : 0;
        }
        __syncthreads();
    }
}

// decision_heuristics.py:159-251.  lb_k [K][4B] = (lb - rhs).max(-1) of look-ahead pass k, rows [cand slot j in 0..2B)
// x (active, inactive)] laid out as the reference's doubled batch: row = half * 2B + j.  score_val / backup_val [B,K].
// out: dec_flat[B] = chosen flat neuron index over the concatenated layers.
__global__ void k_pick_decision(const float* __restrict__ lb_k, const float* __restrict__ score_val,
                                const int32_t* __restrict__ score_idx, const float* __restrict__ backup_val,
                                const int32_t* __restrict__ backup_idx, const float* __restrict__ mask_cat, int n_total,
                                int B, int K, int32_t* __restrict__ dec_flat) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float SMALL = 1e-6f, LARGE = 1e6f;
    float best_s = -FLT_MAX, best_b = -FLT_MAX;
    int ks = 0, kb = 0;
    for (int k = 0; k < K; ++k) {
        const float* row = lb_k + (size_t)k * 4 * B;
        const float inv_s = (score_val[(size_t)b * K + k] <= SMALL) ? LARGE : 0.f;
        const float inv_b = (backup_val[(size_t)b * K + k] >= -SMALL) ? LARGE : 0.f;
        const float vs = fmaxf(row[b] - inv_s, row[2 * B + b] - inv_s);
        const float vb = fmaxf(row[B + b] - inv_b, row[3 * B + b] - inv_b);
        if (vs > best_s) { best_s = vs; ks = k; }          // topk(1, 0): the first maximum
        if (vb > best_b) { best_b = vb; kb = k; }
    }
    int choice = -1;
    if (fmaxf(best_s, best_b) > -LARGE) {
        const int cand = best_s > best_b ? score_idx[(size_t)b * K + ks] : backup_idx[(size_t)b * K + kb];
        if (mask_cat[(size_t)b * n_total + cand] != 0.f) choice = cand;
    }
    if (choice < 0) {
        // no valid candidate: the reference draws a random layer and takes its first unstable neuron; here the first
        // unstable neuron in layer order (deterministic)
        for (int i = 0; i < n_total; ++i)
            if (mask_cat[(size_t)b * n_total + i] != 0.f) { choice = i; break; }
        if (choice < 0) choice = 0;
    }
    dec_flat[b] = choice;
}

}  // namespace

}  // namespace cb

extern "C" {

int cb_store_multi_copy(const cb_copy_desc_t* d_descs, int32_t n_descs, const int32_t* src_map, const int32_t* dst_map,
                        int32_t R, void* stream) {
    if (n_descs <= 0 || R <= 0) return CB_OK;
    cb::Launch _l(cb::K_STORE, (cudaStream_t)stream);
    dim3 grid((unsigned)(R < 2048 ? R : 2048), (unsigned)n_descs);
    cb::k_multi_copy<<<grid, 128, 0, (cudaStream_t)stream>>>(d_descs, src_map, dst_map, R);
    return cudaGetLastError() == cudaSuccess ? CB_OK : CB_ERR_CUDA;
}

int cb_store_apply_split(cb_split_layer_t* d_layers, int32_t n_layers, const int32_t* dec_layer, const int32_t* dec_neuron,
                         const float* dec_side, const float* dec_point, int32_t R, void* stream) {
    if (R <= 0) return CB_OK;
    cb::Launch _l(cb::K_STORE, (cudaStream_t)stream);
    cb::k_apply_split<<<(R + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_layers, n_layers, dec_layer, dec_neuron, dec_side,
                                                                         dec_point, R);
    return cudaGetLastError() == cudaSuccess ? CB_OK : CB_ERR_CUDA;
}

int cb_store_keep_rank(const float* lb, const float* rhs, int32_t R, int32_t S, int32_t base, int32_t* rank, int32_t* out,
                       const int32_t* const* d_hist_cnt, int32_t n_layers, void* stream) {
    cb::Launch _l(cb::K_STORE, (cudaStream_t)stream);
    cb::k_keep_rank<<<1, 1024, 0, (cudaStream_t)stream>>>(lb, rhs, R, S, base, rank, out, d_hist_cnt, n_layers);
    return cudaGetLastError() == cudaSuccess ? CB_OK : CB_ERR_CUDA;
}

int cb_babsr_scores(const float* lA, const float* lower, const float* upper, const float* bias, int32_t B, int32_t S,
                    int32_t n, float* score, float* backup, float* mask_out, int32_t ld, int32_t col0, void* stream) {
    if (B <= 0 || n <= 0) return CB_OK;
    cb::Launch _l(cb::K_BRANCH, (cudaStream_t)stream);
    const size_t total = (size_t)B * n;
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    cb::k_babsr<<<blocks, 256, 0, (cudaStream_t)stream>>>(lA, lower, upper, bias, B, S, n, score, backup, mask_out, ld, col0);
    return cudaGetLastError() == cudaSuccess ? CB_OK : CB_ERR_CUDA;
}

int cb_topk_rows(const float* x, int32_t B, int32_t n, int32_t k, int32_t largest, float* vals, int32_t* idx, void* stream) {
    if (B <= 0 || k <= 0) return CB_OK;
    if (k > 64) return CB_ERR_ARG;
    cb::Launch _l(cb::K_BRANCH, (cudaStream_t)stream);
    cb::k_topk_rows<<<B, 256, 64 * sizeof(float) + 64 * sizeof(int), (cudaStream_t)stream>>>(x, n, k, largest, vals, idx);
    return cudaGetLastError() == cudaSuccess ? CB_OK : CB_ERR_CUDA;
}

int cb_pick_decision(const float* lb_k, const float* score_val, const int32_t* score_idx, const float* backup_val,
                     const int32_t* backup_idx, const float* mask_cat, int32_t n_total, int32_t B, int32_t K,
                     int32_t* dec_flat, void* stream) {
    if (B <= 0) return CB_OK;
    cb::Launch _l(cb::K_BRANCH, (cudaStream_t)stream);
    cb::k_pick_decision<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(lb_k, score_val, score_idx, backup_val, backup_idx,
                                                                           mask_cat, n_total, B, K, dec_flat);
    return cudaGetLastError() == cudaSuccess ? CB_OK : CB_ERR_CUDA;
}

}  // extern "C"
