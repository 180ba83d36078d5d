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
// Kernel.  Persistent: one CTA per SM (and N tile of <= 128 destination channels) walks tiles of n_mt x 128 consecutive
// positions q (M tiles of the MMA: TMEM lane = position).  24 warps, each role on its own:
//   warp 0        weight producer: [tap][channel/16][plane][2][N][8] by bulk copies - the whole packed tensor once when it
//                 fits in shared memory, else a ring of (chunk, tap) blocks;
//   warps 1-4     MMA issuers: (tile parity <-> TMEM buffer) x (half of the accumulator sets); three MMAs per (tap, M
//                 tile, 16 channels) when the three weight planes fit side by side along N, six otherwise, into a main
//                 and a small-terms accumulator per set;
//   11 warps      loaders: read the fp32 source (thread = position, coalesced along x), split it into the three bf16
//                 planes of the bf16x3 scheme (crown_tc.cu) and store it K-major, [plane][channel/8][position][channel%8],
//                 ONCE per K chunk: a tap's operand is the same buffer with the descriptor start address moved by
//                 shift*16 bytes (no im2col copy); the bias dot product of the pass rides here;
//   8 warps       epilogue: tcgen05.ld, main + small (+ bias, + old values when a second writer accumulates), strided
//                 NCHW stores (lanes = consecutive x).  Two TMEM buffers overlap it with the MMAs of the next tile; a
//                 strided pass whose class accumulators fill the TMEM hands them over class by class instead.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crown_kernels.cuh"
#include "crown_tc_common.cuh"

namespace cb {

namespace {

using namespace tcc;

constexpr int CT_LOAD_WARPS = 11;         // loader warps (source -> bf16x3 planes in shared memory)
constexpr int CT_SLOTS = 2;               // positions of a tile one loader thread owns: P <= CT_SLOTS * CT_LOADERS; 24 warps in all: 6 per scheduler leave 80 registers a thread
constexpr int CT_EPI_WARPS = 8;           // epilogue warps: 4 TMEM lane quarters x 2 column groups
constexpr int CT_LOADERS = CT_LOAD_WARPS * 32;
constexpr int CT_EPILOGUE = CT_EPI_WARPS * 32;
constexpr int CT_MMA_WARPS = 4;           // MMA issuers: (tile parity <-> TMEM buffer) x (half of the accumulator sets)
constexpr int CT_MAX_BIAS = 512;          // channels of the bias kept in shared memory (more: read through L1)
constexpr int CT_MAX_GROUPS = 384;        // MMA groups (tap, M tile, 16 channels) of one channel chunk held as a table
constexpr int CT_FIRST_LOADER = 32 * (1 + CT_MMA_WARPS);
constexpr int CT_THREADS = CT_FIRST_LOADER + CT_LOADERS + CT_EPILOGUE;   // warp 0: weight producer, warps 1-4: MMA issuers (warp 1 allocates TMEM)
constexpr int CT_SRC_STAGES = 3;
constexpr int CT_UNROLL = 2;              // source items a loader thread keeps in flight
constexpr int CT_W_STAGES = 24;
constexpr int CT_SMEM_MAX = 227 * 1024 - 9 * 1024 - 128;     // dynamic part; the barriers and the MMA group table are static

__device__ __forceinline__ void ct_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one lane of a converged warp (the CUTLASS elect_one_sync idiom)
__device__ __forceinline__ bool elect_one_ct() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}

// base + stride * i in ONE instruction (IMAD.WIDE.U32); written out because the compiler otherwise carries a 64-bit
// running offset next to the 64-bit base and spends four integer instructions per load / store address
__device__ __forceinline__ char* ct_chan_ptr(const char* base, uint32_t stride, uint32_t i) {
    uint64_t p;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(stride), "r"(i), "l"(reinterpret_cast<uint64_t>(base)));
    return reinterpret_cast<char*>(p);
}

// q / d for 0 <= q < 2^31 with mul = floor(2^32 / d): the estimate is the quotient or one below it
__device__ __forceinline__ int ct_div(int q, int d, uint32_t mul) {
    int r = (int)__umulhi((uint32_t)q, mul);
    if (q - r * d >= d) ++r;
    return r;
}

__device__ __forceinline__ void ct_tmem_ld8_raw(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

// Persistent: CTA b takes the position tiles b, b + gridDim.x, ...; per tile the loader warps fill the source stages
// chunk by chunk, the MMA thread accumulates into one of (up to) two TMEM buffers, the epilogue warps drain the other.
template <bool DBG>
__device__ __forceinline__ void conv_tc_body(const ConvTcArgs& a) {
    if (a.done != nullptr && *a.done != 0) return;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t src_full[CT_SRC_STAGES];
    __shared__ __align__(8) uint64_t src_empty[CT_SRC_STAGES];
    __shared__ __align__(8) uint64_t w_full[CT_W_STAGES];
    __shared__ __align__(8) uint64_t w_empty[CT_W_STAGES];
    __shared__ __align__(8) uint64_t acc_full[2];
    __shared__ __align__(8) uint64_t acc_empty[2];
    __shared__ __align__(8) uint64_t acc_full_c[CT_MAX_CLS];     // per accumulator class (strided pass, one TMEM buffer)
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t s_tap_a[CT_MAX_TAPS];      // per tap: (source buffer offset + shift * 16) >> 4
    __shared__ uint32_t s_tap_d[CT_MAX_TAPS];      // per tap: TMEM column offset of its accumulator set | slot << 28 | last-tap << 30 | first-tap << 31
    // resident weights: the MMA groups of one channel chunk as a flat table, the groups of share 0 first:
    //   x = A offset >> 4 inside a source stage, y = B offset >> 4 inside the weight tensor, z = TMEM column | first << 31
    __shared__ uint4 s_grp[CT_MAX_GROUPS];
    __shared__ int s_grp_n[2];                     // groups of share 0 / share 1
    __shared__ __align__(16) float s_bias[CT_MAX_BIAS];   // the layer bias, zero-padded to Kp (bias dot product of the loaders)

    const ConvTcGeom& g = a.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.y;
    const int Mcta = g.n_mt * 128;
    const int n_tiles = a.n_tiles;                         // position tiles of the launch
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
    const int n_slots = g.dir == 0 ? g.n_slots : 1;
    const int acc_buf_cols = n_slots * g.n_mt * acc_w;     // one TMEM buffer
    uint8_t* const s_src = smem;
    uint8_t* const s_w = smem + (size_t)g.src_stages * stage_bytes;
    const int w_total = g.n_taps * (g.Kp >> 4) * w_kstep;  // resident form: every tap, every channel of this N tile

    // issuers per tile: the accumulator sets (class, M tile) are split over two warps when there are at least two
    const int n_sets = n_slots * g.n_mt;
    const int n_ks = g.KC >> 4;
    const bool flat = g.w_resident && g.n_taps * g.n_mt * n_ks <= CT_MAX_GROUPS;
    const int n_share = (flat && n_sets >= 2) ? 2 : 1;
    // strided pass with a single TMEM buffer: accumulators are handed to the epilogue class by class
    const bool by_class = g.dir == 0 && n_slots > 1 && g.acc_bufs == 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < CT_SRC_STAGES; ++s) { mbar_init(&src_full[s], CT_LOAD_WARPS); mbar_init(&src_empty[s], n_share); }
        for (int s = 0; s < CT_W_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], n_share); mbar_init(&acc_empty[s], CT_EPI_WARPS); }
        for (int s = 0; s < CT_MAX_CLS; ++s) mbar_init(&acc_full_c[s], (n_share == 2 && g.n_mt >= 2) ? 2 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (flat) {
            int n = 0;
            for (int sh = 0; sh < n_share; ++sh) {
                const int n0 = n;
                for (int tap = 0; tap < g.n_taps; ++tap) {
                    const int slot = g.dir == 0 ? g.acc_slot[g.taps[tap].acc] : 0;
                    for (int mt = 0; mt < g.n_mt; ++mt) {
                        const int set = slot * g.n_mt + mt;
                        if (set % n_share != sh) continue;
                        for (int ks = 0; ks < n_ks; ++ks) {
                            uint4 e;
                            e.x = (uint32_t)((g.taps[tap].buf * buf_bytes) >> 4) + (uint32_t)g.taps[tap].shift + (uint32_t)(mt * 128) +
                                  (uint32_t)(ks * 2 * P);
                            e.y = (uint32_t)(((tap * (g.Kp >> 4) + ks) * w_kstep) >> 4);
                            e.z = (uint32_t)(set * acc_w) | (((g.taps[tap].first & 1) && ks == 0) ? 0x80000000u : 0u);
                            e.w = (uint32_t)slot << 8;           // bit 0 (set below): this share's last group of the class
                            s_grp[n++] = e;
                        }
                    }
                }
                unsigned closed = 0;                              // classes whose last group (of this share) is marked
                for (int i = n - 1; i >= n0; --i) {
                    const unsigned sl = s_grp[i].w >> 8;
                    if (!((closed >> sl) & 1u)) { s_grp[i].w |= 1u; closed |= 1u << sl; }
                }
                s_grp_n[sh] = n - n0;
            }
            if (n_share == 1) s_grp_n[1] = 0;
        }
    }
    if (threadIdx.x >= CT_FIRST_LOADER && threadIdx.x < CT_FIRST_LOADER + g.n_taps) {
        const int tap = threadIdx.x - CT_FIRST_LOADER;
        const int slot = g.dir == 0 ? g.acc_slot[g.taps[tap].acc] : 0;
        s_tap_a[tap] = (uint32_t)((g.taps[tap].buf * buf_bytes) >> 4) + (uint32_t)g.taps[tap].shift;
        s_tap_d[tap] = (uint32_t)(slot * g.n_mt * acc_w) | ((g.taps[tap].first & 1) ? 0x80000000u : 0u) |
                       ((g.taps[tap].first & 2) ? 0x40000000u : 0u) | ((uint32_t)slot << 28);
    }
    const bool bias_smem = g.Kp <= CT_MAX_BIAS;
    if (g.dir == 0 && a.bias != nullptr && bias_smem)
        for (int c = threadIdx.x; c < g.Kp; c += CT_THREADS) s_bias[c] = c < g.Csrc ? __ldg(a.bias + c) : 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(g.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0) {
        // ===== weight producer =====
        if (lane == 0) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wp) + (size_t)n_tile * w_total;
            if (g.w_resident) {
                // once per CTA; the mbarrier transaction count is 20 bits: hand the tensor over in slices
                int off = 0, s = 0;
                while (off < w_total) {
                    const int n = min(w_total - off, 64 * 1024);
                    mbar_expect_tx(&w_full[s], (uint32_t)n);
                    bulk_g2s(s_w + off, wsrc + off, (uint32_t)n, &w_full[s]);
                    off += n;
                    ++s;
                }
            } else {
                // one thread, one block per trip: every index is carried incrementally (a run-time division costs this
                // lone lane ~300 cycles, and the ring can only run as fast as this loop)
                const int n_blocks = my_tiles * n_chunks * g.n_taps;
                const size_t tap_stride = (size_t)(g.Kp >> 4) * w_kstep, chunk_stride = (size_t)(g.KC >> 4) * w_kstep;
                int s = 0, round = 0, tap = 0;
                size_t off = 0, chunk_off = 0;
                uint8_t* sdst = s_w;
                for (int b = 0; b < n_blocks; ++b) {
                    if (round > 0) mbar_wait(&w_empty[s], (uint32_t)(round - 1) & 1u);
                    mbar_expect_tx(&w_full[s], (uint32_t)w_block);
                    bulk_g2s(sdst, wsrc + off + chunk_off, (uint32_t)w_block, &w_full[s]);
                    sdst += w_block;
                    if (++s == g.w_stages) { s = 0; ++round; sdst = s_w; }
                    off += tap_stride;
                    if (++tap == g.n_taps) {
                        tap = 0; off = 0;
                        chunk_off += chunk_stride;
                        if (chunk_off == chunk_stride * (size_t)n_chunks) chunk_off = 0;
                    }
                }
            }
        }
    } else if (warp <= CT_MMA_WARPS) {
        // ===== MMA issuers: the whole warp runs the loop (uniform control flow keeps the descriptors in uniform
        // registers), one elected lane issues.  Descriptors are assembled from their low words: the start address >> 4
        // in bits 0-13 is the only part that changes (crown_chain_common.cuh:umma_split3) =====
        const uint32_t a_lbo = (uint32_t)P * 16;
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);                       // SBO = 128 B, descriptor version 1
        const uint32_t a_lo0 = ((a_lbo >> 4) & 0x3fffu) << 16;
        const uint32_t b_lbo = (uint32_t)(g.mma3 ? 3 * g.N16 : g.N16) * 16;
        const uint32_t b_lo0 = ((b_lbo >> 4) & 0x3fffu) << 16;
        const uint32_t idesc1 = umma_idesc_bf16(g.N16);
        const uint32_t nstep = (uint32_t)(g.N16 >> 3) << 17;                      // + N16 columns in the N field
        const uint32_t plane16 = (uint32_t)plane_bytes >> 4;
        const uint32_t bplane16 = (uint32_t)(g.N16 * 32) >> 4;
        if (g.w_resident) {
            const int n_slices = (w_total + 64 * 1024 - 1) / (64 * 1024);
            for (int s = 0; s < n_slices; ++s) mbar_wait(&w_full[s], 0);
        }
        // two accumulator buffers: warp 1 issues the even tiles, warp 2 the odd ones (one instruction stream per
        // buffer: the descriptor arithmetic in front of every MMA is what bounds a single issuer on small layers);
        // one buffer: warp 1 issues everything
        // Two issuers only when a tile is ONE ring item (single channel chunk) and the weights are resident: a
        // phase-parity wait is only meaningful for a consumer at most one phase away from its barrier.  With one item
        // per tile the issuer of tile t has consumed tile t-2, so every earlier fill of the stage it waits on (tile
        // t - stages <= t-2, loaded in order) has completed; with several chunks per tile, or a streamed weight ring
        // shared by both issuers, that no longer holds.
        // warp 1 + mw: tile parity = mw & 1 (only with two issue streams per tile parity, see above), share = mw >> 1
        const int mw = warp - 1;
        const bool two = g.acc_bufs > 1 && g.w_resident && n_chunks == 1;
        const int par = mw & 1, share = mw >> 1;
        const bool active = share < n_share && (two || par == 0);
        const int t_step = two ? 2 : 1;
        const int g_begin = share == 0 ? 0 : s_grp_n[0];
        const int g_count = flat ? s_grp_n[share < 2 ? share : 0] : 0;
        int ws = 0, wround = 0;                                  // weight ring: stage, fill count of that stage
        for (int t = (two ? par : 0); t < my_tiles && active; t += t_step) {
            const int tb = g.acc_bufs > 1 ? (t & 1) : 0;
            const int use = g.acc_bufs > 1 ? (t >> 1) : t;           // how often this buffer has been used before
            if (use > 0) mbar_wait_relaxed(&acc_empty[tb], (uint32_t)(use - 1) & 1u, 32);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (flat) {
                // ---- resident weights: one flat, table-driven loop over this issuer's MMA groups per chunk ----
                for (int chunk = 0; chunk < n_chunks; ++chunk) {
                    const int item = t * n_chunks + chunk;
                    const int cs = item % g.src_stages;
                    mbar_wait(&src_full[cs], (uint32_t)(item / g.src_stages) & 1u);
                    if (DBG && a.dbg && blockIdx.x == 0 && t < 16 && chunk == 0 && lane == 0 && share == 0) a.dbg[t * 8 + 4] = clock64();
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_base = a_lo0 + (smem_u32(s_src + (size_t)cs * stage_bytes) >> 4);
                    const uint32_t b_base = b_lo0 + ((smem_u32(s_w) + (uint32_t)(chunk * (g.KC >> 4) * w_kstep)) >> 4);
                    const uint32_t d_base = tmem_base + (uint32_t)(tb * acc_buf_cols);
                    const uint32_t first_mask = chunk == 0 ? 0x80000000u : 0u;
                    if (elect_one_ct()) {
                        for (int gi = 0; gi < g_count; ++gi) {
                            const uint4 e = s_grp[g_begin + gi];
                            const uint32_t a_lo = a_base + e.x, b_lo = b_base + e.y;
                            const uint32_t d0 = d_base + (e.z & 0x7fffffffu);
                            const uint32_t acc = (e.z & first_mask) ? 0u : 1u;
                            if (DBG && a.dbg_align == 2) continue;
                            if (g.mma3) {
                                asm volatile(
                                    "{\n\t"
                                    ".reg .pred p, q;\n\t"
                                    ".reg .b64 da0, da1, da2, db;\n\t"
                                    ".reg .b32 a1, a2, d1, d2, i2, i3;\n\t"
                                    "setp.ne.b32 p, %6, 0;\n\t"
                                    "setp.eq.b32 q, 0, 0;\n\t"
                                    "add.u32 a1, %1, %7;\n\t"
                                    "add.u32 a2, a1, %7;\n\t"
                                    "mov.b64 da0, {%1, %2};\n\t"
                                    "mov.b64 da1, {a1, %2};\n\t"
                                    "mov.b64 da2, {a2, %2};\n\t"
                                    "mov.b64 db, {%3, %4};\n\t"
                                    "add.u32 d1, %0, %8;\n\t"
                                    "add.u32 d2, d1, %8;\n\t"
                                    "add.u32 i2, %5, %9;\n\t"
                                    "add.u32 i3, i2, %9;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [%0], da0, db, i3, p;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [d1], da1, db, i2, q;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [d2], da2, db, %5, q;\n\t"
                                    "}" ::"r"(d0), "r"(a_lo), "r"(desc_hi), "r"(b_lo), "r"(desc_hi), "r"(idesc1), "r"(acc),
                                    "r"(plane16), "r"((uint32_t)g.N16), "r"(nstep)
                                    : "memory");
                            } else {
                                asm volatile(
                                    "{\n\t"
                                    ".reg .pred p, q;\n\t"
                                    ".reg .b64 da0, da1, da2, db0, db1, db2;\n\t"
                                    ".reg .b32 a1, a2, b1, b2, ds;\n\t"
                                    "setp.ne.b32 p, %6, 0;\n\t"
                                    "setp.eq.b32 q, 0, 0;\n\t"
                                    "add.u32 a1, %1, %7;\n\t"
                                    "add.u32 a2, a1, %7;\n\t"
                                    "add.u32 b1, %3, %8;\n\t"
                                    "add.u32 b2, b1, %8;\n\t"
                                    "mov.b64 da0, {%1, %2};\n\t"
                                    "mov.b64 da1, {a1, %2};\n\t"
                                    "mov.b64 da2, {a2, %2};\n\t"
                                    "mov.b64 db0, {%3, %4};\n\t"
                                    "mov.b64 db1, {b1, %4};\n\t"
                                    "mov.b64 db2, {b2, %4};\n\t"
                                    "add.u32 ds, %0, %9;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [ds], da2, db0, %5, p;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [ds], da1, db1, %5, q;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [ds], da0, db2, %5, q;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [ds], da1, db0, %5, q;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [ds], da0, db1, %5, q;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [%0], da0, db0, %5, p;\n\t"
                                    "}" ::"r"(d0), "r"(a_lo), "r"(desc_hi), "r"(b_lo), "r"(desc_hi), "r"(idesc1), "r"(acc),
                                    "r"(plane16), "r"(bplane16), "r"((uint32_t)g.N16)
                                    : "memory");
                            }
                            // this share's part of the class is complete: hand it to the epilogue
                            if (by_class && chunk == n_chunks - 1 && (e.w & 1u)) umma_commit(&acc_full_c[e.w >> 8]);
                        }
                        if (DBG && a.dbg && blockIdx.x == 0 && t < 16 && chunk == 0 && share == 0) a.dbg[128 + t * 4 + 0] = clock64();
                        umma_commit(&src_empty[cs]);
                    }
                    __syncwarp();
                }
                if (!by_class && elect_one_ct()) umma_commit(&acc_full[tb]);
                __syncwarp();
                if (DBG && a.dbg && blockIdx.x == 0 && t < 16 && lane == 0 && share == 0) a.dbg[t * 8 + 5] = clock64();
                continue;
            }
            for (int chunk = 0; chunk < n_chunks; ++chunk) {
                const int item = t * n_chunks + chunk;               // position in the source ring
                const int cs = item % g.src_stages;
                mbar_wait(&src_full[cs], (uint32_t)(item / g.src_stages) & 1u);
                if (DBG && a.dbg && blockIdx.x == 0 && t < 16 && chunk == 0 && lane == 0) a.dbg[t * 8 + 4] = clock64();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t s_stage16 = smem_u32(s_src + (size_t)cs * stage_bytes) >> 4;
                for (int tap = 0; tap < g.n_taps; ++tap) {
                    uint32_t wb16;
                    if (g.w_resident) {
                        wb16 = (smem_u32(s_w) + (uint32_t)((tap * (g.Kp >> 4) + chunk * (g.KC >> 4)) * w_kstep)) >> 4;
                    } else {
                        // ring position carried across tiles (this warp walks every (tile, chunk, tap) in order)
                        mbar_wait(&w_full[ws], (uint32_t)wround & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        wb16 = smem_u32(s_w + (size_t)ws * w_block) >> 4;
                    }
                    const uint32_t td = s_tap_d[tap];
                    uint32_t sa16 = s_stage16 + s_tap_a[tap];
                    if (DBG && a.dbg_align) sa16 &= ~7u;              // timing experiment only (wrong results): 128-byte aligned operand starts
                    const uint32_t first_tap = (chunk == 0 && (td >> 31)) ? 1u : 0u;
                    const uint32_t dtap = tmem_base + (uint32_t)(tb * acc_buf_cols) + (td & 0x0fffffffu);
                    if (elect_one_ct()) {
                        for (int mt = 0; mt < g.n_mt; ++mt) {
                            const uint32_t d0 = dtap + (uint32_t)(mt * acc_w);
                            for (int ks = 0; ks < (g.KC >> 4); ++ks) {
                                const uint32_t acc = (first_tap && ks == 0) ? 0u : 1u;
                                const uint32_t a_lo = a_lo0 + ((sa16 + (uint32_t)(mt * 128) + (uint32_t)(ks * 2) * (a_lbo >> 4)) & 0x3fffu);
                                const uint32_t b_lo = b_lo0 + ((wb16 + (uint32_t)ks * ((uint32_t)w_kstep >> 4)) & 0x3fffu);
                                if (DBG && a.dbg_align == 2) continue;       // timing experiment: issue loop without MMAs
                                if (g.mma3) {
                                    // the weight planes sit side by side along N: [w1 | w2 | w3] on ONE descriptor, so
                                    //   x1.[w1|w2|w3] -> [main | s1 | s2],  x2.[w1|w2] -> [s1 | s2],  x3.[w1] -> [s2]
                                    // gives the six products of the bf16x3 split in three MMAs
                                    asm volatile(
                                        "{\n\t"
                                        ".reg .pred p, q;\n\t"
                                        ".reg .b64 da0, da1, da2, db;\n\t"
                                        ".reg .b32 a1, a2, d1, d2, i2, i3;\n\t"
                                        "setp.ne.b32 p, %6, 0;\n\t"
                                        "setp.eq.b32 q, 0, 0;\n\t"
                                        "add.u32 a1, %1, %7;\n\t"
                                        "add.u32 a2, a1, %7;\n\t"
                                        "mov.b64 da0, {%1, %2};\n\t"
                                        "mov.b64 da1, {a1, %2};\n\t"
                                        "mov.b64 da2, {a2, %2};\n\t"
                                        "mov.b64 db, {%3, %4};\n\t"
                                        "add.u32 d1, %0, %8;\n\t"
                                        "add.u32 d2, d1, %8;\n\t"
                                        "add.u32 i2, %5, %9;\n\t"
                                        "add.u32 i3, i2, %9;\n\t"
                                        "tcgen05.mma.cta_group::1.kind::f16 [%0], da0, db, i3, p;\n\t"
                                        "tcgen05.mma.cta_group::1.kind::f16 [d1], da1, db, i2, q;\n\t"
                                        "tcgen05.mma.cta_group::1.kind::f16 [d2], da2, db, %5, q;\n\t"
                                        "}" ::"r"(d0), "r"(a_lo), "r"(desc_hi), "r"(b_lo), "r"(desc_hi), "r"(idesc1), "r"(acc),
                                        "r"(plane16), "r"((uint32_t)g.N16), "r"(nstep)
                                        : "memory");
                                } else {
                                    // five correction terms into their own accumulator (see crown_tc.cu), then the main one
                                    asm volatile(
                                        "{\n\t"
                                        ".reg .pred p, q;\n\t"
                                        ".reg .b64 da0, da1, da2, db0, db1, db2;\n\t"
                                        ".reg .b32 a1, a2, b1, b2, ds;\n\t"
                                        "setp.ne.b32 p, %6, 0;\n\t"
                                        "setp.eq.b32 q, 0, 0;\n\t"
                                        "add.u32 a1, %1, %7;\n\t"
                                        "add.u32 a2, a1, %7;\n\t"
                                        "add.u32 b1, %3, %8;\n\t"
                                        "add.u32 b2, b1, %8;\n\t"
                                        "mov.b64 da0, {%1, %2};\n\t"
                                        "mov.b64 da1, {a1, %2};\n\t"
                                        "mov.b64 da2, {a2, %2};\n\t"
                                        "mov.b64 db0, {%3, %4};\n\t"
                                        "mov.b64 db1, {b1, %4};\n\t"
                                        "mov.b64 db2, {b2, %4};\n\t"
                                        "add.u32 ds, %0, %9;\n\t"
                                        "tcgen05.mma.cta_group::1.kind::f16 [ds], da2, db0, %5, p;\n\t"
                                        "tcgen05.mma.cta_group::1.kind::f16 [ds], da1, db1, %5, q;\n\t"
                                        "tcgen05.mma.cta_group::1.kind::f16 [ds], da0, db2, %5, q;\n\t"
                                        "tcgen05.mma.cta_group::1.kind::f16 [ds], da1, db0, %5, q;\n\t"
                                        "tcgen05.mma.cta_group::1.kind::f16 [ds], da0, db1, %5, q;\n\t"
                                        "tcgen05.mma.cta_group::1.kind::f16 [%0], da0, db0, %5, p;\n\t"
                                        "}" ::"r"(d0), "r"(a_lo), "r"(desc_hi), "r"(b_lo), "r"(desc_hi), "r"(idesc1), "r"(acc),
                                        "r"(plane16), "r"(bplane16), "r"((uint32_t)g.N16)
                                        : "memory");
                                }
                            }
                        }
                        if (!g.w_resident) umma_commit(&w_empty[ws]);
                        if (by_class && chunk == n_chunks - 1 && (td & 0x40000000u)) umma_commit(&acc_full_c[(td >> 28) & 3u]);
                    }
                    if (!g.w_resident && ++ws == g.w_stages) { ws = 0; ++wround; }
                    __syncwarp();
                }
                if (DBG && a.dbg && blockIdx.x == 0 && t < 16 && lane == 0 && chunk == 0) a.dbg[128 + t * 4 + 0] = clock64();
                if (elect_one_ct()) umma_commit(&src_empty[cs]);
                __syncwarp();
                if (DBG && a.dbg && blockIdx.x == 0 && t < 16 && lane == 0 && chunk == 0) a.dbg[128 + t * 4 + 1] = clock64();
            }
            if (!by_class && elect_one_ct()) umma_commit(&acc_full[tb]);
            __syncwarp();
            if (DBG && a.dbg && blockIdx.x == 0 && t < 16 && lane == 0) a.dbg[t * 8 + 5] = clock64();
        }
    } else if (warp < 1 + CT_MMA_WARPS + CT_LOAD_WARPS) {
        // ===== loaders: source -> bf16x3 planes in shared memory =====
        // A thread owns up to CT_SLOTS positions of the tile (lanes = consecutive positions: coalesced along x), decodes
        // them once per tile into registers and walks their 8-channel groups chunk by chunk, CT_UNROLL groups of loads
        // in flight.  No shared tables, no barriers between the loader warps.
        const int te = threadIdx.x - CT_FIRST_LOADER;
        const int HWs = g.Hsrc * g.Wsrc;
        const uint32_t cstride = (uint32_t)HWs * 4u;             // bytes between two channels of a position
        const bool do_bias = g.dir == 0 && a.bias != nullptr && a.bias_rows != nullptr && n_tile == 0;
        // several class buffers (strided convolution of the gradient): spread (buffer, position) over the loader threads
        // instead of walking the buffers one after the other with the few threads a 128-position tile occupies
        const bool spread = n_buf > 1 && P * n_buf <= CT_SLOTS * CT_LOADERS;
        const int n_items = spread ? P * n_buf : P;
        int item = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * Mcta;      // rows * G < 2^31 is checked by the launcher
            if (DBG && a.dbg && blockIdx.x == 0 && te == 0 && t < 16) a.dbg[t * 8 + 0] = clock64();
            int pr[CT_SLOTS], py[CT_SLOTS], px[CT_SLOTS];          // sub-domain row (-1: outside the batch), y, x
            int pb[CT_SLOTS], pp[CT_SLOTS];                        // class buffer (spread mode), position inside the tile
#pragma unroll
            for (int sl = 0; sl < CT_SLOTS; ++sl) {
                const int j = te + sl * CT_LOADERS;
                int pl = j, buf = 0;
                if (spread) {
#pragma unroll
                    for (int b = 1; b < CT_MAX_CLS; ++b)
                        if (b < n_buf && j >= b * P) { buf = b; pl = j - b * P; }
                }
                pb[sl] = buf; pp[sl] = pl;
                const int q = m0 + g.dmin + pl;
                pr[sl] = -1; py[sl] = 0; px[sl] = 0;
                if (j < n_items && q >= 0) {
                    const int r = ct_div(q, g.G, g.mulG);
                    if (r < a.rows) {
                        const int rem = q - r * g.G;
                        pr[sl] = r;
                        py[sl] = ct_div(rem, g.Wp, g.mulWp);
                        px[sl] = rem - py[sl] * g.Wp;
                    }
                }
            }
            if (DBG && a.dbg && blockIdx.x == 0 && te == 0 && t < 16) a.dbg[t * 8 + 1] = clock64();
            for (int chunk = 0; chunk < n_chunks; ++chunk, ++item) {
                const int cs = item % g.src_stages;
                if (item >= g.src_stages) mbar_wait_relaxed(&src_empty[cs], (uint32_t)((item / g.src_stages) - 1) & 1u, 200);
                if (DBG && a.dbg && blockIdx.x == 0 && te == 0 && t < 16 && chunk == 0) a.dbg[t * 8 + 2] = clock64();
                uint8_t* const stage = s_src + (size_t)cs * stage_bytes;
                const int cbase = chunk * g.KC;
                const bool full_k = cbase + g.KC <= g.Csrc;          // no padding channels in this chunk
                for (int bi = 0; bi < (spread ? 1 : n_buf); ++bi) {
#pragma unroll
                    for (int sl = 0; sl < CT_SLOTS; ++sl) {
                        const int j = te + sl * CT_LOADERS;
                        if (sl * CT_LOADERS + (te - lane) >= n_items) break;             // warp-uniform
                        const int buf = spread ? pb[sl] : bi;
                        const int pl = pp[sl];
                        uint8_t* const bstage = stage + (size_t)buf * buf_bytes;
                        int hv, wv, oh, ow, sH, sW;
                        if (g.dir == 0) { hv = g.Hsrc; wv = g.Wsrc; oh = 0; ow = 0; sH = 1; sW = 1; }
                        else { hv = g.cls_h[buf]; wv = g.cls_w[buf]; oh = g.cls_oh[buf]; ow = g.cls_ow[buf]; sH = g.sh; sW = g.sw; }
                        const bool live = pr[sl] >= 0 && py[sl] < hv && px[sl] < wv;
                        // byte pointers and a 32-bit channel stride: one widening multiply-add per load address
                        const char* const sp0 = reinterpret_cast<const char*>(
                            a.src + (live ? (size_t)pr[sl] * g.Csrc * HWs + (size_t)(sH * py[sl] + oh) * g.Wsrc + (sW * px[sl] + ow) : 0) +
                            (size_t)cbase * HWs);
                        uint8_t* const d0 = bstage + (size_t)(j < n_items ? pl : 0) * 16;
                        float bsum = 0.f;
                        for (int g0 = 0; g0 < KG; g0 += CT_UNROLL) {
                            float v[CT_UNROLL][8];
#pragma unroll
                            for (int u = 0; u < CT_UNROLL; ++u) {
                                const char* const sp = ct_chan_ptr(sp0, cstride, (uint32_t)((g0 + u) * 8));
                                if (g0 + u < KG && live) {
                                    if (full_k) {
#pragma unroll
                                        for (int i = 0; i < 8; ++i)
                                            v[u][i] = __ldg(reinterpret_cast<const float*>(ct_chan_ptr(sp, cstride, (uint32_t)i)));
                                    } else {
#pragma unroll
                                        for (int i = 0; i < 8; ++i)
                                            v[u][i] = (cbase + (g0 + u) * 8 + i < g.Csrc)
                                                          ? __ldg(reinterpret_cast<const float*>(ct_chan_ptr(sp, cstride, (uint32_t)i))) : 0.f;
                                    }
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
                                }
                            }
#pragma unroll
                            for (int u = 0; u < CT_UNROLL; ++u) {
                                if (g0 + u >= KG) break;
                                if (do_bias && live) {
                                    const int c0 = cbase + (g0 + u) * 8;
                                    if (bias_smem) {
                                        const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c0);
                                        const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c0 + 4);
                                        bsum = fmaf(v[u][0], b0.x, bsum); bsum = fmaf(v[u][1], b0.y, bsum);
                                        bsum = fmaf(v[u][2], b0.z, bsum); bsum = fmaf(v[u][3], b0.w, bsum);
                                        bsum = fmaf(v[u][4], b1.x, bsum); bsum = fmaf(v[u][5], b1.y, bsum);
                                        bsum = fmaf(v[u][6], b1.z, bsum); bsum = fmaf(v[u][7], b1.w, bsum);
                                    } else {
#pragma unroll
                                        for (int i = 0; i < 8; ++i)
                                            if (full_k || c0 + i < g.Csrc) bsum = fmaf(v[u][i], __ldg(a.bias + c0 + i), bsum);
                                    }
                                }
                                if (j < n_items) {
                                    uint4 p1, p2, p3;
                                    pack8(v[u], p1, p2, p3);
                                    uint8_t* d = d0 + (size_t)(g0 + u) * P * 16;       // [kgroup][position]
                                    *reinterpret_cast<uint4*>(d) = p1;
                                    *reinterpret_cast<uint4*>(d + plane_bytes) = p2;
                                    *reinterpret_cast<uint4*>(d + 2 * plane_bytes) = p3;
                                }
                            }
                        }
                        if (do_bias) {
                            // bias dot product of the positions this tile owns: segmented warp reduction keyed by the
                            // sub-domain row (contiguous runs of lanes), one atomic per run
                            int r = -1;
                            if (pl < P && pl >= -g.dmin && pl < -g.dmin + Mcta) r = pr[sl];
                            if (r < 0) bsum = 0.f;
                            const int rp = __shfl_up_sync(0xffffffffu, r, 1);
                            const bool head = lane == 0 || rp != r;
                            const unsigned hb = __ballot_sync(0xffffffffu, head);
                            const int seg = __popc(hb & (0xffffffffu >> (31 - lane)));
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const float vo = __shfl_down_sync(0xffffffffu, bsum, o);
                                const int so = __shfl_down_sync(0xffffffffu, seg, o);
                                if (lane + o < 32 && so == seg) bsum += vo;
                            }
                            if (head && r >= 0 && bsum != 0.f) atomicAdd(a.bias_rows + r, bsum);
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) ct_mbar_arrive(&src_full[cs]);
            }
            if (DBG && a.dbg && blockIdx.x == 0 && te == 0 && t < 16) a.dbg[t * 8 + 3] = clock64();
        }
    } else {
        // ===== epilogue: TMEM lane quarter = warp % 4.  Two M tiles: each half of the epilogue warps drains one of them
        // (every column; the position of a lane is decoded once per tile); otherwise the halves split the columns =====
        const int quarter = warp & 3;
        const int grp = (warp - (1 + CT_MMA_WARPS + CT_LOAD_WARPS)) >> 2;
        const bool by_mt = g.n_mt == 2;
        const int mt_lo = by_mt ? grp : 0, mt_hi = by_mt ? grp + 1 : g.n_mt;
        const int c_first = by_mt ? 0 : grp * 8, c_step = by_mt ? 8 : 8 * (CT_EPI_WARPS / 4);
        const int HWd = g.Hdst * g.Wdst;
        const uint32_t dstride = (uint32_t)HWd * 4u;             // bytes between two channels of a position
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int n_acc = g.dir == 0 ? g.n_cls : 1;
        const bool mma3 = g.mma3 != 0;
        const bool add_bias = g.dir == 1 && a.bias != nullptr;
        const bool plain = !a.accumulate && !add_bias;           // the common case: sum the planes, store
        for (int t = 0; t < my_tiles; ++t) {
            const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * Mcta;
            const int tb = g.acc_bufs > 1 ? (t & 1) : 0;
            const int use = g.acc_bufs > 1 ? (t >> 1) : t;
            if (!by_class) mbar_wait_relaxed(&acc_full[tb], (uint32_t)use & 1u, 200);
            if (DBG && a.dbg && blockIdx.x == 0 && threadIdx.x == CT_FIRST_LOADER + CT_LOADERS && t < 16) a.dbg[t * 8 + 6] = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool stamp = DBG && a.dbg && blockIdx.x == 0 && threadIdx.x == CT_FIRST_LOADER + CT_LOADERS && t == 2;
            int it = 0;
            for (int mt = mt_lo; mt < mt_hi; ++mt) {
                const int q = m0 + mt * 128 + quarter * 32 + lane;
                const int r = ct_div(q, g.G, g.mulG);
                const int rem = q - r * g.G;
                const int y = ct_div(rem, g.Wp, g.mulWp), x = rem - y * g.Wp;
                for (int ak = 0; ak < n_acc; ++ak) {
                    const int ai = g.dir == 0 ? g.cls_order[ak] : 0;
                    const int slot = g.dir == 0 ? g.acc_slot[ai] : 0;
                    if (slot < 0 && a.accumulate) continue;          // nothing reaches this class
                    if (by_class && slot >= 0) {
                        // one TMEM buffer: the classes arrive one by one (a second wait on a completed phase returns at once)
                        mbar_wait_relaxed(&acc_full_c[slot], (uint32_t)t & 1u, 100);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    int hv, wv, ys, xs;
                    if (g.dir == 0) { hv = g.cls_h[ai]; wv = g.cls_w[ai]; ys = g.sh * y + g.cls_oh[ai]; xs = g.sw * x + g.cls_ow[ai]; }
                    else { hv = g.Hdst; wv = g.Wdst; ys = y; xs = x; }
                    const bool valid = r < a.rows && y < hv && x < wv;
                    float* const dp = a.dst + (size_t)r * g.Cdst * HWd + (size_t)ys * g.Wdst + xs;
                    const uint32_t tcol = trow + (uint32_t)(tb * acc_buf_cols + (max(slot, 0) * g.n_mt + mt) * acc_w);
                    for (int c0 = c_first; c0 < g.N16; c0 += c_step) {
                        const int cb = n_tile * g.N16 + c0;
                        if (cb >= g.Cdst) break;                                 // padding columns only (warp-uniform)
                        uint32_t r0[8], r1[8], r2[8];
                        if (stamp && it < 16) a.dbg[192 + it * 4 + 0] = clock64();
                        if (slot >= 0 && !(DBG && a.dbg_align == 4)) {
                            ct_tmem_ld8_raw(tcol + (uint32_t)c0, r0);
                            ct_tmem_ld8_raw(tcol + (uint32_t)(g.N16 + c0), r1);
                            if (mma3) ct_tmem_ld8_raw(tcol + (uint32_t)(2 * g.N16 + c0), r2);
                        }
                        const int nc = min(8, g.Cdst - cb);                      // warp-uniform
                        float val[8];
                        if (plain) {
                            if (stamp && it < 16) a.dbg[192 + it * 4 + 1] = clock64();
                            if (slot >= 0 && !(DBG && a.dbg_align == 4)) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            if (stamp && it < 16) a.dbg[192 + it * 4 + 2] = clock64();
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float sm = __uint_as_float(r1[i]);
                                if (mma3) sm += __uint_as_float(r2[i]);
                                val[i] = slot >= 0 ? __uint_as_float(r0[i]) + sm : 0.f;
                            }
                        } else {
                            // the destination's old values (residual fan-out: a second writer accumulates) and the bias
                            // are requested while the TMEM reads are in flight
                            float old[8], bv[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) { old[i] = 0.f; bv[i] = 0.f; }
                            if (add_bias) {
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (i < nc) bv[i] = __ldg(a.bias + cb + i);
                            }
                            if (a.accumulate && valid) {
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (i < nc) old[i] = dp[(size_t)(cb + i) * HWd];
                            }
                            if (stamp && it < 16) a.dbg[192 + it * 4 + 1] = clock64();
                            if (slot >= 0 && !(DBG && a.dbg_align == 4)) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            if (stamp && it < 16) a.dbg[192 + it * 4 + 2] = clock64();
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float v = old[i];
                                if (slot >= 0) {
                                    float sm = __uint_as_float(r1[i]);
                                    if (mma3) sm += __uint_as_float(r2[i]);
                                    v += __uint_as_float(r0[i]) + sm;
                                }
                                val[i] = v + bv[i];
                            }
                        }
                        // straight-line stores (a branch per channel serialises them)
                        if (valid && !(DBG && a.dbg_align == 3)) {
                            char* const dq = reinterpret_cast<char*>(dp + (size_t)cb * HWd);
                            if (nc == 8) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) *reinterpret_cast<float*>(ct_chan_ptr(dq, dstride, (uint32_t)i)) = val[i];
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (i < nc) *reinterpret_cast<float*>(ct_chan_ptr(dq, dstride, (uint32_t)i)) = val[i];
                            }
                        }
                        if (stamp && it < 16) a.dbg[192 + it * 4 + 3] = clock64();
                        ++it;
                    }
                }
            }
            if (DBG && a.dbg && blockIdx.x == 0 && threadIdx.x == CT_FIRST_LOADER + CT_LOADERS && t < 16) a.dbg[t * 8 + 7] = clock64();
            // this buffer may be overwritten by the MMAs of the tile after next
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) ct_mbar_arrive(&acc_empty[tb]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols) : "memory");
    }
}

__global__ void __launch_bounds__(CT_THREADS, 1) k_conv_tc(const __grid_constant__ ConvTcArgs a) { conv_tc_body<false>(a); }
// CB_CONV_DBG=1 / CB_CONV_ALIGN: the same kernel with the clock64 stamps and the timing-experiment switches compiled in
__global__ void __launch_bounds__(CT_THREADS, 1) k_conv_tc_dbg(const __grid_constant__ ConvTcArgs a) { conv_tc_body<true>(a); }

// W [Cout,Cin,T] -> [n_tile][tap][Kp/16][plane][2][N16][8]; pass: (n, k) = (ci, co), gradient: (n, k) = (co, ci)
// (mma3: [n_tile][tap][Kp/16][2][3*N16][8], the planes side by side along N)
struct TapOrder { short k[CT_MAX_TAPS]; };     // position in the kernel's tap order -> kh*KW + kw

__global__ void k_conv_tc_pack_w(const float* __restrict__ W, int Cout, int Cin, int T, int dir, int Kp, int N16,
                                 int n_ntiles, int mma3, TapOrder order, uint16_t* __restrict__ out) {
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
                w = W[((size_t)co * Cin + ci) * T + order.k[t]];
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

// Launch configuration for `rows` sub-domain rows: M tiles per CTA tile, channel chunk, stage counts, TMEM buffers.
// Preference order: double-buffered accumulators (the epilogue of one tile under the MMAs of the next) with resident
// weights; layers that stream their weights take the largest M per tile instead (every weight block is then reused
// by more positions) and a single accumulator buffer when two do not fit.
bool conv_tc_try(ConvTcGeom& g, int n_mt, int acc_bufs, bool allow_ring = true) {
    const int n_slots = g.dir == 0 ? g.n_slots : 1;
    const int n_buf = g.dir == 0 ? 1 : g.n_cls;
    const int acc_w = (g.mma3 ? 3 : 2) * g.N16;
    const int Mcta = n_mt * 128;
    const int P = (Mcta + g.span + 7) & ~7;
    if (P > CT_SLOTS * CT_LOADERS) return false;
    const int w_kstep = 3 * g.N16 * 32;
    const int w_total = g.n_taps * (g.Kp >> 4) * w_kstep;
    const int cols = acc_bufs * n_slots * n_mt * acc_w;
    int tm = 32;
    while (tm < cols) tm <<= 1;
    if (tm > 512) return false;
    // first a shape that keeps the whole packed weight tensor in shared memory (smaller channel chunks / fewer source
    // stages if that is what it takes: no per-tile weight stream from L2, no producer hand-shakes), then the ring
    static const bool prefer_resident = getenv("CB_CONV_PREFER_RESIDENT") ? atoi(getenv("CB_CONV_PREFER_RESIDENT")) != 0 : true;
    for (int pass = prefer_resident ? 0 : 1; pass < (allow_ring ? 2 : 1); ++pass)
        for (int KC = 64; KC >= 16; KC >>= 1) {
            if (g.Kp % KC != 0) continue;
            const int w_block = (KC >> 4) * w_kstep;
            const int blocks_per_tile = (g.Kp / KC) * g.n_taps;
            int best_stages = 0, best_w = 0, best_res = 0;
            long long best_src = 0;
            for (int stages = CT_SRC_STAGES; stages >= 2; --stages) {
                const long long src_bytes = (long long)stages * n_buf * 3 * (KC >> 3) * P * 16;
                const long long left = (long long)CT_SMEM_MAX - src_bytes - 1024;
                if (left <= 0) continue;
                if (w_total <= left && w_total <= CT_W_STAGES * 64 * 1024) {
                    best_stages = stages; best_w = 0; best_res = 1; best_src = src_bytes;
                    break;
                }
                if (pass == 0) continue;
                // streamed weights: the ring must cover the L2 latency with blocks in flight - small (chunk, tap)
                // blocks need many of them; give up a source stage when that deepens the ring
                int w_stages = (int)(left / w_block);
                if (w_stages > CT_W_STAGES) w_stages = CT_W_STAGES;
                if (w_stages > blocks_per_tile * 2) w_stages = blocks_per_tile * 2;
                if (w_stages < 2) continue;
                const bool ring_short = best_w > 0 && (long long)best_w * w_block < 48 * 1024;
                if (best_stages == 0 || (ring_short && w_stages > best_w)) {
                    best_stages = stages; best_w = w_stages; best_res = 0; best_src = src_bytes;
                }
            }
            if (best_stages == 0) continue;
            g.KC = KC; g.n_mt = n_mt; g.P = P; g.src_stages = best_stages; g.w_stages = best_w; g.w_resident = best_res;
            g.tmem_cols = tm; g.acc_bufs = acc_bufs;
            g.smem_bytes = (int)(best_src + (best_res ? w_total : (long long)best_w * w_block) + 1024);
            return true;
        }
    return false;
}

bool conv_tc_configure(ConvTcGeom& g, int rows) {
    const int n_slots = g.dir == 0 ? g.n_slots : 1;
    const int acc_w = (g.mma3 ? 3 : 2) * g.N16;
    const long long total = (long long)rows * g.G;
    const int w_total = g.n_taps * (g.Kp >> 4) * 3 * g.N16 * 32;
    const bool heavy_w = w_total > 96 * 1024;                // weights will be streamed per tile
    static const int heavy_bufs = getenv("CB_CONV_HEAVY_BUFS") ? atoi(getenv("CB_CONV_HEAVY_BUFS")) : 2;
    for (int acc_bufs = heavy_w ? heavy_bufs : 2; acc_bufs >= 1; --acc_bufs) {
        int n_max = 512 / (acc_bufs * n_slots * acc_w);
        if (n_max > 4) n_max = 4;
        if (!heavy_w && n_max > 2) n_max = 2;                // small tiles pipeline better when the weights are resident
        while (n_max > 1 && (total + n_max * 128 - 1) / (n_max * 128) < 2 * 148) --n_max;    // every SM busy on small batches
        // a smaller M tile that keeps the packed weights resident beats a larger one that has to stream them (light
        // layers only: the streamed ring of a heavy layer wants the large tile)
        if (!heavy_w)
            for (int n_mt = n_max; n_mt >= 1; --n_mt)
                if (conv_tc_try(g, n_mt, acc_bufs, false)) return true;
        for (int n_mt = n_max; n_mt >= 1; --n_mt)
            if (conv_tc_try(g, n_mt, acc_bufs)) return true;
    }
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
    g.mulG = g.G > 1 ? (unsigned)((1ull << 32) / (unsigned)g.G) : 0xffffffffu;
    g.mulWp = g.Wp > 1 ? (unsigned)((1ull << 32) / (unsigned)g.Wp) : 0xffffffffu;
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
    int n_of[CT_MAX_CLS] = {0, 0, 0, 0};
    for (int t = 0; t < g.n_taps; ++t) {
        g.taps[t].shift = deltas[t] - dmin;
        g.taps[t].ktap = (short)t;
        seen[g.taps[t].acc] = true;
        ++n_of[g.taps[t].acc];
    }
    // pass: slots in the order of ascending tap count - the class that is complete first is handed over first
    for (int k = 0; k < CT_MAX_CLS; ++k) g.cls_order[k] = k;
    if (dir == 0) {
        bool used[CT_MAX_CLS] = {false, false, false, false};
        int k = 0;
        for (;;) {
            int best = -1;
            for (int c = 0; c < g.n_cls; ++c)
                if (seen[c] && !used[c] && (best < 0 || n_of[c] < n_of[best])) best = c;
            if (best < 0) break;
            used[best] = true;
            g.acc_slot[best] = g.n_slots++;
            g.cls_order[k++] = best;
        }
        for (int c = 0; c < g.n_cls; ++c)
            if (!seen[c]) g.cls_order[k++] = c;
    }
    if (dir == 1) g.n_slots = 1;
    // strided pass: taps class by class (stable), so that a class's accumulator is complete - and its epilogue can
    // start - while the MMAs of the next classes are still being issued; the packed weights follow this order (ktap)
    if (dir == 0 && g.n_slots > 1) {
        ConvTcTap sorted[CT_MAX_TAPS];
        int n = 0;
        for (int slot = 0; slot < g.n_slots; ++slot)
            for (int t = 0; t < g.n_taps; ++t)
                if (g.acc_slot[g.taps[t].acc] == slot) sorted[n++] = g.taps[t];
        for (int t = 0; t < g.n_taps; ++t) g.taps[t] = sorted[t];
    }
    for (int t = 0; t < g.n_taps; ++t) {
        bool first = true, last = true;
        for (int u = 0; u < t; ++u) first = first && g.taps[u].acc != g.taps[t].acc;
        for (int u = t + 1; u < g.n_taps; ++u) last = last && g.taps[u].acc != g.taps[t].acc;
        g.taps[t].first = (short)((first ? 1 : 0) | (last ? 2 : 0));
    }
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
    TapOrder order;
    for (int t = 0; t < CT_MAX_TAPS; ++t) order.k[t] = t < g.n_taps ? g.taps[t].ktap : 0;
    k_conv_tc_pack_w<<<blocks, 256, 0, st>>>(W, Cout, Cin, g.n_taps, g.dir, g.Kp, g.N16, g.n_ntiles, g.mma3, order, out);
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
        e = cudaFuncSetAttribute(k_conv_tc_dbg, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_MAX);
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
    a.n_tiles = (int)((total + Mcta - 1) / Mcta);
    // persistent CTAs: one per SM (per N tile), each walks its share of the position tiles
    int n_cta = 148 / a.g.n_ntiles;
    if (n_cta < 1) n_cta = 1;
    if (n_cta > a.n_tiles) n_cta = a.n_tiles;
    dim3 grid((unsigned)n_cta, (unsigned)a.g.n_ntiles);
    a.dbg = nullptr;
    a.dbg_align = getenv("CB_CONV_ALIGN") ? atoi(getenv("CB_CONV_ALIGN")) : 0;
    static long long* d_dbg = nullptr;
    const char* edbg = getenv("CB_CONV_DBG");          // self-test: per-tile clock64 stamps of CTA 0, printed to stderr
    if (edbg && edbg[0] == '1') {
        if (!d_dbg) cudaMalloc(&d_dbg, 256 * sizeof(long long));
        cudaMemsetAsync(d_dbg, 0, 256 * sizeof(long long), st);
        a.dbg = d_dbg;
    }
    if (a.dbg || a.dbg_align) k_conv_tc_dbg<<<grid, CT_THREADS, a.g.smem_bytes, st>>>(a);
    else k_conv_tc<<<grid, CT_THREADS, a.g.smem_bytes, st>>>(a);
    if (a.dbg) {
        long long h[256];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[conv_tc dir %d C %d->%d G %d n_mt %d KC %d acc_bufs %d res %d src_stages %d w_stages %d P %d smem %d tiles %d grid %d] tile: table_start table_done src_slot loads_done | mma_src_ready mma_issued | epi_start epi_done (cycles from first stamp)\n",
                a.g.dir, a.g.Csrc, a.g.Cdst, a.g.G, a.g.n_mt, a.g.KC, a.g.acc_bufs, a.g.w_resident, a.g.src_stages, a.g.w_stages, a.g.P, a.g.smem_bytes, a.n_tiles, n_cta);
        for (int t = 0; t < 8 && h[t * 8] != 0; ++t) {
            fprintf(stderr, "  t%d:", t);
            for (int j = 0; j < 8; ++j) fprintf(stderr, " %lld", h[t * 8 + j] ? h[t * 8 + j] - h[0] : -1);
            fprintf(stderr, " | taps_done %lld src_commit_done %lld", h[128 + t * 4] - h[0], h[128 + t * 4 + 1] - h[0]);
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "  epilogue of tile 2 (accumulate %d), per column group: start | old requested | tmem waited | stores issued:", a.accumulate);
        for (int it = 0; it < 16 && h[192 + it * 4] != 0; ++it)
            fprintf(stderr, " [%lld %lld %lld %lld]", h[192 + it * 4] - h[0], h[192 + it * 4 + 1] - h[0], h[192 + it * 4 + 2] - h[0], h[192 + it * 4 + 3] - h[0]);
        fprintf(stderr, "\n");
    }
    return cudaGetLastError();
}

}  // namespace cb
