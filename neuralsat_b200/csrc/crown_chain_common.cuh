// Shared pieces of the whole-network kernels (crown_chain.cu: pass, crown_chain_grad.cu: gradient):
// tile constants, shared-memory operand layouts (see the header comment of crown_chain.cu) and small
// device helpers.
#pragma once
#include "crown_kernels.cuh"
#include "crown_tc_common.cuh"

namespace cb {
namespace chn {

using namespace tcc;

#ifndef CB_CHAIN_TR
#define CB_CHAIN_TR 64
#endif
// Sub-domain rows per CTA = MMA N.  64: one CTA per SM.  32 (-DCB_CHAIN_TR=32) was tried in r1f with two resident CTAs
// per SM that overlap each other's epilogue and MMA phases: SLOWER (pass 161 -> 178 us, grad 232 -> 306 us at 9472
// sub-domains; twice the MMA count and twice the weight stream cost more than the overlap recovers).  The option still
// compiles, but with the 72 KB weight ring and the per-thread bias sums two such CTAs no longer fit one SM.
constexpr int CH_TR = CB_CHAIN_TR;
constexpr int CH_CTAS_PER_SM = CH_TR <= 32 ? 2 : 1;
// Weight ring.  A bulk copy costs its issuing warp ~600-700 cycles whatever its size (scripts/tma_probe.cu: 3 KB ..
// 48 KB copies all take the same time, ring depth does not matter, and copies issued by different warps overlap), so
// the stream is fed in blocks of CH_WBLOCK_KS k-steps (24 KB: consecutive k-steps of an M-tile are contiguous in the
// packed weights) by CH_WSTAGES producer warps, one per ring stage.
#ifndef CB_CHAIN_WSTAGES
#define CB_CHAIN_WSTAGES 3
#endif
constexpr int CH_WSTAGES = CB_CHAIN_WSTAGES;
constexpr int CH_WBLOCK_KS = 2;
constexpr int CH_WKSTEP = 3 * 128 * 16 * 2;     // 12288 B: three planes of a [128 x 16] weight tile = one k-step
constexpr int CH_WSTAGE = CH_WBLOCK_KS * CH_WKSTEP;
constexpr int CH_WPLANE = 128 * 16 * 2;
// Resident row-tile operand X, MN-major, the three bf16 planes SIDE BY SIDE along N inside every k-group:
//   element (row n, k) of plane p at (k / 8) * CH_XKG3 + p * CH_XKG + (n / 8) * 128 + (k % 8) * 16 + (n % 8) * 2
// so that ONE descriptor (LBO = CH_XKG3, SBO = 128) with N = 64 / 128 / 192 addresses [x1], [x1|x2], [x1|x2|x3]:
// the six products of the bf16x3 split are three MMAs per k-step instead of six,
//   w1 . [x1|x2|x3] -> [main | small-1 | small-2],   w2 . [x1|x2] -> [small-1 | small-2],   w3 . [x1] -> [small-2]
// which reads 24 KB of shared memory per k-step instead of 36 KB (these MMAs are shared-memory-bandwidth bound:
// scripts/mma_probe.cu, profiles/README.md) and takes 208 instead of 288 tensor-pipe cycles.
constexpr int CH_XKG = (CH_TR / 8) * 128;        // bytes of one plane of a k-group (8 k-values x CH_TR rows)
constexpr int CH_XKG3 = 3 * CH_XKG;
constexpr int CH_XBYTES = (CHAIN_KMAX / 8) * CH_XKG3;
constexpr int CH_EPI_WARPS = CH_TR / 4;          // 4 TMEM lane quarters x (CH_TR / 16) row groups of 16 rows
// TMEM: one SLOT per M-tile = [main | small-1 | small-2] accumulators of CH_TR columns each; two slots (M-tile parity)
constexpr int CH_TSLOT = 3 * CH_TR;
constexpr int CH_TMEM_COLS = CH_TR == 64 ? 512 : 256;
static_assert(2 * CH_TSLOT <= CH_TMEM_COLS, "TMEM slots");
constexpr int CH_RPW = CH_TR / (CH_EPI_WARPS / 4);     // rows of the tile one epilogue warp owns
constexpr int CH_EPI_THREADS = CH_EPI_WARPS * 32;
// Warps 0 .. CH_EPI_WARPS-1 are the epilogue warps (whole warpgroups, TMEM lane quarter = warp % 4); the last
// warpgroup holds the weight producer, the MMA issuer and two idle warps.  With 64-row tiles (20 warps, launched
// at 96 registers) that warpgroup gives its registers back (setmaxnreg.dec) and the epilogue warpgroups grow to
// CH_EPI_REGS: two operand sets of 24 values in flight per thread do not fit in 96.
constexpr int CH_THREADS = CH_EPI_THREADS + 128;
constexpr int CH_WARP_MMA = CH_EPI_WARPS + 1;
static_assert(CH_WSTAGES <= 3, "producer warps: the three non-MMA warps of the last warpgroup");
// ring stage a warp of the last warpgroup produces, or -1 (the MMA warp)
__device__ __forceinline__ int chain_producer_stage(int warp) {
    const int w = warp - CH_EPI_WARPS;
    const int s = w == 0 ? 0 : w - 1;            // warps +0, +2, +3 -> stages 0, 1, 2
    return (w == 1 || s >= CH_WSTAGES) ? -1 : s;
}
constexpr int CH_EPI_REGS = 112;
constexpr int CH_AUX_REGS = 32;
constexpr int CH_LAUNCH_REGS = 96;      // what __launch_bounds__(640, 1) gives; the setmaxnreg pool is the LAUNCH allocation
static_assert(CH_TR != 64 || (CH_EPI_THREADS * CH_EPI_REGS + 128 * CH_AUX_REGS <= CH_THREADS * CH_LAUNCH_REGS), "register pool");

__device__ __forceinline__ void chain_set_regs(bool epilogue_group) {
    if constexpr (CH_TR == 64) {
        if (epilogue_group) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CH_EPI_REGS));
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CH_AUX_REGS));
    }
}
// Beta records of the layer being processed, per row (shared by both kernels):
//   s_bidx[row][neuron]  (bytes)  1 + index of the row's record at that neuron, 0 = none: ONE byte load per element
//                                 tells an epilogue thread whether (and which) record hits it
//   s_bval[row][jj]      (float)  pass: beta * sign (records at the same neuron pre-summed into the first);
//                                 gradient: sign
//   s_bnext[row][jj]     (bytes)  gradient only: 1 + index of the next record at the same neuron (every record has
//                                 its own d(beta)), 0 = end
//   s_bloc8[row][jj]     (bytes)  staging of the locations while the map is built
constexpr int CH_BIDX_BYTES = CH_TR * CHAIN_KMAX;
constexpr int CH_BVAL_BYTES = CH_TR * CHAIN_JMAX * 4;
constexpr int CH_BLOC8_BYTES = CH_TR * CHAIN_JMAX;
static_assert(CHAIN_KMAX <= 256 && CHAIN_JMAX < 255, "byte-sized neuron / record indices");
// pass: ring | X | s_bidx | s_bval | s_part [4][rows] (its first bytes double as s_bloc8 during a layer's set-up) |
//       s_extra [rows] | s_acc [4][epilogue threads] float4
constexpr int CH_PART_BYTES = 4 * CH_TR * 4 > CH_BLOC8_BYTES ? 4 * CH_TR * 4 : CH_BLOC8_BYTES;
constexpr int CH_SMEM_PASS = CH_WSTAGES * CH_WSTAGE + CH_XBYTES + CH_BIDX_BYTES + CH_BVAL_BYTES + CH_PART_BYTES + CH_TR * 4 +
                             4 * CH_EPI_THREADS * 16;
// gradient: ring | X | s_bidx | s_bval | s_bnext | s_bloc8
constexpr int CH_SMEM_GRAD = CH_WSTAGES * CH_WSTAGE + CH_XBYTES + CH_BIDX_BYTES + CH_BVAL_BYTES + 2 * CH_BLOC8_BYTES;
static_assert(CH_SMEM_PASS + 512 <= 227 * 1024, "shared memory of the pass kernel (dynamic + barriers)");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one lane of a converged warp (the CUTLASS elect_one_sync idiom: keeps tcgen05 / bulk-copy issue uniform)
__device__ __forceinline__ bool elect_one() {
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

__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void epi_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI_THREADS) : "memory");
}

// main + small-1 + small-2 accumulators of 8 columns (t, t + CH_TR, t + 2 CH_TR), one wait; small terms summed first
__device__ __forceinline__ void tmem_ld8x3(uint32_t t, float (&d)[8]) {
    uint32_t a[8], b[8], c[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%25];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%26];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
          "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
          "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7])
        : "r"(t), "r"(t + CH_TR), "r"(t + 2 * CH_TR)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = __uint_as_float(a[i]) + (__uint_as_float(b[i]) + __uint_as_float(c[i]));
}

// The three MMAs of one k-step (see CH_XKG3), issued by the elected lane.  The issuing warp is ONE instruction stream:
// at ~100 cycles of MMA work per instruction every descriptor add and register-file -> uniform-register move in
// front of it is exposed (measured: 375 cycles per k-step with generic 64-bit descriptor arithmetic, 205 for the MMAs
// themselves), so the descriptors are passed as their low words - the only part that changes: start address >> 4 in
// bits 0-13, never carrying into the stride fields - and assembled next to the MMAs.
//   d: the M-tile's TMEM slot; a_lo: weight plane 0 of the k-step (planes 1, 2 follow at CH_WPLANE); b_lo: the row tile's
//   k-step; a_hi / b_hi: the constant high words; idesc1: instruction descriptor with N = CH_TR
__device__ __forceinline__ void umma_split3(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc1, uint32_t accumulate) {
    constexpr uint32_t NSTEP = (uint32_t)(CH_TR >> 3) << 17;       // + CH_TR columns in the N field of the descriptor
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 da0, da1, da2, db;\n\t"
        ".reg .b32 a1, a2, d1, d2, i2, i3;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "add.u32 a1, %1, %7;\n\t"
        "add.u32 a2, %1, %8;\n\t"
        "mov.b64 da0, {%1, %2};\n\t"
        "mov.b64 da1, {a1, %2};\n\t"
        "mov.b64 da2, {a2, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "add.u32 d1, %0, %9;\n\t"
        "add.u32 d2, %0, %10;\n\t"
        "add.u32 i2, %5, %11;\n\t"
        "add.u32 i3, %5, %12;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da0, db, i3, p;\n\t"      // N = 3 CH_TR: first, it may overwrite
        "tcgen05.mma.cta_group::1.kind::f16 [d1], da1, db, i2, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [d2], da2, db, %5, q;\n\t"
        "}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc1), "r"(accumulate),
        "n"(CH_WPLANE >> 4), "n"(2 * (CH_WPLANE >> 4)), "n"(CH_TR), "n"(2 * CH_TR), "n"(NSTEP), "n"(2 * NSTEP)
        : "memory");
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
    uint8_t* p = X + (k >> 3) * CH_XKG3 + (row >> 3) * 128 + (k & 7) * 16 + (row & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(p) = h1;
    *reinterpret_cast<__nv_bfloat16*>(p + CH_XKG) = h2;
    *reinterpret_cast<__nv_bfloat16*>(p + 2 * CH_XKG) = h3;
}

// 8 consecutive rows n0..n0+7 (n0 % 8 == 0) of column k: one 16-byte store per plane
__device__ __forceinline__ void x_store8(uint8_t* X, int k, int n0, const float (&y)[8]) {
    uint4 p1, p2, p3;
    pack8(y, p1, p2, p3);
    uint8_t* p = X + (k >> 3) * CH_XKG3 + (n0 >> 3) * 128 + (k & 7) * 16;
    *reinterpret_cast<uint4*>(p) = p1;
    *reinterpret_cast<uint4*>(p + CH_XKG) = p2;
    *reinterpret_cast<uint4*>(p + 2 * CH_XKG) = p3;
}


}  // namespace chn
}  // namespace cb
