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
// Sub-domain rows per CTA = MMA N.  64: one CTA per SM.  32 (-DCB_CHAIN_TR=32): two resident CTAs per SM (2 x ~107 KB
// of shared memory, 2 x 256 TMEM columns, 2 x 320 threads) that overlap each other's epilogue and MMA phases; measured
// SLOWER on B200 (pass 161 -> 178 us, grad 232 -> 306 us at 9472 sub-domains): twice the MMA count at N = 32 and twice
// the weight stream cost more than the overlap recovers.
constexpr int CH_TR = CB_CHAIN_TR;
constexpr int CH_CTAS_PER_SM = CH_TR <= 32 ? 2 : 1;
constexpr int CH_WSTAGES = 4;
constexpr int CH_WSTAGE = 3 * 128 * 16 * 2;     // 12288 B: three planes of a [128 x 16] weight tile
constexpr int CH_WPLANE = 128 * 16 * 2;
constexpr int CH_XKG = (CH_TR / 8) * 128;        // bytes per (plane, 8 k-values) block: CH_TR / 8 row groups x 128 B
constexpr int CH_XPLANE = (CHAIN_KMAX / 8) * CH_XKG;
constexpr int CH_XBYTES = 3 * CH_XPLANE;
constexpr int CH_EPI_WARPS = CH_TR / 4;          // 4 TMEM lane quarters x (CH_TR / 16) row groups of 16 rows
constexpr int CH_TMT = 2 * CH_TR;                // TMEM columns of one M-tile: main + small-terms accumulator
constexpr int CH_TBUF = 2 * CH_TMT;              // ... of one buffer (two M-tiles)
constexpr int CH_TMEM_COLS = 2 * CH_TBUF;        // two buffers: 512 (64-row tiles) or 256 (32-row tiles)
constexpr int CH_RPW = CH_TR / (CH_EPI_WARPS / 4);     // rows of the tile one epilogue warp owns
constexpr int CH_EPI_THREADS = CH_EPI_WARPS * 32;
// Warps 0 .. CH_EPI_WARPS-1 are the epilogue warps (whole warpgroups, TMEM lane quarter = warp % 4); the last
// warpgroup holds the weight producer, the MMA issuer and two idle warps.  With 64-row tiles (20 warps, launched
// at 96 registers) that warpgroup gives its registers back (setmaxnreg.dec) and the epilogue warpgroups grow to
// CH_EPI_REGS: two operand sets of 24 values in flight per thread do not fit in 96.
constexpr int CH_THREADS = CH_EPI_THREADS + 128;
constexpr int CH_WARP_PRODUCER = CH_EPI_WARPS;
constexpr int CH_WARP_MMA = CH_EPI_WARPS + 1;
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
constexpr int CH_SMEM = CH_WSTAGES * CH_WSTAGE + CH_XBYTES + CH_TR * 8 * 4 + 2 * CH_TR * CHAIN_JMAX * 4 +
                        8 * CH_TR * 4 + CH_TR * 4 + 64 * 4 + 4 * CH_EPI_THREADS * 16;

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
    uint8_t* p = X + (k >> 3) * CH_XKG + (row >> 3) * 128 + (k & 7) * 16 + (row & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(p) = h1;
    *reinterpret_cast<__nv_bfloat16*>(p + CH_XPLANE) = h2;
    *reinterpret_cast<__nv_bfloat16*>(p + 2 * CH_XPLANE) = h3;
}

// 8 consecutive rows n0..n0+7 (n0 % 8 == 0) of column k: one 16-byte store per plane
__device__ __forceinline__ void x_store8(uint8_t* X, int k, int n0, const float (&y)[8]) {
    uint4 p1, p2, p3;
    pack8(y, p1, p2, p3);
    uint8_t* p = X + (k >> 3) * CH_XKG + (n0 >> 3) * 128 + (k & 7) * 16;
    *reinterpret_cast<uint4*>(p) = p1;
    *reinterpret_cast<uint4*>(p + CH_XPLANE) = p2;
    *reinterpret_cast<uint4*>(p + 2 * CH_XPLANE) = p3;
}


}  // namespace chn
}  // namespace cb
