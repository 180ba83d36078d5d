// Micro-probe (not product code): issue rate of tcgen05.mma kind::f16 (bf16 in, fp32 out), M = 128, K = 16,
// cta_group::1, SS mode, as a function of N, operand majorness and shared-memory swizzle.  The chain kernels
// (neuralsat_b200/csrc/crown_chain*.cu) are bound by this rate; the probe tells which operand layout to use.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/mma_probe scripts/mma_probe.cu
// Run:   scripts/mma_probe            (prints cycles per MMA for every variant, 148 CTAs, max over CTAs)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t swz) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)swz << 61;       // 0 none, 2 = 128-byte swizzle
    return d;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}

struct Variant {
    int N;          // MMA N
    int b_mn;       // B MN-major
    int swz;        // 0 none, 2 = 128 B (both operands)
    int pattern;    // 0: distinct A/B per MMA from a ring; 1: the chain kernels' 6-MMA bf16x3 group (3 A planes x 3 B planes)
};

__global__ void __launch_bounds__(128, 1) k_probe(Variant v, int n_groups, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t bar2;          // per-group commits of patterns 4 / 5 arrive here (nobody waits)
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (warp == 1) {
        // the warp runs the loop converged and one elected lane issues (as the chain kernels do): the descriptors
        // stay in uniform registers and the six UTCHMMA of a group are issued back to back
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(v.N >> 3) << 17) | ((128u >> 4) << 24);
        if (v.b_mn) idesc |= 1u << 16;
        // A: 3 planes of [128 x 16] (no swizzle: 4 KB each, LBO 2048 / SBO 128) or [128 x 64] swizzled tiles (16 KB each)
        const uint32_t a_plane = v.swz ? 16384u : 4096u;
        const uint32_t sA = smem_u32(smem);
        // B after 64 KB: planes of N x 16 (none) or N x 64 (swizzled)
        const uint32_t sB = sA + 65536u;
        const uint32_t b_plane = v.swz ? (uint32_t)v.N * 128u : (uint32_t)v.N * 32u;
        uint64_t ad[3], bd[3];
        for (int p = 0; p < 3; ++p) {
            ad[p] = v.swz ? make_desc(sA + p * a_plane, 16, 1024, 2) : make_desc(sA + p * a_plane, 2048, 128, 0);
            if (v.swz) bd[p] = make_desc(sB + p * b_plane, v.b_mn ? 1024 : 16, 1024, 2);
            else bd[p] = v.b_mn ? make_desc(sB + p * b_plane, (v.N / 8) * 128, 128, 0) : make_desc(sB + p * b_plane, v.N * 16, 128, 0);
        }
        const uint32_t d_main = tmem, d_small = tmem + 256;
        const long long t0 = clock64();
        for (int g = 0; g < n_groups; ++g) {
            const uint32_t acc = g ? 1u : 0u;
            if (elect_one()) {
                if (v.pattern == 1 || v.pattern == 5) {
                    mma(d_small, ad[2], bd[0], idesc, acc);
                    mma(d_small, ad[1], bd[1], idesc, 1u);
                    mma(d_small, ad[0], bd[2], idesc, 1u);
                    mma(d_small, ad[1], bd[0], idesc, 1u);
                    mma(d_small, ad[0], bd[1], idesc, 1u);
                    mma(d_main, ad[0], bd[0], idesc, acc);
                } else if (v.pattern == 6) {
                    // the three-MMA group with operands that MOVE like in the chain kernels: the weight block cycles
                    // through 5 ring positions of 12 KB, the row-tile k-step through 16 positions of 6 KB
                    const uint32_t nstep = (uint32_t)(v.N >> 3) << 17;
                    const uint32_t ao = (uint32_t)(g % 5) * (12288u >> 4), bo = (uint32_t)(g & 15) * (6144u >> 4);
                    const uint64_t bcat = make_desc(sB, 3 * (v.N / 8) * 128, 128, 0) + bo;
                    mma(tmem, ad[0] + ao, bcat, idesc + 2 * nstep, acc);
                    mma(tmem + v.N, ad[1] + ao, bcat, idesc + nstep, 1u);
                    mma(tmem + 2 * v.N, ad[2] + ao, bcat, idesc, 1u);
                } else if (v.pattern == 2 || v.pattern == 3 || v.pattern == 4) {
                    // plane-concatenated operand: w1.[x1|x2|x3], w2.[x1|x2], w3.[x1]; N = 3n / 2n / n, one B descriptor
                    // (k-group stride 3 planes); pattern 2: overlapping accumulator columns [0,3n) [n,3n) [2n,3n);
                    // pattern 3: the same MMAs into disjoint columns (is the overlap what costs?)
                    const uint32_t nstep = (uint32_t)(v.N >> 3) << 17;
                    const uint64_t bcat = make_desc(sB, 3 * (v.N / 8) * 128, 128, 0);
                    const uint32_t o1 = v.pattern == 2 ? v.N : 3 * v.N, o2 = v.pattern == 2 ? 2 * v.N : 5 * v.N;
                    mma(tmem, ad[0], bcat, idesc + 2 * nstep, acc);
                    mma(tmem + o1, ad[1], bcat, idesc + nstep, 1u);
                    mma(tmem + o2, ad[2], bcat, idesc, 1u);
                } else {
                    mma(d_main, ad[0], bd[0], idesc, acc);
                    mma(d_main, ad[1], bd[1], idesc, 1u);
                    mma(d_main, ad[2], bd[2], idesc, 1u);
                    mma(d_main, ad[0], bd[1], idesc, 1u);
                    mma(d_main, ad[1], bd[2], idesc, 1u);
                    mma(d_main, ad[2], bd[0], idesc, 1u);
                }
                if (v.pattern == 4 || v.pattern == 5)      // one commit per k-step, as the chain kernels release their weight stage
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        const long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0, 0x200000;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
        const long long t2 = clock64();
        if ((threadIdx.x & 31) == 0) {
            out[2 * blockIdx.x] = t2 - t0;
            out[2 * blockIdx.x + 1] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    const int ctas = 148, n_groups = 200;
    long long* d_out;
    cudaMalloc(&d_out, 2 * ctas * sizeof(long long));
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const Variant vs[] = {
        {64, 1, 0, 1},  {64, 0, 0, 1},  {64, 0, 2, 1},  {64, 1, 2, 1},  {64, 1, 0, 0},
        {32, 1, 0, 1},  {128, 1, 0, 1}, {128, 0, 0, 1}, {128, 0, 2, 1}, {128, 1, 2, 1},
        {256, 1, 0, 1}, {256, 0, 2, 1}, {16, 1, 0, 1},
        {64, 1, 0, 2},  {64, 1, 0, 3},  {192, 1, 0, 0}, {192, 1, 0, 1}, {64, 1, 0, 4},  {64, 1, 0, 5},  {64, 1, 0, 6},
    };
    printf("N b_mn swz pattern : cycles/MMA (total, max over %d CTAs) issue-cycles/MMA | MMAs = %d\n", ctas, 6 * n_groups);
    for (const Variant& v : vs) {
        long long h[2 * ctas];
        for (int rep = 0; rep < 2; ++rep) {
            k_probe<<<ctas, 128, 200 * 1024>>>(v, n_groups, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant N=%d mn=%d swz=%d: %s\n", v.N, v.b_mn, v.swz, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0, mi = 0;
        for (int i = 0; i < ctas; ++i) { if (h[2 * i] > mx) mx = h[2 * i]; if (h[2 * i + 1] > mi) mi = h[2 * i + 1]; }
        const int per = ((v.pattern >= 2 && v.pattern <= 4) || v.pattern == 6) ? 3 : 6;      // MMAs per group
        printf("%3d %d %d %d : %.1f  %.1f\n", v.N, v.b_mn, v.swz, v.pattern, (double)mx / (per * n_groups), (double)mi / (per * n_groups));
    }
    return 0;
}
