// Micro-probe (not product code): rate at which every CTA of a full wave can stream THE SAME weight buffer from
// the L2 into shared memory with bulk copies (cp.async.bulk, one 12 KB block per ring stage), the access pattern of
// the chain kernels' weight producer.  Variants: ring depth, block size, every CTA reading the same buffer vs its
// own copy, and a cluster of 2 / 4 CTAs sharing each block by multicast (every CTA fetches 1/cluster of the block
// and multicasts it to all CTAs of the cluster).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/tma_probe scripts/tma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x200000;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

struct Cfg { int stages, block, private_copy, cluster, producers; };     // producers: warps that issue copies (stage s belongs to warp s % producers)

// One producer thread per CTA; the "consumer" is the same thread: it waits for a stage and immediately reuses it
// (the MMAs would sit here), so the ring always has `stages` blocks in flight.
__global__ void __launch_bounds__(128, 1) k_stream(const uint8_t* src, size_t buf_bytes, Cfg c, int n_blocks, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[16];
    __shared__ __align__(8) uint64_t empty[16];       // cluster variant: every CTA of the cluster has released the stage
    const uint32_t cr = c.cluster > 1 ? cluster_rank() : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < c.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], c.cluster); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (c.cluster > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    const uint8_t* base = src + (c.private_copy ? (size_t)(blockIdx.x / c.cluster) * buf_bytes : 0);
    const int per_buf = (int)(buf_bytes / c.block);
    if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < c.producers) {
        const int pw = threadIdx.x >> 5;
        const long long t0 = clock64();
        const uint32_t part = c.block / c.cluster;
        const uint16_t mask = (uint16_t)((1u << c.cluster) - 1u);
        for (int i = 0; i < n_blocks + c.stages; ++i) {
            const int s = i % c.stages;
            if (s % c.producers != pw) continue;
            if (i >= c.stages) {
                mbar_wait(&full[s], ((i / c.stages) - 1) & 1);             // block i - stages has landed
                if (c.cluster > 1) {
                    // release the stage in every CTA of the cluster (remote arrive), then wait for all releases here
                    for (int r = 0; r < c.cluster; ++r) {
                        uint32_t remote;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&empty[s])), "r"(r));
                        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
                    }
                    mbar_wait(&empty[s], ((i / c.stages) - 1) & 1);
                }
            }
            if (i < n_blocks) {
                const uint8_t* g = base + (size_t)(i % per_buf) * c.block;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((uint32_t)c.block) : "memory");
                if (c.cluster > 1)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                                 ::"r"(smem_u32(smem + (size_t)s * c.block + cr * part)), "l"(g + cr * part), "r"(part), "r"(smem_u32(&full[s])), "h"(mask) : "memory");
                else
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(smem + (size_t)s * c.block)), "l"(g), "r"((uint32_t)c.block), "r"(smem_u32(&full[s])) : "memory");
            }
        }
        if (pw == 0) out[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (c.cluster > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}

int main() {
    const int ctas = 148, n_blocks = 840;            // 840 x 12 KB = 4 passes over a 2.5 MB weight set
    const size_t buf = 210 * 12288;
    uint8_t* d_src;
    long long* d_out;
    cudaMalloc(&d_src, buf * ctas);
    cudaMemset(d_src, 1, buf * ctas);
    cudaMalloc(&d_out, ctas * sizeof(long long));
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const Cfg cs[] = {{4, 12288, 0, 1, 1}, {8, 12288, 0, 1, 1}, {8, 3072, 0, 1, 1},  {8, 6144, 0, 1, 1},  {4, 24576, 0, 1, 1},
                      {4, 49152, 0, 1, 1}, {2, 49152, 0, 1, 1}, {2, 98304, 0, 1, 1}, {4, 12288, 0, 1, 2}, {4, 12288, 0, 1, 4},
                      {8, 12288, 0, 1, 4}, {8, 6144, 0, 1, 4},  {4, 12288, 1, 1, 4}};
    printf("stages block private cluster : producers : cycles per 12 KB (max over CTAs), B/cycle/SM\n");
    for (const Cfg& c : cs) {
        long long h[ctas];
        const int nb = (int)((long long)n_blocks * 12288 / c.block);
        for (int rep = 0; rep < 2; ++rep) {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(ctas); lc.blockDim = dim3(128); lc.dynamicSmemBytes = 200 * 1024;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = c.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&lc, k_stream, (const uint8_t*)d_src, buf, c, nb, d_out);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("cfg %d %d %d %d: %s\n", c.stages, c.block, c.private_copy, c.cluster, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < ctas; ++i) if (h[i] > mx) mx = h[i];
        const double per12k = (double)mx / n_blocks;
        printf("%2d %5d %d %d %d : %.0f  %.1f\n", c.stages, c.block, c.private_copy, c.cluster, c.producers, per12k, 12288.0 / per12k);
    }
    return 0;
}
