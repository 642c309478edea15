// Does TMA multicast raise the per-SM shared-memory fill rate?  Clusters of CL CTAs; every CTA wants the same stream of
// 32 KB tiles in its shared memory.  mode 0: every CTA loads the whole tile itself (unicast).  mode 1: every CTA loads
// 1/CL of the tile and multicasts it to all CTAs of the cluster.  Prints bytes landed per SM per clock.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_multicast_probe tma_multicast_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
constexpr int TILE = 32768, STAGES = 4;
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t *b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}
template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(128, 1) probe(const uint8_t *src, int iters, int mode, long long *out, int same) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *full = (uint64_t *)(sm + STAGES * TILE);
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) { for (int s = 0; s < STAGES; s++) mb_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("barrier.cluster.arrive.release.aligned; barrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint8_t *base = src + (same ? 0 : (size_t)(blockIdx.x / CL) * (size_t)iters * TILE);   // one stream per cluster, or ONE stream for the whole grid (same = 1: the A operand of a GEMM whose CTAs differ in N)
    long long t0 = clock64();
    for (int it0 = 0; it0 < iters; it0 += STAGES) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < STAGES; s++) {
                const uint8_t *g = base + (size_t)(it0 + s) * TILE;
                mb_expect(&full[s], TILE);
                if (mode == 0) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + s * TILE)), "l"(g), "r"(TILE), "r"(s32(&full[s])) : "memory");
                } else {
                    const uint32_t part = TILE / CL;
                    const uint16_t mask = (uint16_t)((1u << CL) - 1);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(s32(sm + s * TILE + rank * part)), "l"(g + rank * part), "r"(part), "r"(s32(&full[s])), "h"(mask) : "memory");
                }
            }
            for (int s = 0; s < STAGES; s++) mb_wait(&full[s], (it0 / STAGES) & 1);
        }
        // a peer must not multicast round r + 1 into this CTA's slots before this CTA has armed them: one cluster barrier per round
        if (CL > 1 && mode == 1) {
            __syncthreads();
            asm volatile("barrier.cluster.arrive.release.aligned; barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned; barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}
template <int CL> void run(const uint8_t *d, long long *dout, int grid, int iters, int same = 0) {
    CK(cudaFuncSetAttribute(probe<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * TILE + 1024));
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) { probe<CL><<<grid, 128, STAGES * TILE + 1024>>>(d, iters, mode, dout, same); CK(cudaDeviceSynchronize()); }
        long long cyc; CK(cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost));
        printf("%s cluster %d grid %d %s: %.1f bytes landed per SM per clock (%lld cycles for %d tiles)\n", same ? "one stream " : "per cluster", CL, grid, mode ? "multicast" : "unicast  ", (double)iters * TILE / cyc, cyc, iters);
    }
}
int main() {
    setvbuf(stdout, NULL, _IOLBF, 0);
    const int iters = 256;   // 8 MB per cluster stream: L2 resident after the first pass
    uint8_t *d; long long *dout;
    CK(cudaMalloc(&d, (size_t)148 * iters * TILE)); CK(cudaMemset(d, 1, (size_t)148 * iters * TILE)); CK(cudaMalloc(&dout, 8));
    run<1>(d, dout, 148, iters);
    run<2>(d, dout, 148, iters);
    run<4>(d, dout, 148, iters);
    run<2>(d, dout, 32, iters);
    run<4>(d, dout, 32, iters);
    // every CTA of the grid wants the SAME L2-resident stream (8 MB)
    run<1>(d, dout, 148, iters, 1); run<1>(d, dout, 64, iters, 1); run<1>(d, dout, 16, iters, 1);
    run<2>(d, dout, 148, iters, 1); run<4>(d, dout, 148, iters, 1);
    run<2>(d, dout, 64, iters, 1); run<4>(d, dout, 64, iters, 1);
    return 0;
}
