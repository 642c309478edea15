// Per-SM shared-memory fill rate of a GEMM-like operand stream: every CTA pulls the SAME activation k-blocks (A hi + A lo, 128 rows x 128 B
// each) and its OWN weight k-blocks (BN rows x 128 B) through a 4-stage ring, one thread issuing, nothing consuming.
//   mode 0: 3D tiled tensor-map loads with 128B swizzle (what gemm_tc / k_flow_mega do)
//   mode 1: 1D bulk copies of the same bytes from a PRE-TILED copy of the operands (one contiguous 16 KB / BN*128 B block per k-block)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_tiled_probe tma_tiled_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
constexpr int STAGES = 4, KB = 16 /* k-blocks: K = 1024 */, BN = 64, ROUNDS = 16;
constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE = 2 * A_BYTES + B_BYTES;
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra W; }" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma3(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(dst)), "l"(m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk(void *dst, const void *g, uint32_t n, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(g), "r"(n), "r"(s32(bar)) : "memory");
}
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mAl, const __grid_constant__ CUtensorMap mB,
                                                const uint8_t *tA, const uint8_t *tAl, const uint8_t *tB, int mode, long long *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *full = (uint64_t *)(sm + STAGES * STAGE);
    if (threadIdx.x == 0) { for (int s = 0; s < STAGES; s++) mb_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        const int n0 = blockIdx.x * BN;
        int n = 0;
        for (int r = 0; r < ROUNDS; r++)
            for (int it = 0; it < KB; it++, n++) {
                const int s = n % STAGES;
                if (n >= STAGES) mb_wait(&full[s], ((n / STAGES) - 1) & 1);   // the previous load into this slot has landed ("consumed")
                uint8_t *st = sm + s * STAGE;
                mb_expect(&full[s], STAGE);
                if (mode == 0) {
                    tma3(st, &mA, &full[s], it * 64, 0, 0);
                    tma3(st + A_BYTES, &mAl, &full[s], it * 64, 0, 0);
                    tma3(st + 2 * A_BYTES, &mB, &full[s], it * 64, n0, 0);
                } else {
                    bulk(st, tA + (size_t)it * A_BYTES, A_BYTES, &full[s]);
                    bulk(st + A_BYTES, tAl + (size_t)it * A_BYTES, A_BYTES, &full[s]);
                    bulk(st + 2 * A_BYTES, tB + ((size_t)blockIdx.x * KB + it) * B_BYTES, B_BYTES, &full[s]);
                }
            }
        for (int k = n - STAGES; k < n; k++) mb_wait(&full[k % STAGES], (k / STAGES) & 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
static int make_map(CUtensorMap *m, void *base, int k, int rows, int box_rows) {
    cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, 1}, strides[2] = {(cuuint64_t)k * 2, (cuuint64_t)rows * k * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1}, es[3] = {1, 1, 1};
    return (int)cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
int main() {
    setvbuf(stdout, NULL, _IOLBF, 0);
    const int K = KB * 64, NW = 148 * BN;
    uint8_t *A, *Al, *B; long long *dout;
    CK(cudaMalloc(&A, 128 * K * 2)); CK(cudaMalloc(&Al, 128 * K * 2)); CK(cudaMalloc(&B, (size_t)NW * K * 2)); CK(cudaMalloc(&dout, 148 * 8));
    CK(cudaMemset(A, 1, 128 * K * 2)); CK(cudaMemset(Al, 1, 128 * K * 2)); CK(cudaMemset(B, 1, (size_t)NW * K * 2));
    CUtensorMap mA, mAl, mB;
    if (make_map(&mA, A, K, 128, 128) || make_map(&mAl, Al, K, 128, 128) || make_map(&mB, B, K, NW, BN)) { printf("tensor map encode failed\n"); return 1; }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * STAGE + 1024));
    const int grids[3] = {148, 64, 16};
    for (int gi = 0; gi < 3; gi++)
        for (int mode = 0; mode < 2; mode++) {
            long long h[148];
            for (int rep = 0; rep < 2; rep++) { probe<<<grids[gi], 128, STAGES * STAGE + 1024>>>(mA, mAl, mB, A, Al, B, mode, dout); CK(cudaDeviceSynchronize()); }
            CK(cudaMemcpy(h, dout, grids[gi] * 8, cudaMemcpyDeviceToHost));
            long long mx = 0; double mean = 0;
            for (int i = 0; i < grids[gi]; i++) { mx = h[i] > mx ? h[i] : mx; mean += (double)h[i] / grids[gi]; }
            const double bytes = (double)ROUNDS * KB * STAGE;
            printf("grid %3d %s: %.1f B/clk per SM (mean), %.1f (slowest CTA); %.0f cycles per %d KB stage\n", grids[gi], mode ? "1D bulk, pre-tiled " : "3D tiled tensor map",
                   bytes / mean, bytes / mx, mean / (ROUNDS * KB), STAGE / 1024);
        }
    return 0;
}
