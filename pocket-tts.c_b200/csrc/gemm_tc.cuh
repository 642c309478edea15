// gemm_tc.cuh - the tcgen05 / TMEM / TMA contraction kernel used for every dense hot op:
// batched FlowLM linears, the flow head, Mimi transformer linears, and Mimi Conv1d /
// ConvTranspose1d as implicit GEMM (taps = shifted TMA row coordinates; causal padding
// comes from TMA out-of-bounds zero fill, so no im2col buffer ever exists).
//
//   D[m, n] = sum_{tap} sum_{k} A[batch][m + shift(tap)][k + tap*a_k_step] * B[n][k + tap*b_k_step]
//
// A and B are both K-major 16-bit operands (fp16 or bf16) fetched by TMA into 128B-swizzled
// shared-memory tiles; accumulation is fp32 in TMEM; one elected thread issues tcgen05.mma.
// SPLIT = 1 / 2 adds a second ("lo") tensor for A / B so that an fp32 activation carried as
// hi + lo 16-bit halves keeps ~16 bits of significand (used where parity needs it).
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> fused bias / layer-scale / residual /
// activation -> global).
#pragma once
#include "common.cuh"

namespace ptts {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;        // 64 x 16-bit = 128 B = one swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_MAX_STAGES = 8;

struct GemmParams {
    int m_tiles_per_batch;   // ceil(a_rows / 128); grid.x = a_batches * m_tiles_per_batch
    int a_rows;              // valid rows of A per batch (time steps, or weight rows)
    int n_total;             // valid rows of B (output channels, or batch rows)
    int kblocks_per_tap;     // K per tap / 64
    int taps;
    int a_row_shift0;        // row shift of tap 0
    int a_row_step;          // additional row shift per tap
    int a_k_step;            // K offset per tap inside A (0 when taps shift rows)
    int b_k_step;            // K offset per tap inside B
    int stages;
    uint32_t idesc;
    // ---- epilogue
    int transposed;          // 0: element (m,n) -> [(batch*a_rows+m)*ld + n]; 1: -> [n*ld + m]
    const float *bias;       // indexed by feature (n if !transposed else m); NULL = none
    int bias_mask;           // 0 = index by feature, else feature & bias_mask (ConvTranspose phases)
    const float *scale;      // per-feature multiplier (layer scale / gate is handled elsewhere); NULL = none
    const float *resid;      // fp32 residual added after scale; NULL = none
    float *out_f32;          // fp32 output (may alias resid); NULL = none
    int ldo;
    size_t split_stride;     // split-K: slice z stores its partial sum at out_f32 + z*split_stride (reduced by the consumer)
    uint16_t *out_h;         // 16-bit output of act(v); NULL = none
    uint16_t *out_l;         // lo half: 16-bit of act(v) - float(hi); NULL = none
    int ldh;
    int out_bf16;            // encoding of out_h / out_l
    int act;
};

template <int BN, int SPLIT>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES * (SPLIT == 1 ? 2 : 1) + B_BYTES * (SPLIT == 2 ? 2 : 1);
    static constexpr int bytes(int stages) { return stages * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/; }
};

template <int BN, int SPLIT>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapBlo,
               const GemmParams p) {
    using SM = GemmSmem<BN, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + p.stages * SM::STAGE_BYTES);
    uint64_t *full = bars;                         // [stages]
    uint64_t *empty = bars + GEMM_MAX_STAGES;      // [stages]
    uint64_t *tmem_full = bars + 2 * GEMM_MAX_STAGES;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * GEMM_MAX_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int batch = blockIdx.x / p.m_tiles_per_batch;
    const int m0 = (blockIdx.x % p.m_tiles_per_batch) * GEMM_BM;
    const int n0 = blockIdx.y * BN;
    const int total_iters = p.taps * p.kblocks_per_tap;
    const int per = (total_iters + gridDim.z - 1) / gridDim.z;
    const int it0 = blockIdx.z * per;
    const int it1 = min(total_iters, it0 + per);
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        if (SPLIT == 1) tma_prefetch_desc(&mapAlo);
        if (SPLIT == 2) tma_prefetch_desc(&mapBlo);
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int it = it0; it < it1; it++) {
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t *st = smem + s * SM::STAGE_BYTES;
                mbar_expect_tx(&full[s], SM::STAGE_BYTES);
                const int tap = it / p.kblocks_per_tap, kb = it - tap * p.kblocks_per_tap;
                const int ak = kb * GEMM_BK + tap * p.a_k_step;
                const int ar = m0 + p.a_row_shift0 + tap * p.a_row_step;
                const int bk = kb * GEMM_BK + tap * p.b_k_step;
                tma_load_3d(st, &mapA, &full[s], ak, ar, batch);
                uint8_t *nx = st + SM::A_BYTES;
                if (SPLIT == 1) { tma_load_3d(nx, &mapAlo, &full[s], ak, ar, batch); nx += SM::A_BYTES; }
                tma_load_3d(nx, &mapB, &full[s], bk, n0, 0);
                if (SPLIT == 2) tma_load_3d(nx + SM::B_BYTES, &mapBlo, &full[s], bk, n0, 0);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one elected lane)
        int s = 0; uint32_t ph = 0;
        for (int it = it0; it < it1; it++) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(smem + s * SM::STAGE_BYTES);
                const uint32_t alo_addr = a_addr + SM::A_BYTES;
                const uint32_t b_addr = a_addr + SM::A_BYTES * (SPLIT == 1 ? 2 : 1);
                const uint32_t blo_addr = b_addr + SM::B_BYTES;
                const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(b_addr);
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; k++) {
                    // advancing 16 elements (32 B) along K inside the swizzled row: +2 in 16-byte units
                    umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, p.idesc, (it > it0 || k > 0) ? 1u : 0u);
                    if (SPLIT == 1) umma_f16(tmem_base, umma_desc_sw128(alo_addr) + 2 * k, bd + 2 * k, p.idesc, 1u);
                    if (SPLIT == 2) umma_f16(tmem_base, ad + 2 * k, umma_desc_sw128(blo_addr) + 2 * k, p.idesc, 1u);
                }
                umma_commit(&empty[s]);                 // smem slot reusable once these MMAs retire
                if (it == it1 - 1) umma_commit(tmem_full);  // accumulator complete
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
    } else if (it1 > it0) {
        // ------------------------------------------------ epilogue warps (TMEM lane quadrant = warp % 4)
        const int quad = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int m = m0 + quad * 32 + lane;
        const bool m_ok = m < p.a_rows;
        const bool bf = p.out_bf16 != 0;
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
        if (!p.transposed) {
            const size_t grow = (size_t)batch * p.a_rows + m;
            float *of = p.out_f32 ? p.out_f32 + blockIdx.z * p.split_stride + grow * p.ldo : nullptr;
            const float *rs = p.resid ? p.resid + grow * p.ldo : nullptr;
            uint16_t *oh = p.out_h ? p.out_h + grow * p.ldh : nullptr;
            uint16_t *ol = p.out_l ? p.out_l + grow * p.ldh : nullptr;
#pragma unroll 1
            for (int c = 0; c < BN; c += 16) {
                uint32_t r[16];
                __syncwarp();
                tmem_ld16(trow + c, r);
                tmem_ld_wait();
                const int nb = n0 + c;
                if (m_ok && nb < p.n_total) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int n = nb + j;
                        float x = __uint_as_float(r[j]);
                        if (p.bias && blockIdx.z == 0) x += p.bias[p.bias_mask ? (n & p.bias_mask) : n];
                        if (p.scale) x *= p.scale[n];
                        v[j] = x;
                    }
                    if (rs) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 t = *(const float4 *)(rs + nb + j);
                            v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
                        }
                    }
                    if (of) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *(float4 *)(of + nb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    if (oh) {
                        uint32_t hw[8], lw[8];
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const float a0 = act_apply(p.act, v[j]), a1 = act_apply(p.act, v[j + 1]);
                            const uint16_t h0 = f2h16(a0, bf), h1 = f2h16(a1, bf);
                            hw[j / 2] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                            if (ol) {
                                const uint16_t l0 = f2h16(a0 - h162f(h0, bf), bf), l1 = f2h16(a1 - h162f(h1, bf), bf);
                                lw[j / 2] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                            }
                        }
                        *(uint4 *)(oh + nb) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *(uint4 *)(oh + nb + 8) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                        if (ol) {
                            *(uint4 *)(ol + nb) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                            *(uint4 *)(ol + nb + 8) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                        }
                    }
                }
            }
        } else {
            // transposed: this thread owns output FEATURE m; columns are batch rows n. Stores are
            // coalesced across the warp (32 consecutive features per batch row).
            float bv = 0.0f, sv = 1.0f;
            if (m_ok) {
                if (p.bias && blockIdx.z == 0) bv = p.bias[p.bias_mask ? (m & p.bias_mask) : m];
                if (p.scale) sv = p.scale[m];
            }
            float *of = p.out_f32 ? p.out_f32 + blockIdx.z * p.split_stride : nullptr;
#pragma unroll 1
            for (int c = 0; c < BN; c += 16) {
                uint32_t r[16];
                __syncwarp();
                tmem_ld16(trow + c, r);
                tmem_ld_wait();
                if (m_ok) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int n = n0 + c + j;
                        if (n < p.n_total) {
                            float x = (__uint_as_float(r[j]) + bv) * sv;
                            const size_t o = (size_t)n * p.ldo + m;
                            if (p.resid) x += p.resid[o];
                            if (of) of[o] = x;
                            if (p.out_h) {
                                const float a = act_apply(p.act, x);
                                const uint16_t h = f2h16(a, bf);
                                const size_t oh = (size_t)n * p.ldh + m;
                                p.out_h[oh] = h;
                                if (p.out_l) p.out_l[oh] = f2h16(a - h162f(h, bf), bf);
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------- host side
struct TmaApi {
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeTiled encode = nullptr;
    int init();
};
extern TmaApi g_tma;

// 3D map over a 16-bit tensor [batches][rows][k] (k contiguous, row stride ld elements, batch stride
// batch_stride elements), box = {64, box_rows, 1}, 128B swizzle, zero OOB fill.
int make_tmap(CUtensorMap *out, const void *base, bool bf16, int k, int rows, int batches, size_t ld, size_t batch_stride,
              int box_rows);

struct GemmOp {
    CUtensorMap mapA, mapAlo, mapB, mapBlo;
    GemmParams p;
    int bn;        // N tile (B-operand box rows)
    int split;     // 0 none, 1 = A has lo, 2 = B has lo
    dim3 grid;
    int smem_bytes;
};

int gemm_setup_attrs();
int gemm_launch(const GemmOp &op, cudaStream_t st);
int gemm_pick_stages(int bn, int split, int iters, bool two_ctas_per_sm);

}  // namespace ptts
