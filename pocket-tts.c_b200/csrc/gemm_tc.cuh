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
constexpr int GEMM_EPI_WARPS = 16;   // 4 per TMEM lane quadrant, each takes every 4th 16-column chunk
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_MAX_STAGES = 8;

struct GemmParams {
    int m_tiles_per_batch;   // ceil(a_rows / 128)
    int m_tiles_total;       // a_batches * m_tiles_per_batch
    int n_tiles;             // ceil(n_total / BN)
    int split_k;             // K slices (>= 1, none empty)
    int iters_per_split;     // K iterations per slice
    int total_tiles;         // m_tiles_total * n_tiles * split_k  (persistent loop bound)
    int a_rows;              // valid rows of A per batch (time steps, or weight rows)
    int n_total;             // valid rows of B (output channels, or batch rows)
    int kblocks_per_tap;     // K per tap / 64
    int taps;
    int a_row_shift0;        // row shift of tap 0
    int a_row_step;          // additional row shift per tap
    int a_k_step;            // K offset per tap inside A (0 when taps shift rows)
    int b_k_step;            // K offset per tap inside B
    int stages;
    int b_static;            // 1: the B operand (weights) is never written by a kernel -> its first tiles are
                             // fetched before the programmatic-dependency wait (overlaps the previous kernel's tail)
    uint32_t idesc;
    // ---- epilogue
    int transposed;          // 0: element (m,n) -> [(batch*a_rows+m)*ld + n]; 1: -> [n*ld + m]
    const float *bias;       // indexed by feature (n if !transposed else m); NULL = none
    int bias_mask;           // 0 = index by feature, else feature & bias_mask (ConvTranspose phases)
    const float *scale;      // per-feature multiplier (layer scale); NULL = none
    const float *resid;      // fp32 residual added after scale; NULL = none
    const uint16_t *resid_h; // 16-bit residual (same encoding as out_h), row stride ldh; NULL = none
    uint16_t *out_raw;       // 16-bit output of v BEFORE the activation (the residual stream of a resblock), stride ldh
    float *out_f32;          // fp32 output (may alias resid); NULL = none
    int ldo;
    size_t split_stride;     // split-K: slice z stores its partial sum at out_f32 + z*split_stride (folded by the consumer)
    uint16_t *out_h;         // 16-bit output of act(v); NULL = none
    uint16_t *out_l;         // lo half: 16-bit of act(v) - float(hi); NULL = none
    int ldh;
    int out_bf16;            // encoding of out_h / out_l
    int act;
    int fast_act;            // 1: hardware exp/tanh approximations (16-bit outputs only)
    // fused Mimi QKV epilogue (ptts_mimi.c:329-341): columns are [q|k|v] x 8 heads x 64; q,k get RoPE at
    // position m (row within the batch item), q is pre-scaled by 1/8; fp16 head-major [B][H][T][64] outputs
    uint16_t *rope_q, *rope_k, *rope_v;
    const float *rope_cos, *rope_sin;   // [T][32]
    int rope_pos0;           // position of row 0 (streaming decode of a window that starts mid-utterance)
    int rope_rows;           // rows per head of the rope_q / rope_k / rope_v buffers (0: a_rows) and
    int rope_row0;           // destination row of A row 0: the streaming decoder appends a chunk's K / V to per-utterance caches
    int tma_store;           // 1: 16-bit outputs (out_h, out_raw) leave through shared-memory staging + TMA tensor stores
                             // (whole 128-byte lines) instead of one 32-byte store per thread and chunk
    int stage_off;           // byte offset (multiple of 1024) of the per-warp staging buffers from the aligned shared-memory base
    int stage_stride;        // staging bytes per epilogue warp: 4096 per staged output tensor
    long long *trace;        // optional [16] clock64 stamps written by CTA 0 (kernel timeline debugging)
};

template <int BN, int SPLIT>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES * (SPLIT == 1 ? 2 : 1) + B_BYTES * (SPLIT == 2 ? 2 : 1);
};

// 256-bit global accesses (sm_100): one full 32-byte sector per thread per instruction.
__device__ __forceinline__ void st_global_v8(void *ptr, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_v8u(const void *ptr, uint32_t (&v)[8]) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(ptr));
}
__device__ __forceinline__ void ld_global_v8(const void *ptr, float (&v)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(ptr));
}

struct TileCoord { int batch, m0, n0, it0, it1, z; };
__device__ __forceinline__ TileCoord gemm_tile(const GemmParams &p, int tile, int BN) {
    // n fastest: CTAs working at the same time share the A (activation) rows through L2
    TileCoord t;
    const int nt = tile % p.n_tiles;
    int r = tile / p.n_tiles;
    const int mt = r % p.m_tiles_total;
    t.z = r / p.m_tiles_total;
    t.batch = mt / p.m_tiles_per_batch;
    t.m0 = (mt - t.batch * p.m_tiles_per_batch) * GEMM_BM;
    t.n0 = nt * BN;
    t.it0 = t.z * p.iters_per_split;
    t.it1 = min(p.taps * p.kblocks_per_tap, t.it0 + p.iters_per_split);
    return t;
}

// Persistent: each CTA loops over tiles (tile = blockIdx.x + i*gridDim.x). The shared-memory ring
// keeps streaming across tile boundaries and the accumulator is double-buffered in TMEM, so the
// epilogue of tile i overlaps the main loop of tile i+1.
// EPI selects the epilogue at compile time:
//   EPI_GENERIC - every fused form, chosen by the run-time fields of GemmParams;
//   EPI_CONV    - the SEANet convolutions (Conv1d / ConvTranspose1d, ptts_kernels.c:128-204): + bias -> fp16 (optional copy of the
//                 pre-activation value) -> ELU or identity -> fp16, both leaving through the staging buffers as TMA tensor stores.
//                 Straight-line code per 64-column group, accumulator chunks in two alternating register sets, shared-space stores:
//                 the generic path spends ~320 instructions per 16-column chunk on these GEMMs, which are bound by exactly that.
//   EPI_RESID   - Mimi out-projection / second feed-forward linear (ptts_mimi.c:343-375): x = resid + layer_scale * acc in fp32
//                 (optional fp16 copy). The residual rows are loaded one chunk ahead, the first two before the wait for the
//                 accumulator; the layer scale comes from a shared-memory table.
//   EPI_GELU    - first feed-forward linear: tanh-GELU -> fp16.
//   EPI_ROPE    - fused QKV epilogue (RoPE + 1/8 on q, head-major fp16 outputs), rotation tables loaded one chunk ahead.
constexpr int EPI_GENERIC = 0, EPI_CONV = 1, EPI_RESID = 2, EPI_GELU = 3, EPI_ROPE = 4;

template <int BN, int SPLIT, int EPI = EPI_GENERIC>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapBlo,
               const __grid_constant__ CUtensorMap mapOH, const __grid_constant__ CUtensorMap mapOR,
               const GemmParams p) {
    using SM = GemmSmem<BN, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + p.stages * SM::STAGE_BYTES);
    uint64_t *full = bars;                          // [stages]   TMA -> MMA
    uint64_t *empty = bars + GEMM_MAX_STAGES;       // [stages]   MMA -> TMA
    // NACC accumulator buffers in TMEM, one per epilogue team: with narrow tiles (short K, the SEANet
    // convolutions) a tile's epilogue is a latency chain (tcgen05.ld -> residual load -> store), so four
    // tiles are drained concurrently by four teams of 4 warps; BN = 256 keeps two teams of 8 warps.
    constexpr int NACC = BN <= 128 ? 4 : 2;
    constexpr int ACC_COLS = BN <= 128 ? 128 : 256;
    constexpr int WPT = GEMM_EPI_WARPS / NACC;      // warps per team
    uint64_t *tfull = bars + 2 * GEMM_MAX_STAGES;   // [NACC]     MMA -> epilogue
    uint64_t *tempty = tfull + 4;                   // [NACC]     epilogue -> MMA
    uint32_t *tmem_slot = (uint32_t *)(tempty + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        if (SPLIT == 1) tma_prefetch_desc(&mapAlo);
        if (SPLIT == 2) tma_prefetch_desc(&mapBlo);
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < NACC; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], WPT); }
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
    if (threadIdx.x == 0) pdl_trigger();   // the next kernel may start its own prologue now
    const bool tr = p.trace != nullptr && blockIdx.x == 0;
    if (tr && threadIdx.x == 0) p.trace[0] = clock64();
    if constexpr (EPI == EPI_RESID) {   // layer scale (a constant tensor) -> shared memory at stage_off (host: n_total <= 512, no staging)
        float *tab = (float *)(smem + p.stage_off);
        for (int i = threadIdx.x; i < p.n_total; i += GEMM_THREADS) tab[i] = p.scale[i];
        __syncthreads();
    }

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            if (tr) p.trace[1] = clock64();
            // Pass 0 (before the dependency wait): arm the first `stages` slots and fetch their weight
            // tiles. Pass 1: everything else, in ring order.
            int pre = 0;
            if (p.b_static) {
                int s = 0;
                for (int tile = blockIdx.x; tile < p.total_tiles && pre < p.stages; tile += gridDim.x) {
                    const TileCoord t = gemm_tile(p, tile, BN);
                    for (int it = t.it0; it < t.it1 && pre < p.stages; it++, pre++, s++) {
                        uint8_t *nx = smem + s * SM::STAGE_BYTES + SM::A_BYTES * (SPLIT == 1 ? 2 : 1);
                        mbar_expect_tx(&full[s], SM::STAGE_BYTES);
                        const int tap = it / p.kblocks_per_tap, kb = it - tap * p.kblocks_per_tap;
                        const int bk = kb * GEMM_BK + tap * p.b_k_step;
                        tma_load_3d(nx, &mapB, &full[s], bk, t.n0, 0);
                        if (SPLIT == 2) tma_load_3d(nx + SM::B_BYTES, &mapBlo, &full[s], bk, t.n0, 0);
                    }
                }
            }
            pdl_wait();
            int s = 0, n = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord t = gemm_tile(p, tile, BN);
                for (int it = t.it0; it < t.it1; it++, n++) {
                    uint8_t *st = smem + s * SM::STAGE_BYTES;
                    const int tap = it / p.kblocks_per_tap, kb = it - tap * p.kblocks_per_tap;
                    const int ak = kb * GEMM_BK + tap * p.a_k_step;
                    const int ar = t.m0 + p.a_row_shift0 + tap * p.a_row_step;
                    if (n >= pre) {
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], SM::STAGE_BYTES);
                    }
                    tma_load_3d(st, &mapA, &full[s], ak, ar, t.batch);
                    uint8_t *nx = st + SM::A_BYTES;
                    if (SPLIT == 1) { tma_load_3d(nx, &mapAlo, &full[s], ak, ar, t.batch); nx += SM::A_BYTES; }
                    if (n >= pre) {
                        const int bk = kb * GEMM_BK + tap * p.b_k_step;
                        tma_load_3d(nx, &mapB, &full[s], bk, t.n0, 0);
                        if (SPLIT == 2) tma_load_3d(nx + SM::B_BYTES, &mapBlo, &full[s], bk, t.n0, 0);
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
            if (tr) p.trace[2] = clock64();
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one elected lane)
        int s = 0; uint32_t ph = 0; int lt = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, lt++) {
            const TileCoord t = gemm_tile(p, tile, BN);
            const int acc = lt % NACC;
            mbar_wait(&tempty[acc], ((lt / NACC) & 1) ^ 1);   // its team has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
            for (int it = t.it0; it < t.it1; it++) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (tr && lane == 0 && it == t.it0 && lt == 0) p.trace[3] = clock64();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smem + s * SM::STAGE_BYTES);
                    const uint32_t alo_addr = a_addr + SM::A_BYTES;
                    const uint32_t b_addr = a_addr + SM::A_BYTES * (SPLIT == 1 ? 2 : 1);
                    const uint32_t blo_addr = b_addr + SM::B_BYTES;
                    const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(b_addr);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; k++) {
                        // advancing 16 elements (32 B) along K inside the swizzled row: +2 in 16-byte units
                        umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, p.idesc, (it > t.it0 || k > 0) ? 1u : 0u);
                        if (SPLIT == 1) umma_f16(d_tmem, umma_desc_sw128(alo_addr) + 2 * k, bd + 2 * k, p.idesc, 1u);
                        if (SPLIT == 2) umma_f16(d_tmem, ad + 2 * k, umma_desc_sw128(blo_addr) + 2 * k, p.idesc, 1u);
                    }
                    umma_commit(&empty[s]);                       // smem slot reusable once these MMAs retire
                    if (it == t.it1 - 1) umma_commit(&tfull[acc]);  // accumulator complete
                    if (tr && it == t.it1 - 1) p.trace[4] = clock64();
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------ epilogue warps
        // TMEM lane quadrant = warp % 4 (hardware rule); the 4 warps of a quadrant interleave the
        // 16-column chunks. tcgen05.ld of the next chunk is in flight while this one is processed.
        const int team = (warp - 2) / WPT, quad = warp & 3, cg = ((warp - 2) % WPT) >> 2;
        constexpr int NCG = WPT / 4;
        pdl_wait();   // residual / outputs belong to earlier kernels
        const bool bf = p.out_bf16 != 0;
        const int act = p.act, ldo = p.ldo, ldh = p.ldh, n_total = p.n_total, bias_mask = p.bias_mask;
        const bool fast = p.fast_act != 0;
        const float *bias = p.bias, *scale = p.scale, *resid = p.resid;
        uint16_t *out_h = p.out_h, *out_l = p.out_l;
        // team `team` owns accumulator `team` and the CTA's tiles team, team + NACC, ...
        for (int lt = team, tile = blockIdx.x + team * gridDim.x; tile < p.total_tiles; tile += NACC * gridDim.x, lt += NACC) {
            const TileCoord t = gemm_tile(p, tile, BN);
            const int acc = team;
            if (p.tma_store) {   // the previous tile's last TMA store has finished reading this warp's staging buffer
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
            }
            // (the specialised epilogues issue their first global loads before this wait)
            auto wait_acc = [&]() {
                mbar_wait(&tfull[acc], (lt / NACC) & 1);
                tc_fence_after();
                if (tr && lt == 0 && warp == 2 && lane == 0) p.trace[5] = clock64();
            };
            const int m = t.m0 + quad * 32 + lane;
            const bool m_ok = m < p.a_rows;
            const uint32_t trow = tmem_base + acc * ACC_COLS + ((uint32_t)(quad * 32) << 16);
            float *out_f32 = p.out_f32 ? p.out_f32 + (size_t)t.z * p.split_stride : nullptr;
            const bool add_bias = bias != nullptr && t.z == 0;
            uint32_t r[16], rn[16];
            if constexpr (EPI == EPI_RESID) {
                constexpr int CW = BN / NCG, NCH = CW / 16;
                static_assert(NCH % 2 == 0, "EPI_RESID: chunks are processed in pairs");
                const int c_lo = cg * CW, nb_lo = t.n0 + c_lo;
                const size_t grow = (size_t)t.batch * p.a_rows + m;
                const float *rs = resid + grow * ldo + nb_lo;
                float *of = p.out_f32 + grow * ldo + nb_lo;
                uint16_t *oh = out_h ? out_h + grow * ldh + nb_lo : nullptr;
                const float *tab = (const float *)(smem + p.stage_off) + nb_lo;
                float ta[16], tb[16];
                auto ld_res = [&](int i, float(&tt)[16]) {
                    if (m_ok && nb_lo + 16 * i < n_total) {
                        float a0[8], a1[8];
                        ld_global_v8(rs + 16 * i, a0); ld_global_v8(rs + 16 * i + 8, a1);
#pragma unroll
                        for (int j = 0; j < 8; j++) { tt[j] = a0[j]; tt[8 + j] = a1[j]; }
                    }
                };
                auto step = [&](int i, uint32_t(&acc)[16], float(&tt)[16]) {
                    if (m_ok && nb_lo + 16 * i < n_total) {
                        float v[16];
                        const float4 *sp = (const float4 *)(tab + 16 * i);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float4 ss = sp[j];
                            // separately rounded product and sum, like the generic epilogue (its two steps sit in different branches and
                            // are never contracted): the specialised path stays bit-identical to it, which a test pins
                            v[4 * j] = __fadd_rn(__fmul_rn(__uint_as_float(acc[4 * j]), ss.x), tt[4 * j]);
                            v[4 * j + 1] = __fadd_rn(__fmul_rn(__uint_as_float(acc[4 * j + 1]), ss.y), tt[4 * j + 1]);
                            v[4 * j + 2] = __fadd_rn(__fmul_rn(__uint_as_float(acc[4 * j + 2]), ss.z), tt[4 * j + 2]);
                            v[4 * j + 3] = __fadd_rn(__fmul_rn(__uint_as_float(acc[4 * j + 3]), ss.w), tt[4 * j + 3]);
                        }
                        if (i + 2 < NCH) ld_res(i + 2, tt);   // this register set's next chunk
                        uint32_t w0[8], w1[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) { w0[j] = __float_as_uint(v[j]); w1[j] = __float_as_uint(v[8 + j]); }
                        st_global_v8(of + 16 * i, w0);
                        st_global_v8(of + 16 * i + 8, w1);
                        if (oh) {
                            uint32_t hw[8], lw[8];
                            act_pack16<ACT_NONE, true, false, false>(v, hw, lw);
                            st_global_v8(oh + 16 * i, hw);
                        }
                    }
                };
                ld_res(0, ta); ld_res(1, tb);
                wait_acc();
                tmem_ld16(trow + c_lo, r);
#pragma unroll 1
                for (int i = 0; i < NCH; i += 2) {
                    tmem_ld_wait16(r);
                    tmem_ld16(trow + c_lo + 16 * (i + 1), rn);
                    step(i, r, ta);
                    tmem_ld_wait16(rn);
                    if (i + 2 < NCH) tmem_ld16(trow + c_lo + 16 * (i + 2), r);
                    step(i + 1, rn, tb);
                }
            } else if constexpr (EPI == EPI_GELU) {
                constexpr int CW = BN / NCG, NCH = CW / 16;
                const int c_lo = cg * CW, nb_lo = t.n0 + c_lo;
                uint16_t *oh = out_h + ((size_t)t.batch * p.a_rows + m) * ldh + nb_lo;
                auto step = [&](int i, uint32_t(&acc)[16]) {
                    if (m_ok && nb_lo + 16 * i < n_total) {
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) v[j] = __uint_as_float(acc[j]);
                        uint32_t hw[8], lw[8];
                        act_pack16<ACT_GELU_TANH, true, false, false>(v, hw, lw);
                        st_global_v8(oh + 16 * i, hw);
                    }
                };
                wait_acc();
                tmem_ld16(trow + c_lo, r);
#pragma unroll 1
                for (int i = 0; i < NCH; i += 2) {
                    tmem_ld_wait16(r);
                    tmem_ld16(trow + c_lo + 16 * (i + 1), rn);
                    step(i, r);
                    tmem_ld_wait16(rn);
                    if (i + 2 < NCH) tmem_ld16(trow + c_lo + 16 * (i + 2), r);
                    step(i + 1, rn);
                }
            } else if constexpr (EPI == EPI_ROPE) {
                constexpr int CW = BN / NCG, NCH = CW / 16;
                const int c_lo = cg * CW, nb_lo = t.n0 + c_lo;
                const int which = nb_lo >> 9;                       // 0 q, 1 k, 2 v: uniform per warp (CW divides 512)
                const float sc = which == 0 ? 0.125f : 1.0f;        // 1/sqrt(64) folded into q
                const float *cp = p.rope_cos + (size_t)(m + p.rope_pos0) * 32, *sp = p.rope_sin + (size_t)(m + p.rope_pos0) * 32;
                uint16_t *base = (which == 0 ? p.rope_q : which == 1 ? p.rope_k : p.rope_v);
                const size_t hrows = (size_t)(p.rope_rows ? p.rope_rows : p.a_rows);
                float ca[8], sa[8], cb[8], sb[8];
                auto ld_tab = [&](int i, float(&cs)[8], float(&sn)[8]) {
                    const int d0 = (nb_lo + 16 * i) & 63;
                    if (which < 2 && m_ok) { ld_global_v8(cp + (d0 >> 1), cs); ld_global_v8(sp + (d0 >> 1), sn); }
                };
                auto step = [&](int i, uint32_t(&acc)[16], float(&cs)[8], float(&sn)[8]) {
                    const int nb = nb_lo + 16 * i, within = nb & 511, head = within >> 6, d0 = within & 63;
                    if (m_ok && nb < n_total) {
                        uint32_t hw[8];
                        if (which < 2) {
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const float a0 = __uint_as_float(acc[2 * j]), a1 = __uint_as_float(acc[2 * j + 1]);
                                const __half2 h2 = __floats2half2_rn((a0 * cs[j] - a1 * sn[j]) * sc, (a0 * sn[j] + a1 * cs[j]) * sc);
                                hw[j] = *(const uint32_t *)&h2;
                            }
                            if (i + 2 < NCH) ld_tab(i + 2, cs, sn);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const __half2 h2 = __floats2half2_rn(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
                                hw[j] = *(const uint32_t *)&h2;
                            }
                        }
                        st_global_v8(base + (((size_t)t.batch * 8 + head) * hrows + m + p.rope_row0) * 64 + d0, hw);
                    }
                };
                ld_tab(0, ca, sa); ld_tab(1, cb, sb);
                wait_acc();
                tmem_ld16(trow + c_lo, r);
#pragma unroll 1
                for (int i = 0; i < NCH; i += 2) {
                    tmem_ld_wait16(r);
                    tmem_ld16(trow + c_lo + 16 * (i + 1), rn);
                    step(i, r, ca, sa);
                    tmem_ld_wait16(rn);
                    if (i + 2 < NCH) tmem_ld16(trow + c_lo + 16 * (i + 2), r);
                    step(i + 1, rn, cb, sb);
                }
            } else if constexpr (EPI == EPI_CONV) {
                wait_acc();
                constexpr int CW = BN / NCG;
                static_assert(CW % 64 == 0, "EPI_CONV: a warp owns whole 64-column groups");
                const int c_lo = cg * CW;
                // this lane's row of the warp's staging buffers ([32 rows][128 B] per tensor, 16-byte piece index ^ (row & 7))
                const uint32_t row_s = smem_u32(smem) + (uint32_t)(p.stage_off + (warp - 2) * p.stage_stride + lane * 128);
                const uint32_t x7 = (uint32_t)(lane & 7) << 4;
                const bool raw = p.out_raw != nullptr, elu = act == ACT_ELU;
                tmem_ld16(trow + c_lo, r);
#pragma unroll 1
                for (int g = 0; g < CW / 64; g++) {
                    const int cb = c_lo + g * 64, nb0 = t.n0 + cb;
                    const bool live = nb0 < n_total;
                    if (g > 0) {   // the previous group's tensor stores have read the staging buffers
                        if (lane == 0) bulk_wait_read0();
                        __syncwarp();
                    }
                    const float4 *bp = (const float4 *)(bias + (bias_mask ? (nb0 & bias_mask) : nb0));
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        uint32_t(&cur)[16] = (q & 1) ? rn : r;
                        uint32_t(&nxt)[16] = (q & 1) ? r : rn;
                        tmem_ld_wait16(cur);
                        if (q < 3 || g + 1 < CW / 64) tmem_ld16(trow + cb + (q + 1) * 16, nxt);   // in flight while this chunk is processed
                        if (live) {
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const float4 bb = bp[4 * q + j];
                                v[4 * j] = __uint_as_float(cur[4 * j]) + bb.x; v[4 * j + 1] = __uint_as_float(cur[4 * j + 1]) + bb.y;
                                v[4 * j + 2] = __uint_as_float(cur[4 * j + 2]) + bb.z; v[4 * j + 3] = __uint_as_float(cur[4 * j + 3]) + bb.w;
                            }
                            const uint32_t o0 = (uint32_t)(q << 5) ^ x7, o1 = o0 ^ 16u;
                            uint32_t hw[8], lw[8];
                            if (raw) {
                                act_pack16<ACT_NONE, true, false, false>(v, hw, lw);
                                sts128(row_s + 4096 + o0, hw[0], hw[1], hw[2], hw[3]);
                                sts128(row_s + 4096 + o1, hw[4], hw[5], hw[6], hw[7]);
                            }
                            if (elu) act_pack16<ACT_ELU, true, false, false>(v, hw, lw);
                            else act_pack16<ACT_NONE, true, false, false>(v, hw, lw);
                            sts128(row_s + o0, hw[0], hw[1], hw[2], hw[3]);
                            sts128(row_s + o1, hw[4], hw[5], hw[6], hw[7]);
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && live) {
                        const uint32_t stg_s = row_s;   // lane 0: the buffer's first row
                        const int row0 = t.m0 + quad * 32;
                        tma_store_3d_s(&mapOH, stg_s, nb0, row0, t.batch);
                        if (raw) tma_store_3d_s(&mapOR, stg_s + 4096, nb0, row0, t.batch);
                        bulk_commit();
                    }
                }
            } else if (!p.transposed) {
                wait_acc();
                // each warp owns a contiguous range of BN / NCG columns of its 32 rows
                constexpr int CW = BN / NCG;
                const int c_lo = cg * CW;
                if (CW >= 16) { tmem_ld16(trow + c_lo, r); tmem_ld_wait16(r); }
                const size_t grow = (size_t)t.batch * p.a_rows + m;
                float *of = out_f32 ? out_f32 + grow * ldo : nullptr;
                const float *rs = resid ? resid + grow * ldo : nullptr;
                uint16_t *oh = out_h ? out_h + grow * ldh : nullptr;
                uint16_t *ol = out_l ? out_l + grow * ldh : nullptr;
                const bool ts = p.tma_store != 0;
                // staging: [32 rows][128 bytes] per output tensor, 128B-swizzled (16-byte piece index ^ (row & 7)): the
                // 8 rows of a quarter-warp write 8 different pieces -> conflict-free; the TMA store undoes the swizzle
                uint8_t *stg = smem + p.stage_off + (warp - 2) * p.stage_stride;
#pragma unroll 1
                for (int c = c_lo; c < c_lo + CW; c += 16) {
                    const bool more = c + 16 < c_lo + CW;
                    if (more) tmem_ld16(trow + c + 16, rn);   // in flight while this chunk is processed
                    const int nb = t.n0 + c;
                    if (ts && (c & 63) == 0 && c > c_lo) {   // previous 64-column group left through TMA: staging reusable?
                        if (lane == 0) bulk_wait_read0();
                        __syncwarp();
                    }
                    if ((m_ok || ts) && nb < n_total) {
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]);
                        if (add_bias) {
                            const float4 *bp = (const float4 *)(bias + (bias_mask ? (nb & bias_mask) : nb));
#pragma unroll
                            for (int j = 0; j < 4; j++) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
                        }
                        if (scale) {
                            const float4 *sp = (const float4 *)(scale + nb);
#pragma unroll
                            for (int j = 0; j < 4; j++) { const float4 ss = sp[j]; v[4 * j] *= ss.x; v[4 * j + 1] *= ss.y; v[4 * j + 2] *= ss.z; v[4 * j + 3] *= ss.w; }
                        }
                        if (p.rope_q) {
                            const int which = nb >> 9, within = nb & 511, head = within >> 6, d0 = within & 63;
                            uint32_t hw[8];
                            if (which < 2) {
                                float cs[8], sn[8];
                                ld_global_v8(p.rope_cos + (size_t)(m + p.rope_pos0) * 32 + (d0 >> 1), cs);
                                ld_global_v8(p.rope_sin + (size_t)(m + p.rope_pos0) * 32 + (d0 >> 1), sn);
                                const float sc = which == 0 ? 0.125f : 1.0f;   // 1/sqrt(64) folded into q
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const float a0 = v[2 * j], a1 = v[2 * j + 1];
                                    const __half2 h2 = __floats2half2_rn((a0 * cs[j] - a1 * sn[j]) * sc, (a0 * sn[j] + a1 * cs[j]) * sc);
                                    hw[j] = *(const uint32_t *)&h2;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; j++) { const __half2 h2 = __floats2half2_rn(v[2 * j], v[2 * j + 1]); hw[j] = *(const uint32_t *)&h2; }
                            }
                            uint16_t *dst = (which == 0 ? p.rope_q : which == 1 ? p.rope_k : p.rope_v) +
                                            (((size_t)t.batch * 8 + head) * (p.rope_rows ? p.rope_rows : p.a_rows) + m + p.rope_row0) * 64 + d0;
                            st_global_v8(dst, hw);
                        }
                        if (rs && m_ok) {
                            float t0[8], t1[8];
                            ld_global_v8(rs + nb, t0);
                            ld_global_v8(rs + nb + 8, t1);
#pragma unroll
                            for (int j = 0; j < 8; j++) { v[j] += t0[j]; v[8 + j] += t1[j]; }
                        }
                        if (p.resid_h && m_ok) {
                            uint32_t hr[8];
                            ld_global_v8u(p.resid_h + grow * ldh + nb, hr);
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const float2 f = bf ? __bfloat1622float2(*(const __nv_bfloat162 *)&hr[j]) : __half22float2(*(const __half2 *)&hr[j]);
                                v[2 * j] += f.x; v[2 * j + 1] += f.y;
                            }
                        }
                        if (of && m_ok) {
                            uint32_t w0[8], w1[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) { w0[j] = __float_as_uint(v[j]); w1[j] = __float_as_uint(v[8 + j]); }
                            st_global_v8(of + nb, w0);
                            st_global_v8(of + nb + 8, w1);
                        }
                        const int pc = (c & 63) >> 3;   // first 16-byte piece of this chunk inside the 128-byte staging row
                        if (p.out_raw) {
                            uint32_t hw[8], lw[8];
                            act_pack16_dispatch(ACT_NONE, fast, bf, false, v, hw, lw);
                            if (ts) {
                                uint8_t *row = stg + 4096 + lane * 128;
                                *(uint4 *)(row + ((pc ^ (lane & 7)) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                *(uint4 *)(row + (((pc + 1) ^ (lane & 7)) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                            } else st_global_v8(p.out_raw + grow * ldh + nb, hw);
                        }
                        if (oh) {
                            uint32_t hw[8], lw[8];
                            act_pack16_dispatch(act, fast, bf, ol != nullptr, v, hw, lw);
                            if (ts) {
                                uint8_t *row = stg + lane * 128;
                                *(uint4 *)(row + ((pc ^ (lane & 7)) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                *(uint4 *)(row + (((pc + 1) ^ (lane & 7)) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                            } else {
                                st_global_v8(oh + nb, hw);
                                if (ol) st_global_v8(ol + nb, lw);
                            }
                        }
                    }
                    if (ts && ((c & 63) == 48 || !more)) {   // a 64-column group of this warp's 32 rows is staged: hand it to TMA
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0 && t.n0 + (c & ~63) < n_total) {
                            const int col0 = t.n0 + (c & ~63), row0 = t.m0 + quad * 32;
                            if (oh) tma_store_3d(&mapOH, stg, col0, row0, t.batch);
                            if (p.out_raw) tma_store_3d(&mapOR, stg + 4096, col0, row0, t.batch);
                            bulk_commit();
                        }
                    }
                    if (more) {
                        tmem_ld_wait16(rn);
#pragma unroll
                        for (int j = 0; j < 16; j++) r[j] = rn[j];
                    }
                }
            } else {
                wait_acc();
                // transposed: this thread owns output FEATURE m; columns are batch rows n. Stores are
                // coalesced across the warp (32 consecutive features per batch row).
                if (cg * 16 < BN) { tmem_ld16(trow + cg * 16, r); tmem_ld_wait16(r); }
                float bv = 0.0f, sv = 1.0f;
                if (m_ok) {
                    if (add_bias) bv = bias[bias_mask ? (m & bias_mask) : m];
                    if (scale) sv = scale[m];
                }
                const int ncols = min(BN, n_total - t.n0);
#pragma unroll 1
                for (int c = cg * 16; c < BN; c += 16 * NCG) {
                    const bool more = c + 16 * NCG < BN;
                    if (more) tmem_ld16(trow + c + 16 * NCG, rn);
                    if (m_ok && c < ncols) {
                        const size_t o0 = (size_t)(t.n0 + c) * ldo + m, h0i = (size_t)(t.n0 + c) * ldh + m;
                        float *of = out_f32 ? out_f32 + o0 : nullptr;
                        const float *rs = resid ? resid + o0 : nullptr;
                        uint16_t *oh = out_h ? out_h + h0i : nullptr;
                        uint16_t *ol = out_l ? out_l + h0i : nullptr;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            if (c + j < ncols) {
                                float x = (__uint_as_float(r[j]) + bv) * sv;
                                if (rs) x += rs[(size_t)j * ldo];
                                if (of) of[(size_t)j * ldo] = x;
                                if (oh) {
                                    const float a = act_apply(act, x);
                                    const uint16_t h = f2h16(a, bf);
                                    oh[(size_t)j * ldh] = h;
                                    if (ol) ol[(size_t)j * ldh] = f2h16(a - h162f(h, bf), bf);
                                }
                            }
                        }
                    }
                    if (more) {
                        tmem_ld_wait16(rn);
#pragma unroll
                        for (int j = 0; j < 16; j++) r[j] = rn[j];
                    }
                }
            }
            if (cg * 16 >= BN) { /* this warp had no chunk of a narrow tile */ }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);   // all epilogue warps -> accumulator reusable
            if (tr && lt == 0 && warp == 2 && lane == 0) p.trace[6] = clock64();
        }
    }
    if (p.tma_store && warp >= 2 && lane == 0) bulk_wait0();   // this thread's tensor stores have landed
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if (tr && threadIdx.x == 0) p.trace[7] = clock64();
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
    CUtensorMap mapA, mapAlo, mapB, mapBlo, mapOH, mapOR;
    GemmParams p;
    int bn;        // N tile (B-operand box rows)
    int split;     // 0 none, 1 = A has lo, 2 = B has lo
    int epi;       // EPI_GENERIC / EPI_CONV
    dim3 grid;       // persistent: min(total_tiles, max_ctas)
    int smem_bytes;
};

int gemm_setup_attrs();
int gemm_launch(const GemmOp &op, cudaStream_t st);
int gemm_pick_stages(int bn, int split, int iters, bool two_ctas_per_sm);

}  // namespace ptts
