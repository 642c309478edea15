// resblock.cuh - one SEANet residual block (ptts_mimi.c:160-177) as ONE tcgen05 kernel:
//     out = ELU( u + W2 . ELU( W1 (*) ELU(u) + b1 ) + b2 )          (*) = causal Conv1d, k = 3
// The middle tensor (C/2 channels at the stage's sample rate - the single largest SEANet intermediate
// per output byte) never leaves the SM: the first contraction's accumulator is read from TMEM by the
// epilogue warps, biased, ELU'd, rounded to fp16 and written straight into shared memory in the
// K-major 128B-swizzled layout the tensor core reads, where it becomes the A operand of the second
// contraction (W2 stays resident in shared memory for the whole kernel).
//
//   warp 0    : TMA producer - the 3 taps of ELU(u) (shifted row coordinates, OOB rows = causal zero
//               padding) and the matching W1 slices through a shared-memory ring
//   warp 1    : MMA issuer   - contraction 1 into TMEM D1, then (once the epilogue has staged the middle
//               tensor) contraction 2 into TMEM D2
//   warps 2-9 : epilogue     - D1 -> +b1 -> ELU -> fp16 -> shared memory;  D2 -> +b2 +u -> ELU -> fp16 -> HBM
// The last block (64 channels) also applies the decoder's final Conv1d 64 -> 1 (k = 3) as a THIRD contraction: its output tile
// goes back into the `mid` operand buffer, D3[128, 16] = out . W3^T (rows of W3 = the three taps as fp16 hi + lo), and the
// epilogue only combines three numbers per row with the two rows before it.
// Persistent over 128-row tiles; contraction 1 of tile i+1 overlaps the output epilogue of tile i.
#pragma once
#include "gemm_tc.cuh"

namespace ptts {

constexpr int RB_EPI_WARPS = 8;              // two per TMEM lane quadrant: each takes half of the columns
constexpr int RB_THREADS = 64 + 32 * RB_EPI_WARPS;

struct ResParams {
    int tiles_per_batch, tiles_total, a_rows;
    int kb_per_tap;          // C2 / 64
    int chp;                 // middle channels, padded to a multiple of 64
    int stages;
    int prefetch_u;          // 1: L2 prefetch of each tile's residual rows by the producer (A/B switch)
    uint32_t idesc1, idesc2, idesc3;   // idesc3: the fused final conv (N = 16)
    const float *b1, *b2;    // [chp] (zero padded), [C2]
    const uint16_t *u;       // residual stream, fp16 [rows][C2]
    uint16_t *out;           // fp16 [rows][C2] (NULL when the final conv is fused)
    // last block only (C2 = 64): the decoder's final Conv1d 64 -> 1, k = 3 (ptts_mimi.c:723-726) is applied to the block's
    // output rows while they are still in registers: pcm[t] = bias + w2.e[t] + w1.e[t-1] + w0.e[t-2]. The two rows before a
    // tile belong to another tile: their contributions are left in `edge` and added by k_dec_edge_fix.
    const float *dw;         // [64][3] (NULL = not fused)
    float dbias;
    float *pcm;              // [batches][a_rows]
    float *edge;             // [tiles][4]: w1.e[127], w0.e[127], w0.e[126] of each tile
};

template <int C2>
__global__ void __launch_bounds__(RB_THREADS, C2 <= 128 ? 2 : 1)
resblock_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW1,
                const __grid_constant__ CUtensorMap mapW2, const ResParams p) {
    constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;                 // 16 KB
    constexpr uint32_t D2_COL = C2 == 64 ? 64 : 128;
    constexpr uint32_t TMEM_COLS = C2 == 64 ? 256 : (C2 == 128 ? 256 : 512);   // C2 == 64: D1 at 0, D2 at 64, D3 (fused final conv, 16 columns) at 128
    constexpr uint32_t D3_COL = 128;
    extern __shared__ uint8_t rb_smem_raw[];
    uint8_t *smem = rb_smem_raw + ((1024u - (smem_u32(rb_smem_raw) & 1023u)) & 1023u);
    const int chp = p.chp, kb2 = chp / 64;
    const int w1_bytes = chp * 128, stage_bytes = A_BYTES + w1_bytes;
    uint8_t *ring = smem;
    uint8_t *w2s = ring + p.stages * stage_bytes;                 // [kb2][C2 rows][128 B]
    uint8_t *a2s = w2s + kb2 * C2 * 128;                           // [kb2][128 rows][128 B]
    uint8_t *a3s = a2s + kb2 * A_BYTES;                            // [128 rows][128 B]: the block's output tile, A operand of the fused final conv (C2 == 64)
    uint8_t *w3s = a3s + (C2 == 64 ? A_BYTES : 0);                 // [16 rows][128 B]: taps of the fused final conv as a B operand (C2 == 64)
    uint64_t *bars = (uint64_t *)(w3s + (C2 == 64 ? 2048 : 0));   // the wider blocks sit at the two-CTAs-per-SM limit: no room, no fused conv
    uint64_t *full = bars, *empty = bars + GEMM_MAX_STAGES;
    uint64_t *w2_full = bars + 2 * GEMM_MAX_STAGES, *d1_full = w2_full + 1, *a2_full = w2_full + 2, *d2_full = w2_full + 3,
             *d2_empty = w2_full + 4, *a3_full = w2_full + 5, *d3_full = w2_full + 6;
    uint32_t *tmem_slot = (uint32_t *)(w2_full + 7);
    float *dpart = (float *)(bars + 32);          // [2 buffers][128 rows][3]: per-row tap sums of the fused final conv
    float *b1s = dpart + 768, *b2s = b1s + 128;   // biases (every lane reads the same address: one broadcast LDS instead of a global load)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool fuse_out = C2 == 64 && p.dw != nullptr;
    if (fuse_out) {
        // W3[n][c]: rows 0..2 = fp16(w[c][tap n]), rows 3..5 = the rounding remainders (their products are added in the epilogue),
        // rows 6..15 zero; K-major, 128B swizzle like every other B operand
        for (int idx = threadIdx.x; idx < 16 * 64; idx += RB_THREADS) {
            const int k = idx >> 6, c = idx & 63;
            float w = 0.0f;
            if (k < 6) {
                const float wf = p.dw[c * 3 + (k % 3)];
                const float hi = __half2float(__float2half_rn(wf));
                w = k < 3 ? hi : wf - hi;
            }
            *(__half *)(w3s + k * 128 + ((((c >> 3) ^ (k & 7))) << 4) + (c & 7) * 2) = __float2half_rn(w);
        }
        fence_proxy_async();
    }

    for (int i = threadIdx.x; i < p.chp; i += RB_THREADS) b1s[i] = p.b1[i];
    for (int i = threadIdx.x; i < C2; i += RB_THREADS) b2s[i] = p.b2[i];
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapW1); tma_prefetch_desc(&mapW2);
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(w2_full, 1); mbar_init(d1_full, 1); mbar_init(a2_full, RB_EPI_WARPS); mbar_init(d2_full, 1); mbar_init(d2_empty, RB_EPI_WARPS);
        mbar_init(a3_full, RB_EPI_WARPS); mbar_init(d3_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pdl_trigger();
    const int iters = 3 * p.kb_per_tap;

    if (warp == 0) {
        if (lane == 0) {
            // weights are constants: W2 (resident) and the first ring slots' W1 slices go out before the dependency wait
            mbar_expect_tx(w2_full, (uint32_t)(kb2 * C2 * 128));
            for (int kb = 0; kb < kb2; kb++) tma_load_3d(w2s + kb * C2 * 128, &mapW2, w2_full, kb * GEMM_BK, 0, 0);
            pdl_wait();
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
                const int batch = tile / p.tiles_per_batch, m0 = (tile - batch * p.tiles_per_batch) * GEMM_BM;
                if (p.prefetch_u) {   // the tile's residual rows (read by the output epilogue with plain loads): towards L2 now
                    const int nr = min(GEMM_BM, p.a_rows - m0);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.u + ((size_t)batch * p.a_rows + m0) * C2), "r"((uint32_t)nr * C2 * 2u) : "memory");
                }
                for (int it = 0; it < iters; it++) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                    const int tap = it / p.kb_per_tap, kb = it - tap * p.kb_per_tap;
                    uint8_t *st = ring + s * stage_bytes;
                    tma_load_3d(st, &mapA, &full[s], kb * GEMM_BK, m0 - 2 + tap, batch);          // row t-2+tap: causal, k = 3
                    tma_load_3d(st + A_BYTES, &mapW1, &full[s], it * GEMM_BK, 0, 0);             // W1 columns tap*C2 + kb*64
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        int s = 0; uint32_t ph = 0; int lt = 0;
        mbar_wait(w2_full, 0);
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, lt++) {
            // contraction 1: D1[128, chp] = sum over taps, channels
            for (int it = 0; it < iters; it++) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(ring + s * stage_bytes);
                    const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; k++) umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, p.idesc1, (it > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty[s]);
                    if (it == iters - 1) umma_commit(d1_full);
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            // contraction 3 of the PREVIOUS tile (fused final conv): D3[128, 16] = out[128, 64] . W3^T, once its output epilogue has
            // written the tile into the operand buffer; issued here so that contraction 1 of this tile is already running
            if (fuse_out && lt > 0) {
                mbar_wait(a3_full, (lt - 1) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint64_t ad = umma_desc_sw128(smem_u32(a3s)), bd = umma_desc_sw128(smem_u32(w3s));
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; k++) umma_f16(tmem_base + D3_COL, ad + 2 * k, bd + 2 * k, p.idesc3, k > 0 ? 1u : 0u);
                    umma_commit(d3_full);
                }
                __syncwarp();
            }
            // contraction 2: D2[128, C2] = mid[128, chp] . W2^T, once the epilogue has staged `mid` and drained the previous D2
            mbar_wait(a2_full, lt & 1);
            mbar_wait(d2_empty, (lt & 1) ^ 1);
            tc_fence_after();
            if (lane == 0) {
                for (int kb = 0; kb < kb2; kb++) {
                    const uint64_t ad = umma_desc_sw128(smem_u32(a2s + kb * A_BYTES)), bd = umma_desc_sw128(smem_u32(w2s + kb * C2 * 128));
                    constexpr int KS = (C2 / 2 < GEMM_BK ? C2 / 2 : GEMM_BK) / 16;   // 64-channel block: 32 real middle channels = 2 k-steps
#pragma unroll
                    for (int k = 0; k < KS; k++) umma_f16(tmem_base + D2_COL, ad + 2 * k, bd + 2 * k, p.idesc2, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(d2_full);
            }
            __syncwarp();
        }
        if (fuse_out && lt > 0) {   // contraction 3 of the last tile
            mbar_wait(a3_full, (lt - 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t ad = umma_desc_sw128(smem_u32(a3s)), bd = umma_desc_sw128(smem_u32(w3s));
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; k++) umma_f16(tmem_base + D3_COL, ad + 2 * k, bd + 2 * k, p.idesc3, k > 0 ? 1u : 0u);
                umma_commit(d3_full);
            }
            __syncwarp();
        }
    } else {
        const int quad = warp & 3, cg = (warp - 2) >> 2;   // column half
        // the middle tensor has C2 / 2 real channels (chp pads them to a multiple of 64 for the operand layout only): contraction 1
        // produces, epilogue 1 converts and contraction 2 reads just those
        constexpr int CH = C2 / 2;
        const int c1_lo = cg * (CH / 2), c1_hi = c1_lo + CH / 2, c2_lo = cg * (C2 / 2), c2_hi = c2_lo + C2 / 2;
        pdl_wait();
        int lt = 0;
        const int r = quad * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
        // epilogue 3 of tile `tl` (the lt3-th of this CTA): D3 -> per-row tap sums -> pcm. Runs one tile LATE (between epilogue 1 and the
        // wait for contraction 2 of the next tile), so neither the round trip of contraction 3 nor that of contraction 2 is idle time.
        auto epilogue3 = [&](int tl, int lt3) {
            mbar_wait(d3_full, lt3 & 1);   // every epilogue warp: the output-tile operand buffer is free again after this
            tc_fence_after();
            if (cg == 0) {
                const int b3 = tl / p.tiles_per_batch, m3 = (tl - b3 * p.tiles_per_batch) * GEMM_BM + r;
                uint32_t y3[16];
                tmem_ld16(trow + D3_COL, y3); tmem_ld_wait16(y3);
                float *pb = dpart + (lt3 & 1) * 384;               // double-buffered: one barrier per tile is enough
                pb[r * 3 + 0] = __uint_as_float(y3[0]) + __uint_as_float(y3[3]);
                pb[r * 3 + 1] = __uint_as_float(y3[1]) + __uint_as_float(y3[4]);
                pb[r * 3 + 2] = __uint_as_float(y3[2]) + __uint_as_float(y3[5]);
                asm volatile("bar.sync 1, 128;" ::: "memory");    // the 4 warps that hold the 128 rows
                auto D = [&](int row, int k) { return pb[row * 3 + k]; };
                float y = p.dbias + D(r, 2);
                if (r >= 1) y += D(r - 1, 1);
                if (r >= 2) y += D(r - 2, 0);
                if (m3 < p.a_rows) p.pcm[(size_t)b3 * p.a_rows + m3] = y;
                if (r == 127) { float *e = p.edge + (size_t)tl * 4; e[0] = D(127, 1); e[1] = D(127, 0); e[2] = D(126, 0); }
            }
            tc_fence_before();
        };
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, lt++) {
            const int batch = tile / p.tiles_per_batch, m0 = (tile - batch * p.tiles_per_batch) * GEMM_BM;
            const int m = m0 + r;
            // ---- epilogue 1: D1 -> +b1 -> ELU -> fp16 -> shared memory (A operand of contraction 2)
            mbar_wait(d1_full, lt & 1);
            tc_fence_after();
            {
                uint32_t rr[16], rn[16];
                tmem_ld16(trow + c1_lo, rr); tmem_ld_wait16(rr);
#pragma unroll 1
                for (int c = c1_lo; c < c1_hi; c += 16) {
                    const bool more = c + 16 < c1_hi;
                    if (more) tmem_ld16(trow + c + 16, rn);
                    float v[16];
                    const float4 *bp = (const float4 *)(b1s + c);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float4 bb = bp[j];
                        v[4 * j] = __uint_as_float(rr[4 * j]) + bb.x; v[4 * j + 1] = __uint_as_float(rr[4 * j + 1]) + bb.y;
                        v[4 * j + 2] = __uint_as_float(rr[4 * j + 2]) + bb.z; v[4 * j + 3] = __uint_as_float(rr[4 * j + 3]) + bb.w;
                    }
                    uint32_t hw[8], lw[8];
                    act_pack16<ACT_ELU, true, false, false>(v, hw, lw);
                    uint8_t *row = a2s + (c >> 6) * A_BYTES + r * 128;
                    const int pc = (c & 63) >> 3;
                    *(uint4 *)(row + ((pc ^ (r & 7)) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *(uint4 *)(row + (((pc + 1) ^ (r & 7)) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                    if (more) {
                        tmem_ld_wait16(rn);
#pragma unroll
                        for (int j = 0; j < 16; j++) rr[j] = rn[j];
                    }
                }
            }
            fence_proxy_async();          // generic-proxy writes of `mid` -> visible to the tensor core's async-proxy reads
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a2_full);
            // ---- epilogue 2: D2 -> +b2 + u -> ELU -> fp16 -> HBM
            // The residual rows do not depend on the contraction: the first two chunks' loads are issued BEFORE the wait for D2 (their
            // latency hides behind the MMA round trip), later ones one chunk ahead. Chunks alternate between two register sets.
            const bool m_ok = m < p.a_rows;
            const size_t grow = (size_t)batch * p.a_rows + m;
            {
                constexpr int NU = (C2 / 2) / 16;   // chunks per warp: 2 / 4 / 8
                const uint16_t *up = p.u + grow * C2 + c2_lo;
                uint16_t *op = p.out ? p.out + grow * C2 + c2_lo : nullptr;
                uint32_t u0[8], u1[8];
                if (m_ok) { ld_global_v8u(up, u0); ld_global_v8u(up + 16, u1); }
                if (fuse_out && lt > 0) epilogue3(tile - (int)gridDim.x, lt - 1);
                mbar_wait(d2_full, lt & 1);
                tc_fence_after();
                uint32_t ra[16], rb[16];
                tmem_ld16(trow + D2_COL + c2_lo, ra);
                auto chunk = [&](int i, uint32_t(&acc)[16], uint32_t(&hr)[8]) {
                    const int c = c2_lo + 16 * i;
                    if (m_ok) {
                        float v[16];
                        const float4 *bp = (const float4 *)(b2s + c);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float4 bb = bp[j];
                            v[4 * j] = __uint_as_float(acc[4 * j]) + bb.x; v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + bb.y;
                            v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + bb.z; v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + bb.w;
                        }
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float2 f = __half22float2(*(const __half2 *)&hr[j]);
                            v[2 * j] += f.x; v[2 * j + 1] += f.y;
                        }
                        if (i + 2 < NU) ld_global_v8u(up + 16 * (i + 2), hr);   // this register set's next chunk
                        uint32_t hw[8], lw[8];
                        act_pack16<ACT_ELU, true, false, false>(v, hw, lw);
                        if (op) st_global_v8(op + 16 * i, hw);
                        if (fuse_out) {   // final conv on the fp16-rounded block output (what the unfused path reads back): the tile goes
                                          // into its own operand buffer as the A operand of contraction 3
                            uint8_t *row = a3s + r * 128;
                            const int pc = c >> 3;
                            *(uint4 *)(row + ((pc ^ (r & 7)) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                            *(uint4 *)(row + (((pc + 1) ^ (r & 7)) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                        }
                    } else if (fuse_out) {   // rows past the end of the utterance: finite operand rows
                        uint8_t *row = a3s + r * 128;
                        const int pc = c >> 3;
                        *(uint4 *)(row + ((pc ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                        *(uint4 *)(row + (((pc + 1) ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                    }
                };
#pragma unroll 1
                for (int i = 0; i < NU; i += 2) {
                    tmem_ld_wait16(ra);
                    tmem_ld16(trow + D2_COL + c2_lo + 16 * (i + 1), rb);
                    chunk(i, ra, u0);
                    tmem_ld_wait16(rb);
                    if (i + 2 < NU) tmem_ld16(trow + D2_COL + c2_lo + 16 * (i + 2), ra);
                    chunk(i + 1, rb, u1);
                }
                if (fuse_out) {
                    fence_proxy_async();          // generic-proxy writes of the output tile -> visible to the tensor core
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(d2_empty); mbar_arrive(a3_full); }   // D2 is drained; the tile is staged
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && !fuse_out) mbar_arrive(d2_empty);
        }
        if (fuse_out && lt > 0) epilogue3(blockIdx.x + (lt - 1) * (int)gridDim.x, lt - 1);   // the last tile's samples
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// Contributions of the two rows before each tile to its first two samples (see ResParams::dw). One thread per tile.
__global__ void k_dec_edge_fix(const float *edge, float *pcm, int tiles_per_batch, int tiles_total, int a_rows) {
    pdl_wait(); pdl_trigger();
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= tiles_total) return;
    const int batch = tile / tiles_per_batch, j = tile - batch * tiles_per_batch;
    if (j == 0) return;                          // causal zero padding before the first row of an utterance
    const float *e = edge + (size_t)(tile - 1) * 4;
    const size_t o = (size_t)batch * a_rows + (size_t)j * GEMM_BM;
    pcm[o] += e[0] + e[2];                       // w1.e[t-1] + w0.e[t-2]
    if (j * GEMM_BM + 1 < a_rows) pcm[o + 1] += e[1];   // w0.e[t-1]
}

}  // namespace ptts
