// attn_tc.cuh - decode attention of the persistent step kernel on warp-level tensor-core MMAs
// (mma.sync.m16n8k16, fp16 operands, fp32 accumulate; ptts_flowlm.c:465-525: max-subtracted softmax in
// fp32, 1/sqrt(64) folded into q). One warp owns a PAIR of (utterance, head) units from start to end:
//
//   scores   S[key][col]  = sum_d K[key][d] * Q[d][col]      A = 16 keys x 16 dims, B = 8 query columns
//            columns: 0 = q(unit 0) hi, 1 = q(unit 0) lo, 2 = q(unit 1) hi, 3 = q(unit 1) lo
//            (q = hi + lo in fp16: ~22 significand bits; K rows are fp16 like the own cache)
//   outputs  O[row][dim]  = sum_key P[row][key] * V[key][d]  A = P rows, B = 16 keys x 8 dims
//            rows: 0 = p(unit 0) hi, 1 = p(unit 0) lo, 2 = p(unit 1) hi, 3 = p(unit 1) lo
//
// Every K / V row block is a [16 keys][128 bytes] tile in shared memory with the 16-byte chunks of row r
// XOR-ed by r & 7, read with ldmatrix.x4 (K: A fragments) / ldmatrix.x4.trans (V: B fragments) - the
// lane addresses are the same for both. The shared voice prefix (fp16 copy made once per voice) is one
// such tile set per CTA; a unit's own rows stream through a per-warp ring of three blocks filled with
// cp.async (the rows in flight hold no registers), in the order K(unit 0), K(unit 1), prefix rows beyond
// the tile, then the same for V. The row appended this step comes from shared memory, so the ring loads
// are issued before it is computed. Score buffers keep every block 16-aligned: prefix scores at
// [0, cpad), unit rows at [cpad, ...), the slack filled with -inf (exp -> 0), so no P masking is needed.
#pragma once

namespace ptts {

constexpr int A5_SLOTS = 3, A5_SLOT_BYTES = 16 * 128;
__host__ __device__ inline size_t attn5_warp_bytes(int pcap) {   // sc[2][pcap], sq[2][64], block ring, new rows [2][k|v][64]
    return (size_t)(2 * pcap + 2 * FL_HD) * sizeof(float) + A5_SLOTS * A5_SLOT_BYTES + 4 * FL_HD * 2;
}
__host__ __device__ inline size_t attn5_smem_bytes(int pcap) { return (size_t)2 * A4_TILE * 128 + A4_WARPS * attn5_warp_bytes(pcap) + 16; }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void a5_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void a5_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// (x0, x1) -> packed fp16 pair: the rounded values (part 0) or the rounding remainders (part 1)
__device__ __forceinline__ uint32_t a5_split_pack(float x0, float x1, int part) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
    return part ? *(const uint32_t *)&l : *(const uint32_t *)&h;
}
__device__ __forceinline__ uint4 a5_pack8(const float4 &a, const float4 &b) {
    return make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(b.x, b.y), pack_h2(b.z, b.w));
}

// fp32 -> fp16 copy of the voice prefix K / V (once per voice, ptts_b200.cu set_voice)
__global__ void __launch_bounds__(256) k_f32_to_f16(const float *src, __half *dst, size_t n8) {
    pdl_wait(); pdl_trigger();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = *(const float4 *)(src + i * 8), b = *(const float4 *)(src + i * 8 + 4);
        *(uint4 *)(dst + i * 8) = a5_pack8(a, b);
    }
}

// Scores of one 16-key block: tile = shared-memory byte address of the block + the lane's row offset,
// x = the lane's chunk XOR (row & 7), mhi = lane >> 4. Lanes t == 0 / 1 end up with unit 0 / 1:
// c[0] + c[1] = key g, c[2] + c[3] = key g + 8.
__device__ __forceinline__ void a5_qk_block(uint32_t tile, int mhi, int x, const uint32_t (&bq)[4][2], float (&c)[4]) {
    c[0] = c[1] = c[2] = c[3] = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t a[4];
        ldsm_x4(tile + (uint32_t)(((2 * j + mhi) ^ x) << 4), a);
        mma_f16_16816(c, a, bq[j][0], bq[j][1]);
    }
}
// P.V of one 16-key block: pa0 / pa2 = the lane's P fragment words (keys 2t, 2t+1 / 2t+8, 2t+9 of its row)
__device__ __forceinline__ void a5_pv_block(uint32_t tile, int mhi, int x, uint32_t pa0, uint32_t pa2, float (&o)[8][4]) {
    const uint32_t a[4] = {pa0, 0u, pa2, 0u};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t r[4];
        ldsm_x4_trans(tile + (uint32_t)(((2 * i + mhi) ^ x) << 4), r);
        mma_f16_16816(o[2 * i], a, r[0], r[1]);
        mma_f16_16816(o[2 * i + 1], a, r[2], r[3]);
    }
}
// The lane's P fragment for the block starting at score index i0 (a multiple of 16): rows g < 4 are
// (unit g >> 1, part g & 1); `on` = this lane's row takes part in the block.
__device__ __forceinline__ void a5_p_frag(const float *s, int i0, int t, int part, bool on, uint32_t &pa0, uint32_t &pa2) {
    float2 p01 = make_float2(0.0f, 0.0f), p23 = p01;
    if (on) { p01 = *(const float2 *)(s + i0 + 2 * t); p23 = *(const float2 *)(s + i0 + 2 * t + 8); }
    pa0 = a5_split_pack(p01.x, p01.y, part);
    pa2 = a5_split_pack(p23.x, p23.y, part);
}

struct A5Pair {
    int n[2], s[2];                     // own rows incl. the new one / index of the new row
    const __half *ok[2], *ov[2];        // own caches of the two units
    // issue cursor
    const __half *cp;                   // the lane's source (row k4 of the next block, chunk l8)
    int rows, seg;                      // rows left in the segment, segment 0..5 = K0 K1 Kfar V0 V1 Vfar
    uint32_t islot, cslot;              // ring slots of the next block issued / consumed
};

// One attention phase: units [u0, u1) of this CTA in head-major order (unit = h * B + b); warp w takes the pairs
// w, w + 16, ... of the CTA's range (one pair per warp for B <= 288 on 148 SMs).
__device__ __forceinline__ void attn5_phase(const AttnDecodeParams &p, int B, int pcap, int step, uint8_t *sm, int tid, unsigned long long *tmw, int flags) {
    uint8_t *kt = sm, *vt = kt + A4_TILE * 128, *scb = vt + A4_TILE * 128;
    const size_t wbytes = attn5_warp_bytes(pcap);
    const int w = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3, k4 = lane >> 3, l8 = lane & 7;
    float *sc0 = (float *)(scb + w * wbytes), *sc1 = sc0 + pcap, *sq = sc0 + 2 * pcap;
    uint8_t *ringw = (uint8_t *)(sq + 2 * FL_HD);
    __half *nrw = (__half *)(ringw + A5_SLOTS * A5_SLOT_BYTES);
    const uint32_t ring_addr = smem_u32(ringw);
    // ldmatrix lane address inside a block: row (k4 & 1) * 8 + l8, chunk (2 * i + (k4 >> 1)) ^ l8
    const uint32_t lrow = (uint32_t)(((k4 & 1) * 8 + l8) * 128);
    const int mhi = k4 >> 1;
    // cp.async lane layout inside a block: rows it * 4 + k4, chunk l8 -> (l8 ^ (row & 7)) * 16
    const uint32_t cpo0 = (uint32_t)(k4 * 128 + ((l8 ^ k4) << 4)), cpo1 = (uint32_t)((4 + k4) * 128 + ((l8 ^ (4 + k4)) << 4));
    const int cond_len = p.cond_len_val;
    const int n_tile = min(cond_len, A4_TILE), n_far = cond_len - n_tile, cpad = (cond_len + 15) & ~15;
    unsigned long long *tm = (tmw && w == 0) ? tmw : nullptr;
    long long c0 = tm ? clock64() : 0;
#define A5_SUB(i) if (tm) { const long long c1 = clock64(); if (lane == 0) tm[i] += (unsigned long long)(c1 - c0); }
#define A5_STAMP(i) if (tm) { const long long c1 = clock64(); if (lane == 0) tm[i] += (unsigned long long)(c1 - c0); c0 = c1; }

    int u0, u1;
    const int cph = (int)gridDim.x / FL_H;
    if (cph >= 1) {
        const int h = (int)blockIdx.x / cph, j = (int)blockIdx.x % cph;
        if (h < FL_H) { u0 = h * B + (int)(((long long)j * B) / cph); u1 = h * B + (int)(((long long)(j + 1) * B) / cph); }
        else { u0 = u1 = 0; }
    } else {
        const int units = FL_H * B;
        u0 = (int)(((long long)blockIdx.x * units) / gridDim.x); u1 = (int)(((long long)(blockIdx.x + 1) * units) / gridDim.x);
    }

    A5Pair P;
    const __half *pk16 = nullptr, *pv16 = nullptr;
    auto seg_set = [&]() {   // source and row count of segment P.seg (empty segments are skipped)
        while (P.seg < 6) {
            const int sg = P.seg;
            const __half *base = sg == 0 ? P.ok[0] : sg == 1 ? P.ok[1] : sg == 2 ? pk16 + (size_t)n_tile * FL_HD
                               : sg == 3 ? P.ov[0] : sg == 4 ? P.ov[1] : pv16 + (size_t)n_tile * FL_HD;
            P.rows = (sg == 0 || sg == 3) ? P.n[0] : (sg == 1 || sg == 4) ? P.n[1] : n_far;
            if (P.rows > 0) { P.cp = base + (size_t)k4 * FL_HD + l8 * 8; return; }
            P.seg++;
        }
    };
    auto issue = [&]() {   // next block of the sequence -> ring slot P.islot; always one commit group
        if (P.seg < 6) {
            const uint32_t dst = ring_addr + P.islot * A5_SLOT_BYTES;
#pragma unroll
            for (int it = 0; it < 4; it++) {
                const bool ok = it * 4 + k4 < P.rows;   // rows past the end are zero-filled
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (it >> 1) * 1024 + ((it & 1) ? cpo1 : cpo0)),
                             "l"(ok ? P.cp + (size_t)it * 4 * FL_HD : P.cp), "r"(ok ? 16u : 0u) : "memory");
            }
            P.cp += 16 * FL_HD; P.rows -= 16;
            if (P.rows <= 0) { P.seg++; seg_set(); }
        }
        a5_commit();
        P.islot = P.islot == A5_SLOTS - 1 ? 0u : P.islot + 1;
    };
    // the new row of unit u replaces row s & 15 of the ring slot holding its last block (lanes 0..7 move 16 bytes each)
    auto patch = [&](int u, bool is_v, uint8_t *slot) {
        const int r = P.s[u] & 15;
        if (lane < 8) *(uint4 *)(slot + r * 128 + ((lane ^ (r & 7)) << 4)) = *(const uint4 *)(nrw + (u * 2 + (is_v ? 1 : 0)) * FL_HD + lane * 8);
        __syncwarp();
    };

    for (int ua = u0; ua < u1;) {
        const int h = ua / B, ub = min(u1, (h + 1) * B);   // sub-range of one head
        const int npairs = (ub - ua + 1) >> 1;
        pk16 = p.pk16 + (size_t)h * p.pstride * FL_HD; pv16 = p.pv16 + (size_t)h * p.pstride * FL_HD;
        for (int pi = w, first = 1; pi < npairs || first; pi += A4_WARPS, first = 0) {
            const bool have = pi < npairs;
            const int b0 = ua - h * B + 2 * pi, nlive = have ? min(2, ub - ua - 2 * pi) : 0;
            // ---- pair prologue. Every global round trip is requested before anything waits: the q / k / v partial sums
            // of the new rows, the row counts, the prefix tile (cp.async, all warps) and the first three ring blocks.
            float2 rq[2][3][4];
            if (have) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int b = u < nlive ? b0 + u : b0;
                    const float *row = p.qkv + (size_t)b * 3 * FL_D + h * FL_HD + 2 * lane;
#pragma unroll
                    for (int c = 0; c < 3; c++)
#pragma unroll
                        for (int z = 0; z < 4; z++)
                            rq[u][c][z] = (z == 0 || z < p.qkv_nsplit) ? *(const float2 *)(row + c * FL_D + z * p.qkv_stride) : make_float2(0.0f, 0.0f);
                    P.s[u] = step + ((p.ob_cache && b >= p.ob_first && b < p.ob_first + p.ob_count) ? p.ob_cache[b - p.ob_first] : p.own_base[b]);
                    P.ok[u] = p.ok + ((size_t)h * p.cap_b + b) * p.max_s * FL_HD;
                    P.ov[u] = p.ov + ((size_t)h * p.cap_b + b) * p.max_s * FL_HD;
                }
            }
            if (first) {
                if (ua != u0) __syncthreads();                      // previous head's tile no longer in use
                const int nt16 = (n_tile + 15) & ~15;
                const uint32_t kt_addr = smem_u32(kt), vt_addr = smem_u32(vt);
                for (int idx = tid; idx < nt16 * 8; idx += MG_THREADS) {
                    const int key = idx >> 3, c = idx & 7;
                    const bool ok = key < n_tile;                   // rows past the prefix are zero-filled
                    const uint32_t off = (uint32_t)(key * 128 + ((c ^ (key & 7)) << 4)), sz = ok ? 16u : 0u;
                    const size_t so = ok ? (size_t)idx * 8 : 0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(kt_addr + off), "l"(pk16 + so), "r"(sz) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vt_addr + off), "l"(pv16 + so), "r"(sz) : "memory");
                }
                a5_commit();   // the tile is a group of its own, ahead of the ring groups
            }
            if (have) {
                P.n[0] = P.s[0] + 1; P.n[1] = nlive == 2 ? P.s[1] + 1 : 0;
                if ((flags & 1) && lane < 4) {   // pull the whole K / V regions of both units towards L2
                    const int u = lane >> 1;
                    const __half *ptr = (lane & 1) ? P.ov[u] : P.ok[u];
                    const uint32_t bytes = (uint32_t)P.s[u] * FL_HD * 2u;
                    if (u < nlive && bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
                }
                P.seg = 0; P.islot = 0; P.cslot = 0;
                seg_set();
                issue(); issue(); issue();
                A5_SUB(6)
                // new positions (ptts_flowlm.c:470-489): RoPE(q) / 8 -> sq, RoPE(k), v -> own caches and nrw; lane = dimension pair
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    if (u < nlive) {
                        const int pos = cond_len + P.s[u];
                        const float cs = p.rope_cos[pos * 32 + lane], sn = p.rope_sin[pos * 32 + lane];
                        float2 f[3];
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            f[c] = make_float2((rq[u][c][0].x + rq[u][c][1].x) + (rq[u][c][2].x + rq[u][c][3].x),
                                               (rq[u][c][0].y + rq[u][c][1].y) + (rq[u][c][2].y + rq[u][c][3].y));
                            for (int z = 4; z < p.qkv_nsplit; z++) {   // more than four partial sums (not produced by the step kernel's QKV split)
                                const float2 e = *(const float2 *)(p.qkv + (size_t)(b0 + u) * 3 * FL_D + h * FL_HD + 2 * lane + c * FL_D + z * p.qkv_stride);
                                f[c].x += e.x; f[c].y += e.y;
                            }
                        }
                        const __half2 kh = __floats2half2_rn(f[1].x * cs - f[1].y * sn, f[1].x * sn + f[1].y * cs);
                        const __half2 vh = __floats2half2_rn(f[2].x, f[2].y);
                        *(__half2 *)(const_cast<__half *>(P.ok[u]) + (size_t)P.s[u] * FL_HD + 2 * lane) = kh;
                        *(__half2 *)(const_cast<__half *>(P.ov[u]) + (size_t)P.s[u] * FL_HD + 2 * lane) = vh;
                        *(__half2 *)(nrw + u * 2 * FL_HD + 2 * lane) = kh;
                        *(__half2 *)(nrw + (u * 2 + 1) * FL_HD + 2 * lane) = vh;
                        *(float2 *)(sq + u * FL_HD + 2 * lane) = make_float2((f[0].x * cs - f[0].y * sn) * 0.125f, (f[0].x * sn + f[0].y * cs) * 0.125f);
                    } else *(float2 *)(sq + u * FL_HD + 2 * lane) = make_float2(0.0f, 0.0f);
                }
                // slack of the score buffers: -inf (prefix tail up to cpad, unit tails up to the next multiple of 16)
                if (lane < 16) {
                    if (cond_len + lane < cpad) { sc0[cond_len + lane] = -INFINITY; sc1[cond_len + lane] = -INFINITY; }
                    sc0[cpad + P.n[0] + lane] = -INFINITY;
                    sc1[cpad + P.n[1] + lane] = -INFINITY;
                }
            }
            A5_SUB(7)
            if (first) {
                if (have) a5_wait<3>(); else a5_wait<0>();
                __syncthreads();                                    // prefix tile complete
            }
            if (!have) continue;
            __syncwarp();
            A5_STAMP(0)
            // ---- scores
            {
                uint32_t bq[4][2];   // query fragments (B operand): lane (g, t), g < 4: unit g >> 1, part g & 1, dims 16j + 2t (+8)
                if (g < 4) {
                    const float *q = sq + (g >> 1) * FL_HD + 2 * t;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float2 x0 = *(const float2 *)(q + 16 * j), x1 = *(const float2 *)(q + 16 * j + 8);
                        bq[j][0] = a5_split_pack(x0.x, x0.y, g & 1);
                        bq[j][1] = a5_split_pack(x1.x, x1.y, g & 1);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) bq[j][0] = bq[j][1] = 0u;
                }
                {   // prefix tile: both units share every K fragment
                    const uint32_t kt_addr = smem_u32(kt) + lrow;
                    float *s = t ? sc1 : sc0;
#pragma unroll 2
                    for (int kb0 = 0; kb0 < n_tile; kb0 += 16) {
                        float c[4];
                        a5_qk_block(kt_addr + kb0 * 128, mhi, l8, bq, c);
                        if (t < 2) {
                            if (kb0 + g < n_tile) s[kb0 + g] = c[0] + c[1];
                            if (kb0 + g + 8 < n_tile) s[kb0 + g + 8] = c[2] + c[3];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; u++) {   // own rows
                    float *s = (u ? sc1 : sc0) + cpad;
                    const int nb = (P.n[u] + 15) >> 4;
                    for (int blk = 0; blk < nb; blk++) {
                        a5_wait<2>();
                        __syncwarp();
                        if (blk == nb - 1) patch(u, false, ringw + P.cslot * A5_SLOT_BYTES);
                        float c[4];
                        a5_qk_block(ring_addr + P.cslot * A5_SLOT_BYTES + lrow, mhi, l8, bq, c);
                        if (t == u) {
                            if (blk * 16 + g < P.n[u]) s[blk * 16 + g] = c[0] + c[1];
                            if (blk * 16 + g + 8 < P.n[u]) s[blk * 16 + g + 8] = c[2] + c[3];
                        }
                        __syncwarp();   // every lane has read the slot
                        issue();
                        P.cslot = P.cslot == A5_SLOTS - 1 ? 0u : P.cslot + 1;
                    }
                }
                for (int r0 = 0; r0 < n_far; r0 += 16) {   // prefix rows beyond the tile (long voice prompts), shared by both units
                    a5_wait<2>();
                    __syncwarp();
                    float c[4];
                    a5_qk_block(ring_addr + P.cslot * A5_SLOT_BYTES + lrow, mhi, l8, bq, c);
                    if (t < 2) {
                        float *s = (t ? sc1 : sc0) + n_tile + r0;
                        if (r0 + g < n_far) s[g] = c[0] + c[1];
                        if (r0 + g + 8 < n_far) s[g + 8] = c[2] + c[3];
                    }
                    __syncwarp();
                    issue();
                    P.cslot = P.cslot == A5_SLOTS - 1 ? 0u : P.cslot + 1;
                }
            }
            __syncwarp();
            A5_STAMP(1)
            // ---- softmax per unit (the first V blocks are in flight); the buffers end on a multiple of 16
            float inv[2];
#pragma unroll
            for (int u = 0; u < 2; u++) inv[u] = u < nlive ? a4_softmax(u ? sc1 : sc0, cpad + ((P.n[u] + 15) & ~15), lane) : 0.0f;
            __syncwarp();
            A5_STAMP(2)
            // ---- P.V
            float o[8][4];
#pragma unroll
            for (int i = 0; i < 8; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
            const float *sp = (g & 2) ? sc1 : sc0;          // the lane's P row: unit g >> 1 (g < 4)
            const bool row_on = g < 4 && (g >> 1) < nlive;
            {
                const uint32_t vt_addr = smem_u32(vt) + lrow;
#pragma unroll 2
                for (int kb0 = 0; kb0 < n_tile; kb0 += 16) {
                    uint32_t pa0, pa2;
                    a5_p_frag(sp, kb0, t, g & 1, row_on, pa0, pa2);
                    a5_pv_block(vt_addr + kb0 * 128, mhi, l8, pa0, pa2, o);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int nb = (P.n[u] + 15) >> 4;
                const bool on = row_on && (g >> 1) == u;
                for (int blk = 0; blk < nb; blk++) {
                    a5_wait<2>();
                    __syncwarp();
                    if (blk == nb - 1) patch(u, true, ringw + P.cslot * A5_SLOT_BYTES);
                    uint32_t pa0, pa2;
                    a5_p_frag(sp, cpad + blk * 16, t, g & 1, on, pa0, pa2);
                    a5_pv_block(ring_addr + P.cslot * A5_SLOT_BYTES + lrow, mhi, l8, pa0, pa2, o);
                    __syncwarp();
                    issue();
                    P.cslot = P.cslot == A5_SLOTS - 1 ? 0u : P.cslot + 1;
                }
            }
            for (int r0 = 0; r0 < n_far; r0 += 16) {
                a5_wait<2>();
                __syncwarp();
                uint32_t pa0, pa2;
                a5_p_frag(sp, n_tile + r0, t, g & 1, row_on, pa0, pa2);
                a5_pv_block(ring_addr + P.cslot * A5_SLOT_BYTES + lrow, mhi, l8, pa0, pa2, o);
                __syncwarp();
                issue();
                P.cslot = P.cslot == A5_SLOTS - 1 ? 0u : P.cslot + 1;
            }
            A5_STAMP(3)
            // ---- rows (hi, lo) of a unit are summed, normalised and transposed through sq[] to 16-byte stores
            {
                const float iv = (g & 2) ? inv[1] : inv[0];
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    const float v0 = o[nt][0] + __shfl_down_sync(0xffffffffu, o[nt][0], 4);
                    const float v1 = o[nt][1] + __shfl_down_sync(0xffffffffu, o[nt][1], 4);
                    if (g < 4 && (g & 1) == 0) *(float2 *)(sq + (g >> 1) * FL_HD + 8 * nt + 2 * t) = make_float2(v0 * iv, v1 * iv);
                }
            }
            __syncwarp();
            if (lane < 16 && k4 < nlive) {   // lanes 0..7 / 8..15: the 64 outputs of unit 0 / 1, 8 consecutive dimensions each
                const float *src = sq + k4 * FL_HD + l8 * 8;
                const float4 a = *(const float4 *)src, b = *(const float4 *)(src + 4);
                const float acc[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                const size_t oi = (size_t)(b0 + k4) * FL_D + h * FL_HD + l8 * 8;
                if (p.out_f32) {
                    *(float4 *)(p.out_f32 + oi) = a;
                    *(float4 *)(p.out_f32 + oi + 4) = b;
                }
                if (p.out_h) {
                    uint4 hv, lv;
                    uint32_t *hp = (uint32_t *)&hv, *lp = (uint32_t *)&lv;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint16_t h0 = f2h16(acc[2 * i], true), h1 = f2h16(acc[2 * i + 1], true);
                        hp[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                        const uint16_t l0 = f2h16(acc[2 * i] - h162f(h0, true), true), l1 = f2h16(acc[2 * i + 1] - h162f(h1, true), true);
                        lp[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                    }
                    *(uint4 *)(p.out_h + oi) = hv;
                    if (p.out_l) *(uint4 *)(p.out_l + oi) = lv;
                }
            }
            if (flags & 4) {   // the same units' rows of the layer that comes next: towards L2 while the GEMM phases run (one 128-byte row per lane)
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    if (u < nlive) {
                        const size_t uo = ((size_t)h * p.cap_b + b0 + u) * p.max_s * FL_HD;
                        const int nr = P.s[u] + p.n_extra;
                        for (int r = lane; r < nr; r += 32) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.nk + uo + (size_t)r * FL_HD) : "memory");
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.nv + uo + (size_t)r * FL_HD) : "memory");
                        }
                    }
                }
            }
            a5_wait<0>();
            __syncwarp();   // scratch reusable by this warp's next pair
            A5_STAMP(4)
            if (tm && lane == 0) tm[5] += (unsigned long long)nlive;
        }
        ua = ub;
    }
#undef A5_STAMP
#undef A5_SUB
}

}  // namespace ptts
