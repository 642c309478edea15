// attn.cuh - decode attention phase of the persistent FlowLM step kernel (included by mega.cuh).
//
// Semantics: ptts_flowlm.c:465-525 - RoPE(q), RoPE(k) at the new position, k / v appended to the cache, then
// softmax(q.K^T / 8).V per head over [shared voice prefix | the utterance's own rows].
//
// A CTA works on one head and a span of utterances (gridDim / 16 CTAs per head), in chunks of up to 32 units
// (unit = one utterance of that head). The two halves of the key set are treated differently:
//   * voice prefix (the SAME K / V rows for every utterance): a dense contraction over the batch.
//     S = Q[32 x 64] . Kp^T[64 x <=128] and O = P[32 x 128] . Vp[128 x 64] run on the tensor cores
//     (mma.sync.m16n8k16, fp16 operands, fp32 accumulate): q and the probabilities are carried as fp16 hi + lo pairs, the
//     prefix K / V as the fp16 copy made once per voice. The work is cut into (16 units) x (16 keys) slices, one per warp;
//     each slice leaves an un-normalised (max, sum, O[64]) partial per unit in shared memory.
//   * own rows (fp16, ~tok_len + 1 + frames rows per utterance, read exactly once): one warp per unit, 8 lanes x 16 B per
//     128-byte row, 4 rows per warp instruction; a warp's two units form ONE stream of 32-row blocks (K blocks, V blocks,
//     next unit's K blocks ...) with two blocks in flight in registers across every stage boundary; fp32 math.
//   The partial softmaxes (8 prefix slices + own rows) are merged by log-sum-exp (exact in exact arithmetic) and the output is
//   written as bf16 hi + lo (the A operand of the out-projection contraction).
// Prefix keys beyond the 128-row tile (voice prompts longer than 128 frames) are read by the unit's own warp from the fp32
// prefix cache in L2.
#pragma once

namespace ptts {

constexpr int A6_TILE = 128;                 // prefix keys on the tensor-core path
constexpr int A6_CHUNK = 32;                 // units per chunk = two m16 row tiles
constexpr int A6_WARPS = MG_THREADS / 32;
// shared-memory map (bytes from the 1024-aligned ring base)
constexpr int A6_SLICES = 8;                 // 16-key slices of the prefix tile, one warp each per row tile
constexpr int A6_KP = 0, A6_VP = 16384, A6_QH = 32768, A6_QL = 36864, A6_QF = 40960, A6_PO = 49152,
              A6_PML = A6_PO + A6_CHUNK * A6_SLICES * FL_HD * 4, A6_SC = A6_PML + A6_CHUNK * A6_SLICES * 2 * 4;
__host__ __device__ inline size_t attn6_smem_bytes(int pcap) { return (size_t)A6_SC + (size_t)A6_WARPS * pcap * sizeof(float); }

__device__ __forceinline__ void ld_global_v8_pred(const float *ptr, bool ok, float (&v)[8]) {
    if (ok) ld_global_v8(ptr, v);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = 0.0f;
    }
}
__device__ __forceinline__ float dot8_reduce(const float (&q)[8], const float (&k)[8]) {
    float a = q[0] * k[0] + q[1] * k[1] + q[2] * k[2] + q[3] * k[3] + q[4] * k[4] + q[5] * k[5] + q[6] * k[6] + q[7] * k[7];
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    return a;
}
__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
    const float2 a = __half22float2(*(const __half2 *)&u.x), b = __half22float2(*(const __half2 *)&u.y),
                 c = __half22float2(*(const __half2 *)&u.z), d = __half22float2(*(const __half2 *)&u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// 32 own rows (fp16, 128 B each) starting at `base`: lane (k4 = lane >> 3, l8 = lane & 7) fetches bytes [16 l8, 16 l8 + 16)
// of rows 4 it + k4; rows past n come back as zeros
__device__ __forceinline__ void load_own32(const __half *base, int n, int k4, int l8, uint4 (&raw)[8]) {
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int key = it * 4 + k4;
        raw[it] = key < n ? *(const uint4 *)(base + (size_t)key * FL_HD + l8 * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
}

// ---- prefix keys beyond the shared-memory tile: fp32 rows straight from the prefix cache (L2)
__device__ __noinline__ void a6_far_qk(const float *kb, int n, const float *sqd, float *sc, int k4, int l8) {
    float q8[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q8[i] = sqd[l8 * 8 + i];
    for (int r0 = 0; r0 < n; r0 += 32) {
        float kr[8][8];
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int key = r0 + it * 4 + k4;
            ld_global_v8_pred(kb + (size_t)key * FL_HD + l8 * 8, key < n, kr[it]);
        }
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int key = r0 + it * 4 + k4;
            const float a = dot8_reduce(q8, kr[it]);
            if (l8 == 0 && key < n) sc[key] = a;
        }
    }
}
__device__ __noinline__ void a6_far_pv(const float *vb, int n, const float *sc, int k4, int l8, float (&acc)[8]) {
    for (int r0 = 0; r0 < n; r0 += 32) {
        float vr[8][8];
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int key = r0 + it * 4 + k4;
            ld_global_v8_pred(vb + (size_t)key * FL_HD + l8 * 8, key < n, vr[it]);
        }
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int key = r0 + it * 4 + k4;
            const float pw = key < n ? sc[key] : 0.0f;
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] += pw * vr[it][i];
        }
    }
}

// ---- fragment loads. Tiles are [rows][64 halves] with 128-byte rows; the 16-byte chunk c of row r sits at chunk c ^ (r & 7)
__device__ __forceinline__ uint32_t a6_swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// fp32 pair -> fp16 hi pair and fp16 lo pair (lo = value - hi, exactly representable difference rounded once)
__device__ __forceinline__ void split_hl(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *(const uint32_t *)&h;
    lo = *(const uint32_t *)&l;
}

// ---- new position of unit (b, h): fold the split-K partial sums of q / k / v, RoPE, append k / v to the own cache, q -> smem
// (fp32 for the own-row part, fp16 hi + lo in the fragment tile for the prefix part). Lane = dimension pair.
__device__ __forceinline__ void a6_new_row(const AttnDecodeParams &p, int b, int h, int u, int step, int cond_len, uint8_t *sm, int lane) {
    const float *row = p.qkv + (size_t)b * 3 * FL_D + h * FL_HD + 2 * lane;
    float2 rq[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int z = 0; z < 4; z++)
            rq[c][z] = (z == 0 || z < p.qkv_nsplit) ? *(const float2 *)(row + c * FL_D + z * p.qkv_stride) : make_float2(0.0f, 0.0f);
    const int s = step + p.own_base[b], pos = cond_len + s;
    const float cs = p.rope_cos[pos * 32 + lane], sn = p.rope_sin[pos * 32 + lane];
    float2 f[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        f[c] = make_float2((rq[c][0].x + rq[c][1].x) + (rq[c][2].x + rq[c][3].x), (rq[c][0].y + rq[c][1].y) + (rq[c][2].y + rq[c][3].y));
        for (int z = 4; z < p.qkv_nsplit; z++) {   // more than four partial sums (not produced by the step kernel's QKV split)
            const float2 e = *(const float2 *)(row + c * FL_D + z * p.qkv_stride);
            f[c].x += e.x; f[c].y += e.y;
        }
    }
    __half *okb = p.ok + ((size_t)h * p.cap_b + b) * p.max_s * FL_HD, *ovb = p.ov + ((size_t)h * p.cap_b + b) * p.max_s * FL_HD;
    *(__half2 *)(okb + (size_t)s * FL_HD + 2 * lane) = __floats2half2_rn(f[1].x * cs - f[1].y * sn, f[1].x * sn + f[1].y * cs);
    *(__half2 *)(ovb + (size_t)s * FL_HD + 2 * lane) = __floats2half2_rn(f[2].x, f[2].y);
    const float qx = (f[0].x * cs - f[0].y * sn) * 0.125f, qy = (f[0].x * sn + f[0].y * cs) * 0.125f;   // 1 / sqrt(64) folded into q
    *(float2 *)((float *)(sm + A6_QF) + u * FL_HD + 2 * lane) = make_float2(qx, qy);
    uint32_t hi, lo;
    split_hl(qx, qy, hi, lo);
    const uint32_t off = a6_swz(u, lane >> 2) + (lane & 3) * 4;
    *(uint32_t *)(sm + A6_QH + off) = hi;
    *(uint32_t *)(sm + A6_QL + off) = lo;
}

// ---- prefix part, sliced: warp w takes row tile mt = w & 1 (16 units) and the 16 keys [16 ke, 16 ke + 16), ke = w >> 1:
// S = Q.Kp^T over those keys, a local softmax, O = P.Vp; the un-normalised partial (max, sum, O[64]) per unit and slice goes to
// shared memory and is merged with the own-row part by log-sum-exp. Rows of absent units hold q = 0 (finite results, unused);
// a slice past the end of the prefix writes (-inf, 0, 0).
__device__ __forceinline__ void a6_prefix_slice(uint8_t *sm, int mt, int ke, int n_tile, int lane) {
    const uint32_t kp = smem_u32(sm + A6_KP), vp = smem_u32(sm + A6_VP), qh = smem_u32(sm + A6_QH), ql = smem_u32(sm + A6_QL);
    const int g = lane >> 2, j = lane & 3, key0 = ke * 16;
    float *po = (float *)(sm + A6_PO), *pml = (float *)(sm + A6_PML);
    const int ua = mt * 16 + g, ub = ua + 8;
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    if (key0 < n_tile) {
        float s[2][4];
        {
            const int qrow = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
                uint32_t kb[2][4];
                const int row = key0 + nt * 8 + (lane & 7);
                ldsm_x4(kp + a6_swz(row, lane >> 3), kb[0]);        // dims 0..31: b0 / b1 of k-steps 0, 1
                ldsm_x4(kp + a6_swz(row, 4 + (lane >> 3)), kb[1]);  // dims 32..63: k-steps 2, 3
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    uint32_t ah[4], al[4];
                    const uint32_t off = a6_swz(qrow, ks * 2 + (lane >> 4));
                    ldsm_x4(qh + off, ah);
                    ldsm_x4(ql + off, al);
                    mma_f16_16816(s[nt], ah, kb[ks >> 1][(ks & 1) * 2], kb[ks >> 1][(ks & 1) * 2 + 1]);
                    mma_f16_16816(s[nt], al, kb[ks >> 1][(ks & 1) * 2], kb[ks >> 1][(ks & 1) * 2 + 1]);
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            const int key = key0 + nt * 8 + 2 * j;
            if (key >= n_tile) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
            if (key + 1 >= n_tile) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
        }
        m0 = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
        m1 = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        // key key0 is valid, so both maxima are finite
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            s[nt][0] = expf(s[nt][0] - m0); s[nt][1] = expf(s[nt][1] - m0);
            s[nt][2] = expf(s[nt][2] - m1); s[nt][3] = expf(s[nt][3] - m1);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        uint32_t ph[4], pl[4];   // the two score tiles are the A fragment of P for this 16-key step
        split_hl(s[0][0], s[0][1], ph[0], pl[0]);
        split_hl(s[0][2], s[0][3], ph[1], pl[1]);
        split_hl(s[1][0], s[1][1], ph[2], pl[2]);
        split_hl(s[1][2], s[1][3], ph[3], pl[3]);
        const int krow = key0 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
        for (int dp = 0; dp < 4; dp++) {   // dimension tiles 2 dp, 2 dp + 1
            uint32_t vb[4];
            ldsm_x4_t(vp + a6_swz(krow, dp * 2 + (lane >> 4)), vb);
            mma_f16_16816(o[2 * dp], ph, vb[0], vb[1]);
            mma_f16_16816(o[2 * dp], pl, vb[0], vb[1]);
            mma_f16_16816(o[2 * dp + 1], ph, vb[2], vb[3]);
            mma_f16_16816(o[2 * dp + 1], pl, vb[2], vb[3]);
        }
    }
    if (j == 0) {
        pml[(ua * A6_SLICES + ke) * 2] = m0; pml[(ua * A6_SLICES + ke) * 2 + 1] = l0;
        pml[(ub * A6_SLICES + ke) * 2] = m1; pml[(ub * A6_SLICES + ke) * 2 + 1] = l1;
    }
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        *(float2 *)(po + (size_t)(ua * A6_SLICES + ke) * FL_HD + nt * 8 + 2 * j) = make_float2(o[nt][0], o[nt][1]);
        *(float2 *)(po + (size_t)(ub * A6_SLICES + ke) * FL_HD + nt * 8 + 2 * j) = make_float2(o[nt][2], o[nt][3]);
    }
}

// ---- own rows of this warp's units (chunk units w and w + 16), as ONE stream of 32-row blocks
//   unit 0: K blocks, V blocks; unit 1: K blocks, V blocks
// with two blocks in flight in registers at all times: the V rows of a unit are on their way while its softmax runs, the next
// unit's K rows while this unit's output is merged and written - HBM never waits for a latency-bound stage of the same warp.
// (unit descriptors are plain scalars selected with ?: - an indexed array of structs would live in local memory)
#define A6_SEL(i, a0, a1) ((i) < total0 ? (a0) : (a1))
#define A6_ISSUE(i, raw)                                                                                              \
    {                                                                                                                 \
        const int li_ = (i) < total0 ? (i) : (i) - total0, nb_ = A6_SEL(i, nb0, nb1), no_ = A6_SEL(i, n_own0, n_own1); \
        const bool isv_ = li_ >= nb_;                                                                                 \
        const int r0_ = (isv_ ? li_ - nb_ : li_) * 32;                                                                \
        load_own32((isv_ ? A6_SEL(i, vp0, vp1) : A6_SEL(i, kp0, kp1)) + (size_t)r0_ * FL_HD, no_ - r0_, k4, l8, raw); \
    }

__device__ __noinline__ void a6_own_stream(const AttnDecodeParams &p, int h, int b0, int nu, int u_first, int step, int cond_len, int n_tile,
                                           uint8_t *sm, float *sc, int lane, unsigned long long *tm) {
    const int k4 = lane >> 3, l8 = lane & 7, n_far = cond_len - n_tile;
    long long c0 = tm ? clock64() : 0;
#define A6_STAMP(i) if (tm) { const long long c1 = clock64(); if (lane == 0) tm[i] += (unsigned long long)(c1 - c0); c0 = c1; }
    const int ua0 = u_first, ua1 = u_first + A6_WARPS;
    const int bb0 = ua0 < nu ? b0 + ua0 : b0, bb1 = ua1 < nu ? b0 + ua1 : b0;
    const int n_own0 = ua0 < nu ? step + p.own_base[bb0] + 1 : 0, n_own1 = ua1 < nu ? step + p.own_base[bb1] + 1 : 0;
    const int nb0 = (n_own0 + 31) >> 5, nb1 = (n_own1 + 31) >> 5;
    const __half *kp0 = p.ok + ((size_t)h * p.cap_b + bb0) * p.max_s * FL_HD, *vp0 = p.ov + ((size_t)h * p.cap_b + bb0) * p.max_s * FL_HD;
    const __half *kp1 = p.ok + ((size_t)h * p.cap_b + bb1) * p.max_s * FL_HD, *vp1 = p.ov + ((size_t)h * p.cap_b + bb1) * p.max_s * FL_HD;
    const int total0 = 2 * nb0, total = total0 + 2 * nb1;
    if (total == 0) return;
    uint4 ra[8], rb[8];
    A6_ISSUE(0, ra)
    if (total > 1) A6_ISSUE(1, rb)
    float q8[8], acc[8], mo = 0.0f, lo = 0.0f;
#pragma unroll 1
    for (int i = 0; i < total; i++) {
        const int li = i < total0 ? i : i - total0, x_nb = A6_SEL(i, nb0, nb1), x_n_own = A6_SEL(i, n_own0, n_own1);
        const int x_u = A6_SEL(i, ua0, ua1), x_b = A6_SEL(i, bb0, bb1);
        const bool isv = li >= x_nb;
        const int r0 = (isv ? li - x_nb : li) * 32;
        const float *qf = (const float *)(sm + A6_QF) + x_u * FL_HD;
        if (li == 0) {
#pragma unroll
            for (int d = 0; d < 8; d++) { q8[d] = qf[l8 * 8 + d]; acc[d] = 0.0f; }
        }
        // ---- consume block i from the buffer it was loaded into (even blocks: ra, odd blocks: rb), then refill that buffer
        if (!isv) {
            if ((i & 1) == 0) {
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    float kf[8];
                    unpack8(ra[it], kf);
                    const float a = dot8_reduce(q8, kf);
                    const int key = r0 + it * 4 + k4;
                    if (l8 == 0 && key < x_n_own) sc[key] = a;
                }
            } else {
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    float kf[8];
                    unpack8(rb[it], kf);
                    const float a = dot8_reduce(q8, kf);
                    const int key = r0 + it * 4 + k4;
                    if (l8 == 0 && key < x_n_own) sc[key] = a;
                }
            }
        } else {
            if ((i & 1) == 0) {
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    const int key = r0 + it * 4 + k4;
                    const float pw = key < x_n_own ? sc[key] : 0.0f;
                    float vf[8];
                    unpack8(ra[it], vf);
#pragma unroll
                    for (int d = 0; d < 8; d++) acc[d] += pw * vf[d];
                }
            } else {
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    const int key = r0 + it * 4 + k4;
                    const float pw = key < x_n_own ? sc[key] : 0.0f;
                    float vf[8];
                    unpack8(rb[it], vf);
#pragma unroll
                    for (int d = 0; d < 8; d++) acc[d] += pw * vf[d];
                }
            }
        }
        if (i + 2 < total) {
            if ((i & 1) == 0) A6_ISSUE(i + 2, ra)
            else A6_ISSUE(i + 2, rb)
        }
        if (!isv && li == x_nb - 1) {
            // ---- all scores of the unit are in: far prefix rows, then softmax over [own rows | far prefix rows]
            A6_STAMP(1)
            if (n_far > 0) a6_far_qk(p.pk + ((size_t)h * p.pstride + n_tile) * FL_HD, n_far, qf, sc + x_n_own, k4, l8);
            __syncwarp();
            const int np = x_n_own + n_far;
            mo = -INFINITY;
            for (int t = lane; t < np; t += 32) mo = fmaxf(mo, sc[t]);
            mo = warp_max(mo);
            lo = 0.0f;
            for (int t = lane; t < np; t += 32) { const float e = expf(sc[t] - mo); sc[t] = e; lo += e; }
            lo = warp_sum(lo);
            __syncwarp();
            A6_STAMP(2)
        }
        if (isv && li == 2 * x_nb - 1) {
            // ---- unit complete: reduce over the 4 row groups, merge with the prefix partials (log-sum-exp), write bf16 hi + lo
            A6_STAMP(3)
            if (n_far > 0) a6_far_pv(p.pv + ((size_t)h * p.pstride + n_tile) * FL_HD, n_far, sc + x_n_own, k4, l8, acc);
#pragma unroll
            for (int d = 0; d < 8; d++) {
                acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], 8);
                acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], 16);
            }
            if (k4 == 0) {   // lanes 0..7 hold the 64 outputs, 8 consecutive dimensions each
                float m = mo;
                const float *pml = (const float *)(sm + A6_PML) + (size_t)x_u * A6_SLICES * 2;
                const float *po = (const float *)(sm + A6_PO) + (size_t)x_u * A6_SLICES * FL_HD + l8 * 8;
                if (n_tile > 0) {
#pragma unroll
                    for (int e = 0; e < A6_SLICES; e++) m = fmaxf(m, pml[2 * e]);
                }
                const float wb = expf(mo - m);
                float den = lo * wb, out[8];
#pragma unroll
                for (int d = 0; d < 8; d++) out[d] = acc[d] * wb;
                if (n_tile > 0) {
#pragma unroll
                    for (int e = 0; e < A6_SLICES; e++) {
                        const float we = expf(pml[2 * e] - m);   // exp(-inf) = 0 for slices past the end of the prefix
                        den += pml[2 * e + 1] * we;
                        const float4 pa = *(const float4 *)(po + e * FL_HD), pb = *(const float4 *)(po + e * FL_HD + 4);
                        out[0] += pa.x * we; out[1] += pa.y * we; out[2] += pa.z * we; out[3] += pa.w * we;
                        out[4] += pb.x * we; out[5] += pb.y * we; out[6] += pb.z * we; out[7] += pb.w * we;
                    }
                }
                const float inv = 1.0f / den;
#pragma unroll
                for (int d = 0; d < 8; d++) out[d] *= inv;
                const size_t oi = (size_t)x_b * FL_D + h * FL_HD + l8 * 8;
                if (p.out_f32) {
                    *(float4 *)(p.out_f32 + oi) = make_float4(out[0], out[1], out[2], out[3]);
                    *(float4 *)(p.out_f32 + oi + 4) = make_float4(out[4], out[5], out[6], out[7]);
                }
                if (p.out_h) {
                    uint4 hv, lv;
                    uint32_t *hp = (uint32_t *)&hv, *lw = (uint32_t *)&lv;
#pragma unroll
                    for (int d = 0; d < 4; d++) {
                        const uint16_t h0 = f2h16(out[2 * d], true), h1 = f2h16(out[2 * d + 1], true);
                        hp[d] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                        const uint16_t l0 = f2h16(out[2 * d] - h162f(h0, true), true), l1 = f2h16(out[2 * d + 1] - h162f(h1, true), true);
                        lw[d] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                    }
                    *(uint4 *)(p.out_h + oi) = hv;
                    if (p.out_l) *(uint4 *)(p.out_l + oi) = lv;
                }
            }
            __syncwarp();   // score scratch reusable by the next unit
            A6_STAMP(4)
            if (tm && lane == 0) tm[5] += 1;
        }
    }
#undef A6_STAMP
#undef A6_ISSUE
#undef A6_SEL
}

// One attention phase: units [u0, u1) of this CTA in head-major order (unit = h * B + b).
__device__ __forceinline__ void attn6_phase(const AttnDecodeParams &p, int B, int pcap, int step, uint8_t *sm, int tid, unsigned long long *tmw = nullptr) {
    const int w = tid >> 5, lane = tid & 31;
    float *sc = (float *)(sm + A6_SC) + (size_t)w * pcap;
    const int cond_len = p.ctrl_cond_len ? *p.ctrl_cond_len : 0;
    const int n_tile = min(cond_len, A6_TILE);
    // with at least one CTA per head, every CTA stays inside one head (one prefix tile per phase): gridDim / 16 CTAs share
    // a head's utterances
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
    for (int ua = u0; ua < u1;) {
        const int h = ua / B, ub = min(u1, (h + 1) * B);   // sub-range of one head
        __syncthreads();                                    // previous head's tiles no longer in use
        if (n_tile > 0) {   // fp16 prefix tiles of head h, rows past the prefix zero-filled (P = 0 times garbage must stay 0)
            const __half *pkh = p.pk16 + (size_t)h * p.pstride * FL_HD, *pvh = p.pv16 + (size_t)h * p.pstride * FL_HD;
            for (int idx = tid; idx < A6_TILE * 8; idx += MG_THREADS) {
                const int row = idx >> 3, c = idx & 7;
                uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
                if (row < n_tile) { kv = *(const uint4 *)(pkh + (size_t)row * FL_HD + c * 8); vv = *(const uint4 *)(pvh + (size_t)row * FL_HD + c * 8); }
                *(uint4 *)(sm + A6_KP + a6_swz(row, c)) = kv;
                *(uint4 *)(sm + A6_VP + a6_swz(row, c)) = vv;
            }
        }
        for (int c0 = ua; c0 < ub; c0 += A6_CHUNK) {
            const int nu = min(A6_CHUNK, ub - c0), b0 = c0 - h * B;
            if (c0 != ua) __syncthreads();                  // previous chunk's q rows and prefix partials no longer in use
            long long t0 = (tmw && w == 2) ? clock64() : 0;
            for (int u = w; u < A6_CHUNK; u += A6_WARPS) {
                if (u < nu) a6_new_row(p, b0 + u, h, u, step, cond_len, sm, lane);
                else {   // absent unit: a zero query keeps the row tile's arithmetic finite
                    const uint32_t off = a6_swz(u, lane >> 2) + (lane & 3) * 4;
                    *(uint32_t *)(sm + A6_QH + off) = 0u;
                    *(uint32_t *)(sm + A6_QL + off) = 0u;
                }
            }
            __syncthreads();
            if (tmw && w == 2 && lane == 0) { const long long t1 = clock64(); tmw[0] += (unsigned long long)(t1 - t0); t0 = t1; }
            if (n_tile > 0) {
                if ((w & 1) * 16 < nu) a6_prefix_slice(sm, w & 1, w >> 1, n_tile, lane);
                __syncthreads();                            // all partials of the chunk are in shared memory
                if (tmw && w == 2 && lane == 0) { tmw[6] += (unsigned long long)(clock64() - t0); tmw[7] += 1; }
            }
            a6_own_stream(p, h, b0, nu, w, step, cond_len, n_tile, sm, sc, lane, (tmw && w == 2) ? tmw : nullptr);
        }
        ua = ub;
    }
}

}  // namespace ptts
