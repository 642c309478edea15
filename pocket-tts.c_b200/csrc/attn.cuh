// attn.cuh - decode attention phase of the persistent FlowLM step kernel (included by mega.cuh).
//
// Semantics: ptts_flowlm.c:465-525 - RoPE(q), RoPE(k) at the new position, k / v appended to the cache, then
// softmax(q.K^T / 8).V per head over [shared voice prefix | the utterance's own rows].
//
// A CTA works on one head and a span of utterances (gridDim / 16 CTAs per head), in chunks of up to 32 units
// (unit = one utterance of that head). The two halves of the key set are treated differently:
//   * voice prefix (the SAME K / V rows for every utterance): a dense contraction over the batch.
//     S = Q[32 x 64] . Kp^T[64 x <=128] and O = P[32 x 128] . Vp[128 x 64] run on the tensor cores
//     (mma.sync.m16n8k16, fp16 operands, fp32 accumulate), flash-style inside ONE warp per 16 units: q and the
//     probabilities are carried as fp16 hi + lo pairs, the prefix K / V as the fp16 copy made once per voice.
//     The result per unit is an un-normalised (max, sum, O[64]) triple in shared memory.
//   * own rows (fp16, ~tok_len + 1 + frames rows per utterance, read exactly once): one warp per unit, 8 lanes x 16 B per
//     128-byte row, 4 rows per warp instruction, register double-buffered batches of 32 rows; fp32 math.
//   The two partial softmaxes are merged by log-sum-exp (exact in exact arithmetic) and the output is written as bf16
//   hi + lo (the A operand of the out-projection contraction).
// Prefix keys beyond the 128-row tile (voice prompts longer than 128 frames) are read by the unit's own warp from the fp32
// prefix cache in L2.
// Warps 0 and 1 first do the prefix part of their 16 units, then join the other 14 warps, which draw units from a per-chunk
// counter; a unit's warp waits for its prefix triple (a shared-memory flag, release / acquire at CTA scope) only at the merge.
#pragma once

namespace ptts {

constexpr int A6_TILE = 128;                 // prefix keys on the tensor-core path
constexpr int A6_CHUNK = 32;                 // units per chunk = two m16 row tiles
constexpr int A6_WARPS = MG_THREADS / 32;
// shared-memory map (bytes from the 1024-aligned ring base)
constexpr int A6_KP = 0, A6_VP = 16384, A6_QH = 32768, A6_QL = 36864, A6_QF = 40960, A6_PO = 49152, A6_PML = 57344, A6_CTL = 57600,
              A6_SC = 57664;
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
__device__ __forceinline__ void st_release_cta(int *p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta(const int *p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
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

// ---- prefix part of the 16 units of row tile mt (one warp): S = Q.Kp^T, online softmax over two halves of 64 keys,
// O = P.Vp; un-normalised (max, sum, O) per unit -> shared memory. Rows of absent units hold q = 0 (finite results, unused).
__device__ __noinline__ void a6_prefix_tile(uint8_t *sm, int mt, int n_tile, int lane) {
    const uint32_t kp = smem_u32(sm + A6_KP), vp = smem_u32(sm + A6_VP), qh = smem_u32(sm + A6_QH), ql = smem_u32(sm + A6_QL);
    const int g = lane >> 2, j = lane & 3;
    uint32_t ah[4][4], al[4][4];
    {
        const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            const uint32_t off = a6_swz(row, ks * 2 + (lane >> 4));
            ldsm_x4(qh + off, ah[ks]);
            ldsm_x4(ql + off, al[ks]);
        }
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.0f;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        const int key0 = half * 64;
        if (key0 >= n_tile) break;
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
            uint32_t kb[2][4];
            const int row = key0 + nt * 8 + (lane & 7);
            ldsm_x4(kp + a6_swz(row, lane >> 3), kb[0]);        // dims 0..31: b0 / b1 of k-steps 0, 1
            ldsm_x4(kp + a6_swz(row, 4 + (lane >> 3)), kb[1]);  // dims 32..63: k-steps 2, 3
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                mma_f16_16816(s[nt], ah[ks], kb[ks >> 1][(ks & 1) * 2], kb[ks >> 1][(ks & 1) * 2 + 1]);
                mma_f16_16816(s[nt], al[ks], kb[ks >> 1][(ks & 1) * 2], kb[ks >> 1][(ks & 1) * 2 + 1]);
            }
        }
        float b0 = -INFINITY, b1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const int key = key0 + nt * 8 + 2 * j;
            if (key >= n_tile) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
            if (key + 1 >= n_tile) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            b0 = fmaxf(b0, fmaxf(s[nt][0], s[nt][1]));
            b1 = fmaxf(b1, fmaxf(s[nt][2], s[nt][3]));
        }
        b0 = fmaxf(b0, __shfl_xor_sync(0xffffffffu, b0, 1)); b0 = fmaxf(b0, __shfl_xor_sync(0xffffffffu, b0, 2));
        b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, 1)); b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, 2));
        // key key0 is valid (key0 < n_tile), so the new maxima are finite
        const float n0 = fmaxf(m0, b0), n1 = fmaxf(m1, b1);
        const float f0 = expf(m0 - n0), f1 = expf(m1 - n1);   // exp(-inf) = 0 on the first half
        float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            s[nt][0] = expf(s[nt][0] - n0); s[nt][1] = expf(s[nt][1] - n0);
            s[nt][2] = expf(s[nt][2] - n1); s[nt][3] = expf(s[nt][3] - n1);
            r0 += s[nt][0] + s[nt][1];
            r1 += s[nt][2] + s[nt][3];
            o[nt][0] *= f0; o[nt][1] *= f0; o[nt][2] *= f1; o[nt][3] *= f1;
        }
        l0 = l0 * f0 + r0; l1 = l1 * f1 + r1;
        m0 = n0; m1 = n1;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {   // 16 keys per k-step: score tiles 2 kk and 2 kk + 1 are the A fragment of P
            uint32_t ph[4], pl[4];
            split_hl(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
            split_hl(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
            split_hl(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
            split_hl(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
            const int krow = key0 + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
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
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float *po = (float *)(sm + A6_PO), *pml = (float *)(sm + A6_PML);
    const int ua = mt * 16 + g, ub = ua + 8;
    if (j == 0) { pml[2 * ua] = m0; pml[2 * ua + 1] = l0; pml[2 * ub] = m1; pml[2 * ub + 1] = l1; }
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        *(float2 *)(po + ua * FL_HD + nt * 8 + 2 * j) = make_float2(o[nt][0], o[nt][1]);
        *(float2 *)(po + ub * FL_HD + nt * 8 + 2 * j) = make_float2(o[nt][2], o[nt][3]);
    }
}

// ---- own rows of unit u = (b, h) + the far prefix rows, merged with the unit's prefix triple; one warp.
__device__ __noinline__ void a6_own_unit(const AttnDecodeParams &p, int b, int h, int u, int step, int cond_len, int n_tile, uint8_t *sm,
                                            float *sc, int seq, int lane, unsigned long long *tm) {
    const int k4 = lane >> 3, l8 = lane & 7;
    long long c0 = tm ? clock64() : 0;
#define A6_STAMP(i) if (tm) { const long long c1 = clock64(); if (lane == 0) tm[i] += (unsigned long long)(c1 - c0); c0 = c1; }
    const int n_own = step + p.own_base[b] + 1, n_far = cond_len - n_tile;
    const __half *okb = p.ok + ((size_t)h * p.cap_b + b) * p.max_s * FL_HD, *ovb = p.ov + ((size_t)h * p.cap_b + b) * p.max_s * FL_HD;
    const float *qf = (const float *)(sm + A6_QF) + u * FL_HD;
    uint4 ra[8], rb[8];
    load_own32(okb, n_own, k4, l8, ra);
    float q8[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q8[i] = qf[l8 * 8 + i];
    // ---- scores: batches of 32 rows, the next batch in flight while this one is reduced
#define A6_SCORE(raw, r0)                                                   \
    _Pragma("unroll") for (int it = 0; it < 8; it++) {                      \
        float kf[8];                                                        \
        unpack8(raw[it], kf);                                               \
        const float a = dot8_reduce(q8, kf);                                \
        const int key = (r0) + it * 4 + k4;                                 \
        if (l8 == 0 && key < n_own) sc[key] = a;                            \
    }
#pragma unroll 1
    for (int r0 = 0; r0 < n_own; r0 += 64) {
        const bool more = r0 + 32 < n_own;
        if (more) load_own32(okb + (size_t)(r0 + 32) * FL_HD, n_own - r0 - 32, k4, l8, rb);
        A6_SCORE(ra, r0)
        if (more) {
            if (r0 + 64 < n_own) load_own32(okb + (size_t)(r0 + 64) * FL_HD, n_own - r0 - 64, k4, l8, ra);
            A6_SCORE(rb, r0 + 32)
        }
    }
#undef A6_SCORE
    if (n_far > 0) a6_far_qk(p.pk + ((size_t)h * p.pstride + n_tile) * FL_HD, n_far, qf, sc + n_own, k4, l8);
    load_own32(ovb, n_own, k4, l8, ra);   // first V batch requested before the softmax
    __syncwarp();
    A6_STAMP(1)
    // ---- softmax over [own rows | far prefix rows]: exp(s - max) in place, sum kept for the merge
    const int np = n_own + n_far;
    float mo = -INFINITY;
    for (int t = lane; t < np; t += 32) mo = fmaxf(mo, sc[t]);
    mo = warp_max(mo);
    float lo = 0.0f;
    for (int t = lane; t < np; t += 32) { const float e = expf(sc[t] - mo); sc[t] = e; lo += e; }
    lo = warp_sum(lo);
    __syncwarp();
    A6_STAMP(2)
    // ---- P.V
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.0f;
#define A6_PV(raw, r0)                                                      \
    _Pragma("unroll") for (int it = 0; it < 8; it++) {                      \
        const int key = (r0) + it * 4 + k4;                                 \
        const float pw = key < n_own ? sc[key] : 0.0f;                      \
        float vf[8];                                                        \
        unpack8(raw[it], vf);                                               \
        _Pragma("unroll") for (int i = 0; i < 8; i++) acc[i] += pw * vf[i]; \
    }
#pragma unroll 1
    for (int r0 = 0; r0 < n_own; r0 += 64) {
        const bool more = r0 + 32 < n_own;
        if (more) load_own32(ovb + (size_t)(r0 + 32) * FL_HD, n_own - r0 - 32, k4, l8, rb);
        A6_PV(ra, r0)
        if (more) {
            if (r0 + 64 < n_own) load_own32(ovb + (size_t)(r0 + 64) * FL_HD, n_own - r0 - 64, k4, l8, ra);
            A6_PV(rb, r0 + 32)
        }
    }
#undef A6_PV
    if (n_far > 0) a6_far_pv(p.pv + ((size_t)h * p.pstride + n_tile) * FL_HD, n_far, sc + n_own, k4, l8, acc);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    A6_STAMP(3)
    // ---- merge with the prefix triple of this unit (log-sum-exp), write bf16 hi + lo
    float wa = 0.0f, wb = 1.0f, lp = 0.0f;
    const float *po = (const float *)(sm + A6_PO) + u * FL_HD + l8 * 8;
    if (n_tile > 0) {
        const int *flag = (const int *)(sm + A6_CTL) + 1 + (u >> 4);
        if (lane == 0) {
            const long long t0 = clock64();
            while (ld_acquire_cta(flag) < seq)
                if (clock64() - t0 > 2000000000LL) { printf("ptts_b200: attention prefix flag timed out (block %d)\n", blockIdx.x); __trap(); }
        }
        __syncwarp();
        const float *pml = (const float *)(sm + A6_PML) + 2 * u;
        const float mp = pml[0];
        lp = pml[1];
        const float m = fmaxf(mp, mo);
        wa = expf(mp - m); wb = expf(mo - m);
    }
    const float inv = 1.0f / (lp * wa + lo * wb);
    if (k4 == 0) {   // lanes 0..7 hold the 64 outputs, 8 consecutive dimensions each
        float out[8];
        if (n_tile > 0) {
            const float4 pa = *(const float4 *)po, pb = *(const float4 *)(po + 4);
            out[0] = pa.x; out[1] = pa.y; out[2] = pa.z; out[3] = pa.w; out[4] = pb.x; out[5] = pb.y; out[6] = pb.z; out[7] = pb.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) out[i] = (out[i] * wa + acc[i] * wb) * inv;
        const size_t oi = (size_t)b * FL_D + h * FL_HD + l8 * 8;
        if (p.out_f32) {
            *(float4 *)(p.out_f32 + oi) = make_float4(out[0], out[1], out[2], out[3]);
            *(float4 *)(p.out_f32 + oi + 4) = make_float4(out[4], out[5], out[6], out[7]);
        }
        if (p.out_h) {
            uint4 hv, lv;
            uint32_t *hp = (uint32_t *)&hv, *lw = (uint32_t *)&lv;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint16_t h0 = f2h16(out[2 * i], true), h1 = f2h16(out[2 * i + 1], true);
                hp[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                const uint16_t l0 = f2h16(out[2 * i] - h162f(h0, true), true), l1 = f2h16(out[2 * i + 1] - h162f(h1, true), true);
                lw[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
            *(uint4 *)(p.out_h + oi) = hv;
            if (p.out_l) *(uint4 *)(p.out_l + oi) = lv;
        }
    }
    __syncwarp();   // score scratch reusable by this warp's next unit
    A6_STAMP(4)
    if (tm && lane == 0) tm[5] += 1;
#undef A6_STAMP
}

// One attention phase: units [u0, u1) of this CTA in head-major order (unit = h * B + b).
__device__ __forceinline__ void attn6_phase(const AttnDecodeParams &p, int B, int pcap, int step, uint8_t *sm, int tid, unsigned long long *tmw = nullptr) {
    const int w = tid >> 5, lane = tid & 31;
    int *ctl = (int *)(sm + A6_CTL);
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
    if (tid == 0) { ctl[1] = 0; ctl[2] = 0; }
    int seq = 0;
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
            if (c0 != ua) __syncthreads();                  // previous chunk's q rows and prefix triples no longer in use
            if (tid == 0) ctl[0] = 0;
            seq++;
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
            if (tmw && w == 2 && lane == 0) tmw[0] += (unsigned long long)(clock64() - t0);
            if (w < 2 && n_tile > 0 && w * 16 < nu) {
                t0 = (tmw && w == 0) ? clock64() : 0;
                a6_prefix_tile(sm, w, n_tile, lane);
                __syncwarp();
                if (lane == 0) st_release_cta(ctl + 1 + w, seq);
                if (tmw && w == 0 && lane == 0) { tmw[6] += (unsigned long long)(clock64() - t0); tmw[7] += 1; }
            }
            while (true) {
                int u = 0;
                if (lane == 0) u = atomicAdd(ctl, 1);
                u = __shfl_sync(0xffffffffu, u, 0);
                if (u >= nu) break;
                a6_own_unit(p, b0 + u, h, u, step, cond_len, n_tile, sm, sc, seq, lane, (tmw && w == 2) ? tmw : nullptr);
            }
        }
        ua = ub;
    }
}

}  // namespace ptts
