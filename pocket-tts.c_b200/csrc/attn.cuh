// attn.cuh - decode attention phase of the persistent FlowLM step kernel (included by mega.cuh).
//
// Semantics: ptts_flowlm.c:465-525 - RoPE(q), RoPE(k) at the new position, k / v appended to the cache, then
// softmax(q.K^T / 8).V per head over [shared voice prefix | the utterance's own rows].
//
// A CTA works on one head and a span of utterances (gridDim / 16 CTAs per head), in chunks of up to 32 units
// (unit = one utterance of that head); warp w owns units w and w + 16 of the chunk. The key set has three parts:
//   * own rows already in the cache (fp16, tok_len + 1 + frames-so-far rows per utterance, ~80 % of the phase's bytes, read
//     exactly once, straight from HBM): each warp streams its units' K and V rows through a private ring of shared-memory slots
//     filled by TMA bulk copies (cp.async.bulk, 8 K rows + 8 V rows = 2 KB per slot, completion on an mbarrier per slot). The
//     old rows depend on nothing computed in this step, so the first copies are issued before anything else happens in the
//     phase and a slot is refilled the moment it has been consumed: the copies stay in flight across every latency-bound stage
//     (new rows, prefix part, merge) and no register holds data in flight. The rows are folded with an ONLINE softmax (running
//     max / sum / accumulator per unit): no score buffer, no separate softmax stage, sequence length not limited by shared memory.
//   * the new row: k / v of this step, taken from shared memory where the warp that appended them to the cache left them.
//   * voice prefix (the SAME K / V rows for every utterance): a dense contraction over the batch.
//     S = Q[32 x 64] . Kp^T[64 x <=128] and O = P[32 x 128] . Vp[128 x 64] run on the tensor cores
//     (mma.sync.m16n8k16, fp16 operands, fp32 accumulate): q and the probabilities are carried as fp16 hi + lo pairs, the
//     prefix K / V as the fp16 copy made once per voice. The work is cut into (16 units) x (32 keys) slices, one per warp
//     (8 warps); each slice leaves an un-normalised (max, sum, O[64]) partial per unit in shared memory.
//   The partial softmaxes (4 prefix slices + own rows) are merged by log-sum-exp (exact in exact arithmetic) and the output is
//   written as bf16 hi + lo (the A operand of the out-projection contraction).
// Prefix keys beyond the 128-row tile (voice prompts longer than 128 frames) are read by the unit's own warp from the fp32
// prefix cache in L2 and folded into the same online softmax.
#pragma once

namespace ptts {

constexpr int A6_TILE = 128;                 // prefix keys on the tensor-core path
constexpr int A6_CHUNK = 32;                 // units per chunk = two m16 row tiles
constexpr int A6_WARPS = MG_THREADS / 32;
constexpr int A6_SLICES = 4;                 // 32-key slices of the prefix tile, one warp each per row tile
constexpr int A6_ROWS = 16;                  // own rows per ring slot (K rows + V rows): 2 KB per bulk copy - the TMA unit's rate of
                                             // small copies, not bytes, bounds the stream (1 KB copies: 17 B/clk per SM measured)
constexpr int A6_SLOTS = 2;                  // ring slots per warp
constexpr int A6_SLOT_BYTES = 2 * A6_ROWS * FL_HD * 2;
// shared-memory map (bytes from the 1024-aligned ring base). Slot 0 of every warp sits in ring A, slot 1 in ring B; the fp16 prefix
// tiles and the fp16 query fragments (dead after the prefix slices) alias ring B, whose first copies are issued after them.
constexpr int A6_QF = 0, A6_KN = 8192, A6_VN = 12288, A6_PML = 16384, A6_PO = A6_PML + A6_CHUNK * A6_SLICES * 2 * 4,
              A6_RING_A = A6_PO + A6_CHUNK * A6_SLICES * FL_HD * 4, A6_RING_B = A6_RING_A + A6_WARPS * A6_SLOT_BYTES,
              A6_KP = A6_RING_B, A6_VP = A6_KP + 16384, A6_QH = A6_VP + 16384, A6_QL = A6_QH + 4096,
              A6_BARS = A6_RING_B + A6_WARPS * A6_SLOT_BYTES, A6_END = A6_BARS + A6_WARPS * A6_SLOTS * 8;
static_assert(A6_QL + 4096 <= A6_BARS, "prefix tiles must fit inside ring B");
__host__ __device__ inline size_t attn6_smem_bytes(int) { return (size_t)A6_END; }   // independent of the sequence length

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
__device__ __forceinline__ void mbar_inval(uint64_t *bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }

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

// ---- new position of unit (b, h): fold the split-K partial sums of q / k / v, RoPE, append k / v to the own cache; q -> smem
// (fp32 for the own-row part, fp16 hi + lo in the fragment tile for the prefix part), the fp16-rounded k / v of the new row -> smem
// (exactly the values later steps read back from the cache). Lane = dimension pair.
// The loads of a warp's two units (partial sums, RoPE rows) are issued together: one L2 round trip instead of two (the L1 is cold
// after every grid barrier).
struct A6NewRowIn { float2 rq[3][4]; float cs, sn; };
__device__ __forceinline__ void a6_new_row_load(const AttnDecodeParams &p, int b, int h, int s, int cond_len, int lane, A6NewRowIn &in) {
    const float *row = p.qkv + (size_t)b * 3 * FL_D + h * FL_HD + 2 * lane;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int z = 0; z < 4; z++)
            in.rq[c][z] = (z == 0 || z < p.qkv_nsplit) ? *(const float2 *)(row + c * FL_D + z * p.qkv_stride) : make_float2(0.0f, 0.0f);
    const int pos = cond_len + s;   // s = rows already in the own cache = index of the new row
    in.cs = p.rope_cos[pos * 32 + lane]; in.sn = p.rope_sin[pos * 32 + lane];
}
__device__ __forceinline__ void a6_new_row(const AttnDecodeParams &p, int b, int h, int u, int s, uint8_t *sm, int lane, const A6NewRowIn &in) {
    const float *row = p.qkv + (size_t)b * 3 * FL_D + h * FL_HD + 2 * lane;
    const float2 (&rq)[3][4] = in.rq;
    const float cs = in.cs, sn = in.sn;
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
    const __half2 kh = __floats2half2_rn(f[1].x * cs - f[1].y * sn, f[1].x * sn + f[1].y * cs), vh = __floats2half2_rn(f[2].x, f[2].y);
    *(__half2 *)(okb + (size_t)s * FL_HD + 2 * lane) = kh;
    *(__half2 *)(ovb + (size_t)s * FL_HD + 2 * lane) = vh;
    *(__half2 *)(sm + A6_KN + (u * FL_HD + 2 * lane) * 2) = kh;
    *(__half2 *)(sm + A6_VN + (u * FL_HD + 2 * lane) * 2) = vh;
    const float qx = (f[0].x * cs - f[0].y * sn) * 0.125f, qy = (f[0].x * sn + f[0].y * cs) * 0.125f;   // 1 / sqrt(64) folded into q
    *(float2 *)((float *)(sm + A6_QF) + u * FL_HD + 2 * lane) = make_float2(qx, qy);
    uint32_t hi, lo;
    split_hl(qx, qy, hi, lo);
    const uint32_t off = a6_swz(u, lane >> 2) + (lane & 3) * 4;
    *(uint32_t *)(sm + A6_QH + off) = hi;
    *(uint32_t *)(sm + A6_QL + off) = lo;
}

// ---- prefix part, sliced: a warp takes row tile mt (16 units) and the 32 keys [32 ke, 32 ke + 32): S = Q.Kp^T over those keys, a
// local softmax, O = P.Vp; the un-normalised partial (max, sum, O[64]) per unit and slice goes to shared memory and is merged
// with the own-row part by log-sum-exp. Rows of absent units hold q = 0 (finite results, unused); a slice past the end of the
// prefix writes (-inf, 0, 0).
__device__ __forceinline__ void a6_prefix_slice(uint8_t *sm, int mt, int ke, int n_tile, int lane) {
    const uint32_t kp = smem_u32(sm + A6_KP), vp = smem_u32(sm + A6_VP), qh = smem_u32(sm + A6_QH), ql = smem_u32(sm + A6_QL);
    const int g = lane >> 2, j = lane & 3, key0 = ke * 32;
    float *po = (float *)(sm + A6_PO), *pml = (float *)(sm + A6_PML);
    const int ua = mt * 16 + g, ub = ua + 8;
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    if (key0 < n_tile) {
        float s[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
        {
            const int qrow = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {   // 16 dimensions per step
                uint32_t ah[4], al[4];
                const uint32_t off = a6_swz(qrow, ks * 2 + (lane >> 4));
                ldsm_x4(qh + off, ah);
                ldsm_x4(ql + off, al);
#pragma unroll
                for (int np = 0; np < 2; np++) {   // key tiles 2 np, 2 np + 1: one x4 load = (b0, b1) of both for this k-step
                    uint32_t kb[4];
                    const int row = key0 + np * 16 + ((lane >> 4) & 1) * 8 + (lane & 7);
                    ldsm_x4(kp + a6_swz(row, ks * 2 + ((lane >> 3) & 1)), kb);
                    mma_f16_16816(s[2 * np], ah, kb[0], kb[1]);
                    mma_f16_16816(s[2 * np], al, kb[0], kb[1]);
                    mma_f16_16816(s[2 * np + 1], ah, kb[2], kb[3]);
                    mma_f16_16816(s[2 * np + 1], al, kb[2], kb[3]);
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const int key = key0 + nt * 8 + 2 * j;
            if (key >= n_tile) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
            if (key + 1 >= n_tile) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
            m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        // key key0 is valid, so both maxima are finite
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            s[nt][0] = __expf(s[nt][0] - m0); s[nt][1] = __expf(s[nt][1] - m0);
            s[nt][2] = __expf(s[nt][2] - m1); s[nt][3] = __expf(s[nt][3] - m1);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {   // 16 keys per step: score tiles 2 kk and 2 kk + 1 are the A fragment of P
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

// ---- own rows. Online softmax state of the unit in flight: m (uniform over the warp), l and acc[8] per lane (partial over the rows
// the lane's row group k4 has seen; reduced over the four groups when the unit is complete). Lane (k4 = lane >> 3, l8 = lane & 7)
// handles rows 4 it + k4 of a block and dimensions [8 l8, 8 l8 + 8).
struct A6Acc { float m, l, acc[8]; };

// fold NIT x 4 rows whose K / V values are given as fp32 registers; rows >= nvalid are ignored
template <int NIT>
__device__ __forceinline__ void a6_fold(A6Acc &st, const float (&q8)[8], const float (&kf)[NIT][8], const float (&vf)[NIT][8], int nvalid, int k4) {
    float sc[NIT], cm = -INFINITY;
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const float a = dot8_reduce(q8, kf[it]);
        sc[it] = it * 4 + k4 < nvalid ? a : -INFINITY;
        cm = fmaxf(cm, sc[it]);
    }
    cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 8));
    cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 16));
    const float mn = fmaxf(st.m, cm);               // finite whenever nvalid >= 1 (row 0 belongs to some lane group)
    const float f = __expf(st.m - mn);                // exp(-inf) = 0 on the first block
    st.l *= f;
#pragma unroll
    for (int d = 0; d < 8; d++) st.acc[d] *= f;
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const bool ok = it * 4 + k4 < nvalid;
        const float pw = ok ? __expf(sc[it] - mn) : 0.0f;
        st.l += pw;
#pragma unroll
        for (int d = 0; d < 8; d++) st.acc[d] += pw * (ok ? vf[it][d] : 0.0f);   // stale shared memory beyond nvalid may hold NaN patterns
    }
    st.m = mn;
}

// far prefix rows [0, n) of fp32 K / V regions in L2 (voice prompts longer than the shared-memory tile)
__device__ __noinline__ void a6_far(A6Acc &st, const float *kb, const float *vb, int n, const float *qf, int k4, int l8) {
    float q8[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q8[i] = qf[l8 * 8 + i];
    for (int r0 = 0; r0 < n; r0 += 16) {
        float kf[4][8], vf[4][8];
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int key = r0 + it * 4 + k4;
            ld_global_v8_pred(kb + (size_t)key * FL_HD + l8 * 8, key < n, kf[it]);
            ld_global_v8_pred(vb + (size_t)key * FL_HD + l8 * 8, key < n, vf[it]);
        }
        a6_fold<4>(st, q8, kf, vf, n - r0, k4);
    }
}

// One warp, its units u_first and u_first + 16 of the chunk: the stream of ring slots (slot c of the stream = 8 old rows of K and V
// of one unit), then per unit the new row, the far prefix rows, the merge with the prefix partials and the write-out.
struct A6Stream {
    const __half *k0, *v0;      // unit 0 cache rows; unit 1 = + unit_stride
    size_t unit_stride;
    int n_old0, n_old1, nch0, total;
};
__device__ __forceinline__ uint8_t *a6_slot(uint8_t *ring, int slot) { return ring + slot * (A6_RING_B - A6_RING_A); }
__device__ __forceinline__ void a6_issue(const A6Stream &S, int c, uint8_t *ring, uint64_t *bars) {   // one lane
    const bool second = c >= S.nch0;
    const int r0 = (second ? c - S.nch0 : c) * A6_ROWS, n_old = second ? S.n_old1 : S.n_old0;
    const uint32_t bytes = (uint32_t)min(A6_ROWS, n_old - r0) * FL_HD * 2;
    const size_t off = (second ? S.unit_stride : 0) + (size_t)r0 * FL_HD;
    const int slot = c % A6_SLOTS;
    uint8_t *dst = a6_slot(ring, slot);
    mbar_expect_tx(&bars[slot], 2 * bytes);
    bulk_load_1d(dst, S.k0 + off, bytes, &bars[slot]);
    bulk_load_1d(dst + A6_SLOT_BYTES / 2, S.v0 + off, bytes, &bars[slot]);
}

// (arguments by value: a reference to a struct built from kernel parameters would force that struct into local memory, and every field
// access of the callee into an L1-cold local load)
struct A6Out { uint16_t *out_h, *out_l; const float *pk_far, *pv_far; };
__device__ __noinline__ void a6_own_stream(const A6Out p, const A6Stream S, int h, int b0, int u_first, int cond_len, int n_tile,
                                           uint8_t *sm, uint8_t *ring, uint64_t *bars, int lane, unsigned long long *tm) {
    const int k4 = lane >> 3, l8 = lane & 7, n_far = cond_len - n_tile;
#ifdef PTTS_MEGA_ATTN_STAMPS   // per-stage cycle stamps of CTA 0 / warp 2 (tools/mega_trace.py prints them): a build option, not a run-time one
    long long c0 = tm ? clock64() : 0;
#define A6_STAMP(i) if (tm) { const long long c1 = clock64(); if (lane == 0) tm[i] += (unsigned long long)(c1 - c0); c0 = c1; }
#else
    (void)tm;
#define A6_STAMP(i)
#endif
    A6Acc st;
    float q8[8];
    int c = 0;
#pragma unroll 1
    for (int x = 0; x < 2; x++) {
        const int u = u_first + x * A6_WARPS, nch = x ? S.total - S.nch0 : S.nch0, n_old = x ? S.n_old1 : S.n_old0;
        if (x == 1 && S.n_old1 < 0) break;   // no second unit
        const float *qf = (const float *)(sm + A6_QF) + u * FL_HD;
        st.m = -INFINITY; st.l = 0.0f;
#pragma unroll
        for (int d = 0; d < 8; d++) { st.acc[d] = 0.0f; q8[d] = qf[l8 * 8 + d]; }
#pragma unroll 1
        for (int i = 0; i < nch; i++, c++) {
            const int slot = c % A6_SLOTS;
            mbar_wait(&bars[slot], (uint32_t)(c / A6_SLOTS) & 1u);
            const uint8_t *src = a6_slot(ring, slot) + (k4 * FL_HD + l8 * 8) * 2;
            constexpr int NIT = A6_ROWS / 4;
            float kf[NIT][8], vf[NIT][8];
#pragma unroll
            for (int it = 0; it < NIT; it++) {
                unpack8(*(const uint4 *)(src + it * 4 * FL_HD * 2), kf[it]);
                unpack8(*(const uint4 *)(src + A6_SLOT_BYTES / 2 + it * 4 * FL_HD * 2), vf[it]);
            }
            __syncwarp();                                                   // every lane has read the slot
            if (lane == 0 && c + A6_SLOTS < S.total) a6_issue(S, c + A6_SLOTS, ring, bars);
            a6_fold<NIT>(st, q8, kf, vf, n_old - i * A6_ROWS, k4);
        }
        A6_STAMP(1)
        {   // the new row of this step, from shared memory (row group 0 takes it)
            float kf[1][8], vf[1][8];
            unpack8(*(const uint4 *)(sm + A6_KN + (u * FL_HD + l8 * 8) * 2), kf[0]);
            unpack8(*(const uint4 *)(sm + A6_VN + (u * FL_HD + l8 * 8) * 2), vf[0]);
            a6_fold<1>(st, q8, kf, vf, 1, k4);
        }
        if (n_far > 0) a6_far(st, p.pk_far, p.pv_far, n_far, qf, k4, l8);
        // ---- unit complete: reduce over the four row groups, merge with the prefix partials (log-sum-exp), write bf16 hi + lo
        st.l += __shfl_xor_sync(0xffffffffu, st.l, 8);
        st.l += __shfl_xor_sync(0xffffffffu, st.l, 16);
#pragma unroll
        for (int d = 0; d < 8; d++) {
            st.acc[d] += __shfl_xor_sync(0xffffffffu, st.acc[d], 8);
            st.acc[d] += __shfl_xor_sync(0xffffffffu, st.acc[d], 16);
        }
        A6_STAMP(3)
        if (k4 == 0) {   // lanes 0..7 hold the 64 outputs, 8 consecutive dimensions each
            float m = st.m;
            const float *pml = (const float *)(sm + A6_PML) + (size_t)u * A6_SLICES * 2;
            const float *po = (const float *)(sm + A6_PO) + (size_t)u * A6_SLICES * FL_HD + l8 * 8;
            if (n_tile > 0) {
#pragma unroll
                for (int e = 0; e < A6_SLICES; e++) m = fmaxf(m, pml[2 * e]);
            }
            const float wb = __expf(st.m - m);
            float den = st.l * wb, out[8];
#pragma unroll
            for (int d = 0; d < 8; d++) out[d] = st.acc[d] * wb;
            if (n_tile > 0) {
#pragma unroll
                for (int e = 0; e < A6_SLICES; e++) {
                    const float we = __expf(pml[2 * e] - m);   // exp(-inf) = 0 for slices past the end of the prefix
                    den += pml[2 * e + 1] * we;
                    const float4 pa = *(const float4 *)(po + e * FL_HD), pb = *(const float4 *)(po + e * FL_HD + 4);
                    out[0] += pa.x * we; out[1] += pa.y * we; out[2] += pa.z * we; out[3] += pa.w * we;
                    out[4] += pb.x * we; out[5] += pb.y * we; out[6] += pb.z * we; out[7] += pb.w * we;
                }
            }
            const float inv = __fdividef(1.0f, den);
#pragma unroll
            for (int d = 0; d < 8; d++) out[d] *= inv;
            const size_t oi = (size_t)(b0 + u) * FL_D + h * FL_HD + l8 * 8;
            {   // bf16 hi + lo pair (the only output form of the persistent kernel)
                uint4 hv, lv;
                uint32_t *hp = (uint32_t *)&hv, *lw = (uint32_t *)&lv;
#pragma unroll
                for (int d = 0; d < 4; d++) {
                    const __nv_bfloat162 hb = __floats2bfloat162_rn(out[2 * d], out[2 * d + 1]);
                    hp[d] = *(const uint32_t *)&hb;
                    const __nv_bfloat162 lb = __floats2bfloat162_rn(out[2 * d] - __low2float(hb), out[2 * d + 1] - __high2float(hb));
                    lw[d] = *(const uint32_t *)&lb;
                }
                *(uint4 *)(p.out_h + oi) = hv;
                *(uint4 *)(p.out_l + oi) = lv;
            }
        }
        A6_STAMP(4)
#ifdef PTTS_MEGA_ATTN_STAMPS
        if (tm && lane == 0) tm[5] += 1;
#endif
    }
#undef A6_STAMP
}

// Units [u0, u1) of this CTA in head-major order (unit = h * B + b): with at least one CTA per head, every CTA stays inside one
// head (one prefix tile per phase) and gridDim / 16 CTAs share a head's utterances.
__device__ __forceinline__ void attn6_units(int B, int &u0, int &u1) {
    const int cph = (int)gridDim.x / FL_H;
    if (cph >= 1) {
        const int h = (int)blockIdx.x / cph, j = (int)blockIdx.x % cph;
        if (h < FL_H) { u0 = h * B + (int)(((long long)j * B) / cph); u1 = h * B + (int)(((long long)(j + 1) * B) / cph); }
        else { u0 = u1 = 0; }
    } else {
        const int units = FL_H * B;
        u0 = (int)(((long long)blockIdx.x * units) / gridDim.x); u1 = (int)(((long long)(blockIdx.x + 1) * units) / gridDim.x);
    }
}

// One attention phase: units [u0, u1) of this CTA in head-major order (unit = h * B + b).
// ob_cache: own_base of this CTA's units [u0, u1) in shared memory (or NULL: read p.own_base), cond_len: the prefix length of the launch.
// (L2 prefetches of the rows - cp.async.bulk.prefetch.L2 at the start of the phase, or from the idle warps of the preceding GEMM
// phases - were measured and lose: issuing them costs more than they save, profiles/r02_attention_experiments.txt.)
__device__ __forceinline__ void attn6_phase(const AttnDecodeParams &p, int B, int step, int cond_len, const int *ob_cache, const int *units, uint8_t *sm, int tid,
                                            unsigned long long *tmw_ = nullptr) {
#ifdef PTTS_MEGA_ATTN_STAMPS
    unsigned long long *const tmw = tmw_;
#else
    constexpr unsigned long long *tmw = nullptr; (void)tmw_;
#endif
    const int w = tid >> 5, lane = tid & 31;
    uint8_t *ring = sm + A6_RING_A + w * A6_SLOT_BYTES;   // slot 0; slot 1 is the same offset inside ring B
    uint64_t *bars = (uint64_t *)(sm + A6_BARS) + w * A6_SLOTS;
    const int n_tile = min(cond_len, A6_TILE);
    const int u0 = units[0], u1 = units[1];   // launch constants, computed once by the kernel prologue (two 64-bit divisions)
    const long long t_entry = (tmw && w == 2) ? clock64() : 0;
    int h = units[2];                         // head of unit u0; a CTA that crosses into the next head continues at its first unit
    for (int ua = u0; ua < u1; h++) {
        const int ub = min(u1, (h + 1) * B);   // sub-range of one head
        for (int c0 = ua; c0 < ub; c0 += A6_CHUNK) {
            const int nu = min(A6_CHUNK, ub - c0), b0 = c0 - h * B;
            __syncthreads();   // the previous chunk (or the previous phase's use of the ring memory) is over for every warp
            // ---- this warp's stream of old own rows: set up the ring barriers and get the first copies going
            A6Stream S;
            {
                const int ua0 = w, ua1 = w + A6_WARPS;
                const int bb0 = b0 + min(ua0, nu - 1);
                S.n_old0 = ua0 < nu ? step + (ob_cache ? ob_cache[c0 - u0 + ua0] : p.own_base[bb0]) : -1;
                S.n_old1 = ua1 < nu ? step + (ob_cache ? ob_cache[c0 - u0 + ua1] : p.own_base[b0 + ua1]) : -1;
                S.k0 = p.ok + ((size_t)h * p.cap_b + bb0) * p.max_s * FL_HD;
                S.v0 = p.ov + ((size_t)h * p.cap_b + bb0) * p.max_s * FL_HD;
                S.unit_stride = (size_t)A6_WARPS * p.max_s * FL_HD;
                S.nch0 = S.n_old0 > 0 ? (S.n_old0 + A6_ROWS - 1) / A6_ROWS : 0;
                S.total = S.nch0 + (S.n_old1 > 0 ? (S.n_old1 + A6_ROWS - 1) / A6_ROWS : 0);
            }
            if (lane == 0) {
                for (int s = 0; s < A6_SLOTS; s++) mbar_init(&bars[s], 1);
                fence_barrier_init();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the ring memory was last touched through the generic proxy
                if (S.total > 0) a6_issue(S, 0, ring, bars);                   // slot 1 (ring B) holds the prefix tiles until the slices are done
            }
            __syncwarp();
            if (n_tile > 0) {   // fp16 prefix tiles of head h (re-staged for every chunk: ring B overwrites them) with cp.async: in flight
                                // while the new rows are computed; rows past the prefix zero-filled (P = 0 times garbage must stay 0)
                const __half *pkh = p.pk16 + (size_t)h * p.pstride * FL_HD, *pvh = p.pv16 + (size_t)h * p.pstride * FL_HD;
                const uint32_t kpa = smem_u32(sm + A6_KP), vpa = smem_u32(sm + A6_VP);
                for (int idx = tid; idx < A6_TILE * 8; idx += MG_THREADS) {
                    const int row = idx >> 3, cc = idx & 7;
                    const bool ok = row < n_tile;
                    const size_t so = ok ? (size_t)row * FL_HD + cc * 8 : 0;
                    const uint32_t sz = ok ? 16u : 0u, off = a6_swz(row, cc);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(kpa + off), "l"(pkh + so), "r"(sz) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vpa + off), "l"(pvh + so), "r"(sz) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            long long t0 = (tmw && w == 2) ? clock64() : 0;
            if (tmw && w == 2 && lane == 0 && c0 == u0) tmw[2] += (unsigned long long)(t0 - t_entry);
            {
                A6NewRowIn in0, in1;
                const bool live0 = w < nu, live1 = w + A6_WARPS < nu;
                if (live0) a6_new_row_load(p, b0 + w, h, S.n_old0, cond_len, lane, in0);
                if (live1) a6_new_row_load(p, b0 + w + A6_WARPS, h, S.n_old1, cond_len, lane, in1);
                if (live0) a6_new_row(p, b0 + w, h, w, S.n_old0, sm, lane, in0);
                if (live1) a6_new_row(p, b0 + w + A6_WARPS, h, w + A6_WARPS, S.n_old1, sm, lane, in1);
                for (int u = w; u < A6_CHUNK; u += A6_WARPS)
                    if (u >= nu) {   // absent unit: a zero query keeps the row tile's arithmetic finite
                        const uint32_t off = a6_swz(u, lane >> 2) + (lane & 3) * 4;
                        *(uint32_t *)(sm + A6_QH + off) = 0u;
                        *(uint32_t *)(sm + A6_QL + off) = 0u;
                    }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            if (tmw && w == 2 && lane == 0) { const long long t1 = clock64(); tmw[0] += (unsigned long long)(t1 - t0); t0 = t1; }
            if (n_tile > 0) {
                if (w < 2 * A6_SLICES && (w & 1) * 16 < nu) a6_prefix_slice(sm, w & 1, w >> 1, n_tile, lane);
                __syncthreads();                            // all partials of the chunk are in shared memory, ring B is free
                if (tmw && w == 2 && lane == 0) { tmw[6] += (unsigned long long)(clock64() - t0); tmw[7] += 1; }
            }
            if (lane == 0 && S.total > 1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                a6_issue(S, 1, ring, bars);
            }
            __syncwarp();
            if (S.n_old0 >= 0) {
                const A6Out o = {p.out_h, p.out_l, p.pk ? p.pk + ((size_t)h * p.pstride + n_tile) * FL_HD : nullptr,
                                 p.pv ? p.pv + ((size_t)h * p.pstride + n_tile) * FL_HD : nullptr};
                a6_own_stream(o, S, h, b0, w, cond_len, n_tile, sm, ring, bars, lane, (tmw && w == 2) ? tmw : nullptr);
            }
            __syncwarp();
            if (lane == 0)
                for (int s = 0; s < A6_SLOTS; s++) mbar_inval(&bars[s]);
        }
        ua = ub;
    }
}

}  // namespace ptts
