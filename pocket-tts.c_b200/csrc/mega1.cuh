// mega1.cuh - batch-1 streaming decode as ONE persistent launch (ptts_flowlm.c:1187-1251 for a single
// utterance): the bandwidth-bound GEMV chain of a frame step (45 matrix-vector products + 6 attention
// layers) runs inside a cooperative kernel, 148 CTAs x 16 warps, phases separated by the grid barrier
// of mega.cuh instead of 47 kernel boundaries.
//   GEMV : one warp per output row, bf16 weights streamed with 16-byte loads (a warp reads one
//          contiguous row), fp32 activation vector in shared memory, warp-shuffle reduction; LayerNorm /
//          adaLN prologue and bias / activation / gate / residual epilogue fused. The weights a warp
//          needs for the NEXT GEMV depend on nothing, so they are requested (into registers) before
//          the warp arrives at the barrier: the HBM round trip overlaps the barrier and the prologue.
//   ATTN : flash-decoding - (head, 32-key chunk) items spread over all warps of the grid, partial
//          (max, sum, acc) triples merged by the prologue of the out-projection GEMV.
//   HEAD / MID / TAIL : the per-utterance phases of mega.cuh on CTA 0.
#pragma once
#include "mega.cuh"

namespace ptts {

constexpr int M1_THREADS = 512, M1_WARPS = M1_THREADS / 32, M1_PF = 16;
constexpr int M1_MAX_K = 4096, M1_REC = 68;
constexpr int M1_DYN_SMEM = M1_WARPS * M1_PF * 32 * 16;   // look-ahead weight buffers: 8 KB per warp

enum Mega1PhaseType : int { M1_HEAD = 0, M1_GEMV, M1_ATTN, M1_MID, M1_TAIL, M1_COUNT };
struct Mega1Phase { int type, idx, next_gemv; };

struct Gemv1 {
    const __nv_bfloat16 *W;      // [N, K], K in {512, 1024, 4096}
    int N, K;
    const float *x;              // [K] input vector (global)
    int pro;                     // 0 none, 1 LayerNorm, 2 LayerNorm + adaLN, 3 merge of attention partials
    const float *ln_w, *ln_b;
    float eps;
    const float *shift, *scale;
    const float *bias;
    int act;
    const float *gate;
    float *acc_out, *out;        // acc_out[n] += y (residual stream) or out[n] = y
    int attn;                    // pro == 3: index of the attention layer whose partials are merged
};

struct Attn1 {
    AttnDecodeParams a;
    float *part;                 // [16 heads][max_chunks][M1_REC]: acc[64], max, sum (records 16-byte aligned)
    int max_chunks;
};

struct Mega1Program {
    MegaProgram base;            // HEAD / MID / TAIL parameters, control block, barrier, timers
    const Mega1Phase *phases;
    int n_phases;
    const Gemv1 *gemvs;
    const Attn1 *attn;
};

__device__ __forceinline__ float block_sum512(float v, float *red, int tid) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float t = (tid & 31) < M1_WARPS ? red[tid & 31] : 0.0f;
    return warp_sum(t);
}

// xs[0..K) = prologue(x): every CTA recomputes it (K floats from L2, two block reductions)
__device__ __forceinline__ void m1_prologue(const Gemv1 &g, const Attn1 *attn, int step, float *xs, float *red, int tid) {
    const int K = g.K;
    if (g.pro == 3) {
        // merge the flash-decoding partials of every head: thread t owns dimensions 2t, 2t+1
        const Attn1 &A = attn[g.attn];
        const int cond_len = A.a.ctrl_cond_len ? *A.a.ctrl_cond_len : 0;
        const int nch = (cond_len + 31) / 32 + (step + A.a.own_base[0] + 1 + 31) / 32;
        const int d = tid * 2, h = d >> 6;
        const float *ph = A.part + (size_t)h * A.max_chunks * M1_REC;
        float M = -INFINITY;
        for (int c = 0; c < nch; c++) M = fmaxf(M, ph[c * M1_REC + 64]);
        float L = 0.0f, o0 = 0.0f, o1 = 0.0f;
        for (int c = 0; c < nch; c++) {
            const float w = expf(ph[c * M1_REC + 64] - M);
            L += w * ph[c * M1_REC + 65];
            const float2 a = *(const float2 *)(ph + c * M1_REC + (d & 63));
            o0 += w * a.x; o1 += w * a.y;
        }
        xs[d] = o0 / L; xs[d + 1] = o1 / L;
        __syncthreads();
        return;
    }
    float v[8];
    const int per = K / M1_THREADS;   // 1, 2 or 8
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = j < per ? g.x[tid + j * M1_THREADS] : 0.0f;
    if (g.pro >= 1) {
        float lw[8], lb[8], lsc[8], lsh[8];   // independent of the statistics: in flight during the reductions
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int i = tid + j * M1_THREADS;
            const bool on = j < per;
            lw[j] = on && g.ln_w ? g.ln_w[i] : 1.0f;
            lb[j] = on && g.ln_b ? g.ln_b[i] : 0.0f;
            lsc[j] = on && g.pro == 2 ? g.scale[i] : 0.0f;
            lsh[j] = on && g.pro == 2 ? g.shift[i] : 0.0f;
        }
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) s += v[j];
        const float mean = block_sum512(s, red, tid) / (float)K;
        float q = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) if (j < per) { const float d = v[j] - mean; q += d * d; }
        const float inv = 1.0f / sqrtf(block_sum512(q, red, tid) / (float)K + g.eps);
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (j < per) v[j] = ((v[j] - mean) * inv * lw[j] + lb[j]) * (1.0f + lsc[j]) + lsh[j];
    }
#pragma unroll
    for (int j = 0; j < 8; j++) if (j < per) xs[tid + j * M1_THREADS] = v[j];
    __syncthreads();
}

__device__ __forceinline__ uint4 ldg_w(const __nv_bfloat16 *W, size_t row, int K, int c, int lane) {
    return __ldcs((const uint4 *)(W + row * K) + c * 32 + lane);
}
__device__ __forceinline__ float dot_chunk(const uint4 &u, const float *xs, int c, int lane) {
    const float4 a = *(const float4 *)(xs + (c * 32 + lane) * 8), b = *(const float4 *)(xs + (c * 32 + lane) * 8 + 4);
    return bf16lo(u.x) * a.x + bf16hi(u.x) * a.y + bf16lo(u.y) * a.z + bf16hi(u.y) * a.w + bf16lo(u.z) * b.x + bf16hi(u.z) * b.y +
           bf16lo(u.w) * b.z + bf16hi(u.w) * b.w;
}
struct M1Epi { float bias, gate, acc; };
__device__ __forceinline__ M1Epi m1_epi_load(const Gemv1 &g, int n) {   // issued before the dot product of the row
    M1Epi e;
    e.bias = g.bias ? g.bias[n] : 0.0f;
    e.gate = g.gate ? g.gate[n] : 1.0f;
    e.acc = g.acc_out ? g.acc_out[n] : 0.0f;
    return e;
}
__device__ __forceinline__ void m1_epilogue(const Gemv1 &g, int n, float y, const M1Epi &e) {
    y = act_apply(g.act, y + e.bias) * e.gate;
    if (g.acc_out) g.acc_out[n] = e.acc + y; else g.out[n] = y;
}

// Look-ahead: the first M1_PF 16-byte chunks per lane (row-major over this warp's rows) of GEMV `g` are copied
// global -> shared with cp.async (no registers held, nothing for the compiler to spill - a spilled register
// prefetch would wait for the load at the spill store), wbuf = this warp's [M1_PF][32] uint4.
template <int CH>   // 16-byte chunks per lane per row: 2, 4 or 16
__device__ __forceinline__ void m1_prefetch_t(const Gemv1 &g, int gw, int nw, int lane, uint4 *wbuf) {
    const uint32_t dst = smem_u32(wbuf + lane);
#pragma unroll
    for (int j = 0; j < M1_PF / CH; j++) {
        const int r = gw + j * nw;
        if (r < g.N) {
            const uint4 *src = (const uint4 *)(g.W + (size_t)r * g.K) + lane;
#pragma unroll
            for (int c = 0; c < CH; c++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (j * CH + c) * 512), "l"(src + c * 32) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void m1_prefetch(const Gemv1 &g, int gw, int nw, int lane, uint4 *wbuf) {
    if (g.K == 512) m1_prefetch_t<2>(g, gw, nw, lane, wbuf);
    else if (g.K == 1024) m1_prefetch_t<4>(g, gw, nw, lane, wbuf);
    else m1_prefetch_t<16>(g, gw, nw, lane, wbuf);
}

template <int CH>
__device__ __forceinline__ void m1_rows_t(const Gemv1 &g, const float *xs, int gw, int nw, int lane, const uint4 *wbuf) {
    constexpr int RP = M1_PF / CH;   // rows whose weights were fetched ahead
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    M1Epi ep[RP];
#pragma unroll
    for (int j = 0; j < RP; j++) { const int r = gw + j * nw; if (r < g.N && lane == 0) ep[j] = m1_epi_load(g, r); }
#pragma unroll
    for (int j = 0; j < RP; j++) {
        const int r = gw + j * nw;
        if (r < g.N) {
            float a = 0.0f;
#pragma unroll
            for (int c = 0; c < CH; c++) a += dot_chunk(wbuf[(j * CH + c) * 32 + lane], xs, c, lane);
            a = warp_sum(a);
            if (lane == 0) m1_epilogue(g, r, a, ep[j]);
        }
    }
    for (int r = gw + RP * nw; r < g.N; r += nw) {   // matrices with more rows per warp than the look-ahead covers
        uint4 w[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) w[c] = ldg_w(g.W, (size_t)r, g.K, c, lane);
        float a = 0.0f;
#pragma unroll
        for (int c = 0; c < CH; c++) a += dot_chunk(w[c], xs, c, lane);
        a = warp_sum(a);
        if (lane == 0) m1_epilogue(g, r, a, m1_epi_load(g, r));
    }
    __syncwarp();   // wbuf is free for the next look-ahead
}

// One 32-key chunk of one head: OWN = false -> rows [k0, k0+32) of the fp32 voice-prefix cache, OWN = true ->
// own rows [k0, k0+32) of the fp16 cache (row `s` is the position being generated: taken from the scratch, where
// it still is in fp32). Writes the chunk's (acc[64], max, sum) record.
template <bool OWN>
__device__ __forceinline__ void m1_attn_chunk(const AttnDecodeParams &p, int h, int k0, int n /*valid rows of the region*/, int s, const float *wsm,
                                              float *po, int lane) {
    const int k4 = lane >> 3, l8 = lane & 7;
    const float *pkh = p.pk + (size_t)h * p.pstride * FL_HD, *pvh = p.pv + (size_t)h * p.pstride * FL_HD;
    const __half *okb = p.ok + (size_t)h * p.cap_b * p.max_s * FL_HD, *ovb = p.ov + (size_t)h * p.cap_b * p.max_s * FL_HD;
    float rows[OWN ? 1 : 8][8];
    uint4 raw[OWN ? 8 : 1];
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int key = k0 + it * 4 + k4;
        if (OWN) raw[it] = key < min(n, s) ? *(const uint4 *)(okb + (size_t)key * FL_HD + l8 * 8) : make_uint4(0u, 0u, 0u, 0u);
        else ld_global_v8_pred(pkh + (size_t)key * FL_HD + l8 * 8, key < n, rows[it]);
    }
    float q8[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q8[i] = wsm[l8 * 8 + i];
    float sc[8], mx = -INFINITY;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int key = k0 + it * 4 + k4;
        float kf[8];
        if (OWN) {
            unpack8(raw[it], kf);
            if (key == s) {
#pragma unroll
                for (int i = 0; i < 8; i++) kf[i] = wsm[64 + l8 * 8 + i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) kf[i] = rows[it][i];
        }
        const float a = dot8_reduce(q8, kf);
        sc[it] = key < n ? a : -INFINITY;
        mx = fmaxf(mx, sc[it]);
    }
#pragma unroll
    for (int it = 0; it < 8; it++) {   // V rows, requested before the softmax math
        const int key = k0 + it * 4 + k4;
        if (OWN) raw[it] = key < min(n, s) ? *(const uint4 *)(ovb + (size_t)key * FL_HD + l8 * 8) : make_uint4(0u, 0u, 0u, 0u);
        else ld_global_v8_pred(pvh + (size_t)key * FL_HD + l8 * 8, key < n, rows[it]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    float sum = 0.0f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.0f;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int key = k0 + it * 4 + k4;
        const float e = key < n ? expf(sc[it] - mx) : 0.0f;
        sum += e;
        float vf[8];
        if (OWN) {
            unpack8(raw[it], vf);
            if (key == s) {
#pragma unroll
                for (int i = 0; i < 8; i++) vf[i] = wsm[128 + l8 * 8 + i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) vf[i] = rows[it][i];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] += e * vf[i];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 8);
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    if (k4 == 0) {
        *(float4 *)(po + l8 * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *(float4 *)(po + l8 * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (lane == 0) { po[64] = mx; po[65] = sum; }
}

// One attention layer for the single utterance: items (head, 32-key chunk of the prefix or of the own rows) over
// all warps of the grid. Record index: prefix chunks first, then own chunks.
__device__ __forceinline__ void m1_attn(const Attn1 &A, int step, float *wsm /*this warp: q[64] knew[64] vnew[64]*/, int gw, int nw, int lane) {
    const AttnDecodeParams &p = A.a;
    const int cond_len = p.ctrl_cond_len ? *p.ctrl_cond_len : 0;
    const int s = step + p.own_base[0], pos = cond_len + s;
    const int nchp = (cond_len + 31) / 32, ncho = (s + 1 + 31) / 32, nch = nchp + ncho, items = FL_H * nch;
    for (int item = gw; item < items; item += nw) {
        const int h = item / nch, c = item - h * nch;
        // new position: RoPE(q), RoPE(k), v (lane = dimension pair) -> this warp's scratch
        {
            const float *row = p.qkv + h * FL_HD;
            const float cs = p.rope_cos[pos * 32 + lane], sn = p.rope_sin[pos * 32 + lane];
            const float2 qq = *(const float2 *)(row + 2 * lane), kk = *(const float2 *)(row + FL_D + 2 * lane),
                         vv = *(const float2 *)(row + 2 * FL_D + 2 * lane);
            const float2 kr = make_float2(kk.x * cs - kk.y * sn, kk.x * sn + kk.y * cs);
            __syncwarp();
            *(float2 *)(wsm + 2 * lane) = make_float2((qq.x * cs - qq.y * sn) * 0.125f, (qq.x * sn + qq.y * cs) * 0.125f);
            *(float2 *)(wsm + 64 + 2 * lane) = kr;
            *(float2 *)(wsm + 128 + 2 * lane) = vv;
            if (c == nch - 1) {   // the item that owns the last chunk appends the row to the cache
                __half *okb = p.ok + (size_t)h * p.cap_b * p.max_s * FL_HD, *ovb = p.ov + (size_t)h * p.cap_b * p.max_s * FL_HD;
                *(__half2 *)(okb + (size_t)s * FL_HD + 2 * lane) = __floats2half2_rn(kr.x, kr.y);
                *(__half2 *)(ovb + (size_t)s * FL_HD + 2 * lane) = __floats2half2_rn(vv.x, vv.y);
            }
            __syncwarp();
        }
        float *po = A.part + ((size_t)h * A.max_chunks + c) * M1_REC;
        if (c < nchp) m1_attn_chunk<false>(p, h, c * 32, cond_len, s, wsm, po, lane);
        else m1_attn_chunk<true>(p, h, (c - nchp) * 32, s + 1, s, wsm, po, lane);
    }
}

__global__ void __launch_bounds__(M1_THREADS, 1) k_flow_mega1(const __grid_constant__ Mega1Program Q) {
    __shared__ __align__(16) float xs[M1_MAX_K];
    __shared__ float red[32];
    __shared__ float scratch[128];                    // row group 0: red[8] + lat[32]
    __shared__ __align__(16) float wsm[M1_WARPS * 192];
    const MegaProgram &P = Q.base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // warp 15 owns no rows and issues no look-ahead loads: its lane 0 runs the grid barrier, whose release fence
    // would otherwise wait for the weight loads the other warps have just put in flight
    constexpr int M1_WORKERS = M1_WARPS - 1;
    const bool worker = warp < M1_WORKERS;
    const int nw = (int)gridDim.x * M1_WORKERS, gw = worker ? (int)blockIdx.x * M1_WORKERS + warp : nw + (1 << 20);
    unsigned int epoch = 0;
    const int step0 = P.ctrl->step;
    const bool prof = P.prof != nullptr && blockIdx.x == 0 && tid == 0;
    extern __shared__ __align__(16) uint8_t m1_dyn[];
    uint4 *wpre = (uint4 *)m1_dyn + (size_t)warp * M1_PF * 32;   // this warp's look-ahead buffer (8 KB)
    // weights of the first GEMV of the program
    for (int pi = 0; pi < Q.n_phases; pi++)
        if (Q.phases[pi].type == M1_GEMV) { if (worker) m1_prefetch(Q.gemvs[Q.phases[pi].idx], gw, nw, lane, wpre); break; }
    int steps_done = 0;
    for (int si = 0; si < P.nsteps; si++) {
        const int step = step0 + si;
        for (int pi = 0; pi < Q.n_phases; pi++) {
            const Mega1Phase ph = Q.phases[pi];
            unsigned long long t0 = 0;
            if (prof) t0 = globaltimer_ns();
            switch (ph.type) {
                case M1_GEMV: {
                    const Gemv1 &g = Q.gemvs[ph.idx];
                    m1_prologue(g, Q.attn, step, xs, red, tid);
                    if (prof) P.prof[16] += globaltimer_ns() - t0;
                    if (worker) {
                        if (g.K == 512) m1_rows_t<2>(g, xs, gw, nw, lane, wpre);
                        else if (g.K == 1024) m1_rows_t<4>(g, xs, gw, nw, lane, wpre);
                        else m1_rows_t<16>(g, xs, gw, nw, lane, wpre);
                        // the next GEMV's weights: requested now, used after the barrier (and whatever phases lie in between)
                        if (prof) P.prof[17] += globaltimer_ns() - t0;
                        if (ph.next_gemv >= 0 && !(ph.next_gemv <= ph.idx && si + 1 == P.nsteps)) m1_prefetch(Q.gemvs[ph.next_gemv], gw, nw, lane, wpre);
                        if (prof) P.prof[18] += globaltimer_ns() - t0;
                    }
                    break;
                }
                case M1_ATTN:
                    if (worker) {
                        m1_attn(Q.attn[ph.idx], step, wsm + warp * 192, gw, nw, lane);
                        if (ph.next_gemv >= 0) m1_prefetch(Q.gemvs[ph.next_gemv], gw, nw, lane, wpre);   // out-projection weights
                    }
                    break;
                case M1_HEAD:
                    if (blockIdx.x == 0 && tid < 256) mega_head(P, 0, tid, step, scratch, scratch + 16, 1);
                    break;
                case M1_MID:
                    if (blockIdx.x == 0 && tid < 256) {
                        mega_flow_update(P, 0, tid, step, ph.idx, scratch + 16);
                        group_bar(1);
                        flow_input_proj(P, 0, tid, scratch + 16);
                    }
                    break;
                case M1_TAIL:
                    if (blockIdx.x == 0 && tid < 256) mega_tail(P, 0, tid, step, scratch, scratch + 16, 1);
                    break;
                default: break;
            }
            grid_sync(P.bar, epoch, false, M1_WORKERS * 32);
            if (prof) {
                P.prof[2 * ph.type] += globaltimer_ns() - t0;
                P.prof[2 * ph.type + 1] += 1;
            }
        }
        steps_done = si + 1;
        if (*(volatile int *)P.done_count >= P.n_live) break;
    }
    if (blockIdx.x == 0 && tid == 0) P.ctrl->step = step0 + steps_done;
}

// ---------------------------------------------------------------- host side
struct Mega1Host {
    uint64_t gen = ~0ull;
    int lsd = 0;
    Mega1Phase *d_phases = nullptr;
    Gemv1 *d_gemvs = nullptr;
    Attn1 *d_attn = nullptr;
    float *d_part = nullptr;
    Mega1Program prog = {};
    double gemv_bytes = 0.0;   // weight bytes of one frame step
};

static int mega1_build(Engine *e, Mega1Host &M, int lsd, const float *cbias) {
    void *old[] = {M.d_phases, M.d_gemvs, M.d_attn, M.d_part};
    for (void *q : old) e->dfree(q);
    M = Mega1Host();
    std::vector<Mega1Phase> ph;
    std::vector<Gemv1> gv;
    std::vector<Attn1> at;
    const int max_chunks = (e->capCond + 31) / 32 + (e->capS + 31) / 32 + 2;
    if (e->dalloc(&M.d_part, (size_t)FL_L * FL_H * max_chunks * M1_REC) != 0) return -1;
    auto add = [&](const __nv_bfloat16 *W, int N, int K, const float *x, Gemv1 g) {
        g.W = W; g.N = N; g.K = K; g.x = x;
        gv.push_back(g);
        ph.push_back({M1_GEMV, (int)gv.size() - 1, -1});
        M.gemv_bytes += 2.0 * N * K;
    };
    ph.push_back({M1_HEAD, 0, -1});
    for (int k = 0; k < lsd; k++) {
        Gemv1 g = {};
        g.bias = cbias + (size_t)k * FN_D; g.act = ACT_SILU; g.out = e->sc;
        add(e->cond_w, FN_D, FL_D, e->normed, g);
        g = {}; g.bias = e->ada_b; g.out = e->ada;
        add(e->ada_w, FN_ADA, FN_D, e->sc, g);
        for (int b = 0; b < FN_DEPTH; b++) {
            const float *ad = e->ada + (size_t)b * 3 * FN_D;
            g = {}; g.pro = 2; g.ln_w = e->FB[b].lnw; g.ln_b = e->FB[b].lnb; g.eps = 1e-6f; g.shift = ad; g.scale = ad + FN_D;
            g.bias = e->FB[b].m0b; g.act = ACT_SILU; g.out = e->fg;
            add(e->FB[b].m0w, FN_D, FN_D, e->fx, g);
            g = {}; g.bias = e->FB[b].m2b; g.gate = ad + 2 * FN_D; g.acc_out = e->fx;
            add(e->FB[b].m2w, FN_D, FN_D, e->fg, g);
        }
        const float *ad = e->ada + (size_t)FN_DEPTH * 3 * FN_D;
        g = {}; g.pro = 2; g.eps = 1e-6f; g.shift = ad; g.scale = ad + FN_D; g.bias = e->fin_b; g.out = e->flow;
        add(e->fin_w, FL_LAT, FN_D, e->fx, g);
        if (k + 1 < lsd) ph.push_back({M1_MID, k, -1});
    }
    ph.push_back({M1_TAIL, 0, -1});
    for (int l = 0; l < FL_L; l++) {
        Gemv1 g = {};
        g.pro = 1; g.ln_w = e->L[l].n1w; g.ln_b = e->L[l].n1b; g.eps = 1e-5f; g.out = e->qkv;
        add(e->L[l].inproj, 3 * FL_D, FL_D, e->x, g);
        Attn1 a = {};
        a.a = make_attn_params(e, l, false, 1);
        a.part = M.d_part + (size_t)l * FL_H * max_chunks * M1_REC; a.max_chunks = max_chunks;
        at.push_back(a);
        ph.push_back({M1_ATTN, (int)at.size() - 1, -1});
        g = {}; g.pro = 3; g.attn = l; g.acc_out = e->x;
        add(e->L[l].outproj, FL_D, FL_D, e->att, g);
        g = {}; g.pro = 1; g.ln_w = e->L[l].n2w; g.ln_b = e->L[l].n2b; g.eps = 1e-5f; g.act = ACT_GELU_ERF; g.out = e->ffh;
        add(e->L[l].l1, FL_FF, FL_D, e->x, g);
        g = {}; g.acc_out = e->x;
        add(e->L[l].l2, FL_D, FL_FF, e->ffh, g);
    }
    // look-ahead: a GEMV fetches the weights of the next GEMV unless an attention phase lies in between (that phase
    // needs the registers); the attention phase fetches them itself once its items are done
    const int np = (int)ph.size();
    for (int i = 0; i < np; i++) {
        if (ph[i].type != M1_GEMV && ph[i].type != M1_ATTN) continue;
        for (int d = 1; d <= np; d++) {
            const Mega1Phase &q = ph[(i + d) % np];
            if (q.type == M1_ATTN) break;
            if (q.type == M1_GEMV) { ph[i].next_gemv = q.idx; break; }
        }
    }
    if (e->dalloc(&M.d_phases, ph.size()) != 0 || e->dalloc(&M.d_gemvs, gv.size()) != 0 || e->dalloc(&M.d_attn, at.size()) != 0) return -1;
    PTTS_CUDA_OK(cudaMemcpy(M.d_phases, ph.data(), ph.size() * sizeof(Mega1Phase), cudaMemcpyHostToDevice));
    PTTS_CUDA_OK(cudaMemcpy(M.d_gemvs, gv.data(), gv.size() * sizeof(Gemv1), cudaMemcpyHostToDevice));
    PTTS_CUDA_OK(cudaMemcpy(M.d_attn, at.data(), at.size() * sizeof(Attn1), cudaMemcpyHostToDevice));
    Mega1Program &Q = M.prog;
    Q.phases = M.d_phases; Q.n_phases = np; Q.gemvs = M.d_gemvs; Q.attn = M.d_attn;
    MegaProgram &P = Q.base;
    P.B = 1;
    P.st = make_state(e, 1, lsd);
    P.ctrl = e->ctrl; P.eos_after = e->d_eos_after;
    RowNormParams r = {};
    r.x = e->x; r.D = FL_D; r.w = e->onorm_w; r.b = e->onorm_b; r.eps = 1e-5f; r.do_norm = 1; r.out_f32 = e->normed;
    P.onorm = r;
    P.eos_w = e->eos_w; P.eos_b = e->eos_b; P.inp_w = e->inp_w; P.inp_b = e->inp_b; P.fx = e->fx;
    P.flow = e->flow; P.flow_nsplit = 1; P.flow_stride = 0;
    P.in_lin = e->in_lin;
    r = {};
    r.x = e->x; r.D = FL_D; r.eps = 1e-5f; r.do_norm = 1;   // TAIL: x = input_linear . latent (LayerNorm 1 is the QKV prologue here)
    P.ln0 = r;
    M.gen = e->gen; M.lsd = lsd;
    return 0;
}

}  // namespace ptts
