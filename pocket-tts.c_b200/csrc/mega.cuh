// mega.cuh - the persistent FlowLM step kernel: the whole lock-step decode loop
// (ptts_flowlm.c:1187-1251, every frame of every utterance of a batch) in ONE cooperative launch.
//
// One CTA per SM walks a small phase program once per frame; phases are separated by a grid-wide
// barrier (a monotonic counter in global memory). A phase is one of
//   HEAD : per utterance - fold the last FFN2 partial sums, out_norm, EOS rule, noise -> latent,
//          flow input projection                               (ptts_flowlm.c:1197-1231, 710)
//   GEMM : one dense contraction on tcgen05 (TMA -> smem ring -> tcgen05.mma -> TMEM -> epilogue);
//          the ring, the mbarrier phases and the TMEM accumulators persist across phases, and the
//          producer fetches the NEXT contraction's weight tiles before it reaches the barrier
//   ROW  : per utterance - split-K fold (+ gate) + residual + LayerNorm / adaLN + bf16 hi/lo split
//   MID  : per utterance - Euler update k, flow input projection for k + 1    (lsd_steps > 1)
//   TAIL : per utterance - Euler update, record the frame, stop rule, x = input_linear . latent,
//          LayerNorm 1 of layer 0                              (ptts_flowlm.c:775-777, 1235-1241)
//   ATTN : RoPE + KV append + decode attention, one warp per (utterance, head) unit, no block barriers
// Nothing is launched per layer or per frame: kernel-boundary drains, TMEM allocation, barrier
// initialisation and cold prologues are paid once per generate call.
#pragma once
#include "flowlm.cuh"

namespace ptts {

constexpr int MG_EPI_WARPS = 12;                  // 3 per TMEM lane quadrant
constexpr int MG_THREADS = 512;                   // warp 0 TMA, warp 1 MMA, warps 2..13 epilogue; 16 warps -> 128 registers per thread
constexpr int MG_STAGES = 4;
constexpr int MG_A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB: A hi at +0, A lo at +16 KB, B at +32 KB
constexpr int MG_SLOT_BYTES = 3 * MG_A_BYTES;      // B tile <= 128 rows
constexpr int MG_RING_BYTES = MG_STAGES * MG_SLOT_BYTES;
constexpr int MG_MAX_BN = 128;
constexpr int MG_ATTN_EXTRA = 0;                  // shared memory beyond the ring would shrink L1 (it holds the rope tables, q/k/v partials and spills)
constexpr int MG_SMEM_BYTES = 1024 /*alignment*/ + 1024 /*barriers + scratch*/ + MG_RING_BYTES + MG_ATTN_EXTRA;

enum MegaPhaseType : int { MP_HEAD = 0, MP_GEMM, MP_ROW, MP_MID, MP_TAIL, MP_ATTN, MP_COUNT };

struct MegaPhase {
    int type;
    int idx;         // GEMM / ROW / ATTN: index into the respective table; MID: LSD step k just finished
    int next_gemm;   // GEMM whose first weight tiles the producer fetches after this phase's own loads (-1: none)
    int seq;         // GEMM phases: ordinal of the contraction within the frame (parity of the ring-state slots)
    int wraps;       // GEMM phases: next_gemm belongs to the NEXT frame
};

// What a contraction phase needs of a GemmOp: the three operand maps and the GemmParams fields the roles read (same names).
struct MegaGemmParams {
    int m_tiles_per_batch, m_tiles_total, n_tiles, iters_per_split, total_tiles, taps, kblocks_per_tap, a_rows, n_total, act, ldo, ldh, split_k;
    uint32_t idesc;
    uint32_t n_magic, m_magic;   // ceil(2^32 / n_tiles), ceil(2^32 / m_tiles_total): x / d = umulhi(x, magic) for x * d < 2^32
    const float *bias;
    float *out_f32;
    size_t split_stride;
    uint16_t *out_h, *out_l;
};
struct MegaGemm {
    CUtensorMap mapA, mapAlo, mapB;
    MegaGemmParams p;
    int bn, split;
};
__device__ __forceinline__ TileCoord mega_tile(const MegaGemmParams &p, int tile, int BN) {   // gemm_tile() on the compact record
    TileCoord t;
    // (divisor 1 has no 32-bit reciprocal: magic == 0 marks it)
    const int r = p.n_magic ? (int)__umulhi((uint32_t)tile, p.n_magic) : tile, nt = tile - r * p.n_tiles;
    t.z = p.m_magic ? (int)__umulhi((uint32_t)r, p.m_magic) : r;
    const int mt = r - t.z * p.m_tiles_total;
    t.batch = 0;                       // mega_build only admits plain linears: one batch, one tap
    t.m0 = mt * GEMM_BM;
    t.n0 = nt * BN;
    t.it0 = t.z * p.iters_per_split;
    t.it1 = min(p.kblocks_per_tap, t.it0 + p.iters_per_split);
    return t;
}
// Table capacities of the kernel parameter space (32 KB in all): lsd_steps <= 6 (the flow-head contractions and row programs of the
// LSD steps share their entries; only cond_embed, whose bias carries the time embedding, differs per step)
constexpr int MG_PHASES_C = 208, MG_ROWS_C = 20, MG_GEMMS_C = 44;

// The whole program travels in the kernel parameter space: a phase starts with its table entries, and a global load behind the grid
// barrier (the L1 is cold there) costs an L2 round trip 65 times per frame - the constant bank answers in a few cycles (measured
// 718 -> 697 us per frame, profiles/r02_step_kernel_experiments.txt). TMA reads its tensor maps from the parameter space as well.
struct MegaProgram {
    MegaPhase phases[MG_PHASES_C];
    RowNormParams rows[MG_ROWS_C];
    MegaGemm gemms[MG_GEMMS_C];
    int n_phases;
    int n_gemm_phases;              // contraction phases per frame (parity of the ring-state slots)
    AttnDecodeParams attn[FL_L];    // per layer, by value: the attention phase reads its parameters straight from the constant bank
    int B, pcap;                    // padded batch, score-buffer length
    FlowState st;
    FlowCtrl *ctrl;
    const int *eos_after;
    // HEAD
    RowNormParams onorm;
    const float *eos_w, *eos_b;
    const __nv_bfloat16 *inp_w;     // [512, 32]
    const float *inp_b;
    float *fx;                      // [B, 512]
    // flow output of the final-layer GEMM: split-K partial sums [flow_nsplit][.] flow_stride floats apart
    const float *flow;
    int flow_nsplit;
    size_t flow_stride;
    // TAIL
    const __nv_bfloat16 *in_lin;    // [1024, 32]
    RowNormParams ln0;              // LayerNorm 1 of layer 0 (x -> xn hi/lo); ln0.x is the residual stream
    // control
    unsigned int *bar;              // grid barrier counter, zeroed by the host before the launch
    int *done_count;                // utterances that reached their stop point, zeroed by the host
    int n_live, nsteps, first_gemm;
    int wrap_prefetch;              // 1: the last contraction of a frame may fetch weights of the next frame's first one
                                    // (only when the loop cannot stop early, i.e. EOS detection is off)
    unsigned long long *prof;       // optional [MP_COUNT][2]: ns and phase count per type (written by CTA 0)
    unsigned long long *trace;      // optional [n_phases][2]: ns of CTA 0's own work / ns including the barrier, summed over frames
};

static_assert(sizeof(MegaProgram) <= 32764, "the step program must fit the kernel parameter space");

// ---------------------------------------------------------------- sync primitives
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Sum over a 256-thread group (8 warps); `red` holds 8 floats owned by the group.
__device__ __forceinline__ float group_sum(float v, float *red, int tid, int bar_id) {
    const int lane = tid & 31, w = tid >> 5;
    v = warp_sum(v);
    group_bar(bar_id);
    if (lane == 0) red[w] = v;
    group_bar(bar_id);
    return warp_sum(lane < 8 ? red[lane] : 0.0f);
}

// Grid-wide barrier: every CTA arrives once on a monotonic counter; `epoch` is the value the counter
// reaches when all CTAs have arrived. Global writes made before the barrier by any thread of any CTA
// are visible after it to generic loads and to TMA (async proxy) loads.
__device__ __forceinline__ void grid_sync(unsigned int *bar, unsigned int &epoch, int bar_tid = 0) {
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncthreads();
    epoch += gridDim.x;
    if ((int)threadIdx.x == bar_tid) {   // this thread's fence waits for ITS outstanding loads too: pick one that has none in flight
        __threadfence();
        atomicAdd(bar, 1u);
        uint32_t spins = 0;   // bounded: an L2 acquire load takes ~0.5 us, 2^23 of them several seconds
        while (ld_acquire_u32(bar) < epoch) {
            if (++spins == (1u << 23)) __trap();
        }
        // no fence here: the acquire load synchronises with the other CTAs' releases and the CTA barrier below extends that to this
        // CTA's threads (an explicit fence.acq_rel.gpu after the loop costs 0.3 us per barrier: 604 vs 583 us per frame)
    }
    __syncthreads();
}

// ---------------------------------------------------------------- row helpers (256 threads per row)
// Thread tid owns the 4 consecutive elements tid*4 .. tid*4+3 of a row (D <= 1024): every access is a
// 16-byte vector, and the split-K partial sums are fetched with independent loads (all in flight at
// once) before they are added.
constexpr int ROW_MAX_SPLIT = 8;
__device__ __forceinline__ void row_load_fold(const RowNormParams &p, size_t row, int tid, float (&v)[4]) {
    const int i = tid * 4;
    v[0] = v[1] = v[2] = v[3] = 0.0f;
    if (i >= p.D) return;
    float *xr = p.x + row * p.D + i;
    float4 t = *(const float4 *)xr;
    if (p.nsplit > 0) {
        float4 a[ROW_MAX_SPLIT];
        const float *pp = p.part + row * p.D + i;
#pragma unroll
        for (int z = 0; z < ROW_MAX_SPLIT; z++) a[z] = z < p.nsplit ? *(const float4 *)(pp + z * p.split_stride) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float4 g = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        if (p.gate) g = *(const float4 *)(p.gate + row * p.ldmod + i);
        float4 sum = a[0];
#pragma unroll
        for (int z = 1; z < ROW_MAX_SPLIT; z++) { sum.x += a[z].x; sum.y += a[z].y; sum.z += a[z].z; sum.w += a[z].w; }
        t.x += sum.x * g.x; t.y += sum.y * g.y; t.z += sum.z * g.z; t.w += sum.w * g.w;
        *(float4 *)xr = t;
    }
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ uint2 pack4_bf16(const float (&o)[4]) {
    uint2 r;   // one packed conversion per pair
    const __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
    r.x = *(const uint32_t *)&a;
    r.y = *(const uint32_t *)&b;
    return r;
}
// LayerNorm (+ adaLN) of the row held in v (ptts_flowlm.c:159-181, 735-737); y returns the outputs.
__device__ __forceinline__ void row_ln_store(const RowNormParams &p, size_t row, int tid, const float (&v)[4], float *red, int bar_id,
                                             float (&y)[4]) {
    const int i = tid * 4;
    const bool on = i < p.D;
    // affine / modulation vectors do not depend on the statistics: fetch them before the reductions
    float4 w4 = make_float4(1.0f, 1.0f, 1.0f, 1.0f), b4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), sc4 = b4, sh4 = b4;
    if (on) {
        if (p.w) w4 = *(const float4 *)(p.w + i);
        if (p.b) b4 = *(const float4 *)(p.b + i);
        if (p.scale) { sc4 = *(const float4 *)(p.scale + row * p.ldmod + i); sh4 = *(const float4 *)(p.shift + row * p.ldmod + i); }
    }
    // (fast division / rsqrt / exp intrinsics throughout the persistent kernel: 1-2 ulp, far inside the bf16 hi + lo operand rounding,
    // and none of the IEEE slow paths that cost instructions and registers)
    const float inv_d = __fdividef(1.0f, (float)p.D);
    const float mean = group_sum(on ? v[0] + v[1] + v[2] + v[3] : 0.0f, red, tid, bar_id) * inv_d;
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) { const float d = v[j] - mean; q += d * d; }
    const float var = group_sum(on ? q : 0.0f, red, tid, bar_id) * inv_d;
    const float inv = rsqrtf(var + p.eps);
    y[0] = y[1] = y[2] = y[3] = 0.0f;
    if (!on) return;
    const float ww[4] = {w4.x, w4.y, w4.z, w4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w}, ss[4] = {sc4.x, sc4.y, sc4.z, sc4.w},
                hh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float o = (v[j] - mean) * inv;
        if (p.w) o *= ww[j];
        if (p.b) o += bb[j];
        if (p.scale) o = o * (1.0f + ss[j]) + hh[j];
        y[j] = o;
    }
    if (p.out_f32) *(float4 *)(p.out_f32 + row * p.D + i) = make_float4(y[0], y[1], y[2], y[3]);
    if (p.out_h) {
        const uint2 hi = pack4_bf16(y);
        *(uint2 *)(p.out_h + row * p.D + i) = hi;
        // (the persistent kernel only runs with hi + lo activation pairs: out_l is always set with out_h)
        const float lo[4] = {y[0] - bf16lo(hi.x), y[1] - bf16hi(hi.x), y[2] - bf16lo(hi.y), y[3] - bf16hi(hi.y)};
        *(uint2 *)(p.out_l + row * p.D + i) = pack4_bf16(lo);
    }
}

// fx[b, n] = input_proj . lat + bias for n = tid, tid + 256 (ptts_flowlm.c:710); lat in shared memory
__device__ __forceinline__ void flow_input_proj(const MegaProgram &P, int b, int tid, const float *lat) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int n = tid + r * 256;
        const uint4 *wr = (const uint4 *)(P.inp_w + (size_t)n * FL_LAT);
        float a = P.inp_b[n];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint4 u = wr[c];
            const float *l = lat + c * 8;
            a += bf16lo(u.x) * l[0] + bf16hi(u.x) * l[1] + bf16lo(u.y) * l[2] + bf16hi(u.y) * l[3] + bf16lo(u.z) * l[4] +
                 bf16hi(u.z) * l[5] + bf16lo(u.w) * l[6] + bf16hi(u.w) * l[7];
        }
        P.fx[(size_t)b * FN_D + n] = a;
    }
}

// HEAD for utterance b at decode iteration `step` (k_rownorm + k_eos_state + k_flow_init(k = 0) fused).
__device__ __forceinline__ void mega_head(const MegaProgram &P, int b, int tid, int step, float *red, float *lat, int bar_id) {
    const FlowState &st = P.st;
    float v[4], y[4];
    row_load_fold(P.onorm, b, tid, v);
    row_ln_store(P.onorm, b, tid, v, red, bar_id, y);
    const float4 ew = *(const float4 *)(P.eos_w + tid * 4);
    const float eos = group_sum(ew.x * y[0] + ew.y * y[1] + ew.z * y[2] + ew.w * y[3], red, tid, bar_id) + P.eos_b[0];
    const bool valid = step < st.max_frames[b] && !st.done[b];
    if (valid) {
        if (step == 0) *(float4 *)(st.first_cond + (size_t)b * FL_D + tid * 4) = make_float4(y[0], y[1], y[2], y[3]);
        if (tid == 0) {
            if (step == 0) st.first_eos[b] = eos;
            if (st.eos_log) st.eos_log[(size_t)b * st.max_frames_cap + step] = eos;
            if (P.ctrl->eos_enabled && step + 1 >= P.ctrl->eos_min_frames && eos >= P.ctrl->eos_threshold && st.eos_step[b] < 0)
                st.eos_step[b] = step;
        }
    }
    if (tid < FL_LAT) {
        const int fi = min(max(step, 0), st.max_frames_cap - 1);
        const float z = st.noise[((size_t)b * st.max_frames_cap + fi) * FL_LAT + tid];
        st.latent[b * FL_LAT + tid] = z;
        lat[tid] = z;
    }
    group_bar(bar_id);
    flow_input_proj(P, b, tid, lat);
}

// Euler update of LSD step k for utterance b (k_flow_update): returns the new latent in lat[] (threads < 32).
__device__ __forceinline__ void mega_flow_update(const MegaProgram &P, int b, int tid, int step, int k, float *lat) {
    const FlowState &st = P.st;
    if (tid < FL_LAT) {
        const bool valid = step < st.max_frames[b] && !st.done[b];
        float f = 0.0f;
        for (int z = 0; z < P.flow_nsplit; z++) f += P.flow[z * P.flow_stride + (size_t)b * FL_LAT + tid];
        const float v = st.latent[b * FL_LAT + tid] + __fdividef(f, (float)st.lsd_steps);
        st.latent[b * FL_LAT + tid] = v;
        lat[tid] = v;
        if (valid) {
            if (step == 0 && k == 0) st.first_flow[b * FL_LAT + tid] = f;
            if (k == st.lsd_steps - 1) st.out_latents[((size_t)b * st.max_frames_cap + step) * FL_LAT + tid] = v;
        }
        __syncwarp();
        if (k == st.lsd_steps - 1 && tid == 0 && valid) {
            st.used[b] = step + 1;
            const int es = st.eos_step[b];
            if ((es >= 0 && step >= es + P.eos_after[b]) || step + 1 >= st.max_frames[b]) {
                st.done[b] = 1;
                atomicAdd(P.done_count, 1);
            }
        }
    }
}

// TAIL for utterance b (k_flow_update(last k) + k_prepare_x + LayerNorm 1 of layer 0 fused).
__device__ __forceinline__ void mega_tail(const MegaProgram &P, int b, int tid, int step, float *red, float *lat, int bar_id) {
    mega_flow_update(P, b, tid, step, P.st.lsd_steps - 1, lat);
    group_bar(bar_id);
    float v[4], y[4];
    float *xr = P.ln0.x + (size_t)b * FL_D;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = tid * 4 + j;
        const uint4 *wr = (const uint4 *)(P.in_lin + (size_t)i * FL_LAT);
        float a = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint4 u = wr[c];
            const float *l = lat + c * 8;
            a += bf16lo(u.x) * l[0] + bf16hi(u.x) * l[1] + bf16lo(u.y) * l[2] + bf16hi(u.y) * l[3] + bf16lo(u.z) * l[4] +
                 bf16hi(u.z) * l[5] + bf16lo(u.w) * l[6] + bf16hi(u.w) * l[7];
        }
        v[j] = a;
    }
    *(float4 *)(xr + tid * 4) = make_float4(v[0], v[1], v[2], v[3]);
    row_ln_store(P.ln0, b, tid, v, red, bar_id, y);
}


}  // namespace ptts
#include "attn.cuh"
namespace ptts {

// ---------------------------------------------------------------- GEMM phase
struct MegaPipe {
    int s;          // ring slot (producer and MMA threads keep their own copy)
    uint32_t ph;    // ring phase parity
    int lt;         // accumulator tiles processed by this CTA (MMA and epilogue threads)
    int pre;        // producer: leading items of the coming GEMM whose B tile is already in flight
};

__device__ __forceinline__ uint32_t mg_stage_bytes(const MegaGemm &g) {
    return (uint32_t)(MG_A_BYTES * 2 + g.bn * GEMM_BK * 2);   // A hi + A lo + B
}

__device__ __forceinline__ void mega_gemm(const MegaGemm &g, const MegaGemm *next, MegaPipe &pipe, uint8_t *ring, uint64_t *full,
                                          uint64_t *empty, uint64_t *tfull, uint64_t *tempty, uint32_t tmem_base) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const MegaGemmParams &p = g.p;
    const int BN = g.bn;
    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            asm volatile("fence.proxy.async.global;" ::: "memory");
            const uint32_t sb = mg_stage_bytes(g);
            int n = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord t = mega_tile(p, tile, BN);
                for (int it = t.it0; it < t.it1; it++, n++) {
                    uint8_t *st = ring + pipe.s * MG_SLOT_BYTES;
                    if (n >= pipe.pre) {
                        mbar_wait(&empty[pipe.s], pipe.ph ^ 1);
                        mbar_expect_tx(&full[pipe.s], sb);
                    }
                    tma_load_3d(st, &g.mapA, &full[pipe.s], it * GEMM_BK, t.m0, 0);
                    tma_load_3d(st + MG_A_BYTES, &g.mapAlo, &full[pipe.s], it * GEMM_BK, t.m0, 0);
                    if (n >= pipe.pre) tma_load_3d(st + 2 * MG_A_BYTES, &g.mapB, &full[pipe.s], it * GEMM_BK, t.n0, 0);
                    if (++pipe.s == MG_STAGES) { pipe.s = 0; pipe.ph ^= 1; }
                }
            }
            // weights of the next contraction do not depend on anything computed in between
            pipe.pre = 0;
            if (next) {
                // the next contraction's activation maps: into the TMA descriptor cache now (117 maps per frame thrash it; its
                // first operand load would otherwise start with a descriptor fetch behind the grid barrier): 584 -> 582 us per frame
                tma_prefetch_desc(&next->mapA); tma_prefetch_desc(&next->mapAlo);
                const uint32_t sbn = mg_stage_bytes(*next);
                int s2 = pipe.s; uint32_t ph2 = pipe.ph;
                for (int tile = blockIdx.x; tile < next->p.total_tiles && pipe.pre < MG_STAGES; tile += gridDim.x) {
                    const TileCoord t = mega_tile(next->p, tile, next->bn);
                    for (int it = t.it0; it < t.it1 && pipe.pre < MG_STAGES; it++, pipe.pre++) {
                        mbar_wait(&empty[s2], ph2 ^ 1);
                        mbar_expect_tx(&full[s2], sbn);
                        tma_load_3d(ring + s2 * MG_SLOT_BYTES + 2 * MG_A_BYTES, &next->mapB, &full[s2], it * GEMM_BK, t.n0, 0);
                        if (++s2 == MG_STAGES) { s2 = 0; ph2 ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one elected lane)
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, pipe.lt++) {
            const TileCoord t = mega_tile(p, tile, BN);
            const int acc = pipe.lt & 1;
            mbar_wait(&tempty[acc], ((pipe.lt >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * MG_MAX_BN;
            for (int it = t.it0; it < t.it1; it++) {
                mbar_wait(&full[pipe.s], pipe.ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(ring + pipe.s * MG_SLOT_BYTES);
                    const uint64_t ad = umma_desc_sw128(a_addr), ald = umma_desc_sw128(a_addr + MG_A_BYTES),
                                   bd = umma_desc_sw128(a_addr + 2 * MG_A_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; k++) {
                        umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, p.idesc, (it > t.it0 || k > 0) ? 1u : 0u);
                        umma_f16(d_tmem, ald + 2 * k, bd + 2 * k, p.idesc, 1u);
                    }
                    umma_commit(&empty[pipe.s]);
                    if (it == t.it1 - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++pipe.s == MG_STAGES) { pipe.s = 0; pipe.ph ^= 1; }
            }
        }
    } else if (warp < 2 + MG_EPI_WARPS) {
        // ------------------------------------------------ epilogue warps (row-major outputs only)
        const int quad = warp & 3, cg = (warp - 2) >> 2;
        constexpr int NCG = MG_EPI_WARPS / 4;
        const int act = p.act, ldo = p.ldo, ldh = p.ldh, n_total = p.n_total;
        const float *bias = p.bias;
        uint16_t *out_h = p.out_h, *out_l = p.out_l;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, pipe.lt++) {
            const TileCoord t = mega_tile(p, tile, BN);
            const int acc = pipe.lt & 1;
            mbar_wait(&tfull[acc], (pipe.lt >> 1) & 1);
            tc_fence_after();
            const int m = t.m0 + quad * 32 + lane;
            const bool m_ok = m < p.a_rows;
            const uint32_t trow = tmem_base + acc * MG_MAX_BN + ((uint32_t)(quad * 32) << 16);
            float *of = p.out_f32 ? p.out_f32 + (size_t)t.z * p.split_stride + (size_t)m * ldo : nullptr;
            uint16_t *oh = out_h ? out_h + (size_t)m * ldh : nullptr;
            uint16_t *ol = out_l + (size_t)m * ldh;   // set whenever out_h is (hi + lo pairs only)
            const bool add_bias = bias != nullptr && t.z == 0;
            uint32_t r[16];
#pragma unroll 1
            for (int c = cg * 16; c < BN; c += 16 * NCG) {
                tmem_ld16(trow + c, r);
                tmem_ld_wait16(r);
                const int nb = t.n0 + c;
                if (m_ok && nb < n_total) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]);
                    if (add_bias) {
                        const float4 *bp = (const float4 *)(bias + nb);
#pragma unroll
                        for (int j = 0; j < 4; j++) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
                    }
                    if (of) {
                        uint32_t w0[8], w1[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) { w0[j] = __float_as_uint(v[j]); w1[j] = __float_as_uint(v[8 + j]); }
                        st_global_v8(of + nb, w0);
                        st_global_v8(of + nb + 8, w1);
                    }
                    if (oh) {
                        uint32_t hw[8], lw[8];
                        if (act == ACT_GELU_ERF) act_pack16<ACT_GELU_ERF, false, true, true>(v, hw, lw);
                        else if (act == ACT_SILU) act_pack16<ACT_SILU, true, true, true>(v, hw, lw);   // __expf / __fdividef
                        else act_pack16<ACT_NONE, false, true, true>(v, hw, lw);
                        st_global_v8(oh + nb, hw);
                        st_global_v8(ol + nb, lw);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
}

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(MG_THREADS, 1) k_flow_mega(const __grid_constant__ MegaProgram P) {
    extern __shared__ uint8_t mg_smem_raw[];
    // aligned by pointer + offset (not through an integer round trip): the compiler keeps the shared address space,
    // so tile / score / scratch accesses below compile to LDS / STS instead of generic loads and stores
    uint8_t *sbase = mg_smem_raw + ((1024u - (smem_u32(mg_smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = (uint64_t *)sbase;
    uint64_t *full = bars, *empty = bars + MG_STAGES, *tfull = bars + 2 * MG_STAGES, *tempty = tfull + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty + 2);
    float *scratch = (float *)(sbase + 256);                  // per row group: red[8] + lat[32]
    uint8_t *ring = sbase + 1024;                             // 1024-byte aligned; the attention phases also use MG_ATTN_EXTRA beyond it
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    constexpr uint32_t TMEM_COLS = 2 * MG_MAX_BN;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < MG_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], MG_EPI_WARPS); mbar_init(&tempty[1], MG_EPI_WARPS);
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

    // Ring / accumulator state of the GEMM roles lives in shared memory between the contractions (two slots, alternating per
    // contraction: a role writes the slot the NEXT contraction reads, so a late reader of this one never sees it), not in
    // registers that would stay allocated - or be spilled - across every other phase: [slot][0..2] producer s, ph, pre;
    // [4..6] MMA warp s, ph, lt; [8] epilogue lt
    int *pst = (int *)(sbase + 128);
    if (tid < 32) pst[tid] = 0;
    __syncthreads();
    unsigned int epoch = 0;
    const int step0 = P.ctrl->step;
    const int grp = tid >> 8;                       // row group 0, 1
    const int gt = tid & 255;
    float *red = scratch + grp * 64, *lat = red + 16;
    const bool prof = P.prof != nullptr && blockIdx.x == 0 && tid == 0;
    int steps_done = 0;

    // launch constants of the attention phases: the prefix length, and own_base of this CTA's units (fixed for the whole launch)
    // cached in shared memory - every grid barrier leaves the L1 cold, and a phase must not start with dependent L2 round trips
    const int cond_len0 = P.ctrl->cond_len;
    int *ob_cache = nullptr;
    int *a_units = (int *)(sbase + 100);   // this CTA's unit range of the attention phases and the head of its first unit
    {
        int au0, au1;
        attn6_units(P.B, au0, au1);
        if (tid == 0) { a_units[0] = au0; a_units[1] = au1; a_units[2] = au0 / P.B; }
        if (au1 - au0 <= 64 && au1 > au0 && au0 / P.B == (au1 - 1) / P.B) {
            ob_cache = (int *)(sbase + 768);
            if (tid < au1 - au0) ob_cache[tid] = P.attn[0].own_base[au0 - (au0 / P.B) * P.B + tid];
        }
        __syncthreads();
    }

    // the first contraction's weight tiles can be requested right away
    if (warp == 0 && lane == 0 && P.first_gemm >= 0) {
        const MegaGemm &g0 = P.gemms[P.first_gemm];
        const uint32_t sb0 = mg_stage_bytes(g0);
        int pre = 0;
        for (int tile = blockIdx.x; tile < g0.p.total_tiles && pre < MG_STAGES; tile += gridDim.x) {
            const TileCoord t = mega_tile(g0.p, tile, g0.bn);
            for (int it = t.it0; it < t.it1 && pre < MG_STAGES; it++, pre++) {
                mbar_expect_tx(&full[pre], sb0);
                tma_load_3d(ring + pre * MG_SLOT_BYTES + 2 * MG_A_BYTES, &g0.mapB, &full[pre], it * GEMM_BK, t.n0, 0);
            }
        }
        pst[2] = pre;   // read by the first contraction of the launch (slot parity 0)
    }

    for (int si = 0; si < P.nsteps; si++) {
        const int step = step0 + si;
        for (int pi = 0; pi < P.n_phases; pi++) {
            const MegaPhase ph = P.phases[pi];
            // phase start stamp of the profiling thread: parked in shared memory, not in two registers that would stay live (for
            // every thread) across the whole phase
            if (prof) *(volatile unsigned long long *)(sbase + 112) = globaltimer_ns();
            switch (ph.type) {
                case MP_GEMM: {
                    // a fetch that wraps into the next frame must be consumed: never past the end of the launch
                    const bool wraps = ph.wraps != 0;
                    const MegaGemm *nx = ph.next_gemm >= 0 && !(wraps && (!P.wrap_prefetch || si + 1 == P.nsteps)) ? &P.gemms[ph.next_gemm] : nullptr;
                    const int par = (si * P.n_gemm_phases + ph.seq) & 1;
                    const int *rd = pst + par * 16;
                    int *wr = pst + (par ^ 1) * 16;
                    MegaPipe pipe = {0, 0u, 0, 0};
                    if (warp == 0) { pipe.s = rd[0]; pipe.ph = (uint32_t)rd[1]; pipe.pre = rd[2]; }
                    else if (warp == 1) { pipe.s = rd[4]; pipe.ph = (uint32_t)rd[5]; pipe.lt = rd[6]; }
                    else pipe.lt = rd[8];
                    mega_gemm(P.gemms[ph.idx], nx, pipe, ring, full, empty, tfull, tempty, tmem_base);
                    if (lane == 0) {
                        if (warp == 0) { wr[0] = pipe.s; wr[1] = (int)pipe.ph; wr[2] = pipe.pre; }
                        else if (warp == 1) { wr[4] = pipe.s; wr[5] = (int)pipe.ph; wr[6] = pipe.lt; }
                        else if (warp == 2) wr[8] = pipe.lt;
                    }
                    break;
                }
                case MP_ROW:
                    {   // (all MG_THREADS = 512 threads: two row groups of 256)
                        const RowNormParams &rp = P.rows[ph.idx];
                        for (int b = (int)blockIdx.x + grp * (int)gridDim.x; b < P.B; b += 2 * (int)gridDim.x) {
                            float v[4], y[4];
                            row_load_fold(rp, b, gt, v);
                            if (rp.do_norm) row_ln_store(rp, b, gt, v, red, 1 + grp, y);
                        }
                    }
                    break;
                case MP_HEAD:
                        for (int b = (int)blockIdx.x + grp * (int)gridDim.x; b < P.B; b += 2 * (int)gridDim.x) {
                            mega_head(P, b, gt, step, red, lat, 1 + grp);
                            group_bar(1 + grp);   // lat[] reusable
                        }
                    break;
                case MP_MID:
                        for (int b = (int)blockIdx.x + grp * (int)gridDim.x; b < P.B; b += 2 * (int)gridDim.x) {
                            mega_flow_update(P, b, gt, step, ph.idx, lat);
                            group_bar(1 + grp);
                            flow_input_proj(P, b, gt, lat);
                            group_bar(1 + grp);
                        }
                    break;
                case MP_TAIL:
                        for (int b = (int)blockIdx.x + grp * (int)gridDim.x; b < P.B; b += 2 * (int)gridDim.x) {
                            mega_tail(P, b, gt, step, red, lat, 1 + grp);
                            group_bar(1 + grp);
                        }
                    break;
                case MP_ATTN:
                    attn6_phase(P.attn[ph.idx], P.B, step, cond_len0, ob_cache, a_units, ring, tid, (P.trace && blockIdx.x == 0) ? P.trace + 2 * P.n_phases : nullptr);
                    break;
                default: break;
            }
            if (prof && P.trace) *(volatile unsigned long long *)(sbase + 120) = globaltimer_ns();
            grid_sync(P.bar, epoch);
            if (prof) {
                const unsigned long long t2 = globaltimer_ns(), t0 = *(volatile unsigned long long *)(sbase + 112),
                                         t1 = *(volatile unsigned long long *)(sbase + 120);
                P.prof[2 * ph.type] += t2 - t0;
                P.prof[2 * ph.type + 1] += 1;
                if (P.trace) { P.trace[2 * pi] += t1 - t0; P.trace[2 * pi + 1] += t2 - t0; }
            }
        }
        steps_done = si + 1;
        // every utterance has reached its stop point: done_count only changes in TAIL phases, several
        // barriers before this read, so all CTAs take the same branch
        if (*(volatile int *)P.done_count >= P.n_live) break;
    }
    if (blockIdx.x == 0 && tid == 0) P.ctrl->step = step0 + steps_done;

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------- host side: program construction and launch
struct MegaHost {
    uint64_t gen = ~0ull;
    int B = 0, lsd = 0, precise = -1;
    bool fits = false;              // the program fits the parameter-space tables (otherwise the caller runs the multi-kernel path)
    MegaProgram prog = {};
    std::vector<MegaPhase> phases;
    std::vector<GemmOp> gemms;
    // algorithmic work of one frame step, per phase type (profile accounting)
    double flops[MP_COUNT] = {0}, bytes[MP_COUNT] = {0};
};

static AttnDecodeParams make_attn_params(Engine *e, int layer, bool out16, int qkv_nsplit) {
    AttnDecodeParams a = {};
    a.qkv = e->qkv; a.qkv_nsplit = qkv_nsplit; a.qkv_stride = (size_t)e->capB * 3 * FL_D;
    a.rope_cos = e->rope_cos; a.rope_sin = e->rope_sin;
    a.ctrl_cond_len = &e->ctrl->cond_len;
    a.own_base = e->d_own_base;
    a.pstride = e->capCond;
    a.pk = e->pk + (size_t)layer * FL_D * e->capCond; a.pv = e->pv + (size_t)layer * FL_D * e->capCond;
    a.pk16 = e->pk16 + (size_t)layer * FL_D * e->capCond; a.pv16 = e->pv16 + (size_t)layer * FL_D * e->capCond;
    { const int nl = (layer + 1) % FL_L; a.nk = e->ok + (size_t)nl * e->capB * FL_D * e->capS; a.nv = e->ov + (size_t)nl * e->capB * FL_D * e->capS; a.n_extra = nl == 0 ? 1 : 0; }
    a.ok = e->ok + (size_t)layer * e->capB * FL_D * e->capS; a.ov = e->ov + (size_t)layer * e->capB * FL_D * e->capS;
    a.max_s = e->capS; a.cap_b = e->capB;
    a.step = &e->ctrl->step;
    if (out16) { a.out_h = e->att_h; a.out_l = e->att_l; } else a.out_f32 = e->att;
    return a;
}

// Returns 0 (M.fits tells whether the program can run as k_flow_mega) or -1 on error.
// precise == 0 (single bf16 activations) is not a program of this kernel: M.fits = false and the caller takes the multi-kernel path
// (the kernel carries one operand form only - every variant it does not have to hold is register-allocation headroom).
static int mega_build(Engine *e, MegaHost &M, const FlowGemms &G, int B, int lsd, int precise) {
    M = MegaHost();
    std::vector<RowNormParams> rows;
    std::vector<AttnDecodeParams> attn;
    auto &ph = M.phases;
    auto &gm = M.gemms;
    const size_t ps = (size_t)e->capB * FL_D;
    // Table entries are shared where the work is identical: the flow-head contractions and row programs of every LSD step use
    // the same buffers and weights (only cond_embed differs, its bias carries the time embedding of the step).
    auto new_gemm = [&](const GemmOp &op) -> int {
        if (op.bn > MG_MAX_BN || op.split == 2 || op.p.taps != 1 || op.p.transposed || op.p.m_tiles_per_batch != op.p.m_tiles_total) {
            set_error("mega_build: contraction shape not supported by the persistent step kernel");
            return -1;
        }
        gm.push_back(op);
        return (int)gm.size() - 1;
    };
    auto gemm_phase = [&](int idx) {
        ph.push_back({MP_GEMM, idx, -1, 0, 0});
        const GemmOp &op = gm[idx];
        const double rows_ = op.p.a_rows, n = op.p.n_total, k = (double)op.p.kblocks_per_tap * GEMM_BK;
        M.flops[MP_GEMM] += 2.0 * rows_ * n * k;
        M.bytes[MP_GEMM] += 2.0 * n * k;   // weights once; activations stay in L2
    };
    auto add_gemm = [&](const GemmOp &op) -> int {
        const int idx = new_gemm(op);
        if (idx < 0) return -1;
        gemm_phase(idx);
        return 0;
    };
    auto new_row = [&](const RowNormParams &r) { rows.push_back(r); return (int)rows.size() - 1; };
    auto row_phase = [&](int idx) { ph.push_back({MP_ROW, idx, -1, 0, 0}); };
    auto add_row = [&](const RowNormParams &r) { row_phase(new_row(r)); };
    ph.push_back({MP_HEAD, 0, -1, 0, 0});
    int i_ada = -1, i_m0[FN_DEPTH], i_m2[FN_DEPTH], i_fin = -1, i_row[FN_DEPTH + 1];
    for (int k = 0; k < lsd; k++) {
        if (add_gemm(G.cond[k]) != 0) return -1;
        if (k == 0 && (i_ada = new_gemm(G.ada)) < 0) return -1;
        gemm_phase(i_ada);
        for (int b = 0; b <= FN_DEPTH; b++) {
            if (k == 0) {
                RowNormParams r = {};
                r.x = e->fx; r.D = FN_D; r.eps = 1e-6f; r.do_norm = 1; r.ldmod = FN_ADA;
                if (b > 0) { r.part = e->ft; r.nsplit = G.m2[b - 1].p.split_k; r.split_stride = (size_t)e->capB * FN_D; r.gate = e->ada + (size_t)(b - 1) * 3 * FN_D + 2 * FN_D; }
                if (b < FN_DEPTH) { r.w = e->FB[b].lnw; r.b = e->FB[b].lnb; }
                r.shift = e->ada + (size_t)b * 3 * FN_D; r.scale = r.shift + FN_D;
                r.out_h = e->fh_h; r.out_l = precise ? e->fh_l : nullptr; r.out_bf16 = 1;
                i_row[b] = new_row(r);
                if (b < FN_DEPTH && ((i_m0[b] = new_gemm(G.m0[b])) < 0 || (i_m2[b] = new_gemm(G.m2[b])) < 0)) return -1;
            }
            row_phase(i_row[b]);
            if (b < FN_DEPTH) { gemm_phase(i_m0[b]); gemm_phase(i_m2[b]); }
        }
        if (k == 0 && (i_fin = new_gemm(G.fin)) < 0) return -1;
        gemm_phase(i_fin);
        if (k + 1 < lsd) ph.push_back({MP_MID, k, -1, 0, 0});
    }
    ph.push_back({MP_TAIL, 0, -1, 0, 0});
    for (int l = 0; l < FL_L; l++) {
        RowNormParams r = {};
        if (l > 0) {
            r.x = e->x; r.D = FL_D; r.w = e->L[l].n1w; r.b = e->L[l].n1b; r.eps = 1e-5f; r.do_norm = 1;
            r.out_h = e->xn_h; r.out_l = precise ? e->xn_l : nullptr; r.out_bf16 = 1;
            r.part = e->partB; r.nsplit = (int)G.ff2[l - 1].p.split_k; r.split_stride = ps;
            add_row(r);
        }
        if (add_gemm(G.qkv[l]) != 0) return -1;
        attn.push_back(make_attn_params(e, l, true, G.qkv[l].p.split_k));
        ph.push_back({MP_ATTN, (int)attn.size() - 1, -1, 0, 0});
        if (add_gemm(G.outp[l]) != 0) return -1;
        r = {};
        r.x = e->x; r.D = FL_D; r.w = e->L[l].n2w; r.b = e->L[l].n2b; r.eps = 1e-5f; r.do_norm = 1;
        r.part = e->partA; r.nsplit = (int)G.outp[l].p.split_k; r.split_stride = ps;
        r.out_h = e->xn_h; r.out_l = precise ? e->xn_l : nullptr; r.out_bf16 = 1;
        add_row(r);
        if (add_gemm(G.ff1[l]) != 0 || add_gemm(G.ff2[l]) != 0) return -1;
    }
    // per contraction phase: its ordinal within the frame, and the next contraction reachable without an attention phase in between
    // (attention aliases the ring) - whose first weight tiles the producer requests early
    const int np = (int)ph.size();
    int first_gemm = -1, n_gp = 0;
    for (int i = 0; i < np; i++) {
        if (ph[i].type != MP_GEMM) continue;
        if (first_gemm < 0) first_gemm = ph[i].idx;
        ph[i].seq = n_gp++;
        for (int d = 1; d < np; d++) {
            const MegaPhase &q = ph[(i + d) % np];
            if (q.type == MP_ATTN) break;
            if (q.type == MP_GEMM) { ph[i].next_gemm = q.idx; ph[i].wraps = i + d >= np ? 1 : 0; break; }
        }
    }
    const int pcap = ((e->capCond + e->capS + 63) / 32) * 32;
    PTTS_CHECK(attn6_smem_bytes(pcap) <= (size_t)MG_RING_BYTES + MG_ATTN_EXTRA,
               "attention score buffers of %zu bytes exceed the persistent kernel's shared memory (sequence too long)",
               attn6_smem_bytes(pcap));
    M.gen = e->gen; M.B = B; M.lsd = lsd; M.precise = precise;
    M.fits = precise != 0 && np <= MG_PHASES_C && (int)rows.size() <= MG_ROWS_C && (int)gm.size() <= MG_GEMMS_C;
    if (!M.fits) return 0;
    MegaProgram &P = M.prog;
    for (int i = 0; i < np; i++) P.phases[i] = ph[i];
    for (size_t i = 0; i < rows.size(); i++) P.rows[i] = rows[i];
    for (size_t i = 0; i < gm.size(); i++) {
        MegaGemm &g = P.gemms[i];
        const GemmParams &q = gm[i].p;
        g.mapA = gm[i].mapA; g.mapAlo = gm[i].mapAlo; g.mapB = gm[i].mapB; g.bn = gm[i].bn; g.split = gm[i].split;
        g.p.m_tiles_per_batch = q.m_tiles_per_batch; g.p.m_tiles_total = q.m_tiles_total; g.p.n_tiles = q.n_tiles;
        g.p.iters_per_split = q.iters_per_split; g.p.total_tiles = q.total_tiles; g.p.taps = q.taps; g.p.kblocks_per_tap = q.kblocks_per_tap;
        g.p.a_rows = q.a_rows; g.p.n_total = q.n_total; g.p.act = q.act; g.p.ldo = q.ldo; g.p.ldh = q.ldh; g.p.split_k = q.split_k;
        g.p.n_magic = q.n_tiles > 1 ? (uint32_t)(((1ull << 32) + (uint64_t)q.n_tiles - 1) / (uint64_t)q.n_tiles) : 0u;
        g.p.m_magic = q.m_tiles_total > 1 ? (uint32_t)(((1ull << 32) + (uint64_t)q.m_tiles_total - 1) / (uint64_t)q.m_tiles_total) : 0u;
        g.p.idesc = q.idesc; g.p.bias = q.bias; g.p.out_f32 = q.out_f32; g.p.split_stride = q.split_stride; g.p.out_h = q.out_h; g.p.out_l = q.out_l;
    }
    P.n_phases = np; P.n_gemm_phases = n_gp;
    P.B = B; P.pcap = pcap;
    for (int l = 0; l < FL_L; l++) P.attn[l] = attn[l];
    P.st = make_state(e, B, lsd);
    P.ctrl = e->ctrl; P.eos_after = e->d_eos_after;
    RowNormParams r = {};
    r.x = e->x; r.D = FL_D; r.w = e->onorm_w; r.b = e->onorm_b; r.eps = 1e-5f; r.do_norm = 1;
    r.part = e->partB; r.nsplit = (int)G.ff2[FL_L - 1].p.split_k; r.split_stride = ps;
    r.out_f32 = e->normed; r.out_h = e->normed_h; r.out_l = precise ? e->normed_l : nullptr; r.out_bf16 = 1;
    P.onorm = r;
    P.eos_w = e->eos_w; P.eos_b = e->eos_b; P.inp_w = e->inp_w; P.inp_b = e->inp_b; P.fx = e->fx;
    P.flow = e->flow; P.flow_nsplit = G.fin.p.split_k; P.flow_stride = (size_t)e->capB * FL_LAT;
    P.in_lin = e->in_lin;
    r = {};
    r.x = e->x; r.D = FL_D; r.w = e->L[0].n1w; r.b = e->L[0].n1b; r.eps = 1e-5f; r.do_norm = 1;
    r.out_h = e->xn_h; r.out_l = precise ? e->xn_l : nullptr; r.out_bf16 = 1;
    P.ln0 = r;
    P.first_gemm = first_gemm;
    return 0;
}

}  // namespace ptts
