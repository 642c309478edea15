// kernels.cuh - the non-GEMM kernels of the hot path: bf16-weight GEMV (batch-1 streaming decode),
// row LayerNorm / adaLN, RoPE + KV-append + decode attention, FlowLM state machine, flow-head
// glue, Mimi front end (quantizer proj + depthwise upsample), Mimi RoPE / banded attention,
// final conv to PCM.
#pragma once
#include "common.cuh"

namespace ptts {

constexpr int FL_D = 1024, FL_H = 16, FL_HD = 64, FL_L = 6, FL_FF = 4096, FL_LAT = 32;
constexpr int FN_D = 512, FN_DEPTH = 6, FN_ADA = FN_DEPTH * 3 * FN_D + 2 * FN_D;  // 10240 adaLN rows
constexpr int MI_D = 512, MI_H = 8, MI_L = 2, MI_FF = 2048, MI_CTX = 250;

// Block-wide sum for up to 1024 threads; `red` must hold 32 floats. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = lane < nw ? red[lane] : 0.0f;
    return warp_sum(t);
}
__device__ __forceinline__ float block_max(float v, float *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = lane < nw ? red[lane] : -INFINITY;
    return warp_max(t);
}

// =====================================================================================
// Row kernel: optional fold of split-K partial sums (+ gate) into the fp32 residual
// stream, LayerNorm (ptts_flowlm.c:159-181: biased variance, two-pass), optional adaLN
// modulation (ptts_flowlm.c:735-737), outputs fp32 and/or 16-bit hi(+lo).
// One block (256 threads) per row, D <= 1024.
// =====================================================================================
struct RowNormParams {
    float *x;                  // [R, D] fp32, updated in place when nsplit > 0
    int D;
    const float *part;         // [nsplit][R][D] partial sums to fold: x += gate * sum_z part_z
    int nsplit;
    size_t split_stride;
    const float *gate;         // optional [R, ldmod] multiplier of the folded sum (flow-head gate)
    const float *w, *b;        // LN affine (NULL = none)
    float eps;
    const float *shift, *scale;  // optional adaLN rows [R, ldmod]
    int ldmod;
    int do_norm;               // 0 = only fold
    float *out_f32;            // [R, D] or NULL
    uint16_t *out_h, *out_l;   // [R, D] or NULL
    int out_bf16;
};

__global__ void __launch_bounds__(256) k_rownorm(const RowNormParams p) {
    pdl_wait(); pdl_trigger();
    __shared__ float red[32];
    const size_t row = blockIdx.x;
    float v[4];
    float *xr = p.x + row * p.D;
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = threadIdx.x + j * 256;
        v[j] = 0.0f;
        if (i < p.D) {
            float t = xr[i];
            if (p.nsplit > 0) {
                float a = 0.0f;
                for (int z = 0; z < p.nsplit; z++) a += p.part[z * p.split_stride + row * p.D + i];
                if (p.gate) a *= p.gate[row * p.ldmod + i];
                t += a;
                xr[i] = t;
            }
            v[j] = t;
            s += t;
        }
    }
    if (!p.do_norm) return;
    const float mean = block_sum(s, red) / (float)p.D;
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = threadIdx.x + j * 256;
        if (i < p.D) { const float d = v[j] - mean; q += d * d; }
    }
    const float var = block_sum(q, red) / (float)p.D;
    const float inv = 1.0f / sqrtf(var + p.eps);
    const bool bf = p.out_bf16 != 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = threadIdx.x + j * 256;
        if (i < p.D) {
            float y = (v[j] - mean) * inv;
            if (p.w) y *= p.w[i];
            if (p.b) y += p.b[i];
            if (p.scale) y = y * (1.0f + p.scale[row * p.ldmod + i]) + p.shift[row * p.ldmod + i];
            if (p.out_f32) p.out_f32[row * p.D + i] = y;
            if (p.out_h) {
                const uint16_t h = f2h16(y, bf);
                p.out_h[row * p.D + i] = h;
                if (p.out_l) p.out_l[row * p.D + i] = f2h16(y - h162f(h, bf), bf);
            }
        }
    }
}

// =====================================================================================
// Batch-1 GEMV: y[n] = sum_k W[n,k] x[k], bf16 weights streamed once with 128-bit loads,
// fp32 activations and accumulation, warp-shuffle reduction. K is split across the CTA's
// threads (one 8-element chunk per thread per pass) so a warp reads 512 contiguous bytes of
// a weight row; each CTA owns RPC consecutive rows, giving RPC * (K/8/THREADS) independent
// 16-byte loads in flight per thread.
// Prologue (every CTA recomputes it; x is tiny): none | LayerNorm | LayerNorm + adaLN.
// Epilogue: + bias (+ bias2), activation, optional gate, store or in-place residual add.
// =====================================================================================
struct GemvParams {
    const __nv_bfloat16 *W;    // [N, K]
    int N;
    const float *x;            // [K]
    const float *ln_w, *ln_b;  // PRO >= 1
    float eps;
    const float *shift, *scale;  // PRO == 2
    const float *bias, *bias2;
    int act;
    const float *gate;         // y *= gate[n]
    float *acc_out;            // acc_out[n] += y   (residual stream), or
    float *out;                // out[n] = y
    int *inc;                  // optional: *inc += 1 by one thread of the grid (the frame counter, after the last GEMV of a step)
};

template <int K, int RPC, int PRO>
__global__ void __launch_bounds__((K / 8 < 128 ? K / 8 : 128)) k_gemv(const GemvParams p) {
    constexpr int THREADS = K / 8 < 128 ? K / 8 : 128;
    constexpr int NCH = K / 8 / THREADS;  // chunks per thread
    constexpr int NW = THREADS / 32;
    __shared__ float red[32];
    __shared__ float part[NW][RPC];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int r0 = blockIdx.x * RPC;

    // issue the weight loads first: they do not depend on the prologue
    uint4 wv[RPC][NCH];
#pragma unroll
    for (int r = 0; r < RPC; r++)
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const int row = min(r0 + r, p.N - 1);
            wv[r][c] = __ldcs((const uint4 *)(p.W + (size_t)row * K) + tid + c * THREADS);
        }
    pdl_wait();      // the weights above are constants; everything below depends on the previous kernel
    pdl_trigger();

    float xv[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const float4 a = *(const float4 *)(p.x + (tid + c * THREADS) * 8);
        const float4 b = *(const float4 *)(p.x + (tid + c * THREADS) * 8 + 4);
        xv[c][0] = a.x; xv[c][1] = a.y; xv[c][2] = a.z; xv[c][3] = a.w;
        xv[c][4] = b.x; xv[c][5] = b.y; xv[c][6] = b.z; xv[c][7] = b.w;
    }
    if (PRO >= 1) {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < NCH; c++)
#pragma unroll
            for (int j = 0; j < 8; j++) s += xv[c][j];
        const float mean = block_sum(s, red) / (float)K;
        float q = 0.0f;
#pragma unroll
        for (int c = 0; c < NCH; c++)
#pragma unroll
            for (int j = 0; j < 8; j++) { const float d = xv[c][j] - mean; q += d * d; }
        const float inv = 1.0f / sqrtf(block_sum(q, red) / (float)K + p.eps);
#pragma unroll
        for (int c = 0; c < NCH; c++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = (tid + c * THREADS) * 8 + j;
                float y = (xv[c][j] - mean) * inv;
                if (p.ln_w) y *= p.ln_w[i];
                if (p.ln_b) y += p.ln_b[i];
                if (PRO == 2) y = y * (1.0f + p.scale[i]) + p.shift[i];
                xv[c][j] = y;
            }
    }
    float acc[RPC];
#pragma unroll
    for (int r = 0; r < RPC; r++) {
        float a = 0.0f;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const uint4 u = wv[r][c];
            a += bf16lo(u.x) * xv[c][0] + bf16hi(u.x) * xv[c][1] + bf16lo(u.y) * xv[c][2] + bf16hi(u.y) * xv[c][3] +
                 bf16lo(u.z) * xv[c][4] + bf16hi(u.z) * xv[c][5] + bf16lo(u.w) * xv[c][6] + bf16hi(u.w) * xv[c][7];
        }
        acc[r] = warp_sum(a);
    }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < RPC; r++) part[w][r] = acc[r];
    }
    __syncthreads();
    if (tid < RPC && r0 + tid < p.N) {
        const int n = r0 + tid;
        float y = 0.0f;
#pragma unroll
        for (int i = 0; i < NW; i++) y += part[i][tid];
        if (p.bias) y += p.bias[n];
        if (p.bias2) y += p.bias2[n];
        y = act_apply(p.act, y);
        if (p.gate) y *= p.gate[n];
        if (p.acc_out) p.acc_out[n] += y; else p.out[n] = y;
    }
    if (p.inc && blockIdx.x == 0 && tid == 0) *p.inc += 1;
}

// =====================================================================================
// FlowLM generation state machine (ptts_flowlm.c:1138-1251 restructured for a batch).
// Prefill (voice prefix once per voice; text tokens + BOS of every utterance in ONE batched pass,
// k_pf_* below) leaves x[b] = hidden state at the BOS position and the own KV rows
// 0..tok_len[b]. Decode iteration j (device counter `step`) then does, for all utterances in
// lock-step:   head: out_norm(x) -> EOS rule -> noise -> LSD/flow -> latent_j   (frame j)
//              body: x = input_linear . latent_j -> transformer step at own position
//                    tok_len[b] + 1 + j  -> x for frame j+1
// =====================================================================================
struct FlowState {
    int *step;                 // device decode-iteration counter j (= frame index)
    int B, max_tok, max_frames_cap, cond_len;
    const int *tokens;         // [B, max_tok]
    const int *tok_len;        // [B]
    const int *max_frames;     // [B]
    const float *noise;        // [B, max_frames_cap, 32]
    float *latent;             // [B, 32] running latent of the current frame
    float *out_latents;        // [B, max_frames_cap, 32]
    int *eos_step, *done, *used;     // [B]
    float *first_eos, *first_cond, *first_flow;  // [B], [B,1024], [B,32]
    float *eos_log;            // optional [B, max_frames_cap]
    int lsd_steps;
};

// Device-resident per-call scalars, so that a captured step graph stays valid across calls.
struct FlowCtrl {
    int step, cond_len, eos_enabled, eos_min_frames, eos_after_default;
    float eos_threshold;
};

// x[b,:] = input_linear . latent[b] (ptts_flowlm.c:1240-1241). grid B, block 256.
__global__ void __launch_bounds__(256) k_prepare_x(const FlowState st, const __nv_bfloat16 *in_lin, float *x) {
    pdl_wait(); pdl_trigger();
    const int b = blockIdx.x;
    float *xr = x + (size_t)b * FL_D;
    __shared__ float lat[FL_LAT];
    if (threadIdx.x < FL_LAT) lat[threadIdx.x] = st.latent[b * FL_LAT + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.x; i < FL_D; i += 256) {
        const uint4 *wr = (const uint4 *)(in_lin + (size_t)i * FL_LAT);
        float a = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint4 u = wr[c];
            const float *l = lat + c * 8;
            a += bf16lo(u.x) * l[0] + bf16hi(u.x) * l[1] + bf16lo(u.y) * l[2] + bf16hi(u.y) * l[3] +
                 bf16lo(u.z) * l[4] + bf16hi(u.z) * l[5] + bf16lo(u.w) * l[6] + bf16hi(u.w) * l[7];
        }
        xr[i] = a;
    }
}

// EOS logit + stop rule (ptts_flowlm.c:1202-1209). grid B, block 256. normed [B,1024].
__global__ void __launch_bounds__(256) k_eos_state(const FlowState st, const FlowCtrl *ctrl, const int *eos_after, const float *normed,
                                                   const float *eos_w, const float *eos_b) {
    pdl_wait(); pdl_trigger();
    __shared__ float red[32];
    const int b = blockIdx.x, s = *st.step;
    const float *nr = normed + (size_t)b * FL_D;
    float a = 0.0f;
    for (int i = threadIdx.x; i < FL_D; i += 256) a += eos_w[i] * nr[i];
    const float eos = block_sum(a, red) + eos_b[0];
    const int i = s;
    const bool valid = i < st.max_frames[b] && !st.done[b];
    if (!valid) return;
    if (i == 0) {
        for (int k = threadIdx.x; k < FL_D; k += 256) st.first_cond[(size_t)b * FL_D + k] = nr[k];
    }
    if (threadIdx.x == 0) {
        if (i == 0) st.first_eos[b] = eos;
        if (st.eos_log) st.eos_log[(size_t)b * st.max_frames_cap + i] = eos;
        if (ctrl->eos_enabled && i + 1 >= ctrl->eos_min_frames && eos >= ctrl->eos_threshold && st.eos_step[b] < 0) st.eos_step[b] = i;
    }
}

// Flow-head entry for LSD step k: latent = noise (k == 0) and fx = input_proj . latent + b
// (ptts_flowlm.c:710). grid B, block 512.
__global__ void __launch_bounds__(512) k_flow_init(const FlowState st, int k, const __nv_bfloat16 *inp_w /*[512,32]*/,
                                                   const float *inp_b, float *fx) {
    pdl_wait(); pdl_trigger();
    __shared__ float lat[FL_LAT];
    const int b = blockIdx.x, i = *st.step;
    if (threadIdx.x < FL_LAT) {
        float v;
        if (k == 0) {
            const int fi = min(max(i, 0), st.max_frames_cap - 1);
            v = st.noise[((size_t)b * st.max_frames_cap + fi) * FL_LAT + threadIdx.x];
            st.latent[b * FL_LAT + threadIdx.x] = v;
        } else {
            v = st.latent[b * FL_LAT + threadIdx.x];
        }
        lat[threadIdx.x] = v;
    }
    __syncthreads();
    const int n = threadIdx.x;
    const uint4 *wr = (const uint4 *)(inp_w + (size_t)n * FL_LAT);
    float a = inp_b[n];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint4 u = wr[c];
        const float *l = lat + c * 8;
        a += bf16lo(u.x) * l[0] + bf16hi(u.x) * l[1] + bf16lo(u.y) * l[2] + bf16hi(u.y) * l[3] + bf16lo(u.z) * l[4] +
             bf16hi(u.z) * l[5] + bf16lo(u.w) * l[6] + bf16hi(u.w) * l[7];
    }
    fx[(size_t)b * FN_D + n] = a;
}

// Euler update latent += flow / N (ptts_flowlm.c:775-777); on the last LSD step record the frame,
// apply the stop rule (ptts_flowlm.c:1235-1238). grid B, block 32.
__global__ void __launch_bounds__(32) k_flow_update(const FlowState st, const FlowCtrl *ctrl, const int *eos_after, int k,
                                                    const float *flow /*[B,32]*/) {
    pdl_wait(); pdl_trigger();
    const int b = blockIdx.x, i = *st.step, d = threadIdx.x;
    const bool valid = i < st.max_frames[b] && !st.done[b];
    const float f = flow[b * FL_LAT + d];
    const float v = st.latent[b * FL_LAT + d] + f / (float)st.lsd_steps;
    st.latent[b * FL_LAT + d] = v;
    if (valid) {
        if (i == 0 && k == 0) st.first_flow[b * FL_LAT + d] = f;
        if (k == st.lsd_steps - 1) st.out_latents[((size_t)b * st.max_frames_cap + i) * FL_LAT + d] = v;
    }
    __syncwarp();
    if (k == st.lsd_steps - 1 && d == 0) {
        if (valid) {
            st.used[b] = i + 1;
            const int es = st.eos_step[b];
            if ((es >= 0 && i >= es + eos_after[b]) || i + 1 >= st.max_frames[b]) st.done[b] = 1;
        }
    }
}
__global__ void k_advance_step(int *step) {
    pdl_wait(); pdl_trigger(); *step += 1; }

// Batch-1 decode: the three single-CTA kernels at the start of a frame step as ONE graph node - out_norm (ptts_flowlm.c:1197),
// EOS logit + stop rule (:1202-1209), noise -> latent and the flow input projection of LSD step 0 (:1211-1231, 710).
// grid 1, block 256. x [1024] -> normed [1024] (input of the cond_embed GEMV), fx [512].
__global__ void __launch_bounds__(256) k_b1_head(const FlowState st, const FlowCtrl *ctrl, const float *x, const float *nw, const float *nb,
                                                 const float *eos_w, const float *eos_b, const __nv_bfloat16 *inp_w, const float *inp_b,
                                                 float *normed, float *fx) {
    pdl_wait(); pdl_trigger();
    __shared__ float red[32];
    __shared__ float lat[FL_LAT];
    const int tid = threadIdx.x, i = *st.step;
    const float4 v = *(const float4 *)(x + tid * 4);
    const float4 w4 = *(const float4 *)(nw + tid * 4), b4 = *(const float4 *)(nb + tid * 4), e4 = *(const float4 *)(eos_w + tid * 4);
    if (tid < FL_LAT) {   // noise of this frame -> running latent (LSD step 0)
        const int fi = min(max(i, 0), st.max_frames_cap - 1);
        const float z = st.noise[(size_t)fi * FL_LAT + tid];
        st.latent[tid] = z;
        lat[tid] = z;
    }
    const float mean = block_sum(v.x + v.y + v.z + v.w, red) / (float)FL_D;
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    const float inv = 1.0f / sqrtf(block_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3, red) / (float)FL_D + 1e-5f);
    const float4 y = make_float4(d0 * inv * w4.x + b4.x, d1 * inv * w4.y + b4.y, d2 * inv * w4.z + b4.z, d3 * inv * w4.w + b4.w);
    *(float4 *)(normed + tid * 4) = y;
    const float eos = block_sum(e4.x * y.x + e4.y * y.y + e4.z * y.z + e4.w * y.w, red) + eos_b[0];
    const bool valid = i < st.max_frames[0] && !st.done[0];
    if (valid) {
        if (i == 0) *(float4 *)(st.first_cond + tid * 4) = y;
        if (tid == 0) {
            if (i == 0) st.first_eos[0] = eos;
            if (st.eos_log) st.eos_log[i] = eos;
            if (ctrl->eos_enabled && i + 1 >= ctrl->eos_min_frames && eos >= ctrl->eos_threshold && st.eos_step[0] < 0) st.eos_step[0] = i;
        }
    }
    // (block_sum's barriers made lat[] visible)
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int n = tid + r * 256;
        const uint4 *wr = (const uint4 *)(inp_w + (size_t)n * FL_LAT);
        float a = inp_b[n];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint4 u = wr[c];
            const float *l = lat + c * 8;
            a += bf16lo(u.x) * l[0] + bf16hi(u.x) * l[1] + bf16lo(u.y) * l[2] + bf16hi(u.y) * l[3] + bf16lo(u.z) * l[4] +
                 bf16hi(u.z) * l[5] + bf16lo(u.w) * l[6] + bf16hi(u.w) * l[7];
        }
        fx[n] = a;
    }
}

// Batch-1 decode: the three kernels at the end of the flow head as ONE graph node - Euler update of the last LSD step + frame record +
// stop rule (ptts_flowlm.c:775-777, 1235-1238), x = input_linear . latent (:1240-1241). grid 1, block 256.
__global__ void __launch_bounds__(256) k_b1_tail(const FlowState st, const int *eos_after, const float *flow, const __nv_bfloat16 *in_lin, float *x) {
    pdl_wait(); pdl_trigger();
    __shared__ float lat[FL_LAT];
    const int tid = threadIdx.x, i = *st.step, k = st.lsd_steps - 1;
    if (tid < FL_LAT) {
        const bool valid = i < st.max_frames[0] && !st.done[0];
        const float f = flow[tid];
        const float v = st.latent[tid] + f / (float)st.lsd_steps;
        st.latent[tid] = v;
        lat[tid] = v;
        if (valid) {
            if (i == 0 && k == 0) st.first_flow[tid] = f;
            st.out_latents[(size_t)i * FL_LAT + tid] = v;
        }
        __syncwarp();
        if (tid == 0 && valid) {
            st.used[0] = i + 1;
            const int es = st.eos_step[0];
            if ((es >= 0 && i >= es + eos_after[0]) || i + 1 >= st.max_frames[0]) st.done[0] = 1;
        }
    }
    __syncthreads();
    for (int n = tid; n < FL_D; n += 256) {
        const uint4 *wr = (const uint4 *)(in_lin + (size_t)n * FL_LAT);
        float a = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint4 u = wr[c];
            const float *l = lat + c * 8;
            a += bf16lo(u.x) * l[0] + bf16hi(u.x) * l[1] + bf16lo(u.y) * l[2] + bf16hi(u.y) * l[3] + bf16lo(u.z) * l[4] +
                 bf16hi(u.z) * l[5] + bf16lo(u.w) * l[6] + bf16hi(u.w) * l[7];
        }
        x[n] = a;
    }
}

// =====================================================================================
// RoPE + KV append + single-query attention over [shared voice prefix | own positions]
// (ptts_flowlm.c:465-525).  grid (16 heads, B), block 128.
// cache layout: prefix [H][pstride][64] fp32, own [H][B][max_s][64] fp16 (measured effect on the latents: 2.7e-5
// max-relative after 48 frames against a 1e-3 bar, tools/parity_long.py).
// =====================================================================================
struct AttnDecodeParams {
    const float *qkv;          // [B, 3072], or split-K partial sums [qkv_nsplit][..] qkv_stride floats apart
    int qkv_nsplit;            // <= 1: plain
    size_t qkv_stride;
    const float *rope_cos, *rope_sin;  // [max_pos][32]
    const float *pk, *pv;      // prefix cache of this layer (NULL when cond_len == 0)
    const __half *pk16, *pv16; // fp16 copy of the prefix cache (tensor-core attention of the persistent kernel)
    const __half *nk, *nv;     // own caches of the layer whose attention phase comes next (L2 prefetch), n_extra rows beyond `step`
    int n_extra;
    // persistent kernel only: values fixed for the whole launch, filled in by the kernel itself
    int cond_len_val;          // *ctrl_cond_len
    const int *ob_cache;       // own_base[ob_first .. ob_first + ob_count) in shared memory (NULL: read own_base)
    int ob_first, ob_count;
    __half *ok, *ov;           // own cache of this layer (fp16), head-major [H][cap_b][max_s][64]: the utterances one CTA works on
                               // (one head, consecutive b) are one contiguous span of memory
    int cap_b;                 // utterance capacity of the own cache
    const int *ctrl_cond_len;  // device cond_len
    const int *own_base;       // [B] own-cache rows already filled by the prefill (tok_len + 1)
    int pstride;               // row capacity of the prefix buffers
    int max_s;
    const int *step;
    float *out_f32;            // [B,1024] (GEMV path)
    uint16_t *out_h, *out_l;   // bf16 hi/lo (tensor path)
};

// =====================================================================================
// Decode attention: one CTA = one head x G
// utterances, 8 warps, WPU = 8/G warps per utterance.
//   * the shared voice-prefix K / V tiles are staged ONCE per CTA in shared memory and reused
//     by all G utterances (the prefix is utterance-independent, ptts_flowlm.c:1159-1166);
//   * per-utterance keys are read straight from HBM with fully coalesced 256-byte rows
//     (lane = dimension pair), 8 rows in flight per warp;
//   * keys are dealt to the warps of an utterance in blocks of 32; partial (max, sum, acc)
//     triples are merged flash-decoding style.
// grid (16, ceil(B/G)), block 256. smem: tile[128][65] + scores[8][pcap] + sq[G][64] + comb[8][66].
// =====================================================================================
constexpr int AD_TILE = 128, AD_LD = 65;
__host__ __device__ inline size_t attn2_smem_bytes(int pcap, int G) {
    return (size_t)(AD_TILE * AD_LD + 8 * pcap + G * FL_HD + 8 * 66) * sizeof(float);
}

// Sum of up to 8 split-K partial values; the loads are independent (all in flight together).
__device__ __forceinline__ float2 ld_fold2(const float *ptr, int nsplit, size_t stride) {
    float2 a[8];
    a[0] = *(const float2 *)ptr;
#pragma unroll
    for (int z = 1; z < 8; z++) a[z] = z < nsplit ? *(const float2 *)(ptr + z * stride) : make_float2(0.0f, 0.0f);
#pragma unroll
    for (int z = 1; z < 8; z++) { a[0].x += a[z].x; a[0].y += a[z].y; }
    return a[0];
}

// One work item = (head h, utterance group gi) processed by 256 threads (tid 0..255); `sync` is the
// barrier of exactly those threads (the whole CTA in k_attn_decode2, a named barrier in the
// persistent step kernel); `step` is the decode iteration.
template <int WPU, typename Sync>
__device__ __forceinline__ void attn_decode_item(const AttnDecodeParams &p, int B, int pcap, int h, int gi, int tid, float *ad_smem,
                                                 int step, Sync sync) {
    constexpr int G = 8 / WPU;
    float *tile = ad_smem;                       // [128][65]
    float *scores = tile + AD_TILE * AD_LD;      // [8][pcap]
    float *sq = scores + 8 * pcap;               // [G][64]
    float *comb = sq + G * FL_HD;                // [8][66]
    const int w = tid >> 5, lane = tid & 31;
    const int g = w / WPU, wu = w % WPU;
    const int b = gi * G + g;
    const bool live = b < B;
    const int cond_len = p.ctrl_cond_len ? *p.ctrl_cond_len : 0;
    const int s = step + (live ? p.own_base[b] : 0);   // own-cache row written by this step
    const int pos = cond_len + s, P = pos + 1;
    __half *okb = p.ok + ((size_t)h * p.cap_b + (live ? b : 0)) * p.max_s * FL_HD;
    __half *ovb = p.ov + ((size_t)h * p.cap_b + (live ? b : 0)) * p.max_s * FL_HD;
    float *sc = scores + w * pcap;
    float2 q = make_float2(0.0f, 0.0f);
    if (live) {
        const float *row = p.qkv + (size_t)b * 3 * FL_D + h * FL_HD;
        const float c = p.rope_cos[pos * 32 + lane], sn = p.rope_sin[pos * 32 + lane];
        const float2 qq = ld_fold2(row + 2 * lane, p.qkv_nsplit, p.qkv_stride);
        // 1/sqrt(64) = 0.125 is a power of two: scaling q equals scaling every score
        q = make_float2((qq.x * c - qq.y * sn) * 0.125f, (qq.x * sn + qq.y * c) * 0.125f);
        if (wu == 0) {
            const float2 kk = ld_fold2(row + FL_D + 2 * lane, p.qkv_nsplit, p.qkv_stride);
            *(__half2 *)(okb + (size_t)s * FL_HD + 2 * lane) = __floats2half2_rn(kk.x * c - kk.y * sn, kk.x * sn + kk.y * c);
            const float2 vnew = ld_fold2(row + 2 * FL_D + 2 * lane, p.qkv_nsplit, p.qkv_stride);
            *(__half2 *)(ovb + (size_t)s * FL_HD + 2 * lane) = __floats2half2_rn(vnew.x, vnew.y);
            *(float2 *)(sq + g * FL_HD + 2 * lane) = q;
        }
    }
    sync();   // appended row + sq visible to every warp of the CTA
    const float *pkh = p.pk ? p.pk + (size_t)h * p.pstride * FL_HD : nullptr;
    const float *pvh = p.pv ? p.pv + (size_t)h * p.pstride * FL_HD : nullptr;
    const float *qs = sq + g * FL_HD;

    // ---- scores, prefix part: lane = key, K tile in shared memory (row stride 65: conflict-free)
    for (int t0 = 0; t0 < cond_len; t0 += AD_TILE) {
        const int nt = min(AD_TILE, cond_len - t0);
        for (int idx = tid; idx < nt * 16; idx += 256) {
            const int r = idx >> 4, c4 = (idx & 15) * 4;
            const float4 v = *(const float4 *)(pkh + (size_t)(t0 + r) * FL_HD + c4);
            float *d = tile + r * AD_LD + c4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        sync();
        if (live) {
            for (int rr = 0; rr < AD_TILE / 32; rr++) {
                const int blk = t0 / 32 + rr;
                if (blk % WPU != wu) continue;
                const int r = rr * 32 + lane;
                if (r < nt) {
                    const float *kr = tile + r * AD_LD;
                    float a = 0.0f;
#pragma unroll 16
                    for (int d = 0; d < FL_HD; d++) a += qs[d] * kr[d];
                    sc[t0 + r] = a;
                }
            }
        }
        sync();
    }
    // ---- scores, own part: lane = dimension pair, coalesced 256 B rows from HBM, 8 in flight
    float mx = -INFINITY;
    if (live) {
        for (int blk = cond_len / 32; blk * 32 < P; blk++) {
            if (blk % WPU != wu) continue;
            const int ta = max(blk * 32, cond_len), tb = min(blk * 32 + 32, P);
            for (int t = ta; t < tb; t += 8) {
                float2 kv[8];
#pragma unroll
                for (int j = 0; j < 8; j++)
                    kv[j] = (t + j < tb) ? __half22float2(*(const __half2 *)(okb + (size_t)(t + j - cond_len) * FL_HD + 2 * lane)) : make_float2(0.0f, 0.0f);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float a = warp_sum(q.x * kv[j].x + q.y * kv[j].y);
                    if (lane == 0 && t + j < tb) sc[t + j] = a;
                }
            }
        }
        __syncwarp();
        // ---- per-warp softmax partials over the blocks dealt to this warp
        for (int blk = wu; blk * 32 < P; blk += WPU) {
            const int t = blk * 32 + lane;
            if (t < P) mx = fmaxf(mx, sc[t]);
        }
        mx = warp_max(mx);
    }
    float sum = 0.0f;
    if (live) {
        for (int blk = wu; blk * 32 < P; blk += WPU) {
            const int t = blk * 32 + lane;
            if (t < P) { const float e = expf(sc[t] - mx); sc[t] = e; sum += e; }
        }
        sum = warp_sum(sum);
        __syncwarp();
    }
    // ---- P.V, prefix part from shared memory
    float2 acc = make_float2(0.0f, 0.0f);
    for (int t0 = 0; t0 < cond_len; t0 += AD_TILE) {
        const int nt = min(AD_TILE, cond_len - t0);
        for (int idx = tid; idx < nt * 16; idx += 256) {
            const int r = idx >> 4, c4 = (idx & 15) * 4;
            *(float4 *)(tile + r * FL_HD + c4) = *(const float4 *)(pvh + (size_t)(t0 + r) * FL_HD + c4);
        }
        sync();
        if (live) {
            for (int rr = 0; rr < AD_TILE / 32; rr++) {
                const int blk = t0 / 32 + rr;
                if (blk % WPU != wu) continue;
                const int r1 = min(rr * 32 + 32, nt);
#pragma unroll 4
                for (int r = rr * 32; r < r1; r++) {
                    const float pw = sc[t0 + r];
                    const float2 v = *(const float2 *)(tile + r * FL_HD + 2 * lane);
                    acc.x += pw * v.x; acc.y += pw * v.y;
                }
            }
        }
        sync();
    }
    // ---- P.V, own part from HBM
    if (live) {
        for (int blk = cond_len / 32; blk * 32 < P; blk++) {
            if (blk % WPU != wu) continue;
            const int ta = max(blk * 32, cond_len), tb = min(blk * 32 + 32, P);
            for (int t = ta; t < tb; t += 8) {
                float2 vv[8];
#pragma unroll
                for (int j = 0; j < 8; j++)
                    vv[j] = (t + j < tb) ? __half22float2(*(const __half2 *)(ovb + (size_t)(t + j - cond_len) * FL_HD + 2 * lane)) : make_float2(0.0f, 0.0f);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float pw = (t + j < tb) ? sc[t + j] : 0.0f;
                    acc.x += pw * vv[j].x; acc.y += pw * vv[j].y;
                }
            }
        }
    }
    // ---- merge the WPU partials of each utterance
    float *cw = comb + w * 66;
    if (lane == 0) { cw[64] = mx; cw[65] = sum; }
    *(float2 *)(cw + 2 * lane) = acc;
    sync();
    if (live && wu == 0) {
        float M = -INFINITY;
#pragma unroll
        for (int i = 0; i < WPU; i++) M = fmaxf(M, comb[(g * WPU + i) * 66 + 64]);
        float L = 0.0f;
        float2 o = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int i = 0; i < WPU; i++) {
            const float *ci = comb + (g * WPU + i) * 66;
            const float f = ci[64] == -INFINITY ? 0.0f : expf(ci[64] - M);
            L += ci[65] * f;
            const float2 a = *(const float2 *)(ci + 2 * lane);
            o.x += a.x * f; o.y += a.y * f;
        }
        o.x /= L; o.y /= L;
        const size_t oi = (size_t)b * FL_D + h * FL_HD + 2 * lane;
        if (p.out_f32) *(float2 *)(p.out_f32 + oi) = o;
        if (p.out_h) {
            const uint16_t h0 = f2h16(o.x, true), h1 = f2h16(o.y, true);
            *(uint32_t *)(p.out_h + oi) = (uint32_t)h0 | ((uint32_t)h1 << 16);
            if (p.out_l) {
                const uint16_t l0 = f2h16(o.x - h162f(h0, true), true), l1 = f2h16(o.y - h162f(h1, true), true);
                *(uint32_t *)(p.out_l + oi) = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
        }
    }
}

template <int WPU>
__global__ void __launch_bounds__(256) k_attn_decode2(const AttnDecodeParams p, int B, int pcap) {
    pdl_wait(); pdl_trigger();
    extern __shared__ __align__(16) float ad_smem_dyn[];
    attn_decode_item<WPU>(p, B, pcap, blockIdx.x, blockIdx.y, threadIdx.x, ad_smem_dyn, *p.step, [] { __syncthreads(); });
}

// =====================================================================================
// Batched prefill: all text tokens + BOS of all utterances (or all voice
// frames) are packed into M rows and go through each layer as one GEMM pass; this kernel set is
// the glue around those GEMMs.
// =====================================================================================
struct PrefillRows {
    const int *row_b;      // [M] utterance of each packed row
    const int *row_i;      // [M] own position of each packed row (0 .. tok_len[b])
    const int *row_start;  // [n+1] first packed row of each utterance
};

// X[r,:] = embedding of token i of utterance b, or input_linear.bos_emb for the last row
// (ptts_flowlm.c:1168-1181); ext (voice frames, ptts_flowlm.c:1159-1161) overrides. grid M, block 256.
__global__ void __launch_bounds__(256) k_pf_gather(const PrefillRows pr, const int *tokens, int max_tok, const int *tok_len,
                                                   const __nv_bfloat16 *embed, const float *bos_proj, const float *ext, float *X) {
    pdl_wait(); pdl_trigger();
    const int r = blockIdx.x;
    float *xr = X + (size_t)r * FL_D;
    if (ext) {
        for (int i = threadIdx.x; i < FL_D; i += 256) xr[i] = ext[(size_t)r * FL_D + i];
        return;
    }
    const int b = pr.row_b[r], i = pr.row_i[r];
    if (i < tok_len[b]) {
        int id = tokens[(size_t)b * max_tok + i];
        if (id < 0 || id > 4000) id = 0;  // ptts_flowlm.c:1170
        for (int k = threadIdx.x; k < FL_D; k += 256) xr[k] = __bfloat162float(embed[(size_t)id * FL_D + k]);
    } else {
        for (int k = threadIdx.x; k < FL_D; k += 256) xr[k] = bos_proj[k];
    }
}

// x[b,:] = X[last row of utterance b]. grid n, block 256.
__global__ void __launch_bounds__(256) k_pf_last(const PrefillRows pr, const float *X, float *x) {
    pdl_wait(); pdl_trigger();
    const int b = blockIdx.x, r = pr.row_start[b + 1] - 1;
    for (int k = threadIdx.x; k < FL_D; k += 256) x[(size_t)b * FL_D + k] = X[(size_t)r * FL_D + k];
}

// Causal attention for the packed prefill rows of one (head, utterance): RoPE at pos = cond_len + i,
// K/V appended to the own cache rows 0..nq-1, keys = shared voice prefix + own rows <= i
// (ptts_flowlm.c:465-525 applied to every prefill position at once). grid (16, n), block 256.
// smem: Kp[cond][65] Vp[cond][64] Ko[nq][65] Vo[nq][64] sc[8][cond+nq] q[8][64].
struct PrefillAttnParams {
    const float *qkv;          // [M, 3072]
    const float *rope_cos, *rope_sin;
    const float *pk, *pv;      // prefix cache of this layer (NULL when cond_len == 0)
    void *ok, *ov;             // destination rows [H][cap_b][max_s][64]: the fp16 own cache, or (voice pass) the fp32 prefix cache
    int own_f16;               // 1: ok / ov are __half, 0: float
    int cap_b;
    int cond_len, pstride, max_s, max_q;
    uint16_t *out_h, *out_l;   // [M,1024] bf16 hi/lo
};
__host__ __device__ inline size_t pf_attn_smem_bytes(int cond_len, int max_q) {
    const size_t rows = (size_t)cond_len + max_q, kfl = (rows * AD_LD + 3) & ~(size_t)3, scf = (rows + 3) & ~(size_t)3;
    return (kfl + rows * FL_HD + 8 * scf + 8 * FL_HD) * sizeof(float);
}
__global__ void __launch_bounds__(256) k_pf_attn(const PrefillAttnParams p, const PrefillRows pr) {
    pdl_wait(); pdl_trigger();
    extern __shared__ __align__(16) float pf_smem[];
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int r0 = pr.row_start[b], nq = pr.row_start[b + 1] - r0, cl = p.cond_len, P = cl + nq;
    const size_t nrows = (size_t)cl + p.max_q, scf = (nrows + 3) & ~(size_t)3;
    float *K = pf_smem;                                        // [rows][65]  prefix rows then own rows
    float *V = K + ((nrows * AD_LD + 3) & ~(size_t)3);         // [rows][64], 16-byte aligned
    float *sc = V + nrows * FL_HD;                             // [8][scf]
    float *qs = sc + 8 * scf;                                  // [8][64]
    const size_t obase = ((size_t)h * p.cap_b + b) * p.max_s * FL_HD;
    for (int idx = tid; idx < cl * 16; idx += 256) {
        const int r = idx >> 4, c4 = (idx & 15) * 4;
        const float4 kv = *(const float4 *)(p.pk + ((size_t)h * p.pstride + r) * FL_HD + c4);
        float *d = K + r * AD_LD + c4;
        d[0] = kv.x; d[1] = kv.y; d[2] = kv.z; d[3] = kv.w;
        *(float4 *)(V + r * FL_HD + c4) = *(const float4 *)(p.pv + ((size_t)h * p.pstride + r) * FL_HD + c4);
    }
    for (int idx = tid; idx < nq * 32; idx += 256) {
        const int i = idx >> 5, l = idx & 31, pos = cl + i;
        const float *row = p.qkv + (size_t)(r0 + i) * 3 * FL_D + h * FL_HD;
        const float c = p.rope_cos[pos * 32 + l], sn = p.rope_sin[pos * 32 + l];
        const float2 kk = *(const float2 *)(row + FL_D + 2 * l), vv = *(const float2 *)(row + 2 * FL_D + 2 * l);
        const float2 kr = make_float2(kk.x * c - kk.y * sn, kk.x * sn + kk.y * c);
        K[(cl + i) * AD_LD + 2 * l] = kr.x; K[(cl + i) * AD_LD + 2 * l + 1] = kr.y;
        *(float2 *)(V + (cl + i) * FL_HD + 2 * l) = vv;
        const size_t o = obase + (size_t)i * FL_HD + 2 * l;
        if (p.own_f16) {
            *(__half2 *)((__half *)p.ok + o) = __floats2half2_rn(kr.x, kr.y);
            *(__half2 *)((__half *)p.ov + o) = __floats2half2_rn(vv.x, vv.y);
        } else {
            *(float2 *)((float *)p.ok + o) = kr;
            *(float2 *)((float *)p.ov + o) = vv;
        }
    }
    __syncthreads();
    float *scw = sc + w * scf, *qw = qs + w * FL_HD;
    for (int i = w; i < nq; i += 8) {
        const int pos = cl + i, nk = pos + 1;
        const float *row = p.qkv + (size_t)(r0 + i) * 3 * FL_D + h * FL_HD;
        const float c = p.rope_cos[pos * 32 + lane], sn = p.rope_sin[pos * 32 + lane];
        const float2 qq = *(const float2 *)(row + 2 * lane);
        __syncwarp();
        qw[2 * lane] = (qq.x * c - qq.y * sn) * 0.125f;
        qw[2 * lane + 1] = (qq.x * sn + qq.y * c) * 0.125f;
        __syncwarp();
        float mx = -INFINITY;
        for (int t = lane; t < nk; t += 32) {
            const float *kr = K + t * AD_LD;
            float a = 0.0f;
#pragma unroll 16
            for (int d = 0; d < FL_HD; d++) a += qw[d] * kr[d];
            scw[t] = a;
            mx = fmaxf(mx, a);
        }
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int t = lane; t < nk; t += 32) { const float e = expf(scw[t] - mx); scw[t] = e; sum += e; }
        sum = warp_sum(sum);
        __syncwarp();
        float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll 4
        for (int t = 0; t < nk; t++) {
            const float pw = scw[t];
            const float2 v = *(const float2 *)(V + t * FL_HD + 2 * lane);
            acc.x += pw * v.x; acc.y += pw * v.y;
        }
        acc.x /= sum; acc.y /= sum;
        const size_t oi = (size_t)(r0 + i) * FL_D + h * FL_HD + 2 * lane;
        const uint16_t h0 = f2h16(acc.x, true), h1 = f2h16(acc.y, true);
        *(uint32_t *)(p.out_h + oi) = (uint32_t)h0 | ((uint32_t)h1 << 16);
        const uint16_t l0 = f2h16(acc.x - h162f(h0, true), true), l1 = f2h16(acc.y - h162f(h1, true), true);
        *(uint32_t *)(p.out_l + oi) = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
}

// cos/sin table: freq_i = exp(-ln(10000) * 2i/64) (ptts_flowlm.c:361-364), angle = freq_i * pos.
__global__ void k_rope_table(float *cs, float *sn, int max_pos) {
    pdl_wait(); pdl_trigger();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= max_pos * 32) return;
    const int pos = idx >> 5, i = idx & 31;
    const float f = expf(-logf(10000.0f) * (2.0f * (float)i / 64.0f));
    const float a = f * (float)pos;
    cs[idx] = cosf(a);
    sn[idx] = sinf(a);
}

// Time embedding MLP for one scalar (ptts_flowlm.c:572-586), fp32 weights, single block of 512.
// Accumulates 0.5 * embed into out (call for s and for t), out pre-initialised with cond_embed.bias.
__global__ void __launch_bounds__(512) k_time_embed(const float *l0w, const float *l0b, const float *l2w, const float *l2b,
                                                    const float *alpha, const float *freqs, float tval, float *out) {
    pdl_wait(); pdl_trigger();
    __shared__ float e[256], h[FN_D], red[32];
    const int n = threadIdx.x;
    if (n < 128) {
        const float f = freqs ? freqs[n] : expf(-logf(10000.0f) * ((float)n / 128.0f));
        e[n] = cosf(f * tval);
        e[n + 128] = sinf(f * tval);
    }
    __syncthreads();
    float a = l0b[n];
    for (int i = 0; i < 256; i++) a += l0w[n * 256 + i] * e[i];
    h[n] = a / (1.0f + expf(-a));
    __syncthreads();
    float y = l2b[n];
    for (int i = 0; i < FN_D; i++) y += l2w[n * FN_D + i] * h[i];
    // variance-style "RMSNorm": unbiased variance, x not centred (ptts_flowlm.c:183-197)
    const float mean = block_sum(y, red) / (float)FN_D;
    const float d = y - mean;
    const float var = block_sum(d * d, red) / (float)(FN_D - 1);
    out[n] += 0.5f * (y * alpha[n] * (1.0f / sqrtf(var + 1e-5f)));
}

// =====================================================================================
// Mimi front end: latent scaling (optional), quantizer.output_proj (ptts_mimi.c:577-586) and the
// depthwise ConvTranspose1d k=32 s=16 (ptts_mimi.c:588-593; two taps per output), written
// time-major: x[b, t*16+r, c].  grid (F, B), block 512 (thread = channel).
// =====================================================================================
__global__ void __launch_bounds__(512) k_mimi_front(const float *lat /*[B,F,32], item stride lat_stride*/, size_t lat_stride, int F,
                                                    const float *emb_std, const float *emb_mean,
                                                    int apply_scale, const float *quant_wt /*[32,512] transposed*/,
                                                    const float *ups_t /*[32,512] transposed*/, float *x /*[B,16F,512]*/, int prev_ok) {
    pdl_wait(); pdl_trigger();
    __shared__ float l0[32], l1[32];
    const int t = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
    if (c < 32) {
        float v = lat[(size_t)b * lat_stride + (size_t)t * 32 + c];
        if (apply_scale) v = v * emb_std[c] + emb_mean[c];
        l0[c] = v;
    } else if (c < 64) {
        float v = 0.0f;
        if (t > 0 || prev_ok) {   // prev_ok: `lat` points into the middle of an utterance (streaming), frame -1 exists
            v = lat[(size_t)b * lat_stride + (size_t)t * 32 - 32 + (c - 32)];
            if (apply_scale) v = v * emb_std[c - 32] + emb_mean[c - 32];
        }
        l1[c - 32] = v;
    }
    __syncthreads();
    float q0 = 0.0f, q1 = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const float w = quant_wt[i * MI_D + c];
        q0 += w * l0[i];
        q1 += w * l1[i];
    }
    float *xo = x + ((size_t)b * F + t) * 16 * MI_D + c;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        float v = ups_t[r * MI_D + c] * q0;
        if (t > 0 || prev_ok) v += ups_t[(16 + r) * MI_D + c] * q1;
        xo[(size_t)r * MI_D] = v;
    }
}

// Warp-per-row LayerNorm for 512-wide rows (Mimi transformer norms, ptts_mimi.c:184-206): two-pass in
// registers, shuffle reductions only, fp16 output. grid ceil(R/8), block 256.
__global__ void __launch_bounds__(256) k_ln512_f16(const float *x, size_t R, const float *w, const float *b, float eps, __half *out) {
    pdl_wait(); pdl_trigger();
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= R) return;
    const int lane = threadIdx.x & 31;
    const float4 *xr = (const float4 *)(x + row * MI_D);
    float4 v[4];
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) { v[j] = xr[lane + 32 * j]; s += v[j].x + v[j].y + v[j].z + v[j].w; }
    const float mean = warp_sum(s) / (float)MI_D;
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float a = v[j].x - mean, c = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
        q += a * a + c * c + d * d + e * e;
    }
    const float inv = 1.0f / sqrtf(warp_sum(q) / (float)MI_D + eps);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = (lane + 32 * j) * 4;
        const float4 ww = *(const float4 *)(w + i), bb = *(const float4 *)(b + i);
        const __half2 h0 = __floats2half2_rn((v[j].x - mean) * inv * ww.x + bb.x, (v[j].y - mean) * inv * ww.y + bb.y);
        const __half2 h1 = __floats2half2_rn((v[j].z - mean) * inv * ww.z + bb.z, (v[j].w - mean) * inv * ww.w + bb.w);
        uint2 pk;
        pk.x = *(const uint32_t *)&h0; pk.y = *(const uint32_t *)&h1;
        *(uint2 *)(out + row * MI_D + i) = pk;
    }
}

// Mimi: split qkv rows [R,1536] fp32 into per-head q (scaled by 1/8), k (both RoPE-rotated at
// pos = t) and v, stored fp16 [B][H][T][64] (ptts_mimi.c:329-341). grid R, block 256.
__global__ void __launch_bounds__(256) k_mimi_rope_heads(const float *qkv, int T, const float *rope_cos, const float *rope_sin,
                                                         __half *q, __half *k, __half *v) {
    pdl_wait(); pdl_trigger();
    const size_t row = blockIdx.x;
    const int b = (int)(row / T), t = (int)(row % T);
    const float *r = qkv + row * 3 * MI_D;
    for (int pi = threadIdx.x; pi < 3 * MI_D / 2; pi += 256) {
        const int which = pi / (MI_D / 2), pr = pi % (MI_D / 2);
        const int h = pr / 32, i = pr % 32;
        const float a0 = r[which * MI_D + 2 * pr], a1 = r[which * MI_D + 2 * pr + 1];
        float o0 = a0, o1 = a1;
        if (which < 2) {
            const float c = rope_cos[t * 32 + i], s = rope_sin[t * 32 + i];
            o0 = a0 * c - a1 * s;
            o1 = a0 * s + a1 * c;
            if (which == 0) { o0 *= 0.125f; o1 *= 0.125f; }
        }
        __half *dst = (which == 0 ? q : which == 1 ? k : v) + (((size_t)b * MI_H + h) * T + t) * 64 + 2 * i;
        *(__half2 *)dst = __floats2half2_rn(o0, o1);
    }
}

// Mimi banded causal attention: key tk visible to query tq iff 0 <= tq - tk < 250
// (ptts_mimi.c:264-299; the reference is an O(T^2) masked loop, this is O(T*250)).
// Flash-style: one CTA = 64 consecutive queries of one (batch, head), 4 warps x 16 query rows;
// key/value blocks of 64 rows are staged in shared memory, S = Q.K^T and O += P.V run on
// mma.sync.m16n8k16 (fp16 in, fp32 accumulate) with an online softmax in registers. The attention
// is 3 % of Mimi's FLOPs (16.4 of 541.5 MFLOP per frame), so the warp-level MMA path is enough to
// take it off the profile; the GEMM-shaped 97 % run on tcgen05.
// q is pre-scaled by 1/8 and q, k are RoPE-rotated by k_mimi_rope_heads. grid (ceil(T/64), H, B), block 128.
constexpr int FA_LD = 72;   // padded row stride in halves: conflict-free 32-bit fragment loads
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *(const uint32_t *)&h;
}
// Four 8x8 b16 matrices from shared memory (lane i supplies the address of row i % 8 of matrix i / 8): register j of lane l holds
// M_j[l / 4][2 (l % 4) .. +1] - the B-fragment layout of mma.m16n8k16 for a [n][k] operand; .trans gives M_j[2 (l % 4) .. +1][l / 4],
// the layout for a [k][n] operand.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
// Streaming form: queries are rows [t_begin, T) of buffers that hold t_stride rows per head (the K / V caches of the utterance so
// far); out row 0 = query t_begin. One-shot form: t_begin = 0, t_stride = T.
__global__ void __launch_bounds__(128) k_mimi_win_attn(const __half *q, const __half *k, const __half *v, int T,
                                                       __half *out /*[B*(T - t_begin), 512]*/, int t_begin, int t_stride) {
    pdl_wait(); pdl_trigger();
    __shared__ __align__(16) __half sQ[64 * FA_LD], sKb[2][64 * FA_LD], sVb[2][64 * FA_LD];   // K / V blocks double-buffered (cp.async)
    const int q0 = t_begin + blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const size_t base = ((size_t)b * MI_H + h) * t_stride;
    const int kb_end = min(T, q0 + 64);   // keys beyond the tile's last query are never visible
    const int kb_start = max(0, q0 - (MI_CTX - 1)) & ~63;
    // K / V block kb -> buffer `buf` with cp.async (rows past T are zero-filled); one commit group per block
    auto stage = [&](int kb, int buf) {
        const uint32_t ka = smem_u32(sKb[buf]), va = smem_u32(sVb[buf]);
        for (int idx = tid; idx < 512; idx += 128) {
            const int r = idx >> 3, c = (idx & 7) * 8;
            const bool ok = kb + r < T;
            const size_t so = ok ? (base + kb + r) * 64 + c : base * 64;
            const uint32_t off = (uint32_t)((r * FA_LD + c) * 2), sz = ok ? 16u : 0u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ka + off), "l"(k + so), "r"(sz) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(va + off), "l"(v + so), "r"(sz) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(kb_start, 0);
    for (int idx = tid; idx < 512; idx += 128) {
        const int r = idx >> 3, c = (idx & 7) * 8;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (q0 + r < T) val = *(const uint4 *)(q + (base + q0 + r) * 64 + c);
        *(uint4 *)(sQ + r * FA_LD + c) = val;
    }
    __syncthreads();
    uint32_t qa[4][4];
    {
        const __half *r0 = sQ + (warp * 16 + g) * FA_LD, *r1 = r0 + 8 * FA_LD;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            qa[ks][0] = *(const uint32_t *)(r0 + ks * 16 + 2 * t);
            qa[ks][1] = *(const uint32_t *)(r1 + ks * 16 + 2 * t);
            qa[ks][2] = *(const uint32_t *)(r0 + ks * 16 + 8 + 2 * t);
            qa[ks][3] = *(const uint32_t *)(r1 + ks * 16 + 8 + 2 * t);
        }
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    const int tq0 = q0 + warp * 16 + g, tq1 = tq0 + 8;
    int buf = 0;
    for (int kb = kb_start; kb < kb_end; kb += 64, buf ^= 1) {
        const bool more = kb + 64 < kb_end;
        if (more) stage(kb + 64, buf ^ 1);          // the next block lands while this one is multiplied
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // fragment addresses for ldmatrix: K rows nt*8 + (lane & 7), 8-column block lane >> 3 (+ 4 for the second half of d);
        // V rows ks*16 + ((lane >> 3) & 1) * 8 + (lane & 7), 8-column block 2*ndp + (lane >> 4)
        const uint32_t sK_s = smem_u32(sKb[buf]) + (uint32_t)(((lane & 7) * FA_LD + (lane >> 3) * 8) * 2);
        const uint32_t sV_s = smem_u32(sVb[buf]) + (uint32_t)(((((lane >> 3) & 1) * 8 + (lane & 7)) * FA_LD + (lane >> 4) * 8) * 2);
        float sc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.0f;
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                uint32_t kf[4];
                ldsm_x4(kf, sK_s + (uint32_t)((nt * 8 * FA_LD + hf * 32) * 2));
                mma_f16_16816(sc[nt], qa[2 * hf], kf[0], kf[1]);
                mma_f16_16816(sc[nt], qa[2 * hf + 1], kf[2], kf[3]);
            }
        }
        // blocks wholly inside every query's window need no mask (block-uniform test)
        const bool need_mask = !(kb + 63 <= q0 && q0 + 63 - kb < MI_CTX);
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            if (need_mask) {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int key = kb + nt * 8 + 2 * t + (e & 1);
                    const int tq = e < 2 ? tq0 : tq1;
                    const bool ok = key <= tq && tq - key < MI_CTX;
                    if (!ok) sc[nt][e] = -INFINITY;
                }
            }
            bm0 = fmaxf(bm0, fmaxf(sc[nt][0], sc[nt][1]));
            bm1 = fmaxf(bm1, fmaxf(sc[nt][2], sc[nt][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
        // a row whose keys so far are all masked keeps m = -inf: scale 1, p = 0 (no NaN from inf - inf)
        const float f0 = mn0 == -INFINITY ? 1.0f : __expf(m0 - mn0), f1 = mn1 == -INFINITY ? 1.0f : __expf(m1 - mn1);
        const float sub0 = mn0 == -INFINITY ? 0.0f : mn0, sub1 = mn1 == -INFINITY ? 0.0f : mn1;
        float rs0 = 0.0f, rs1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            sc[nt][0] = __expf(sc[nt][0] - sub0); sc[nt][1] = __expf(sc[nt][1] - sub0);
            sc[nt][2] = __expf(sc[nt][2] - sub1); sc[nt][3] = __expf(sc[nt][3] - sub1);
            rs0 += sc[nt][0] + sc[nt][1];
            rs1 += sc[nt][2] + sc[nt][3];
            o[nt][0] *= f0; o[nt][1] *= f0; o[nt][2] *= f1; o[nt][3] *= f1;
        }
        l0 = l0 * f0 + rs0; l1 = l1 * f1 + rs1;
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            uint32_t pa[4];
            pa[0] = pack_h2(sc[2 * ks][0], sc[2 * ks][1]);
            pa[1] = pack_h2(sc[2 * ks][2], sc[2 * ks][3]);
            pa[2] = pack_h2(sc[2 * ks + 1][0], sc[2 * ks + 1][1]);
            pa[3] = pack_h2(sc[2 * ks + 1][2], sc[2 * ks + 1][3]);
#pragma unroll
            for (int ndp = 0; ndp < 4; ndp++) {
                uint32_t vf[4];
                ldsm_x4_t(vf, sV_s + (uint32_t)((ks * 16 * FA_LD + ndp * 16) * 2));
                mma_f16_16816(o[2 * ndp], pa, vf[0], vf[1]);
                mma_f16_16816(o[2 * ndp + 1], pa, vf[2], vf[3]);
            }
        }
        __syncthreads();   // all warps are done with this buffer before the stage of the next iteration overwrites it
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.0f ? 1.0f / l0 : 0.0f, i1 = l1 > 0.0f ? 1.0f / l1 : 0.0f;
#pragma unroll
    for (int nd = 0; nd < 8; nd++) {
        if (tq0 < T) *(__half2 *)(out + ((size_t)b * (T - t_begin) + tq0 - t_begin) * MI_D + h * 64 + nd * 8 + 2 * t) = __floats2half2_rn(o[nd][0] * i0, o[nd][1] * i0);
        if (tq1 < T) *(__half2 *)(out + ((size_t)b * (T - t_begin) + tq1 - t_begin) * MI_D + h * 64 + nd * 8 + 2 * t) = __floats2half2_rn(o[nd][2] * i1, o[nd][3] * i1);
    }
}

// =====================================================================================
// Prompt prefill attention on warp-level tensor cores (same contract as k_pf_attn, ptts_flowlm.c:465-525 applied to
// every prompt position at once): one CTA = up to 64 prompt rows of one (head, utterance), 4 warps x 16 query rows,
// flash-style over key blocks of 64 = [voice prefix (fp16 copy) | the utterance's own rows]. The own K / V rows are
// produced on the fly from the QKV projection (RoPE at pos = cond_len + i), rounded to fp16 - exactly what lands in the
// own cache and what every later decode step reads - and appended to the cache by the CTA whose query tile holds the row.
// Queries and probabilities enter the MMAs as fp16 hi + lo pairs (two MMAs each), accumulation and softmax in fp32.
// grid (ceil(max_q / 64), 16, n), block 128.
// =====================================================================================
__device__ __forceinline__ void split_h2(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *(const uint32_t *)&h; lo = *(const uint32_t *)&l;
}
__global__ void __launch_bounds__(128) k_pf_attn_tc(const PrefillAttnParams p, const PrefillRows pr, const __half *pk16, const __half *pv16) {
    pdl_wait(); pdl_trigger();
    __shared__ __align__(16) __half sQh[64 * FA_LD], sQl[64 * FA_LD], sK[64 * FA_LD], sV[64 * FA_LD];
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int r0 = pr.row_start[b], nq = pr.row_start[b + 1] - r0, cl = p.cond_len, i0 = qt * 64;
    if (i0 >= nq) return;
    const size_t obase = ((size_t)h * p.cap_b + b) * p.max_s * FL_HD;
    for (int idx = tid; idx < 64 * 32; idx += 128) {
        const int r = idx >> 5, l = idx & 31, i = i0 + r;
        float2 q = make_float2(0.0f, 0.0f);
        if (i < nq) {
            const float *row = p.qkv + (size_t)(r0 + i) * 3 * FL_D + h * FL_HD;
            const float c = p.rope_cos[(cl + i) * 32 + l], sn = p.rope_sin[(cl + i) * 32 + l];
            const float2 qq = *(const float2 *)(row + 2 * l);
            q = make_float2((qq.x * c - qq.y * sn) * 0.125f, (qq.x * sn + qq.y * c) * 0.125f);
        }
        uint32_t hi, lo;
        split_h2(q.x, q.y, hi, lo);
        *(uint32_t *)(sQh + r * FA_LD + 2 * l) = hi;
        *(uint32_t *)(sQl + r * FA_LD + 2 * l) = lo;
    }
    __syncthreads();
    uint32_t qh[4][4], ql[4][4];
    {
        const int ro = (warp * 16 + g) * FA_LD;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            qh[ks][0] = *(const uint32_t *)(sQh + ro + ks * 16 + 2 * t);
            qh[ks][1] = *(const uint32_t *)(sQh + ro + 8 * FA_LD + ks * 16 + 2 * t);
            qh[ks][2] = *(const uint32_t *)(sQh + ro + ks * 16 + 8 + 2 * t);
            qh[ks][3] = *(const uint32_t *)(sQh + ro + 8 * FA_LD + ks * 16 + 8 + 2 * t);
            ql[ks][0] = *(const uint32_t *)(sQl + ro + ks * 16 + 2 * t);
            ql[ks][1] = *(const uint32_t *)(sQl + ro + 8 * FA_LD + ks * 16 + 2 * t);
            ql[ks][2] = *(const uint32_t *)(sQl + ro + ks * 16 + 8 + 2 * t);
            ql[ks][3] = *(const uint32_t *)(sQl + ro + 8 * FA_LD + ks * 16 + 8 + 2 * t);
        }
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    const int tq0 = i0 + warp * 16 + g, tq1 = tq0 + 8;          // prompt rows of this lane's two accumulator rows
    const bool warp_live = i0 + warp * 16 < nq;
    const int kend = cl + min(nq, i0 + 64);                     // keys [0, kend) can be visible to this tile
    for (int kb = 0; kb < kend; kb += 64) {
        __syncthreads();   // all warps are done with the previous K/V block
        for (int idx = tid; idx < 512; idx += 128) {
            const int r = idx >> 3, c = (idx & 7) * 8, kk = kb + r;
            uint4 kv = make_uint4(0, 0, 0, 0), vv = kv;
            if (kk < cl) {
                kv = *(const uint4 *)(pk16 + ((size_t)h * p.pstride + kk) * FL_HD + c);
                vv = *(const uint4 *)(pv16 + ((size_t)h * p.pstride + kk) * FL_HD + c);
            } else if (kk - cl < nq) {
                const int i = kk - cl;
                const float *row = p.qkv + (size_t)(r0 + i) * 3 * FL_D + h * FL_HD + c;
                const float4 k0 = *(const float4 *)(row + FL_D), k1 = *(const float4 *)(row + FL_D + 4);
                const float4 v0 = *(const float4 *)(row + 2 * FL_D), v1 = *(const float4 *)(row + 2 * FL_D + 4);
                const float4 cs = *(const float4 *)(p.rope_cos + (cl + i) * 32 + (c >> 1)), sn = *(const float4 *)(p.rope_sin + (cl + i) * 32 + (c >> 1));
                const float kr[8] = {k0.x * cs.x - k0.y * sn.x, k0.x * sn.x + k0.y * cs.x, k0.z * cs.y - k0.w * sn.y, k0.z * sn.y + k0.w * cs.y,
                                     k1.x * cs.z - k1.y * sn.z, k1.x * sn.z + k1.y * cs.z, k1.z * cs.w - k1.w * sn.w, k1.z * sn.w + k1.w * cs.w};
                kv = make_uint4(pack_h2(kr[0], kr[1]), pack_h2(kr[2], kr[3]), pack_h2(kr[4], kr[5]), pack_h2(kr[6], kr[7]));
                vv = make_uint4(pack_h2(v0.x, v0.y), pack_h2(v0.z, v0.w), pack_h2(v1.x, v1.y), pack_h2(v1.z, v1.w));
                if ((i >> 6) == qt) {   // this tile owns the row: append it to the cache
                    const size_t o_ = obase + (size_t)i * FL_HD + c;
                    if (p.own_f16) {
                        *(uint4 *)((__half *)p.ok + o_) = kv;
                        *(uint4 *)((__half *)p.ov + o_) = vv;
                    } else {
                        *(float4 *)((float *)p.ok + o_) = make_float4(kr[0], kr[1], kr[2], kr[3]);
                        *(float4 *)((float *)p.ok + o_ + 4) = make_float4(kr[4], kr[5], kr[6], kr[7]);
                        *(float4 *)((float *)p.ov + o_) = v0;
                        *(float4 *)((float *)p.ov + o_ + 4) = v1;
                    }
                }
            }
            *(uint4 *)(sK + r * FA_LD + c) = kv;
            *(uint4 *)(sV + r * FA_LD + c) = vv;
        }
        __syncthreads();
        if (!warp_live) continue;   // padding rows only: nothing to compute (the barriers above are still reached)
        float sc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.0f;
            const __half *kr = sK + (nt * 8 + g) * FA_LD + 2 * t;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const uint32_t b0 = *(const uint32_t *)(kr + ks * 16), b1 = *(const uint32_t *)(kr + ks * 16 + 8);
                mma_f16_16816(sc[nt], qh[ks], b0, b1);
                mma_f16_16816(sc[nt], ql[ks], b0, b1);
            }
        }
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int key = kb + nt * 8 + 2 * t + (e & 1);
                if (key > cl + (e < 2 ? tq0 : tq1)) sc[nt][e] = -INFINITY;   // causal: prefix + own rows <= i
            }
            bm0 = fmaxf(bm0, fmaxf(sc[nt][0], sc[nt][1]));
            bm1 = fmaxf(bm1, fmaxf(sc[nt][2], sc[nt][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
        const float f0 = mn0 == -INFINITY ? 1.0f : expf(m0 - mn0), f1 = mn1 == -INFINITY ? 1.0f : expf(m1 - mn1);
        const float sub0 = mn0 == -INFINITY ? 0.0f : mn0, sub1 = mn1 == -INFINITY ? 0.0f : mn1;
        float rs0 = 0.0f, rs1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            sc[nt][0] = expf(sc[nt][0] - sub0); sc[nt][1] = expf(sc[nt][1] - sub0);
            sc[nt][2] = expf(sc[nt][2] - sub1); sc[nt][3] = expf(sc[nt][3] - sub1);
            rs0 += sc[nt][0] + sc[nt][1];
            rs1 += sc[nt][2] + sc[nt][3];
            o[nt][0] *= f0; o[nt][1] *= f0; o[nt][2] *= f1; o[nt][3] *= f1;
        }
        l0 = l0 * f0 + rs0; l1 = l1 * f1 + rs1;
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            uint32_t ph[4], pl[4];
            split_h2(sc[2 * ks][0], sc[2 * ks][1], ph[0], pl[0]);
            split_h2(sc[2 * ks][2], sc[2 * ks][3], ph[1], pl[1]);
            split_h2(sc[2 * ks + 1][0], sc[2 * ks + 1][1], ph[2], pl[2]);
            split_h2(sc[2 * ks + 1][2], sc[2 * ks + 1][3], ph[3], pl[3]);
            const __half *v0 = sV + (ks * 16 + 2 * t) * FA_LD + g, *v1 = v0 + 8 * FA_LD;
#pragma unroll
            for (int nd = 0; nd < 8; nd++) {
                const uint32_t b0 = (uint32_t)__half_as_ushort(v0[nd * 8]) | ((uint32_t)__half_as_ushort(v0[FA_LD + nd * 8]) << 16);
                const uint32_t b1 = (uint32_t)__half_as_ushort(v1[nd * 8]) | ((uint32_t)__half_as_ushort(v1[FA_LD + nd * 8]) << 16);
                mma_f16_16816(o[nd], ph, b0, b1);
                mma_f16_16816(o[nd], pl, b0, b1);
            }
        }
    }
    if (!warp_live) return;
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0_ = 1.0f / l0, i1_ = 1.0f / l1;
#pragma unroll
    for (int nd = 0; nd < 8; nd++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int tq = e ? tq1 : tq0;
            if (tq < nq) {
                const float a = o[nd][2 * e] * (e ? i1_ : i0_), c = o[nd][2 * e + 1] * (e ? i1_ : i0_);
                const size_t oi = (size_t)(r0 + tq) * FL_D + h * FL_HD + nd * 8 + 2 * t;
                const uint16_t h0 = f2h16(a, true), h1 = f2h16(c, true);
                *(uint32_t *)(p.out_h + oi) = (uint32_t)h0 | ((uint32_t)h1 << 16);
                const uint16_t l0_ = f2h16(a - h162f(h0, true), true), l1_ = f2h16(c - h162f(h1, true), true);
                *(uint32_t *)(p.out_l + oi) = (uint32_t)l0_ | ((uint32_t)l1_ << 16);
            }
        }
    }
}

// Final Conv1d 64 -> 1, k = 3 on ELU'd fp16 activations [B, T, 64] -> pcm [B, T] fp32
// (ptts_mimi.c:723-726 with ptts_kernels.c:128-153). 256 samples per block; the 258 input rows are
// staged once in shared memory (coalesced), then one thread per sample.
__global__ void __launch_bounds__(256) k_dec_out(const __half *e, int T, const float *w /*[64*3] as [c][kk]*/, const float *bias,
                                                 float *pcm) {
    pdl_wait(); pdl_trigger();
    __shared__ float sw[3][64];
    __shared__ __align__(16) __half se[258 * FA_LD];
    if (threadIdx.x < 192) sw[threadIdx.x % 3][threadIdx.x / 3] = w[threadIdx.x];
    const int b = blockIdx.y, t0 = blockIdx.x * 256;
    for (int idx = threadIdx.x; idx < 258 * 8; idx += 256) {
        const int r = idx >> 3, c = (idx & 7) * 8, tt = t0 - 2 + r;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (tt >= 0 && tt < T) val = *(const uint4 *)(e + ((size_t)b * T + tt) * 64 + c);
        *(uint4 *)(se + r * FA_LD + c) = val;
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= T) return;
    float a = bias[0];
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
        const uint4 *row = (const uint4 *)(se + (threadIdx.x + kk) * FA_LD);
#pragma unroll
        for (int c8 = 0; c8 < 8; c8++) {
            const uint4 u = row[c8];
            const __half2 *hp = (const __half2 *)&u;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __half22float2(hp[j]);
                a += f.x * sw[kk][c8 * 8 + 2 * j] + f.y * sw[kk][c8 * 8 + 2 * j + 1];
            }
        }
    }
    pcm[(size_t)b * T + t] = a;
}

// ---------------------------------------------------------------- ptts_mimi_forward_one (ptts_mimi.c:550-564)
// The reference's --mimi-test parity hook (main.c:444): quantizer projection of ONE latent, then the decoder transformer over that
// single token (no upsampling). With one token the softmax is over one key, so the attention output is the value row itself and
// RoPE at position 0 is the identity: each layer is LN -> v = in_proj[1024:1536] -> out_proj -> layer-scale residual -> LN -> MLP.
// One CTA of 512 threads; a debug entry, not a hot path.
struct MimiOneLayer { const __half *inproj, *outproj, *l1, *l2; const float *n1w, *n1b, *n2w, *n2b, *ls1, *ls2; };
struct MimiOneParams { const float *quant_w /*[32][512]*/; MimiOneLayer L[MI_L]; };
__device__ __forceinline__ float dot_h16_row(const __half *w, const float *x, int K) {
    float a = 0.0f;
    for (int k = 0; k < K; k += 8) {
        const uint4 u = *(const uint4 *)(w + k);
        const float2 p0 = __half22float2(*(const __half2 *)&u.x), p1 = __half22float2(*(const __half2 *)&u.y),
                     p2 = __half22float2(*(const __half2 *)&u.z), p3 = __half22float2(*(const __half2 *)&u.w);
        a += p0.x * x[k] + p0.y * x[k + 1] + p1.x * x[k + 2] + p1.y * x[k + 3] + p2.x * x[k + 4] + p2.y * x[k + 5] + p3.x * x[k + 6] + p3.y * x[k + 7];
    }
    return a;
}
__global__ void __launch_bounds__(512) k_mimi_forward_one(const MimiOneParams p, const float *lat, float *out) {
    __shared__ float x[MI_D], xn[MI_D], hb[MI_FF], red[32], l32[32];
    const int c = threadIdx.x;
    if (c < 32) l32[c] = lat[c];
    __syncthreads();
    {
        float a = 0.0f;
        for (int i = 0; i < 32; i++) a += p.quant_w[i * MI_D + c] * l32[i];
        x[c] = a;
    }
    __syncthreads();
    for (int l = 0; l < MI_L; l++) {
        const MimiOneLayer &L = p.L[l];
        for (int half = 0; half < 2; half++) {
            // LayerNorm (biased variance, eps 1e-5: ptts_mimi.c:184-206)
            const float mean = block_sum(x[c], red) / MI_D;
            const float d = x[c] - mean;
            const float var = block_sum(d * d, red) / MI_D;
            xn[c] = d * rsqrtf(var + 1e-5f) * (half ? L.n2w[c] : L.n1w[c]) + (half ? L.n2b[c] : L.n1b[c]);
            __syncthreads();
            if (half == 0) {
                const float v = dot_h16_row(L.inproj + (size_t)(2 * MI_D + c) * MI_D, xn, MI_D);
                __syncthreads();
                xn[c] = v;
                __syncthreads();
                x[c] += L.ls1[c] * dot_h16_row(L.outproj + (size_t)c * MI_D, xn, MI_D);
            } else {
                for (int j = c; j < MI_FF; j += 512) {
                    const float h = dot_h16_row(L.l1 + (size_t)j * MI_D, xn, MI_D);
                    hb[j] = 0.5f * h * (1.0f + tanhf(0.79788456f * (h + 0.044715f * h * h * h)));
                }
                __syncthreads();
                x[c] += L.ls2[c] * dot_h16_row(L.l2 + (size_t)c * MI_FF, hb, MI_FF);
            }
            __syncthreads();
        }
    }
    out[c] = x[c];
}

// fp32 -> fp16 copy of the voice prefix K / V (once per voice, ptts_b200.cu set_voice): 8 elements per thread and iteration
__global__ void __launch_bounds__(256) k_f32_to_f16(const float *src, __half *dst, size_t n8) {
    pdl_wait(); pdl_trigger();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = *(const float4 *)(src + i * 8), b = *(const float4 *)(src + i * 8 + 4);
        __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y), __floats2half2_rn(b.z, b.w)};
        *(uint4 *)(dst + i * 8) = *(const uint4 *)h;
    }
}

__global__ void k_f32_to_h16(const float *in, uint16_t *out, size_t n, int bf) {
    pdl_wait(); pdl_trigger();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f2h16(in[i], bf != 0);
}

}  // namespace ptts
