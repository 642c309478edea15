// enc.cuh - voice-cloning encoder: 24 kHz mono PCM -> FlowLM conditioning rows [frames, 1024].
// The reference only shape-checks these tensors (ptts.c:914-944: SEANet encoder, encoder_transformer, downsample.conv) and
// loads speaker_proj_weight without using it (ptts_flowlm.c:874); the forward pass is the mirror image of the decoder it does
// implement, built from the same primitives (causal Conv1d ptts_kernels.c:114-156, ResBlock ptts_mimi.c:160-177, transformer
// ptts_mimi.c:301-376) and checked against oracle/ptts_oracle.c orc_voice_encode.
//
// Every contraction reuses the decoder's machinery: a causal Conv1d with kernel 2s and stride s reads x[(t-1)s .. (t+1)s) for
// output t, i.e. rows t-1 and t of the SAME activation buffer viewed as [T/s][s*Cin] - a two-tap implicit GEMM with K = s*Cin per
// tap (the ConvTranspose trick of mimi.cuh, transposed); the residual blocks are the fused resblock_kernel; the transformer is
// mimi_transformer_layers with the encoder's weights. fp16 operands, fp32 accumulation, like the decoder.
#pragma once
#include "mimi.cuh"

namespace ptts {

// Conv1d weight [out, in, 2s] (stride s) -> B matrix [out][2 taps][s*in]: column tap*(s*in) + phase*in + ic = w[o][ic][tap*s + phase]
static int enc_load_strided(Engine *e, const HostWeights &hw, const char *name, ConvW &c, int cin, int cout, int stride, bool bias) {
    std::vector<float> w, b;
    char nm[256];
    const int k = 2 * stride;
    snprintf(nm, sizeof nm, "%s.weight", name);
    if (hw.get(nm, MIMI_PFX, (int64_t)cin * cout * k, w) != 0) return -1;
    b.assign(cout, 0.0f);
    if (bias) {
        snprintf(nm, sizeof nm, "%s.bias", name);
        if (hw.get(nm, MIMI_PFX, cout, b) != 0) return -1;
    }
    std::vector<float> m((size_t)cout * 2 * stride * cin);
    for (int o = 0; o < cout; o++)
        for (int tap = 0; tap < 2; tap++)
            for (int ph = 0; ph < stride; ph++)
                for (int i = 0; i < cin; i++)
                    m[(((size_t)o * 2 + tap) * stride + ph) * cin + i] = w[((size_t)o * cin + i) * k + tap * stride + ph];
    c.cin = stride * cin; c.cout = cout; c.k = k; c.stride = stride; c.n_rows = cout; c.k_per_tap = stride * cin; c.transposed_conv = false;
    if (e->up_f16(&c.w, m) != 0 || e->up_f32(&c.b, b) != 0) return -1;
    return 0;
}

// Optional: returns 0 with has_encoder = false when the checkpoint has no encoder tensors.
static int enc_load_weights(Engine *e, const HostWeights &hw) {
    std::vector<float> v;
    char nm[256];
    e->has_encoder = false;
    if (!hw.find("encoder.model.0.conv.weight", MIMI_PFX) || !hw.find("speaker_proj_weight", FLOW_PFX)) return 0;
    if (hw.get("encoder.model.0.conv.weight", MIMI_PFX, 64 * 7, v) != 0 || e->up_f32(&e->enc_in_w, v) != 0) return -1;
    if (hw.get("encoder.model.0.conv.bias", MIMI_PFX, 64, v) != 0 || e->up_f32(&e->enc_in_b, v) != 0) return -1;
    static const int idx[3] = {1, 4, 7}, ratio[3] = {4, 5, 6};
    int ch = 64;
    for (int s = 0; s < 3; s++) {
        const int hid = ch / 2, hp = hid < 64 ? 64 : hid;   // the 32-channel bottleneck is zero-padded to 64 like the decoder's
        snprintf(nm, sizeof nm, "encoder.model.%d.block.1.conv", idx[s]);
        if (mimi_load_conv(e, hw, nm, e->eres1[s], ch, hid, 3, ch, hp) != 0) return -1;
        snprintf(nm, sizeof nm, "encoder.model.%d.block.3.conv", idx[s]);
        if (mimi_load_conv(e, hw, nm, e->eres2[s], hid, ch, 1, hp, ch) != 0) return -1;
        snprintf(nm, sizeof nm, "encoder.model.%d.conv", idx[s] + 2);
        if (enc_load_strided(e, hw, nm, e->edown[s], ch, 2 * ch, ratio[s], true) != 0) return -1;
        ch *= 2;
    }
    if (mimi_load_conv(e, hw, "encoder.model.11.conv", e->enc_out, 512, 512, 3, 512, 512) != 0) return -1;
    if (enc_load_strided(e, hw, "downsample.conv.conv", e->down, 512, 512, 16, false) != 0) return -1;
    for (int l = 0; l < MI_L; l++) {
#define EW(field, suffix, ne, up) snprintf(nm, sizeof nm, "encoder_transformer.transformer.layers.%d." suffix, l); \
        if (hw.get(nm, MIMI_PFX, ne, v) != 0 || e->up(&e->EL[l].field, v) != 0) return -1
        EW(inproj, "self_attn.in_proj.weight", 3 * MI_D * MI_D, up_f16);
        EW(outproj, "self_attn.out_proj.weight", MI_D * MI_D, up_f16);
        EW(l1, "linear1.weight", MI_FF * MI_D, up_f16);
        EW(l2, "linear2.weight", MI_D * MI_FF, up_f16);
        EW(n1w, "norm1.weight", MI_D, up_f32); EW(n1b, "norm1.bias", MI_D, up_f32);
        EW(n2w, "norm2.weight", MI_D, up_f32); EW(n2b, "norm2.bias", MI_D, up_f32);
        EW(ls1, "layer_scale_1.scale", MI_D, up_f32); EW(ls2, "layer_scale_2.scale", MI_D, up_f32);
#undef EW
    }
    if (hw.get("speaker_proj_weight", FLOW_PFX, (int64_t)FL_D * MI_D, v) != 0 || e->up_f16(&e->spk16, v) != 0) return -1;
    e->has_encoder = true;
    return 0;
}

// First conv of the encoder (1 -> 64 channels, k 7, causal): pcm [T] fp32 -> u [T][64] fp16 and ELU(u). grid ceil(T / 4), block 256.
__global__ void __launch_bounds__(256) k_enc_front(const float *pcm, int T, const float *w /*[64][7]*/, const float *b, __half *u, __half *eu) {
    pdl_wait(); pdl_trigger();
    const int t = blockIdx.x * 4 + (threadIdx.x >> 6), oc = threadIdx.x & 63;
    if (t >= T) return;
    float a = b[oc];
#pragma unroll
    for (int kk = 0; kk < 7; kk++) {
        const int i = t - 6 + kk;
        if (i >= 0) a += w[oc * 7 + kk] * pcm[i];
    }
    u[(size_t)t * 64 + oc] = __float2half_rn(a);
    eu[(size_t)t * 64 + oc] = __float2half_rn(a >= 0.0f ? a : expm1f(a));
}

// pcm: device buffer of frames * 1920 samples (zero padded); d_cond: device [frames][1024] fp32.
static int enc_run(Engine *e, const float *d_pcm, int frames, float *d_cond) {
    const int T0 = frames * 1920, T200 = frames * 16;
    if (ensure_rope(e, T200) != 0) return -1;
    // workspace: per 12.5 Hz frame 1920 rows x 64 ch x (u, eu, vv) + 480 x 128 x 3 + 96 x 256 x 3 + 16 x 512 x (fp32 x, fp16 xn/q/k/v/att/eu3, ffn)
    const size_t per_frame = (size_t)1920 * 64 * 2 * 3 + 480 * 128 * 2 * 3 + 96 * 256 * 2 * 3 + 16 * ((size_t)MI_D * 4 + MI_D * 2 * 7 + MI_FF * 2) + MI_D * 2 + 8192;
    const size_t need = per_frame * (size_t)frames + (1 << 20);
    if (need > e->mimi_ws_bytes) {
        cudaStreamSynchronize(e->st);
        e->mimi_ws_bytes = 0;
        e->dfree(e->mimi_ws); e->mimi_ws = nullptr;
        uint8_t *p = nullptr;
        if (e->dalloc(&p, need) != 0) return -1;
        e->mimi_ws = p; e->mimi_ws_bytes = need;
    }
    Bump ws{(uint8_t *)e->mimi_ws};
    __half *u = ws.take<__half>((size_t)T0 * 64), *eu = ws.take<__half>((size_t)T0 * 64), *vv = ws.take<__half>((size_t)T0 * 64);
    {
        LaunchScope ls(e, KC_MIMI_MISC);
        launch_k(k_enc_front, dim3((unsigned)((T0 + 3) / 4)), dim3(256), 0, e->st, d_pcm, T0, (const float *)e->enc_in_w, (const float *)e->enc_in_b, u, eu);
        if (launch_check("k_enc_front") != 0) return -1;
    }
    int T = T0, C = 64;
    const __half *cur = nullptr;
    for (int s = 0; s < 3; s++) {
        // vv = ELU(u + conv_k1(ELU(conv_k3(ELU(u)))))  (resblock, then the ELU in front of the strided conv)
        if (resblock_launch(e, e->eres1[s], e->eres2[s], eu, u, vv, 1, T, C) != 0) return -1;
        const ConvW &D = e->edown[s];
        const int T2 = T / D.stride, C2 = D.cout;
        GemmParams g = {};
        g.bias = D.b;
        __half *u2 = nullptr, *eu2 = ws.take<__half>((size_t)T2 * C2), *vv2 = nullptr;
        if (s < 2) {
            u2 = ws.take<__half>((size_t)T2 * C2); vv2 = ws.take<__half>((size_t)T2 * C2);
            g.out_raw = (uint16_t *)u2;                       // the next block's residual stream
        }
        g.out_h = (uint16_t *)eu2; g.ldh = C2; g.act = ACT_ELU;   // ELU(conv): input of the next block / of the last conv
        // rows t-1 and t of vv viewed as [T2][stride * C]
        if (mimi_gemm(e, vv, 1, T2, D.k_per_tap, D.w, D.n_rows, 2, -1, 1, C2 >= 256 ? 256 : 128, g) != 0) return -1;
        u = u2; eu = eu2; vv = vv2; T = T2; C = C2;
        cur = eu2;
    }
    // conv k3 512 -> 512 -> fp32 residual stream of the encoder transformer
    float *x = ws.take<float>((size_t)T200 * MI_D);
    __half *xn = ws.take<__half>((size_t)T200 * MI_D);
    __half *q = ws.take<__half>((size_t)T200 * MI_D), *k = ws.take<__half>((size_t)T200 * MI_D), *v = ws.take<__half>((size_t)T200 * MI_D);
    __half *att = ws.take<__half>((size_t)T200 * MI_D), *hb = ws.take<__half>((size_t)T200 * MI_FF);
    {
        GemmParams g = {};
        g.out_f32 = x; g.ldo = MI_D;
        if (conv_gemm(e, e->enc_out, cur, 1, T200, 256, g) != 0) return -1;
    }
    if (mimi_transformer_layers(e, e->EL, x, xn, q, k, v, att, hb, 1, T200, 0) != 0) return -1;
    // learnt downsampler 200 Hz -> 12.5 Hz (k 32, stride 16, no bias), then the speaker projection 512 -> 1024
    __half *lat = ws.take<__half>((size_t)frames * MI_D);
    {
        GemmParams g = {};
        g.bias = e->down.b; g.out_h = (uint16_t *)lat; g.ldh = MI_D; g.act = ACT_NONE;
        if (mimi_gemm(e, xn, 1, frames, e->down.k_per_tap, e->down.w, e->down.n_rows, 2, -1, 1, 128, g) != 0) return -1;
        g = {};
        g.out_f32 = d_cond; g.ldo = FL_D;
        if (mimi_gemm(e, lat, 1, frames, MI_D, e->spk16, FL_D, 1, 0, 0, 128, g) != 0) return -1;
    }
    return 0;
}

}  // namespace ptts
