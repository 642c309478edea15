// mimi.cuh - Mimi decoder (ptts_mimi.c:570-738) on the device, time-major end to end:
// front end -> 2-layer windowed transformer -> SEANet conv stack, every linear / Conv1d /
// ConvTranspose1d running as a tcgen05 implicit GEMM (gemm_tc.cuh) with bias / layer-scale /
// residual / ELU / GELU fused in the epilogue. fp16 operands, fp32 accumulation, fp32 residual
// streams. No [C,T] <-> [T,C] transposes exist (ptts_mimi.c:134-150 eliminated by layout choice).
#pragma once
#include "engine.cuh"
#include "resblock.cuh"

namespace ptts {

// Conv1d weight [out,in,k] -> GEMM B matrix [n_rows][k*cin_pad] with column tap*cin_pad + ic
// (tap kk reads input row t-(k-1)+kk: ptts_kernels.c:128-153 with stride 1).
static int mimi_load_conv(Engine *e, const HostWeights &hw, const char *name, ConvW &c, int cin, int cout, int k, int cin_pad,
                          int cout_pad) {
    std::vector<float> w, b;
    char nm[256];
    snprintf(nm, sizeof nm, "%s.weight", name);
    if (hw.get(nm, MIMI_PFX, (int64_t)cin * cout * k, w) != 0) return -1;
    snprintf(nm, sizeof nm, "%s.bias", name);
    if (hw.get(nm, MIMI_PFX, cout, b) != 0) return -1;
    std::vector<float> m((size_t)cout_pad * k * cin_pad, 0.0f), bp(cout_pad, 0.0f);
    for (int o = 0; o < cout; o++) {
        bp[o] = b[o];
        for (int i = 0; i < cin; i++)
            for (int kk = 0; kk < k; kk++) m[((size_t)o * k + kk) * cin_pad + i] = w[((size_t)o * cin + i) * k + kk];
    }
    c.cin = cin_pad; c.cout = cout_pad; c.k = k; c.stride = 1; c.n_rows = cout_pad; c.k_per_tap = cin_pad; c.transposed_conv = false;
    if (e->up_f16(&c.w, m) != 0 || e->up_f32(&c.b, bp) != 0) return -1;
    return 0;
}

// ConvTranspose1d weight [in,out,k], k = 2*stride -> B matrix [stride*cout][2*cin]:
// row r*cout+oc, column tap*cin+ic = w[ic][oc][r + tap*stride]; tap 0 reads x[t], tap 1 reads x[t-1]
// (each output sample sees exactly two taps per input channel: ptts_kernels.c:172-204).
static int mimi_load_convtr(Engine *e, const HostWeights &hw, const char *name, ConvW &c, int cin, int cout, int k, int stride) {
    std::vector<float> w, b;
    char nm[256];
    PTTS_CHECK(k == 2 * stride, "convtr %s: kernel %d != 2*stride %d", name, k, stride);
    snprintf(nm, sizeof nm, "%s.weight", name);
    if (hw.get(nm, MIMI_PFX, (int64_t)cin * cout * k, w) != 0) return -1;
    snprintf(nm, sizeof nm, "%s.bias", name);
    if (hw.get(nm, MIMI_PFX, cout, b) != 0) return -1;
    std::vector<float> m((size_t)stride * cout * 2 * cin);
    for (int r = 0; r < stride; r++)
        for (int o = 0; o < cout; o++)
            for (int tap = 0; tap < 2; tap++)
                for (int i = 0; i < cin; i++)
                    m[((size_t)(r * cout + o) * 2 + tap) * cin + i] = w[((size_t)i * cout + o) * k + r + tap * stride];
    c.cin = cin; c.cout = cout; c.k = k; c.stride = stride; c.n_rows = stride * cout; c.k_per_tap = cin; c.transposed_conv = true;
    if (e->up_f16(&c.w, m) != 0 || e->up_f32(&c.b, b) != 0) return -1;
    return 0;
}

static int mimi_load_weights(Engine *e, const HostWeights &hw) {
    std::vector<float> v;
    char nm[256];
    auto transpose_512x32 = [](const std::vector<float> &in) {   // [512,32] -> [32,512]: coalesced per-channel reads
        std::vector<float> o(in.size());
        for (int c = 0; c < MI_D; c++) for (int i = 0; i < 32; i++) o[(size_t)i * MI_D + c] = in[(size_t)c * 32 + i];
        return o;
    };
    if (hw.get("quantizer.output_proj.weight", MIMI_PFX, MI_D * 32, v) != 0 || e->up_f32(&e->quant_w, transpose_512x32(v)) != 0) return -1;
    int r = hw.get("upsample.convtr.weight", MIMI_PFX, MI_D * 32, v, true);   // ptts_mimi.c:385-388
    if (r < 0) return -1;
    if (r == 1 && hw.get("upsample.convtr.convtr.weight", MIMI_PFX, MI_D * 32, v) != 0) return -1;
    if (e->up_f32(&e->ups_w, transpose_512x32(v)) != 0) return -1;
    for (int l = 0; l < MI_L; l++) {
#define MW(field, suffix, ne, up) snprintf(nm, sizeof nm, "decoder_transformer.transformer.layers.%d." suffix, l); \
        if (hw.get(nm, MIMI_PFX, ne, v) != 0 || e->up(&e->ML[l].field, v) != 0) return -1
        MW(inproj, "self_attn.in_proj.weight", 3 * MI_D * MI_D, up_f16);
        MW(outproj, "self_attn.out_proj.weight", MI_D * MI_D, up_f16);
        MW(l1, "linear1.weight", MI_FF * MI_D, up_f16);
        MW(l2, "linear2.weight", MI_D * MI_FF, up_f16);
        MW(n1w, "norm1.weight", MI_D, up_f32); MW(n1b, "norm1.bias", MI_D, up_f32);
        MW(n2w, "norm2.weight", MI_D, up_f32); MW(n2b, "norm2.bias", MI_D, up_f32);
        MW(ls1, "layer_scale_1.scale", MI_D, up_f32); MW(ls2, "layer_scale_2.scale", MI_D, up_f32);
#undef MW
    }
    // geometry is hard-coded in the reference (ptts_mimi.c:389-486)
    if (mimi_load_conv(e, hw, "decoder.model.0.conv", e->dec_in, 512, 512, 7, 512, 512) != 0) return -1;
    if (mimi_load_convtr(e, hw, "decoder.model.2.convtr", e->up[0], 512, 256, 12, 6) != 0) return -1;
    if (mimi_load_conv(e, hw, "decoder.model.3.block.1.conv", e->res1[0], 256, 128, 3, 256, 128) != 0) return -1;
    if (mimi_load_conv(e, hw, "decoder.model.3.block.3.conv", e->res2[0], 128, 256, 1, 128, 256) != 0) return -1;
    if (mimi_load_convtr(e, hw, "decoder.model.5.convtr", e->up[1], 256, 128, 10, 5) != 0) return -1;
    if (mimi_load_conv(e, hw, "decoder.model.6.block.1.conv", e->res1[1], 128, 64, 3, 128, 64) != 0) return -1;
    if (mimi_load_conv(e, hw, "decoder.model.6.block.3.conv", e->res2[1], 64, 128, 1, 64, 128) != 0) return -1;
    if (mimi_load_convtr(e, hw, "decoder.model.8.convtr", e->up[2], 128, 64, 8, 4) != 0) return -1;
    // 32-channel bottleneck padded to 64 channels (zero weights) so that K stays a multiple of 64
    if (mimi_load_conv(e, hw, "decoder.model.9.block.1.conv", e->res1[2], 64, 32, 3, 64, 64) != 0) return -1;
    if (mimi_load_conv(e, hw, "decoder.model.9.block.3.conv", e->res2[2], 32, 64, 1, 64, 64) != 0) return -1;
    if (hw.get("decoder.model.11.conv.weight", MIMI_PFX, 64 * 3, v) != 0 || e->up_f32(&e->dec_out_w, v) != 0) return -1;
    if (hw.get("decoder.model.11.conv.bias", MIMI_PFX, 1, v) != 0 || e->up_f32(&e->dec_out_b, v) != 0) return -1;
    e->h_dec_out_b = v[0];
    return 0;
}

// Workspace bytes per latent frame (16 tokens, 96 / 480 / 1920 rows in the three SEANet stages).
static size_t mimi_ws_per_frame() {
    const size_t tok = 16 * (size_t)(MI_D * 4 + MI_D * 2 + 3 * MI_D * 2 + MI_D * 2 + MI_FF * 2 + MI_D * 2);
    const size_t s0 = 96 * (size_t)(256 * 2 + 256 * 2 + 128 * 2 + 256 * 2);
    const size_t s1 = 480 * (size_t)(128 * 2 + 128 * 2 + 64 * 2 + 128 * 2);
    const size_t s2 = 1920 * (size_t)(64 * 2 + 64 * 2 + 64 * 2 + 64 * 2);
    return tok + s0 + s1 + s2 + 4096;
}

struct Bump {
    uint8_t *base; size_t off = 0;
    template <typename T> T *take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T *p = (T *)(base + off);
        off += count * sizeof(T);
        return p;
    }
};

// Activations [Bc][T][C] fp16 as the A operand, weights [N][taps*K] fp16 as the B operand.
static int mimi_gemm(Engine *e, const __half *A, int Bc, int T, int C, const __half *W, int n_rows, int taps, int shift0, int step,
                     int bn, GemmParams epi) {
    GemmDesc d;
    d.A = A; d.a_k = C; d.a_rows = T; d.a_batches = Bc; d.a_ld = C; d.a_bstride = (size_t)T * C;
    d.B = W; d.b_k = taps * C; d.b_rows = n_rows; d.b_ld = (size_t)taps * C;
    d.bf16 = false; d.taps = taps; d.k_per_tap = C; d.shift0 = shift0; d.step = step; d.a_k_step = 0; d.b_k_step = C;
    d.bn = bn; d.split_k = 1; d.two_cta = true; d.num_sms = e->num_sms;
    d.epi = epi;
    d.epi.fast_act = 1;   // fp16 outputs: hardware exp/tanh approximations are below the rounding error
    GemmOp op;
    if (gemm_build(op, d) != 0) return -1;
    return run_gemm(e, op, KC_GEMM_MIMI);
}

static int conv_gemm(Engine *e, const ConvW &c, const __half *A, int Bc, int T, int bn, GemmParams epi) {
    epi.bias = c.b;
    if (c.transposed_conv) {
        epi.bias_mask = c.cout - 1;
        return mimi_gemm(e, A, Bc, T, c.cin, c.w, c.n_rows, 2, 0, -1, bn, epi);
    }
    return mimi_gemm(e, A, Bc, T, c.cin, c.w, c.n_rows, c.k, -(c.k - 1), 1, bn, epi);
}

static int mimi_set_attrs() {
    PTTS_CUDA_OK(cudaFuncSetAttribute(resblock_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PTTS_CUDA_OK(cudaFuncSetAttribute(resblock_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PTTS_CUDA_OK(cudaFuncSetAttribute(resblock_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return 0;
}

// One fused residual block (resblock.cuh): vv = ELU(u + conv1(ELU(conv3(eu)))) for Bc items of T rows, C2 channels.
template <int C2>
static int resblock_launch_t(Engine *e, const ConvW &c1, const ConvW &c2, const __half *eu, const __half *u, __half *vv, int Bc, int T,
                             float *pcm = nullptr, float *edge = nullptr) {
    // (the kernel's dynamic shared-memory attribute is set per device by ptts_b200_create: mimi_set_attrs)
    const int chp = c1.cout;   // padded to a multiple of 64 at load time
    PTTS_CHECK(chp % 64 == 0 && chp <= 128 && c1.cin == C2 && c1.k == 3 && c2.cin == chp && c2.cout == C2 && c2.k == 1,
               "resblock_launch: unexpected block geometry");
    CUtensorMap mapA, mapW1, mapW2;
    if (make_tmap(&mapA, eu, false, C2, T, Bc, (size_t)C2, (size_t)T * C2, GEMM_BM) != 0) return -1;
    if (make_tmap(&mapW1, c1.w, false, 3 * C2, chp, 1, (size_t)3 * C2, 0, chp) != 0) return -1;
    if (make_tmap(&mapW2, c2.w, false, chp, C2, 1, (size_t)chp, 0, C2) != 0) return -1;
    ResParams p = {};
    p.tiles_per_batch = (T + GEMM_BM - 1) / GEMM_BM;
    const long long tiles = (long long)Bc * p.tiles_per_batch;
    PTTS_CHECK(tiles < (1ll << 31), "resblock_launch: too many tiles");
    p.tiles_total = (int)tiles; p.a_rows = T; p.kb_per_tap = C2 / 64; p.chp = chp;
    const int stage = GEMM_BM * GEMM_BK * 2 + chp * 128, kb2 = chp / 64;
    // C2 == 64: + the output-tile operand buffer of the fused final conv (16 KB)
    const int fixed = kb2 * C2 * 128 + kb2 * GEMM_BM * GEMM_BK * 2 + (C2 == 64 ? 2048 + GEMM_BM * GEMM_BK * 2 : 0) + 256 + (192 + 2 * 768) * 4 + 1024;
    // C2 <= 128: two CTAs per SM (one's output epilogue overlaps the other's contractions) -> at most 113 KB each
    p.stages = C2 <= 128 ? std::min(3, (113 * 1024 - fixed) / stage) : std::min(GEMM_MAX_STAGES, (227 * 1024 - fixed) / stage);
    PTTS_CHECK(p.stages >= 2, "resblock_launch: shared memory too small");
    p.idesc1 = umma_idesc_f16(GEMM_BM, std::min(chp, C2 / 2), false, false);   // N = the real middle channels (32 for the 64-channel block)
    p.idesc2 = umma_idesc_f16(GEMM_BM, C2, false, false);
    p.idesc3 = umma_idesc_f16(GEMM_BM, 16, false, false);
    p.b1 = c1.b; p.b2 = c2.b; p.u = (const uint16_t *)u; p.out = (uint16_t *)vv;
    p.prefetch_u = getenv("PTTS_B200_RB_NO_PREFETCH") == nullptr ? 1 : 0;   // measured: -4 % (64 ch) / -6 % (128 ch) per launch
    if (pcm) {   // last block: the final 64 -> 1 conv is applied in the epilogue, the block's output tensor is never written
        PTTS_CHECK(C2 == 64 && edge, "resblock_launch: the final conv can only be fused into the 64-channel block");
        p.dbias = e->h_dec_out_b;
        p.dw = e->dec_out_w; p.pcm = pcm; p.edge = edge; p.out = nullptr;
    }
    const int smem = p.stages * stage + fixed;
    if (e->prof.on) {
        const double rows = (double)Bc * T;
        e->prof.flops[KC_GEMM_MIMI] += 2.0 * rows * (3.0 * C2 * chp + (double)chp * C2);
        e->prof.bytes[KC_GEMM_MIMI] += rows * C2 * 2.0 * 3.0;   // ELU(u) in, u in, out: the middle tensor stays on chip
    }
    {
        LaunchScope ls(e, KC_GEMM_MIMI);
        launch_k(resblock_kernel<C2>, dim3((unsigned)std::min<long long>(tiles, (long long)e->num_sms * (C2 <= 128 ? 2 : 1))), dim3(RB_THREADS), (size_t)smem,
                 e->st, mapA, mapW1, mapW2, p);
    }
    if (e->prof.on && !e->prof.recs.empty()) {
        Prof::Rec &r = e->prof.recs.back();
        r.rows = Bc * T; r.n = C2; r.k = 3 * C2 + chp; r.bn = -chp;   // bn < 0 marks the fused block in the per-launch log
        r.flops = 2.0 * (double)Bc * T * (3.0 * C2 * chp + (double)chp * C2);
    }
    if (launch_check("resblock_kernel") != 0) return -1;
    if (pcm) {
        LaunchScope ls(e, KC_MIMI_MISC);
        launch_k(k_dec_edge_fix, dim3((unsigned)((tiles + 255) / 256)), dim3(256), 0, e->st, (const float *)edge, pcm, p.tiles_per_batch, p.tiles_total, T);
        return launch_check("k_dec_edge_fix");
    }
    return 0;
}
static int resblock_launch(Engine *e, const ConvW &c1, const ConvW &c2, const __half *eu, const __half *u, __half *vv, int Bc, int T, int C2,
                           float *pcm = nullptr, float *edge = nullptr) {
    if (C2 == 64) return resblock_launch_t<64>(e, c1, c2, eu, u, vv, Bc, T, pcm, edge);
    if (C2 == 128) return resblock_launch_t<128>(e, c1, c2, eu, u, vv, Bc, T);
    if (C2 == 256) return resblock_launch_t<256>(e, c1, c2, eu, u, vv, Bc, T);
    set_error("resblock_launch: unsupported channel count %d", C2);
    return -1;
}

// SEANet conv stack (ptts_mimi.c:676-737) over Bc items of T transformer-output rows xn [Bc][T][512] fp16 -> d_pcm [Bc][T*120].
// d0 = scratch for the first conv's output; the other intermediates come from the bump allocator.
static int mimi_seanet(Engine *e, Bump &ws, int Bc, int T, const __half *xn, __half *d0, float *d_pcm) {
    GemmParams g = {};
    g.out_h = (uint16_t *)d0; g.ldh = MI_D; g.act = ACT_ELU;                               // ELU(dec_in(x))
    if (conv_gemm(e, e->dec_in, xn, Bc, T, 256, g) != 0) return -1;
    const __half *cur = d0;
    int Tc = T;
    __half *last = nullptr;
    bool pcm_done = false;
    for (int s = 0; s < 3; s++) {
        const ConvW &U = e->up[s];
        const int T2 = Tc * U.stride, C2 = U.cout, Ch = e->res1[s].cout;
        const size_t R2 = (size_t)Bc * T2;
        // the resblock's residual stream u is kept in fp16 like every other SEANet activation (2 B instead of 4 B
        // per element on the two widest tensors of the decoder; the GEMM operands are fp16 anyway)
        __half *u = ws.take<__half>(R2 * C2);
        __half *eu = ws.take<__half>(R2 * C2), *rm = ws.take<__half>(R2 * Ch), *vv = ws.take<__half>(R2 * C2);
        g = {}; g.out_raw = (uint16_t *)u; g.out_h = (uint16_t *)eu; g.ldh = U.n_rows; g.act = ACT_ELU;
        if (conv_gemm(e, U, cur, Bc, Tc, U.n_rows >= 256 ? 256 : 128, g) != 0) return -1;  // u = convT(ELU(prev)), eu = ELU(u)
        static const bool fused = getenv("PTTS_B200_NO_RESBLOCK") == nullptr;
        static const bool fuse_out = getenv("PTTS_B200_NO_FUSED_OUT") == nullptr;
        if (fused && s == 2 && fuse_out) {   // last block: + the final 64 -> 1 conv, straight to PCM
            float *edge = ws.take<float>((size_t)Bc * ((T2 + GEMM_BM - 1) / GEMM_BM) * 4);
            if (resblock_launch(e, e->res1[s], e->res2[s], eu, u, vv, Bc, T2, C2, d_pcm, edge) != 0) return -1;
            pcm_done = true;
        } else if (fused) {   // conv_k3 -> ELU -> conv_k1 -> +u -> ELU in one kernel, the middle tensor stays in shared memory
            if (resblock_launch(e, e->res1[s], e->res2[s], eu, u, vv, Bc, T2, C2) != 0) return -1;
        } else {
            g = {}; g.out_h = (uint16_t *)rm; g.ldh = Ch; g.act = ACT_ELU;                      // ELU(conv_k3(ELU(u)))
            if (conv_gemm(e, e->res1[s], eu, Bc, T2, Ch >= 128 ? 128 : 64, g) != 0) return -1;
            g = {}; g.resid_h = (const uint16_t *)u; g.out_h = (uint16_t *)vv; g.ldh = C2; g.act = ACT_ELU;  // ELU(u + conv_k1(.))
            if (conv_gemm(e, e->res2[s], rm, Bc, T2, C2 >= 256 ? 256 : C2, g) != 0) return -1;
        }
        cur = vv; Tc = T2; last = vv;
    }
    if (!pcm_done) {
        LaunchScope ls(e, KC_MIMI_MISC);
        launch_k(k_dec_out, dim3((Tc + 255) / 256, Bc), 256, 0, e->st, last, Tc, e->dec_out_w, e->dec_out_b, d_pcm);
        if (launch_check("k_dec_out") != 0) return -1;
    }
    return 0;
}


// The 2-layer windowed transformer (ptts_mimi.c:301-376) over x [Bc][T][512] fp32 (residual stream, updated in place) with the
// layer weights Lw (decoder_transformer, or encoder_transformer for the voice-cloning encoder). Leaves an fp16 copy of the result in xn.
static int mimi_transformer_layers(Engine *e, const MimiLayerW *Lw, float *x, __half *xn, __half *q, __half *k, __half *v, __half *att, __half *hb,
                                   int Bc, int T, int pos0) {
    const size_t R = (size_t)Bc * T;
    for (int l = 0; l < MI_L; l++) {
        {
            LaunchScope ls(e, KC_MIMI_MISC);
            launch_k(k_ln512_f16, (unsigned)((R + 7) / 8), 256, 0, e->st, x, R, Lw[l].n1w, Lw[l].n1b, 1e-5f, xn);
            if (launch_check("k_ln512_f16") != 0) return -1;
        }
        GemmParams g = {};
        g.rope_q = (uint16_t *)q; g.rope_k = (uint16_t *)k; g.rope_v = (uint16_t *)v;   // RoPE + head split fused in the epilogue
        g.rope_cos = e->rope_cos; g.rope_sin = e->rope_sin; g.rope_pos0 = pos0;
        if (mimi_gemm(e, xn, Bc, T, MI_D, Lw[l].inproj, 3 * MI_D, 1, 0, 0, 256, g) != 0) return -1;
        {
            if (e->prof.on) {   // banded attention: <= 250 keys per query, QK^T and PV
                double pairs = 0.0;
                for (int t = 0; t < T; t++) pairs += std::min(t + 1, MI_CTX);
                e->prof.flops[KC_MIMI_ATTN] += 4.0 * 64.0 * MI_H * pairs * Bc;
                e->prof.bytes[KC_MIMI_ATTN] += (double)R * MI_D * 2.0 * 4.0;
            }
            LaunchScope ls(e, KC_MIMI_ATTN);
            launch_k(k_mimi_win_attn, dim3((T + 63) / 64, MI_H, Bc), 128, 0, e->st, q, k, v, T, att, 0, T);
            if (launch_check("k_mimi_win_attn") != 0) return -1;
        }
        g = {}; g.scale = Lw[l].ls1; g.resid = x; g.out_f32 = x; g.ldo = MI_D;          // x += ls1 * out_proj(att)
        if (mimi_gemm(e, att, Bc, T, MI_D, Lw[l].outproj, MI_D, 1, 0, 0, 256, g) != 0) return -1;
        {
            LaunchScope ls(e, KC_MIMI_MISC);
            launch_k(k_ln512_f16, (unsigned)((R + 7) / 8), 256, 0, e->st, x, R, Lw[l].n2w, Lw[l].n2b, 1e-5f, xn);
            if (launch_check("k_ln512_f16") != 0) return -1;
        }
        g = {}; g.out_h = (uint16_t *)hb; g.ldh = MI_FF; g.act = ACT_GELU_TANH;
        if (mimi_gemm(e, xn, Bc, T, MI_D, Lw[l].l1, MI_FF, 1, 0, 0, 256, g) != 0) return -1;
        g = {}; g.scale = Lw[l].ls2; g.resid = x; g.out_f32 = x; g.ldo = MI_D;          // x += ls2 * linear2(h)
        if (l == MI_L - 1) { g.out_h = (uint16_t *)xn; g.ldh = MI_D; g.act = ACT_NONE; }   // fp16 copy feeds dec_in
        if (mimi_gemm(e, hb, Bc, T, MI_FF, Lw[l].l2, MI_D, 1, 0, 0, 256, g) != 0) return -1;
    }
    return 0;
}

// Decode Bc equal-length items. d_lat [Bc,F,32] fp32 on the device; d_pcm [Bc, F*1920].
// pos0 = token position of the first row (RoPE is absolute: ptts_mimi.c:341).
static int mimi_run_chunk(Engine *e, int Bc, int F, const float *d_lat, size_t lat_stride, int apply_scale, float *d_pcm, int pos0 = 0) {
    const int T = 16 * F;
    const size_t R = (size_t)Bc * T;
    if (ensure_rope(e, T + pos0) != 0) return -1;
    const size_t need = mimi_ws_per_frame() * (size_t)Bc * F + (1 << 20);
    if (need > e->mimi_ws_bytes) {
        cudaStreamSynchronize(e->st);
        e->mimi_ws_bytes = 0;
        e->dfree(e->mimi_ws); e->mimi_ws = nullptr;
        uint8_t *p = nullptr;
        if (e->dalloc(&p, need) != 0) return -1;
        e->mimi_ws = p; e->mimi_ws_bytes = need;
    }
    Bump ws{(uint8_t *)e->mimi_ws};
    float *x = ws.take<float>(R * MI_D);
    __half *xn = ws.take<__half>(R * MI_D);
    __half *q = ws.take<__half>(R * MI_D), *k = ws.take<__half>(R * MI_D), *v = ws.take<__half>(R * MI_D);
    __half *att = ws.take<__half>(R * MI_D);
    __half *hb = ws.take<__half>(R * MI_FF);
    __half *d0 = ws.take<__half>(R * MI_D);
    {
        LaunchScope ls(e, KC_MIMI_MISC);
        launch_k(k_mimi_front, dim3(F, Bc), 512, 0, e->st, d_lat, lat_stride, F, e->emb_std, e->emb_mean, apply_scale, e->quant_w, e->ups_w, x, 0);
        if (launch_check("k_mimi_front") != 0) return -1;
    }
    if (mimi_transformer_layers(e, e->ML, x, xn, q, k, v, att, hb, Bc, T, pos0) != 0) return -1;
    return mimi_seanet(e, ws, Bc, T, xn, d0, d_pcm);
}

// ---------------------------------------------------------------- streaming decode (single utterance, carried state)
// The decoder is causal, so a chunk of nf new latent frames needs only state from the past (README.md:162-167 roadmap item,
// SURVEY.md section 8 f2):
//   * the previous latent frame (second tap of the k = 32 / s = 16 depthwise upsampler, ptts_mimi.c:588-593): read in place from
//     the utterance's latent buffer;
//   * per transformer layer, the RoPE'd K / V rows of the last 249 tokens (window 250, ptts_mimi.c:264-299): the chunk's QKV
//     epilogue appends its rows at their absolute positions to per-utterance caches ms_k / ms_v (RoPE positions are absolute);
//     the banded attention kernel runs for the chunk's query rows only;
//   * the causal conv tails of the SEANet stack (ptts_kernels.c:128-153, 172-204): its receptive field is exactly 8 transformer
//     rows (dec_in k = 7: 6 rows; each ConvTranspose k = 2s: one input row; each residual block's k = 3 conv: two rows of its own
//     rate; the final k = 3 conv: two samples - 2 rows in total after the three up-sampling stages), so the stack is run over
//     [8 carried transformer-output rows | the chunk's rows] and only the chunk's samples are kept. The 8 rows are the conv
//     tails' source; recomputing the tails from them costs half a frame of conv work per chunk and keeps one copy of the state.
// A chunk therefore costs nf (+ 0.5 conv) frames of Mimi work whatever the position in the utterance, and the samples equal the
// one-shot decode (same kernels, same operands; tests/test_gpu_measured_configs.py checks against the oracle's one-shot PCM).
static int mimi_stream_reserve(Engine *e, int frames) {
    const int cap = (16 * frames + 63) / 64 * 64;
    if (cap <= e->ms_cap) return 0;
    cudaStreamSynchronize(e->st);
    e->ms_cap = 0;
    e->dfree(e->ms_q); e->dfree(e->ms_xn); e->ms_q = nullptr; e->ms_xn = nullptr;
    for (int l = 0; l < MI_L; l++) { e->dfree(e->ms_k[l]); e->dfree(e->ms_v[l]); e->ms_k[l] = e->ms_v[l] = nullptr; }
    if (e->dalloc(&e->ms_q, (size_t)cap * MI_D) != 0 || e->dalloc(&e->ms_xn, (size_t)cap * MI_D) != 0) return -1;
    for (int l = 0; l < MI_L; l++)
        if (e->dalloc(&e->ms_k[l], (size_t)cap * MI_D) != 0 || e->dalloc(&e->ms_v[l], (size_t)cap * MI_D) != 0) return -1;
    e->ms_cap = cap;
    return 0;
}

constexpr int MIMI_CONV_HALO = 8;   // transformer rows the SEANet stack looks back (derivation above)

// Decode latent frames [f0, f0 + nf) of the streamed utterance; d_lat points at frame f0 of its latent buffer (raw FlowLM latents
// when apply_scale). *pcm_out = device pointer to the chunk's nf * 1920 samples (inside the workspace, valid until the next call).
static int mimi_stream_chunk(Engine *e, int f0, int nf, const float *d_lat, int apply_scale, float **pcm_out) {
    const int pos0 = 16 * f0, Tn = 16 * nf, T_end = pos0 + Tn;
    PTTS_CHECK(T_end <= e->ms_cap, "mimi_stream_chunk: stream state holds %d tokens, chunk ends at %d", e->ms_cap, T_end);
    if (ensure_rope(e, T_end) != 0) return -1;
    const int w0 = std::max(0, pos0 - MIMI_CONV_HALO), Tw = T_end - w0;
    const size_t need = mimi_ws_per_frame() * (size_t)(nf + 1) + (size_t)Tw * 120 * 4 + (1 << 20);
    if (need > e->mimi_ws_bytes) {
        cudaStreamSynchronize(e->st);
        e->mimi_ws_bytes = 0;
        e->dfree(e->mimi_ws); e->mimi_ws = nullptr;
        uint8_t *p = nullptr;
        if (e->dalloc(&p, need) != 0) return -1;
        e->mimi_ws = p; e->mimi_ws_bytes = need;
    }
    Bump ws{(uint8_t *)e->mimi_ws};
    const size_t R = (size_t)Tn;
    float *x = ws.take<float>(R * MI_D);
    __half *xn = ws.take<__half>(R * MI_D), *att = ws.take<__half>(R * MI_D), *hb = ws.take<__half>(R * MI_FF);
    __half *d0 = ws.take<__half>((size_t)Tw * MI_D);
    float *pcm = ws.take<float>((size_t)Tw * 120);
    {
        LaunchScope ls(e, KC_MIMI_MISC);
        launch_k(k_mimi_front, dim3(nf, 1), 512, 0, e->st, d_lat, (size_t)0, nf, e->emb_std, e->emb_mean, apply_scale, e->quant_w, e->ups_w, x, f0 > 0 ? 1 : 0);
        if (launch_check("k_mimi_front") != 0) return -1;
    }
    for (int l = 0; l < MI_L; l++) {
        {
            LaunchScope ls(e, KC_MIMI_MISC);
            launch_k(k_ln512_f16, (unsigned)((R + 7) / 8), 256, 0, e->st, x, R, e->ML[l].n1w, e->ML[l].n1b, 1e-5f, xn);
            if (launch_check("k_ln512_f16") != 0) return -1;
        }
        GemmParams g = {};
        g.rope_q = (uint16_t *)e->ms_q; g.rope_k = (uint16_t *)e->ms_k[l]; g.rope_v = (uint16_t *)e->ms_v[l];
        g.rope_cos = e->rope_cos; g.rope_sin = e->rope_sin; g.rope_pos0 = pos0; g.rope_rows = e->ms_cap; g.rope_row0 = pos0;
        if (mimi_gemm(e, xn, 1, Tn, MI_D, e->ML[l].inproj, 3 * MI_D, 1, 0, 0, 256, g) != 0) return -1;
        {
            LaunchScope ls(e, KC_MIMI_ATTN);
            launch_k(k_mimi_win_attn, dim3((Tn + 63) / 64, MI_H, 1), 128, 0, e->st, (const __half *)e->ms_q, (const __half *)e->ms_k[l],
                     (const __half *)e->ms_v[l], T_end, att, pos0, e->ms_cap);
            if (launch_check("k_mimi_win_attn") != 0) return -1;
        }
        g = {}; g.scale = e->ML[l].ls1; g.resid = x; g.out_f32 = x; g.ldo = MI_D;
        if (mimi_gemm(e, att, 1, Tn, MI_D, e->ML[l].outproj, MI_D, 1, 0, 0, 256, g) != 0) return -1;
        {
            LaunchScope ls(e, KC_MIMI_MISC);
            launch_k(k_ln512_f16, (unsigned)((R + 7) / 8), 256, 0, e->st, x, R, e->ML[l].n2w, e->ML[l].n2b, 1e-5f, xn);
            if (launch_check("k_ln512_f16") != 0) return -1;
        }
        g = {}; g.out_h = (uint16_t *)hb; g.ldh = MI_FF; g.act = ACT_GELU_TANH;
        if (mimi_gemm(e, xn, 1, Tn, MI_D, e->ML[l].l1, MI_FF, 1, 0, 0, 256, g) != 0) return -1;
        g = {}; g.scale = e->ML[l].ls2; g.resid = x; g.out_f32 = x; g.ldo = MI_D;
        if (l == MI_L - 1) { g.out_h = (uint16_t *)(e->ms_xn + (size_t)pos0 * MI_D); g.ldh = MI_D; g.act = ACT_NONE; }   // carried: feeds dec_in
        if (mimi_gemm(e, hb, 1, Tn, MI_FF, e->ML[l].l2, MI_D, 1, 0, 0, 256, g) != 0) return -1;
    }
    if (mimi_seanet(e, ws, 1, Tw, e->ms_xn + (size_t)w0 * MI_D, d0, pcm) != 0) return -1;
    *pcm_out = pcm + (size_t)(pos0 - w0) * 120;
    return 0;
}

// latents on the device [n,F,32] (item stride lat_stride floats) -> pcm on the device [n, F*1920], chunked over items so that the
// workspace stays within the budget (PTTS_B200_MIMI_WS_GB, default 24 GB). When host_out is given, the PCM of
// chunk i travels to the host on the copy stream while chunk i+1 is being decoded (at least 4 chunks, so that
// only the last quarter of the device->host transfer is exposed); the caller waits on e->ev_copy.
static int mimi_run(Engine *e, int n, int F, const float *d_lat, size_t lat_stride, int apply_scale, float *d_pcm, float *host_out = nullptr) {
    static const size_t budget = [] {   // initialised once, thread-safe
        const char *s = getenv("PTTS_B200_MIMI_WS_GB");
        return (size_t)((s && atof(s) > 0 ? atof(s) : 24.0) * (double)(1ull << 30));
    }();
    int per = (int)(budget / (mimi_ws_per_frame() * (size_t)F));
    if (per < 1) per = 1;
    if (per > n) per = n;
    if (host_out && n >= 8) per = std::min(per, (n + 3) / 4);
    // the copy of the LAST chunk cannot hide behind a later decode: keep it short (a sixteenth of the job)
    const int tail = (host_out && n >= 64) ? std::max(8, n / 16) : 0;
    for (int i = 0, bc = 0; i < n; i += bc) {
        bc = std::min(per, n - i);
        if (tail > 0 && n - i > tail && n - i - bc < tail) bc = n - i - tail;   // leave exactly `tail` utterances for the last chunk
        float *dst = d_pcm + (size_t)i * F * 1920;
        if (mimi_run_chunk(e, bc, F, d_lat + (size_t)i * lat_stride, lat_stride, apply_scale, dst) != 0) return -1;
        if (host_out) {
            PTTS_CUDA_OK(cudaEventRecord(e->ev_chunk, e->st));
            PTTS_CUDA_OK(cudaStreamWaitEvent(e->st_copy, e->ev_chunk, 0));
            PTTS_CUDA_OK(cudaMemcpyAsync(host_out + (size_t)i * F * 1920, dst, (size_t)bc * F * 1920 * 4, cudaMemcpyDeviceToHost, e->st_copy));
        }
    }
    if (host_out) {
        PTTS_CUDA_OK(cudaEventRecord(e->ev_copy, e->st_copy));
        e->pcm_on_host = host_out;
    }
    return 0;
}

}  // namespace ptts
