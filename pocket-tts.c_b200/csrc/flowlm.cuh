// flowlm.cuh - FlowLM side of the engine: weight upload, lock-step generation loop
// (ptts_flowlm.c:1138-1251) in two data paths:
//   * batch-1 streaming decode: fp32 activations, bf16 weight-streaming GEMV kernels;
//   * batched decode (B >= 2, padded to 16): tcgen05 GEMMs with the batch as the N operand
//     ("weights as A": D[feature, utterance]), activations carried as bf16 hi+lo.
// One frame step is captured once into a CUDA graph and replayed.
#pragma once
#include "engine.cuh"

namespace ptts {

static int flow_load_weights(Engine *e, const HostWeights &hw) {
    std::vector<float> v, v2;
    char nm[256];
#define GETF(name, ne) if (hw.get(name, FLOW_PFX, ne, v) != 0) return -1
    GETF("conditioner.embed.weight", 4001 * FL_D); if (e->up_bf16(&e->embed, v) != 0) return -1;
    GETF("emb_std", FL_LAT); memcpy(e->h_emb_std, v.data(), sizeof e->h_emb_std); if (e->up_f32(&e->emb_std, v) != 0) return -1;
    GETF("emb_mean", FL_LAT); memcpy(e->h_emb_mean, v.data(), sizeof e->h_emb_mean); if (e->up_f32(&e->emb_mean, v) != 0) return -1;
    GETF("input_linear.weight", FL_D * FL_LAT); if (e->up_bf16(&e->in_lin, v) != 0) return -1;
    {   // bos_proj = input_linear . bos_emb (ptts_flowlm.c:1179-1181), input-independent -> precompute (fp32)
        if (hw.get("bos_emb", FLOW_PFX, FL_LAT, v2) != 0) return -1;
        std::vector<float> bp(FL_D);
        for (int n = 0; n < FL_D; n++) {
            float a = 0.0f;
            for (int i = 0; i < FL_LAT; i++) a += h_bf2f(h_f2bf(v[n * FL_LAT + i])) * v2[i];
            bp[n] = a;
        }
        if (e->up_f32(&e->bos_proj, bp) != 0) return -1;
    }
    GETF("out_norm.weight", FL_D); if (e->up_f32(&e->onorm_w, v) != 0) return -1;
    GETF("out_norm.bias", FL_D); if (e->up_f32(&e->onorm_b, v) != 0) return -1;
    GETF("out_eos.weight", FL_D); if (e->up_f32(&e->eos_w, v) != 0) return -1;
    GETF("out_eos.bias", 1); if (e->up_f32(&e->eos_b, v) != 0) return -1;
    for (int l = 0; l < FL_L; l++) {
#define LW(field, suffix, ne, up) snprintf(nm, sizeof nm, "transformer.layers.%d." suffix, l); GETF(nm, ne); if (e->up(&e->L[l].field, v) != 0) return -1
        LW(inproj, "self_attn.in_proj.weight", 3 * FL_D * FL_D, up_bf16);
        LW(outproj, "self_attn.out_proj.weight", FL_D * FL_D, up_bf16);
        LW(l1, "linear1.weight", FL_FF * FL_D, up_bf16);
        LW(l2, "linear2.weight", FL_D * FL_FF, up_bf16);
        LW(n1w, "norm1.weight", FL_D, up_f32); LW(n1b, "norm1.bias", FL_D, up_f32);
        LW(n2w, "norm2.weight", FL_D, up_f32); LW(n2b, "norm2.bias", FL_D, up_f32);
#undef LW
    }
    GETF("flow_net.cond_embed.weight", FN_D * FL_D); if (e->up_bf16(&e->cond_w, v) != 0) return -1;
    GETF("flow_net.cond_embed.bias", FN_D); if (e->up_f32(&e->cond_b, v) != 0) return -1;
    GETF("flow_net.input_proj.weight", FN_D * FL_LAT); if (e->up_bf16(&e->inp_w, v) != 0) return -1;
    GETF("flow_net.input_proj.bias", FN_D); if (e->up_f32(&e->inp_b, v) != 0) return -1;
    for (int t = 0; t < 2; t++) {
#define TW(field, suffix, ne) snprintf(nm, sizeof nm, "flow_net.time_embed.%d." suffix, t); GETF(nm, ne); if (e->up_f32(&e->te[t].field, v) != 0) return -1
        TW(l0w, "mlp.0.weight", FN_D * 256); TW(l0b, "mlp.0.bias", FN_D);
        TW(l2w, "mlp.2.weight", FN_D * FN_D); TW(l2b, "mlp.2.bias", FN_D);
        TW(alpha, "mlp.3.alpha", FN_D);
#undef TW
        snprintf(nm, sizeof nm, "flow_net.time_embed.%d.freqs", t);
        int r = hw.get(nm, FLOW_PFX, 128, v, true);   // optional (ptts_flowlm.c:575)
        if (r < 0) return -1;
        e->te[t].freqs = nullptr;
        if (r == 0 && e->up_f32(&e->te[t].freqs, v) != 0) return -1;
    }
    // all adaLN projections stacked: rows [b*1536, (b+1)*1536) = shift|scale|gate of block b,
    // rows [9216, 10240) = shift|scale of the final layer. One GEMV/GEMM per flow pass instead of 7.
    std::vector<float> adaw((size_t)FN_ADA * FN_D), adab(FN_ADA);
    for (int b = 0; b < FN_DEPTH; b++) {
#define BW(field, suffix, ne, up) snprintf(nm, sizeof nm, "flow_net.res_blocks.%d." suffix, b); GETF(nm, ne); if (e->up(&e->FB[b].field, v) != 0) return -1
        BW(lnw, "in_ln.weight", FN_D, up_f32); BW(lnb, "in_ln.bias", FN_D, up_f32);
        BW(m0w, "mlp.0.weight", FN_D * FN_D, up_bf16); BW(m0b, "mlp.0.bias", FN_D, up_f32);
        BW(m2w, "mlp.2.weight", FN_D * FN_D, up_bf16); BW(m2b, "mlp.2.bias", FN_D, up_f32);
#undef BW
        snprintf(nm, sizeof nm, "flow_net.res_blocks.%d.adaLN_modulation.1.weight", b);
        GETF(nm, 3 * FN_D * FN_D);
        memcpy(&adaw[(size_t)b * 3 * FN_D * FN_D], v.data(), sizeof(float) * v.size());
        snprintf(nm, sizeof nm, "flow_net.res_blocks.%d.adaLN_modulation.1.bias", b);
        GETF(nm, 3 * FN_D);
        memcpy(&adab[(size_t)b * 3 * FN_D], v.data(), sizeof(float) * v.size());
    }
    GETF("flow_net.final_layer.adaLN_modulation.1.weight", 2 * FN_D * FN_D);
    memcpy(&adaw[(size_t)FN_DEPTH * 3 * FN_D * FN_D], v.data(), sizeof(float) * v.size());
    GETF("flow_net.final_layer.adaLN_modulation.1.bias", 2 * FN_D);
    memcpy(&adab[(size_t)FN_DEPTH * 3 * FN_D], v.data(), sizeof(float) * v.size());
    if (e->up_bf16(&e->ada_w, adaw) != 0 || e->up_f32(&e->ada_b, adab) != 0) return -1;
    GETF("flow_net.final_layer.linear.weight", FL_LAT * FN_D); if (e->up_bf16(&e->fin_w, v) != 0) return -1;
    GETF("flow_net.final_layer.linear.bias", FL_LAT); if (e->up_f32(&e->fin_b, v) != 0) return -1;
#undef GETF
    return 0;
}

static int ensure_rope(Engine *e, int len) {
    if (len <= e->rope_len) return 0;
    int n = 1024;
    while (n < len) n *= 2;
    e->dfree(e->rope_cos); e->dfree(e->rope_sin);
    if (e->dalloc(&e->rope_cos, (size_t)n * 32) != 0 || e->dalloc(&e->rope_sin, (size_t)n * 32) != 0) return -1;
    launch_k(k_rope_table, (n * 32 + 255) / 256, 256, 0, e->st, e->rope_cos, e->rope_sin, n);
    if (launch_check("k_rope_table") != 0) return -1;
    e->rope_len = n;
    e->invalidate_graphs();
    return 0;
}

// cbias[k] = cond_embed.bias + (time_embed0(k/N) + time_embed1((k+1)/N)) / 2: input-independent
// (ptts_flowlm.c:572-586, 715-719, 768-769), computed once per lsd_steps value.
static int ensure_cbias(Engine *e, int steps, float **out) {
    auto it = e->cbias.find(steps);
    if (it != e->cbias.end()) { *out = it->second; return 0; }
    float *buf = nullptr;
    if (e->dalloc(&buf, (size_t)steps * FN_D) != 0) return -1;
    for (int k = 0; k < steps; k++) {
        PTTS_CUDA_OK(cudaMemcpyAsync(buf + (size_t)k * FN_D, e->cond_b, FN_D * sizeof(float), cudaMemcpyDeviceToDevice, e->st));
        const float s = (float)k / (float)steps, t = (float)(k + 1) / (float)steps;
        launch_k(k_time_embed, 1, 512, 0, e->st, e->te[0].l0w, e->te[0].l0b, e->te[0].l2w, e->te[0].l2b, e->te[0].alpha, e->te[0].freqs, s,
                                          buf + (size_t)k * FN_D);
        launch_k(k_time_embed, 1, 512, 0, e->st, e->te[1].l0w, e->te[1].l0b, e->te[1].l2w, e->te[1].l2b, e->te[1].alpha, e->te[1].freqs, t,
                                          buf + (size_t)k * FN_D);
        if (launch_check("k_time_embed") != 0) return -1;
    }
    e->cbias[steps] = buf;
    *out = buf;
    return 0;
}

// (Re)allocate the run state for batch B (padded), own positions S, tokens, frames, voice length.
static int flow_ensure_capacity(Engine *e, int B, int S, int tok, int F, int cond) {
    if (B <= e->capB && S <= e->capS && tok <= e->capTok && F <= e->capF && cond <= e->capCond) return 0;
    cudaStreamSynchronize(e->st);
    e->invalidate_graphs();
    const bool regrow_prefix = cond > e->capCond;
    // capacities grow in steps (32 cache rows, 16 tokens, 8 frames): a sweep over batches whose longest prompt differs by a few
    // tokens must not re-allocate the caches and rebuild the step program on every call
    S = (S + 31) / 32 * 32; tok = (tok + 15) / 16 * 16; F = (F + 7) / 8 * 8;
    B = std::max(B, e->capB); S = std::max(S, e->capS); tok = std::max(tok, e->capTok); F = std::max(F, e->capF);
    cond = std::max(cond, e->capCond);
    void **bufs[] = {(void **)&e->d_tokens, (void **)&e->d_tok_len, (void **)&e->d_max_frames, (void **)&e->d_eos_after,
                     (void **)&e->d_noise, (void **)&e->d_latent, (void **)&e->d_out_latents, (void **)&e->d_eos_step,
                     (void **)&e->d_done, (void **)&e->d_used, (void **)&e->d_first_eos, (void **)&e->d_first_cond,
                     (void **)&e->d_first_flow, (void **)&e->x, (void **)&e->qkv, (void **)&e->att, (void **)&e->ffh,
                     (void **)&e->normed, (void **)&e->xn_h, (void **)&e->xn_l, (void **)&e->att_h, (void **)&e->att_l,
                     (void **)&e->ffh_h, (void **)&e->ffh_l, (void **)&e->normed_h, (void **)&e->normed_l, (void **)&e->partA,
                     (void **)&e->partB, (void **)&e->fx, (void **)&e->ada, (void **)&e->ft, (void **)&e->flow, (void **)&e->sc,
                     (void **)&e->fg, (void **)&e->sc_h, (void **)&e->sc_l, (void **)&e->fh_h, (void **)&e->fh_l,
                     (void **)&e->fg_h, (void **)&e->fg_l, (void **)&e->ok, (void **)&e->ov, (void **)&e->d_own_base};
    // the recorded capacities drop to zero BEFORE anything is freed: if an allocation below fails, the next call does not
    // pass the "fits" test above with null buffers but comes back here and re-allocates everything
    e->capB = e->capS = e->capTok = e->capF = 0;
    for (void **p : bufs) { e->dfree(*p); *p = nullptr; }
    const size_t b = (size_t)B;
#define DA(ptr, count) if (e->dalloc(&e->ptr, (count)) != 0) { e->voice_len = -1; return -1; }
    DA(d_own_base, b); DA(d_tokens, b * tok); DA(d_tok_len, b); DA(d_max_frames, b); DA(d_eos_after, b);
    DA(d_noise, b * F * FL_LAT); DA(d_latent, b * FL_LAT); DA(d_out_latents, b * F * FL_LAT);
    DA(d_eos_step, b); DA(d_done, b); DA(d_used, b); DA(d_first_eos, b); DA(d_first_cond, b * FL_D); DA(d_first_flow, b * FL_LAT);
    DA(x, b * FL_D); DA(qkv, 8 * b * 3 * FL_D); DA(att, b * FL_D); DA(ffh, b * FL_FF); DA(normed, b * FL_D);
    DA(xn_h, b * FL_D); DA(xn_l, b * FL_D); DA(att_h, b * FL_D); DA(att_l, b * FL_D); DA(ffh_h, b * FL_FF); DA(ffh_l, b * FL_FF);
    DA(normed_h, b * FL_D); DA(normed_l, b * FL_D); DA(partA, 8 * b * FL_D); DA(partB, 8 * b * FL_D);
    DA(fx, b * FN_D); DA(ada, b * FN_ADA); DA(ft, 8 * b * FN_D); DA(flow, b * FL_LAT); DA(sc, b * FN_D); DA(fg, b * FN_D);
    DA(sc_h, b * FN_D); DA(sc_l, b * FN_D); DA(fh_h, b * FN_D); DA(fh_l, b * FN_D); DA(fg_h, b * FN_D); DA(fg_l, b * FN_D);
    DA(ok, (size_t)FL_L * b * FL_D * S); DA(ov, (size_t)FL_L * b * FL_D * S);
    if (regrow_prefix || !e->pk) {
        e->capCond = 0; e->voice_len = -1;
        e->dfree(e->pk); e->dfree(e->pv); e->dfree(e->d_cond); e->dfree(e->pk16); e->dfree(e->pv16);
        e->pk = e->pv = e->d_cond = nullptr; e->pk16 = e->pv16 = nullptr;
        DA(pk, (size_t)FL_L * FL_D * cond + 16); DA(pv, (size_t)FL_L * FL_D * cond + 16); DA(d_cond, (size_t)cond * FL_D + 16);
        DA(pk16, (size_t)FL_L * FL_D * cond + 16); DA(pv16, (size_t)FL_L * FL_D * cond + 16);
        e->voice_len = -1;
    }
    if (!e->ctrl) DA(ctrl, 1);
#undef DA
    // zero the 16-bit operand buffers once: padded batch rows must hold finite values
    PTTS_CUDA_OK(cudaMemsetAsync(e->x, 0, b * FL_D * 4, e->st));
    PTTS_CUDA_OK(cudaMemsetAsync(e->d_latent, 0, b * FL_LAT * 4, e->st));
    PTTS_CUDA_OK(cudaMemsetAsync(e->d_noise, 0, b * F * FL_LAT * 4, e->st));
    e->capB = B; e->capS = S; e->capTok = tok; e->capF = F; e->capCond = cond;
    return 0;
}

static FlowState make_state(Engine *e, int B, int lsd_steps) {
    FlowState s = {};
    s.step = &e->ctrl->step;
    s.B = B; s.max_tok = e->capTok; s.max_frames_cap = e->capF; s.cond_len = 0;
    s.tokens = e->d_tokens; s.tok_len = e->d_tok_len; s.max_frames = e->d_max_frames;
    s.noise = e->d_noise; s.latent = e->d_latent; s.out_latents = e->d_out_latents;
    s.eos_step = e->d_eos_step; s.done = e->d_done; s.used = e->d_used;
    s.first_eos = e->d_first_eos; s.first_cond = e->d_first_cond; s.first_flow = e->d_first_flow;
    s.eos_log = nullptr;
    s.lsd_steps = lsd_steps;
    return s;
}

// ---------------------------------------------------------------- batch-1 GEMV path
template <int K, int RPC, int PRO>
static int gemv(Engine *e, const __nv_bfloat16 *W, int N, const float *x, GemvParams p) {
    p.W = W; p.N = N; p.x = x;
    constexpr int THREADS = K / 8 < 128 ? K / 8 : 128;
    if (e->prof.on) { e->prof.bytes[KC_GEMV] += 2.0 * N * K + 4.0 * (K + N); e->prof.flops[KC_GEMV] += 2.0 * N * K; }
    LaunchScope ls(e, KC_GEMV);
    launch_k_window(k_gemv<K, RPC, PRO>, (N + RPC - 1) / RPC, THREADS, 0, e->st, (const void *)W, std::min((size_t)N * K * 2, e->l2_max_window),
                    e->l2_hit_ratio, p);
    return launch_check("k_gemv");
}

static int attn_decode(Engine *e, int B, int layer, bool out16, int qkv_nsplit = 1) {
    AttnDecodeParams a = {};
    a.qkv = e->qkv; a.qkv_nsplit = qkv_nsplit; a.qkv_stride = (size_t)e->capB * 3 * FL_D;
    a.rope_cos = e->rope_cos; a.rope_sin = e->rope_sin;
    a.ctrl_cond_len = &e->ctrl->cond_len;
    a.own_base = e->d_own_base;
    a.pstride = e->capCond;
    a.pk = e->pk + (size_t)layer * FL_D * e->capCond; a.pv = e->pv + (size_t)layer * FL_D * e->capCond;
    a.ok = e->ok + (size_t)layer * e->capB * FL_D * e->capS; a.ov = e->ov + (size_t)layer * e->capB * FL_D * e->capS;
    a.max_s = e->capS; a.cap_b = e->capB;
    a.step = &e->ctrl->step;
    if (out16) { a.out_h = e->att_h; a.out_l = e->att_l; } else a.out_f32 = e->att;
    const int pcap = ((e->capCond + e->capS + 63) / 32) * 32;
    const int G = B >= 128 ? 8 : (B >= 64 ? 4 : (B >= 32 ? 2 : 1));
    const size_t smem = attn2_smem_bytes(pcap, G);
    PTTS_CHECK(smem <= 200 * 1024, "attention score buffer of %zu bytes exceeds shared memory (sequence too long)", smem);
    if (e->prof.on) {
        // algorithmic bytes: shared prefix K/V once, each utterance's own K/V rows once, the new K/V
        // row written, q/k/v read and the output written (own rows: mean prompt length + frames so far)
        const double own = e->prof.mean_own_base + e->prof.cur_step + 1, pre = e->voice_len;
        e->prof.bytes[KC_ATTN_DECODE] += 2.0 * FL_D * (4.0 * pre + 2.0 * (double)B * own) + (double)B * FL_D * (2.0 * 2 + 4.0 * (3 + 1));
        e->prof.flops[KC_ATTN_DECODE] += 4.0 * FL_D * (double)B * (pre + own);
    }
    LaunchScope ls(e, KC_ATTN_DECODE);
    const dim3 grid(FL_H, (B + G - 1) / G);
    if (G == 8) launch_k(k_attn_decode2<1>, grid, 256, smem, e->st, a, B, pcap);
    else if (G == 4) launch_k(k_attn_decode2<2>, grid, 256, smem, e->st, a, B, pcap);
    else if (G == 2) launch_k(k_attn_decode2<4>, grid, 256, smem, e->st, a, B, pcap);
    else launch_k(k_attn_decode2<8>, grid, 256, smem, e->st, a, B, pcap);
    return launch_check("k_attn_decode2");
}

static int rownorm(Engine *e, int rows, RowNormParams p, int cls = KC_ROWNORM) {
    LaunchScope ls(e, cls);
    launch_k(k_rownorm, rows, 256, 0, e->st, p);
    return launch_check("k_rownorm");
}

// One decode iteration, batch-1 path: head (frame j) then transformer body for frame j+1.
static int flow_step_gemv(Engine *e, const FlowState &st, const float *cbias) {
    {   // out_norm + EOS rule + noise -> latent + flow input projection of LSD step 0: one node instead of three
        LaunchScope ls(e, KC_FLOW_MISC);
        launch_k(k_b1_head, 1, 256, 0, e->st, st, e->ctrl, (const float *)e->x, (const float *)e->onorm_w, (const float *)e->onorm_b,
                 (const float *)e->eos_w, (const float *)e->eos_b, (const __nv_bfloat16 *)e->inp_w, (const float *)e->inp_b, e->normed, e->fx);
        if (launch_check("k_b1_head") != 0) return -1;
    }
    for (int k = 0; k < st.lsd_steps; k++) {
        if (k > 0) {
            LaunchScope ls(e, KC_FLOW_MISC);
            launch_k(k_flow_init, 1, 512, 0, e->st, st, k, e->inp_w, e->inp_b, e->fx);
            if (launch_check("k_flow_init") != 0) return -1;
        }
        GemvParams g = {};
        g.bias = cbias + (size_t)k * FN_D; g.act = ACT_SILU; g.out = e->sc;      // silu(c), c = cond_embed + time
        if (gemv<FL_D, 4, 0>(e, e->cond_w, FN_D, e->normed, g) != 0) return -1;
        g = {}; g.bias = e->ada_b; g.out = e->ada;
        if (gemv<FN_D, 8, 0>(e, e->ada_w, FN_ADA, e->sc, g) != 0) return -1;
        for (int b = 0; b < FN_DEPTH; b++) {
            const float *ad = e->ada + (size_t)b * 3 * FN_D;
            g = {}; g.ln_w = e->FB[b].lnw; g.ln_b = e->FB[b].lnb; g.eps = 1e-6f; g.shift = ad; g.scale = ad + FN_D;
            g.bias = e->FB[b].m0b; g.act = ACT_SILU; g.out = e->fg;
            if (gemv<FN_D, 4, 2>(e, e->FB[b].m0w, FN_D, e->fx, g) != 0) return -1;
            g = {}; g.bias = e->FB[b].m2b; g.gate = ad + 2 * FN_D; g.acc_out = e->fx;
            if (gemv<FN_D, 4, 0>(e, e->FB[b].m2w, FN_D, e->fg, g) != 0) return -1;
        }
        const float *ad = e->ada + (size_t)FN_DEPTH * 3 * FN_D;
        g = {}; g.eps = 1e-6f; g.shift = ad; g.scale = ad + FN_D; g.bias = e->fin_b; g.out = e->flow;
        if (gemv<FN_D, 4, 2>(e, e->fin_w, FL_LAT, e->fx, g) != 0) return -1;
        if (k + 1 < st.lsd_steps) {
            LaunchScope ls(e, KC_FLOW_MISC);
            launch_k(k_flow_update, 1, 32, 0, e->st, st, e->ctrl, e->d_eos_after, k, e->flow);
            if (launch_check("k_flow_update") != 0) return -1;
        }
    }
    {   // last Euler update + frame record + stop rule + x = input_linear . latent: one node instead of two
        LaunchScope ls(e, KC_FLOW_MISC);
        launch_k(k_b1_tail, 1, 256, 0, e->st, st, (const int *)e->d_eos_after, (const float *)e->flow, (const __nv_bfloat16 *)e->in_lin, e->x);
        if (launch_check("k_b1_tail") != 0) return -1;
    }
    for (int l = 0; l < FL_L; l++) {
        GemvParams g = {};
        g.ln_w = e->L[l].n1w; g.ln_b = e->L[l].n1b; g.eps = 1e-5f; g.out = e->qkv;
        if (gemv<FL_D, 8, 1>(e, e->L[l].inproj, 3 * FL_D, e->x, g) != 0) return -1;
        if (attn_decode(e, 1, l, false) != 0) return -1;
        g = {}; g.acc_out = e->x;
        if (gemv<FL_D, 4, 0>(e, e->L[l].outproj, FL_D, e->att, g) != 0) return -1;
        g = {}; g.ln_w = e->L[l].n2w; g.ln_b = e->L[l].n2b; g.eps = 1e-5f; g.act = ACT_GELU_ERF; g.out = e->ffh;
        if (gemv<FL_D, 8, 1>(e, e->L[l].l1, FL_FF, e->x, g) != 0) return -1;
        g = {}; g.acc_out = e->x;
        if (l == FL_L - 1) g.inc = &e->ctrl->step;   // the frame counter moves with the last kernel of the step (every reader of this
                                                     // step's value - the attention kernels above - has finished by then)
        if (gemv<FL_FF, 4, 0>(e, e->L[l].l2, FL_D, e->ffh, g) != 0) return -1;
    }
    return 0;
}

// ---------------------------------------------------------------- batched tcgen05 path
struct FlowGemms {
    GemmOp qkv[FL_L], outp[FL_L], ff1[FL_L], ff2[FL_L];
    std::vector<GemmOp> cond;   // per LSD step (bias differs)
    GemmOp ada, m0[FN_DEPTH], m2[FN_DEPTH], fin;
    int B = 0, precise = -1, lsd = 0;
    uint64_t gen = 0;
};

constexpr int FLOW_MAX_SPLIT = 8;   // split-K slices a consumer kernel folds (partA / partB / ft / qkv are sized for it)

// Decode-step GEMM: activations [B, K] (bf16 hi + lo, the A operand, M = utterances) x weights
// [n_w, K] (the B operand), output rows = utterances. The weights are streamed once per M tile; K is
// split across CTAs until the grid covers the SMs, and the consumer kernel folds the partial sums.
static int flow_gemm_build(Engine *e, GemmOp &op, const __nv_bfloat16 *W, int n_w, int K, const uint16_t *xh, const uint16_t *xl, int B,
                           int bn, int max_split, const float *bias, float *out_f32, int ldo, size_t split_stride, uint16_t *oh,
                           uint16_t *ol, int ldh, int act) {
    GemmDesc d;
    d.A = xh; d.Alo = xl; d.a_k = K; d.a_rows = B; d.a_batches = 1; d.a_ld = K;
    d.B = W; d.b_k = K; d.b_rows = n_w; d.b_ld = K;
    d.bf16 = true; d.taps = 1; d.k_per_tap = K; d.num_sms = e->num_sms;
    d.bn = std::min(bn, std::max(32, n_w));
    const int tiles = ((B + GEMM_BM - 1) / GEMM_BM) * ((n_w + d.bn - 1) / d.bn);
    d.split_k = std::max(1, std::min(std::min(max_split, e->num_sms / tiles), K / (2 * GEMM_BK)));
    d.epi.transposed = 0; d.epi.bias = bias; d.epi.out_f32 = out_f32; d.epi.ldo = ldo; d.epi.split_stride = split_stride;
    d.epi.out_h = oh; d.epi.out_l = ol; d.epi.ldh = ldh; d.epi.out_bf16 = 1; d.epi.act = act;
    return gemm_build(op, d);
}

static int flow_gemms_build(Engine *e, FlowGemms &G, int B, int precise, int lsd, const float *cbias) {
    const size_t ps = (size_t)e->capB * FL_D;
    const uint16_t *xl = precise ? e->xn_l : nullptr, *al = precise ? e->att_l : nullptr, *fl = precise ? e->ffh_l : nullptr;
    for (int l = 0; l < FL_L; l++) {
        if (flow_gemm_build(e, G.qkv[l], e->L[l].inproj, 3 * FL_D, FL_D, e->xn_h, xl, B, 128, FLOW_MAX_SPLIT, nullptr, e->qkv, 3 * FL_D,
                            (size_t)e->capB * 3 * FL_D, nullptr, nullptr, 0, ACT_NONE) != 0) return -1;
        if (flow_gemm_build(e, G.outp[l], e->L[l].outproj, FL_D, FL_D, e->att_h, al, B, 128, FLOW_MAX_SPLIT, nullptr, e->partA, FL_D, ps,
                            nullptr, nullptr, 0, ACT_NONE) != 0) return -1;
        if (flow_gemm_build(e, G.ff1[l], e->L[l].l1, FL_FF, FL_D, e->xn_h, xl, B, 64, 1, nullptr, nullptr, 0, 0, e->ffh_h,
                            precise ? e->ffh_l : nullptr, FL_FF, ACT_GELU_ERF) != 0) return -1;
        if (flow_gemm_build(e, G.ff2[l], e->L[l].l2, FL_D, FL_FF, e->ffh_h, fl, B, 128, FLOW_MAX_SPLIT, nullptr, e->partB, FL_D, ps, nullptr,
                            nullptr, 0, ACT_NONE) != 0) return -1;
    }
    G.cond.resize(lsd);
    for (int k = 0; k < lsd; k++)
        if (flow_gemm_build(e, G.cond[k], e->cond_w, FN_D, FL_D, e->normed_h, precise ? e->normed_l : nullptr, B, 32, 1,
                            cbias + (size_t)k * FN_D, nullptr, 0, 0, e->sc_h, precise ? e->sc_l : nullptr, FN_D, ACT_SILU) != 0)
            return -1;
    if (flow_gemm_build(e, G.ada, e->ada_w, FN_ADA, FN_D, e->sc_h, precise ? e->sc_l : nullptr, B, 128, 1, e->ada_b, e->ada, FN_ADA, 0,
                        nullptr, nullptr, 0, ACT_NONE) != 0) return -1;
    for (int b = 0; b < FN_DEPTH; b++) {
        if (flow_gemm_build(e, G.m0[b], e->FB[b].m0w, FN_D, FN_D, e->fh_h, precise ? e->fh_l : nullptr, B, 32, 1, e->FB[b].m0b, nullptr, 0,
                            0, e->fg_h, precise ? e->fg_l : nullptr, FN_D, ACT_SILU) != 0) return -1;
        if (flow_gemm_build(e, G.m2[b], e->FB[b].m2w, FN_D, FN_D, e->fg_h, precise ? e->fg_l : nullptr, B, 64, FLOW_MAX_SPLIT, e->FB[b].m2b,
                            e->ft, FN_D, (size_t)e->capB * FN_D, nullptr, nullptr, 0, ACT_NONE) != 0) return -1;
    }
    if (flow_gemm_build(e, G.fin, e->fin_w, FL_LAT, FN_D, e->fh_h, precise ? e->fh_l : nullptr, B, 32, 1, e->fin_b, e->flow, FL_LAT, 0,
                        nullptr, nullptr, 0, ACT_NONE) != 0) return -1;
    G.B = B; G.precise = precise; G.lsd = lsd; G.gen = e->gen;
    return 0;
}

// One decode iteration, batched tcgen05 path: head (frame j) then transformer body for frame j+1.
// The split-K partials of the last FFN2 of iteration j-1 (zeroed before the first iteration) are
// folded into x by the out_norm row kernel.
static int flow_step_tc(Engine *e, const FlowState &st, const FlowGemms &G, int precise) {
    const int B = st.B;
    const size_t ps = (size_t)e->capB * FL_D;
    {
        RowNormParams r = {};
        r.x = e->x; r.D = FL_D; r.w = e->onorm_w; r.b = e->onorm_b; r.eps = 1e-5f; r.do_norm = 1;
        r.part = e->partB; r.nsplit = (int)G.ff2[FL_L - 1].p.split_k; r.split_stride = ps;
        r.out_f32 = e->normed; r.out_h = e->normed_h; r.out_l = precise ? e->normed_l : nullptr; r.out_bf16 = 1;
        if (rownorm(e, B, r) != 0) return -1;
        LaunchScope ls(e, KC_FLOW_MISC);
        launch_k(k_eos_state, B, 256, 0, e->st, st, e->ctrl, e->d_eos_after, e->normed, e->eos_w, e->eos_b);
        if (launch_check("k_eos_state") != 0) return -1;
    }
    for (int k = 0; k < st.lsd_steps; k++) {
        {
            LaunchScope ls(e, KC_FLOW_MISC);
            launch_k(k_flow_init, B, 512, 0, e->st, st, k, e->inp_w, e->inp_b, e->fx);
            if (launch_check("k_flow_init") != 0) return -1;
        }
        if (run_gemm(e, G.cond[k], KC_GEMM_FLOW) != 0) return -1;
        if (run_gemm(e, G.ada, KC_GEMM_FLOW) != 0) return -1;
        for (int b = 0; b <= FN_DEPTH; b++) {
            // fold the previous block's gated MLP output, then LN + adaLN for this block / the final layer
            RowNormParams r = {};
            r.x = e->fx; r.D = FN_D; r.eps = 1e-6f; r.do_norm = 1; r.ldmod = FN_ADA;
            if (b > 0) { r.part = e->ft; r.nsplit = G.m2[b - 1].p.split_k; r.split_stride = (size_t)e->capB * FN_D; r.gate = e->ada + (size_t)(b - 1) * 3 * FN_D + 2 * FN_D; }
            if (b < FN_DEPTH) { r.w = e->FB[b].lnw; r.b = e->FB[b].lnb; }
            r.shift = e->ada + (size_t)b * 3 * FN_D; r.scale = r.shift + FN_D;
            r.out_h = e->fh_h; r.out_l = precise ? e->fh_l : nullptr; r.out_bf16 = 1;
            if (rownorm(e, B, r) != 0) return -1;
            if (b < FN_DEPTH) {
                if (run_gemm(e, G.m0[b], KC_GEMM_FLOW) != 0) return -1;
                if (run_gemm(e, G.m2[b], KC_GEMM_FLOW) != 0) return -1;
            }
        }
        if (run_gemm(e, G.fin, KC_GEMM_FLOW) != 0) return -1;
        LaunchScope ls(e, KC_FLOW_MISC);
        launch_k(k_flow_update, B, 32, 0, e->st, st, e->ctrl, e->d_eos_after, k, e->flow);
        if (launch_check("k_flow_update") != 0) return -1;
    }
    {
        LaunchScope ls(e, KC_FLOW_MISC);
        launch_k(k_prepare_x, B, 256, 0, e->st, st, e->in_lin, e->x);
        if (launch_check("k_prepare_x") != 0) return -1;
    }
    for (int l = 0; l < FL_L; l++) {
        RowNormParams r = {};
        r.x = e->x; r.D = FL_D; r.w = e->L[l].n1w; r.b = e->L[l].n1b; r.eps = 1e-5f; r.do_norm = 1;
        r.out_h = e->xn_h; r.out_l = precise ? e->xn_l : nullptr; r.out_bf16 = 1;
        if (l > 0) { r.part = e->partB; r.nsplit = (int)G.ff2[l - 1].p.split_k; r.split_stride = ps; }
        if (rownorm(e, B, r) != 0) return -1;
        if (run_gemm(e, G.qkv[l], KC_GEMM_FLOW) != 0) return -1;
        if (attn_decode(e, B, l, true, G.qkv[l].p.split_k) != 0) return -1;
        if (run_gemm(e, G.outp[l], KC_GEMM_FLOW) != 0) return -1;
        r = {};
        r.x = e->x; r.D = FL_D; r.w = e->L[l].n2w; r.b = e->L[l].n2b; r.eps = 1e-5f; r.do_norm = 1;
        r.part = e->partA; r.nsplit = (int)G.outp[l].p.split_k; r.split_stride = ps;
        r.out_h = e->xn_h; r.out_l = precise ? e->xn_l : nullptr; r.out_bf16 = 1;
        if (rownorm(e, B, r) != 0) return -1;
        if (run_gemm(e, G.ff1[l], KC_GEMM_FLOW) != 0) return -1;
        if (run_gemm(e, G.ff2[l], KC_GEMM_FLOW) != 0) return -1;
    }
    LaunchScope ls(e, KC_FLOW_MISC);
    launch_k(k_advance_step, 1, 1, 0, e->st, &e->ctrl->step);
    return launch_check("k_advance_step");
}

// ---------------------------------------------------------------- batched prefill
// Grow the packed-row workspace.
static int prefill_ensure(Engine *e, int rows, int n) {
    if (rows <= e->capRows && e->d_row_start && e->pf_ffh_l) return 0;
    cudaStreamSynchronize(e->st);
    rows = std::max((rows + 2047) / 2048 * 2048, e->capRows);
    e->capRows = 0;   // a failed grow must not leave a capacity that passes the test above with null buffers
    void **bufs[] = {(void **)&e->d_row_b, (void **)&e->d_row_i, (void **)&e->d_row_start, (void **)&e->pf_x, (void **)&e->pf_qkv,
                     (void **)&e->pf_xn_h, (void **)&e->pf_xn_l, (void **)&e->pf_att_h, (void **)&e->pf_att_l,
                     (void **)&e->pf_ffh_h, (void **)&e->pf_ffh_l};
    for (void **p : bufs) { e->dfree(*p); *p = nullptr; }
    const size_t r = (size_t)rows;
#define DA(ptr, count) if (e->dalloc(&e->ptr, (count)) != 0) return -1
    DA(d_row_b, r); DA(d_row_i, r); DA(d_row_start, 4096); DA(pf_x, r * FL_D); DA(pf_qkv, r * 3 * FL_D);
    DA(pf_xn_h, r * FL_D); DA(pf_xn_l, r * FL_D); DA(pf_att_h, r * FL_D); DA(pf_att_l, r * FL_D);
    DA(pf_ffh_h, r * FL_FF); DA(pf_ffh_l, r * FL_FF);
#undef DA
    (void)n;
    e->capRows = rows;
    return 0;
}

// activations [M, K] (bf16 hi + lo, the A operand) x weights [N, K] (B operand), normal orientation
static int prefill_gemm(Engine *e, const uint16_t *ah, const uint16_t *al, int M, int K, const __nv_bfloat16 *W, int N,
                        GemmParams epi) {
    GemmDesc d;
    d.A = ah; d.Alo = al; d.a_k = K; d.a_rows = M; d.a_batches = 1; d.a_ld = K;
    d.B = W; d.b_k = K; d.b_rows = N; d.b_ld = K;
    d.bf16 = true; d.taps = 1; d.k_per_tap = K; d.num_sms = e->num_sms;
    d.bn = M <= 256 ? 64 : (M <= 1024 ? 128 : 256);   // short prompts: more, narrower tiles to spread the weight stream
    d.epi = epi; d.epi.out_bf16 = 1;
    GemmOp op;
    if (gemm_build(op, d) != 0) return -1;
    return run_gemm(e, op, KC_GEMM_FLOW);
}

// Run the transformer over M packed rows. voice_mode: rows are the voice frames (d_cond), K/V go to
// the prefix buffers, no prefix is attended. Otherwise rows are text tokens + BOS of n utterances,
// K/V go to the own caches and x[b] receives the BOS-position hidden state.
static int flow_prefill(Engine *e, int n, int M, int max_q, int cond_len, bool voice_mode) {
    PrefillRows pr = {e->d_row_b, e->d_row_i, e->d_row_start};
    const size_t smem = pf_attn_smem_bytes(cond_len, max_q);
    PTTS_CHECK(smem <= 200 * 1024, "prefill attention needs %zu bytes of shared memory (voice prefix %d + prompt %d rows too long)",
               smem, cond_len, max_q);
    {
        LaunchScope ls(e, KC_FLOW_MISC);
        launch_k(k_pf_gather, M, 256, 0, e->st, pr, e->d_tokens, e->capTok, e->d_tok_len, e->embed, e->bos_proj,
                                          voice_mode ? e->d_cond : nullptr, e->pf_x);
        if (launch_check("k_pf_gather") != 0) return -1;
    }
    for (int l = 0; l < FL_L; l++) {
        RowNormParams r = {};
        r.x = e->pf_x; r.D = FL_D; r.w = e->L[l].n1w; r.b = e->L[l].n1b; r.eps = 1e-5f; r.do_norm = 1;
        r.out_h = e->pf_xn_h; r.out_l = e->pf_xn_l; r.out_bf16 = 1;
        if (rownorm(e, M, r) != 0) return -1;
        GemmParams g = {};
        g.out_f32 = e->pf_qkv; g.ldo = 3 * FL_D;
        if (prefill_gemm(e, e->pf_xn_h, e->pf_xn_l, M, FL_D, e->L[l].inproj, 3 * FL_D, g) != 0) return -1;
        {
            PrefillAttnParams a = {};
            a.qkv = e->pf_qkv; a.rope_cos = e->rope_cos; a.rope_sin = e->rope_sin;
            a.cond_len = cond_len; a.pstride = e->capCond; a.max_q = max_q;
            if (voice_mode) {
                a.ok = e->pk + (size_t)l * FL_D * e->capCond; a.ov = e->pv + (size_t)l * FL_D * e->capCond; a.max_s = e->capCond; a.cap_b = 1;
                a.own_f16 = 0;
            } else {
                a.pk = e->pk + (size_t)l * FL_D * e->capCond; a.pv = e->pv + (size_t)l * FL_D * e->capCond;
                a.ok = e->ok + (size_t)l * e->capB * FL_D * e->capS; a.ov = e->ov + (size_t)l * e->capB * FL_D * e->capS;
                a.max_s = e->capS; a.cap_b = e->capB; a.own_f16 = 1;
            }
            a.out_h = e->pf_att_h; a.out_l = e->pf_att_l;
            LaunchScope ls(e, KC_ATTN_DECODE);
            // prompts: tensor-core kernel over [fp16 prefix copy | own rows]; the voice pass (once per voice, fp32 prefix cache
            // as its product) keeps the fp32 scalar kernel. PTTS_B200_PF_ATTN_SCALAR=1 selects the scalar kernel for A/B runs.
            static const bool pf_scalar = getenv("PTTS_B200_PF_ATTN_SCALAR") != nullptr;
            if (voice_mode || pf_scalar) launch_k(k_pf_attn, dim3(FL_H, n), 256, smem, e->st, a, pr);
            else launch_k(k_pf_attn_tc, dim3((max_q + 63) / 64, FL_H, n), 128, 0, e->st, a, pr,
                          (const __half *)(e->pk16 + (size_t)l * FL_D * e->capCond), (const __half *)(e->pv16 + (size_t)l * FL_D * e->capCond));
            if (launch_check("k_pf_attn") != 0) return -1;
        }
        g = {}; g.resid = e->pf_x; g.out_f32 = e->pf_x; g.ldo = FL_D;
        if (prefill_gemm(e, e->pf_att_h, e->pf_att_l, M, FL_D, e->L[l].outproj, FL_D, g) != 0) return -1;
        r = {};
        r.x = e->pf_x; r.D = FL_D; r.w = e->L[l].n2w; r.b = e->L[l].n2b; r.eps = 1e-5f; r.do_norm = 1;
        r.out_h = e->pf_xn_h; r.out_l = e->pf_xn_l; r.out_bf16 = 1;
        if (rownorm(e, M, r) != 0) return -1;
        g = {}; g.out_h = e->pf_ffh_h; g.out_l = e->pf_ffh_l; g.ldh = FL_FF; g.act = ACT_GELU_ERF;
        if (prefill_gemm(e, e->pf_xn_h, e->pf_xn_l, M, FL_D, e->L[l].l1, FL_FF, g) != 0) return -1;
        g = {}; g.resid = e->pf_x; g.out_f32 = e->pf_x; g.ldo = FL_D;
        if (prefill_gemm(e, e->pf_ffh_h, e->pf_ffh_l, M, FL_FF, e->L[l].l2, FL_D, g) != 0) return -1;
    }
    if (!voice_mode) {
        LaunchScope ls(e, KC_FLOW_MISC);
        launch_k(k_pf_last, n, 256, 0, e->st, pr, e->pf_x, e->x);
        if (launch_check("k_pf_last") != 0) return -1;
    }
    return 0;
}

}  // namespace ptts
