/*
 * ptts_oracle.c - CPU restatement of the Pocket-TTS generation hot path.
 * TEST INFRASTRUCTURE ONLY (see ptts_oracle.h). Parity: pinned against the
 * compiled reference (oracle/_ref) and tests/golden/.
 *
 * This is not the reference's code: Mimi runs time-major ([T,C]) end to end,
 * convolutions are expressed as per-row dot products over contiguous windows,
 * and summation order differs.  The arithmetic definitions are the reference's
 * (cited per function).
 */
#include "ptts_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define D 1024
#define NH 16
#define HD 64
#define NL 6
#define FF 4096
#define LAT 32
#define FD 512
#define FDEPTH 6
#define MD 512
#define MH 8
#define ML 2
#define MFF 2048
#define MCTX 250

static char g_err[256];
const char *orc_last_error(void) { return g_err; }

typedef struct { const float *w, *b; int cin, cout, k, stride; float *wt; /* re-laid taps */ } conv_t;

struct orc_model {
    const float *embed, *emb_std, *emb_mean, *bos, *in_lin, *onorm_w, *onorm_b, *eos_w, *eos_b;
    struct { const float *inproj, *outproj, *n1w, *n1b, *n2w, *n2b, *l1, *l2; } L[NL];
    const float *cond_w, *cond_b, *inp_w, *inp_b;
    struct { const float *l0w, *l0b, *l2w, *l2b, *alpha, *freqs; } te[2];
    struct { const float *lnw, *lnb, *m0w, *m0b, *m2w, *m2b, *adaw, *adab; } rb[FDEPTH];
    const float *fin_w, *fin_b, *fin_adaw, *fin_adab;
    /* mimi */
    const float *quant, *ups;
    struct { const float *inproj, *outproj, *n1w, *n1b, *n2w, *n2b, *l1, *l2, *ls1, *ls2; } M[ML];
    conv_t dec_in, res1[3], res2[3], dec_out;
    conv_t up[3];
    /* voice-cloning encoder (optional tensors; shapes ptts.c:914-944, speaker projection ptts_flowlm.c:874) */
    int has_encoder;
    conv_t enc_in, eres1[3], eres2[3], edown[3], enc_out, down;
    struct { const float *inproj, *outproj, *n1w, *n1b, *n2w, *n2b, *l1, *l2, *ls1, *ls2; } E[ML];
    const float *spk;
};

/* ---------------------------------------------------------------- lookup */
static int ends_with(const char *s, const char *suf) {
    size_t a = strlen(s), b = strlen(suf);
    return a >= b && strcmp(s + a - b, suf) == 0;
}

/* ptts_flowlm.c:107-122 / ptts_mimi.c:77-95 lookup order */
static const float *find(const orc_tensor *t, int n, const char *const *prefixes, const char *name,
                         int64_t numel, int required) {
    char buf[512];
    const orc_tensor *hit = NULL;
    for (int i = 0; i < n && !hit; i++) if (strcmp(t[i].name, name) == 0) hit = &t[i];
    for (int p = 0; prefixes[p] && !hit; p++) {
        snprintf(buf, sizeof buf, "%s%s", prefixes[p], name);
        for (int i = 0; i < n && !hit; i++) if (strcmp(t[i].name, buf) == 0) hit = &t[i];
    }
    for (int i = 0; i < n && !hit; i++) if (ends_with(t[i].name, name)) hit = &t[i];
    if (!hit) {
        if (required) snprintf(g_err, sizeof g_err, "missing tensor %s", name);
        return NULL;
    }
    if (numel >= 0 && hit->numel != numel) {
        snprintf(g_err, sizeof g_err, "size mismatch for %s: got %lld want %lld", hit->name,
                 (long long)hit->numel, (long long)numel);
        return NULL;
    }
    return hit->data;
}

static const char *const FLOW_PFX[] = {"flow_lm.", NULL};
static const char *const MIMI_PFX[] = {"mimi.", "model.", NULL};

#define FGET(dst, nm, ne) do { (dst) = find(t, n, FLOW_PFX, (nm), (ne), 1); if (!(dst)) goto fail; } while (0)
#define MGET(dst, nm, ne) do { (dst) = find(t, n, MIMI_PFX, (nm), (ne), 1); if (!(dst)) goto fail; } while (0)

/* Conv1d weights [out,in,k] -> [out][k][in]; ConvTranspose weights [in,out,k] -> [k][out][in].
 * With time-major activations a causal window is then one contiguous span. */
static int relay_conv(conv_t *c) {
    c->wt = (float *)malloc(sizeof(float) * (size_t)c->cout * c->cin * c->k);
    if (!c->wt) return -1;
    for (int o = 0; o < c->cout; o++)
        for (int i = 0; i < c->cin; i++)
            for (int kk = 0; kk < c->k; kk++)
                c->wt[((size_t)o * c->k + kk) * c->cin + i] = c->w[((size_t)o * c->cin + i) * c->k + kk];
    return 0;
}
static int relay_convtr(conv_t *c) {
    c->wt = (float *)malloc(sizeof(float) * (size_t)c->cout * c->cin * c->k);
    if (!c->wt) return -1;
    for (int i = 0; i < c->cin; i++)
        for (int o = 0; o < c->cout; o++)
            for (int kk = 0; kk < c->k; kk++)
                c->wt[((size_t)kk * c->cout + o) * c->cin + i] = c->w[((size_t)i * c->cout + o) * c->k + kk];
    return 0;
}

orc_model *orc_create(const orc_tensor *t, int n) {
    orc_model *m = (orc_model *)calloc(1, sizeof *m);
    char nm[256];
    if (!m) return NULL;
    g_err[0] = 0;
    FGET(m->embed, "conditioner.embed.weight", 4001 * D);
    FGET(m->emb_std, "emb_std", LAT);
    FGET(m->emb_mean, "emb_mean", LAT);
    FGET(m->bos, "bos_emb", LAT);
    FGET(m->in_lin, "input_linear.weight", D * LAT);
    FGET(m->onorm_w, "out_norm.weight", D);
    FGET(m->onorm_b, "out_norm.bias", D);
    FGET(m->eos_w, "out_eos.weight", D);
    FGET(m->eos_b, "out_eos.bias", 1);
    for (int l = 0; l < NL; l++) {
#define LN_(field, suffix, ne) snprintf(nm, sizeof nm, "transformer.layers.%d." suffix, l); FGET(m->L[l].field, nm, ne)
        LN_(inproj, "self_attn.in_proj.weight", 3 * D * D);
        LN_(outproj, "self_attn.out_proj.weight", D * D);
        LN_(n1w, "norm1.weight", D); LN_(n1b, "norm1.bias", D);
        LN_(n2w, "norm2.weight", D); LN_(n2b, "norm2.bias", D);
        LN_(l1, "linear1.weight", FF * D); LN_(l2, "linear2.weight", D * FF);
#undef LN_
    }
    FGET(m->cond_w, "flow_net.cond_embed.weight", FD * D);
    FGET(m->cond_b, "flow_net.cond_embed.bias", FD);
    FGET(m->inp_w, "flow_net.input_proj.weight", FD * LAT);
    FGET(m->inp_b, "flow_net.input_proj.bias", FD);
    for (int i = 0; i < 2; i++) {
#define TE_(field, suffix, ne) snprintf(nm, sizeof nm, "flow_net.time_embed.%d." suffix, i); FGET(m->te[i].field, nm, ne)
        TE_(l0w, "mlp.0.weight", FD * 256); TE_(l0b, "mlp.0.bias", FD);
        TE_(l2w, "mlp.2.weight", FD * FD); TE_(l2b, "mlp.2.bias", FD);
        TE_(alpha, "mlp.3.alpha", FD);
#undef TE_
        snprintf(nm, sizeof nm, "flow_net.time_embed.%d.freqs", i);
        m->te[i].freqs = find(t, n, FLOW_PFX, nm, 128, 0); /* optional, ptts_flowlm.c:575 */
    }
    for (int b = 0; b < FDEPTH; b++) {
#define RB_(field, suffix, ne) snprintf(nm, sizeof nm, "flow_net.res_blocks.%d." suffix, b); FGET(m->rb[b].field, nm, ne)
        RB_(lnw, "in_ln.weight", FD); RB_(lnb, "in_ln.bias", FD);
        RB_(m0w, "mlp.0.weight", FD * FD); RB_(m0b, "mlp.0.bias", FD);
        RB_(m2w, "mlp.2.weight", FD * FD); RB_(m2b, "mlp.2.bias", FD);
        RB_(adaw, "adaLN_modulation.1.weight", 3 * FD * FD); RB_(adab, "adaLN_modulation.1.bias", 3 * FD);
#undef RB_
    }
    FGET(m->fin_w, "flow_net.final_layer.linear.weight", LAT * FD);
    FGET(m->fin_b, "flow_net.final_layer.linear.bias", LAT);
    FGET(m->fin_adaw, "flow_net.final_layer.adaLN_modulation.1.weight", 2 * FD * FD);
    FGET(m->fin_adab, "flow_net.final_layer.adaLN_modulation.1.bias", 2 * FD);

    MGET(m->quant, "quantizer.output_proj.weight", MD * LAT);
    m->ups = find(t, n, MIMI_PFX, "upsample.convtr.weight", MD * 32, 0);
    if (!m->ups) MGET(m->ups, "upsample.convtr.convtr.weight", MD * 32);
    for (int l = 0; l < ML; l++) {
#define ML_(field, suffix, ne) snprintf(nm, sizeof nm, "decoder_transformer.transformer.layers.%d." suffix, l); MGET(m->M[l].field, nm, ne)
        ML_(inproj, "self_attn.in_proj.weight", 3 * MD * MD);
        ML_(outproj, "self_attn.out_proj.weight", MD * MD);
        ML_(n1w, "norm1.weight", MD); ML_(n1b, "norm1.bias", MD);
        ML_(n2w, "norm2.weight", MD); ML_(n2b, "norm2.bias", MD);
        ML_(l1, "linear1.weight", MFF * MD); ML_(l2, "linear2.weight", MD * MFF);
        ML_(ls1, "layer_scale_1.scale", MD); ML_(ls2, "layer_scale_2.scale", MD);
#undef ML_
    }
    /* hard-coded geometry, ptts_mimi.c:389-486 */
    {
        struct { conv_t *c; const char *nm; int cin, cout, k, s, tr; } G[] = {
            {&m->dec_in, "decoder.model.0.conv", 512, 512, 7, 1, 0},
            {&m->up[0], "decoder.model.2.convtr", 512, 256, 12, 6, 1},
            {&m->res1[0], "decoder.model.3.block.1.conv", 256, 128, 3, 1, 0},
            {&m->res2[0], "decoder.model.3.block.3.conv", 128, 256, 1, 1, 0},
            {&m->up[1], "decoder.model.5.convtr", 256, 128, 10, 5, 1},
            {&m->res1[1], "decoder.model.6.block.1.conv", 128, 64, 3, 1, 0},
            {&m->res2[1], "decoder.model.6.block.3.conv", 64, 128, 1, 1, 0},
            {&m->up[2], "decoder.model.8.convtr", 128, 64, 8, 4, 1},
            {&m->res1[2], "decoder.model.9.block.1.conv", 64, 32, 3, 1, 0},
            {&m->res2[2], "decoder.model.9.block.3.conv", 32, 64, 1, 1, 0},
            {&m->dec_out, "decoder.model.11.conv", 64, 1, 3, 1, 0},
        };
        for (size_t i = 0; i < sizeof G / sizeof G[0]; i++) {
            conv_t *c = G[i].c;
            c->cin = G[i].cin; c->cout = G[i].cout; c->k = G[i].k; c->stride = G[i].s;
            snprintf(nm, sizeof nm, "%s.weight", G[i].nm);
            MGET(c->w, nm, (int64_t)c->cin * c->cout * c->k);
            snprintf(nm, sizeof nm, "%s.bias", G[i].nm);
            MGET(c->b, nm, c->cout);
            if ((G[i].tr ? relay_convtr(c) : relay_conv(c)) != 0) goto fail;
        }
    }
    /* optional: encoder side. Present only when the checkpoint carries the tensors the reference's --verify expects. */
    m->spk = find(t, n, FLOW_PFX, "speaker_proj_weight", D * MD, 0);
    if (m->spk && find(t, n, MIMI_PFX, "encoder.model.0.conv.weight", 64 * 7, 0)) {
        struct { conv_t *c; const char *nm; int cin, cout, k, s, bias; } G[] = {
            {&m->enc_in, "encoder.model.0.conv", 1, 64, 7, 1, 1},
            {&m->eres1[0], "encoder.model.1.block.1.conv", 64, 32, 3, 1, 1},
            {&m->eres2[0], "encoder.model.1.block.3.conv", 32, 64, 1, 1, 1},
            {&m->edown[0], "encoder.model.3.conv", 64, 128, 8, 4, 1},
            {&m->eres1[1], "encoder.model.4.block.1.conv", 128, 64, 3, 1, 1},
            {&m->eres2[1], "encoder.model.4.block.3.conv", 64, 128, 1, 1, 1},
            {&m->edown[1], "encoder.model.6.conv", 128, 256, 10, 5, 1},
            {&m->eres1[2], "encoder.model.7.block.1.conv", 256, 128, 3, 1, 1},
            {&m->eres2[2], "encoder.model.7.block.3.conv", 128, 256, 1, 1, 1},
            {&m->edown[2], "encoder.model.9.conv", 256, 512, 12, 6, 1},
            {&m->enc_out, "encoder.model.11.conv", 512, 512, 3, 1, 1},
            {&m->down, "downsample.conv.conv", 512, 512, 32, 16, 0},
        };
        for (size_t i = 0; i < sizeof G / sizeof G[0]; i++) {
            conv_t *c = G[i].c;
            c->cin = G[i].cin; c->cout = G[i].cout; c->k = G[i].k; c->stride = G[i].s;
            snprintf(nm, sizeof nm, "%s.weight", G[i].nm);
            MGET(c->w, nm, (int64_t)c->cin * c->cout * c->k);
            c->b = NULL;
            if (G[i].bias) { snprintf(nm, sizeof nm, "%s.bias", G[i].nm); MGET(c->b, nm, c->cout); }
        }
        for (int l = 0; l < ML; l++) {
#define EL_(field, suffix, ne) snprintf(nm, sizeof nm, "encoder_transformer.transformer.layers.%d." suffix, l); MGET(m->E[l].field, nm, ne)
            EL_(inproj, "self_attn.in_proj.weight", 3 * MD * MD);
            EL_(outproj, "self_attn.out_proj.weight", MD * MD);
            EL_(n1w, "norm1.weight", MD); EL_(n1b, "norm1.bias", MD);
            EL_(n2w, "norm2.weight", MD); EL_(n2b, "norm2.bias", MD);
            EL_(l1, "linear1.weight", MFF * MD); EL_(l2, "linear2.weight", MD * MFF);
            EL_(ls1, "layer_scale_1.scale", MD); EL_(ls2, "layer_scale_2.scale", MD);
#undef EL_
        }
        m->has_encoder = 1;
    }
    return m;
fail:
    orc_destroy(m);
    return NULL;
}

void orc_destroy(orc_model *m) {
    if (!m) return;
    free(m->dec_in.wt); free(m->dec_out.wt);
    for (int i = 0; i < 3; i++) { free(m->up[i].wt); free(m->res1[i].wt); free(m->res2[i].wt); }
    free(m);
}

/* ---------------------------------------------------------------- primitives */
static inline float dotf(const float *a, const float *b, int n) {
    float s = 0.0f;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* y[n,out] = x[n,in] W[out,in]^T + b   (ptts_kernels.c:95-110) */
void orc_linear(float *y, const float *x, const float *w, const float *b, int n, int in, int out) {
#pragma omp parallel for collapse(2) schedule(static) if ((long)n * in * out > 200000)
    for (int t = 0; t < n; t++)
        for (int o = 0; o < out; o++)
            y[(size_t)t * out + o] = (b ? b[o] : 0.0f) + dotf(x + (size_t)t * in, w + (size_t)o * in, in);
}

/* Reference layouts (channel-major), ptts_kernels.c:128-153: causal, left pad k-stride. */
void orc_conv1d(float *y, const float *x, const float *w, const float *b,
                int in_ch, int out_ch, int T, int k, int stride, int groups) {
    int out_len = T / stride, ipg = in_ch / groups, opg = out_ch / groups, pad = k - stride;
#pragma omp parallel for schedule(static)
    for (int oc = 0; oc < out_ch; oc++) {
        int g = oc / opg;
        for (int t = 0; t < out_len; t++) {
            double s = b ? b[oc] : 0.0f;
            for (int ic = 0; ic < ipg; ic++)
                for (int kk = 0; kk < k; kk++) {
                    int idx = t * stride - pad + kk;
                    if (idx >= 0 && idx < T)
                        s += (double)w[((size_t)oc * ipg + ic) * k + kk] * x[(size_t)(g * ipg + ic) * T + idx];
                }
            y[(size_t)oc * out_len + t] = (float)s;
        }
    }
}

/* ptts_kernels.c:172-204: scatter form, keep the first T*stride outputs. */
void orc_convtr1d(float *y, const float *x, const float *w, const float *b,
                  int in_ch, int out_ch, int T, int k, int stride, int groups) {
    int out_len = (T - 1) * stride + k - (k - stride), ipg = in_ch / groups, opg = out_ch / groups;
    for (int oc = 0; oc < out_ch; oc++) {
        int g = oc / opg, ocg = oc % opg;
        for (int j = 0; j < out_len; j++) {
            double s = b ? b[oc] : 0.0f;
            for (int ic = g * ipg; ic < (g + 1) * ipg; ic++)
                for (int kk = j % stride; kk < k; kk += stride) {
                    int t = (j - kk) / stride;
                    if (t >= 0 && t < T && t * stride + kk == j)
                        s += (double)w[((size_t)ic * opg + ocg) * k + kk] * x[(size_t)ic * T + t];
                }
            y[(size_t)oc * out_len + j] = (float)s;
        }
    }
}

float orc_silu(float x) { return x / (1.0f + expf(-x)); } /* ptts_flowlm.c:199-201 */

/* ptts_flowlm.c:159-181 / ptts_mimi.c:184-206: biased variance, two-pass. */
static void layernorm(const float *x, int d, const float *w, const float *b, float eps, float *y) {
    float mean = 0.0f, var = 0.0f;
    for (int i = 0; i < d; i++) mean += x[i];
    mean /= (float)d;
    for (int i = 0; i < d; i++) { float v = x[i] - mean; var += v * v; }
    var /= (float)d;
    float inv = 1.0f / sqrtf(var + eps);
    for (int i = 0; i < d; i++) {
        float v = (x[i] - mean) * inv;
        if (w) v *= w[i];
        if (b) v += b[i];
        y[i] = v;
    }
}

/* ptts_flowlm.c:183-197: unbiased variance, x itself is not centred. */
static void var_rmsnorm(const float *x, int d, const float *alpha, float eps, float *y) {
    float mean = 0.0f, var = 0.0f;
    for (int i = 0; i < d; i++) mean += x[i];
    mean /= (float)d;
    for (int i = 0; i < d; i++) { float v = x[i] - mean; var += v * v; }
    if (d > 1) var /= (float)(d - 1);
    float inv = 1.0f / sqrtf(var + eps);
    for (int i = 0; i < d; i++) y[i] = x[i] * (alpha ? alpha[i] : 1.0f) * inv;
}

static float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.7071067811865475f)); }  /* ptts_flowlm.c:207-210 */
static float gelu_tanh(float x) { return 0.5f * x * (1.0f + tanhf(0.7978845608f * (x + 0.044715f * x * x * x))); } /* ptts_mimi.c:208-211 */
static float elu(float v) { return v >= 0.0f ? v : (expf(v) - 1.0f); } /* ptts_kernels.c:215-218 */

/* Interleaved-pair rotation at integer position, ptts_flowlm.c:356-388 / ptts_mimi.c:230-262. */
static void rope(float *q, float *k, int heads, int pos) {
    float lmp = logf(10000.0f);
    for (int i = 0; i < HD / 2; i++) {
        float f = expf(-lmp * (2.0f * i / HD));
        float a = f * (float)pos, c = cosf(a), s = sinf(a);
        for (int h = 0; h < heads; h++) {
            float *qq = q + h * HD + 2 * i, *kk = k + h * HD + 2 * i;
            float q0 = qq[0], q1 = qq[1], k0 = kk[0], k1 = kk[1];
            qq[0] = q0 * c - q1 * s; qq[1] = q0 * s + q1 * c;
            kk[0] = k0 * c - k1 * s; kk[1] = k0 * s + k1 * c;
        }
    }
}

/* softmax(q.K/sqrt(64)).V over keys [lo,hi] of a [pos][heads][64] cache
 * (ptts_flowlm.c:502-525, ptts_mimi.c:264-299). */
static void attend(const float *q, const float *K, const float *V, int heads, int lo, int hi,
                   float *sc, float *out) {
    int dm = heads * HD;
    for (int h = 0; h < heads; h++) {
        float mx = -INFINITY;
        for (int t = lo; t <= hi; t++) {
            float s = dotf(q + h * HD, K + (size_t)t * dm + h * HD, HD) / sqrtf((float)HD);
            sc[t - lo] = s;
            if (s > mx) mx = s;
        }
        float sum = 0.0f;
        for (int t = lo; t <= hi; t++) { sc[t - lo] = expf(sc[t - lo] - mx); sum += sc[t - lo]; }
        float inv = 1.0f / sum;
        float *o = out + h * HD;
        for (int d0 = 0; d0 < HD; d0++) o[d0] = 0.0f;
        for (int t = lo; t <= hi; t++) {
            float w = sc[t - lo] * inv;
            const float *v = V + (size_t)t * dm + h * HD;
            for (int d0 = 0; d0 < HD; d0++) o[d0] += w * v[d0];
        }
    }
}

/* ---------------------------------------------------------------- FlowLM */
typedef struct { int max_len, len; float *k[NL], *v[NL], *sc; } kvcache;

static void kv_free(kvcache *c) { for (int l = 0; l < NL; l++) { free(c->k[l]); free(c->v[l]); } free(c->sc); }
static int kv_init(kvcache *c, int max_len) {
    memset(c, 0, sizeof *c);
    c->max_len = max_len;
    for (int l = 0; l < NL; l++) {
        c->k[l] = (float *)malloc(sizeof(float) * (size_t)max_len * D);
        c->v[l] = (float *)malloc(sizeof(float) * (size_t)max_len * D);
        if (!c->k[l] || !c->v[l]) { kv_free(c); return -1; }
    }
    c->sc = (float *)malloc(sizeof(float) * (size_t)max_len);
    return c->sc ? 0 : -1;
}

/* transformer_forward_step_cached, ptts_flowlm.c:440-566 */
static int flowlm_step(const orc_model *m, kvcache *c, float *x) {
    static float xn[D], qkv[3 * D], att[D], h1[FF], h2[D];
    if (c->len >= c->max_len) return -1;
    int pos = c->len;
    for (int l = 0; l < NL; l++) {
        layernorm(x, D, m->L[l].n1w, m->L[l].n1b, 1e-5f, xn);
        orc_linear(qkv, xn, m->L[l].inproj, NULL, 1, D, 3 * D);
        rope(qkv, qkv + D, NH, pos);
        memcpy(c->k[l] + (size_t)pos * D, qkv + D, sizeof(float) * D);
        memcpy(c->v[l] + (size_t)pos * D, qkv + 2 * D, sizeof(float) * D);
        attend(qkv, c->k[l], c->v[l], NH, 0, pos, c->sc, att);
        orc_linear(h2, att, m->L[l].outproj, NULL, 1, D, D);
        for (int i = 0; i < D; i++) x[i] += h2[i];
        layernorm(x, D, m->L[l].n2w, m->L[l].n2b, 1e-5f, xn);
        orc_linear(h1, xn, m->L[l].l1, NULL, 1, D, FF);
        for (int i = 0; i < FF; i++) h1[i] = gelu_erf(h1[i]);
        orc_linear(h2, h1, m->L[l].l2, NULL, 1, FF, D);
        for (int i = 0; i < D; i++) x[i] += h2[i];
    }
    c->len++;
    return 0;
}

/* timestep_embed, ptts_flowlm.c:572-586 */
static void time_embed(const orc_model *m, int which, float t, float *out) {
    float e[256], h[FD];
    for (int i = 0; i < 128; i++) {
        float f = m->te[which].freqs ? m->te[which].freqs[i] : expf(-logf(10000.0f) * ((float)i / 128.0f));
        e[i] = cosf(f * t);
        e[i + 128] = sinf(f * t);
    }
    orc_linear(h, e, m->te[which].l0w, m->te[which].l0b, 1, 256, FD);
    for (int i = 0; i < FD; i++) h[i] = orc_silu(h[i]);
    orc_linear(out, h, m->te[which].l2w, m->te[which].l2b, 1, FD, FD);
    var_rmsnorm(out, FD, m->te[which].alpha, 1e-5f, out);
}

/* flow_net_forward, ptts_flowlm.c:655-762 */
void orc_flow_net(orc_model *m, const float *cond, float s, float t, const float *xin, float *out) {
    float x[FD], ts[FD], tt[FD], c[FD], sc[FD], h[FD], g[FD], ada[3 * FD];
    orc_linear(x, xin, m->inp_w, m->inp_b, 1, LAT, FD);
    time_embed(m, 0, s, ts);
    time_embed(m, 1, t, tt);
    orc_linear(c, cond, m->cond_w, m->cond_b, 1, D, FD);
    for (int i = 0; i < FD; i++) { c[i] = (ts[i] + tt[i]) * 0.5f + c[i]; sc[i] = orc_silu(c[i]); }
    for (int b = 0; b < FDEPTH; b++) {
        layernorm(x, FD, m->rb[b].lnw, m->rb[b].lnb, 1e-6f, h);
        orc_linear(ada, sc, m->rb[b].adaw, m->rb[b].adab, 1, FD, 3 * FD);
        for (int i = 0; i < FD; i++) h[i] = h[i] * (1.0f + ada[FD + i]) + ada[i];
        orc_linear(g, h, m->rb[b].m0w, m->rb[b].m0b, 1, FD, FD);
        for (int i = 0; i < FD; i++) g[i] = orc_silu(g[i]);
        orc_linear(h, g, m->rb[b].m2w, m->rb[b].m2b, 1, FD, FD);
        for (int i = 0; i < FD; i++) x[i] += ada[2 * FD + i] * h[i];
    }
    layernorm(x, FD, NULL, NULL, 1e-6f, h);
    orc_linear(ada, sc, m->fin_adaw, m->fin_adab, 1, FD, 2 * FD);
    for (int i = 0; i < FD; i++) h[i] = h[i] * (1.0f + ada[FD + i]) + ada[i];
    orc_linear(out, h, m->fin_w, m->fin_b, 1, FD, LAT);
}

/* ptts_flowlm.c:1013-1025 */
static uint32_t xorshift_u32(uint64_t *st) {
    uint64_t x = *st;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    *st = x;
    return (uint32_t)((x * 2685821657736338717ULL) >> 32);
}
static float rng_f01(uint64_t *st) { return (xorshift_u32(st) + 1.0f) / 4294967296.0f; }

/* ptts_flowlm.c:1187-1231 */
void orc_noise(int64_t seed, float temp, float clampv, int frames, float *out) {
    if (seed == -1) seed = (int64_t)time(NULL);
    uint64_t st = (uint64_t)seed;
    float sd = temp > 0.0f ? sqrtf(temp) : 0.0f;
    for (int f = 0; f < frames; f++)
        for (int d0 = 0; d0 < LAT; d0 += 2) {
            float z0 = 0.0f, z1 = 0.0f;
            if (sd > 0.0f) {
                float u1 = rng_f01(&st), u2 = rng_f01(&st);
                float r = sqrtf(-2.0f * logf(u1)), th = 2.0f * (float)M_PI * u2;
                z0 = r * cosf(th) * sd;
                z1 = r * sinf(th) * sd;
            }
            if (clampv > 0.0f) {
                z0 = z0 < -clampv ? -clampv : (z0 > clampv ? clampv : z0);
                z1 = z1 < -clampv ? -clampv : (z1 > clampv ? clampv : z1);
            }
            out[f * LAT + d0] = z0;
            out[f * LAT + d0 + 1] = z1;
        }
}

/* ptts_flowlm_generate_latents, ptts_flowlm.c:1138-1251 */
int orc_flowlm_generate(orc_model *m, const int *tokens, int token_len,
                        const float *cond_prefix, int cond_len,
                        int max_frames, int lsd_steps, float temp, float noise_clamp,
                        int64_t seed, int eos_enabled, float eos_threshold,
                        int eos_min_frames, int eos_after,
                        float *out_latents, int *out_frames_used,
                        float *out_first_eos_logit, float *out_first_cond, float *out_first_flow,
                        float *all_cond, float *all_eos, const float *noise_override) {
    if (!m || !tokens || token_len <= 0 || !out_latents || !out_frames_used) return -1;
    if (max_frames < 1 || cond_len < 0 || (cond_len > 0 && !cond_prefix)) return -1;
    if (eos_min_frames < 1) eos_min_frames = 1;
    if (eos_after < 0) eos_after = 0;
    kvcache c;
    if (kv_init(&c, token_len + cond_len + 1 + max_frames) != 0) return -1;
    float x[D], normed[D], lat[LAT], flow[LAT];
    float *noise = (float *)malloc(sizeof(float) * (size_t)max_frames * LAT);
    if (!noise) { kv_free(&c); return -1; }
    if (noise_override) memcpy(noise, noise_override, sizeof(float) * (size_t)max_frames * LAT);
    else orc_noise(seed, temp, noise_clamp, max_frames, noise);

    for (int t = 0; t < cond_len; t++) {
        memcpy(x, cond_prefix + (size_t)t * D, sizeof x);
        flowlm_step(m, &c, x);
    }
    for (int t = 0; t < token_len; t++) {
        int id = tokens[t];
        if (id < 0 || id > 4000) id = 0;  /* ptts_flowlm.c:1170 */
        memcpy(x, m->embed + (size_t)id * D, sizeof x);
        flowlm_step(m, &c, x);
    }
    orc_linear(x, m->bos, m->in_lin, NULL, 1, LAT, D);
    flowlm_step(m, &c, x);

    int eos_step = -1, used = 0;
    for (int i = 0; i < max_frames; i++) {
        layernorm(x, D, m->onorm_w, m->onorm_b, 1e-5f, normed);
        if (i == 0 && out_first_cond) memcpy(out_first_cond, normed, sizeof normed);
        if (all_cond) memcpy(all_cond + (size_t)i * D, normed, sizeof normed);
        float eos = dotf(m->eos_w, normed, D) + m->eos_b[0];
        if (i == 0 && out_first_eos_logit) *out_first_eos_logit = eos;
        if (all_eos) all_eos[i] = eos;
        if (eos_enabled && i + 1 >= eos_min_frames && eos >= eos_threshold && eos_step < 0) eos_step = i;
        memcpy(lat, noise + (size_t)i * LAT, sizeof lat);
        /* lsd_decode, ptts_flowlm.c:764-779 */
        for (int k = 0; k < lsd_steps; k++) {
            float s = (float)k / (float)lsd_steps, t = (float)(k + 1) / (float)lsd_steps;
            orc_flow_net(m, normed, s, t, lat, flow);
            if (i == 0 && k == 0 && out_first_flow) memcpy(out_first_flow, flow, sizeof flow);
            for (int d0 = 0; d0 < LAT; d0++) lat[d0] += flow[d0] / (float)lsd_steps;
        }
        memcpy(out_latents + (size_t)i * LAT, lat, sizeof lat);
        used = i + 1;
        if (eos_step >= 0 && i >= eos_step + eos_after) break;
        orc_linear(x, lat, m->in_lin, NULL, 1, LAT, D);
        if (flowlm_step(m, &c, x) != 0) { free(noise); kv_free(&c); return -1; }
    }
    *out_frames_used = used;
    free(noise);
    kv_free(&c);
    return 0;
}

void orc_scale_latents(orc_model *m, const float *in, int frames, float *out) {
    for (int i = 0; i < frames; i++)
        for (int d0 = 0; d0 < LAT; d0++) out[i * LAT + d0] = in[i * LAT + d0] * m->emb_std[d0] + m->emb_mean[d0];
}

/* ---------------------------------------------------------------- Mimi (time-major) */

/* Causal Conv1d stride 1 on [T,C]: y[t,o] = b[o] + sum_{kk,i} w[o,i,kk] x[t-(k-1)+kk, i]
 * (ptts_kernels.c:128-153 with stride 1 => left pad k-1). act_in applies ELU to the input. */
static void conv_tm(const conv_t *c, const float *x, int T, float *y) {
    int k = c->k, ci = c->cin, co = c->cout;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T; t++) {
        int kk0 = (k - 1) - t; if (kk0 < 0) kk0 = 0;            /* taps that fall before t=0 read zeros */
        const float *xs = x + (size_t)(t - (k - 1) + kk0) * ci;  /* contiguous span of (k-kk0)*ci floats */
        for (int o = 0; o < co; o++)
            y[(size_t)t * co + o] = c->b[o] + dotf(xs, c->wt + ((size_t)o * k + kk0) * ci, (k - kk0) * ci);
    }
}

/* ConvTranspose1d with k = 2*stride on [T,C] -> [T*s,Cout]:
 * y[t*s+r,o] = b[o] + sum_i w[i,o,r] x[t,i] + w[i,o,r+s] x[t-1,i]   (ptts_kernels.c:172-204) */
static void convtr_tm(const conv_t *c, const float *x, int T, float *y) {
    int s = c->stride, ci = c->cin, co = c->cout;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T; t++)
        for (int r = 0; r < s; r++)
            for (int o = 0; o < co; o++) {
                float v = c->b[o] + dotf(x + (size_t)t * ci, c->wt + ((size_t)r * co + o) * ci, ci);
                if (t > 0) v += dotf(x + (size_t)(t - 1) * ci, c->wt + ((size_t)(r + s) * co + o) * ci, ci);
                y[((size_t)t * s + r) * co + o] = v;
            }
}

static void elu_copy(const float *x, size_t n, float *y) { for (size_t i = 0; i < n; i++) y[i] = elu(x[i]); }

/* Mimi transformer_forward, ptts_mimi.c:301-376 (window ptts_mimi.c:279) */
typedef struct { const float *inproj, *outproj, *n1w, *n1b, *n2w, *n2b, *l1, *l2, *ls1, *ls2; } mimi_layer_t;
static int mimi_transformer_w(const mimi_layer_t *Lw, float *x, int T);
static int mimi_transformer(const orc_model *m, float *x, int T) { return mimi_transformer_w((const mimi_layer_t *)m->M, x, T); }
static int mimi_transformer_w(const mimi_layer_t *Lw, float *x, int T) {
    float *xn = (float *)malloc(sizeof(float) * (size_t)T * MD);
    float *qkv = (float *)malloc(sizeof(float) * (size_t)T * 3 * MD);
    float *K = (float *)malloc(sizeof(float) * (size_t)T * MD);
    float *V = (float *)malloc(sizeof(float) * (size_t)T * MD);
    float *att = (float *)malloc(sizeof(float) * (size_t)T * MD);
    float *h1 = (float *)malloc(sizeof(float) * (size_t)T * MFF);
    if (!xn || !qkv || !K || !V || !att || !h1) { free(xn); free(qkv); free(K); free(V); free(att); free(h1); return -1; }
    for (int l = 0; l < ML; l++) {
        for (int t = 0; t < T; t++) layernorm(x + (size_t)t * MD, MD, Lw[l].n1w, Lw[l].n1b, 1e-5f, xn + (size_t)t * MD);
        orc_linear(qkv, xn, Lw[l].inproj, NULL, T, MD, 3 * MD);
        for (int t = 0; t < T; t++) {
            float *row = qkv + (size_t)t * 3 * MD;
            rope(row, row + MD, MH, t);
            memcpy(K + (size_t)t * MD, row + MD, sizeof(float) * MD);
            memcpy(V + (size_t)t * MD, row + 2 * MD, sizeof(float) * MD);
        }
#pragma omp parallel
        {
            float sc[MCTX];
#pragma omp for schedule(static)
            for (int t = 0; t < T; t++) {
                int lo = t - (MCTX - 1); if (lo < 0) lo = 0;
                attend(qkv + (size_t)t * 3 * MD, K, V, MH, lo, t, sc, att + (size_t)t * MD);
            }
        }
        orc_linear(xn, att, Lw[l].outproj, NULL, T, MD, MD);
        for (size_t i = 0; i < (size_t)T * MD; i++) x[i] += Lw[l].ls1[i % MD] * xn[i];
        for (int t = 0; t < T; t++) layernorm(x + (size_t)t * MD, MD, Lw[l].n2w, Lw[l].n2b, 1e-5f, xn + (size_t)t * MD);
        orc_linear(h1, xn, Lw[l].l1, NULL, T, MD, MFF);
        for (size_t i = 0; i < (size_t)T * MFF; i++) h1[i] = gelu_tanh(h1[i]);
        orc_linear(xn, h1, Lw[l].l2, NULL, T, MFF, MD);
        for (size_t i = 0; i < (size_t)T * MD; i++) x[i] += Lw[l].ls2[i % MD] * xn[i];
    }
    free(xn); free(qkv); free(K); free(V); free(att); free(h1);
    return 0;
}

/* ptts_mimi_decode, ptts_mimi.c:570-738 */
int orc_mimi_decode(orc_model *m, const float *latents, int frames, float *out_audio, int *out_len) {
    if (!m || !latents || !out_audio || !out_len || frames < 1) return -1;
    int T = frames * 16;
    float *q = (float *)malloc(sizeof(float) * (size_t)frames * MD);
    float *x = (float *)malloc(sizeof(float) * (size_t)T * MD);
    if (!q || !x) { free(q); free(x); return -1; }
    /* quantizer.output_proj (ptts_mimi.c:577-586), time-major */
    orc_linear(q, latents, m->quant, NULL, frames, LAT, MD);
    /* depthwise ConvTranspose1d k=32 s=16 (ptts_mimi.c:588-593): two taps per output */
    for (int t = 0; t < frames; t++)
        for (int r = 0; r < 16; r++)
            for (int c = 0; c < MD; c++) {
                float v = m->ups[c * 32 + r] * q[(size_t)t * MD + c];
                if (t > 0) v += m->ups[c * 32 + r + 16] * q[(size_t)(t - 1) * MD + c];
                x[((size_t)t * 16 + r) * MD + c] = v;
            }
    free(q);
    if (mimi_transformer(m, x, T) != 0) { free(x); return -1; }

    /* SEANet, ptts_mimi.c:676-737 */
    float *y = (float *)malloc(sizeof(float) * (size_t)T * 512);
    if (!y) { free(x); return -1; }
    conv_tm(&m->dec_in, x, T, y);
    free(x);
    x = y;
    int C = 512;
    for (int st = 0; st < 3; st++) {
        size_t n = (size_t)T * C;
        float *a = (float *)malloc(sizeof(float) * n);
        elu_copy(x, n, a);
        int T2 = T * m->up[st].stride, C2 = m->up[st].cout;
        float *u = (float *)malloc(sizeof(float) * (size_t)T2 * C2);
        convtr_tm(&m->up[st], a, T, u);
        free(a); free(x);
        /* resblock, ptts_mimi.c:160-177: u + conv_k1(ELU(conv_k3(ELU(u)))) */
        size_t n2 = (size_t)T2 * C2;
        float *e = (float *)malloc(sizeof(float) * n2);
        float *hmid = (float *)malloc(sizeof(float) * (size_t)T2 * (C2 / 2));
        float *r2 = (float *)malloc(sizeof(float) * n2);
        elu_copy(u, n2, e);
        conv_tm(&m->res1[st], e, T2, hmid);
        elu_copy(hmid, (size_t)T2 * (C2 / 2), hmid);
        conv_tm(&m->res2[st], hmid, T2, r2);
        for (size_t i = 0; i < n2; i++) u[i] += r2[i];
        free(e); free(hmid); free(r2);
        x = u; T = T2; C = C2;
    }
    {
        size_t n = (size_t)T * C;
        float *a = (float *)malloc(sizeof(float) * n);
        elu_copy(x, n, a);
        conv_tm(&m->dec_out, a, T, out_audio);
        free(a);
    }
    free(x);
    *out_len = T;
    return 0;
}

/* ---------------------------------------------------------------- voice-cloning encoder
 * PARITY UNPINNED END TO END: the reference only shape-checks these tensors (ptts.c:914-944) and loads speaker_proj_weight
 * without using it (ptts_flowlm.c:874); it has no forward pass to compare with. The composition below is the mirror image of
 * the decoder the reference does implement, built from the reference's own primitives, each of which IS pinned against the
 * compiled reference (tests/test_oracle_vs_reference.py):
 *   Conv1d, causal, any stride: left pad k - stride, out_len = T / stride            (ptts_kernels.c:114-156 -> orc_conv1d)
 *   ResBlock: x + conv_k1(ELU(conv_k3(ELU(x))))                                       (ptts_mimi.c:160-177)
 *   transformer: 2 layers, window 250, RoPE from position 0, layer scale              (ptts_mimi.c:301-376)
 * order of the modules: ptts.c:919-944 (model.0 conv, then per ratio 4, 5, 6: resblock, ELU, strided conv; ELU, conv k3),
 * encoder_transformer, downsample.conv (k 32, stride 16, no bias, ptts.c:914), speaker projection [1024,512]. */
int orc_voice_encode(orc_model *m, const float *pcm, int n_samples, float *out_cond, int max_frames, int *out_frames) {
    if (!m || !pcm || !out_cond || !out_frames || n_samples < 1) return -1;
    if (!m->has_encoder) { snprintf(g_err, sizeof g_err, "checkpoint has no encoder tensors"); return -1; }
    const int F = (n_samples + 1919) / 1920;       /* whole 12.5 Hz frames, zero padded on the right */
    if (F > max_frames) { snprintf(g_err, sizeof g_err, "voice_encode: %d frames exceed the output capacity %d", F, max_frames); return -1; }
    int T = F * 1920, C = 64;
    float *x = (float *)calloc((size_t)T, sizeof(float)), *y = NULL;
    if (!x) return -1;
    memcpy(x, pcm, sizeof(float) * (size_t)n_samples);
    y = (float *)malloc(sizeof(float) * (size_t)C * T);
    orc_conv1d(y, x, m->enc_in.w, m->enc_in.b, 1, 64, T, 7, 1, 1);          /* channel-major [C,T] from here */
    free(x); x = y;
    for (int st = 0; st < 3; st++) {
        const size_t n = (size_t)C * T;
        float *e = (float *)malloc(sizeof(float) * n), *h = (float *)malloc(sizeof(float) * n / 2), *r = (float *)malloc(sizeof(float) * n);
        elu_copy(x, n, e);
        orc_conv1d(h, e, m->eres1[st].w, m->eres1[st].b, C, C / 2, T, 3, 1, 1);
        elu_copy(h, n / 2, h);
        orc_conv1d(r, h, m->eres2[st].w, m->eres2[st].b, C / 2, C, T, 1, 1, 1);
        for (size_t i = 0; i < n; i++) x[i] += r[i];
        elu_copy(x, n, e);
        const int s = m->edown[st].stride, T2 = T / s, C2 = 2 * C;
        float *d = (float *)malloc(sizeof(float) * (size_t)C2 * T2);
        orc_conv1d(d, e, m->edown[st].w, m->edown[st].b, C, C2, T, m->edown[st].k, s, 1);
        free(e); free(h); free(r); free(x);
        x = d; T = T2; C = C2;
    }
    {   /* ELU, conv k3 512 -> 512 */
        const size_t n = (size_t)C * T;
        float *e = (float *)malloc(sizeof(float) * n), *o = (float *)malloc(sizeof(float) * n);
        elu_copy(x, n, e);
        orc_conv1d(o, e, m->enc_out.w, m->enc_out.b, 512, 512, T, 3, 1, 1);
        free(e); free(x); x = o;
    }
    /* [512,T] -> [T,512] (ptts_mimi.c:134-150), encoder transformer, back */
    float *tm = (float *)malloc(sizeof(float) * (size_t)T * MD);
    for (int c = 0; c < MD; c++) for (int t = 0; t < T; t++) tm[(size_t)t * MD + c] = x[(size_t)c * T + t];
    if (mimi_transformer_w((const mimi_layer_t *)m->E, tm, T) != 0) { free(tm); free(x); return -1; }
    for (int c = 0; c < MD; c++) for (int t = 0; t < T; t++) x[(size_t)c * T + t] = tm[(size_t)t * MD + c];
    free(tm);
    /* learnt downsample 200 Hz -> 12.5 Hz */
    float *lat = (float *)malloc(sizeof(float) * (size_t)MD * F);
    orc_conv1d(lat, x, m->down.w, NULL, 512, 512, T, 32, 16, 1);
    free(x);
    /* speaker projection: cond[t, :] = W[1024,512] . lat[:, t] */
    for (int t = 0; t < F; t++)
        for (int o = 0; o < D; o++) {
            float a = 0.0f;
            for (int c = 0; c < MD; c++) a += m->spk[(size_t)o * MD + c] * lat[(size_t)c * F + t];
            out_cond[(size_t)t * D + o] = a;
        }
    free(lat);
    *out_frames = F;
    return 0;
}
