/*
 * ptts_b200_models.c - host C layer that makes the B200 engine a drop-in for the reference's L3
 * model files.  It provides the public symbols of ptts_flowlm.c and ptts_mimi.c
 * (ptts_flowlm.h:19-55, ptts_mimi.h:14-27) on top of the C ABI in include/ptts_b200.h, so the
 * UNMODIFIED reference sources main.c, ptts.c, ptts_audio.c, ptts_safetensors.c, ptts_spm.c link
 * against it instead of ptts_flowlm.c + ptts_mimi.c + ptts_kernels.c + ptts_cuda.c:
 *
 *     ptts_generate (ptts.c:1011)  ->  ptts_flowlm_load / ptts_mimi_load      (engine lookup, no reload)
 *                                  ->  ptts_flowlm_generate_latents (ptts.c:1083)  -> ptts_b200_flowlm_generate_latents
 *                                  ->  ptts_flowlm_scale_latents   (ptts.c:1111)  -> ptts_b200_flowlm_scale_latents
 *                                  ->  ptts_mimi_decode            (ptts.c:1128)  -> ptts_b200_mimi_decode
 *
 * ptts.h, ptts_params, the safetensors layout, tokenizer, voice loading and WAV writer are untouched.
 * The engine is created once per ptts_ctx (keyed by its mmap'd weights file) and lives until exit,
 * which removes the per-call weight reload of ptts.c:1047-1053 and the stale host-pointer weight
 * cache hazard of ptts_cuda.c:617-644.
 * There is no CPU fallback: every failure propagates as -1 / NULL with the message on stderr.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptts_b200.h"
#include "ptts_flowlm.h"   /* reference headers, -I$(REF) */
#include "ptts_internal.h"
#include "ptts_mimi.h"

struct ptts_flowlm { ptts_ctx *ctx; ptts_b200_engine *eng; };
struct ptts_mimi { ptts_ctx *ctx; ptts_b200_engine *eng; };

#define MAX_ENGINES 8
static struct { const safetensors_file_t *key; ptts_b200_engine *eng; } g_engines[MAX_ENGINES];
static int g_n_engines = 0, g_atexit = 0;

static void shutdown_engines(void) {
    for (int i = 0; i < g_n_engines; i++) ptts_b200_destroy(g_engines[i].eng);
    g_n_engines = 0;
}

static ptts_b200_engine *engine_for(ptts_ctx *ctx) {
    if (!ctx || !ctx->weights) return NULL;
    for (int i = 0; i < g_n_engines; i++)
        if (g_engines[i].key == ctx->weights) return g_engines[i].eng;
    if (g_n_engines == MAX_ENGINES) { fprintf(stderr, "[ptts] B200: too many model contexts\n"); return NULL; }
    const safetensors_file_t *sf = ctx->weights;
    ptts_b200_tensor *tab = (ptts_b200_tensor *)calloc((size_t)sf->num_tensors, sizeof *tab);
    if (!tab) return NULL;
    int n = 0;
    for (int i = 0; i < sf->num_tensors; i++) {
        const safetensor_t *t = &sf->tensors[i];
        if (t->dtype != DTYPE_F32 && t->dtype != DTYPE_F16 && t->dtype != DTYPE_BF16) continue;
        int64_t numel = 1;
        for (int d = 0; d < t->ndim; d++) numel *= t->shape[d];
        tab[n].name = t->name;
        tab[n].data = safetensors_data(sf, t);   /* zero-copy view of the mmap (ptts_safetensors.c:213) */
        tab[n].dtype = (int)t->dtype;            /* same numbering as safetensor_dtype_t */
        tab[n].numel = numel;
        n++;
    }
    const char *dev = getenv("PTTS_B200_DEVICE");
    ptts_b200_engine *eng = NULL;
    int rc = ptts_b200_create(&eng, dev ? atoi(dev) : 0, tab, n);
    free(tab);
    if (rc != 0) {
        fprintf(stderr, "[ptts] B200 engine creation failed: %s\n", ptts_b200_last_error());
        return NULL;
    }
    if (!g_atexit) { atexit(shutdown_engines); g_atexit = 1; }
    g_engines[g_n_engines].key = sf;
    g_engines[g_n_engines].eng = eng;
    g_n_engines++;
    return eng;
}

/* ---- ptts_flowlm.h ---------------------------------------------------------------- */
ptts_flowlm *ptts_flowlm_load(ptts_ctx *ctx) {
    ptts_b200_engine *eng = engine_for(ctx);
    if (!eng) return NULL;
    ptts_flowlm *fm = (ptts_flowlm *)calloc(1, sizeof *fm);
    if (fm) { fm->ctx = ctx; fm->eng = eng; }
    return fm;
}
void ptts_flowlm_free(ptts_flowlm *fm) { free(fm); }

int ptts_flowlm_generate_latents(ptts_flowlm *fm, const int *tokens, int token_len, const float *cond_prefix, int cond_len,
                                 int max_frames, int lsd_steps, float temp, float noise_clamp, int64_t seed, int eos_enabled,
                                 float eos_threshold, int eos_min_frames, int eos_after, float *out_latents,
                                 int *out_frames_used, float *out_first_eos_logit, float *out_first_cond, float *out_first_flow) {
    if (!fm) return -1;
    int rc = ptts_b200_flowlm_generate_latents(fm->eng, tokens, token_len, cond_prefix, cond_len, max_frames, lsd_steps, temp,
                                               noise_clamp, seed, eos_enabled, eos_threshold, eos_min_frames, eos_after,
                                               out_latents, out_frames_used, out_first_eos_logit, out_first_cond, out_first_flow);
    if (rc != 0) fprintf(stderr, "[ptts] B200: %s\n", ptts_b200_last_error());
    return rc;
}

int ptts_flowlm_forward_one(ptts_flowlm *fm, const int *tokens, int token_len, const float *cond_prefix, int cond_len,
                            int lsd_steps, float temp, float noise_clamp, int64_t seed, float *out_latent, float *out_eos_logit) {
    /* single step == first frame of the cached generation loop (ptts_flowlm.c:1129-1136) */
    int used = 0;
    return ptts_flowlm_generate_latents(fm, tokens, token_len, cond_prefix, cond_len, 1, lsd_steps, temp, noise_clamp, seed, 0,
                                        0.0f, 1, 0, out_latent, &used, out_eos_logit, NULL, NULL);
}

int ptts_flowlm_forward_next(ptts_flowlm *fm, const int *tokens, int token_len, const float *cond_prefix, int cond_len,
                             const float *prev_latents, int prev_len, int lsd_steps, float temp, float noise_clamp,
                             int64_t *seed_io, float *out_latent, float *out_eos_logit) {
    if (prev_len == 0 && !prev_latents)
        return ptts_flowlm_forward_one(fm, tokens, token_len, cond_prefix, cond_len, lsd_steps, temp, noise_clamp,
                                       seed_io ? *seed_io : -1, out_latent, out_eos_logit);
    fprintf(stderr, "[ptts] B200: ptts_flowlm_forward_next (naive non-cached debug path) is not on the accelerated path\n");
    return -1;
}

void ptts_flowlm_scale_latents(const ptts_flowlm *fm, const float *in_latents, int frames, float *out_latents) {
    if (!fm || !in_latents || !out_latents || frames <= 0) return;
    ptts_b200_flowlm_scale_latents(fm->eng, in_latents, frames, out_latents);
}

/* ---- ptts_mimi.h ------------------------------------------------------------------ */
ptts_mimi *ptts_mimi_load(ptts_ctx *ctx) {
    ptts_b200_engine *eng = engine_for(ctx);
    if (!eng) return NULL;
    ptts_mimi *mm = (ptts_mimi *)calloc(1, sizeof *mm);
    if (mm) { mm->ctx = ctx; mm->eng = eng; }
    return mm;
}
void ptts_mimi_free(ptts_mimi *mm) { free(mm); }

int ptts_mimi_decode(ptts_mimi *mm, const float *latents, int frames, float *out_audio, int *out_len) {
    if (!mm) return -1;
    int rc = ptts_b200_mimi_decode(mm->eng, latents, frames, out_audio, out_len);
    if (rc != 0) fprintf(stderr, "[ptts] B200: %s\n", ptts_b200_last_error());
    return rc;
}
int ptts_mimi_decode_one(ptts_mimi *mm, const float *latent, float *out_audio, int *out_len) {
    return ptts_mimi_decode(mm, latent, 1, out_audio, out_len);
}
int ptts_mimi_forward_one(ptts_mimi *mm, const float *latent, float *out_embed) {
    (void)mm; (void)latent; (void)out_embed;
    fprintf(stderr, "[ptts] B200: ptts_mimi_forward_one (--mimi-test debug statistic) is not on the accelerated path\n");
    return -1;
}
