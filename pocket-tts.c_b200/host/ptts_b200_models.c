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
 * Engine lifetime.  ptts_generate loads and frees both model handles on EVERY call (ptts.c:1047-1053, 1157-1158), and
 * ptts_free (ptts.c:414) cannot be hooked without patching ptts.c, so the engine must outlive the handles to keep the
 * weights resident, yet must never be handed to a different model:
 *   - every handle (ptts_flowlm / ptts_mimi) holds a reference on its engine; ptts_flowlm_free / ptts_mimi_free drop it;
 *   - an engine is identified by the CONTENT of the weights file it was built from (file size, tensor count, a hash of the
 *     safetensors JSON header and of 64 data samples spread over the file), never by the safetensors_file_t address alone:
 *     ptts_free + ptts_load_dir of another model at a recycled address yields a different fingerprint and a new engine
 *     (the stale host-pointer cache hazard of ptts_cuda.c:617-644 does not exist here);
 *   - engines without handles stay cached (weights resident between ptts_generate calls), at most MAX_IDLE of them: the
 *     least recently used idle engine is destroyed when a new one is needed, and all of them at exit.
 * There is no CPU fallback: every failure propagates as -1 / NULL with the message on stderr.
 * Threading: like the reference host code (ptts.c has none), this layer is not re-entrant; the engines themselves are
 * independent of each other (include/ptts_b200.h).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptts_b200.h"
#include "ptts_flowlm.h"   /* reference headers, -I$(REF) */
#include "ptts_internal.h"
#include "ptts_mimi.h"

typedef struct {
    ptts_b200_pool *pool;   /* one engine per device of PTTS_B200_DEVICES ("all" or "0,1,..."); default: the single device PTTS_B200_DEVICE */
    ptts_b200_engine *eng;  /* engine of pool slot 0: the single-utterance drop-in entries run here */
    uint64_t fp;            /* content fingerprint of the weights file */
    size_t file_size;
    int num_tensors;
    int refs;               /* live ptts_flowlm / ptts_mimi handles */
    uint64_t last_use;
} engine_slot;

struct ptts_flowlm { ptts_ctx *ctx; engine_slot *slot; };
struct ptts_mimi { ptts_ctx *ctx; engine_slot *slot; };

#define MAX_ENGINES 8
#define MAX_IDLE 2
static engine_slot g_slots[MAX_ENGINES];
static uint64_t g_clock = 0;
static int g_atexit = 0;

static void shutdown_engines(void) {
    for (int i = 0; i < MAX_ENGINES; i++)
        if (g_slots[i].eng) { ptts_b200_pool_destroy(g_slots[i].pool); g_slots[i].pool = NULL; g_slots[i].eng = NULL; g_slots[i].refs = 0; }
}

static uint64_t fnv1a(uint64_t h, const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; }
    return h;
}

static uint64_t weights_fingerprint(const safetensors_file_t *sf) {
    uint64_t h = 1469598103934665603ULL;
    h = fnv1a(h, &sf->file_size, sizeof sf->file_size);
    h = fnv1a(h, &sf->num_tensors, sizeof sf->num_tensors);
    if (sf->header_json) h = fnv1a(h, sf->header_json, sf->header_size);
    const size_t data0 = 8 + sf->header_size;
    if (sf->data && sf->file_size > data0 + 64) {
        const size_t span = sf->file_size - data0 - 64;
        for (int i = 0; i < 64; i++) h = fnv1a(h, (const unsigned char *)sf->data + data0 + (span / 63) * (size_t)i, 64);
    }
    return h;
}

static void drop_idle(int keep) {   /* destroy least recently used idle engines until at most `keep` remain */
    for (;;) {
        int idle = 0, lru = -1;
        for (int i = 0; i < MAX_ENGINES; i++)
            if (g_slots[i].eng && g_slots[i].refs == 0) {
                idle++;
                if (lru < 0 || g_slots[i].last_use < g_slots[lru].last_use) lru = i;
            }
        if (idle <= keep || lru < 0) return;
        ptts_b200_pool_destroy(g_slots[lru].pool);
        memset(&g_slots[lru], 0, sizeof g_slots[lru]);
    }
}

static engine_slot *engine_acquire(ptts_ctx *ctx) {
    if (!ctx || !ctx->weights) return NULL;
    const safetensors_file_t *sf = ctx->weights;
    const uint64_t fp = weights_fingerprint(sf);
    for (int i = 0; i < MAX_ENGINES; i++) {
        engine_slot *s = &g_slots[i];
        if (s->eng && s->fp == fp && s->file_size == sf->file_size && s->num_tensors == sf->num_tensors) {
            s->refs++; s->last_use = ++g_clock;
            return s;
        }
    }
    drop_idle(MAX_IDLE - 1);   /* room for the new engine's device memory before it is allocated */
    int slot = -1;
    for (int i = 0; i < MAX_ENGINES; i++) if (!g_slots[i].eng) { slot = i; break; }
    if (slot < 0) { drop_idle(0); for (int i = 0; i < MAX_ENGINES; i++) if (!g_slots[i].eng) { slot = i; break; } }
    if (slot < 0) { fprintf(stderr, "[ptts] B200: %d model contexts are in use at once\n", MAX_ENGINES); return NULL; }
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
    /* devices: PTTS_B200_DEVICES = "all" | "0,1,2" (utterance sharding of ptts_generate_batch); default one device, PTTS_B200_DEVICE */
    int devs[64], nd = 0, all = 0;
    const char *dl = getenv("PTTS_B200_DEVICES"), *dev = getenv("PTTS_B200_DEVICE");
    if (dl && strcmp(dl, "all") == 0) all = 1;
    else if (dl && *dl) {
        for (const char *q = dl; *q && nd < 64;) {
            devs[nd++] = atoi(q);
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
    }
    if (!all && nd == 0) devs[nd++] = dev ? atoi(dev) : 0;
    ptts_b200_pool *pool = NULL;
    int rc = ptts_b200_pool_create(&pool, all ? NULL : devs, all ? 0 : nd, tab, n);
    free(tab);
    if (rc != 0) {
        fprintf(stderr, "[ptts] B200 engine creation failed: %s\n", ptts_b200_last_error());
        return NULL;
    }
    if (!g_atexit) { atexit(shutdown_engines); g_atexit = 1; }
    engine_slot *s = &g_slots[slot];
    s->pool = pool; s->eng = ptts_b200_pool_engine(pool, 0); s->fp = fp; s->file_size = sf->file_size; s->num_tensors = sf->num_tensors; s->refs = 1; s->last_use = ++g_clock;
    return s;
}

static void engine_release(engine_slot *s) {
    if (!s || s->refs <= 0) return;
    s->refs--;
    if (s->refs == 0) drop_idle(MAX_IDLE);
}

/* test / diagnostics hook: engines currently alive in this process (live + idle) */
int ptts_b200_host_engine_count(void) {
    int n = 0;
    for (int i = 0; i < MAX_ENGINES; i++) n += g_slots[i].eng != NULL;
    return n;
}

/* ---- additive batched entry at the ptts.h level ------------------------------------------------------------------------
 * ptts.h is single-utterance (ptts_generate, ptts.h:75-76); configs 3-5 need a batch. ptts_generate_batch mirrors ptts_generate
 * (ptts.c:1011-1161) step by step with the reference's own host functions - ptts_prepare_text (ptts.c:219-283), ptts_tokenize
 * (the reference SentencePiece reader), ptts_estimate_frames (ptts.c:285-291), ptts_load_voice_conditioning (ptts.c:293-364) - then
 * hands ALL utterances to the engine pool: bucketed by estimated frames, sharded over the visible devices (one host thread +
 * engine per device), PCM gathered in the original order. Utterance i draws its noise from seed params->seed + i
 * (params->seed == -1: time based, as ptts_flowlm.c:1188). Returns n ptts_audio objects (free with ptts_audio_batch_free), NULL on
 * failure. ptts.h, ptts_params and every existing symbol are untouched. */
ptts_audio **ptts_generate_batch(ptts_ctx *ctx, const char *const *texts, int n, const char *voice_path, const ptts_params *params);
void ptts_audio_batch_free(ptts_audio **audio, int n);

#include <time.h>
#include "ptts.h"

void ptts_audio_batch_free(ptts_audio **audio, int n) {
    if (!audio) return;
    for (int i = 0; i < n; i++) ptts_audio_free(audio[i]);
    free(audio);
}

ptts_audio **ptts_generate_batch(ptts_ctx *ctx, const char *const *texts, int n, const char *voice_path, const ptts_params *params) {
    if (!ctx || !texts || n < 1) return NULL;
    ptts_params p = PTTS_PARAMS_DEFAULT;
    if (params) p = *params;
    if (p.sample_rate <= 0) p.sample_rate = 24000;
    if (p.num_steps <= 0) p.num_steps = 1;
    engine_slot *slot = engine_acquire(ctx);
    if (!slot) return NULL;
    int *offs = (int *)calloc((size_t)n + 1, sizeof(int)), *frames = (int *)calloc((size_t)n, sizeof(int));
    int *eos_after = (int *)calloc((size_t)n, sizeof(int)), *used = (int *)calloc((size_t)n, sizeof(int));
    int64_t *seeds = (int64_t *)calloc((size_t)n, sizeof(int64_t));
    int *flat = NULL, flat_cap = 0, fcap = 0, ok = offs && frames && eos_after && used && seeds;
    float *voice = NULL, *pcm = NULL;
    int voice_len = 0;
    ptts_audio **out = NULL;
    const int64_t seed0 = p.seed == -1 ? (int64_t)time(NULL) : p.seed;
    for (int i = 0; ok && i < n; i++) {
        int words = 0, guess = 0, *ids = NULL, len = 0;
        char *prep = ptts_prepare_text(texts[i], &words, &guess);
        if (!prep || ptts_tokenize(ctx, prep, &ids, &len) != 0 || len <= 0) { fprintf(stderr, "[ptts] B200: cannot tokenize utterance %d\n", i); free(prep); free(ids); ok = 0; break; }
        free(prep);
        if (offs[i] + len > flat_cap) {
            flat_cap = (offs[i] + len) * 2 + 64;
            int *nf = (int *)realloc(flat, (size_t)flat_cap * sizeof(int));
            if (!nf) { free(ids); ok = 0; break; }
            flat = nf;
        }
        memcpy(flat + offs[i], ids, (size_t)len * sizeof(int));
        free(ids);
        offs[i + 1] = offs[i] + len;
        frames[i] = p.num_frames > 0 ? p.num_frames : ptts_estimate_frames(words);
        eos_after[i] = p.eos_after > 0 ? p.eos_after : guess;
        seeds[i] = seed0 + i;
        if (frames[i] > fcap) fcap = frames[i];
    }
    if (ok && ptts_load_voice_conditioning(ctx, voice_path, &voice, &voice_len) != 0) ok = 0;
    if (ok) {
        pcm = (float *)malloc((size_t)n * fcap * 1920 * sizeof(float));
        ok = pcm != NULL;
    }
    if (ok) {
        ptts_b200_gen_params gp = {p.num_steps, p.temp, p.noise_clamp, p.eos_enabled, p.eos_threshold, p.eos_min_frames > 0 ? p.eos_min_frames : 1,
                                   p.eos_after > 0 ? p.eos_after : 3, 1};
        if (ptts_b200_pool_generate(slot->pool, n, flat, offs, voice, voice_len, frames, seeds, eos_after, &gp, 0, NULL, used, pcm) != 0) {
            fprintf(stderr, "[ptts] B200: %s\n", ptts_b200_last_error());
            ok = 0;
        }
    }
    if (ok) {
        out = (ptts_audio **)calloc((size_t)n, sizeof *out);
        for (int i = 0; out && i < n; i++) {
            out[i] = ptts_audio_create(p.sample_rate, 1, used[i] * 1920);
            if (!out[i]) { ptts_audio_batch_free(out, n); out = NULL; break; }
            memcpy(out[i]->samples, pcm + (size_t)i * fcap * 1920, (size_t)used[i] * 1920 * sizeof(float));
        }
    }
    free(pcm); free(voice); free(flat); free(offs); free(frames); free(eos_after); free(used); free(seeds);
    engine_release(slot);
    return out;
}

/* ---- additive voice-cloning entry at the ptts.h level -------------------------------------------------
 * 24 kHz mono PCM -> conditioning rows [*out_len][1024] (malloc'd, caller frees), the tensor ptts_load_voice_conditioning
 * (ptts.c:293-364) otherwise reads from a voice file as "audio_prompt". The reference loads the tensors involved
 * (encoder.*, encoder_transformer.*, downsample.conv, speaker_proj_weight) only to shape-check them (ptts.c:914-944,
 * ptts_flowlm.c:874); the engine runs them (ptts_b200_voice_encode). Returns 0, or -1 (message on stderr). */
int ptts_voice_from_pcm(ptts_ctx *ctx, const float *pcm, int n_samples, float **out_cond, int *out_len) {
    if (!ctx || !pcm || n_samples < 1 || !out_cond || !out_len) return -1;
    engine_slot *slot = engine_acquire(ctx);
    if (!slot) return -1;
    const int frames = (n_samples + 1919) / 1920;
    float *cond = (float *)malloc((size_t)frames * 1024 * sizeof(float));
    int got = 0, rc = -1;
    if (cond && ptts_b200_voice_encode(slot->eng, pcm, n_samples, cond, frames, &got) == 0) {
        *out_cond = cond; *out_len = got; rc = 0;
    } else {
        fprintf(stderr, "[ptts] B200: %s\n", cond ? ptts_b200_last_error() : "out of memory");
        free(cond);
    }
    engine_release(slot);
    return rc;
}

/* ---- ptts_flowlm.h ---------------------------------------------------------------- */
ptts_flowlm *ptts_flowlm_load(ptts_ctx *ctx) {
    engine_slot *s = engine_acquire(ctx);
    if (!s) return NULL;
    ptts_flowlm *fm = (ptts_flowlm *)calloc(1, sizeof *fm);
    if (!fm) { engine_release(s); return NULL; }
    fm->ctx = ctx; fm->slot = s;
    return fm;
}
void ptts_flowlm_free(ptts_flowlm *fm) {
    if (!fm) return;
    engine_release(fm->slot);
    free(fm);
}

int ptts_flowlm_generate_latents(ptts_flowlm *fm, const int *tokens, int token_len, const float *cond_prefix, int cond_len,
                                 int max_frames, int lsd_steps, float temp, float noise_clamp, int64_t seed, int eos_enabled,
                                 float eos_threshold, int eos_min_frames, int eos_after, float *out_latents,
                                 int *out_frames_used, float *out_first_eos_logit, float *out_first_cond, float *out_first_flow) {
    if (!fm) return -1;
    int rc = ptts_b200_flowlm_generate_latents(fm->slot->eng, tokens, token_len, cond_prefix, cond_len, max_frames, lsd_steps, temp,
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
    /* teacher-forced single step over a caller-supplied latent history (ptts_flowlm.c:1027-1127): no caller in the reference
     * tree (main.c and ptts.c use ptts_flowlm_generate_latents); documented as not provided in INTEGRATION.md */
    fprintf(stderr, "[ptts] B200: ptts_flowlm_forward_next with a latent history is not provided by the engine\n");
    return -1;
}

void ptts_flowlm_scale_latents(const ptts_flowlm *fm, const float *in_latents, int frames, float *out_latents) {
    if (!fm || !in_latents || !out_latents || frames <= 0) return;
    ptts_b200_flowlm_scale_latents(fm->slot->eng, in_latents, frames, out_latents);
}

/* ---- ptts_mimi.h ------------------------------------------------------------------ */
ptts_mimi *ptts_mimi_load(ptts_ctx *ctx) {
    engine_slot *s = engine_acquire(ctx);
    if (!s) return NULL;
    ptts_mimi *mm = (ptts_mimi *)calloc(1, sizeof *mm);
    if (!mm) { engine_release(s); return NULL; }
    mm->ctx = ctx; mm->slot = s;
    return mm;
}
void ptts_mimi_free(ptts_mimi *mm) {
    if (!mm) return;
    engine_release(mm->slot);
    free(mm);
}

int ptts_mimi_decode(ptts_mimi *mm, const float *latents, int frames, float *out_audio, int *out_len) {
    if (!mm) return -1;
    int rc = ptts_b200_mimi_decode(mm->slot->eng, latents, frames, out_audio, out_len);
    if (rc != 0) fprintf(stderr, "[ptts] B200: %s\n", ptts_b200_last_error());
    return rc;
}
int ptts_mimi_decode_one(ptts_mimi *mm, const float *latent, float *out_audio, int *out_len) {
    return ptts_mimi_decode(mm, latent, 1, out_audio, out_len);
}
int ptts_mimi_forward_one(ptts_mimi *mm, const float *latent, float *out_embed) {
    if (!mm) return -1;
    int rc = ptts_b200_mimi_forward_one(mm->slot->eng, latent, out_embed);   /* the --mimi-test parity hook (main.c:444) */
    if (rc != 0) fprintf(stderr, "[ptts] B200: %s\n", ptts_b200_last_error());
    return rc;
}
