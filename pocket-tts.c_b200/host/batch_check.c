/*
 * batch_check.c - exercises the additive batched entry of the host layer, ptts_generate_batch (host/ptts_b200_models.c), through the
 * reference's UNMODIFIED public API for everything else (ptts_load_dir, ptts_audio_save_wav, ptts.h): N texts -> N WAV files.
 *   batch_check <model_dir> <outdir> <text> [<text> ...]
 * Utterance i uses seed 123 + i, temp 0.7, EOS disabled (threshold 1e9), frames from ptts_estimate_frames; tests/test_gpu_cli_dropin.py
 * runs the reference CPU CLI once per text with the same settings and compares the WAVs.
 */
#include <stdio.h>
#include <stdlib.h>

#include "ptts.h"

ptts_audio **ptts_generate_batch(ptts_ctx *ctx, const char *const *texts, int n, const char *voice_path, const ptts_params *params);
void ptts_audio_batch_free(ptts_audio **audio, int n);

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s model_dir outdir text...\n", argv[0]); return 2; }
    ptts_ctx *ctx = ptts_load_dir(argv[1]);
    if (!ctx) { fprintf(stderr, "load: %s\n", ptts_get_error()); return 1; }
    ptts_params p = PTTS_PARAMS_DEFAULT;
    p.temp = 0.7f; p.seed = 123; p.eos_threshold = 1e9f;
    const int n = argc - 3;
    ptts_audio **a = ptts_generate_batch(ctx, (const char *const *)(argv + 3), n, "alba", &p);
    if (!a) { fprintf(stderr, "ptts_generate_batch failed\n"); ptts_free(ctx); return 1; }
    int rc = 0;
    for (int i = 0; i < n; i++) {
        char path[1024];
        snprintf(path, sizeof path, "%s/b%d.wav", argv[2], i);
        if (ptts_audio_save_wav(a[i], path) != 0) rc = 1;
        printf("utterance %d: %d samples\n", i, a[i]->num_samples);
    }
    ptts_audio_batch_free(a, n);
    ptts_free(ctx);
    return rc;
}
