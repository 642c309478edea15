/*
 * reload_check.c - lifetime test of the host drop-in layer (host/ptts_b200_models.c) through the reference's UNMODIFIED
 * public API (ptts.h): load model A, generate; free; load model B (other weights), generate; free; load A again, generate.
 * The hazard under test is the host-pointer-keyed cache of ptts_cuda.c:617-644 against ptts_free + ptts_load_dir
 * (ptts.c:370-414): after the free, malloc may hand the next context the same safetensors_file_t address.
 * Writes a.wav / b.wav / a2.wav into <outdir> and prints the number of engines alive; tests/test_gpu_cli_dropin.py compares
 * the WAVs with the reference CPU CLI run on the same two models.
 *
 *   reload_check <model_dir_A> <model_dir_B> <outdir>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptts.h"

int ptts_b200_host_engine_count(void);

static int gen(const char *dir, const char *out, int repeats) {
    ptts_ctx *ctx = ptts_load_dir(dir);
    if (!ctx) { fprintf(stderr, "load %s: %s\n", dir, ptts_get_error()); return -1; }
    ptts_params p = PTTS_PARAMS_DEFAULT;
    p.temp = 0.7f; p.seed = 123; p.num_frames = 5; p.eos_threshold = 1e9f;
    int rc = 0;
    for (int r = 0; r < repeats && rc == 0; r++) {   /* several ptts_generate calls on one context: one engine, weights resident */
        ptts_audio *a = ptts_generate(ctx, "Hello world!", "alba", &p);
        if (!a) { fprintf(stderr, "generate: %s\n", ptts_get_error()); rc = -1; break; }
        if (r == repeats - 1 && ptts_audio_save_wav(a, out) != 0) rc = -1;
        ptts_audio_free(a);
    }
    printf("after %s: engines alive = %d\n", out, ptts_b200_host_engine_count());
    ptts_free(ctx);
    return rc;
}

int main(int argc, char **argv) {
    if (argc != 4) { fprintf(stderr, "usage: %s dirA dirB outdir\n", argv[0]); return 2; }
    char pa[1024], pb[1024], pa2[1024];
    snprintf(pa, sizeof pa, "%s/a.wav", argv[3]);
    snprintf(pb, sizeof pb, "%s/b.wav", argv[3]);
    snprintf(pa2, sizeof pa2, "%s/a2.wav", argv[3]);
    if (gen(argv[1], pa, 2) != 0) return 1;
    if (gen(argv[2], pb, 1) != 0) return 1;
    if (gen(argv[1], pa2, 1) != 0) return 1;
    /* a fourth and fifth model context in a row must not accumulate engines (idle engines are bounded) */
    if (gen(argv[2], pb, 1) != 0) return 1;
    if (gen(argv[1], pa2, 1) != 0) return 1;
    const int alive = ptts_b200_host_engine_count();
    printf("final engines alive = %d\n", alive);
    return alive <= 2 ? 0 : 3;
}
