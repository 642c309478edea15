/*
 * ptts_b200.h - C ABI of the B200 (sm_100a) engine for Pocket-TTS's per-frame generation path.
 *
 * Plain C, plain pointers and sizes; no CUDA / torch types.  This library replaces ptts_cuda.c
 * (reference ptts_cuda.h:9-96) as the dispatch target of ptts_flowlm.c and ptts_mimi.c, but it is
 * entered at whole-loop granularity: the per-op host-pointer ABI of
 * ptts_cuda.h is itself the bottleneck (one blocking H2D/D2H round trip per GEMV).
 *
 * Conventions
 *   - every function returns 0 on success and -1 on failure; the message is available from
 *     ptts_b200_last_error() (thread-local), mirroring ptts_get_error() (ptts.h:52).
 *   - there is NO CPU fallback: if no CUDA device / kernel image is available the call fails.
 *   - host pointers are owned by the caller; the engine owns all device memory.
 *   - one engine per device; an engine is single-threaded; distinct engines may be driven from
 *     distinct host threads (utterances shard across GPUs with no collective).
 */
#ifndef PTTS_B200_H
#define PTTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptts_b200_engine ptts_b200_engine;

/* Host weight table entry.  dtype follows safetensor_dtype_t (ptts_safetensors.h:20-28):
 * 0 = F32, 1 = F16, 2 = BF16.  Names are checkpoint tensor names; lookup order is the reference's
 * (exact, then "flow_lm." / "mimi." / "model." prefix, then suffix: ptts_flowlm.c:107-122,
 * ptts_mimi.c:77-95). */
typedef struct {
    const char *name;
    const void *data;
    int dtype;
    int64_t numel;
} ptts_b200_tensor;

/* Engine lifetime (replaces per-call ptts_flowlm_load / ptts_mimi_load, ptts.c:1047-1053, and the
 * host-pointer-keyed device weight cache of ptts_cuda.c:617-644).  Weights are converted once to
 * bf16 (FlowLM) / fp16 (Mimi) device copies. */
int ptts_b200_create(ptts_b200_engine **out, int device, const ptts_b200_tensor *tensors, int n_tensors);
void ptts_b200_destroy(ptts_b200_engine *e);
const char *ptts_b200_last_error(void);

/* ---- drop-in for ptts_flowlm_generate_latents (ptts_flowlm.h:43-51, ptts_flowlm.c:1138-1251).
 * Same arguments after the engine handle, same outputs (raw latents [frames,32], frames used,
 * optional first EOS logit / first cond[1024] / first flow[32]).  The voice prefix KV is cached
 * per engine and reused while cond_prefix is unchanged. fp32 activations, bf16 weights. */
int ptts_b200_flowlm_generate_latents(ptts_b200_engine *e, const int *tokens, int token_len,
                                      const float *cond_prefix, int cond_len,
                                      int max_frames, int lsd_steps, float temp, float noise_clamp,
                                      int64_t seed, int eos_enabled, float eos_threshold,
                                      int eos_min_frames, int eos_after,
                                      float *out_latents, int *out_frames_used,
                                      float *out_first_eos_logit, float *out_first_cond,
                                      float *out_first_flow);

/* ---- drop-in for ptts_flowlm_scale_latents (ptts_flowlm.h:54-55, ptts_flowlm.c:1253-1263). */
int ptts_b200_flowlm_scale_latents(ptts_b200_engine *e, const float *in_latents, int frames, float *out_latents);

/* ---- drop-in for ptts_mimi_decode (ptts_mimi.h:26-27, ptts_mimi.c:570-738):
 * latents [frames,32] (already scaled) -> out_audio [frames*1920], *out_len = frames*1920. */
int ptts_b200_mimi_decode(ptts_b200_engine *e, const float *latents, int frames, float *out_audio, int *out_len);

/* ---- drop-in for ptts_mimi_forward_one (ptts_mimi.h:20, ptts_mimi.c:550-564), the parity hook behind the reference CLI's
 * --mimi-test (main.c:444): quantizer projection of one scaled latent[32] + the decoder transformer over that single token
 * -> out_embed[512]. */
int ptts_b200_mimi_forward_one(ptts_b200_engine *e, const float *latent, float *out_embed);

/* ---- additive: voice-cloning encoder. 24 kHz mono PCM -> conditioning rows out_cond [frames, 1024]
 * at 12.5 Hz (frames = ceil(n_samples / 1920), the input is zero padded to whole frames), i.e. the tensor the reference loads
 * from a voice file as "audio_prompt" (ptts.c:293-364) and passes to ptts_flowlm_generate_latents as cond_prefix. The reference
 * itself only shape-checks the tensors involved (ptts.c:914-944: encoder.model.*, encoder_transformer.*, downsample.conv) and
 * loads flow_lm.speaker_proj_weight without using it (ptts_flowlm.c:874); this entry runs them: SEANet encoder (causal Conv1d,
 * ptts_kernels.c:114-156; ResBlock, ptts_mimi.c:160-177) -> encoder transformer (ptts_mimi.c:301-376) -> stride-16 downsampler
 * -> speaker projection. Fails (-1) when the checkpoint has no encoder tensors or frames > max_frames. */
int ptts_b200_voice_encode(ptts_b200_engine *e, const float *pcm, int n_samples, float *out_cond, int max_frames, int *out_frames);

/* ---- additive batched entries (the reference is single-utterance, ptts.h:75-76). */

/* Mimi decode of n equal-length latent sequences: latents [n, frames, 32] (scaled) ->
 * out_audio [n, frames*1920]. */
int ptts_b200_mimi_decode_batch(ptts_b200_engine *e, const float *latents, int n, int frames, float *out_audio);

/* The same decode split in three, so that the middle part can be timed with the latents already resident in HBM and the PCM
 * left there (configs[3] device-resident timing): stage (H2D), run (kernels only, asynchronous on the engine stream),
 * fetch (D2H of [n, frames*1920]). mimi_decode_batch == stage + run (PCM chunks overlapped with the decode) + wait. */
int ptts_b200_mimi_stage(ptts_b200_engine *e, const float *latents, int n, int frames);
int ptts_b200_mimi_run(ptts_b200_engine *e);
int ptts_b200_mimi_fetch(ptts_b200_engine *e, float *out_audio);

typedef struct {
    int lsd_steps;        /* ptts_params.num_steps (ptts.h:29) */
    float temp;           /* ptts_params.temp */
    float noise_clamp;
    int eos_enabled;
    float eos_threshold;
    int eos_min_frames;
    int eos_after;        /* used when per-utterance eos_after[] is NULL */
    int precise;          /* 1: activations fed to the tensor cores as bf16 hi+lo pairs (default, the parity-tested form, runs in the
                             persistent step kernel); 0: hi only - a diagnostic form (3e-2 against the oracle) that runs on the
                             multi-kernel graph path */
} ptts_b200_gen_params;

/* Batched generation, tokens -> latents -> PCM, n utterances sharing one voice prefix.
 *   tokens_flat / token_offsets[n+1] : concatenated token ids
 *   max_frames[n], seeds[n]          : per utterance (seed -1 is NOT time-seeded here: caller decides)
 *   eos_after[n]                     : optional per-utterance override (NULL = params->eos_after)
 *   out_latents [n, frames_cap, 32]  : raw latents (optional, may be NULL)
 *   out_frames_used [n]
 *   out_audio [n, frames_cap*1920]   : PCM, zero beyond frames_used*1920 (optional, may be NULL)
 * frames_cap = max over max_frames[]. Host buffers; H2D/D2H copies happen inside the call.
 * n <= 1024 per call (larger jobs are sharded by the caller: more calls, more engines, more GPUs); returns -1 otherwise. */
int ptts_b200_generate_batch(ptts_b200_engine *e, int n, const int *tokens_flat, const int *token_offsets,
                             const float *cond_prefix, int cond_len, const int *max_frames,
                             const int64_t *seeds, const int *eos_after, const ptts_b200_gen_params *params,
                             float *out_latents, int *out_frames_used, float *out_audio);

/* The same job split in three so that the middle part can be timed with all inputs already
 * resident in HBM: stage (H2D of tokens / noise, state reset), run (kernels only, asynchronous on
 * the engine stream), fetch (D2H). generate_batch == stage + run + fetch. */
int ptts_b200_batch_stage(ptts_b200_engine *e, int n, const int *tokens_flat, const int *token_offsets,
                          const float *cond_prefix, int cond_len, const int *max_frames,
                          const int64_t *seeds, const int *eos_after, const ptts_b200_gen_params *params);
int ptts_b200_batch_run(ptts_b200_engine *e, int with_mimi);
int ptts_b200_batch_fetch(ptts_b200_engine *e, float *out_latents, int *out_frames_used, float *out_audio);

/* ---- additive multi-GPU entry: a pool = one engine per device, weights replicated, one host thread per
 * device inside pool_generate. The utterance list is sorted by max_frames (the caller's ptts_estimate_frames value, ptts.c:285-291),
 * cut into calls of at most max_batch utterances (<= 0: 1024; at least one call per device when there is enough work), the calls
 * are dealt to the devices longest-first, and latents / frames_used / PCM are gathered in the ORIGINAL order with the layout of
 * ptts_b200_generate_batch (frames_cap = max over max_frames[]). No collective and no peer traffic on the data path.
 * devices == NULL: every visible device. */
typedef struct ptts_b200_pool ptts_b200_pool;
int ptts_b200_pool_create(ptts_b200_pool **out, const int *devices, int n_devices, const ptts_b200_tensor *tensors, int n_tensors);
void ptts_b200_pool_destroy(ptts_b200_pool *p);
int ptts_b200_pool_size(const ptts_b200_pool *p);
ptts_b200_engine *ptts_b200_pool_engine(ptts_b200_pool *p, int i);   /* engine of pool slot i (single-utterance entries, measurement hooks) */
int ptts_b200_pool_generate(ptts_b200_pool *p, int n, const int *tokens_flat, const int *token_offsets, const float *cond_prefix,
                            int cond_len, const int *max_frames, const int64_t *seeds, const int *eos_after,
                            const ptts_b200_gen_params *params, int max_batch, float *out_latents, int *out_frames_used,
                            float *out_audio);

/* ---- additive streaming entry (roadmap item of the reference, README.md:162-167):
 * single utterance, audio handed out as it is generated, so first-audio latency is observable
 * through the API. begin = stage + prompt prefill; next = up to want_frames more FlowLM frames, then
 * the Mimi decoder over those frames ONLY, with carried state: per-layer K / V caches of the 250-token
 * attention window (RoPE at absolute positions), the previous latent frame for the upsampler and the 8
 * transformer-output rows the causal conv tails of the SEANet stack are rebuilt from. A call costs
 * want_frames (+ 0.5) frames of decoder work wherever it is in the utterance, and the PCM equals the
 * one-shot decode. out_pcm receives *out_frames * 1920 samples; *out_done != 0 after the last frame
 * (EOS rule or max_frames). */
int ptts_b200_stream_begin(ptts_b200_engine *e, const int *tokens, int token_len, const float *cond_prefix, int cond_len,
                           int max_frames, int64_t seed, const ptts_b200_gen_params *params);
int ptts_b200_stream_next(ptts_b200_engine *e, int want_frames, float *out_pcm, int *out_frames, int *out_done);

/* ---- measurement hooks (CUDA events on the engine's own stream). */
int ptts_b200_sync(ptts_b200_engine *e);
int ptts_b200_timer_start(ptts_b200_engine *e);
int ptts_b200_timer_stop(ptts_b200_engine *e, float *out_ms);      /* syncs, returns elapsed ms */
/* Duration of the most recent launch of the persistent FlowLM step kernel (k_flow_mega: all frames of the last batched run),
 * CUDA events recorded on the engine stream immediately around that one launch (prompt prefill and Mimi excluded), and the
 * number of frame steps it executed. Waits for the launch to finish. -1 if no batched run has happened yet. */
int ptts_b200_step_kernel_ms(ptts_b200_engine *e, float *out_ms, int *out_steps);
/* Per-kernel-class profiling: when enabled, launches bypass CUDA graphs and each launch is
 * bracketed by events. classes: see ptts_b200_profile_class_name. */
int ptts_b200_profile_enable(ptts_b200_engine *e, int on);
int ptts_b200_profile_read(ptts_b200_engine *e, int max_classes, float *ms_per_class, int64_t *launches_per_class,
                           int *n_classes);
/* Algorithmic work (FLOPs, bytes) issued per class while profiling, from launch shapes (host-side
 * accounting): the numerators of the roofline fractions. */
int ptts_b200_profile_work(ptts_b200_engine *e, int max_classes, double *flops, double *bytes);
const char *ptts_b200_profile_class_name(int cls);
/* Total kernel launches issued by the engine since creation (graph nodes count per replay). */
int64_t ptts_b200_launch_count(ptts_b200_engine *e);

/* ---- unit-test hook: one tcgen05 GEMM  D[m,n] = sum_k A[m,k] B[n,k]  on device copies of the
 * host arrays (A, B given as f32, rounded to bf16 (bf16 != 0) or fp16 inside).  `split`:
 * 0 none, 1 A carried as hi+lo, 2 B carried as hi+lo.  transposed != 0 stores out[n*M + m].
 * taps > 1 runs the implicit-GEMM conv form: A is [batches, rows, K/taps] and tap j reads row
 * m + shift0 + j*step (zero outside [0, rows)), B is [N, K]. */
int ptts_b200_test_gemm(ptts_b200_engine *e, const float *A, const float *B, float *out,
                        int batches, int rows, int N, int K, int taps, int shift0, int step,
                        int bf16, int split, int transposed, int bn, int split_k);

#ifdef __cplusplus
}
#endif
#endif /* PTTS_B200_H */
