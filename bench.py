#!/usr/bin/env python
"""bench.py - audio-seconds per second of the Pocket-TTS generation hot path on N B200s.

Workload (BASELINE.json configs[2], the configuration the `metric` is quoted on): batched
generation of 256 synthetic ~100-character prompts per GPU, voice "alba" (125-frame synthetic
prefix), 100 frames (8 s of 24 kHz audio) per prompt, temp 0.7, EOS disabled, LSD steps 1,
random-init weights of the named architecture (no network for the HF checkpoint).
A "step" = tokens -> FlowLM lock-step decode -> flow head -> Mimi decoder -> PCM for one batch.

  value : inputs already resident in HBM (tokens + noise staged), kernels only, CUDA events on the
          engine stream, max over ranks.
  e2e   : the same job through the public C-ABI call sequence with HOST buffers: stage (H2D) +
          run + fetch (D2H of PCM into pinned memory) inside the timed region.
  --impl reference : the UNMODIFIED reference (oracle/_ref/libptts_ref.so, built by oracle/Makefile
          with the flags of `make cpu-opt`) through its own ptts_generate on the host cores.

Multi-GPU: one process per GPU (torchrun), utterances sharded across ranks, weights replicated, no
collective on the data path (torch.distributed is used only for the barrier and the max-reduce of
the timings) -> "scaling": "weak".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BATCH = 256
FRAMES = 100
FRAME_SEC = 0.08
VOICE_FRAMES = 125
WORKLOAD = "batched generation: 256 synthetic ~100-char prompts x 100 frames per GPU, voice alba (configs[2])"


def src_sha16() -> str:
    """Identity of the kernel sources the loaded library was built from (stamped into profiles/*traffic*.json by
    tools/ncu_traffic.py): an ncu capture of another build is refused instead of silently going stale."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / "pocket-tts.c_b200" / "csrc").glob("*.cu*")):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


# parameters streamed once per FlowLM frame step at lsd_steps = 1, as bf16
FLOW_PARAMS_PER_STEP = 75_522_048 + 35_841 + (9_759_008 - 789_760)     # transformer + head/misc + flow net without the constant time embedding
MIMI_GEMM_FLOP_PER_FRAME = (201.3 + 323.8) * 1e6                        # transformer linears + SEANet / projections (window attention 16.4 M is separate)


def flow_step_kernel_work(tok_lens, frames, cond_len):
    """Algorithmic bytes / FLOPs of ONE launch of the persistent step kernel (all `frames` frame steps of the batch), section
    8(d) terms only: weights once per step (bf16), every utterance's own K/V rows read once per step (fp16 cache: e = 2), the
    new K/V row written, the shared voice-prefix K/V (fp32) read once per step. Split-K partial sums, activations and anything
    else the implementation moves are NOT algorithmic and are not counted."""
    B = len(tok_lens)
    own0 = np.asarray(tok_lens, np.float64) + 1.0            # rows the prompt prefill left in each own cache
    s = np.arange(frames, dtype=np.float64)
    rows = (own0[:, None] + s[None, :] + 1.0).sum()          # keys each utterance attends to in its own cache, summed over steps
    w = FLOW_PARAMS_PER_STEP * 2.0 * frames
    kv_read = 2.0 * rows * 1024 * 6 * 2.0
    kv_write = 2.0 * 1024 * 6 * 2.0 * B * frames
    prefix = 2.0 * cond_len * 1024 * 6 * 4.0 * frames
    flops = 2.0 * FLOW_PARAMS_PER_STEP * B * frames + 4.0 * 1024 * 6 * (rows + cond_len * B * frames)
    return dict(bytes=w + kv_read + kv_write + prefix, flops=flops, weights=w, kv_read=kv_read, kv_write=kv_write, prefix=prefix)


# ---- sharding of the work over the ranks (one process per GPU, no collective on the data path). Pure functions, also exercised
# by the world_size-2 gloo test on CPU (tests/test_dist_gloo.py).
def rank_slice(rank: int, per_rank: int):
    """Utterances [lo, hi) of a list of world * per_rank that rank `rank` processes (weak scaling: per_rank is fixed)."""
    return rank * per_rank, (rank + 1) * per_rank


def sweep_plan(total_prompts: int, world: int, sweep_batch: int):
    """configs[4]: a fixed list of total_prompts is dealt to the ranks (strong scaling) and cut into calls of at most sweep_batch
    utterances: returns (utterances per call, calls per rank)."""
    per_rank = total_prompts // world
    sb = min(sweep_batch, per_rank)
    return sb, max(1, per_rank // sb)


def reduce_max(value: float, world: int, device: str) -> float:
    """MAX over the ranks of a host scalar (device-timed milliseconds / seconds): the job is as slow as its slowest rank."""
    if world == 1:
        return value
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if "Active" in r[col] and "Not" not in r[col]:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


def tokenize_prompts(pkg, model_dir, prompts):
    """Host glue kept from the reference: ptts_prepare_text + SentencePiece (token ids are bit-exact
    with the reference tokenizer, tests/test_oracle_golden.py::test_tokenizer_ids_bit_exact)."""
    import sentencepiece as spm
    sp = spm.SentencePieceProcessor(model_file=str(model_dir / "tokenizer.model"))
    return [np.asarray(sp.encode(pkg.prepare_text(t)[0]), np.int32) for t in prompts]


def run_reference(args, rank):
    """The reference's own CPU implementation through its own public API on all host cores."""
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "oracle"))
    import pyoracle
    from __graft_entry__ import load_package
    pkg = load_package()
    md = pkg.synth.model_dir(0, VOICE_FRAMES)
    cores = os.cpu_count() or 1
    ref = pyoracle.Reference(md, threads=cores)
    prompts = pkg.synth.make_prompts(args.warmup + args.steps, seed=123)
    frames = FRAMES if args.steps <= 20 else 25
    for i in range(args.warmup):
        ref.generate(prompts[i], "alba", num_frames=frames, seed=123 + i, temp=0.7)
    t0 = time.perf_counter()
    for i in range(args.steps):
        pcm = ref.generate(prompts[args.warmup + i], "alba", num_frames=frames, seed=123 + i, temp=0.7)
        assert pcm.shape[0] == frames * 1920
    dt = time.perf_counter() - t0
    val = args.steps * frames * FRAME_SEC / dt
    sample = f"{args.steps} prompts x {frames} frames, sequential ptts_generate (reference has no batching)"
    print(json.dumps({
        "impl": "reference", "metric": "audio-sec/sec", "value": val, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fast", type=int, default=0, help="1: activations as single bf16 (precise=0); default hi+lo pairs")
    ap.add_argument("--skip-mimi-only", action="store_true", help="skip the configs[3] Mimi decode-only measurement")
    ap.add_argument("--ragged", action="store_true", default=True,
                    help="also time the ragged variants (SURVEY.md 8d): one lock-step batch whose frame counts are drawn from U[60, 140], "
                         "and a ragged prompt list through the product's length-bucketing pool driver (default: on)")
    ap.add_argument("--no-ragged", dest="ragged", action="store_false")
    ap.add_argument("--total-prompts", type=int, default=8192,
                    help="configs[4]: a FIXED list of this many prompts (sharded over the ranks, calls of --sweep-batch utterances) through the "
                         "host-buffer API, so that the line carries a strong-scaling point at every N (0 disables)")
    ap.add_argument("--sweep-batch", type=int, default=1024, help="utterances per call in the --total-prompts sweep (the engine takes up to 1024)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries exactly ONE JSON line: everything libraries print there (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        return reduce_max(v, world, "cuda")

    if rank == 0:
        md = pkg.synth.model_dir(0, VOICE_FRAMES)
    barrier()
    md = pkg.synth.model_dir(0, VOICE_FRAMES)
    W = pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors")
    voice = pkg.synth.read_safetensors(md / "embeddings" / "alba.safetensors")["audio_prompt"].reshape(-1, 1024)
    eng = pkg.Engine(W, local_rank)
    del W

    B, F = args.batch, args.frames
    lo, hi = rank_slice(rank, B)
    prompts = pkg.synth.make_prompts(B * world, seed=123)[lo:hi]
    toks = tokenize_prompts(pkg, md, prompts)
    seeds = np.arange(B, dtype=np.int64) + 123 + rank * B
    params = pkg.default_params(lsd_steps=1, temp=0.7, eos_enabled=0, precise=0 if args.fast else 1)
    audio = torch.empty((B, F * 1920), dtype=torch.float32).pin_memory().numpy()   # pinned host PCM buffer
    audio_sec = B * F * FRAME_SEC

    # ---------------- warm-up (also builds the step graph, tensor maps and workspaces)
    eng.stage(toks, voice, F, seeds, None, params)
    for _ in range(args.warmup):
        eng.run(True)
    eng.sync()

    # ---------------- timed region 1: inputs resident in HBM, kernels only
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    l0 = eng.launch_count()
    eng.timer_start()
    for _ in range(args.steps):
        eng.run(True)
    ms = eng.timer_stop()
    launches = eng.launch_count() - l0
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = world * audio_sec / (ms_per_step / 1000.0)

    # ---------------- timed region 2: end to end through the host-buffer C ABI
    eng.generate_batch(toks, voice, F, seeds, None, params, True, audio)   # warm
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = eng.generate_batch(toks, voice, F, seeds, None, params, True, audio)
    eng.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    assert np.isfinite(out["audio"][0, :1920]).all() and int(out["frames"].min()) == F
    e2e_val = world * audio_sec * args.steps / e2e_s
    h2d = int(sum(len(t) for t in toks) * 4 + B * F * 32 * 4 + voice.nbytes * 0 + 8 * B * 4)
    d2h = int(B * F * 1920 * 4 + B * F * 32 * 4 + B * 4)

    line = {
        "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16 weights / bf16 hi+lo or fp16 activations / f32 accumulate",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "frames": F, "voice_frames": int(voice.shape[0]),
                   "lsd_steps": 1, "temp": 0.7, "eos": "disabled", "precise": int(not args.fast),
                   "l2": "inputs larger than L2: per-step working set (KV caches + Mimi workspace) >> 126 MB",
                   "parallelism": f"utterance sharding x{world}, weights replicated, no collective"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "realtime_factor_per_gpu": value / world,
    }

    if args.ragged:
        # ---------------- ragged variant: per-utterance frame counts U[60, 140]; the lock-step loop runs until the longest
        # utterance is done (finished utterances keep their batch slot), audio seconds count the frames actually produced
        rf = np.random.default_rng(99 + rank).integers(60, 141, size=B).astype(np.int32)
        raudio = torch.empty((B, int(rf.max()) * 1920), dtype=torch.float32).pin_memory().numpy()
        eng.stage(toks, voice, rf, seeds, None, params)
        eng.run(True); eng.sync()
        barrier()
        eng.timer_start()
        for _ in range(args.steps):
            eng.run(True)
        rms = max_over_ranks(eng.timer_stop()) / args.steps
        out = eng.fetch(True, True, raudio)
        assert (out["frames"] == rf).all()
        line["ragged"] = {"workload": f"{B} prompts per GPU, frames ~ U[60, 140] (mean {float(rf.mean()):.1f}, max {int(rf.max())})",
                          "ms_per_step": rms, "audio_s_per_s": world * float(rf.sum()) * FRAME_SEC / (rms / 1000.0),
                          "note": "inputs resident in HBM; value counts produced frames only"}
        eng.stage(toks, voice, F, seeds, None, params)   # restore the uniform batch for the profile section below

    if args.total_prompts > 0:
        # ---------------- configs[4]: a long prompt list sharded over the ranks, every batch through the host-buffer call
        # (tokens and noise up, PCM down), different prompts in every batch; tokenisation is outside the timed region.
        SB, nb = sweep_plan(args.total_prompts, world, args.sweep_batch)
        lo, hi = rank_slice(rank, nb * SB)
        allp = pkg.synth.make_prompts(nb * SB * world, seed=777)[lo:hi]
        alltok = tokenize_prompts(pkg, md, allp)
        sseeds = np.arange(SB, dtype=np.int64) + 9000 + rank * nb * SB
        saudio = torch.empty((SB, F * 1920), dtype=torch.float32).pin_memory().numpy()
        # capacities (KV rows, token slots, prefill rows) are reserved once, outside the timed region: warm up on the SB longest
        # prompts of this rank's list, so that no later call has to grow the caches and rebuild the step program
        longest = sorted(range(len(alltok)), key=lambda i: -len(alltok[i]))[:SB]
        eng.generate_batch([alltok[i] for i in longest], voice, F, sseeds, None, params, True, saudio)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for j in range(nb):
            eng.generate_batch(alltok[j * SB:(j + 1) * SB], voice, F, sseeds + j * SB, None, params, True, saudio)
        eng.sync()
        sweep_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        B_, B = B, SB   # (the summary below reports the sweep's batch size)
        line["sweep"] = {"workload": f"throughput sweep: {nb * B * world} synthetic prompts x {F} frames sharded across {world} B200 (configs[4])",
                         "prompts": nb * B * world, "batches_per_gpu": nb, "seconds": sweep_s,
                         "audio_s_per_s": nb * B * world * F * FRAME_SEC / sweep_s,
                         "batch_per_call": SB,
                         "note": "host token lists -> host PCM through ptts_b200 generate_batch, H2D / D2H inside"}
        B = B_
        if args.ragged:
            # ---------------- the same kind of list with RAGGED lengths (frames ~ U[60, 140]) through ptts_b200_pool_generate: the
            # product's host driver sorts the utterances by length and cuts the sorted list into calls, so every lock-step call
            # (and its Mimi decode) runs only to its own longest utterance. Efficiency = produced audio-s/s vs the uniform sweep.
            nr = min(nb, 8) * SB
            rfr = np.random.default_rng(4242 + rank).integers(60, 141, size=nr).astype(np.int32)
            Wp = pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors")
            pool = pkg.Pool(Wp, [local_rank])
            del Wp
            rpcm = torch.empty((nr, int(rfr.max()) * 1920), dtype=torch.float32).pin_memory().numpy()
            rseeds = np.arange(nr, dtype=np.int64) + 70000 + rank * nr
            pool.generate(alltok[:nr], voice, rfr, rseeds, None, params, SB, True, rpcm, False)   # shapes / capacities
            barrier()
            t0 = time.perf_counter()
            out = pool.generate(alltok[:nr], voice, rfr, rseeds, None, params, SB, True, rpcm, False)
            rs = max_over_ranks(time.perf_counter() - t0)
            assert (out["frames"] == rfr).all()
            pool.close()
            uniform_rate = line["sweep"]["audio_s_per_s"]
            ragged_rate = world * float(rfr.sum()) * FRAME_SEC / rs
            line["ragged_sweep"] = {"workload": f"{nr * world} prompts, frames ~ U[60, 140], ptts_b200_pool_generate (length bucketing, calls of <= {SB})",
                                    "seconds": rs, "audio_s_per_s": ragged_rate, "vs_uniform_sweep": ragged_rate / uniform_rate,
                                    "note": "host token lists -> host PCM; audio seconds count produced frames only"}
            barrier()
        eng.stage(toks, voice, F, seeds, None, params)   # restore the staged configs[2] batch for the profile section below

    if rank == 0:
        # ---------------- per-class profile of one step. The FlowLM decode loop is ONE launch (k_flow_mega, the
        # persistent step kernel): its classes are timed inside the kernel (%globaltimer stamps of CTA 0 around every
        # phase, grid barrier included); prefill and Mimi kernels are bracketed by CUDA events on the engine stream.
        pk = peaks()
        eng.profile(True)
        eng.run(True)
        prof = eng.profile_read()
        eng.profile(False)
        tot = sum(v["ms"] for v in prof.values()) or 1.0
        shares = {k: round(v["ms"] / tot, 4) for k, v in prof.items() if v["launches"]}
        line["kernel_shares"] = shares
        line["kernel_classes"] = {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                      "TFLOPs": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2),
                                      "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in prof.items() if v["launches"]}
        # dominant kernel = the persistent FlowLM step kernel (ONE launch per step). Timed alone: CUDA events recorded on the engine
        # stream immediately around that launch (ptts_b200_step_kernel_ms: prompt prefill and Mimi excluded), averaged over 3 runs.
        eng.stage(toks, voice, F, seeds, None, params)
        eng.run(False); eng.sync()
        mk = []
        for _ in range(3):
            eng.run(False)
            ms_k, steps_k = eng.step_kernel_ms()
            assert steps_k == F
            mk.append(ms_k)
        ms_flow = sum(mk) / len(mk)
        work = flow_step_kernel_work([len(t) for t in toks], F, int(voice.shape[0]))
        mem_frac = work["bytes"] / (ms_flow / 1e3) / 1e9 / pk["hbm"]
        flop_frac = work["flops"] / (ms_flow / 1e3) / 1e12 / pk["tf_sust"]
        traffic, traffic_note = None, "no ncu --set full capture committed for this build"
        sha = src_sha16()
        for tr in sorted((ROOT / "profiles").glob("r*_flow_mega_traffic.json"), reverse=True):
            try:
                tj = json.loads(tr.read_text())
            except Exception:
                continue
            if tj.get("src_sha16") == sha:
                traffic, traffic_note = tj.get("dram_bytes_per_launch"), f"{tr.name} (ncu --set full, same kernel sources {sha})"
                break
            traffic_note = f"{tr.name} is from other kernel sources ({tj.get('src_sha16')} != {sha}): refused as stale"
        roof = {"bound": "hbm", "achieved": work["bytes"] / (ms_flow / 1e3) / 1e9, "peak": pk["hbm"], "unit": "GB/s", "frac": mem_frac,
                "kernel": "k_flow_mega (persistent FlowLM step kernel: all frame steps of the batch in one launch)",
                "launches_per_step": 1, "avg_launch_us": 1000.0 * ms_flow, "us_per_frame_step": 1000.0 * ms_flow / F,
                "peak_source": pk["src"] + " (sustained: the kernel runs for tens of ms)", "traffic": traffic, "traffic_source": traffic_note,
                "share_of_step": round(ms_flow / ms_per_step, 4), "tensor_frac": flop_frac,
                "algorithmic_bytes_per_launch": work["bytes"], "algorithmic_flops_per_launch": work["flops"],
                "algorithmic_bytes_terms": {k: work[k] for k in ("weights", "kv_read", "kv_write", "prefix")},
                "timing": "CUDA events on the engine stream around the k_flow_mega launch only, mean of 3 launches"}
        line["roofline"] = roof
        # second kernel family of the step: the Mimi tcgen05 kernels (implicit-GEMM convolutions, linears, fused residual blocks)
        if "gemm_tc_mimi" in prof and prof["gemm_tc_mimi"]["launches"]:
            m = prof["gemm_tc_mimi"]
            mflop = MIMI_GEMM_FLOP_PER_FRAME * B * F
            line["roofline_mimi_gemm"] = {"bound": "tensor", "achieved": mflop / (m["ms"] / 1e3) / 1e12, "peak": pk["tf_sust"],
                                          "unit": "TFLOP/s", "frac": mflop / (m["ms"] / 1e3) / 1e12 / pk["tf_sust"], "traffic": None,
                                          "kernel": "gemm_tc_kernel + resblock_kernel (all Mimi tcgen05 launches of one step)",
                                          "algorithmic_flops_per_step": mflop, "issued_flops_per_step": m["flops"],
                                          "launches_per_step": m["launches"], "ms_per_step": m["ms"],
                                          "avg_launch_us": 1000.0 * m["ms"] / m["launches"], "share_of_step": shares.get("gemm_tc_mimi"),
                                          "timing": "CUDA events around every launch of the class (profile mode, one step)"}

        # ---------------- configs[3]: Mimi decode-only, 64 x 4096 latent frames -> 24 kHz PCM. Two numbers: host latents -> host PCM
        # through ptts_b200_mimi_decode_batch (H2D + 2 GB of D2H inside), and the device-resident decode (latents staged, PCM left in
        # HBM) timed with CUDA events. The whole output is checked for finiteness; tests/test_gpu_measured_configs.py checks slices
        # of 4096-frame items against the oracle.
        if not args.skip_mimi_only:
            lat = np.random.default_rng(0).standard_normal((64, 4096, 32)).astype(np.float32)
            pcm = torch.empty((64, 4096 * 1920), dtype=torch.float32).pin_memory().numpy()
            eng.mimi_decode_batch(lat[:2], pcm[:2])
            t0 = time.perf_counter()
            eng.mimi_decode_batch(lat, pcm)
            dt = time.perf_counter() - t0
            finite = bool(np.isfinite(pcm).all()) and float(np.abs(pcm).max()) > 0.0
            assert finite, "configs[3]: non-finite or all-zero PCM"
            eng.mimi_stage(lat)
            eng.mimi_run(); eng.sync()
            eng.timer_start()
            eng.mimi_run()
            dms = eng.timer_stop()
            mflop4 = (MIMI_GEMM_FLOP_PER_FRAME + 16.4e6) * 64 * 4096
            line["mimi_decode_only"] = {"workload": "64 x 4096 latent frames (configs[3])", "seconds": dt, "audio_s_per_s": 64 * 4096 * FRAME_SEC / dt,
                                        "device_resident_seconds": dms / 1e3, "device_resident_audio_s_per_s": 64 * 4096 * FRAME_SEC / (dms / 1e3),
                                        "roofline": {"bound": "tensor", "achieved": mflop4 / (dms / 1e3) / 1e12, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                                                     "frac": mflop4 / (dms / 1e3) / 1e12 / pk["tf_sust"], "traffic": None},
                                        "output_finite": finite,
                                        "note": "seconds: host latents -> host PCM (H2D + 2 GB D2H inside); device_resident: ptts_b200_mimi_run between CUDA events"}
            del lat, pcm

        # ---------------- batch-1 streaming decode (configs[1]) and first-audio latency
        one = [toks[0]]
        eng.generate_batch(one, voice, 1, [123], None, params, True)           # warm (graph capture)
        lat = []
        for _ in range(5):
            t0 = time.perf_counter()
            eng.generate_batch(one, voice, 1, [123], None, params, True)
            lat.append(1000 * (time.perf_counter() - t0))
        lat_stream = []
        for _ in range(5):
            t0 = time.perf_counter()
            eng.stream_begin(one[0], voice, 200, 123, params)
            pcm1, _done = eng.stream_next(1)
            lat_stream.append(1000 * (time.perf_counter() - t0))
            assert len(pcm1) == 1920
        eng.stage(one, voice, 200, [123], None, params)
        eng.run(False)
        eng.sync()
        eng.timer_start()
        eng.run(False)
        ms1 = eng.timer_stop()
        steps1 = 200   # decode iterations; the prompt is prefilled in one batched pass
        line["batch1"] = {"first_audio_ms": statistics.median(lat), "first_audio_stream_api_ms": statistics.median(lat_stream), "first_audio_note": "tokens on host -> first 1920 PCM samples on host, voice prefix KV cached",
                          "decode_200_frames_ms": ms1, "us_per_step": 1000 * ms1 / steps1,
                          "audio_s_per_s": 200 * FRAME_SEC / (ms1 / 1000),
                          "roofline": {"bound": "hbm", "achieved": FLOW_PARAMS_PER_STEP * 2.0 / (ms1 / steps1 / 1e3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                       "frac": FLOW_PARAMS_PER_STEP * 2.0 / (ms1 / steps1 / 1e3) / 1e9 / pk["hbm"], "traffic": None,
                                       "kernel": "one CUDA graph of GEMV / attention kernels per frame step (weights streamed once per step)"}}

        # ---------------- CPU baseline: the compiled reference on the host cores, bounded sample
        if world == 1 and not args.no_cpu_baseline:
            try:
                sys.path.insert(0, str(ROOT / "oracle"))
                import pyoracle
                cores = os.cpu_count() or 1
                ref = pyoracle.Reference(md, threads=cores)
                nprompt = 2
                t0 = time.perf_counter()
                for i in range(nprompt):
                    ref.generate(prompts[i], "alba", num_frames=F, seed=int(seeds[i]), temp=0.7)
                dt = time.perf_counter() - t0
                ref.close()
                line["cpu_baseline"] = {"value": nprompt * F * FRAME_SEC / dt, "unit": "audio-s/s", "cores": cores, "kind": "reference",
                                        "sample": f"{nprompt} of the {B} prompts x {F} frames, sequential ptts_generate, OpenMP {cores} threads"}
            except Exception as ex:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"unavailable: {ex}"}
        # ---------------- second baseline: the reference's OWN GPU backend (`make cuda-opt`: ptts_cuda.c on cuBLAS + NVRTC kernels, unmodified,
        # built by oracle/Makefile) on this same B200: one prompt x 100 frames, timings as the reference prints them (PTTS_TIMING=1)
        if world == 1 and not args.no_cpu_baseline:
            exe = ROOT / "oracle" / "_ref" / "ptts_ref_cuda"
            try:
                import re
                wav = Path(tempfile.gettempdir()) / "ptts_ref_cuda.wav"
                cmd = [str(exe), "-d", str(md), "-p", prompts[0], "--voice", "alba", "-t", "0.7", "-S", "123", "--frames", str(F), "--eos-threshold", "1e9",
                       "-o", str(wav)]
                env = dict(os.environ, PTTS_TIMING="1")
                t0 = time.perf_counter()
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
                wall = time.perf_counter() - t0
                fm = re.search(r"FlowLM latents: ([0-9.]+) ms", r.stderr)
                mm = re.search(r"Mimi decode: ([0-9.]+) ms", r.stderr)
                if r.returncode == 0 and fm and mm:
                    gen_s = (float(fm.group(1)) + float(mm.group(1))) / 1e3
                    line["reference_cuda_baseline"] = {"value": F * FRAME_SEC / gen_s, "unit": "audio-s/s", "flowlm_ms": float(fm.group(1)), "mimi_ms": float(mm.group(1)),
                                                       "wall_s_incl_load_and_nvrtc": wall, "kind": "reference (make cuda-opt: cuBLAS + NVRTC, unmodified)",
                                                       "sample": f"1 prompt x {F} frames, single utterance (the reference has no batching)"}
                else:
                    line["reference_cuda_baseline"] = {"value": None, "sample": f"failed: rc={r.returncode} {r.stderr[-200:]}"}
            except Exception as ex:
                line["reference_cuda_baseline"] = {"value": None, "sample": f"unavailable: {ex}"}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
