"""Small job for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of the engine once -
batch-1 decode, batched decode through the persistent step kernel (grid barrier, mbarrier pipelines, TMEM hand-offs, TMA rings),
Mimi one-shot and streaming decode. Usage: compute-sanitizer --tool X python tools/sanitize_job.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
md = pkg.synth.model_dir(0)
W = pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors")
voice = pkg.synth.read_safetensors(md / "embeddings" / "alba.safetensors")["audio_prompt"].reshape(-1, 1024)[:40]
eng = pkg.Engine(W, 0)
ids = np.array([3, 3976, 66, 7, 14, 9, 21], np.int32)
g = eng.flowlm_generate_latents(ids, voice, 3, lsd_steps=1, temp=0.7)
pcm = eng.mimi_decode(eng.scale_latents(g["latents"]))
toks = [ids[: 3 + (i % 5)] for i in range(16)]
b = eng.generate_batch(toks, voice, [8 - (i % 3) for i in range(16)], seeds=list(range(16)), params=pkg.default_params(temp=0.7))
chunks = list(eng.stream(ids, voice, 3, seed=5, params=pkg.default_params(temp=0.7), chunk=2))
assert np.isfinite(pcm).all() and np.isfinite(b["audio"]).all() and sum(len(c) for c in chunks) == 3 * 1920
print("sanitize job ok: launches", eng.launch_count())
eng.close()
