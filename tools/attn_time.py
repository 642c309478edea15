"""Per-class kernel times of one batched step (profile mode): python tools/attn_time.py [B] [frames]   (A/B of the attention kernels)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100
md = pkg.synth.model_dir(0, 125)
W = pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors")
voice = pkg.synth.read_safetensors(md / "embeddings" / "alba.safetensors")["audio_prompt"].reshape(-1, 1024)
eng = pkg.Engine(W, 0)
toks = bench.tokenize_prompts(pkg, md, pkg.synth.make_prompts(B, seed=7))
params = pkg.default_params(lsd_steps=1, temp=0.7, eos_enabled=False)
seeds = np.arange(B, dtype=np.int64)
audio = np.zeros((B, F * 1920), np.float32)
eng.generate_batch(toks, voice, F, seeds, None, params, True, audio)
eng.profile(True)
eng.generate_batch(toks, voice, F, seeds, None, params, True, audio)
pr = eng.profile_read()
print({k: round(v["ms"], 3) for k, v in pr.items() if v["launches"]})
eng.close()
