"""Batch-1 decode time per frame step: python tools/b1_time.py [frames]  (PTTS_B200_L2_PERSIST_MB=N for the L2 set-aside)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
F = int(sys.argv[1]) if len(sys.argv) > 1 else 200
md = pkg.synth.model_dir(0)
W = pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors")
voice = pkg.synth.read_safetensors(md / "embeddings" / "alba.safetensors")["audio_prompt"].reshape(-1, 1024)
eng = pkg.Engine(W, 0)
import sentencepiece as spm
sp = spm.SentencePieceProcessor(model_file=str(md / "tokenizer.model"))
tok = np.asarray(sp.encode(pkg.prepare_text(pkg.synth.make_prompts(1, seed=5)[0])[0]), np.int32)
eng.stage([tok], voice, F, np.array([7]), None, pkg.default_params(temp=0.7))
eng.run(False); eng.sync()
ms = []
for _ in range(5):
    eng.sync(); eng.timer_start(); eng.run(False); ms.append(eng.timer_stop())
print(f"batch-1 decode, {F} frames (prefill included): ms {['%.2f' % m for m in ms]}  -> {min(ms) * 1e3 / F:.1f} us per frame step")
eng.close()
