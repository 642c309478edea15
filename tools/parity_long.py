"""Long-horizon parity of the batched path against the CPU oracle: python tools/parity_long.py [frames] [n]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
from __graft_entry__ import load_package  # noqa: E402
from conftest import max_rel  # noqa: E402
import pyoracle  # noqa: E402
pkg = load_package()
F = int(sys.argv[1]) if len(sys.argv) > 1 else 48
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
md = pkg.synth.model_dir(0)
W = pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors")
voice = pkg.synth.read_safetensors(md / "embeddings" / "alba.safetensors")["audio_prompt"].reshape(-1, 1024)
eng = pkg.Engine(W, 0)
import sentencepiece as spm
sp = spm.SentencePieceProcessor(model_file=str(md / "tokenizer.model"))
toks = [np.asarray(sp.encode(pkg.prepare_text(t)[0]), np.int32) for t in pkg.synth.make_prompts(n, seed=7)]
out = eng.generate_batch(toks, voice, F, seeds=np.arange(n) + 50, params=pkg.default_params(temp=0.7), want_audio=False)
orc = pyoracle.Oracle(W)
worst = 0.0
for i in range(n):
    o = orc.flowlm_generate(toks[i], voice, F, lsd_steps=1, temp=0.7, seed=50 + i)
    errs = [max_rel(out["latents"][i, :k], o["latents"][:k]) for k in (8, 16, 32, F)]
    worst = max(worst, errs[-1])
    print(i, " ".join(f"{e:.2e}" for e in errs))
print("worst max_rel over", F, "frames:", f"{worst:.2e}")
eng.close()
