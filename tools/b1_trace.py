"""Phase-class timing of the batch-1 persistent kernel: python tools/b1_trace.py [frames]"""
import sys
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
toks = [np.asarray(sp.encode(pkg.prepare_text(t)[0]), np.int32) for t in pkg.synth.make_prompts(1, seed=123)]
eng.stage(toks, voice, F, [123], None, pkg.default_params(temp=0.7))
eng.run(False); eng.sync()
eng.timer_start(); eng.run(False); print("decode ms", eng.timer_stop(), "frames", F)
eng.profile(True)
eng.run(False)
for k, v in eng.profile_read().items():
    if v["launches"]: print(k, "ms %.3f" % v["ms"], "phases", v["launches"], "us/phase %.2f" % (1000 * v["ms"] / v["launches"]))
eng.profile(False)
eng.close()
