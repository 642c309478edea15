"""Per-phase timeline of the persistent FlowLM step kernel: python tools/mega_trace.py [B] [F]  (PTTS_B200_MEGA_TRACE=1)."""
import os, sys
from pathlib import Path
import numpy as np
os.environ["PTTS_B200_MEGA_TRACE"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100
md = pkg.synth.model_dir(0)
W = pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors")
voice = pkg.synth.read_safetensors(md / "embeddings" / "alba.safetensors")["audio_prompt"].reshape(-1, 1024)
eng = pkg.Engine(W, 0)
import sentencepiece as spm
sp = spm.SentencePieceProcessor(model_file=str(md / "tokenizer.model"))
toks = [np.asarray(sp.encode(pkg.prepare_text(t)[0]), np.int32) for t in pkg.synth.make_prompts(B, seed=123)]
eng.stage(toks, voice, F, np.arange(B) + 123, None, pkg.default_params(temp=0.7))
eng.run(False); eng.sync()
eng.profile(True)
eng.run(False)
print({k: v for k, v in eng.profile_read().items() if v["launches"]})
eng.profile(False)
eng.close()
