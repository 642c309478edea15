"""Time of the persistent FlowLM step kernel alone: python tools/step_time.py [B] [F] [reps]  (CUDA events around the launch)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
md = pkg.synth.model_dir(0)
W = pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors")
voice = pkg.synth.read_safetensors(md / "embeddings" / "alba.safetensors")["audio_prompt"].reshape(-1, 1024)
eng = pkg.Engine(W, 0)
import sentencepiece as spm
sp = spm.SentencePieceProcessor(model_file=str(md / "tokenizer.model"))
toks = [np.asarray(sp.encode(pkg.prepare_text(t)[0]), np.int32) for t in pkg.synth.make_prompts(B, seed=123)]
eng.stage(toks, voice, F, np.arange(B) + 123, None, pkg.default_params(temp=0.7))
ms = []
for _ in range(reps + 1):
    eng.run(False); eng.sync()
    ms.append(eng.step_kernel_ms()[0])
print(f"B={B} F={F}: k_flow_mega ms per launch {['%.2f' % m for m in ms[1:]]}  us per frame {min(ms[1:]) * 1e3 / F:.1f}")
eng.close()
