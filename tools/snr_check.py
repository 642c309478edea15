"""PCM SNR of the engine against the committed reference golden vectors (margin over the 40 dB bar)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from __graft_entry__ import load_package  # noqa: E402
from conftest import snr_db  # noqa: E402
pkg = load_package()
md = pkg.synth.model_dir(0)
eng = pkg.Engine(pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors"), 0)
g = np.load(ROOT / "tests" / "golden" / "reference_vectors.npz")
print("mimi vs reference golden: %.1f dB" % snr_db(eng.mimi_decode(g["hello_scaled"]), g["hello_pcm"]))
for k in g.files:
    if "pcm" in k: print(k, g[k].shape)
eng.close()
