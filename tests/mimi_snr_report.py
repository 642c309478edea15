"""Development report (not a test): PCM SNR of the loaded libptts_b200.so against the oracle for a few
decode lengths, and against the reference golden PCM.  python tests/mimi_snr_report.py"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import snr_db  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402
sys.path.insert(0, str(ROOT / "oracle"))
import pyoracle  # noqa: E402

pkg = load_package()
md = pkg.synth.model_dir(0)
W = pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors")
eng = pkg.Engine(W, 0)
pyoracle.build(ref=False)
orc = pyoracle.Oracle(W)
g = np.load(ROOT / "tests" / "golden" / "reference_vectors.npz")
out = ["golden %.1f" % snr_db(eng.mimi_decode(g["hello_scaled"]), g["hello_pcm"])]
for frames in (2, 17, 40):
    lat = np.random.default_rng(frames).standard_normal((frames, 32)).astype(np.float32)
    out.append("%d frames %.1f" % (frames, snr_db(eng.mimi_decode(lat), orc.mimi_decode(lat))))
print("PCM SNR dB vs oracle:", ", ".join(out))
eng.close()
