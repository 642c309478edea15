"""Per-GEMM timing of one Mimi decode: python tools/mimi_log.py [n] [frames]   (PTTS_B200_GEMM_LOG=1)."""
import os, sys
from pathlib import Path
import numpy as np
os.environ["PTTS_B200_GEMM_LOG"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100
md = pkg.synth.model_dir(0)
W = pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors")
eng = pkg.Engine(W, 0)
lat = np.random.default_rng(0).standard_normal((n, F, 32)).astype(np.float32)
out = np.zeros((n, F * 1920), np.float32)
eng.mimi_decode_batch(lat, out)
eng.sync(); eng.timer_start(); eng.mimi_decode_batch(lat, out); print("mimi_decode_batch ms (incl. H2D/D2H)", eng.timer_stop())
eng.profile(True)
eng.mimi_decode_batch(lat, out)
print({k: v for k, v in eng.profile_read().items() if v["launches"]})
eng.profile(False)
eng.close()
