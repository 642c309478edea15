"""One Mimi decode for ncu: python tools/mimi_once.py [n] [frames]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100
md = pkg.synth.model_dir(0)
eng = pkg.Engine(pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors"), 0)
lat = np.random.default_rng(0).standard_normal((n, F, 32)).astype(np.float32)
out = np.zeros((n, F * 1920), np.float32)
eng.mimi_decode_batch(lat, out)
eng.close()
