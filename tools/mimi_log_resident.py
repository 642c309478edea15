"""Per-launch GEMM timing of one DEVICE-RESIDENT Mimi decode (one chunk): python tools/mimi_log_resident.py [n] [frames]"""
import os, sys
from pathlib import Path
import numpy as np
os.environ["PTTS_B200_GEMM_LOG"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 144
F = int(sys.argv[2]) if len(sys.argv) > 2 else 100
md = pkg.synth.model_dir(0)
eng = pkg.Engine(pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors"), 0)
lat = np.random.default_rng(0).standard_normal((n, F, 32)).astype(np.float32)
eng.mimi_stage(lat); eng.mimi_run(); eng.sync()
eng.profile(True)
eng.mimi_run(); eng.sync()
print({k: v for k, v in eng.profile_read().items() if v["launches"]})
eng.profile(False)
eng.close()
