"""What bounds the short-K SEANet convolutions? One-CTA timelines (PTTS_B200_GEMM_TRACE=1) of the stage-2 shape with
one thing changed at a time: taps vs plain K, N tile, number of rows. test_gemm writes fp32 outputs (no activation)."""
import os, sys
from pathlib import Path
import numpy as np
os.environ["PTTS_B200_GEMM_TRACE"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package
pkg = load_package()
md = pkg.synth.model_dir(0)
eng = pkg.Engine(pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors"), 0)
rng = np.random.default_rng(0)
def run(name, rows, N, K, bn, taps=1, shift0=0, step=0):
    A = rng.standard_normal((rows, K // taps)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    print("==", name, "tiles per CTA %.1f" % (rows / 128 * ((N + bn - 1) // bn) / 148), file=sys.stderr)
    eng.test_gemm(A, B, 1, rows, N, K, taps=taps, shift0=shift0, step=step, bf16=False, split=0, transposed=False, bn=bn, split_k=1)
R = 128 * 148 * 40
run("k3 conv 64->64 (taps=3)", R, 64, 192, 64, taps=3, shift0=-2, step=1)
run("plain K=192 N=64", R, 64, 192, 64)
run("plain K=64 N=64", R, 64, 64, 64)
run("plain K=64 N=128", R, 128, 64, 128)
run("plain K=256 N=256 (bn 256)", R // 4, 256, 256, 256)
run("plain K=512 N=64", R, 64, 512, 64)
