"""Timeline of one CTA of the tcgen05 GEMM for the FlowLM / Mimi shapes (PTTS_B200_GEMM_TRACE=1)."""
import os, sys, time
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
def run(name, **kw):
    rows, N, K = kw["rows"], kw["N"], kw["K"]
    b = kw.get("batches", 1)
    A = rng.standard_normal((b * rows, K // kw.get("taps", 1))).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    print("==", name, file=sys.stderr)
    for _ in range(2):
        t0 = time.perf_counter()
        eng.test_gemm(A, B, b, rows, N, K, taps=kw.get("taps", 1), shift0=kw.get("shift0", 0), step=kw.get("step", 0),
                      bf16=kw.get("bf16", True), split=kw.get("split", 0), transposed=kw.get("transposed", False), bn=kw["bn"],
                      split_k=kw.get("split_k", 1))
run("flow qkv 3072x1024, B=256", rows=3072, N=256, K=1024, bn=64, split=2, transposed=True)
run("flow ffn2 1024x4096 split4", rows=1024, N=256, K=4096, bn=64, split=2, transposed=True, split_k=4)
run("flow mlp 512x512", rows=512, N=256, K=512, bn=64, split=2, transposed=True)
run("mimi linear T=16384 512->1536", rows=16384, N=1536, K=512, bn=256, bf16=False)
run("mimi dec_in T=16384 k7", rows=16384, N=512, K=3584, taps=7, shift0=-6, step=1, bn=256, bf16=False)
run("mimi res conv T=122880 64->64 k3", rows=122880, N=64, K=192, taps=3, shift0=-2, step=1, bn=64, bf16=False)
