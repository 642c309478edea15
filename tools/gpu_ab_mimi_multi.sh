# usage: bash tools/gpu_ab_mimi_multi.sh NAME...  - same-box comparison of Mimi decoder builds build_ab/NAME.so (device-resident decode of
# 144 x 100 frames, two rounds), then the per-launch log (36 x 100 frames) and the PCM SNR against the oracle for the LAST name
L=pocket-tts.c_b200/libptts_b200.so
cp $L /tmp/keep.so
cat > /tmp/mimi_ab.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
pkg = load_package()
md = pkg.synth.model_dir(0)
eng = pkg.Engine(pkg.synth.read_safetensors(md / "tts_b6369a24.safetensors"), 0)
lat = np.random.default_rng(0).standard_normal((144, 100, 32)).astype(np.float32)
eng.mimi_stage(lat); eng.mimi_run(); eng.sync()
ms = []
for _ in range(5):
    eng.timer_start(); eng.mimi_run(); ms.append(eng.timer_stop())
print("mimi_run 144 x 100 frames, device resident: ms", ["%.2f" % m for m in ms])
PY
for rep in 1 2; do
  for n in "$@"; do cp build_ab/$n.so $L; echo -n "$n: "; timeout 200 python /tmp/mimi_ab.py 2>&1 | tail -1; done
done
for n in "$@"; do :; done
cp build_ab/$n.so $L; echo "== $n"; timeout 300 python tests/mimi_snr_report.py 2>&1 | tail -1
timeout 200 python tools/mimi_log.py 36 100 > gpurun_out/mimi_log_$n.txt 2>&1; grep "^gemm" gpurun_out/mimi_log_$n.txt | tail -7
cp /tmp/keep.so $L
