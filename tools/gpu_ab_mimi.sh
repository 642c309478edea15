# usage: bash tools/gpu_ab_mimi.sh  - same-box A/B of the Mimi decoder: build_ab/base.so vs build_ab/new.so (device-resident decode of 144 x 100 frames)
L=pocket-tts.c_b200/libptts_b200.so
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
  cp build_ab/base.so $L; echo -n "base: "; timeout 200 python /tmp/mimi_ab.py 2>&1 | tail -1
  cp build_ab/new.so $L;  echo -n "new : "; timeout 200 python /tmp/mimi_ab.py 2>&1 | tail -1
done
cp build_ab/new.so $L
