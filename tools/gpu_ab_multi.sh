# usage: bash tools/gpu_ab_multi.sh NAME...  - same-box comparison of builds build_ab/NAME.so: persistent step kernel (256 x 100 frames)
# and device-resident Mimi decode (144 x 100 frames), two rounds each
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
  for n in "$@"; do
    cp build_ab/$n.so $L
    echo -n "$n step: "; timeout 200 python tools/step_time.py 256 100 3 2>&1 | tail -1
    echo -n "$n mimi: "; timeout 200 python /tmp/mimi_ab.py 2>&1 | tail -1
  done
done
cp /tmp/keep.so $L
