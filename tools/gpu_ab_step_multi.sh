L=pocket-tts.c_b200/libptts_b200.so
cp $L /tmp/keep.so
for rep in 1 2 3; do
  for n in "$@"; do cp build_ab/$n.so $L; echo -n "$n step: "; timeout 200 python tools/step_time.py 256 100 3 2>&1 | tail -1; done
done
cp /tmp/keep.so $L
