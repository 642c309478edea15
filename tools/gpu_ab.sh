# usage: bash tools/gpu_ab.sh [FLAGS_NEW]  - same-box A/B of the persistent step kernel: build_ab/base.so vs build_ab/new.so (+ optional PTTS_B200_MEGA_FLAGS for new)
L=pocket-tts.c_b200/libptts_b200.so
for rep in 1 2; do
  cp build_ab/base.so $L; echo -n "base: "; timeout 200 python tools/step_time.py 256 100 3 2>&1 | tail -1
  cp build_ab/new.so $L;  echo -n "new : "; timeout 200 python tools/step_time.py 256 100 3 2>&1 | tail -1
  if [ -n "$1" ]; then echo -n "new flags=$1: "; PTTS_B200_MEGA_FLAGS=$1 timeout 200 python tools/step_time.py 256 100 3 2>&1 | tail -1; fi
done
cp build_ab/new.so $L
