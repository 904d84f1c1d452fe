#!/bin/bash
# The host replays of the device cores (tests/emul/*.cpp compile the product's __host__ __device__ headers) under
# AddressSanitizer + UBSan, fed from exact-size heap buffers: full, truncated and damaged streams through the serial
# state machine and the rounds (32 / 128 lanes); chunks with and without a dictionary through the deflate phases.
# CPU only.  Last run (end of round 1): 738 + 540 runs, no report.
set -e
cd "$(dirname "$0")/../.."
for n in inf_emul def_emul; do
  g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -shared -std=c++17 -Wno-unknown-pragmas \
      -o /tmp/lib${n}_asan.so tests/emul/$n.cpp
done
export LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0
python tools/asan/inflate_replay.py
python tools/asan/deflate_replay.py
