#!/bin/bash
# Build and run the thread-per-lane emulation tests of the HashAgg kernels: tools/emu/run.sh [name filter]
set -e
HERE=$(cd "$(dirname "$0")" && pwd); OUT=${TMPDIR:-/tmp}/b200q_emu
python "$HERE/build_emu.py" "$OUT" > /dev/null
g++ -std=c++20 -O1 -g -pthread -w -I"$HERE/include" -I"$OUT/blaze_b200/csrc" "$HERE/hashagg_emu_test.cc" -o "$OUT/hashagg_emu_test"
"$OUT/hashagg_emu_test" "$@"
