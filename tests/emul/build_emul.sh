#!/bin/sh
# Build the HOST EMULATION of the kernel sources (test infrastructure only):
# the same .cu files, compiled by g++ with -DAEW_EMUL (see csrc/aew_emul.h).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/../../ae_waveforms_b200/csrc
OUT=$HERE/libaewave_emul.so
FLAGS="-O1 -g -std=c++17 -fPIC -DAEW_EMUL -ffp-contract=off -fno-fast-math -pthread -Wall -Wno-unused-function -Wno-unknown-pragmas -Wno-unused-variable"
OBJS=""
for f in "$SRC"/*.cu; do
  o=$HERE/$(basename "$f" .cu).emul.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find "$SRC" -name '*.h' -newer "$o" -o -name '*.cuh' -newer "$o")" ]; then
    g++ $FLAGS -x c++ -c "$f" -o "$o"
  fi
  OBJS="$OBJS $o"
done
for f in "$SRC"/*.cpp; do
  [ -f "$f" ] || continue
  o=$HERE/$(basename "$f" .cpp).emulcpp.o
  g++ $FLAGS -c "$f" -o "$o"
  OBJS="$OBJS $o"
done
g++ -shared -pthread -o "$OUT" $OBJS
echo "built $OUT"
