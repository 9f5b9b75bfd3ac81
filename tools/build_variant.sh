#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFLAG ..."   -> b-shot-slam_b200/libbshot_b200_NAME.so
# Tuning builds of the product library with extra nvcc flags; select one with BSHOT_LIB=<path>.
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/b-shot-slam_b200/csrc
OBJ=$ROOT/build/$NAME
mkdir -p "$OBJ"
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in "$SRC"/*.cu; do
  b=$(basename "$f" .cu)
  /usr/local/cuda/bin/nvcc $FLAGS -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -c "$f" -o "$OBJ/$b.o" &
done
wait
/usr/local/cuda/bin/nvcc $ARCH -shared -o "$ROOT/b-shot-slam_b200/libbshot_b200_$NAME.so" "$OBJ"/*.o
echo "built libbshot_b200_$NAME.so"
