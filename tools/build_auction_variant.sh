#!/bin/bash
# A/B builds of the exact-assignment kernel only: recompiles csrc/auction.cu with extra -D flags and links it with the
# product build's other objects into tools/variants/<name>.so (diagnostics; run with SHWD_B200_LIB=tools/variants/<name>.so).
#   tools/build_auction_variant.sh f01 -DSHWD_AU_EPS_FACTOR=0.1
set -e
name=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG="$ROOT/sphere-homeomorphic-wasserstein-distance-for-point-cloud-registration_b200"
OUT="$ROOT/tools/variants"; mkdir -p "$OUT"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 "$@" -c "$PKG/csrc/auction.cu" -o "$OUT/auction_$name.o"
objs=$(ls "$PKG"/build/*.o | grep -v "/auction.o")
nvcc -shared --cudart=static -o "$OUT/$name.so" $objs "$OUT/auction_$name.o" 2>/dev/null
rm -f "$OUT/auction_$name.o"
echo "$OUT/$name.so"
