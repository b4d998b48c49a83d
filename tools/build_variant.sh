#!/bin/bash
# Build a variant of libshwd_b200.so with extra -D flags into gpurun-visible tools/variants/<name>.so (diagnostics only).
#   tools/build_variant.sh name -DSHWD_OFFSET_LSE=0 ...
set -e
name=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG="$ROOT/sphere-homeomorphic-wasserstein-distance-for-point-cloud-registration_b200"
OUT="$ROOT/tools/variants"; mkdir -p "$OUT/obj_$name"
pids=()
for f in "$PKG"/csrc/*.cu; do
  b=$(basename "$f" .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 "$@" -c "$f" -o "$OUT/obj_$name/$b.o" 2> >(grep -i "error" -A3 >&2) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared --cudart=static -o "$OUT/$name.so" "$OUT"/obj_$name/*.o
rm -rf "$OUT/obj_$name"
echo "$OUT/$name.so"
