#!/bin/bash
# Build the in-tree library and, only if that succeeded, run the given command on the GPU box:  tools/brun.sh '<command>'
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
python sphere-homeomorphic-wasserstein-distance-for-point-cloud-registration_b200/build.py > /tmp/brun_build.log 2>&1 || { grep -i "error" -A6 /tmp/brun_build.log | head -40; echo "BUILD FAILED"; exit 1; }
cuobjdump -res-usage sphere-homeomorphic-wasserstein-distance-for-point-cloud-registration_b200/build/sinkhorn.o 2>/dev/null | grep -A1 "sinkhorn_..._kernelILi[01]" | grep -o "Function [^:]*\|REG:[0-9]*\|STACK:[0-9]*" | paste - - - | sed 's/_ZN4shwd19//; s/EEEvNS_10SinkParamsE//'
/usr/local/graft/bin/gpurun --timeout ${BRUN_TIMEOUT:-600} -- "$1" 2>&1 | tail -${BRUN_TAIL:-12}
