#!/bin/bash
# tools/build_variant.sh NAME -DCG_...   -> /root/repo/gpurun_tmp_NAME.so  (tuning builds; select with CG_B200_LIB)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
make -s -C "$ROOT/coulombgalore_b200/csrc" -j8 BUILD="$ROOT/coulombgalore_b200/csrc/build_$name" LIB="$ROOT/gpurun_tmp_$name.so" EXTRA_NVFLAGS="$*" 2>&1 | grep -E "error" || true
ls -la "$ROOT/gpurun_tmp_$name.so" | awk '{print $5, $9}'
rm -rf "$ROOT/coulombgalore_b200/csrc/build_$name"  # objects only; the snapshot sent to the GPU box is size-limited
