#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's native library from its own,
# UNMODIFIED sources where they lie under /root/reference into
# oracle/_ref/libTrajectoryConstraints.so (git-ignored, travels with gpurun).
# Eigen3 and GoogleTest are not installed in this image, so the sources are
# compiled against the header stand-ins in oracle/ref_shim/ (an Eigen subset +
# a FRIEND_TEST macro).  No reference source is copied into the repo.
#
# Reference build being replaced: CC/CMakeLists.txt:1-18 and
# CC/src/CMakeLists.txt:1-14 (CC = trajectory_generation/constraint_functions/
# TrajectoryConstraintsCCode), one SHARED library, C++11.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${TG_REFERENCE_ROOT:-/root/reference}"
CC_DIR="$REF/trajectory_generation/constraint_functions/TrajectoryConstraintsCCode"
OUT="$HERE/_ref"
if [ ! -d "$CC_DIR/src" ]; then
    echo "build_ref: reference sources not found at $CC_DIR (using prebuilt $OUT if present)" >&2
    exit 0
fi
mkdir -p "$OUT"
g++ -std=c++11 -O2 -fPIC -shared -w \
    -I "$HERE/ref_shim" -I "$CC_DIR/include" \
    "$CC_DIR"/src/*.cpp \
    -o "$OUT/libTrajectoryConstraints.so"
echo "build_ref: wrote $OUT/libTrajectoryConstraints.so"
