#!/usr/bin/env bash
# Builds the UNMODIFIED reference extensions (diff-gaussian-rasterization, simple-knn) for sm_100 into
# baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun) -- the reference arm of bench.py and the
# drop-in comparison tests import them from there.  Sources are compiled from a /tmp copy because
# /root/reference is read-only; nothing of the reference enters the repository's history.
#   usage: tools/install_reference.sh [/root/reference]
set -euo pipefail
REF="${1:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/baseline/_ref"
TMP="$(mktemp -d /tmp/dgs_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF/submodules/diff-gaussian-rasterization" "$TMP/DGR"
cp -r "$REF/submodules/simple-knn" "$TMP/KNN"
# the reference omits <cstdint> / <cfloat> (rasterizer_impl.h:24, simple_knn.cu:90): force-include them
export NVCC_APPEND_FLAGS="-include cstdint -include cfloat" TORCH_CUDA_ARCH_LIST=10.0 MAX_JOBS="${MAX_JOBS:-8}"
(cd "$TMP/DGR" && python setup.py build_ext --inplace >"$TMP/dgr.log" 2>&1) || { tail -30 "$TMP/dgr.log"; exit 1; }
(cd "$TMP/KNN" && python setup.py build_ext --inplace >"$TMP/knn.log" 2>&1) || { tail -30 "$TMP/knn.log"; exit 1; }
mkdir -p "$OUT"
rm -rf "$OUT/diff_gaussian_rasterization" "$OUT/simple_knn"
cp -r "$TMP/DGR/diff_gaussian_rasterization" "$OUT/diff_gaussian_rasterization"
cp -r "$TMP/KNN/simple_knn" "$OUT/simple_knn"
chmod -R u+w "$OUT"
find "$OUT" -name __pycache__ -prune -exec rm -rf {} +
ls -la "$OUT/diff_gaussian_rasterization" "$OUT/simple_knn"
