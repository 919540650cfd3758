#!/usr/bin/env bash
# Copies the few reference files the same-box comparison needs into the git-ignored baseline/_ref/ so that they travel
# to the GPU box with gpurun (the box has no /root/reference).  Nothing under baseline/_ref/ is ever committed or
# imported by the product path: it is the measured opponent (tools/bench_reference_gpu.py, bench.py --impl reference).
# Usage: tools/vendor_reference.sh [/root/reference]
set -euo pipefail
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
DST=$ROOT/baseline/_ref
[ -d "$REF/octic_vits" ] || { echo "no reference at $REF (this only works in the build container)"; exit 1; }
rm -rf "$DST"
mkdir -p "$DST/deit" "$DST/experiments"
cp -r "$REF/octic_vits" "$DST/octic_vits"
cp "$REF/deit/__init__.py" "$REF/deit/vit.py" "$DST/deit/"
cp "$REF/experiments/test_equivariance.py" "$REF/experiments/complexity.py" "$DST/experiments/"
cp "$REF/LICENSE" "$DST/LICENSE"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$DST" && find . -type f | sort | xargs sha256sum ) > "$DST/MANIFEST.sha256"
echo "vendored $(find "$DST" -name '*.py' | wc -l) reference files into $DST"
