#!/bin/bash
# Install the UNMODIFIED reference (newton-raphson/sdf-representation) for the reference arm of bench.py and for the
# like-for-like GPU baseline (tools/bench_reference_gpu.py).  The reference is a script collection without packaging
# metadata (no setup.py / pyproject.toml), so `pip install --target baseline/_ref /root/reference` has nothing to
# build: this recipe copies its Python packages as they are.  baseline/_ref/ is git-ignored (never part of the
# history) but travels to the GPU box with gpurun.  Run from the repo root in the build container:
#     bash baseline/install_ref.sh
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -d "$SRC/model" ]; then
  echo "reference not found at $SRC (baseline/_ref left as it is)"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
for pkg in model executor configgen dataloader utils evaluations; do
  [ -d "$SRC/$pkg" ] && cp -r "$SRC/$pkg" "$DST/$pkg"
done
cp "$SRC/main.py" "$DST/main.py" 2>/dev/null || true
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && find model executor configgen dataloader utils evaluations main.py -type f -name '*.py' | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
echo "installed $(wc -l < "$DST/SHA256SUMS") unmodified reference files into $DST"
