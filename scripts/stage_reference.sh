#!/bin/bash
# Stages the UNMODIFIED reference (dynaroars/neuralsat, neuralsat-pt201) into baseline/_ref/ so that it travels to the
# GPU box with the snapshot (baseline/_ref is git-ignored, NOT gpurun-ignored).  Used only by the reference arm of
# bench.py (`--impl reference`, oracle/ref_arm.py) and by nothing in the product path.
# The reference is a source tree without setup.py / pyproject.toml, so `pip install --target` does not apply; the
# Python packages are copied as they are (everything except the 107 MB of example networks).
set -e
SRC=${1:-/root/reference/neuralsat-pt201}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref/neuralsat-pt201"
[ -d "$SRC/auto_LiRPA" ] || { echo "no reference tree at $SRC"; exit 0; }
mkdir -p "$DST"
for d in auto_LiRPA abstractor heuristic util solver verifier attacker onnx2pytorch; do
  rm -rf "$DST/$d"
  cp -r "$SRC/$d" "$DST/$d"
done
cp "$SRC"/*.py "$DST"/
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "staged $(du -sh "$DST" | cut -f1) at $DST"
