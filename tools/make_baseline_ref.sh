#!/bin/bash
# Copies the UNMODIFIED reference tree to baseline/_ref (git-ignored; it travels to the GPU box with the gpurun
# snapshot, /root/reference does not).  Used only by the reference arms: `bench.py --impl reference`,
# `python -m cim_quantization_b200.launcher` (which runs the reference's own main_lsq.py), tests/golden generators.
set -e
SRC=${1:-/root/reference}
DST="$(dirname "$0")/../baseline/_ref"
[ -d "$SRC" ] || { echo "no reference tree at $SRC"; exit 0; }
mkdir -p "$DST"
for d in examples models proto utils test; do rm -rf "$DST/$d"; cp -r "$SRC/$d" "$DST/$d"; done
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "reference copied to $DST"
