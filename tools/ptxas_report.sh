#!/bin/bash
# Regenerate profiles/ptxas_registers.txt: one line per kernel (registers, spills, static shared memory) from
# `nvcc -Xptxas=-v` of csrc/kernels.cu with the product's flags, plus the bulk-copy SASS check and the id of the in-tree
# library it describes.  Run from anywhere; cross-compiles without a GPU.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC="$ROOT/splendor-ai_b200/csrc"
OUT="$ROOT/profiles/ptxas_registers.txt"
TMP=$(mktemp -d)
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
nvcc $FLAGS -Xptxas=-v -o "$TMP/lib.so" "$CSRC/kernels.cu" 2> "$TMP/ptxas.log"
python - "$TMP/ptxas.log" > "$OUT" <<'EOF'
import re, subprocess, sys
name = None
spill = (0, 0)
for line in open(sys.argv[1]):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*\)$", "", name)
        continue
    m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        spill = (int(m.group(1)), int(m.group(2)))
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        sm = re.search(r"(\d+) bytes smem", line)
        print(f"{name:<62} regs {int(m.group(1)):3d}  spill st/ld {spill[0]:3d}/{spill[1]:3d}  static smem {int(sm.group(1)) if sm else 0:6d}")
        name = None
EOF
{
  echo
  echo "# built with: nvcc $FLAGS -Xptxas=-v ($(nvcc --version | grep -o 'V[0-9][0-9.]*'))"
  echo "# SASS check of the same build:  cuobjdump -sass | grep -c UBLKCP  ->  $(cuobjdump -sass "$TMP/lib.so" | grep -c UBLKCP) (the mask kernels' bulk-copy stores, one per player count)"
  echo "# source: git $(git -C "$ROOT" rev-parse --short HEAD)$(git -C "$ROOT" diff --quiet -- splendor-ai_b200/csrc || echo '+local changes'), sha256 of kernels.cu $(sha256sum "$CSRC/kernels.cu" | cut -c1-16), engine.cuh $(sha256sum "$CSRC/engine.cuh" | cut -c1-16)"
} >> "$OUT"
rm -rf "$TMP"
tail -4 "$OUT"
