#!/bin/bash
# usage: tools/ab_variants.sh [-g "65536 1048576"] [-p "2 4"] variant...   -- perf_playout for each csrc/variants/<variant>.so (same GPU call)
games="65536 1048576"; players="2"
while [ "$1" = "-g" ] || [ "$1" = "-p" ]; do
  if [ "$1" = "-g" ]; then games="$2"; else players="$2"; fi; shift 2
done
for v in "$@"; do echo "== $v"; SPL_LIB=/root/repo/splendor-ai_b200/csrc/variants/$v.so python tools/perf_playout.py --games $games --players $players --reps 10; done
