#!/bin/bash
# usage: tools/ab_variants.sh [-g "65536 1048576"] variant...   -- perf_playout for each csrc/variants/<variant>.so (same GPU call)
games="65536 1048576"
if [ "$1" = "-g" ]; then games="$2"; shift 2; fi
for v in "$@"; do echo "== $v"; SPL_LIB=/root/repo/splendor-ai_b200/csrc/variants/$v.so python tools/perf_playout.py --games $games --players 2 --reps 10; done
