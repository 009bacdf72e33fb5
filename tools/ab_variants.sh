for v in "$@"; do echo "== $v"; SPL_LIB=/root/repo/splendor-ai_b200/csrc/variants/$v.so python tools/perf_playout.py --games 65536 1048576 --players 2 --reps 10; done
