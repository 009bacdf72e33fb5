#!/bin/bash
# Time the row kernels of one library build under explicit (warps per block)x(blocks per SM) shapes.
# usage: tools/sweep_row_shapes.sh LIB.so "obs shapes" "mask shapes" [envs]
ROOT=$(cd "$(dirname "$0")/.." && pwd)
LIB=$1; OBS=$2; MASK=$3; ENVS=${4:-262144}
export SPL_LIB=$LIB
for s in $OBS; do echo "== obs $s"; SPL_OBS_SHAPE=$s python $ROOT/tools/perf_kernels.py --envs $ENVS --only observe; done
for s in $MASK; do echo "== mask $s"; SPL_MASK_SHAPE=$s python $ROOT/tools/perf_kernels.py --envs $ENVS --only legal_mask; done
