#!/bin/bash
# Developer helper (GPU box): times the given variant libraries (variants/libblitzen_cull_<name>.so; "tree" = the in-tree build) on the bench workload.
#   usage: scripts/ab_sweep.sh <out-file> <cases> <name> [<name> ...]
OUT=$1; CASES=$2; shift 2
: > $OUT
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = "tree" ]; then LIB=blitzen_b200/libblitzen_cull.so; else LIB=variants/libblitzen_cull_$v.so; fi
  echo "== $v (rep $rep)" >> $OUT
  BLZ_CULL_LIB=$LIB timeout 300 python scripts/kernel_sweep.py --opts "stream_cfg=2" --cases $CASES --iters 30 >> $OUT 2>&1
done
done
