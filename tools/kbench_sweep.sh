#!/bin/bash
# run every build/kbench_* variant over a few shapes (development tool)
for exe in "$@"; do
  echo "== $exe"
  for blk in 32 64 128; do $exe 65536 20 $blk 0.05; done
  $exe 1048576 5 64 0.05
  $exe 1048576 5 128 0.05
  $exe 65536 10 64 0
done
