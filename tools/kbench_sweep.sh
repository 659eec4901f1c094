#!/bin/bash
# run every given kbench variant over a few shapes (development tool)
for exe in "$@"; do
  echo "== $exe"
  $exe 65536 20 64 0.05
  $exe 1048576 5 64 0.05
  $exe 1048576 5 128 0.05
  $exe 65536 10 64 0
done
