#!/bin/bash
for exe in "$@"; do echo "== $exe"; for n in 4096 65536 1048576; do $exe $n 10 64 0.05; done; $exe 1048576 5 128 0.05; $exe 65536 10 64 0; done
