#!/bin/bash
for exe in build/kbench_nopair build/kbench_pair; do echo "== $exe"; for n in 4096 16384 65536 131072 262144 1048576; do $exe $n 10 64 0.05; done; $exe 65536 10 128 0.05; $exe 65536 10 64 0; done
