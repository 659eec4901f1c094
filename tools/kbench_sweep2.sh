#!/bin/bash
for exe in build/kbench_v3 build/kbench_pi; do echo "== $exe"; $exe 65536 20 64 0.05; $exe 1048576 5 128 0.05; $exe 65536 10 64 0; done
echo "== nosmem"; for b in 32 64 128; do build/kbench_nosmem 65536 20 $b 0.05; done; build/kbench_nosmem 1048576 5 128 0.05
