#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"k_pair_small" -s 4 -c 1 -o gpurun_out/r02_pair_small_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/r02_pair_small_full.ncu-rep
