#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"k_verify|k_pair_small" -s 6 -c 2 -o gpurun_out/r02_verify_pair -f python bench.py --steps 2 --warmup 3 --genome-mbp 375 --no-cpu-baseline --no-e2e --no-bitparallel > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
