#!/bin/bash
# round 2 profiles, second batch: the pigeonhole scan kernel at mm 3 and the final pairing kernel
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"k_seed_scan" -s 3 -c 1 -o gpurun_out/r02_seed_scan_mm3_full -f python bench.py --mm 3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant > gpurun_out/ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_pair_small" -s 4 -c 1 -o gpurun_out/r02_pair_small_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/r02_seed_scan_mm3_full.ncu-rep gpurun_out/r02_pair_small_final.ncu-rep
