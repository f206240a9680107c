#!/bin/bash
# round 2: full ncu page of the small-panel scan kernel as it runs by default (k_seed_scan<2, 0, 4>, 3 Gbp x 1 pair)
mkdir -p gpurun_out
FLAGS="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-e2e-fasta --no-n-variant --no-bitparallel --pairs 1"
python bench.py $FLAGS > gpurun_out/r02_pairs1.json 2>/dev/null || exit 1
ncu --set full --import-source on --clock-control none -k regex:"k_seed_scan" -s 3 -c 1 -o gpurun_out/r02_seed_scan_t4_2bit -f python bench.py $FLAGS > gpurun_out/ncu_t4b.log 2>&1
ls -la gpurun_out/r02_seed_scan_t4_2bit.ncu-rep
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_pairs1.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['stage_ms'], d['roofline'])"
