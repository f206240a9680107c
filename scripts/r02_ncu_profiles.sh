#!/bin/bash
# round 2 profiles: launch list of the bench step (asynchronous chain) + full capture of the dominant kernel
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant > gpurun_out/r02_plain.json 2>gpurun_out/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_ncu_launches_chain.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant > gpurun_out/ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_seed_scan" -s 3 -c 1 -o gpurun_out/r02_seed_scan_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/r02_seed_scan_full.ncu-rep
grep -v "^==" gpurun_out/r02_ncu_launches_chain.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['ID'], r['Kernel Name'][:44], r['Block Size'], r['Grid Size'], r['Metric Value'])
" | tail -8
