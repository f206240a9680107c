#!/bin/bash
# round 2: strided first level -- what bounds it?  L2 prefetch of a warp's next chunk on/off, + one full ncu capture
mkdir -p gpurun_out
out=gpurun_out/r02_seed_stride2.txt
: > $out
FLAGS="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-e2e-fasta --no-n-variant --no-bitparallel"
run() {
  env IPCR_SEED_DEBUG=1 "$@" timeout 300 python bench.py $FLAGS $EXTRA > gpurun_out/stride.json 2>gpurun_out/stride.err
  python - "$* $EXTRA" >> $out <<'PY'
import sys,json
ix=[l.strip() for l in open('gpurun_out/stride.err') if l.startswith('seed index')]
for l in open('gpurun_out/stride.json'):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], '|', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()}, d['candidates'], d['sites'], d['products'], '|', ix[0] if ix else '')
PY
}
EXTRA="--pairs 1"
for T in 1 2 4; do for PF in 0 1; do run IPCR_SEED_STRIDE_FORCE=$T IPCR_SEED_PREFETCH=$PF; done; done
EXTRA="--pairs 10"
for T in 2 4; do for PF in 0 1; do run IPCR_SEED_STRIDE_FORCE=$T IPCR_SEED_PREFETCH=$PF; done; done
EXTRA=""
run IPCR_SEED_PREFETCH=0
run IPCR_SEED_PREFETCH=1
cat $out
IPCR_SEED_PREFETCH=0 ncu --set full --import-source on --clock-control none -k regex:"k_seed_scan" -s 3 -c 1 -o gpurun_out/r02_seed_scan_t4 -f python bench.py $FLAGS --steps 2 --pairs 1 > gpurun_out/ncu_t4.log 2>&1
ls -la gpurun_out/r02_seed_scan_t4.ncu-rep
