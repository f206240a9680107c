#!/bin/bash
# round 2: mm 3 x 400 orientations -- batched second-level lookups, then the pigeonhole first level
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "mm3" > gpurun_out/r02_pytest_mm3.log 2>&1; echo "pytest mm3 rc=$?"; tail -3 gpurun_out/r02_pytest_mm3.log
IPCR_SEED_PIGEON=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q > gpurun_out/r02_pytest_pigeon.log 2>&1; echo "pytest forced-pigeonhole rc=$?"; tail -3 gpurun_out/r02_pytest_pigeon.log
out=gpurun_out/r02_mm3.txt
: > $out
run() {
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant $EXTRA 2>>gpurun_out/r02_mm3.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$* $EXTRA', 'step', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items()}, d['candidates'], d['sites'], 'products', d['products'])
" >> $out
}
EXTRA="--mm 3" run IPCR_NOOP=1
EXTRA="--mm 3" run IPCR_SEED_PIGEON=0
EXTRA="--mm 3 --n-per-mbp 10" run IPCR_NOOP=1
EXTRA="--mm 2" run IPCR_SEED_PIGEON=1
EXTRA="--mm 2" run IPCR_NOOP=1
cat $out
