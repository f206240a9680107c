#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "mm3" > gpurun_out/r02_pytest_mm3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_mm3.log
IPCR_SEED_BLOOM=0 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "randomized or known or n_runs or dirty or hit_cap" > gpurun_out/r02_pytest_nobloom.log 2>&1; echo "pytest nobloom rc=$?"; tail -3 gpurun_out/r02_pytest_nobloom.log
for extra in "--mm 3" "--mm 3 --n-per-mbp 10"; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant $extra 2>>gpurun_out/r02_mm3.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$extra', 'step', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items()}, d['candidates'], d['sites'], 'products', d['products'])
"
done
