#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q > gpurun_out/r02_pytest_c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_c.log
for extra in "" "--genome-mbp 375" "--genome-mbp 4.6 --contigs 1 --pairs 1"; do
  IPCR_CHAIN_DEBUG=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant $extra 2>gpurun_out/r02_b.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$extra', round(d['ms_per_step'],4), round(d['wall_ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items()}, d['sites'], d['products'], d['gpu_launches'])
"
  tail -2 gpurun_out/r02_b.err | cut -c1-300
done
