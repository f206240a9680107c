#!/bin/bash
# round 2: GPU tests + quick bench of the asynchronous chain
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest3.log
for extra in "" "--n-per-mbp 10" "--pairs 1" "--genome-mbp 4.6 --contigs 1 --pairs 1"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel $extra 2>>gpurun_out/r02_b.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$extra', round(d['ms_per_step'],4), round(d['wall_ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items()}, d['candidates'], d['sites'], d['products'], d['gpu_launches'])
"
  IPCR_NO_CHAIN=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel $extra 2>>gpurun_out/r02_b.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('NOCHAIN $extra', round(d['ms_per_step'],4), round(d['wall_ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items()}, d['candidates'], d['sites'], d['products'], d['gpu_launches'])
"
done
tail -5 gpurun_out/r02_b.err
