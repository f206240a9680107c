#!/bin/bash
mkdir -p gpurun_out
for extra in "" "--pairs 1"; do
  IPCR_CHAIN_DEBUG=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel $extra 2>gpurun_out/r02_b.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$extra', round(d['ms_per_step'],4), round(d['wall_ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items()}, d['candidates'], d['sites'], d['products'], d['gpu_launches'])
"
  tail -2 gpurun_out/r02_b.err
done
