#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e') or {}
        print('$1', 'n', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'wall', round(d['wall_ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items()}, d['rank0_ms'], 'prod', d['products'], 'launches', d['gpu_launches'], 'e2e_ms', e.get('ms_per_step'), '|', (d.get('exchange') or '')[:60])
"; }
for n in "$@"; do
  for ex in peer nccl; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --exchange $ex 2>gpurun_out/r02_b${n}_$ex.err | tee gpurun_out/r02_bench_strong_n${n}_$ex.json | show $ex
    grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r02_b${n}_$ex.err | tail -4
  done
done
