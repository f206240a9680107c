#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bd.err ) 2>&1 | tail -4
tail -3 gpurun_out/r02_bd.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_default_n1.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','stage_ms','e2e','e2e_fasta','with_scattered_n','cpu_baseline','cpu_baseline_default_flags','reference_sample_mbp','roofline','clocks','gpu_launches'):
    print(k, json.dumps(d.get(k))[:900])
print('bitparallel', json.dumps(d.get('bitparallel'))[:600])
PY
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bd.err ) 2>&1 | tail -4
cut -c1-1200 gpurun_out/r02_bench_ref.json
