#!/bin/bash
# round 2: the other BASELINE config shapes through the same bench.py (one random pair + self = 4 orientations)
mkdir -p gpurun_out
out=gpurun_out/r02_config_shapes.txt
: > $out
run() {
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bitparallel --no-e2e-fasta --no-n-variant "$@" 2>>gpurun_out/r02_cfg.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print('$*', '| device ms', round(d['ms_per_step'],4), '| Gbp/s', round(d['gbp_per_s'],1), '|', {k:round(v,3) for k,v in d['stage_ms'].items()}, '| launches/step', d['gpu_launches']//d['steps'], '| e2e packed ms', round(e['ms_per_step'],3), '| e2e ascii ms', round(e['ascii']['ms_per_step'],3), '| products', d['products'])
" >> $out
}
run --genome-mbp 4.6 --contigs 1 --pairs 1 --mm 2
run --genome-mbp 100 --contigs 50 --pairs 1 --mm 2
run --genome-mbp 500 --contigs 10 --pairs 1 --mm 1
run --mm 3
cat $out
