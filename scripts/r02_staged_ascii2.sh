#!/bin/bash
mkdir -p gpurun_out
run() {
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 4 --no-cpu-baseline --no-bitparallel --no-n-variant --no-e2e-fasta --genome-mbp 500 --contigs 10 --pairs 1 --mm 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        e=json.loads(l)['e2e']; print('$*', '500 Mbp: packed', round(e['ms_per_step'],2), 'ascii pinned', round(e['ascii']['ms_per_step'],2), 'pageable', round(e['ascii'].get('pageable_ms_per_step'),2), e['ascii'].get('pageable_stage_ms'))
"
}
run IPCR_NOOP=1
run IPCR_STREAM_SLABS=4
run IPCR_STAGED_H2D_MIN=999999999999
