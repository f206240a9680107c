#!/bin/bash
# round 2: pageable ASCII records (what a Go []byte is) through the staging ring: parity, then e2e with / without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -x -k "streamed or packed or multi_device or worker" > gpurun_out/r02_staged_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r02_staged_tests.log
for mode in staged driver; do
  if [ $mode = driver ]; then export IPCR_STAGED_H2D_MIN=999999999999; else unset IPCR_STAGED_H2D_MIN; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bitparallel --no-n-variant --no-e2e-fasta > gpurun_out/r02_staged_$mode.json 2> gpurun_out/r02_staged_$mode.err
  echo "bench $mode rc=$?"
  python - $mode <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r02_staged_{sys.argv[1]}.json') if l.startswith('{')][-1])
e=d['e2e']; print(sys.argv[1], 'e2e packed ms', round(e['ms_per_step'],2), 'ascii pinned ms', round(e['ascii']['ms_per_step'],2), 'ascii pageable ms', e['ascii'].get('pageable_ms_per_step'))
PY
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bitparallel --no-n-variant --no-e2e-fasta --genome-mbp 500 --contigs 10 --pairs 1 --mm 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        e=json.loads(l)['e2e']; print('500 Mbp: packed', round(e['ms_per_step'],2), 'ascii pinned', round(e['ascii']['ms_per_step'],2), 'pageable', e['ascii'].get('pageable_ms_per_step'))
"
