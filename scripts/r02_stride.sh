#!/bin/bash
# round 2: anchor stride of the seed index's first level (IPCR_SEED_STRIDE = largest stride the cost model may pick,
# IPCR_SEED_STRIDE_FORCE = this stride whatever the model says) over panel sizes, 3 Gbp, mm 2 (run under gpurun)
mkdir -p gpurun_out
out=gpurun_out/r02_seed_stride.txt
: > $out
run() {
  env IPCR_SEED_DEBUG=1 "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-e2e-fasta --no-n-variant --no-bitparallel $EXTRA 2>gpurun_out/stride.err | python -c "
import sys,json,re
ix=[l.strip() for l in open('gpurun_out/stride.err') if l.startswith('seed index')]
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$* $EXTRA |', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()}, d['candidates'], d['sites'], d['products'], '|', ix[0] if ix else '')
" >> $out
}
if [ "$1" != "sweep-only" ]; then
  timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_stride_parity.log 2>&1
  echo "parity rc=$?"; tail -3 gpurun_out/r02_stride_parity.log
fi
for P in 1 10 25 50; do
  EXTRA="--pairs $P"
  run IPCR_SEED_STRIDE=1
  run IPCR_SEED_STRIDE_FORCE=2
  run IPCR_SEED_STRIDE_FORCE=4
  run IPCR_NOOP=1
done
EXTRA="" run IPCR_NOOP=1
EXTRA="--pairs 1 --mm 3" run IPCR_SEED_STRIDE=1
EXTRA="--pairs 1 --mm 3" run IPCR_NOOP=1
cat $out
