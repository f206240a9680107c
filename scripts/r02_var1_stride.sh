#!/bin/bash
# round 2: shuffle look-ahead (IPCR_SEED_VAR=1) in the strided scan kernels, where the streaming loads are a quarter of the LSU pipe
mkdir -p gpurun_out
out=gpurun_out/r02_var1_stride.txt
: > $out
FLAGS="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-e2e-fasta --no-n-variant --no-bitparallel"
run() {
  env "$@" timeout 200 python bench.py $FLAGS $EXTRA 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$* $EXTRA |', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()}, d['candidates'], d['sites'], d['products'])
" >> $out
}
EXTRA="--pairs 1"; run IPCR_SEED_VAR=0; run IPCR_SEED_VAR=1
EXTRA="--pairs 10"; run IPCR_SEED_VAR=0; run IPCR_SEED_VAR=1
cat $out
IPCR_SEED_VAR=1 timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "known_answers or equivalence or randomized_parity_small or hit_cap or streamed" 2>&1 | tail -2
