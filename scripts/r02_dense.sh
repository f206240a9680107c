#!/bin/bash
# round 2, SURVEY 8f row 4: does a dense-hit panel make the HOST collector / --sort / TSV writer the bottleneck?
mkdir -p gpurun_out
python - <<'PY'
import numpy as np
rng=np.random.default_rng(7)
with open('/dev/shm/dense.fa','wb') as f:
    for c in range(10):
        n=50_000_000
        seq=np.frombuffer(b"ACGT",dtype=np.uint8)[rng.integers(0,4,n)]
        f.write(b">chr%d\n"%(c+1))
        body=np.empty((n//80,81),dtype=np.uint8); body[:,:80]=seq.reshape(-1,80); body[:,80]=10
        f.write(body.tobytes())
PY
ls -la /dev/shm/dense.fa
for pr in "ACGTTGCAAGCT TTGACCGTAAGC" "ACGTTGCAAGC TTGACCGTAAG"; do
  set -- $pr
  ( time IPCR_CLI_TIMING=1 ./ipcr_b200/ipcr_gpu_cli ipcr --forward $1 --reverse $2 --mismatches 3 --terminal-window 0 --hit-cap 0 --max-length 2000 --sort /dev/shm/dense.fa > /dev/shm/dense.tsv ) 2>&1 | tail -5
  wc -l /dev/shm/dense.tsv; head -3 /dev/shm/dense.tsv
done
rm -f /dev/shm/dense.fa /dev/shm/dense.tsv
