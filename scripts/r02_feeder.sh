#!/bin/bash
python - <<'PY'
import numpy as np
rng=np.random.default_rng(7)
with open('/dev/shm/f.fa','wb') as f:
    for c in range(24):
        n=125_000_000
        seq=np.frombuffer(b"ACGT",dtype=np.uint8)[rng.integers(0,4,n)]
        f.write(b">chr%d\n"%(c+1))
        body=np.empty((n//80,81),dtype=np.uint8); body[:,:80]=seq.reshape(-1,80); body[:,80]=10
        f.write(body.tobytes())
PY
ls -la /dev/shm/f.fa
python - <<'PY'
import os, time, sys
sys.path.insert(0,'.')
os.environ['IPCR_FEEDER_TIMING']='1'
from ipcr_b200.fasta import FastaFile
from ipcr_b200 import _capi
L=_capi.lib(); err=_capi.errbuf(); L.ipcr_gpu_init(1,err,len(err))
for kw in (dict(pin=False,pack=True), dict(pin=False,pack=False), dict(pin=True,pack=False), dict(pin=False,pack=True,threads=8), dict(pin=False,pack=True,threads=32)):
    for rep in range(2):
        t0=time.perf_counter(); ff=FastaFile('/dev/shm/f.fa', **kw); t1=time.perf_counter(); ff.close(); t2=time.perf_counter()
        print(kw, 'open %.3f s close %.3f s'%(t1-t0, t2-t1), flush=True)
PY
rm -f /dev/shm/f.fa
