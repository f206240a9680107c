#!/bin/bash
# round 2: the pack-only feeder on the GPU box's host: feeder + CLI tests, then FASTA on disk -> sorted TSV
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fasta_feeder.py tests/test_cli_host.py -q -x > gpurun_out/r02_feeder2_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r02_feeder2_tests.log
nproc; grep -m1 "model name" /proc/cpuinfo
IPCR_FEEDER_TIMING=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-bitparallel --no-n-variant > gpurun_out/r02_feeder2_bench.json 2> gpurun_out/r02_feeder2_bench.err
echo "bench rc=$?"
grep "feeder" gpurun_out/r02_feeder2_bench.err | tail -8
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_feeder2_bench.json') if l.startswith('{')][-1])
print('e2e_fasta', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['e2e_fasta'].items() if k!='api'})
print('e2e', d['e2e']['ms_per_step'])
PY
