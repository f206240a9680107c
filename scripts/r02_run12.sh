#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_cli_host.py tests/test_host_apps.py tests/test_fasta_feeder.py tests/test_gpu_multi.py -x -q > gpurun_out/r02_pytest_b.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_b.log
IPCR_FEEDER_TIMING=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-bitparallel --no-cpu-baseline 2>gpurun_out/r02_b1.err > gpurun_out/r02_bench_n1c.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1c.json') if l.startswith('{')][-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'e2e_fasta', {k:d['e2e_fasta'].get(k) for k in ('total_s','open_parse_pack_s','scan_s','sort_format_s','scan_device_ms')})
PY
grep feeder gpurun_out/r02_b1.err | tail -3
