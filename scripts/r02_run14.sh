#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_fasta_feeder.py tests/test_cli_host.py -x -q > gpurun_out/r02_pytest_d.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_d.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-bitparallel --no-cpu-baseline 2>gpurun_out/r02_b1.err > gpurun_out/r02_bench_n1d.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1d.json') if l.startswith('{')][-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e_stage_ms'], 'ascii', d['e2e']['ascii']['ms_per_step'], 'e2e_fasta', {k:d['e2e_fasta'].get(k) for k in ('total_s','open_parse_pack_s','scan_s')})
PY
tail -2 gpurun_out/r02_b1.err
