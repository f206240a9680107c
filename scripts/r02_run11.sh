#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_fasta_feeder.py -x -q -m "gpu or not gpu" > gpurun_out/r02_pytest_a.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_a.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-bitparallel 2>gpurun_out/r02_b1.err > gpurun_out/r02_bench_n1b.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1b.json') if l.startswith('{')][-1])
print('ms', d['ms_per_step'], d['stage_ms'], 'e2e', d['e2e']['ms_per_step'], 'e2e_fasta', {k:d['e2e_fasta'].get(k) for k in ('total_s','open_parse_pack_s','scan_s','sort_format_s')}, 'n-variant', d['with_scattered_n']['ms_per_step'])
PY
tail -3 gpurun_out/r02_b1.err
