#!/bin/bash
# round 2 final: full GPU suite, smoke, default bench (as the driver runs it), reference arm, ncu launch list
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_final.log 2>&1 ) 2>&1 | tail -3; tail -3 gpurun_out/r02_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bf.err ) 2>&1 | tail -3
tail -2 gpurun_out/r02_bf.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref_final.json 2>> gpurun_out/r02_bf.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_ncu_launches_chain_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bitparallel --no-e2e-fasta --no-n-variant > gpurun_out/ncu.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1_final.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','stage_ms','gpu_launches','clocks'):
    print(k, json.dumps(d.get(k))[:300])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'ascii', d['e2e']['ascii']['ms_per_step'])
print('e2e_fasta', {k:d['e2e_fasta'].get(k) for k in ('total_s','open_parse_pack_s','scan_s','sort_format_s')})
print('n', d['with_scattered_n']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline_default_flags']['value'], 'roofline', d['roofline']['frac'], 'bitpar', d['bitparallel']['ms_per_step'], d['bitparallel']['roofline']['frac'])
r=json.loads([l for l in open('gpurun_out/r02_bench_ref_final.json') if l.startswith('{')][-1])
print('ref', r['value'], r['ms_per_step'], r['reference_sample_mbp'])
PY
