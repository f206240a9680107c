#!/bin/bash
# round 2: pageable packed records through the multi-threaded pinned staging ring: tests, then FASTA -> TSV with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_fasta_feeder.py tests/test_cli_host.py -q -x -k "packed or feeder or cli or fasta" > gpurun_out/r02_feeder3_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r02_feeder3_tests.log
for mode in staged driver; do
  if [ $mode = driver ]; then export IPCR_STAGED_H2D_MIN=999999999999; else unset IPCR_STAGED_H2D_MIN; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bitparallel --no-n-variant --no-e2e > gpurun_out/r02_feeder3_$mode.json 2> gpurun_out/r02_feeder3_$mode.err
  echo "bench $mode rc=$?"
  python - $mode <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r02_feeder3_{sys.argv[1]}.json') if l.startswith('{')][-1])
print(sys.argv[1], 'e2e_fasta', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['e2e_fasta'].items() if k in ('total_s','open_parse_pack_s','scan_s','scan_device_ms','rows')})
PY
done
