#!/bin/bash
# the packed ipcr_gpu_scan_fasta call spread over the devices of ONE process: parity test, then the timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -m gpu -k "multi_device" > gpurun_out/r02_packed_multi_test.log 2>&1
echo "pytest rc=$?"
tail -3 gpurun_out/r02_packed_multi_test.log
n=$(nvidia-smi -L | wc -l)
for k in 2 $n; do
  [ "$k" -gt "$n" ] && continue
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e-fasta --no-n-variant --in-process-devices $k \
      > gpurun_out/r02_packed_multi_k$k.json 2> gpurun_out/r02_packed_multi_k$k.err
  echo "bench k=$k rc=$?"
  python - <<PY
import json
l=[x for x in open("gpurun_out/r02_packed_multi_k$k.json") if x.startswith("{")]
d=json.loads(l[-1]); e=d["e2e"]
print("one device e2e ms", e["ms_per_step"], "one_process", e.get("one_process"))
PY
done
