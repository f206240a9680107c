#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r02_pytest_multi2.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r02_pytest_multi2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/r02_b2.err > gpurun_out/r02_bench_strong_n2.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_strong_n2.json') if l.startswith('{')][-1])
print('n2 ms', d['ms_per_step'], d['stage_ms'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], d['exchange'][:30])
PY
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r02_b2.err | tail -3
