#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02_pytest_gpu_all.log 2>&1 ) 2>&1 | tail -3; echo "rc=$?"
tail -16 gpurun_out/r02_pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
