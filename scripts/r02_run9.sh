#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli_host.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02_pytest_cli.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_cli.log
bash scripts/r02_dense.sh 2>&1 | grep -v "^source_file\|^/dev/shm/dense.fa"
