#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_fullsize.py -x -q --durations=10 -k "full_text or headline" > gpurun_out/r02_pytest_full.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_pytest_full.log
