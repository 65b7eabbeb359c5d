#!/bin/bash
# 1 GPU: parity of the restructured merge (virtual ranks); then 2 GPUs would follow in a separate call
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_session.py -m gpu -q > gpurun_out/pytest_r02h.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r02h.log
