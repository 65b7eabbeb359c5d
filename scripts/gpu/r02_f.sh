#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_session.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/pytest_r02f.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r02f.log
timeout 300 python scripts/gpu/kbench.py cfg2 5 > gpurun_out/kbench_r02f.json 2>/dev/null; cat gpurun_out/kbench_r02f.json
DDN_LIB_PATH=$PWD/depthdensifier_b200/libddn_b200_nodefer.so timeout 300 python scripts/gpu/kbench.py cfg2 5 > gpurun_out/kbench_r02f_nodefer.json 2>/dev/null; cat gpurun_out/kbench_r02f_nodefer.json
