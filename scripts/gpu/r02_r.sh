#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/experiments/e2e_probe.py cfg2 > gpurun_out/r02_e2e_probe.json 2> gpurun_out/r02_e2e_probe.err; echo "rc=$?"; cat gpurun_out/r02_e2e_probe.json; tail -3 gpurun_out/r02_e2e_probe.err
