#!/bin/bash
mkdir -p gpurun_out
timeout 150 python scripts/experiments/k3_tma_probe.py > gpurun_out/r02_k3_tma_probe.json 2> gpurun_out/r02_k3_tma_probe.err; echo "rc=$?"; cat gpurun_out/r02_k3_tma_probe.json; tail -4 gpurun_out/r02_k3_tma_probe.err
