#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/experiments/peer_read_probe.py > gpurun_out/peer_read_probe_g2.json 2> gpurun_out/peer_read_probe_g2.err; echo "rc=$?"; cat gpurun_out/peer_read_probe_g2.json; tail -3 gpurun_out/peer_read_probe_g2.err
