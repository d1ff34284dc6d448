#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/alg1_probe.py > gpurun_out/r02_alg1_probe.log 2>&1; tail -2 gpurun_out/r02_alg1_probe.log
