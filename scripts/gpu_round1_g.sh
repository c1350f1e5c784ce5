#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/gpu_configs.py > gpurun_out/configs.log 2>&1
echo "exit $?" >> gpurun_out/configs.log
cat gpurun_out/configs.log
