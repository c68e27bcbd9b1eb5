#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_drift2.py > gpurun_out/k_drift2.log 2>&1; echo "rc=$?"; tail -70 gpurun_out/k_drift2.log
