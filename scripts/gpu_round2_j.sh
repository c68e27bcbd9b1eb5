#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_drift.py 96 200 30 > gpurun_out/j_drift.log 2>&1; echo "rc=$?"; cat gpurun_out/j_drift.log | head -60
SB_SPLIT_ROTATION=0 timeout 300 python scripts/debug_drift.py 96 200 30 > gpurun_out/j_drift_nosplit.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/j_drift_nosplit.log
