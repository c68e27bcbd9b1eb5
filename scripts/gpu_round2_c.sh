#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_compact.py > gpurun_out/c_debug.log 2>&1; echo "debug rc=$?"
tail -80 gpurun_out/c_debug.log
