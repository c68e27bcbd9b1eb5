#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_slab.py 256 > gpurun_out/g_slab.log 2>&1; echo "rc=$?"
cat gpurun_out/g_slab.log | tail -40
