#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_internal.py > gpurun_out/w_debug.log 2>&1; echo "debug rc=$?"; tail -14 gpurun_out/w_debug.log
timeout 1500 python -m pytest tests/test_internal_pes.py tests/test_gpu_kernels.py -m gpu -q --tb=short > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/w_pytest.log
