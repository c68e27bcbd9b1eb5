#!/bin/bash
# internal-coordinate engine bring-up + the EMT Davidson-cap fix
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_internal_pes.py -m gpu -x -q --tb=long > gpurun_out/v_pytest_internal.log 2>&1; echo "internal rc=$?"; tail -40 gpurun_out/v_pytest_internal.log
timeout 900 python -m pytest tests/test_emt.py -m gpu -x -q --tb=short > gpurun_out/v_pytest_emt.log 2>&1; echo "emt rc=$?"; tail -5 gpurun_out/v_pytest_emt.log
