#!/bin/bash
# after the ql_kernel barrier fix: racecheck of the internal-coordinate step again, then the whole GPU test-suite
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/debug_internal.py > gpurun_out/hh_racecheck_internal.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/hh_racecheck_internal.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/hh_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/hh_pytest.log
