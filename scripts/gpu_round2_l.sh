#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_drift.py 96 200 30 > gpurun_out/l_drift.log 2>&1; echo "rc=$?"; grep "jump" gpurun_out/l_drift.log | head; tail -2 gpurun_out/l_drift.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/l_pytest.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_launches.csv python scripts/profile_step.py > gpurun_out/l_prof.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/l_launches.csv > gpurun_out/l_launches.txt; head -30 gpurun_out/l_launches.txt; tail -1 gpurun_out/l_launches.txt
