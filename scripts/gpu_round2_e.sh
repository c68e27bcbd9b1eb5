#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/e_pytest.log | grep -v "^$" | tail -25
SB_PROFILER_RANGE=1 SB_NO_SAMPLER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 6 --warmup 19 --no-cpu-baseline --parity-systems 0 --long-steps 0 > gpurun_out/e_ncu_bench.json 2> gpurun_out/e_ncu_bench.err; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/e_launches.csv > gpurun_out/e_launches.txt; cat gpurun_out/e_launches.txt
