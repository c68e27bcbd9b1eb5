#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/i_pytest.log
SB_PROFILER_RANGE=1 SB_NO_SAMPLER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 6 --warmup 19 --no-cpu-baseline --parity-systems 0 --long-steps 0 > gpurun_out/i_ncu_bench.json 2> gpurun_out/i_ncu_bench.err; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/i_launches.csv > gpurun_out/i_launches.txt; head -24 gpurun_out/i_launches.txt; tail -1 gpurun_out/i_launches.txt
timeout 600 python scripts/debug_slab.py 64 26 > gpurun_out/i_slab.log 2>&1; echo "slab rc=$?"
grep "step 1[0-9]\|step 2[0-9]\|===" gpurun_out/i_slab.log
