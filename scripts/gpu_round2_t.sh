#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/t_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --long-steps 200 --no-cpu-baseline > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t_bench.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["max_dx"], d["parity"]["max_rel_lam"], d["kernel_ms"])
print(d["long_run"]["decile_ms_per_step"], d["long_run"]["rows_at_decile_end"])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t_launches.csv python scripts/profile_step.py > gpurun_out/t_prof.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/t_launches.csv > gpurun_out/t_launches.txt; head -16 gpurun_out/t_launches.txt; tail -1 gpurun_out/t_launches.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hv_tma_kernel -c 3 -o gpurun_out/t_ncu_hv -f python scripts/profile_step.py --warm 20 --steps 2 > gpurun_out/t_ncu_hv.log 2>&1; echo "ncu hv rc=$?"
ncu -i gpurun_out/t_ncu_hv.ncu-rep --page raw --csv > gpurun_out/t_ncu_hv.csv 2>/dev/null; ls -la gpurun_out/t_ncu_hv.csv
SB_SPLIT_MIN_ROWS=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/debug_compact.py > gpurun_out/t_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/t_racecheck.log
