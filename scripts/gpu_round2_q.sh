#!/bin/bash
mkdir -p gpurun_out
SB_SPLIT_MIN_ROWS=1 timeout 300 python scripts/debug_compact.py > gpurun_out/q_debug.log 2>&1; echo "debug rc=$?"
awk '/dx/ {print $4}' gpurun_out/q_debug.log | sort -g | tail -2; grep "raised\|diverged" gpurun_out/q_debug.log | head -3
timeout 300 python scripts/debug_drift.py 96 200 30 > gpurun_out/q_drift.log 2>&1; grep -c jump gpurun_out/q_drift.log; tail -1 gpurun_out/q_drift.log
timeout 300 python scripts/secular_phases2.py > gpurun_out/q_phases.log 2>&1; tail -3 gpurun_out/q_phases.log
timeout 900 python bench.py --steps 20 --warmup 5 --long-steps 200 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["max_dx"], d["parity"]["max_rel_lam"], d["kernel_ms"])
print(d["long_run"]["decile_ms_per_step"], d["long_run"]["rows_at_decile_end"])
PY
