#!/bin/bash
mkdir -p gpurun_out
SB_SPLIT_MIN_ROWS=1 timeout 300 python scripts/debug_compact.py > gpurun_out/h_debug.log 2>&1; echo "debug rc=$?"
grep -c "dx" gpurun_out/h_debug.log; awk '/dx/ {print $4}' gpurun_out/h_debug.log | sort -g | tail -3; grep "raised\|diverged" gpurun_out/h_debug.log | head
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/h_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --long-steps 200 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/h_bench.err
timeout 900 python bench.py --workload emt-slab --steps 20 --warmup 5 --long-steps 0 > gpurun_out/h_bench_slab.json 2> gpurun_out/h_bench_slab.err; echo "bench slab rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/h_bench.json","gpurun_out/h_bench_slab.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["max_dx"], d["parity"]["max_rel_lam"], d["kernel_ms"])
        print(d.get("long_run",{}) and d["long_run"]["decile_ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
