#!/bin/bash
# round-2 closing evidence, part 2: full GPU test-suite, the EMT configurations after the Davidson-cap fix, C3 as named
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/y_pytest.log
run_bench () { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/bench_r2_$name.json 2> gpurun_out/bench_r2_$name.err; echo "bench $name rc=$?"; tail -c 300 gpurun_out/bench_r2_$name.err; }
run_bench C2_emt-cluster --workload emt-cluster --steps 20 --warmup 5 --long-steps 0
run_bench C2_emt-cluster_noprojrot --workload emt-cluster --no-proj-rot --steps 20 --warmup 5 --long-steps 0
run_bench C3_emt-slab --workload emt-slab --steps 20 --warmup 5 --long-steps 0
run_bench C3_emt-slab_internal --workload emt-slab --internal --steps 6 --warmup 3
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r2_C*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        p=d.get("parity") or {}
        print(f, "%.0f"%d["value"], "%.3f ms"%d["ms_per_step"], "e2e %.0f"%d["e2e"]["value"], "parity", p.get("max_dx"), p.get("max_rel_lam"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("systems_flagged"))
    except Exception as e:
        print(f, "ERR", e)
PY
