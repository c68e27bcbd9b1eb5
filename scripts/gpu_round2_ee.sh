#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dense.py tests/test_internal_pes.py -m gpu -q --tb=short > gpurun_out/ee_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/ee_pytest.log
run_bench () { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/bench_r2_$name.json 2> gpurun_out/bench_r2_$name.err; echo "bench $name rc=$?"; tail -c 300 gpurun_out/bench_r2_$name.err; }
run_bench C3_emt-slab_internal --workload emt-slab --internal --steps 6 --warmup 3
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r2_C3_emt-slab_internal.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        p=d.get("parity") or {}
        print(f, "%.0f"%d["value"], "%.1f ms"%d["ms_per_step"], "e2e %.0f"%d["e2e"]["value"], "parity", p.get("max_dx"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("systems_flagged"), "rk/step", d.get("geodesic_steps_per_call"), {k: round(v,1) for k,v in (d.get("phase_ms") or {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
