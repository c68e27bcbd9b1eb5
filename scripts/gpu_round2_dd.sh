#!/bin/bash
# closing run of round 2: whole GPU test-suite, smoke(), C3 as named (both geodesic modes), headline line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/dd_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/dd_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/dd_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/dd_smoke.log
run_bench () { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/bench_r2_$name.json 2> gpurun_out/bench_r2_$name.err; echo "bench $name rc=$?"; tail -c 300 gpurun_out/bench_r2_$name.err; }
run_bench C3_emt-slab_internal --workload emt-slab --internal --steps 6 --warmup 3
run_bench C3_emt-slab_internal_frozenBinv --workload emt-slab --internal --inexact-geodesic --steps 6 --warmup 3 --no-cpu-baseline
timeout 900 python bench.py > gpurun_out/dd_bench_default.json 2> gpurun_out/dd_bench_default.err; echo "default bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r2_C3_emt-slab_internal*.json"))+["gpurun_out/dd_bench_default.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        p=d.get("parity") or {}
        print(f, "%.0f"%d["value"], "%.1f ms"%d["ms_per_step"], "e2e %.0f"%d["e2e"]["value"], "parity", p.get("max_dx"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("systems_flagged"), "rk/step", d.get("geodesic_steps_per_call"), {k: round(v,1) for k,v in (d.get("phase_ms") or {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
