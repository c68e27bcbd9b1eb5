#!/bin/bash
# final state of round 2: whole GPU test-suite, C3 as named in both geodesic modes, launch list of the internal engine
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/ff_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ff_pytest.log
run_bench () { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/bench_r2_$name.json 2> gpurun_out/bench_r2_$name.err; echo "bench $name rc=$?"; tail -c 300 gpurun_out/bench_r2_$name.err; }
run_bench C3_emt-slab_internal --workload emt-slab --internal --steps 6 --warmup 3
run_bench C3_emt-slab_internal_frozenBinv --workload emt-slab --internal --inexact-geodesic --steps 6 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r2_C3_emt-slab_internal*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        p=d.get("parity") or {}
        print(f, "%.0f"%d["value"], "%.1f ms"%d["ms_per_step"], "e2e %.0f"%d["e2e"]["value"], "parity", p.get("max_dx"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("systems_flagged"), "rk/step", d.get("geodesic_steps_per_call"), {k: round(v,1) for k,v in (d.get("phase_ms") or {}).items()}, (d.get("roofline_gemm") or {}).get("achieved"))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ff_launches_internal.csv python scripts/profile_internal.py --batch 128 --warm 3 --steps 2 > gpurun_out/ff_prof_internal.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/ff_launches_internal.csv > gpurun_out/ff_launches_internal.txt; head -22 gpurun_out/ff_launches_internal.txt; tail -1 gpurun_out/ff_launches_internal.txt
rm -f gpurun_out/ff_launches_internal.csv
