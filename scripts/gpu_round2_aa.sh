#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_internal_pes.py -m gpu -q --tb=short > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/aa_pytest.log
timeout 1500 python bench.py --workload emt-slab --internal --steps 6 --warmup 3 > gpurun_out/bench_r2_C3_emt-slab_internal.json 2> gpurun_out/bench_r2_C3_emt-slab_internal.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_r2_C3_emt-slab_internal.err
timeout 900 python bench.py --workload emt-slab --internal --inexact-geodesic --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_C3_emt-slab_internal_frozenBinv.json 2> gpurun_out/bench_r2_C3_emt-slab_internal_frozenBinv.err; echo "bench2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r2_C3_emt-slab_internal*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.0f"%d["value"], "%.1f ms"%d["ms_per_step"], "e2e %.0f"%d["e2e"]["value"], "parity", d["parity"]["max_dx"], "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("systems_flagged"), "rk/step", d["geodesic_steps_per_call"], {k: round(v,1) for k,v in d["phase_ms"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/aa_launches_internal.csv python scripts/profile_internal.py --batch 256 --warm 3 --steps 3 > gpurun_out/aa_prof_internal.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/aa_launches_internal.csv > gpurun_out/aa_launches_internal.txt; head -24 gpurun_out/aa_launches_internal.txt; tail -1 gpurun_out/aa_launches_internal.txt
rm -f gpurun_out/aa_launches_internal.csv
