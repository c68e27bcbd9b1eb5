#!/bin/bash
mkdir -p gpurun_out
timeout 270 python bench.py --workload emt-slab --internal --steps 6 --warmup 3 > gpurun_out/ii_bench_internal.json 2> gpurun_out/ii_bench_internal.err; echo "bench rc=$?"; tail -c 300 gpurun_out/ii_bench_internal.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/ii_bench_internal.json").read().strip().splitlines()[-1])
print("%.0f"%d["value"], "%.1f ms"%d["ms_per_step"], "e2e %.0f"%d["e2e"]["value"], "parity", d["parity"]["max_dx"], "cpu", (d.get("cpu_baseline") or {}).get("value"), d["roofline_gemm"]["traffic"], d["roofline_gemm"]["achieved"], d["roofline_gemm"]["frac"])
PY
