#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_internal_pes.py tests/test_gpu_kernels.py -m gpu -q --tb=short -x > gpurun_out/bb_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/bb_pytest.log
timeout 900 python bench.py --workload emt-slab --internal --batch 256 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bb_bench_256.json 2> gpurun_out/bb_bench_256.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bb_bench_256.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bb_bench_256.json").read().strip().splitlines()[-1])
    print("%.0f"%d["value"], "%.1f ms"%d["ms_per_step"], "e2e %.0f"%d["e2e"]["value"], "parity", d["parity"]["max_dx"], d.get("systems_flagged"), "rk/step", d["geodesic_steps_per_call"], {k: round(v,1) for k,v in d["phase_ms"].items()})
except Exception as e:
    print("ERR", e)
PY
