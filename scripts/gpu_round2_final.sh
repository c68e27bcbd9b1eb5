#!/bin/bash
# final evidence of round 2 (1 GPU): tests, bench lines for the headline and the named configurations
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/f_pytest.log
run_bench () { name=$1; shift; timeout 2400 python bench.py "$@" > gpurun_out/bench_r2_$name.json 2> gpurun_out/bench_r2_$name.err; echo "bench $name rc=$?"; tail -c 300 gpurun_out/bench_r2_$name.err; }
run_bench quadratic --steps 20 --warmup 5 --long-steps 200
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; echo "reference rc=$?"
run_bench quadratic_dense --steps 20 --warmup 5 --long-steps 200 --spectrum dense --no-cpu-baseline
run_bench C2_emt-cluster --workload emt-cluster --steps 20 --warmup 5 --long-steps 0
run_bench C2_emt-cluster_noprojrot --workload emt-cluster --no-proj-rot --steps 20 --warmup 5 --long-steps 0
run_bench C3_emt-slab --workload emt-slab --steps 20 --warmup 5 --long-steps 0
run_bench C4_512x768 --batch 512 --n 768 --kdiag 5 --steps 20 --warmup 5 --long-steps 100
run_bench C5_1024x1536 --batch 1024 --n 1536 --kdiag 5 --steps 8 --warmup 3 --long-steps 0 --parity-systems 2 --cpu-systems 16
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        p=d.get("parity") or {}
        print(f, "%.0f"%d["value"], "%.3f ms"%d["ms_per_step"], "e2e %.0f"%d["e2e"]["value"], "parity", p.get("max_dx"), p.get("max_rel_lam"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["config"].get("spectrum"))
    except Exception as e:
        print(f, "ERR", e)
PY
