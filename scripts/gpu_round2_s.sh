#!/bin/bash
# evidence run: named configurations, ncu captures, sanitizer
mkdir -p gpurun_out
run_bench () { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/s_bench_$name.json 2> gpurun_out/s_bench_$name.err; echo "bench $name rc=$?"; tail -c 300 gpurun_out/s_bench_$name.err; }
run_bench quadratic --steps 20 --warmup 5 --long-steps 200
run_bench C2_emt-cluster --workload emt-cluster --steps 20 --warmup 5 --long-steps 0
run_bench C2_emt-cluster_noprojrot --workload emt-cluster --no-proj-rot --steps 20 --warmup 5 --long-steps 0
run_bench C3_emt-slab --workload emt-slab --steps 20 --warmup 5 --long-steps 0
run_bench C4_512x768 --batch 512 --n 768 --kdiag 5 --steps 20 --warmup 5 --long-steps 100
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.0f"%d["value"], "%.3f ms"%d["ms_per_step"], "e2e %.0f"%d["e2e"]["value"], "parity", d["parity"]["max_dx"], d["parity"]["max_rel_lam"], "cpu", d.get("cpu_baseline",{}).get("value"), d["config"].get("spectrum"))
    except Exception as e:
        print(f, "ERR", e)
PY
# ncu --set full captures (raw page as csv)
for spec in "hv:hv_tma_kernel<1>:2" "apply:secular_apply_kernel:2" "solve:secular_update_kernel:2" "append:append_:2"; do
  tag=${spec%%:*}; rest=${spec#*:}; rx=${rest%%:*}; cnt=${rest##*:}
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o gpurun_out/s_ncu_$tag -f python scripts/profile_step.py --warm 20 --steps 2 > gpurun_out/s_ncu_$tag.log 2>&1; echo "ncu $tag rc=$?"
  ncu -i gpurun_out/s_ncu_$tag.ncu-rep --page raw --csv > gpurun_out/s_ncu_$tag.csv 2>/dev/null
done
ls -la gpurun_out/s_ncu_* | head
# sanitizer on the small compact cases
SB_SPLIT_MIN_ROWS=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/debug_compact.py > gpurun_out/s_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/s_memcheck.log
SB_SPLIT_MIN_ROWS=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/debug_compact.py > gpurun_out/s_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/s_racecheck.log
