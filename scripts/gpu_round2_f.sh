#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_compact.py > gpurun_out/f_debug.log 2>&1; echo "debug rc=$?"
grep -A12 "nfix\|=== n=30 qn tr {'c\|raised\|diverged" gpurun_out/f_debug.log | tail -60
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/f_pytest.log
for wl in emt-slab; do
timeout 900 python bench.py --workload $wl --steps 20 --warmup 5 --long-steps 0 > gpurun_out/f_bench_$wl.json 2> gpurun_out/f_bench_$wl.err; echo "bench $wl rc=$?"
tail -c 600 gpurun_out/f_bench_$wl.json; tail -3 gpurun_out/f_bench_$wl.err
done
