#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_compact.py > gpurun_out/d_debug.log 2>&1; echo "debug rc=$?"
grep -c "dx" gpurun_out/d_debug.log; grep "step 13\|step  9\|raised\|diverged" gpurun_out/d_debug.log | head
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/d_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --long-steps 200 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/d_bench.json; tail -5 gpurun_out/d_bench.err
