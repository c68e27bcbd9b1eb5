#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_internal_pes.py -m gpu -q --tb=short > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/x_pytest.log
timeout 900 python bench.py --workload emt-slab --internal --batch 64 --steps 3 --warmup 2 > gpurun_out/x_bench_small.json 2> gpurun_out/x_bench_small.err; echo "bench small rc=$?"; tail -c 1500 gpurun_out/x_bench_small.err; head -c 3000 gpurun_out/x_bench_small.json
