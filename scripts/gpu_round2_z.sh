#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_internal_pes.py -m gpu -q --tb=short > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/z_pytest.log
timeout 1500 python bench.py --workload emt-slab --internal --steps 6 --warmup 3 > gpurun_out/bench_r2_C3_emt-slab_internal.json 2> gpurun_out/bench_r2_C3_emt-slab_internal.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2_C3_emt-slab_internal.err; head -c 6000 gpurun_out/bench_r2_C3_emt-slab_internal.json
