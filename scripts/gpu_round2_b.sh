#!/bin/bash
# baseline (dense representation) bench with in-run parity + 200-step deciles
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --long-steps 200 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/b_bench.json
tail -5 gpurun_out/b_bench.err
