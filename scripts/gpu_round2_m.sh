#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/secular_phases2.py > gpurun_out/m_phases.log 2>&1; echo "rc=$?"; cat gpurun_out/m_phases.log
timeout 900 python -m pytest tests/test_gpu_host_mirror.py -m gpu -q -x > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/m_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --long-steps 0 --no-cpu-baseline --parity-systems 0 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/m_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['kernel_ms'])"
