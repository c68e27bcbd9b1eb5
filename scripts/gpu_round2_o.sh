#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/secular_phases2.py > gpurun_out/o_phases.log 2>&1; tail -4 gpurun_out/o_phases.log
