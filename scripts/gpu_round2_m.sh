#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/secular_phases2.py > gpurun_out/m_phases.log 2>&1; echo "rc=$?"; cat gpurun_out/m_phases.log
