#!/bin/bash
# evidence for the kernels added in round 2's second half: ncu --set full of the Wilson factorisation, sanitizer passes
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_kernel|chol_panel" -c 8 -o gpurun_out/gg_ncu_wilson -f python scripts/profile_wilson.py > gpurun_out/gg_ncu_wilson.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/gg_ncu_wilson.log
ncu -i gpurun_out/gg_ncu_wilson.ncu-rep --page raw --csv > gpurun_out/gg_ncu_wilson_raw.csv 2>/dev/null
python scripts/reduce_ncu.py gpurun_out/gg_ncu_wilson_raw.csv > gpurun_out/gg_ncu_wilson.csv; cut -c1-400 gpurun_out/gg_ncu_wilson.csv | head -12
rm -f gpurun_out/gg_ncu_wilson_raw.csv gpurun_out/gg_ncu_wilson.ncu-rep
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/debug_internal.py > gpurun_out/gg_memcheck_internal.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/gg_memcheck_internal.log; tail -3 gpurun_out/gg_memcheck_internal.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/debug_internal.py > gpurun_out/gg_racecheck_internal.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/gg_racecheck_internal.log
