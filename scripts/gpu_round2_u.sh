#!/bin/bash
mkdir -p gpurun_out
for spec in "slab:--workload emt-slab --batch 1024 --n 384" "cluster:--workload emt-cluster --batch 256 --n 192 --kdiag 2 --no-proj-rot" "clusterrot:--workload emt-cluster --batch 256 --n 192 --kdiag 2"; do
  tag=${spec%%:*}; args=${spec#*:}
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/u_launches_$tag.csv python scripts/profile_step.py $args --warm 13 --steps 6 > gpurun_out/u_prof_$tag.log 2>&1; echo "ncu $tag rc=$?"
  python scripts/summarize_launches.py gpurun_out/u_launches_$tag.csv > gpurun_out/u_launches_$tag.txt; head -14 gpurun_out/u_launches_$tag.txt; tail -1 gpurun_out/u_launches_$tag.txt
done
