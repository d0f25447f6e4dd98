#!/bin/bash
# gpu_job_c.sh <tag>: racecheck + memcheck of a short bench, then the A/B bench line of the in-tree build
tag=$1
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --mc-missions 0 --closed-loop-steps 0 --dyn-obstacles 0 > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "Race reported|RACECHECK SUMMARY" gpurun_out/${tag}_racecheck.log | sed "s/0x[0-9a-f]*//g" | sort | uniq -c | sort -rn | cut -c1-250 | head
bash scripts/ab_run.sh main
