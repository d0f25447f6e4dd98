#!/bin/bash
# gpu_job_ncu.sh <tag> <kernel regex> [skip] [count] [variant]: one ncu --set full capture (with source) of the named kernels in transit
tag=$1; rx=$2; skip=${3:-60}; cnt=${4:-4}; var=$5
mkdir -p gpurun_out
if [ -n "$var" ]; then export DLSC_B200_LIB=$PWD/gpurun_variants/$var/libdlsc_b200.so; fi
timeout 900 env DLSC_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/${tag}_prof -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mc-missions 0 --closed-loop-steps 0 --dyn-obstacles 0 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
