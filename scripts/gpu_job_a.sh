#!/bin/bash
# gpu_job_a.sh <tag>: GPU tests, default bench, ncu launch list + one full capture (with source) of the hot kernels,
# compute-sanitizer memcheck / racecheck over smoke().  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py 2> gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json; echo "bench rc=$?"
timeout 600 env DLSC_OVERLAP=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mc-missions 0 --closed-loop-steps 0 --dyn-obstacles 0 > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 env DLSC_OVERLAP=0 ncu --set full --clock-control none --import-source on \
    -k regex:'k_lsc|k_sfc|k_qp_gi|k_qp_fast|k_nbr_bin|k_nbr_search|k_goal|k_predict|k_advance' -s 270 -c 9 -o gpurun_out/${tag}_prof -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mc-missions 0 --closed-loop-steps 0 --dyn-obstacles 0 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/${tag}_memcheck.log gpurun_out/${tag}_racecheck.log
cat gpurun_out/${tag}_bench.json | head -c 600
