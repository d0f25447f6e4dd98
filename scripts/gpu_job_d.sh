#!/bin/bash
# gpu_job_d.sh <tag>: GPU tests (all), default bench, memcheck + racecheck over a dynamic-obstacle rollout.
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py 2> gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json; echo "bench rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_dynamic_obstacles.py -m gpu -x -q -k "empty10 or spin4" > gpurun_out/${tag}_memcheck_dyn.log 2>&1; echo "memcheck rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_dynamic_obstacles.py -m gpu -x -q -k "empty10 or spin4" > gpurun_out/${tag}_racecheck_dyn.log 2>&1; echo "racecheck rc=$?"
tail -n 3 gpurun_out/${tag}_memcheck_dyn.log gpurun_out/${tag}_racecheck_dyn.log
head -c 400 gpurun_out/${tag}_bench.json
