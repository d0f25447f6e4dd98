#!/bin/bash
# gpu_job_b.sh <tag> <variants...>: GPU tests (fail-fast) then A/B bench of the in-tree build and the named variants
tag=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
bash scripts/ab_run.sh main "$@" 2>&1 | tee gpurun_out/${tag}_ab.txt
