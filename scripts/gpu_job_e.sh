#!/bin/bash
# gpu_job_e.sh <tag>: smoke(), selected GPU tests, default bench (all legs)
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python -m pytest tests/test_compat_cpp.py tests/test_dynamic_obstacles.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py 2> gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json; echo "bench rc=$?"
python - "$tag" <<'P'
import json, sys
d = json.load(open("gpurun_out/%s_bench.json" % sys.argv[1]))
print(d["ms_per_step"], d["e2e"]["ms_per_step"], json.dumps(d.get("dynamic_obstacles")))
P
tail -3 gpurun_out/${tag}_bench.err
