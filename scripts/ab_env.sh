#!/bin/bash
# ab_env.sh VAR v1 v2 ...: bench the in-tree build with VAR set to each value
var=$1; shift
for v in "$@"; do
  export $var=$v
  timeout 300 python bench.py --no-cpu-baseline --mc-missions 0 --closed-loop-steps 0 --dyn-obstacles 0 --steps 30 2> gpurun_out/ab_env.err | tail -1 > gpurun_out/ab_env.json
  python -c "
import json
d=json.load(open('gpurun_out/ab_env.json'))
print('$var=$v', round(d['ms_per_step'],4), 'p50', round(d['p50_step_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['pilot']['replay_exact'], d['e2e']['replay_exact'], 'launches', d['gpu_launches'])" || tail -3 gpurun_out/ab_env.err
done
