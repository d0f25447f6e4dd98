#!/bin/bash
# ab_agents.sh <agents> <variant>...: stage times of each prebuilt variant at a given swarm size (1 GPU proxy of a shard)
n=$1; shift
for v in "$@"; do
  if [ "$v" = main ]; then unset DLSC_B200_LIB; else export DLSC_B200_LIB=$PWD/gpurun_variants/$v/libdlsc_b200.so; fi
  timeout 300 python bench.py --no-cpu-baseline --mc-missions 0 --closed-loop-steps 0 --dyn-obstacles 0 --steps 30 --agents $n --max-nbr 192 2> gpurun_out/ab_$v.err | tail -1 > gpurun_out/ab_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/ab_$v.json'))
print('$v agents $n', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['stages_ms'].items()}, d['pilot']['replay_exact'])" || tail -3 gpurun_out/ab_$v.err
done
