#!/bin/bash
# ab_run.sh <variant>...: bench each prebuilt variant (gpurun_variants/<v>/libdlsc_b200.so; "main" = the in-tree build)
for v in "$@"; do
  if [ "$v" = main ]; then unset DLSC_B200_LIB; else export DLSC_B200_LIB=$PWD/gpurun_variants/$v/libdlsc_b200.so; fi
  timeout 300 python bench.py --no-cpu-baseline --mc-missions 0 --closed-loop-steps 0 --dyn-obstacles 0 --steps 30 2> gpurun_out/ab_$v.err | tail -1 > gpurun_out/ab_$v.json
  python -c "
import sys,json
d=json.load(open('gpurun_out/ab_$v.json'))
print('$v', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['stages_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],4), d['pilot']['replay_exact'], d['e2e']['replay_exact'])" || tail -3 gpurun_out/ab_$v.err
done
