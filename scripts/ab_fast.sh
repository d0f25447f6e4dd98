#!/bin/bash
# ab_fast.sh <agents...>: QP first-scan path A/B (DLSC_QP_FAST=1 warp per agent, 0 CTA per agent) at several swarm sizes
for a in "$@"; do
  for v in 1 0; do
    DLSC_QP_FAST=$v python bench.py --agents $a --no-cpu-baseline --mc-missions 0 --closed-loop-steps 0 --dyn-obstacles 0 --steps 30 2> gpurun_out/ab_fast.err | tail -1 > gpurun_out/ab_fast.json
    python - "$a" "$v" <<'P'
import json, sys
d = json.load(open("gpurun_out/ab_fast.json"))
print("agents", sys.argv[1], "fast", sys.argv[2], round(d["ms_per_step"], 4), "qp", round(d["stages_ms"]["qp"], 4), d["pilot"]["replay_exact"])
P
  done
done
