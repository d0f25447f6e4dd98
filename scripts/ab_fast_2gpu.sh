#!/bin/bash
# ab_fast_2gpu.sh: QP first-scan path A/B on 2 GPUs (2048 agents per rank of the 4096-agent forest)
for v in 1 0; do
  DLSC_QP_FAST=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$v bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline --mc-missions 0 2> gpurun_out/ab_fast2.err | tail -1 > gpurun_out/ab_fast2.json
  python - "$v" <<'P'
import json, sys
d = json.load(open("gpurun_out/ab_fast2.json"))
print("2 GPUs fast", sys.argv[1], round(d["ms_per_step"], 4), "qp", round(d["stages_ms"]["qp"], 4), d["pilot"]["replay_exact"])
P
done
