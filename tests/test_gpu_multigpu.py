"""Multi-GPU tier (needs >= 2 GPUs on the box, skipped otherwise): the sharded bench under torchrun with both
record exchanges -- stores over NVLink peer memory (dlsc_exchange_records) and the NCCL all-gather -- must replay
the single-context pilot rollout bit for bit (resident chain and host-buffer e2e path)."""
import json
import os
import subprocess
import sys

import pytest

import _parity
from dlsc_gc_planner_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_rank_bench_replays_exactly(cuda_lib, exchange):
    if cuda_lib.dlsc_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531" if exchange == "p2p" else "29532", os.path.join(_parity.ROOT, "bench.py"), "--gpus", "2",
           "--agents", "512", "--max-nbr", "192", "--steps", "6", "--warmup", "3", "--settle", "8", "--no-cpu-baseline",
           "--mc-missions", "4", "--mc-steps", "4", "--exchange", exchange]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=_parity.ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2 and d["pilot"]["replay_exact"] and d["e2e"]["replay_exact"]
    assert d["pilot"]["qp_failsafe_agents"] == 0
    assert ("peer-memory" if exchange == "p2p" else "NCCL") in d["exchange"]
    assert d["montecarlo"]["replay_exact"] and d["montecarlo"]["missions"] == 8
