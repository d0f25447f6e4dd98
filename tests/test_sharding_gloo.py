"""world_size-2 gloo run on CPU: two ranks each own half of the agents (host simulator backend), exchange
records through RecordExchange after every step, and must reproduce the single-context rollout bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _parity
from dlsc_gc_planner_b200 import capi, sharding


def test_agent_block_partition():
    for n, w in ((10, 2), (4096, 8), (11, 4), (3, 5)):
        blocks = [sharding.agent_block(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        for (b0, c0), (b1, _) in zip(blocks, blocks[1:]):
            assert b0 + c0 == b1


def _worker(rank, world, port, steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = capi.load_library(_parity.HOSTSIM_SO)
    cfg, m = _parity.load_case("empty10")
    begin, nl = sharding.agent_block(m.n_agents, world, rank)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, begin=begin, n_local=nl, lib=lib)
    ex = sharding.RecordExchange(pl, world, rank)
    ex.gather()                                   # the reset records of the other block
    wp = m.start.copy()
    trajs = []
    for step in range(steps):
        wp = wp + np.float32(0.1) * np.sign(m.goal - wp)
        pl.set_agents(waypoint=wp[begin:begin + nl])
        pl.plan()
        pl.advance()
        ex.gather()
        trajs.append(pl.traj())
    np.save(os.path.join(out_dir, "traj_%d.npy" % rank), np.array(trajs))
    np.save(os.path.join(out_dir, "rec_%d.npy" % rank), pl.get_records())
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_context(hostsim, tmp_path):
    steps, world = 8, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, steps, str(tmp_path)), nprocs=world, join=True)
    cfg, m = _parity.load_case("empty10")
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=hostsim)
    wp = m.start.copy()
    ref = []
    for step in range(steps):
        wp = wp + np.float32(0.1) * np.sign(m.goal - wp)
        pl.set_agents(waypoint=wp)
        pl.plan()
        pl.advance()
        ref.append(pl.traj())
    ref = np.array(ref)
    t0 = np.load(tmp_path / "traj_0.npy")
    t1 = np.load(tmp_path / "traj_1.npy")
    assert np.array_equal(np.concatenate([t0, t1], axis=1), ref)
    assert np.array_equal(np.load(tmp_path / "rec_0.npy"), pl.get_records())
    assert np.array_equal(np.load(tmp_path / "rec_1.npy"), pl.get_records())
