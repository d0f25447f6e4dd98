"""BASELINE configs[4]: a Monte-Carlo batch of independent empty50 missions replanning in lockstep inside one
context.  Every mission of the batch must replan exactly as it does alone: each slice is compared with its own
oracle (bit-exact geometry, QP within tolerance), and no agent may see an agent of another mission."""
import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi, missions
from test_hostsim_parity import check_worst


class BatchedOracle:
    """The oracles of the missions presented as one swarm (the attributes _parity.force_state / compare_step read)."""

    CAT = ("traj", "pos", "vel", "goal_cur", "radius", "downwash", "acc", "waypoint", "disturbed", "init_traj",
           "nbr_cnt", "lsc_normal", "lsc_d", "lsc_anchor", "status", "cost", "qp_x", "qp_iters")

    def __init__(self, oracles):
        self.o = oracles
        self.p = oracles[0].p
        self.N = sum(o.N for o in oracles)
        self.off = np.cumsum([0] + [o.N for o in oracles])

    def __getattr__(self, name):
        if name in BatchedOracle.CAT:
            return np.concatenate([getattr(o, name) for o in self.o])
        if name == "pred_traj":
            return np.concatenate([o.pred_traj for o in self.o])
        if name == "nbr_idx":
            return np.concatenate([o.nbr_idx + self.off[i] for i, o in enumerate(self.o)])
        if name == "seq":
            return self.o[0].seq
        raise AttributeError(name)


def run_batch(lib, n_missions, n_each, steps, device=0):
    cfg, base = _parity.load_case("empty50")
    base = _parity.subset(base, n_each)
    ms = [missions.add_goal_noise(base, 0.2, cfg.dim, seed=i) for i in range(n_missions)]
    batch, group = missions.concat_missions(ms)
    K = n_each - 1
    oracles = [_parity.make_oracle(cfg, m, K) for m in ms]
    sw = BatchedOracle(oracles)
    pl = capi.SwarmPlanner(cfg, batch, max_nbr=K, lib=lib, device=device)
    wfs = [_parity.default_waypoints(cfg, m) for m in ms]
    worst = {}
    for _ in range(steps):
        for o, wf in zip(oracles, wfs):
            o.waypoint = wf(o)
        rec = _parity.records_from_oracle(sw, pl.rec_floats)
        rec[:, cfg.M * (cfg.n + 1) * 3 + 11] = group                  # the mission index travels in the record
        pl.set_records(0, rec)
        pl.set_agents(acc=sw.acc, waypoint=sw.waypoint, disturbed=sw.disturbed)
        pl.seq = sw.seq
        for o in oracles:
            o.step()
        pl.plan()
        d = _parity.compare_step(pl, sw)
        idx, cnt = pl.neighbours()
        valid = np.arange(pl.K)[None, :] < cnt[:, None]
        assert (group[np.where(valid, idx, 0)] == group[:, None])[valid].all(), "neighbour from another mission"
        _parity.merge_max(worst, d)
        for o in oracles:
            o.advance()
    pl.close()
    return worst


def test_groups_via_api_match_records(hostsim):
    """dlsc_set_groups writes the same record slot the batch test fills by hand."""
    cfg, base = _parity.load_case("empty10")
    batch, group = missions.concat_missions([base, base])
    pl = capi.SwarmPlanner(cfg, batch, max_nbr=9, lib=hostsim)
    pl.set_groups(group)
    rec = pl.get_records()
    assert np.array_equal(rec[:, cfg.M * (cfg.n + 1) * 3 + 11], group.astype(np.float32))
    pl.plan()
    idx, cnt = pl.neighbours()
    assert (cnt == 9).all()                                            # range -1: everyone of the own mission, nobody else
    assert (idx[:10, :9] < 10).all() and (idx[10:, :9] >= 10).all()
    pl.close()


def test_montecarlo_batch_cpu(hostsim):
    check_worst(run_batch(hostsim, n_missions=3, n_each=12, steps=8))


@pytest.mark.gpu
def test_montecarlo_batch_gpu(cuda_lib):
    """8 x empty50 (400 agents, K = 49) on the GPU, every mission against its own oracle."""
    check_worst(run_batch(cuda_lib, n_missions=8, n_each=50, steps=6))
