"""Multi-GPU plumbing: contiguous agent blocks + the per-step all-gather of the agent records.

The reference's only "communication" is MultiSyncSimulator::broadcastMsgs copying every agent's Obstacle
record (state, goal, previous trajectory) into every other agent (src/multi_sync_simulator.cpp:468-514).
Sharded over GPUs that copy becomes one all-gather per replan step of the fixed-size records
(include/dlsc_b200.h "Records"): NCCL on device memory in production, gloo on host memory in the CPU tests.
"""
import os

import numpy as np


def agent_block(n_agents, world_size, rank):
    """Contiguous block [begin, begin + n_local) owned by `rank` (blocks differ by at most one agent)."""
    base, extra = divmod(int(n_agents), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


class RecordExchange:
    """All-gather of the records of one SwarmPlanner.

    p2p mode:    (default on GPUs) every rank stores its slice straight into the peers' record arrays over NVLink
                 (dlsc_p2p_export / dlsc_p2p_connect / dlsc_exchange_records); torch.distributed only carries the
                 64-byte IPC handles once;
    device mode: the planner's record array is bound to a torch CUDA tensor and gathered in place (NCCL);
    host mode:   records are fetched / stored through the C ABI and gathered with the default (gloo) group.
    """

    def __init__(self, planner, world_size, rank, device=None, mode="p2p"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.pl, self.world, self.rank = planner, int(world_size), int(rank)
        self.blocks = [agent_block(planner.N, world_size, r) for r in range(world_size)]
        self.equal = len({n for _, n in self.blocks}) == 1
        rf = planner.rec_floats
        self.device = device
        self.mode = "host" if device is None else mode
        if device is not None and mode == "p2p" and self.world > 1:
            # every rank must end up in the same mode: agree on success after each phase, else fall back to NCCL
            def all_ok(ok):
                t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                return bool(t.item())
            handle, err = None, None
            try:
                handle = planner.p2p_export()
            except Exception as e:                       # e.g. IPC not permitted in this container
                err = e
            if all_ok(handle is not None):
                handles = [None] * self.world
                dist.all_gather_object(handles, handle)
                try:
                    if os.environ.get("DLSC_P2P_FORCE_FAIL") == str(self.rank):      # test hook for the fallback path
                        raise RuntimeError("forced failure (DLSC_P2P_FORCE_FAIL)")
                    planner.p2p_connect(self.world, self.rank, handles)
                except Exception as e:
                    err = e
                if all_ok(err is None):
                    return
                planner.p2p_disconnect()
            self.fallback_reason = repr(err) if err is not None else "a peer could not set up peer-memory access"
            mode = "nccl"
        if device is not None and self.world == 1:
            self.mode = "single"
            return
        self.mode = "nccl" if device is not None else "host"
        if device is not None:
            self.rec = torch.zeros(planner.N * rf, dtype=torch.float32, device=device)
            planner.bind_records(self.rec.data_ptr())
            b, n = self.blocks[rank]
            self.local = self.rec[b * rf:(b + n) * rf]
            self.parts = [self.rec[b2 * rf:(b2 + n2) * rf] for b2, n2 in self.blocks]

    def gather(self):
        if self.world == 1:
            return
        if self.mode == "p2p":
            self.pl.exchange_records()
            return
        if self.device is not None:
            if self.equal:
                self.dist.all_gather_into_tensor(self.rec, self.local)
            else:
                self.dist.all_gather(self.parts, self.local.clone())
            return
        b, n = self.blocks[self.rank]
        mine = self.torch.from_numpy(np.ascontiguousarray(self.pl.get_records(b, n)))
        outs = [self.torch.zeros(n2, self.pl.rec_floats) for _, n2 in self.blocks]
        self.dist.all_gather(outs, mine)
        for r, (b2, n2) in enumerate(self.blocks):
            if r != self.rank:
                self.pl.set_records(b2, outs[r].numpy())
