"""Generates tests/golden/world_bt.npz from the reference tree (run in the build container; the tests never read
/root/reference): mission maze10_tro2022 #1 (inputs only) and the occupied leaves of world/maze_tro2022/maze9_1.bt as
parsed by dlsc_gc_planner_b200.missions.load_world_bt (derived data: integer cubes, not the file).

    python tests/golden/make_bt_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dlsc_gc_planner_b200 import missions  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

m = missions.load_mission(f"{REF}/missions/maze10_tro2022/maze10_1.json", dim=2)
res, cubes = missions.load_world_bt(f"{REF}/world/maze_tro2022/maze9_1.bt")
head = open(f"{REF}/world/maze_tro2022/maze9_1.bt", "rb").read(200).decode("ascii", "replace")
n_nodes = int([l.split()[1] for l in head.splitlines() if l.startswith("size")][0])
out = dict(world_min=m.world_min, world_max=m.world_max, start=m.start, goal=m.goal, radius=m.radius, downwash=m.downwash,
           max_vel=m.max_vel, max_acc=m.max_acc, nominal_vel=m.nominal_vel, cubes=cubes, res=res, n_nodes=n_nodes)
np.savez_compressed(os.path.join(OUT, "world_bt.npz"), **out)
print("world_bt.npz:", len(cubes), "occupied leaves,", n_nodes, "nodes, res", res)
