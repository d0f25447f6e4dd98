"""Generates the committed fixtures under tests/golden/ from the reference tree (run in the build
container, where /root/reference exists; the GPU box and the tests never read /root/reference):

  missions_all.npz  missions #1..#30 of empty10 / empty50 / empty70 / forest10 (+ forest_tro2022 worlds) / maze10_dense
                    (+ maze_icra2023/dense worlds), same fields
  missions.npz      the reference missions / worlds the BASELINE configs name (inputs only: start, goal,
                    agent properties, world box, obstacle boxes)           <- missions/*.json, world/*.csv
  golden_log.npz    the first rows of the reference's only recorded run     <- log/result_...csv
  result_head.csv   the same rows verbatim (header + 3 rows: `head -4`)       <- log/result_...csv
  golden_log_full.npz  every row of that log (t, pos/vel/acc of the 10 agents)   <- log/result_...csv
                    (maze10_dense #1, 2-D, M=10, CPLEX path): per agent pos/vel/acc at t = 0, 0.1, 0.2
  gjk_ref.npz       known-answer vectors of the reference's own openGJK object code (oracle/_ref):
                    hull point sets -> witness vector, distance, simplex size

    python tests/golden/make_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dlsc_gc_planner_b200 import missions  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def pack(prefix, m, out):
    out[prefix + "world_min"] = m.world_min
    out[prefix + "world_max"] = m.world_max
    out[prefix + "start"] = m.start
    out[prefix + "goal"] = m.goal
    for f in ("radius", "downwash", "max_vel", "max_acc", "nominal_vel"):
        out[prefix + f] = getattr(m, f)
    out[prefix + "boxes"] = m.boxes


def main():
    out = {}
    m = missions.load_mission(f"{REF}/missions/empty10/multi_random_10agents_1.json"); pack("empty10/", m, out)
    m = missions.load_mission(f"{REF}/missions/empty50/multi_random_50agents_1.json"); pack("empty50/", m, out)
    m = missions.load_mission(f"{REF}/missions/empty70/multi_random_70agents_1.json"); pack("empty70/", m, out)
    m = missions.load_mission(f"{REF}/missions/forest10/forest10_1.json")
    m.boxes = missions.load_world_csv(f"{REF}/world/forest_tro2022/forest1.csv"); pack("forest10/", m, out)
    m = missions.load_mission(f"{REF}/missions/maze10_dense/maze10_1.json", dim=2)
    m.boxes = missions.load_world_csv(f"{REF}/world/maze_icra2023/dense/maze1.csv"); pack("maze10/", m, out)
    np.savez_compressed(os.path.join(OUT, "missions.npz"), **out)

    # every mission of the families the BASELINE configs name (#1..#30; testall_DLSCGC_*.launch pairs mission k with world k)
    allm = {}
    for k in range(1, 31):
        m = missions.load_mission(f"{REF}/missions/empty10/multi_random_10agents_{k}.json"); pack(f"empty10/{k}/", m, allm)
        m = missions.load_mission(f"{REF}/missions/empty50/multi_random_50agents_{k}.json"); pack(f"empty50/{k}/", m, allm)
        m = missions.load_mission(f"{REF}/missions/empty70/multi_random_70agents_{k}.json"); pack(f"empty70/{k}/", m, allm)
        m = missions.load_mission(f"{REF}/missions/forest10/forest10_{k}.json")
        m.boxes = missions.load_world_csv(f"{REF}/world/forest_tro2022/forest{k}.csv"); pack(f"forest10/{k}/", m, allm)
        m = missions.load_mission(f"{REF}/missions/maze10_dense/maze10_{k}.json", dim=2)
        m.boxes = missions.load_world_csv(f"{REF}/world/maze_icra2023/dense/maze{k}.csv"); pack(f"maze10/{k}/", m, allm)
    np.savez_compressed(os.path.join(OUT, "missions_all.npz"), **allm)

    rows = [l.strip().split(",") for l in open(f"{REF}/log/result_1742185870.978562_DLSCGC_10agents.csv")][1:4]
    log = np.array(rows, dtype=np.float64).reshape(3, 10, 12)
    np.savez_compressed(os.path.join(OUT, "golden_log.npz"), t=log[:, 0, 1], state=log[:, :, 2:11])
    full = [l.strip().split(",") for l in open(f"{REF}/log/result_1742185870.978562_DLSCGC_10agents.csv")][1:]
    full = np.array(full, dtype=np.float64).reshape(len(full), 10, 12)
    np.savez_compressed(os.path.join(OUT, "golden_log_full.npz"), t=full[:, 0, 1], state=full[:, :, 2:11])     # all 342 rows
    with open(f"{REF}/log/result_1742185870.978562_DLSCGC_10agents.csv") as f, open(os.path.join(OUT, "result_head.csv"), "w") as o:
        for _ in range(4):
            o.write(f.readline())

    O.build()
    if O.ref_lib() is None:
        raise SystemExit("oracle/_ref not built")
    rng = np.random.default_rng(20261017)
    pts, vs, ds, sn = [], [], [], []
    for t in range(4000):
        c = rng.normal(size=3) * rng.uniform(0, 2)
        p = c + rng.normal(size=(6, 3)) * rng.uniform(0.01, 1.0)
        if t % 7 == 0:      # nearly collinear hulls like Bernstein control polygons of short segments
            d = rng.normal(size=3)
            p = c + np.outer(np.linspace(0, 1, 6) ** rng.uniform(0.5, 2), d) + rng.normal(size=(6, 3)) * 1e-3
        if t % 11 == 0:     # hulls containing / touching the origin
            p = rng.normal(size=(6, 3)) * rng.uniform(0.05, 1.0)
        p = p.astype(np.float32).astype(np.float64)
        d, v, s = O.ref_gjk(p)
        pts.append(p); vs.append(v); ds.append(d); sn.append(s)
    np.savez_compressed(os.path.join(OUT, "gjk_ref.npz"), pts=np.array(pts), v=np.array(vs), d=np.array(ds),
                        simplex=np.array(sn, np.int32))
    print("fixtures written to", OUT)


if __name__ == "__main__":
    main()
