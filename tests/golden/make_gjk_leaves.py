"""Generates tests/golden/gjk_leaves.npz: known-answer vectors of the reference's own openGJK object code
(oracle/_ref, compiled unmodified from /root/reference/src/openGJK/openGJK.cpp) chosen so that every reachable
leaf of the distance sub-algorithm's decision tree (S1D / S2D / S3D, openGJK.cpp:243-631; leaf numbering in
dlsc_gc_planner_b200/csrc/dlsc_math.cuh) is pinned by at least WANT hulls.

The leaf a hull goes through is read from the kernel core itself (gjk::hull_origin with the MaskTrace tracer,
executed on the CPU by tests/hostsim); the expected witness vector / distance / simplex size come from the
reference object code only.  Run in the build container (needs /root/reference for oracle/_ref):

    python tests/golden/make_gjk_leaves.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _parity  # noqa: E402
from dlsc_gc_planner_b200 import capi  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N_LEAVES = 57
WANT = 60
BATCH = 200000


def families(rng, n):
    """n hulls [n][6][3] (float32-representable doubles) from a mix of geometric families."""
    out = []
    k = n // 10
    c = rng.normal(size=(k, 1, 3)) * rng.uniform(0, 2, size=(k, 1, 1))
    out.append(c + rng.normal(size=(k, 6, 3)) * rng.uniform(0.01, 1.0, size=(k, 1, 1)))                 # generic
    out.append(rng.normal(size=(k, 6, 3)) * rng.uniform(0.05, 1.0, size=(k, 1, 1)))                      # around the origin
    d = rng.normal(size=(k, 1, 3))
    t = np.linspace(0, 1, 6)[None, :, None] ** rng.uniform(0.5, 2, size=(k, 1, 1))
    out.append(c + t * d + rng.normal(size=(k, 6, 3)) * 1e-3)                                            # needles
    # flat hulls: points of a plane, tiny thickness
    u = rng.normal(size=(k, 1, 3)); w = rng.normal(size=(k, 1, 3))
    ab = rng.normal(size=(k, 6, 2))
    out.append(c * 0.3 + ab[..., :1] * u + ab[..., 1:] * w + rng.normal(size=(k, 6, 3)) * 10.0 ** rng.uniform(-7, -2, size=(k, 1, 1)))
    # lattice coordinates: exact ties in the half-space tests
    out.append(rng.integers(-3, 4, size=(k, 6, 3)) * 0.5 + rng.integers(-2, 3, size=(k, 1, 3)) * 0.25)
    out.append(rng.integers(-2, 3, size=(k, 6, 3)).astype(np.float64))
    # widely different scales per point (obtuse, sliver tetrahedra)
    out.append(c * 0.2 + rng.normal(size=(k, 6, 3)) * 10.0 ** rng.uniform(-3, 0.5, size=(k, 6, 1)))
    # far vertices behind a near face: origin close to a face / edge of a long thin body
    e = rng.normal(size=(k, 1, 3)); e /= np.linalg.norm(e, axis=2, keepdims=True)
    out.append(e * rng.uniform(1e-4, 0.3, size=(k, 1, 1)) + rng.normal(size=(k, 6, 3)) * np.array([1.0, 1.0, 0.02]) * rng.uniform(0.1, 3, size=(k, 1, 1)))
    # duplicates and repeated coordinates (Bernstein polygons of a hovering agent: all six points equal or nearly)
    base = rng.normal(size=(k, 1, 3)) * rng.uniform(0, 1, size=(k, 1, 1))
    dup = np.repeat(base, 6, axis=1)
    sel = rng.integers(0, 6, size=(k, 3))
    for j in range(3):
        dup[np.arange(k), sel[:, j]] += rng.normal(size=(k, 3)) * 10.0 ** rng.uniform(-6, 0, size=(k, 1))
    out.append(dup)
    # hulls touching the origin: a vertex / edge midpoint / face point exactly (or almost) at the origin
    kk = n - 9 * k
    p = rng.normal(size=(kk, 6, 3)) * rng.uniform(0.05, 1.0, size=(kk, 1, 1))
    mode = rng.integers(0, 3, size=kk)
    shift = np.where(mode[:, None] == 0, p[:, 0], np.where(mode[:, None] == 1, 0.5 * (p[:, 0] + p[:, 1]), (p[:, 0] + p[:, 1] + p[:, 2]) / 3))
    out.append(p - shift[:, None, :] * (1 + rng.choice([0.0, 1e-7, -1e-7, 1e-4], size=(kk, 1, 1))))
    pts = np.concatenate(out, axis=0)
    return pts.astype(np.float32).astype(np.float64)


def main():
    O.build()
    if O.ref_lib() is None:
        raise SystemExit("oracle/_ref not built (needs /root/reference)")
    lib = _parity.hostsim_lib()
    cfg, m = _parity.load_case("empty10")
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=lib)
    rng = np.random.default_rng(20261018)
    have = np.zeros(N_LEAVES, np.int64)
    keep = []
    for rnd in range(25):
        pts = families(rng, BATCH)
        _, _, _, lv = pl.gjk_batch(pts)
        bits = ((lv[:, None] >> np.arange(N_LEAVES, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
        order = np.argsort(have)                      # rarest leaves pick first
        taken = np.zeros(len(pts), bool)
        for b in order:
            if have[b] >= WANT:
                continue
            idx = np.flatnonzero(bits[:, b] & ~taken)[: WANT - have[b]]
            taken[idx] = True
        sel = np.flatnonzero(taken)
        have += bits[sel].sum(axis=0)
        keep.append(pts[sel])
        missing = [int(b) for b in range(N_LEAVES) if have[b] < WANT]
        print(f"round {rnd}: kept {sum(len(k) for k in keep)} hulls, leaves below {WANT}: {missing}", flush=True)
        if not missing:
            break
    # rare leaves: rotated / rescaled / slightly perturbed variants of the hulls that reached them
    pts = np.concatenate(keep, axis=0)
    _, _, _, lv = pl.gjk_batch(pts)
    for b in [b for b in range(N_LEAVES) if 0 < have[b] < WANT]:
        seeds = pts[((lv >> np.uint64(b)) & np.uint64(1)).astype(bool)]
        for _ in range(40):
            k = 4000
            src = seeds[rng.integers(0, len(seeds), size=k)]
            q, _ = np.linalg.qr(rng.normal(size=(k, 3, 3)))
            var = np.einsum("kij,kpj->kpi", q, src) * rng.uniform(0.3, 3, size=(k, 1, 1))
            var = var * (1 + rng.normal(size=(k, 6, 3)) * 10.0 ** rng.uniform(-6, -2, size=(k, 1, 1)))
            var = var.astype(np.float32).astype(np.float64)
            _, _, _, lv2 = pl.gjk_batch(var)
            idx = np.flatnonzero(((lv2 >> np.uint64(b)) & np.uint64(1)).astype(bool))[: WANT - have[b]]
            have[b] += len(idx)
            keep.append(var[idx])
            if have[b] >= WANT:
                break
        print(f"leaf {b}: {have[b]} after the variant search", flush=True)
    pts = np.concatenate(keep, axis=0)
    vs, ds, sn = [], [], []
    for p in pts:
        d, v, s = O.ref_gjk(p)
        vs.append(v); ds.append(d); sn.append(s)
    v2, it, sn2, lv = pl.gjk_batch(pts)
    np.savez_compressed(os.path.join(OUT, "gjk_leaves.npz"), pts=pts.astype(np.float32), v=np.array(vs), d=np.array(ds),
                        simplex=np.array(sn, np.int32), leaves=lv, iters=it)
    bits = ((lv[:, None] >> np.arange(N_LEAVES, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    print("hulls", len(pts), "leaf histogram", bits.sum(axis=0).tolist())
    print("kernel core vs reference object code: mismatching hulls", int((v2 != np.array(vs)).any(axis=1).sum()))


if __name__ == "__main__":
    main()
