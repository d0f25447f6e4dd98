"""Host-side construction of the distance grid the SFC stage reads (data provider, once per mission).

Restates what the reference's MapManager does at start-up -- CSV boxes -> occupied voxels
(src/map_manager.cpp:264-316) -> DynamicEDTOctomap(maxdist = 1.0) (src/map_manager.cpp:61-82) -- with an
exact Euclidean distance transform (scipy) instead of dynamicEDT3D's brushfire.  Output layout is the one
`dlsc_set_edt` takes: dist [ncell] float32 metres, obst [ncell][3] int32 (nearest occupied cell or -1),
cell (x, y, z) -> (x*ny + y)*nz + z, map x = floor(coord/res) - min_key.
"""
import numpy as np


def grid_dims(world_min, world_max, res):
    inv = 1.0 / res
    lo = [int(np.floor(inv * float(np.float32(world_min[k])))) for k in range(3)]
    hi = [int(np.floor(inv * float(np.float32(world_max[k])))) for k in range(3)]
    return tuple(hi[k] - lo[k] + 1 for k in range(3)), tuple(lo)


def _round_half_away(x):
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def occupancy(world_min, world_max, res, boxes):
    dims, mk = grid_dims(world_min, world_max, res)
    occ = np.zeros(dims, bool)
    inv = 1.0 / res
    for r in np.asarray(boxes, np.float32).reshape(-1, 6).astype(np.float64):
        s = [int(_round_half_away((r[k] - 0.5 * r[3 + k]) / res)) for k in range(3)]
        e = [int(_round_half_away((r[k] + 0.5 * r[3 + k]) / res)) for k in range(3)]
        idx = []
        for k in range(3):
            i = np.arange(s[k], e[k])
            c = ((i + 0.5) * res).astype(np.float32).astype(np.float64)     # voxel centre as float32
            m = np.floor(inv * c).astype(np.int64) - mk[k]
            idx.append(m[(m >= 0) & (m < dims[k])])
        if all(len(i) for i in idx):
            occ[np.ix_(idx[0], idx[1], idx[2])] = True
    return occ, dims, mk


def build_edt(world_min, world_max, res, boxes, maxdist=1.0):
    """-> dist float32 [ncell], obst int32 [ncell,3], dims, min_key"""
    from scipy import ndimage
    occ, dims, mk = occupancy(world_min, world_max, res, boxes)
    maxd = int(maxdist / res + 1)
    cap = np.float32(np.float32(maxd) * res)
    if not occ.any():
        dist = np.full(occ.size, cap, np.float32)
        obst = np.full((occ.size, 3), -1, np.int32)
        return dist, obst, dims, mk
    d, ind = ndimage.distance_transform_edt(~occ, return_indices=True)
    sq = np.rint(d * d).astype(np.int64)
    near = sq < maxd * maxd
    dist = np.where(near, (np.sqrt(sq.astype(np.float64)).astype(np.float32).astype(np.float64) * res), cap)
    obst = np.where(near[None], ind, -1).astype(np.int32)
    return (np.ascontiguousarray(dist.reshape(-1), np.float32),
            np.ascontiguousarray(np.moveaxis(obst, 0, -1).reshape(-1, 3)), dims, mk)
