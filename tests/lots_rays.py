"""Ray sets of MOOSE's LotsOfRaysRayStudy regression (reference:
moose/modules/ray_tracing/test/src/userobjects/LotsOfRaysRayStudy.C:126-232) rebuilt on our own
GeneratedMesh-numbered grids: vertex-to-vertex, centroid-to-vertex, centroid-to-centroid and
edge-to-edge rays started on every boundary side."""
import numpy as np

from macaw_b200 import meshgen


def boundary_sides(mesh):
    e, s = np.nonzero(mesh.neighbors < 0)
    return list(zip(e.tolist(), s.tolist()))


def lots(mesh, kind):
    sn = meshgen.SIDE_NODES[mesh.elem_type]
    nv = mesh.conn.shape[1]
    P = mesh.xyz[mesh.conn]  # (n_elem, nv, 3)
    starts, dirs, elems, sides, maxd = [], [], [], [], []
    bsides = boundary_sides(mesh)
    if kind == "ctc":
        on_bnd = np.unique([e for e, _ in bsides])
        bcent = P[on_bnd].mean(axis=1)
    for e, s in bsides:
        side_nodes = sn[s]
        sc = P[e, side_nodes].mean(axis=0)
        others = [v for v in range(nv) if v not in side_nodes]
        if kind == "vtv":
            for n in side_nodes:
                for v in others:
                    starts.append(P[e, n]); dirs.append(P[e, v] - P[e, n]); elems.append(e); sides.append(s)
                    maxd.append(np.inf)
        elif kind == "ctv":
            for v in others:
                starts.append(sc); dirs.append(P[e, v] - sc); elems.append(e); sides.append(s); maxd.append(np.inf)
        elif kind == "ctc":
            for c in bcent:
                d = c - sc
                starts.append(sc); dirs.append(d); elems.append(e); sides.append(-1); maxd.append(np.linalg.norm(d))
        elif kind == "ete":
            en = meshgen.EDGE_NODES[mesh.elem_type]
            ecent = P[e][en].mean(axis=1)
            for k, (a, b) in enumerate(en):
                if a in side_nodes and b in side_nodes:
                    for k2 in range(len(en)):
                        if k2 != k:
                            starts.append(ecent[k]); dirs.append(ecent[k2] - ecent[k]); elems.append(e)
                            sides.append(s); maxd.append(np.inf)
    return (np.array(starts), np.array(dirs), np.array(elems, np.int32), np.array(sides, np.int32),
            np.array(maxd))


def box_exit_distance(start, dirs, lo, hi, dim):
    """Analytic distance from `start` along `dirs` to the bounding box (what the reference's
    BoundingBoxIntersectionHelper computes for expected_distance)."""
    d = dirs / np.linalg.norm(dirs, axis=1, keepdims=True)
    t = np.full(len(start), np.inf)
    for a in range(dim):
        with np.errstate(divide="ignore", invalid="ignore"):
            ta = np.where(d[:, a] > 1e-14, (hi[a] - start[:, a]) / d[:, a],
                          np.where(d[:, a] < -1e-14, (lo[a] - start[:, a]) / d[:, a], np.inf))
        t = np.minimum(t, ta)
    return t
