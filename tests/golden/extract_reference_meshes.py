#!/usr/bin/env python
"""Derive small mesh fixtures from the reference's own regression meshes
(reg_tests/mesh, NetCDF-3 Exodus files readable with scipy) and write them to
tests/golden/mesh_*.npz: node coordinates, global node ids, the edge list
stk::mesh::create_edges would make (unique element edges, each ordered by
ascending global id, first-visit order) and the element connectivity per
topology (elems_tet / elems_pyr / elems_wed / elems_hex / elems_qua).  Run in the build container only (the GPU box has no
/root/reference):  python tests/golden/extract_reference_meshes.py

  multiElemTypeCylinder.g   TETRA4 / HEX8 / WEDGE6 / PYRAMID5   (BASELINE configs[4])
  hybrid.g.8.0              tetra / pyramid / hex, rank 0 of 8   (BASELINE configs[4])
  hybrid.g.8.0 .. .7        all eight parts of the reference's own decomposition,
                            with node / edge ownership (multi-rank parity)
  airfoilRANSEdgeTrilinos.rst  2-D QUAD4, 49 536 nodes            (BASELINE configs[3])
"""
import os
import sys

import numpy as np
from scipy.io import netcdf_file

REF = os.environ.get("NALU_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

# stk::topology edge node ordinals
EDGES = {
    "tet": [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)],
    "pyr": [(0, 1), (1, 2), (2, 3), (3, 0), (0, 4), (1, 4), (2, 4), (3, 4)],
    "wed": [(0, 1), (1, 2), (2, 0), (3, 4), (4, 5), (5, 3), (0, 3), (1, 4), (2, 5)],
    "hex": [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4),
            (0, 4), (1, 5), (2, 6), (3, 7)],
    "qua": [(0, 1), (1, 2), (2, 3), (3, 0)],
}


def convert(fname, out, keep_elems=False):
    f = netcdf_file(os.path.join(REF, "reg_tests", "mesh", fname), "r", mmap=False)
    ndim = int(f.dimensions["num_dim"])
    coords = np.stack([np.array(f.variables["coord" + "xyz"[d]][:], dtype=np.float64)
                       for d in range(ndim)], axis=1)
    gid = np.array(f.variables["node_num_map"][:], dtype=np.int64)
    pairs, blocks, topo = [], {}, {}
    for b in range(1, int(f.dimensions["num_el_blk"]) + 1):
        v = f.variables["connect%d" % b]
        conn = np.array(v[:], dtype=np.int64) - 1
        kind = v.elem_type.decode().lower()[:3]
        topo[kind] = topo.get(kind, 0) + len(conn)
        for a, c in EDGES[kind]:
            pairs.append(np.stack([conn[:, a], conn[:, c]], axis=1))
        if keep_elems:
            blocks.setdefault(kind, []).append(conn)
    p = np.concatenate(pairs)
    lo, hi = np.minimum(p[:, 0], p[:, 1]), np.maximum(p[:, 0], p[:, 1])
    key = lo * len(coords) + hi
    _, first = np.unique(key, return_index=True)
    first.sort()  # first-visit order
    e = np.stack([lo[first], hi[first]], axis=1)
    swap = gid[e[:, 0]] > gid[e[:, 1]]  # L = lower global id
    e[swap] = e[swap][:, ::-1]
    data = dict(coords=coords, gid=gid, edges=e.astype(np.int32))
    for kind, lst in blocks.items():
        data["elems_" + kind] = np.concatenate(lst).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, out), **data)
    print(out, "nodes", len(coords), "edges", len(e), topo,
          "%.0f kB" % (os.path.getsize(os.path.join(HERE, out)) / 1e3))


def read_part(fname):
    """(coords, gid, local unique edges in first-visit order with L = lower
    global id) of one Exodus part file"""
    f = netcdf_file(os.path.join(REF, "reg_tests", "mesh", fname), "r", mmap=False)
    ndim = int(f.dimensions["num_dim"])
    coords = np.stack([np.array(f.variables["coord" + "xyz"[d]][:], dtype=np.float64)
                       for d in range(ndim)], axis=1)
    gid = np.array(f.variables["node_num_map"][:], dtype=np.int64)
    pairs = []
    for b in range(1, int(f.dimensions["num_el_blk"]) + 1):
        if "connect%d" % b not in f.variables:
            continue  # this part holds no element of the block
        v = f.variables["connect%d" % b]
        conn = np.array(v[:], dtype=np.int64) - 1
        kind = v.elem_type.decode().lower()[:3]
        for a, c in EDGES[kind]:
            pairs.append(np.stack([conn[:, a], conn[:, c]], axis=1))
    p = np.concatenate(pairs)
    lo, hi = np.minimum(p[:, 0], p[:, 1]), np.maximum(p[:, 0], p[:, 1])
    _, first = np.unique(lo * len(coords) + hi, return_index=True)
    first.sort()
    e = np.stack([lo[first], hi[first]], axis=1)
    swap = gid[e[:, 0]] > gid[e[:, 1]]
    e[swap] = e[swap][:, ::-1]
    return coords, gid, e.astype(np.int32)


def convert_decomposition(stem, nparts, out):
    """All parts of a pre-split Exodus mesh (the decomposition the reference's
    regression test runs on, reg_tests/mesh/<stem>.<nparts>.<rank>): per part
    the node coordinates, global ids, local edges, and -- recorded explicitly
    -- the STK ownership of shared nodes and edges: the lowest rank that holds
    the entity (every element lives in exactly one part; a node / edge on a
    partition interface appears in several)."""
    parts = [read_part("%s.%d.%d" % (stem, nparts, r)) for r in range(nparts)]
    node_owner, edge_owner = {}, {}
    for r, (coords, gid, e) in enumerate(parts):
        for g in gid.tolist():
            node_owner.setdefault(g, r)
        for a, c in zip(gid[e[:, 0]].tolist(), gid[e[:, 1]].tolist()):
            edge_owner.setdefault((a, c), r)
    data = {"nparts": np.int64(nparts)}
    for r, (coords, gid, e) in enumerate(parts):
        data["coords_%d" % r] = coords
        data["gid_%d" % r] = gid
        data["edges_%d" % r] = e
        data["node_owner_%d" % r] = np.array(
            [node_owner[g] for g in gid.tolist()], dtype=np.int32)
        data["edge_owner_%d" % r] = np.array(
            [edge_owner[(a, c)] for a, c in zip(gid[e[:, 0]].tolist(),
                                                gid[e[:, 1]].tolist())], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, out), **data)
    nn = len(node_owner)
    print(out, "parts", nparts, "global nodes", nn, "global edges", len(edge_owner),
          "shared node copies", sum(len(p[1]) for p in parts) - nn,
          "%.0f kB" % (os.path.getsize(os.path.join(HERE, out)) / 1e3))


def main():
    convert_decomposition("hybrid.g", 8, "mesh_hybrid_g_8_parts.npz")
    convert("multiElemTypeCylinder.g", "mesh_multiElemTypeCylinder.npz", keep_elems=True)
    convert("hybrid.g.8.0", "mesh_hybrid_g_8_0.npz", keep_elems=True)
    convert("airfoilRANSEdgeTrilinos.rst", "mesh_airfoilRANSEdge.npz", keep_elems=True)


if __name__ == "__main__":
    sys.exit(main())
