#!/usr/bin/env python
"""Run the REFERENCE'S OWN code on seeded inputs and write the outputs to
tests/golden/reference_runs.json.

What runs: oracle/_ref/libnalu_ref.so, built by oracle/Makefile.ref from the
reference's source files where they lie (src/master_element/{Hex8,Tet4,Pyr5,
Wed6,Quad42D,Tri32D}CVFEM.C, MasterElement.C, src/PecletFunction.C,
include/edge_kernels/EdgeKernelUtils.h), unmodified, against the stand-in
headers of oracle/ref_shim/ for the Kokkos / STK / MPI names they touch.

Run in the build container (needs /root/reference, which does not exist on the
GPU box):  python tests/golden/extract_reference_runs.py

Output (all doubles as C99 hex strings, so the file holds the exact bits):
  master_elements[topo]: sizes, ipNodeMap, adjacentNodes, and per element the
      nodal coordinates with the reference's scv volumes and scs area vectors
      (double overload; the DoubleType overload GeometryInteriorAlg calls is
      checked to give the same bits before anything is written)
  peclet_classic / peclet_tanh / van_leer: argument tuples with the
      reference's return values
and tests/golden/reference_edge_runs.npz: a 2x2x2-element warped two-phase case
(27 nodes, 54 edges) with all inputs and, for the option matrices of
tests/test_reference_edge_runs.py, the local lhs / rhs blocks of every edge as
the reference's own MomentumEdgeSolverAlg / ScalarEdgeSolverAlg /
ContinuityEdgeSolverAlg leave them, and MdotEdgeAlg's mass_flow_rate.
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("NALU_REFERENCE", "/root/reference")

TOPO_ID = {"hex": 0, "tet": 1, "pyr": 2, "wed": 3, "quad": 4, "tri": 5}
# parent-element node coordinates in the Exodus / STK node order
PARENT = {
    "hex": [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0],
            [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]],
    "tet": [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]],
    "pyr": [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1]],
    "wed": [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]],
    "quad": [[0, 0], [1, 0], [1, 1], [0, 1]],
    "tri": [[0, 0], [1, 0], [0, 1]],
}


def hx(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def load():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s",
                           "-f", "Makefile.ref", "REF=" + REF])
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libnalu_ref.so"))
    d = C.c_double
    for n in ("ref_peclet_classic", "ref_peclet_tanh", "ref_van_leer"):
        getattr(L, n).restype = d
        getattr(L, n).argtypes = [d, d, d]
    return L


def elements(topo, rng, n):
    """the parent element, affine images of it, and warped (non-affine) ones at
    an offset from the origin (the Grandy hex volume is not translation
    invariant in rounding)"""
    x0 = np.array(PARENT[topo], dtype=np.float64)
    nd = x0.shape[1]
    out = [x0.copy()]
    for k in range(n - 1):
        while True:
            A = np.eye(nd) + 0.35 * rng.standard_normal((nd, nd))
            if np.linalg.det(A) > 0.4:
                break
        x = x0 @ A.T + 3.0 * rng.standard_normal(nd)
        if k % 2:  # bend the faces
            x = x + 0.08 * rng.standard_normal(x.shape)
        out.append(x)
    return out


def main():
    L = load()
    rng = np.random.default_rng(20261018)
    vp = C.c_void_p
    me = {}
    for topo, tid in TOPO_ID.items():
        sz = (C.c_int * 4)()
        assert L.ref_me_sizes(tid, sz) == 0
        nd, npe, nscv, nscs = list(sz)
        ipn = (C.c_int * nscv)()
        lr = (C.c_int * (2 * nscs))()
        assert L.ref_me_maps(tid, ipn, lr) == 0
        els = []
        for x in elements(topo, rng, 8):
            x = np.ascontiguousarray(x)
            vol, vol_s = np.zeros(nscv), np.zeros(nscv)
            av, av_s = np.zeros((nscs, nd)), np.zeros((nscs, nd))
            assert L.ref_scv_volume(tid, vp(x.ctypes.data), 0, vp(vol.ctypes.data)) == 0
            assert L.ref_scv_volume(tid, vp(x.ctypes.data), 1, vp(vol_s.ctypes.data)) == 0
            assert L.ref_scs_areav(tid, vp(x.ctypes.data), 0, vp(av.ctypes.data)) == 0
            assert L.ref_scs_areav(tid, vp(x.ctypes.data), 1, vp(av_s.ctypes.data)) == 0
            # the two overloads of the reference agree to the bit
            assert np.array_equal(vol, vol_s) and np.array_equal(av, av_s), topo
            assert np.all(vol > 0.0), (topo, vol)
            els.append({"coords": hx(x), "scv_volume": hx(vol), "scs_areav": hx(av)})
        me[topo] = {"ndim": nd, "nodes_per_element": npe, "num_scv_ip": nscv,
                    "num_scs_ip": nscs, "ip_node_map": list(ipn),
                    "adjacent_nodes": list(lr), "elements": els}

    pec = np.concatenate([[0.0, 1e-300, 1e-16, 1e-8, 0.5, 1.0, 2.0, 5.0, 1e3, 1e8],
                          10.0 ** rng.uniform(-6, 4, 40)])
    classic = []
    for A, hf in ((5.0, 1.0), (5.0, 0.0), (5.0, 0.37), (2.5, 1.0)):
        for p in pec:
            classic.append([float(A).hex(), float(hf).hex(), float(p).hex(),
                            L.ref_peclet_classic(A, hf, p).hex()])
    tanh = []
    for c1, c2 in ((2.0, 1.0), (5000.0, 200.0), (0.0, 0.25), (10.0, 3.0)):
        for p in np.concatenate([pec, c1 + c2 * rng.uniform(-4, 4, 20)]):
            tanh.append([float(c1).hex(), float(c2).hex(), float(p).hex(),
                         L.ref_peclet_tanh(c1, c2, p).hex()])
    vl = []
    vals = np.concatenate([[0.0, 1.0, -1.0, 1e-20, -1e-20, 3.5],
                           rng.standard_normal(30) * 10.0 ** rng.uniform(-8, 3, 30)])
    for a in vals:
        for b in vals[::3]:
            for eps in (1e-16, 0.0 if (a + b) != 0.0 else 1e-16):
                vl.append([float(a).hex(), float(b).hex(), float(eps).hex(),
                           L.ref_van_leer(a, b, eps).hex()])
    out = {
        "_source": "outputs of the reference's own code run in the build container: "
                   "oracle/_ref/libnalu_ref.so (oracle/Makefile.ref, "
                   "oracle/ref_driver.cpp); regenerate with "
                   "tests/golden/extract_reference_runs.py",
        "_reference_files": [
            "src/master_element/Hex8CVFEM.C", "src/master_element/Tet4CVFEM.C",
            "src/master_element/Pyr5CVFEM.C", "src/master_element/Wed6CVFEM.C",
            "src/master_element/Quad42DCVFEM.C", "src/master_element/Tri32DCVFEM.C",
            "src/master_element/MasterElement.C", "src/PecletFunction.C",
            "include/edge_kernels/EdgeKernelUtils.h"],
        "master_elements": me, "peclet_classic": classic, "peclet_tanh": tanh,
        "van_leer": vl,
    }
    path = os.path.join(HERE, "reference_runs.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes;",
          {k: len(v["elements"]) for k, v in me.items()},
          len(classic), len(tanh), len(vl))

    # the reference's own edge algorithms (oracle/ref_edge_driver.cpp) on a small
    # two-phase case: inputs and the local blocks of every edge
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import test_reference_edge_runs as T
    epath = os.path.join(HERE, "reference_edge_runs.npz")
    print("wrote", epath, T.write_fixture(epath), os.path.getsize(epath), "bytes")


if __name__ == "__main__":
    sys.exit(main())
