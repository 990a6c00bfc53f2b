"""The oracle and the product's host-buildable arithmetic against OUTPUTS OF THE
REFERENCE'S OWN CODE (tests/golden/reference_runs.json).

The fixture is written by tests/golden/extract_reference_runs.py in the build
container: oracle/Makefile.ref compiles the reference's master-element sources
(Hex8 / Tet4 / Pyr5 / Wed6 / Quad4-2D CVFEM), PecletFunction.C and the van Leer
limiter of EdgeKernelUtils.h, unmodified and from where they lie, against
stand-in headers for Kokkos / STK (oracle/ref_shim), and the script runs them on
seeded inputs.  This pins what no in-tree gold of the reference pins:

  * GeometryInteriorAlg for Tet4 / Wed6 / Pyr5 (and Hex8 / Quad4 on warped
    elements): the oracle's dual nodal volumes and edge area vectors, and the
    product's element arithmetic (csrc/geometry_cvfem.h, CPU replay), against
    the reference's scv volumes / scs area vectors assembled as
    src/ngp_algorithms/GeometryInteriorAlg.C:99-106, 196-221 assembles them;
  * the integration-point tables (ipNodeMap, adjacentNodes);
  * the Peclet blending functions and the limiter over a sweep of arguments,
    bit for bit.

The device kernels (nw_geometry_interior_*) run on the same elements in
tests/test_zzz_reference_runs_gpu.py (`-m gpu`).
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import oracle_py as orc  # noqa: E402
import parity_util as pu  # noqa: E402

with open(os.path.join(HERE, "golden", "reference_runs.json")) as _f:
    R = json.load(_f)

FH = float.fromhex
TOPOS_3D = ["hex", "tet", "pyr", "wed"]
EMU_TOPO = {"tet": 0, "wed": 1, "pyr": 2}


def _unhex(a, shape):
    return np.array([FH(v) for v in a], dtype=np.float64).reshape(shape)


def _block(topo):
    """the fixture's elements of one topology as one element block without
    shared nodes; expected dual volumes / edge areas assembled from the
    reference's per-element outputs the way GeometryInteriorAlg does"""
    m = R["master_elements"][topo]
    nd, npe = m["ndim"], m["nodes_per_element"]
    nscv, nscs = m["num_scv_ip"], m["num_scs_ip"]
    ipn = m["ip_node_map"]
    lr = np.array(m["adjacent_nodes"]).reshape(nscs, 2)
    # the element's edges, oriented from the lower to the higher local node
    loc_edges = sorted({(min(a, b), max(a, b)) for a, b in lr})
    nel = len(m["elements"])
    coords = np.zeros((nel * npe, nd))
    conn = np.zeros((nel, npe), dtype=np.int32)
    edges = np.zeros((nel * len(loc_edges), 2), dtype=np.int32)
    dnv = np.zeros(nel * npe)
    ev = np.zeros(nel)
    area = np.zeros((len(edges), nd))
    for k, el in enumerate(m["elements"]):
        base = k * npe
        coords[base:base + npe] = _unhex(el["coords"], (npe, nd))
        conn[k] = base + np.arange(npe)
        vol = _unhex(el["scv_volume"], (nscv,))
        av = _unhex(el["scs_areav"], (nscs, nd))
        for ip in range(nscv):  # GeometryInteriorAlg.C:99-106
            dnv[base + ipn[ip]] += vol[ip]
            ev[k] += vol[ip]
        for j, (a, b) in enumerate(loc_edges):
            edges[k * len(loc_edges) + j] = (base + a, base + b)
        for ip in range(nscs):  # GeometryInteriorAlg.C:196-221
            a, b = lr[ip]
            j = loc_edges.index((min(a, b), max(a, b)))
            sign = 1.0 if a == loc_edges[j][0] else -1.0
            area[k * len(loc_edges) + j] += sign * av[ip]
    return m, coords, conn, edges, dnv, ev, area


def _oracle_geometry(topo, conn, coords, edges, n):
    if topo == "quad":
        return orc.geometry_interior_quad4(conn, coords, edges, n)
    return orc.geometry_interior_3d(topo, conn, coords, edges, n)


def _ulps(got, want):
    """largest difference in units of the last place of the wanted value"""
    got, want = np.asarray(got).ravel(), np.asarray(want).ravel()
    sp = np.spacing(np.maximum(np.abs(want), np.finfo(float).tiny))
    return float(np.max(np.abs(got - want) / sp))


def test_fixture_is_what_the_reference_build_gives_now():
    """where oracle/_ref exists (the build container), the committed fixture is
    exactly what the reference's code returns"""
    so = os.path.join(HERE, "..", "oracle", "_ref", "libnalu_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    L = C.CDLL(so)
    vp = C.c_void_p
    ids = {"hex": 0, "tet": 1, "pyr": 2, "wed": 3, "quad": 4}
    for topo, tid in ids.items():
        m = R["master_elements"][topo]
        nd, npe = m["ndim"], m["nodes_per_element"]
        for el in m["elements"]:
            x = _unhex(el["coords"], (npe, nd))
            vol = np.zeros(m["num_scv_ip"])
            av = np.zeros((m["num_scs_ip"], nd))
            assert L.ref_scv_volume(tid, vp(x.ctypes.data), 1, vp(vol.ctypes.data)) == 0
            assert L.ref_scs_areav(tid, vp(x.ctypes.data), 1, vp(av.ctypes.data)) == 0
            assert np.array_equal(vol, _unhex(el["scv_volume"], vol.shape))
            assert np.array_equal(av, _unhex(el["scs_areav"], av.shape))
    L.ref_peclet_tanh.restype = C.c_double
    L.ref_peclet_tanh.argtypes = [C.c_double] * 3
    for c1, c2, p, want in R["peclet_tanh"][:50]:
        assert L.ref_peclet_tanh(FH(c1), FH(c2), FH(p)) == FH(want)


@pytest.mark.parametrize("name", ["multiElemTypeCylinder", "hybrid_g_8_0"])
def test_oracle_geometry_vs_reference_on_its_own_meshes(name):
    """every element of the reference's mixed tet / pyramid / wedge / hex
    regression meshes through the reference's master elements, live (needs
    oracle/_ref: skipped where the reference tree never was)"""
    so = os.path.join(HERE, "..", "oracle", "_ref", "libnalu_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    L = C.CDLL(so)
    vp = C.c_void_p
    L.ref_geometry_block.argtypes = [C.c_int, C.c_long, vp, vp, vp, vp]
    ids = {"hex": 0, "tet": 1, "pyr": 2, "wed": 3}
    msh = pu.load_reference_mesh(name)
    coords = np.ascontiguousarray(msh["coords"])
    edges = np.ascontiguousarray(msh["edges"])
    n = len(coords)
    blocks = pu.mesh_blocks(msh)
    odnv, oarea, oev = pu.oracle_mesh_geometry(blocks, coords, edges)
    ekey = {}
    for e, (a, b) in enumerate(edges):
        ekey[(int(a), int(b))] = (e, 1.0)
        ekey[(int(b), int(a))] = (e, -1.0)
    dnv, area = np.zeros(n), np.zeros((len(edges), 3))
    dmag, amag = np.zeros(n), np.zeros(len(edges))
    total = 0
    for t, conn in blocks.items():
        m = R["master_elements"][t]
        nscv, nscs = m["num_scv_ip"], m["num_scs_ip"]
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        vol = np.zeros((len(conn), nscv))
        av = np.zeros((len(conn), nscs, 3))
        assert L.ref_geometry_block(ids[t], len(conn), conn.ctypes.data,
                                    coords.ctypes.data, vol.ctypes.data,
                                    av.ctypes.data) == 0
        # element volumes: the same sum in the same order, bit for bit
        ev = np.zeros(len(conn))
        for ip in range(nscv):
            ev += vol[:, ip]
        assert np.array_equal(ev, oev[t]), t
        ipn = np.array(m["ip_node_map"])
        np.add.at(dnv, conn[:, ipn].ravel(), vol.ravel())
        np.add.at(dmag, conn[:, ipn].ravel(), np.abs(vol).ravel())
        lr = np.array(m["adjacent_nodes"]).reshape(nscs, 2)
        for ip in range(nscs):
            L_, R_ = conn[:, lr[ip, 0]], conn[:, lr[ip, 1]]
            es = np.array([ekey[(int(a), int(b))] for a, b in zip(L_, R_)])
            idx, sgn = es[:, 0].astype(np.int64), es[:, 1]
            np.add.at(area, idx, sgn[:, None] * av[:, ip, :])
            np.add.at(amag, idx, np.abs(av[:, ip, :]).max(axis=1))
        total += len(conn)
    assert total > 20000 or name != "multiElemTypeCylinder"
    # sums over the elements around a node / an edge, accumulated in another
    # order than the oracle's: a few ulp of the sum of magnitudes
    assert np.max(np.abs(dnv - odnv) / dmag) <= 1e-15
    assert np.max(np.abs(area - oarea) / amag[:, None]) <= 1e-15


@pytest.mark.parametrize("topo", TOPOS_3D + ["quad"])
def test_oracle_geometry_vs_reference_master_elements(topo):
    m, coords, conn, edges, dnv, ev, area = _block(topo)
    n = len(coords)
    odnv, oev, oarea = _oracle_geometry(topo, conn, coords, edges, n)
    # same formulas in the same operation order, no contraction on either side
    # (-ffp-contract=off): the oracle reproduces the reference's bits
    assert np.array_equal(odnv, dnv), topo
    assert np.array_equal(oev, ev), topo
    assert np.array_equal(oarea, area), topo
    # reversed mesh-edge orientation flips the sign (GeometryInteriorAlg.C:213)
    _, _, orev = _oracle_geometry(topo, conn, coords, edges[:, ::-1].copy(), n)
    assert np.array_equal(orev, -oarea)


@pytest.mark.parametrize("topo", ["tet", "wed", "pyr"])
def test_product_element_arithmetic_vs_reference_master_elements(topo):
    """csrc/geometry_cvfem.h (the header the device kernels are built from),
    replayed on the CPU by tests/emul"""
    m, coords, conn, edges, dnv, ev, area = _block(topo)
    L = pu.emu_lib()
    vp = C.c_void_p
    L.emu_geometry_cvfem.argtypes = [C.c_int, C.c_int64, vp, vp, C.c_int64, vp, vp, vp]
    pd, pa = np.zeros(len(coords)), np.zeros((len(edges), 3))
    assert L.emu_geometry_cvfem(EMU_TOPO[topo], len(conn), conn.ctypes.data,
                                coords.ctypes.data, len(edges), edges.ctypes.data,
                                pd.ctypes.data, pa.ctypes.data) == 0
    vscale = np.repeat(ev, m["nodes_per_element"])
    assert np.max(np.abs(pd - dnv) / vscale) <= 1e-15, topo
    assert np.max(np.abs(pa - area)) <= 1e-15 * np.max(np.abs(area)), topo


def test_integration_point_tables_vs_reference():
    """ipNodeMap / adjacentNodes of the reference's master elements against what
    the oracle's geometry does with a one-hot probe: moving one node changes
    exactly the sub-control volumes / surfaces the tables attach to it"""
    for topo in TOPOS_3D:
        m = R["master_elements"][topo]
        npe = m["nodes_per_element"]
        assert sorted(m["ip_node_map"]) == list(range(npe)) or topo == "pyr"
        lr = np.array(m["adjacent_nodes"]).reshape(-1, 2)
        assert np.all(lr[:, 0] != lr[:, 1]) and lr.min() == 0 and lr.max() == npe - 1
        # every element edge of the topology appears as an (L, R) pair
        want = {"hex": 12, "tet": 6, "pyr": 8, "wed": 9}[topo]
        assert len({(min(a, b), max(a, b)) for a, b in lr}) == want


def test_peclet_functions_bitwise_vs_reference():
    L = pu.emu_lib()
    L.emu_peclet_eval.restype = C.c_double
    L.emu_peclet_eval.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    n = 0
    for A, hf, p, want in R["peclet_classic"]:
        hf, p, want = FH(hf), FH(p), FH(want)
        assert orc.peclet_eval(orc.peclet("classic", hf), p) == want, (hf, p)
        assert L.emu_peclet_eval(0, hf, 1.0, p) == want, (hf, p)
        n += 1
    worst = 0.0
    for c1, c2, p, want in R["peclet_tanh"]:
        c1, c2, p, want = FH(c1), FH(c2), FH(p), FH(want)
        # same libm, same expression: the oracle reproduces the bits
        assert orc.peclet_eval(orc.peclet("tanh", c1, c2), p) == want, (c1, c2, p)
        # the product forms (p - c1) / c2 with its reciprocal-multiply: <= 2 ulp
        # of the argument, which tanh' <= 1 carries into the result
        got = L.emu_peclet_eval(1, c1, c2, p)
        worst = max(worst, abs(got - want))
        n += 1
    assert worst <= 4e-16, worst
    assert n == len(R["peclet_classic"]) + len(R["peclet_tanh"]) and n > 400


def test_van_leer_bitwise_vs_reference():
    Lo = orc.lib()
    Lo.orc_van_leer.restype = C.c_double
    Lo.orc_van_leer.argtypes = [C.c_double] * 3
    Le = pu.emu_lib()
    Le.emu_van_leer.restype = C.c_double
    Le.emu_van_leer.argtypes = [C.c_double] * 3
    worst = 0.0
    for a, b, eps, want in R["van_leer"]:
        a, b, eps, want = FH(a), FH(b), FH(eps), FH(want)
        got = Lo.orc_van_leer(a, b, eps)
        assert got == want or (np.isnan(got) and np.isnan(want)), (a, b, eps)
        gp = Le.emu_van_leer(a, b, eps)
        if np.isnan(want):
            assert np.isnan(gp)
        else:
            # the product multiplies by a reciprocal where the reference divides
            worst = max(worst, abs(gp - want) / max(abs(want), 1e-300))
    assert worst <= 4.5e-16, worst
