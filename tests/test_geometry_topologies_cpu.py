"""GeometryInteriorAlg for Tet4 / Wed6 / Pyr5: the oracle's
restatement and the product's per-element arithmetic (csrc/geometry_cvfem.h,
replayed on the CPU by tests/emul), pinned by properties -- the reference holds
no known-answer test for these three topologies:

  * affine images of the parent elements: every sub-control volume scales with
    the determinant, their sum is the element volume;
  * the reference's own mixed regression meshes: element volumes against an
    independent face-based formula, closed dual cells at interior nodes (across
    tet / pyramid / wedge / hex interfaces), linear-exact Green-Gauss gradient
    in the all-tetrahedra regions."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import oracle_py as orc  # noqa: E402
import parity_util as pu  # noqa: E402

I32 = lambda a: np.array(a, dtype=np.int32)  # noqa: E731
PARENT = {
    "tet": (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]]),
            [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)], 1.0 / 6.0),
    "wed": (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0], [1, 0, 1], [0, 1, 1]]),
            [(0, 1), (1, 2), (0, 2), (3, 4), (4, 5), (3, 5), (0, 3), (1, 4), (2, 5)], 0.5),
    "pyr": (np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1.0]]),
            [(0, 1), (1, 2), (2, 3), (0, 3), (0, 4), (1, 4), (2, 4), (3, 4)], 1.0 / 3.0),
}
EMU_TOPO = {"tet": 0, "wed": 1, "pyr": 2}


def emu_geometry(topo, conn, coords, edges):
    """the product's element arithmetic on the CPU (tests/emul)"""
    L = pu.emu_lib()
    vp = C.c_void_p
    L.emu_geometry_cvfem.argtypes = [C.c_int, C.c_int64, vp, vp, C.c_int64, vp, vp, vp]
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    edges = np.ascontiguousarray(edges, dtype=np.int32)
    dnv, area = np.zeros(len(coords)), np.zeros((len(edges), 3))
    assert L.emu_geometry_cvfem(EMU_TOPO[topo], len(conn), conn.ctypes.data,
                                coords.ctypes.data, len(edges), edges.ctypes.data,
                                dnv.ctypes.data, area.ctypes.data) == 0
    return dnv, area


@pytest.mark.parametrize("topo", ["tet", "wed", "pyr"])
def test_parent_and_affine_elements(topo):
    x0, edges, v0 = PARENT[topo]
    npe = len(x0)
    conn = I32([list(range(npe))])
    dnv0, ev0, area0 = orc.geometry_interior_3d(topo, conn, x0, I32(edges), npe)
    assert abs(ev0[0] - v0) <= 2e-16
    if topo != "pyr":
        # the control volumes of a simplex / prism split it evenly
        assert np.max(np.abs(dnv0 - v0 / npe)) <= 2e-16
    else:
        assert np.max(np.abs(dnv0[:4] - dnv0[0])) <= 2e-16 and dnv0[4] > dnv0[0]
    # every sub-control surface points from its left to its right node
    dx0 = x0[I32(edges)[:, 1]] - x0[I32(edges)[:, 0]]
    assert np.all(np.einsum("ij,ij->i", area0, dx0) > 0.0)
    rng = np.random.default_rng(7 + npe)
    for _ in range(5):
        A = np.eye(3) + 0.4 * rng.standard_normal((3, 3))
        if np.linalg.det(A) < 0:
            A[:, 0] = -A[:, 0]
        det = np.linalg.det(A)
        x = x0 @ A.T + rng.standard_normal(3)
        dnv, ev, area = orc.geometry_interior_3d(topo, conn, x, I32(edges), npe)
        assert np.max(np.abs(dnv - det * dnv0)) <= 1e-13 * det
        # area vectors transform with the cofactor matrix
        cof = det * np.linalg.inv(A).T
        assert np.max(np.abs(area - area0 @ cof.T)) <= 1e-13 * np.max(np.abs(area))
        # product arithmetic == oracle
        pd, pa = emu_geometry(topo, conn, x, I32(edges))
        assert np.max(np.abs(pd - dnv)) <= 1e-15 * det
        assert np.max(np.abs(pa - area)) <= 1e-15 * np.max(np.abs(area))
        # reversed edge orientation flips the sign
        rev = I32(edges)[:, ::-1]
        _, _, arev = orc.geometry_interior_3d(topo, conn, x, rev, npe)
        assert np.array_equal(arev, -area)


@pytest.mark.parametrize("name", ["multiElemTypeCylinder", "hybrid_g_8_0"])
def test_reference_mixed_mesh_dual_geometry(name):
    m = pu.load_reference_mesh(name)
    coords = np.ascontiguousarray(m["coords"])
    edges = np.ascontiguousarray(m["edges"])
    blocks = pu.mesh_blocks(m)
    assert set(blocks) >= {"tet", "pyr", "hex"}
    n = len(coords)
    dnv, area, ev = pu.oracle_mesh_geometry(blocks, coords, edges)
    # (1) element volumes against the independent face-based formula (the
    # quadrilateral faces of these meshes are planar to rounding; what is left
    # is the rounding of Grandy's formula in absolute coordinates, eps |x|^3 / V)
    for t, conn in blocks.items():
        ref = pu.element_volumes_by_faces(t, conn, coords)
        assert np.all(ev[t] > 0.0)
        tol = 1e-9 if t == "tet" else 2e-3
        assert np.max(np.abs(ev[t] - ref) / ref) <= tol, t
    tot = sum(v.sum() for v in ev.values())
    assert abs(dnv.sum() - tot) <= 1e-12 * tot and np.all(dnv > 0.0)
    # (2) every element edge is a mesh edge: all sub-control surfaces landed
    # (3) dual cells of interior nodes are closed, whatever topologies meet there
    acc = np.zeros((n, 3))
    mag = np.zeros(n)
    np.add.at(acc, edges[:, 0], area)
    np.add.at(acc, edges[:, 1], -area)
    amag = np.linalg.norm(area, axis=1)
    np.add.at(mag, edges[:, 0], amag)
    np.add.at(mag, edges[:, 1], amag)
    inner = ~pu.boundary_nodes(blocks, n)
    assert inner.sum() > n // 3
    assert np.max(np.linalg.norm(acc[inner], axis=1) / mag[inner]) <= 1e-13
    # a boundary node's cell is open
    assert np.min(np.linalg.norm(acc[~inner], axis=1) / mag[~inner]) > 1e-3
    # (4) edge-based Green-Gauss gradient (NodalGradEdgeAlg) of a linear field is
    # exact where only tetrahedra meet (median dual of a simplicial mesh)
    only_tet = np.ones(n, dtype=bool)
    for t, conn in blocks.items():
        if t != "tet":
            only_tet[conn.ravel()] = False
    sel = inner & only_tet
    assert sel.sum() > 100
    g = np.array([0.7, -1.3, 0.45])
    phi = coords @ g + 2.0
    grad = orc.nodal_grad_edge(1, 3, edges, phi, area, dnv, n)
    scale = np.linalg.norm(g)
    assert np.max(np.abs(grad[sel] - g)) <= 1e-10 * scale
    # (5) the product's element arithmetic reproduces the oracle on every block
    pd, pa = np.zeros(n), np.zeros((len(edges), 3))
    od, oa = np.zeros(n), np.zeros((len(edges), 3))
    for t in ("tet", "wed", "pyr"):
        if t not in blocks:
            continue
        d, a = emu_geometry(t, blocks[t], coords, edges)
        pd += d
        pa += a
        d, _, a = orc.geometry_interior_3d(t, blocks[t], coords, edges, n)
        od += d
        oa += a
    assert np.max(np.abs(pd - od)) <= 1e-14 * np.max(od)
    assert np.max(np.abs(pa - oa)) <= 1e-14 * np.max(np.abs(oa))


@pytest.mark.parametrize("topo", ["tet", "wed", "pyr"])
def test_warped_elements_product_arithmetic_vs_oracle(topo):
    """200 randomly warped (non-affine: bent quadrilateral faces) elements in
    one block, sharing no nodes: product header == oracle to rounding, volumes
    positive and close to the planar-fan estimate, area vectors along their edges"""
    x0, edges0, _ = PARENT[topo]
    npe, ne = len(x0), len(edges0)
    rng = np.random.default_rng(20261017 + npe)
    nel = 200
    coords = np.concatenate([
        (x0 + 0.12 * rng.standard_normal(x0.shape)) * rng.uniform(0.5, 3.0)
        + 10.0 * rng.standard_normal(3) for _ in range(nel)])
    conn = I32(np.arange(nel * npe).reshape(nel, npe))
    edges = I32(np.concatenate([np.array(edges0) + npe * k for k in range(nel)]))
    # random edge orientation, as global-id ordering gives on a real mesh
    flip = rng.random(len(edges)) < 0.5
    edges[flip] = edges[flip][:, ::-1]
    n = len(coords)
    dnv, ev, area = orc.geometry_interior_3d(topo, conn, coords, edges, n)
    pdv, parea = emu_geometry(topo, conn, coords, edges)
    assert np.max(np.abs(pdv - dnv)) <= 1e-13 * np.max(dnv)
    assert np.max(np.abs(parea - area)) <= 1e-13 * np.max(np.abs(area))
    assert np.all(dnv > 0.0)
    ref = pu.element_volumes_by_faces(topo, conn, coords)
    assert np.max(np.abs(ev - ref) / ref) <= (1e-11 if topo == "tet" else 0.05)
    # orientation: an edge's area vector points from its first to its second node
    dx = coords[edges[:, 1]] - coords[edges[:, 0]]
    assert np.all(np.einsum("ij,ij->i", area, dx) > 0.0)
    assert len(edges) == nel * ne
