"""GeometryInteriorAlg for Tet4 / Wed6 / Pyr5 blocks on the device
(nw_geometry_interior_tet4 / _wed6 / _pyr5, csrc/nw_geometry.cu) against the
oracle, on the reference's own mixed-element regression meshes (BASELINE
configs[4]) with their true CVFEM dual geometry.  Needs a B200: `pytest -m gpu`.

Own module, collected after tests/test_gpu_parity.py.  The per-element
arithmetic (csrc/geometry_cvfem.h) is also replayed on the CPU against the
oracle by tests/test_geometry_topologies_cpu.py."""
import numpy as np
import pytest

import oracle_py as orc
import parity_util as pu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    return pu.pkg()


@pytest.fixture(scope="module")
def ctx(P):
    c = P.Context(0)
    yield c
    c.close()


def _device_geometry(P, mesh, blocks, reps=1):
    mesh.register("dual_nodal_volume", P.NW_NODE, 1)
    mesh.register("edge_area_vector", P.NW_EDGE, 3)
    for _ in range(reps):  # later calls: cached per-topology tables
        mesh.fill("dual_nodal_volume", 0.0)
        mesh.fill("edge_area_vector", 0.0)
        for conn in blocks.values():  # one GeometryInteriorAlg per topology
            mesh.geometry_interior(conn, dnv="dual_nodal_volume",
                                   area="edge_area_vector")
    return (mesh.download("dual_nodal_volume"),
            mesh.download("edge_area_vector").reshape(-1, 3))


@pytest.mark.parametrize("name,tile", [("multiElemTypeCylinder", 0),
                                       ("hybrid_g_8_0", 64)])
def test_mixed_mesh_geometry_vs_oracle(P, ctx, name, tile):
    m = pu.load_reference_mesh(name)
    coords = np.ascontiguousarray(m["coords"])
    edges = np.ascontiguousarray(m["edges"])
    blocks = pu.mesh_blocks(m)
    n = len(coords)
    odnv, oarea, _ = pu.oracle_mesh_geometry(blocks, coords, edges)
    mesh = P.Mesh(ctx, 3, edges, np.arange(n, dtype=np.int64), coords,
                  tile_nodes=tile)
    dnv, area = _device_geometry(P, mesh, blocks, reps=2)
    # fp64 atomics add the element shares in any order, FMA contraction moves
    # single ulps: 1e-12 of the largest share
    assert np.max(np.abs(dnv - odnv)) <= 1e-12 * np.max(odnv)
    assert np.max(np.abs(area - oarea)) <= 1e-12 * np.max(np.abs(oarea))
    # closed dual cells at interior nodes, from the device's own numbers
    acc, mag = np.zeros((n, 3)), np.zeros(n)
    np.add.at(acc, edges[:, 0], area)
    np.add.at(acc, edges[:, 1], -area)
    am = np.linalg.norm(area, axis=1)
    np.add.at(mag, edges[:, 0], am)
    np.add.at(mag, edges[:, 1], am)
    inner = ~pu.boundary_nodes(blocks, n)
    assert np.max(np.linalg.norm(acc[inner], axis=1) / mag[inner]) <= 1e-12
    # NodalGradEdgeAlg on the device-made geometry: exact for a linear field
    # where only tetrahedra meet
    only_tet = np.ones(n, dtype=bool)
    for t, conn in blocks.items():
        if t != "tet":
            only_tet[conn.ravel()] = False
    g = np.array([0.7, -1.3, 0.45])
    mesh.put("phi", P.NW_NODE, coords @ g + 2.0)
    mesh.register("dphidx", P.NW_NODE, 3)
    mesh.nodal_grad_edge("phi", "dphidx")
    grad = mesh.download("dphidx").reshape(-1, 3)
    sel = inner & only_tet
    assert sel.sum() > 100
    assert np.max(np.abs(grad[sel] - g)) <= 1e-10 * np.linalg.norm(g)
    ograd = orc.nodal_grad_edge(1, 3, edges, coords @ g + 2.0, oarea, odnv, n)
    assert np.max(np.abs(grad - ograd)) <= 1e-10 * np.max(np.abs(ograd))
    mesh.close()


@pytest.mark.parametrize("topo", ["tet", "wed", "pyr"])
def test_single_block_owned_flags_and_volume_only(P, ctx, topo):
    """one block at a time: the locally-owned selector gates the volumes only,
    either output may be left out"""
    m = pu.load_reference_mesh("multiElemTypeCylinder")
    coords = np.ascontiguousarray(m["coords"])
    edges = np.ascontiguousarray(m["edges"])
    conn = pu.mesh_blocks(m)[topo]
    n = len(coords)
    owned = (np.arange(len(conn)) % 3 != 0).astype(np.uint8)
    odnv, _, oarea = orc.geometry_interior_3d(topo, conn, coords, edges, n,
                                              elem_owned=owned)
    mesh = P.Mesh(ctx, 3, edges, np.arange(n, dtype=np.int64), coords)
    mesh.register("dual_nodal_volume", P.NW_NODE, 1)
    mesh.register("edge_area_vector", P.NW_EDGE, 3)
    mesh.fill("dual_nodal_volume", 0.0)
    mesh.fill("edge_area_vector", 0.0)
    mesh.geometry_interior(conn, dnv="dual_nodal_volume", elem_owned=owned)
    mesh.geometry_interior(conn, area="edge_area_vector", elem_owned=owned)
    dnv = mesh.download("dual_nodal_volume")
    area = mesh.download("edge_area_vector").reshape(-1, 3)
    assert np.max(np.abs(dnv - odnv)) <= 1e-12 * np.max(odnv)
    assert np.max(np.abs(area - oarea)) <= 1e-12 * np.max(np.abs(oarea))
    mesh.close()
