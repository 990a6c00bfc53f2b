"""The reference's own regression meshes (BASELINE configs[3], [4]) through the
product's host logic and the CPU walk-through of the tile kernels: CSR graph and
edge->slot map bit-exact with the oracle, assembled values within 1e-12.
Fixtures: tests/golden/mesh_*.npz (tests/golden/extract_reference_meshes.py).
No GPU."""
import numpy as np
import pytest

import oracle_py as orc
import parity_util as pu

MESHES_3D = ["multiElemTypeCylinder", "hybrid_g_8_0"]


def test_fixture_edge_counts():
    """node / edge counts of the derived fixtures (Euler-type sanity: every
    element edge once, L = lower global id)"""
    for name, nn, ne in (("multiElemTypeCylinder", 11978, 50547),
                         ("hybrid_g_8_0", 9755, 35965),
                         ("airfoilRANSEdge", 49536, 98688)):
        m = pu.load_reference_mesh(name)
        e, gid = m["edges"], m["gid"]
        assert (len(m["coords"]), len(e)) == (nn, ne)
        assert np.all(gid[e[:, 0]] < gid[e[:, 1]])
        key = np.minimum(e[:, 0], e[:, 1]).astype(np.int64) * nn + np.maximum(e[:, 0], e[:, 1])
        assert len(np.unique(key)) == ne
    # 2-D structured quad mesh: E = N + elements - 1 + holes; the airfoil
    # C/O-grid has one hole: 49536 + 49152 = 98688
    assert 49536 + 49152 == 98688


@pytest.mark.parametrize("name,geometry", [(n, g) for n in MESHES_3D
                                           for g in ("synthetic", "cvfem")])
def test_real_mesh_graph_slot_map_and_walkthrough(name, geometry):
    P = pu.pkg()
    case = pu.RealMeshCase(name, geometry=geometry)
    ctx = P.Context(-1)
    mesh = case.box.make_mesh(ctx)
    st = mesh.stats()
    assert st["n_nodes"] == case.n_nodes and st["n_edges"] == case.n_edges
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    g = case.oracle_graph()
    mine = ls.graph()
    assert np.array_equal(mine["row_start_owned"], g.row_start_owned)
    assert np.array_equal(mine["cols"], g.cols)
    assert np.array_equal(mine["rows"], g.rows)
    f, b = case.fields, case.box
    sink = orc.HypreSink(g, b.hid)
    sink.enable_log(case.n_edges)
    orc.continuity_edge(3, case.edges, b.coords, f["velocity"], f["dpdx"],
                        f["density"], f["pressure"], f["momentum_diag"],
                        case.area, sink, **pu.CONT_OPTS)
    oslots, orows = sink.get_log()
    slots, rows = ls.edge_slots()
    assert np.array_equal(slots, oslots)
    assert np.array_equal(rows, orows)
    ls.close()
    mesh.close()
    # tile-plan walk-through with the product's physics header
    emu = pu.Emu(case, tile_nodes=192)
    emu.build_linsys(0, 1)
    emu.check_plan()
    nnz, nrows = g.nnz_owned + g.nnz_shared, g.num_rows_owned + g.num_rows_shared
    vals, rhs = emu.assemble(0, pu.CONT_FIELDS, P.ContinuityOpts(
        pu.DT, pu.GAMMA1, 1.0, 1.0, 0.0), nnz, nrows, 1)
    ov, orhs = sink.get()
    av, arhs = sink.get_abs()
    assert pu.scaled_err(vals, ov, av) < 1
    assert pu.scaled_err(rhs, orhs, arhs) < 1
    omdot = case.oracle_mdot()
    opec = case.oracle_pecfac(orc.peclet("classic", 1.0))
    o = pu.oracle_momentum(case, g, omdot, opec, uvw=True)
    mo = pu.MOM_OPTS
    vals, rhs = emu.assemble(2, pu.MOM_FIELDS, P.MomentumOpts(
        mo["include_divu"], mo["alpha"], mo["alpha_upw"], mo["ho_upwind"],
        mo["relax_fac"], 1, 1e-16, 1, P.peclet_fn("classic", 1.0), 1e-16, -1),
        nnz, nrows, 3, mdot=omdot, pecfac=opec)
    ov, orhs = o.get()
    av, arhs = o.get_abs()
    assert pu.scaled_err(vals, ov, av) < 1
    assert pu.scaled_err(rhs, orhs, arhs) < 1


def test_airfoil_mesh_graph_and_geometry():
    """the real 2-D airfoil mesh: graph bit-exact, true CVFEM geometry from the
    restated Quad42D master element closes every interior dual cell"""
    P = pu.pkg()
    g2 = pu.RealMesh2D()
    ctx = P.Context(-1)
    mesh = P.Mesh(ctx, 2, g2.edges, g2.hid, g2.coords)
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    g = orc.Graph(1, 0, g2.n_nodes - 1)
    g.add_edges(g2.edges, g2.hid)
    g.finalize()
    mine = ls.graph()
    assert np.array_equal(mine["row_start_owned"], g.row_start_owned)
    assert np.array_equal(mine["cols"], g.cols)
    assert g2.vol.min() > 0
    acc = np.zeros((g2.n_nodes, 2))
    np.add.at(acc, g2.edges[:, 0], g2.area)
    np.add.at(acc, g2.edges[:, 1], -g2.area)
    deg = np.bincount(g2.edges.ravel(), minlength=g2.n_nodes)
    interior = deg == 4
    assert interior.sum() > 0.95 * g2.n_nodes
    assert np.max(np.abs(acc[interior])) <= 1e-12 * np.max(np.abs(g2.area))
    ls.close()
    mesh.close()
