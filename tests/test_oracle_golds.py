"""Pin the CPU oracle against the reference's own golden vectors
(tests/golden/reference_golds.json, extracted by
tests/golden/extract_reference_golds.py)."""
import json
import os

import numpy as np

import oracle_py as orc
import parity_util as pu
import unit_cube as uc

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden",
                                "reference_golds.json")))


def test_mdot_edge_gold():
    # UnitTestMdotAlg.C:21-78: rho=1, u=(10,10,10), p=0, dpdx=0 => 2.5, tol 1e-14
    c, e = uc.mesh()
    n = len(c)
    av = uc.edge_area(c, e)
    mdot = orc.mdot_edge(3, e, c, np.full((n, 3), 10.0), np.zeros((n, 3)),
                         np.ones(n), np.zeros(n), np.ones(n), av)
    assert len(mdot) == 12
    assert np.max(np.abs(mdot - G["mdot_edge_value"])) <= 1e-14


def test_nodal_grad_scalar_gold():
    # UnitTestNodalGradAlg.C:22-70, phi = 2x+2y+2z, tol 1e-16 (exact here)
    c, e = uc.mesh()
    phi = 2 * c[:, 0] + 2 * c[:, 1] + 2 * c[:, 2]
    g = orc.nodal_grad_edge(1, 3, e, phi, uc.edge_area(c, e),
                            np.full(len(c), 0.125), len(c))
    assert np.max(np.abs(g.ravel() - np.array(G["nodal_grad_scalar"]))) <= 1e-16


def test_nodal_grad_vector_gold():
    # UnitTestNodalGradAlg.C:72-123: velocity_i = 2 x_i
    # (unit_tests/ngp_algorithms/UnitTestNgpAlgUtils.C:41-48), diagonal of dudx
    c, e = uc.mesh()
    phi = 2.0 * c
    g = orc.nodal_grad_edge(3, 3, e, phi, uc.edge_area(c, e),
                            np.full(len(c), 0.125), len(c))
    diag = g[:, [0, 4, 8]].ravel()
    assert np.max(np.abs(diag - np.array(G["nodal_grad_vector_diag"]))) <= 1e-16


def test_continuity_gold():
    # UnitTestContinuityAdvEdge.C:109-141, dt = gamma1 = 1, tol 1e-12
    c, e = uc.mesh()
    n = len(c)
    sink = orc.DenseSink(n, 1)
    orc.continuity_edge(3, e, c, uc.velocity(c), uc.dpdx(c), np.ones(n),
                        uc.pressure(c), np.ones(n), uc.edge_area(c, e), sink,
                        dt=1.0, gamma1=1.0, noc_fac=1.0, interp_together=1.0)
    lhs, rhs = sink.get()
    assert np.max(np.abs(rhs - np.array(G["continuity_adv"]["rhs"]))) <= 1e-12
    assert np.max(np.abs(lhs - np.array(G["continuity_adv"]["lhs"]))) <= 1e-12


def test_momentum_gold():
    # UnitTestMomentumAdvDiffEdge.C:232-266: alpha=alpha_upw=upw=0, tol 1e-12
    c, e = uc.mesh()
    n = len(c)
    av = uc.edge_area(c, e)
    vel, rho = uc.velocity(c), np.ones(n)
    sink = orc.DenseSink(n, 3)
    orc.momentum_edge(3, e, c, vel, uc.dudx(c), np.full(n, 0.1), rho,
                      np.ones(n), av, uc.fixture_mdot(e, vel, rho, av),
                      np.zeros(len(e)), sink, include_divu=0.0, alpha=0.0,
                      alpha_upw=0.0, ho_upwind=0.0, relax_fac=1.0)
    lhs, rhs = sink.get()
    assert np.max(np.abs(rhs - np.array(G["momentum_adv_diff"]["rhs"]))) <= 1e-12
    assert np.max(np.abs(lhs - np.array(G["momentum_adv_diff"]["lhs"]))) <= 1e-12


def _scalar_case(nz):
    c, e = uc.mesh(nz)
    av = uc.edge_area(c, e, nz)
    z, rho, visc = uc.mixture_fraction_fields(c)
    vel = uc.velocity(c)
    return c, e, av, z, rho, visc, vel


def _run_scalar(c, e, av, z, rho, visc, vel, mdot, sink):
    orc.scalar_edge(3, e, c, vel, z, np.zeros((len(c), 3)), rho, visc, av,
                    mdot, sink, alpha=0.0, alpha_upw=0.0, ho_upwind=0.0,
                    relax_fac=1.0, pf=orc.peclet("classic", 0.0))


def test_scalar_serial_csr_gold():
    # UnitTestScalarAdvDiffEdge.C:24-42, 78-80 (serial CSR through the real
    # graph + column-walk scatter)
    c, e, av, z, rho, visc, vel = _scalar_case(1)
    n = len(c)
    hid = np.arange(n, dtype=np.int64)
    g = orc.Graph(1, 0, n - 1)
    g.add_edges(e, hid)
    g.finalize()
    gold = G["scalar_adv_diff"]["serial"]
    assert g.row_start_owned.tolist() == gold["rowOffsets"]
    assert g.cols.tolist() == gold["cols"]
    sink = orc.HypreSink(g, hid)
    _run_scalar(c, e, av, z, rho, visc, vel, uc.fixture_mdot(e, vel, rho, av),
                sink)
    vals, rhs = sink.get()
    assert np.max(np.abs(vals - np.array(gold["vals"]))) <= 1e-12
    assert np.max(np.abs(rhs[0] - np.array(gold["rhs"]))) <= 1e-12


def test_scalar_two_rank_csr_gold():
    """UnitTestScalarAdvDiffEdge.C:91-143: generated:1x1x2 on 2 ranks.  Rank 0
    owns nodes 1-8 (element 1), rank 1 owns nodes 9-12; interface rows 4-7 get
    rank 1's contributions through the shared-row tail + owner add (the halo
    sum), which is what the 1.85e-5 diagonals of the P0 gold contain."""
    c, e, av, z, rho, visc, vel = _scalar_case(2)
    n = len(c)
    hid = np.arange(n, dtype=np.int64)
    # STK ownership: element el -> rank el; an edge on the shared face (z=1)
    # is owned by the lower rank.
    zmid = 0.5 * (c[e[:, 0], 2] + c[e[:, 1], 2])
    edge_rank = np.where(zmid <= 1.0, 0, 1)
    own = [(0, 7), (8, 11)]
    # Effective mdot (see unit_cube.fixture_mdot): in parallel STK keeps shared
    # and non-shared edges in separate buckets, and the stride-3 read of the
    # packed mass_flow_rate then only picks edges with zero velocity along them
    # (x-edges at y=0, y-edges at x=0, vertical edges) on both ranks -- the P0 /
    # P1 golds are indeed pure diffusion.  So the kernels see mdot == 0 here.
    graphs, sinks, locs = [], [], []
    for r in range(2):
        er = e[edge_rank == r]
        avr = av[edge_rank == r]
        g = orc.Graph(1, own[r][0], own[r][1])
        g.add_edges(er, hid)
        g.finalize()
        s = orc.HypreSink(g, hid)
        _run_scalar(c, er, avr, z, rho, visc, vel, np.zeros(len(er)), s)
        graphs.append(g)
        sinks.append(s)
    # halo sum: owner adds the other rank's shared rows (hypre AddToValues2 +
    # Assemble, src/HypreLinearSystem.C:1585-1590, 1791)
    for r in range(2):
        g, s = graphs[r], sinks[r]
        vals, rhs = s.get()
        vals, rhs = vals[:g.nnz_owned].copy(), rhs[0, :g.num_rows_owned].copy()
        o = 1 - r
        go, so = graphs[o], sinks[o]
        vo, ro = so.get()
        for i, row in enumerate(go.row_indices_shared):
            if not (own[r][0] <= row <= own[r][1]):
                continue
            a = go.nnz_owned + go.row_start_shared[i]
            b = go.nnz_owned + go.row_start_shared[i + 1]
            lr = row - own[r][0]
            for k in range(a, b):
                col = go.cols[k]
                seg = g.cols[g.row_start_owned[lr]:g.row_start_owned[lr + 1]]
                pos = g.row_start_owned[lr] + int(np.searchsorted(seg, col))
                if pos < g.row_start_owned[lr + 1] and g.cols[pos] == col:
                    vals[pos] += vo[k]
                else:
                    # column not in the owner's local graph: Tpetra/hypre
                    # append it; collect for the gold comparison below
                    locs.append((r, row, col, vo[k]))
            rhs[lr] += ro[0, go.num_rows_owned + i]
        gold = G["scalar_adv_diff"]["P%d" % r]
        # gold columns are Tpetra local ids: owned rows first, then ghosts
        ghosts = sorted(set(g.cols[:g.nnz_owned].tolist()) -
                        set(range(own[r][0], own[r][1] + 1)))
        extra = sorted(set(cc for (rr, _, cc, _) in locs if rr == r) -
                       set(range(own[r][0], own[r][1] + 1)) - set(ghosts))
        lid = {gid: i for i, gid in enumerate(
            list(range(own[r][0], own[r][1] + 1)) + ghosts + extra)}
        # assemble final rows as dicts and compare to the gold CSR
        for lr in range(g.num_rows_owned):
            a, b = g.row_start_owned[lr], g.row_start_owned[lr + 1]
            rowd = {lid[int(cg)]: vals[k] for k, cg in
                    zip(range(a, b), g.cols[a:b])}
            for (rr, row, col, v) in locs:
                if rr == r and row - own[r][0] == lr:
                    rowd[lid[int(col)]] = rowd.get(lid[int(col)], 0.0) + v
            ga, gb = gold["rowOffsets"][lr], gold["rowOffsets"][lr + 1]
            gd = dict(zip(gold["cols"][ga:gb], gold["vals"][ga:gb]))
            assert sorted(rowd) == sorted(gd), (r, lr, rowd, gd)
            for k in gd:
                assert abs(rowd[k] - gd[k]) <= 1e-12, (r, lr, k)
        assert np.max(np.abs(rhs - np.array(gold["rhs"]))) <= 1e-12


def test_scalar_fix_pressure_gold():
    """UnitTestScalarAdvDiffEdge.C:236-300 (NGP_adv_diff_edge_tpetra_fix_pressure_at_node):
    after the edge assembly FixPressureAtNodeAlgorithm::execute resets the row of
    STK node 1 (CoeffApplier::resetRows) and sums the 1x1 block lhs = 1,
    rhs = refPressure - p into it (src/FixPressureAtNodeAlgorithm.C:103-116; the
    test passes the density field as 'pressure', refPressure = 1)."""
    c, e, av, z, rho, visc, vel = _scalar_case(1)
    n = len(c)
    hid = np.arange(n, dtype=np.int64)
    g = orc.Graph(1, 0, n - 1)
    g.add_edges(e, hid)
    g.finalize()
    sink = orc.HypreSink(g, hid)
    _run_scalar(c, e, av, z, rho, visc, vel, uc.fixture_mdot(e, vel, rho, av),
                sink)
    sink.reset_rows([0])
    sink.apply([[0]], [[[1.0]]], [[1.0 - rho[0]]])
    vals, rhs = sink.get()
    gold = G["scalar_adv_diff"]["fixed_serial"]
    assert np.max(np.abs(vals - np.array(gold["vals"]))) <= 1e-12
    assert np.max(np.abs(rhs[0] - np.array(gold["rhs"]))) <= 1e-12


def test_scalar_dirichlet_gold():
    """UnitTestScalarAdvDiffEdge.C:302-380 (NGP_adv_diff_edge_tpetra_dirichlet):
    every node of the part is a Dirichlet node, solution = 2, bc = 1 at STK node
    1 and 0 elsewhere.  The gold is the Tpetra system (full rows, zeroed); the
    hypre system of this path keeps Dirichlet rows as skipped diagonal-only rows
    (src/HypreLinearSystem.C:1032-1034) that no kernel assembles into and
    applyDirichletBCs sets to (1, bc - solution) (:2446-2456): the same matrix,
    compared densely."""
    c, e, av, z, rho, visc, vel = _scalar_case(1)
    n = len(c)
    hid = np.arange(n, dtype=np.int64)
    g = orc.Graph(1, 0, n - 1)
    g.set_skipped(hid)
    g.add_edges(e, hid)
    g.finalize()
    assert g.nnz_owned == n  # diagonal-only rows
    sink = orc.HypreSink(g, hid)
    _run_scalar(c, e, av, z, rho, visc, vel, uc.fixture_mdot(e, vel, rho, av),
                sink)
    sol = np.full(n, 2.0)
    bc = np.zeros(n)
    bc[0] = 1.0
    sink.apply_dirichlet(np.arange(n), sol, bc)
    vals, rhs = sink.get()
    gold = G["scalar_adv_diff"]["dirichlet_serial"]
    ref = G["scalar_adv_diff"]["serial"]
    dense_gold = np.zeros((n, n))
    for r in range(n):
        for k in range(ref["rowOffsets"][r], ref["rowOffsets"][r + 1]):
            dense_gold[r, ref["cols"][k]] = gold["vals"][k]
    dense = np.zeros((n, n))
    dense[np.arange(n), g.cols[:n]] = vals
    assert np.array_equal(dense, dense_gold)
    assert np.max(np.abs(rhs[0] - np.array(gold["rhs"]))) <= 1e-12


def _bdf_states(x):
    """the unit-test fixtures only initialise StateNP1; N / NM1 stay zero"""
    z = np.zeros_like(x)
    return (z, z, x)


def test_scalar_mass_bdf_node_gold():
    """UnitTestScalarMassBDFNodeKernel.C:106-150 (dt = 0.1, gamma = 1, -1, 0)"""
    c, e, av, z, rho, visc, vel = _scalar_case(1)
    n = len(c)
    dnv = np.full(n, 0.125)
    sink = orc.DenseSink(n, 1)
    orc.scalar_mass_bdf_node(np.arange(n), _bdf_states(z), _bdf_states(rho),
                             (dnv, dnv, dnv), 0.1, 1.0, -1.0, 0.0, sink)
    lhs, rhs = sink.get()
    gold = G["scalar_mass_bdf_node"]
    assert np.max(np.abs(lhs - np.array(gold["lhs"]))) <= 1e-12
    assert np.max(np.abs(rhs - np.array(gold["rhs"]))) <= 1e-12


def test_momentum_mass_bdf_node_gold():
    """UnitTestMomentumMassBDFNodeKernel.C:51-100: lhs = 1.25 I, rhs gold"""
    c, e = uc.mesh(1)
    n = len(c)
    vel, dp = uc.velocity(c), uc.dpdx(c)
    rho, dnv = np.ones(n), np.full(n, 0.125)
    sink = orc.DenseSink(n, 3)
    orc.momentum_mass_bdf_node(3, np.arange(n), _bdf_states(vel),
                               _bdf_states(rho), (dnv, dnv, dnv), dp, 0.1, 1.0,
                               -1.0, 0.0, sink)
    lhs, rhs = sink.get()
    gold = G["momentum_mass_bdf_node"]
    assert np.max(np.abs(lhs - gold["lhs_diag"] * np.eye(3 * n))) <= 1e-12
    assert np.max(np.abs(rhs - np.array(gold["rhs"]))) <= 1e-12


def test_continuity_mass_bdf_node_gold():
    """UnitTestContinuityMassBDFNodeKernel.C:17-46: rhs = -12.5 everywhere"""
    c, e = uc.mesh(1)
    n = len(c)
    rho, dnv = np.ones(n), np.full(n, 0.125)
    sink = orc.DenseSink(n, 1)
    orc.continuity_mass_bdf_node(np.arange(n), _bdf_states(rho), (dnv, dnv, dnv),
                                 0.1, 1.0, -1.0, 0.0, sink)
    lhs, rhs = sink.get()
    assert np.max(np.abs(lhs)) == 0.0
    assert np.max(np.abs(rhs - G["continuity_mass_bdf_node"]["rhs_all"])) <= 1e-12


def test_wall_dist_edge_gold():
    """UnitTestWallDistEdgeSolver.C:103-125: 2x2 Laplacian blocks, no rhs"""
    c, e = uc.mesh(1)
    n = len(c)
    sink = orc.DenseSink(n, 1)
    orc.wall_dist_edge(3, e, c, uc.edge_area(c, e), sink)
    lhs, rhs = sink.get()
    assert np.max(np.abs(lhs - np.array(G["wall_dist_edge"]["lhs"]))) <= 1e-12
    assert np.max(np.abs(rhs)) == 0.0


def test_geometry_interior_hex8_gold():
    """UnitTestGeometryAlg.C:25-101 (NGP_geometry_interior, generated:1x1x1):
    dual volume 0.125 on 8 nodes, element volume 1, |edge area|^2 = 0.25^2 on
    12 edges, tol 1e-16; directions follow the L->R (ascending id) rule that
    the fixture's edge_area_vector uses"""
    c, e = uc.mesh(1)
    elem = np.array([uc.HEX_LOCAL_TO_ID], dtype=np.int32)
    dnv, ev, area = orc.geometry_interior_hex8(elem, c, e, len(c))
    assert np.max(np.abs(dnv - 0.125)) <= 1e-16
    assert abs(ev[0] - 1.0) <= 1e-16
    assert np.max(np.abs(np.sum(area * area, axis=1) - 0.0625)) <= 1e-16
    assert np.max(np.abs(area - uc.edge_area(c, e))) <= 1e-16


def test_geometry_interior_hex8_vs_generator():
    """on a warped, stretched box the restated Grandy volumes / triangulated
    SCS areas agree with the mesh generator's independent evaluation of the
    same dual mesh"""
    case = pu.Case(dims=(6, 5, 4), warp=0.15, zstretch=1.1)
    b = case.box
    elems = pu.box_hex_elements(b)
    dnv, ev, area = orc.geometry_interior_hex8(elems, b.coords, b.edges, b.n_nodes)
    assert np.max(np.abs(dnv - b.vol)) <= 1e-12 * np.max(b.vol)
    assert np.max(np.abs(area - b.area)) <= 1e-12 * np.max(np.abs(b.area))


def test_geometry_interior_quad4_properties():
    """Quad42DSCV / Quad42DSCS restatement: the unit square gives 0.25 per
    node and 0.5 L->R normals; on the curvilinear O-grid the sub-volumes tile
    the elements and every interior dual cell closes"""
    c = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=float)
    e = np.array([[0, 1], [1, 2], [2, 3], [0, 3]], dtype=np.int32)
    dnv, ev, area = orc.geometry_interior_quad4([[0, 1, 2, 3]], c, e, 4)
    assert np.max(np.abs(dnv - 0.25)) <= 1e-9  # 9-digit Gauss points (:146-147)
    assert np.array_equal(area, [[0.5, 0], [0, 0.5], [-0.5, 0], [0, 0.5]])
    g = pu.OGrid2D()
    dnv, ev, area = orc.geometry_interior_quad4(g.elems, g.coords, g.edges, g.n_nodes)
    assert ev.min() > 0 and abs(dnv.sum() - ev.sum()) <= 1e-12 * ev.sum()
    acc = np.zeros((g.n_nodes, 2))
    np.add.at(acc, g.edges[:, 0], area)
    np.add.at(acc, g.edges[:, 1], -area)
    j = np.arange(g.n_nodes) // 48
    assert np.max(np.abs(acc[(j > 0) & (j < 19)])) <= 1e-14


def test_gold_peclet_function():
    """UnitTestPecletFunction.C:36-100: known answers of the classic and tanh
    blending functions (tolerance 1e-6 as there) -- the oracle's restatement and
    the host build of the product's peclet_eval (edge_physics.h)"""
    import ctypes as C
    g = G["peclet_function"]
    L = pu.emu_lib()
    L.emu_peclet_eval.restype = C.c_double
    L.emu_peclet_eval.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    n = 0
    for tag, form, a, b in (("classic", "classic", g["classic"]["hybridFactor"], 1.0),
                            ("tanh", "tanh", g["tanh"]["c1"], g["tanh"]["c2"]),
                            ("tanh_simd", "tanh", g["tanh_simd"]["c1"], g["tanh_simd"]["c2"])):
        c = g[tag]
        pf = orc.peclet(form, a, b)
        for pec, want in zip(c["peclet_numbers"], c["peclet_factors"]):
            assert abs(orc.peclet_eval(pf, pec) - want) <= c["tolerance"], (tag, pec)
            got = L.emu_peclet_eval(0 if form == "classic" else 1, a, b, pec)
            assert abs(got - want) <= c["tolerance"], (tag, pec, got)
            # the two restatements agree far tighter than the reference's bar
            assert abs(got - orc.peclet_eval(pf, pec)) <= 1e-15
            n += 1
    assert n == 10
