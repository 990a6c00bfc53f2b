"""CPU tests of the product's host logic (plan builder, CSR graph, edge->slot
map) against the oracle, plus a host walk-through of the tile plans.  No GPU."""
import numpy as np
import pytest

import oracle_py as orc
import parity_util as pu

CASES = [
    dict(dims=(6, 5, 4)),
    dict(dims=(9, 7, 5), periodic=(True, True), lengths=(5000.0, 5000.0, 1000.0)),
    dict(dims=(8, 6, 6), warp=0.15, shuffle_bucket=64),
    dict(dims=(7, 6, 9), nranks=2, rank=0),
    dict(dims=(7, 6, 9), nranks=2, rank=1),
    dict(dims=(5, 4, 9), nranks=3, rank=1, periodic=(True, False)),
    dict(dims=(6, 5, 4), tet_split=True),
]


def _ids(c):
    return "-".join("%s=%s" % kv for kv in sorted(c.items()))


@pytest.mark.parametrize("kw", CASES, ids=_ids)
def test_graph_and_slot_map_bit_exact(kw):
    """CSR graph + edge->slot map of the product's host code must be bit-exact
    with the reference algorithm restated in the oracle
    (src/HypreLinearSystem.C:211-269, 412-478, 999-1236, 2165-2239)."""
    P = pu.pkg()
    case = pu.Case(**kw)
    ctx = P.Context(-1)  # host-only: plan building, no compute
    mesh = case.box.make_mesh(ctx, tile_nodes=48)
    for kind, ndof, uvw in ((P.NW_LINSYS_HYPRE, 1, 0), (P.NW_LINSYS_HYPRE_UVW, 3, 3),
                            (P.NW_LINSYS_HYPRE, 3, 0)):
        ls = P.LinearSystem(mesh, kind, ndof)
        ls.buildEdgeToNodeGraph()
        ls.finalizeLinearSystem()
        gnd = 1 if uvw else ndof
        g = case.oracle_graph(num_dof=gnd)
        mine = ls.graph()
        s = ls.sizes
        assert (s.num_rows_owned, s.num_nonzeros_owned, s.num_rows_shared,
                s.num_nonzeros_shared, s.num_periodic_rows) == (
            g.num_rows_owned, g.nnz_owned, g.num_rows_shared, g.nnz_shared,
            g.num_periodic)
        assert np.array_equal(mine["row_start_owned"], g.row_start_owned)
        assert np.array_equal(mine["row_start_shared"], g.row_start_shared)
        assert np.array_equal(mine["cols"], g.cols)
        assert np.array_equal(mine["rows"], g.rows)
        assert np.array_equal(mine["row_indices_shared"], g.row_indices_shared)
        assert np.array_equal(mine["periodic_rows"], g.periodic_rows)
        # slot map: log where the reference's column walk writes
        sink = orc.HypreSink(g, case.box.hid, uvw_ndim=uvw)
        sink.enable_log(case.n_edges)
        n = 2 * (3 if (uvw or ndof == 3) else 1)
        lib = orc.lib()
        import ctypes as C
        # drive the sink with dummy blocks through the momentum / continuity
        # oracle so that every edge makes exactly one apply() call
        if n == 2:
            pu.orc.continuity_edge(
                3, case.edges, case.box.coords, case.fields["velocity"],
                case.fields["dpdx"], case.fields["density"],
                case.fields["pressure"], case.fields["momentum_diag"],
                case.area, sink, **pu.CONT_OPTS)
        else:
            pu.orc.momentum_edge(
                3, case.edges, case.box.coords, case.fields["velocity"],
                case.fields["dudx"], case.fields["viscosity"],
                case.fields["density"],
                case.fields["abl_wall_no_slip_wall_func_node_mask"], case.area,
                np.zeros(case.n_edges), np.zeros(case.n_edges), sink,
                **pu.MOM_OPTS)
        oslots, orows = sink.get_log()
        slots, rows = ls.edge_slots()
        assert np.array_equal(slots, oslots)
        assert np.array_equal(rows, orows)
        ls.close()
    mesh.close()


def test_skipped_rows_graph():
    """Dirichlet (skipped) rows: owned -> single diagonal entry, shared ->
    dropped (src/HypreLinearSystem.C:1032-1041, 1157-1158)."""
    P = pu.pkg()
    case = pu.Case(dims=(6, 5, 8), nranks=2, rank=1)
    b = case.box
    lo, hi = int(b.offsets[1]), int(b.offsets[2]) - 1
    skipped = np.array([lo, lo + 3, hi, lo - 2, lo - 7], dtype=np.int64)
    ctx = P.Context(-1)
    mesh = b.make_mesh(ctx, tile_nodes=32)
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    ls.set_skipped_rows(skipped)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    g = case.oracle_graph(skipped=skipped)
    mine = ls.graph()
    for k, o in (("row_start_owned", g.row_start_owned),
                 ("row_start_shared", g.row_start_shared), ("cols", g.cols),
                 ("rows", g.rows), ("row_indices_shared", g.row_indices_shared)):
        assert np.array_equal(mine[k], o), k
    sink = orc.HypreSink(g, b.hid)
    sink.enable_log(case.n_edges)
    pu.orc.continuity_edge(3, case.edges, b.coords, case.fields["velocity"],
                           case.fields["dpdx"], case.fields["density"],
                           case.fields["pressure"],
                           case.fields["momentum_diag"], case.area, sink,
                           **pu.CONT_OPTS)
    oslots, orows = sink.get_log()
    slots, rows = ls.edge_slots()
    assert np.array_equal(slots, oslots)
    assert np.array_equal(rows, orows)


@pytest.mark.parametrize("fma", [False, pytest.param(True, marks=pytest.mark.skipif(
    not pu.host_has_fma(), reason="host CPU without FMA"))], ids=["plain", "fma"])
@pytest.mark.parametrize("kw", CASES, ids=_ids)
def test_tile_plan_walkthrough_matches_oracle(kw, fma):
    """Replay the tile kernels' phases on the CPU from the same plan arrays and
    the same physics header; every matrix / rhs entry within 1e-12 of the
    oracle (scaled by the entry's sum of |contributions|).  fma: the header
    compiled with fused multiply-add contraction, as nvcc compiles it for the
    device (the oracle is compiled -ffp-contract=off)."""
    P = pu.pkg()
    case = pu.Case(**kw)
    emu = pu.Emu(case, tile_nodes=40, fma=fma)
    emu.build_linsys(0, 1)
    emu.check_plan()
    g = case.oracle_graph()
    nnz, rows = g.nnz_owned + g.nnz_shared, g.num_rows_owned + g.num_rows_shared

    omdot = case.oracle_mdot()
    mdot = emu.mdot()
    assert pu.scaled_err(mdot, omdot, np.abs(omdot) + 1e-3 * np.max(np.abs(omdot))) < 1

    o = pu.oracle_continuity(case, g)
    vals, rhs = emu.assemble(0, pu.CONT_FIELDS, P.ContinuityOpts(
        pu.DT, pu.GAMMA1, 1.0, 1.0, 0.0), nnz, rows, 1)
    ov, orhs = o.get()
    av, arhs = o.get_abs()
    assert pu.scaled_err(vals, ov, av) < 1
    assert pu.scaled_err(rhs, orhs, arhs) < 1

    o = pu.oracle_scalar(case, g, omdot)
    so = pu.SCAL_OPTS
    vals, rhs = emu.assemble(1, pu.SCAL_FIELDS, P.ScalarOpts(
        so["alpha"], so["alpha_upw"], so["ho_upwind"], so["relax_fac"], 1,
        1e-16, P.peclet_fn("tanh", 2.0, 1.0)), nnz, rows, 1, mdot=omdot)
    ov, orhs = o.get()
    av, arhs = o.get_abs()
    assert pu.scaled_err(vals, ov, av) < 1
    assert pu.scaled_err(rhs, orhs, arhs) < 1

    opec = case.oracle_pecfac(orc.peclet("classic", 1.0))
    o = pu.oracle_momentum(case, g, omdot, opec, uvw=True)
    mo = pu.MOM_OPTS
    for fuse in (0, 1):
        vals, rhs = emu.assemble(2, pu.MOM_FIELDS, P.MomentumOpts(
            mo["include_divu"], mo["alpha"], mo["alpha_upw"], mo["ho_upwind"],
            mo["relax_fac"], 1, 1e-16, fuse, P.peclet_fn("classic", 1.0),
            1e-16, -1), nnz, rows, 3, mdot=omdot, pecfac=opec)
        ov, orhs = o.get()
        av, arhs = o.get_abs()
        assert pu.scaled_err(vals, ov, av) < 1
        assert pu.scaled_err(rhs, orhs, arhs) < 1

    # monolithic 3-dof momentum (src/HypreLinearSystem.C:2059-2161) through the
    # same node-graph plan: three rows per node, blocks formed in the row walk
    g3 = case.oracle_graph(num_dof=3)
    o3 = pu.oracle_momentum(case, g3, omdot, opec, uvw=False)
    vals, rhs = emu.assemble_mono(pu.MOM_FIELDS, P.MomentumOpts(
        mo["include_divu"], mo["alpha"], mo["alpha_upw"], mo["ho_upwind"],
        mo["relax_fac"], 1, 1e-16, 0, P.peclet_fn("classic", 1.0), 1e-16, -1),
        g3.nnz_owned + g3.nnz_shared, g3.num_rows_owned + g3.num_rows_shared,
        mdot=omdot, pecfac=opec)
    ov, orhs = o3.get()
    av, arhs = o3.get_abs()
    assert pu.scaled_err(vals, ov, av) < 1
    assert pu.scaled_err(rhs, orhs.ravel(), arhs.ravel()) < 1
    if not kw.get("periodic") and kw.get("nranks", 1) == 1:
        # with Dirichlet nodes: the twin skips their node rows
        nodes = np.array([2, 11, 30], dtype=np.int64)
        sk3 = (3 * nodes[:, None] + np.arange(3)).ravel()
        emu_s = pu.Emu(case, tile_nodes=40, fma=fma)
        emu_s.build_linsys(0, 1, skipped=nodes)
        g3s = case.oracle_graph(num_dof=3, skipped=sk3)
        o3s = pu.oracle_momentum(case, g3s, omdot, opec, uvw=False)
        vals, rhs = emu_s.assemble_mono(pu.MOM_FIELDS, P.MomentumOpts(
            mo["include_divu"], mo["alpha"], mo["alpha_upw"], mo["ho_upwind"],
            mo["relax_fac"], 1, 1e-16, 0, P.peclet_fn("classic", 1.0), 1e-16, -1),
            g3s.nnz_owned + g3s.nnz_shared,
            g3s.num_rows_owned + g3s.num_rows_shared, mdot=omdot, pecfac=opec,
            skipped3=sk3)
        ov, orhs = o3s.get()
        av, arhs = o3s.get_abs()
        assert pu.scaled_err(vals, ov, av) < 1
        assert pu.scaled_err(rhs, orhs.ravel(), arhs.ravel()) < 1

    f = case.fields
    for phi, d1 in (("pressure", 1), ("velocity", 3)):
        got = emu.nodal_grad(f[phi], d1)
        ref = orc.nodal_grad_edge(d1, 3, case.edges, f[phi], case.area,
                                  f["dual_nodal_volume"], case.n_nodes)
        mag = np.abs(orc.nodal_grad_edge(
            d1, 3, case.edges, np.abs(f[phi]), np.abs(case.area),
            f["dual_nodal_volume"], case.n_nodes)) + 1e-3 * np.max(np.abs(ref))
        assert pu.scaled_err(got, ref, mag) < 1


@pytest.mark.parametrize("kw", [CASES[2], CASES[6], CASES[4]], ids=_ids)
def test_tile_cut_refinement_keeps_the_plan_valid(kw, monkeypatch):
    """NW_TILE_REFINE=1 (opt-in greedy cut refinement after the RCB,
    plan.cpp): tiles change, the plan invariants and the assembled systems do
    not; the cut never grows."""
    P = pu.pkg()
    case = pu.Case(**kw)
    stats = {}
    for on in ("0", "1"):
        monkeypatch.setenv("NW_TILE_REFINE", on)
        emu = pu.Emu(case, tile_nodes=40)
        emu.build_linsys(0, 1)
        emu.check_plan()
        g = case.oracle_graph()
        nnz = g.nnz_owned + g.nnz_shared
        rows = g.num_rows_owned + g.num_rows_shared
        o = pu.oracle_continuity(case, g)
        vals, rhs = emu.assemble(0, pu.CONT_FIELDS, P.ContinuityOpts(
            pu.DT, pu.GAMMA1, 1.0, 1.0, 0.0), nnz, rows, 1)
        ov, orhs = o.get()
        av, arhs = o.get_abs()
        assert pu.scaled_err(vals, ov, av) < 1
        assert pu.scaled_err(rhs, orhs, arhs) < 1
        ctx = P.Context(-1)
        mesh = case.box.make_mesh(ctx, tile_nodes=40)
        stats[on] = mesh.stats()
        mesh.close()
    assert stats["1"]["n_tile_edges"] <= stats["0"]["n_tile_edges"]


@pytest.mark.parametrize("kw", CASES, ids=_ids)
def test_monolithic_system_shares_the_node_graph_plan(kw):
    """The numDof = ndim system (src/HypreLinearSystem.C:2059-2161) takes the tile
    path when its graph is the exact blow-up of the node graph (checked row by
    row by the builder: ndim entries per node-row entry, same column order, the
    ndim rows of a node contiguous); skipped rows keep the atomic kernel."""
    P = pu.pkg()
    case = pu.Case(**kw)
    ctx = P.Context(-1)
    mesh = case.box.make_mesh(ctx, tile_nodes=48)
    one = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    one.buildEdgeToNodeGraph()
    one.finalizeLinearSystem()
    mono = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 3)
    mono.buildEdgeToNodeGraph()
    mono.finalizeLinearSystem()
    assert one.uses_tile_path() and mono.uses_tile_path()
    # the blow-up the builder relies on, restated with the exported graphs
    g1, g3 = one.graph(), mono.graph()
    s1, s3 = one.sizes, mono.sizes
    assert s3.num_rows_owned == 3 * s1.num_rows_owned
    assert s3.num_rows_shared == 3 * s1.num_rows_shared
    per3 = set(int(r) - s3.i_lower for r in g3["periodic_rows"])
    rs1, rs3 = g1["row_start_owned"], g3["row_start_owned"]
    for r in range(s1.num_rows_owned):
        if 3 * r in per3:
            continue  # periodic slave: a lone diagonal in both graphs
        c1 = g1["cols"][rs1[r]:rs1[r + 1]]
        for i in range(3):
            c3 = g3["cols"][rs3[3 * r + i]:rs3[3 * r + i + 1]]
            assert np.array_equal(c3, (3 * c1[:, None] + np.arange(3)).ravel())
    # Dirichlet nodes (all dofs of a node, as applyDirichletBCs lists them):
    # skipped node rows of the twin -- still the tile path; a list that covers
    # only the first dof of a node keeps the atomic kernel
    lo = int(case.box.offsets[case.box.rank]) * 3
    skip = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 3)
    skip.set_skipped_rows(np.array([lo, lo + 1, lo + 2, lo + 9, lo + 10, lo + 11],
                                   dtype=np.int64))
    skip.buildEdgeToNodeGraph()
    skip.finalizeLinearSystem()
    assert skip.uses_tile_path()
    part = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 3)
    part.set_skipped_rows(np.array([lo], dtype=np.int64))
    part.buildEdgeToNodeGraph()
    part.finalizeLinearSystem()
    assert not part.uses_tile_path()
    for ls in (one, mono, skip, part):
        ls.close()
    mesh.close()


def test_abi_library_exports_every_declared_symbol():
    """the C-ABI library loads without a GPU and exports every function that
    include/nalu_edge_b200.h declares; compute calls fail loudly (no fallback)"""
    import re
    import os
    P = pu.pkg()
    L = P.lib()
    hdr = open(os.path.join(pu.ROOT, "include", "nalu_edge_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(
        r"^(?:int|const char\*|void\*)\s+(nw_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert declared == set(P.ABI_SYMBOLS), declared ^ set(P.ABI_SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    import torch
    if not torch.cuda.is_available():
        ctx = P.Context(-1)
        case = pu.Case(dims=(3, 3, 3))
        mesh = case.box.make_mesh(ctx)
        with pytest.raises(P.NwError, match="no CPU fallback"):
            mesh.mdot_edge()
        with pytest.raises(P.NwError):
            P.Context(0)
