"""Seeded random unstructured meshes through the product's host logic (CSR
graph, edge->slot map, tile plans) and the CPU walk-through of the tile kernels,
against the oracle: ragged rows, isolated nodes, periodic aliases that make
several edges hit one (row, column) slot, Dirichlet rows, arbitrary tile sizes.
No GPU."""
import numpy as np
import pytest

import oracle_py as orc
import parity_util as pu


class _Box:
    pass


class FuzzCase:
    """random graph with the attributes parity_util.Case exposes"""

    def __init__(self, seed):
        rng = np.random.default_rng(seed)
        P = pu.pkg()
        synth = __import__("nalu_wind_b200.synth", fromlist=["state"])
        n = int(rng.integers(6, 140))
        coords = rng.random((n, 3)) * np.array([4.0, 3.0, 2.0])
        # mesh-like connectivity: a few nearest neighbours + random long edges
        d2 = ((coords[:, None, :] - coords[None, :, :]) ** 2).sum(-1)
        np.fill_diagonal(d2, np.inf)
        k = int(rng.integers(1, 5))
        pairs = set()
        for a in range(n):
            if rng.random() < 0.08:
                continue  # isolated unless someone else links to it
            for b_ in np.argsort(d2[a])[:k]:
                pairs.add((min(a, int(b_)), max(a, int(b_))))
        for _ in range(int(rng.integers(0, n // 3 + 1))):
            a, b_ = rng.integers(0, n, 2)
            if a != b_:
                pairs.add((min(int(a), int(b_)), max(int(a), int(b_))))
        gid = rng.permutation(n).astype(np.int64) + 1
        own_hid = np.arange(n, dtype=np.int64)
        hid = own_hid.copy()
        # periodic aliases: slave rows resolve to their master's row
        n_slaves = int(rng.integers(0, max(1, n // 8)))
        slaves = rng.choice(n, n_slaves, replace=False)
        masters = {}
        for s in slaves:
            m = int(rng.integers(0, n))
            if m in slaves or m == s:
                continue
            masters[int(s)] = m
            hid[s] = own_hid[m]
        # an edge may not join two nodes of one row
        pairs = [(a, b_) for (a, b_) in sorted(pairs) if hid[a] != hid[b_]]
        edges = np.array(pairs, dtype=np.int32).reshape(-1, 2)
        swap = gid[edges[:, 0]] > gid[edges[:, 1]] if len(edges) else np.zeros(0, bool)
        edges[swap] = edges[swap][:, ::-1]
        edges = edges[rng.permutation(len(edges))]
        b = _Box()
        b.n_nodes, b.n_edges = n, len(edges)
        b.coords, b.gid, b.hid, b.own_hid = coords, gid, hid, own_hid
        b.edges = np.ascontiguousarray(edges)
        b.offsets = np.array([0, n], dtype=np.int64)
        b.rank, b.nranks = 0, 1
        b.periodic = (bool(masters), False)
        dx = coords[edges[:, 1]] - coords[edges[:, 0]]
        b.area = np.ascontiguousarray(
            0.4 * dx + 0.05 * rng.standard_normal(dx.shape))
        b.vol = 0.05 + rng.random(n)

        def make_mesh(ctx, tile_nodes=0, b=b):
            return P.Mesh(ctx, 3, b.edges, b.hid, b.coords,
                          hypre_offsets=b.offsets,
                          node_own_hypre_id=b.own_hid if masters else None,
                          tile_nodes=tile_nodes)
        b.make_mesh = make_mesh
        self.box = b
        pg = hid + 1  # aliases carry their master's state
        self.fields = synth.state(coords, gid, (4.0, 3.0, 2.0), pu.DT, pu.GAMMA1,
                                  periodic_gid=pg)
        self.fields["dual_nodal_volume"] = b.vol
        self.edges, self.area = b.edges, b.area
        self.n_nodes, self.n_edges = n, len(edges)
        self.masters = masters

    oracle_graph = pu.Case.oracle_graph
    oracle_mdot = pu.Case.oracle_mdot
    oracle_pecfac = pu.Case.oracle_pecfac


SEEDS = list(range(24))


@pytest.mark.parametrize("seed", SEEDS)
def test_fuzz_graph_slot_map_and_tile_walkthrough(seed):
    P = pu.pkg()
    case = FuzzCase(seed)
    rng = np.random.default_rng(1000 + seed)
    tile = int(rng.choice([1, 2, 3, 5, 8, 16, 33, 64, 256]))
    rows_all = np.unique(case.box.hid)
    skipped = (np.sort(rng.choice(rows_all, min(len(rows_all), int(rng.integers(0, 4))),
                                  replace=False)).astype(np.int64))
    # ---- graph + slot map, bit-exact ----
    ctx = P.Context(-1)
    mesh = case.box.make_mesh(ctx, tile_nodes=tile)
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    ls.set_skipped_rows(skipped)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    g = case.oracle_graph(skipped=skipped)
    mine = ls.graph()
    for key, ref in (("row_start_owned", g.row_start_owned), ("cols", g.cols),
                     ("rows", g.rows), ("periodic_rows", g.periodic_rows)):
        assert np.array_equal(mine[key], ref), key
    sink = orc.HypreSink(g, case.box.hid)
    sink.enable_log(max(case.n_edges, 1))
    f, b = case.fields, case.box
    orc.continuity_edge(3, case.edges, b.coords, f["velocity"], f["dpdx"],
                        f["density"], f["pressure"], f["momentum_diag"],
                        case.area, sink, **pu.CONT_OPTS)
    if case.n_edges:
        oslots, orows = sink.get_log()
        slots, rows = ls.edge_slots()
        assert np.array_equal(slots, oslots[:case.n_edges])
        assert np.array_equal(rows, orows[:case.n_edges])
    ls.close()
    mesh.close()
    # ---- tile plan walk-through (no skipped rows in the emulator's system) ----
    emu = pu.Emu(case, tile_nodes=tile)
    emu.build_linsys(0, 1)
    emu.check_plan()
    g = case.oracle_graph()
    nnz, nrows = g.nnz_owned + g.nnz_shared, g.num_rows_owned + g.num_rows_shared
    o = pu.oracle_continuity(case, g)
    vals, rhs = emu.assemble(0, pu.CONT_FIELDS, P.ContinuityOpts(
        pu.DT, pu.GAMMA1, 1.0, 1.0, 0.0), nnz, nrows, 1)
    ov, orhs = o.get()
    av, arhs = o.get_abs()
    # rows nothing assembles into keep the reference's reset value (diag 1);
    # the walk-through covers what the tiles write
    touched = av > 0
    assert pu.scaled_err(vals[touched], ov[touched], av[touched]) < 1
    assert pu.scaled_err(rhs, orhs, arhs) < 1
    if case.n_edges:
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
        touched = av > 0
        assert pu.scaled_err(vals[touched], ov[touched], av[touched]) < 1
        assert pu.scaled_err(rhs, orhs, arhs) < 1
        got = emu.nodal_grad(f["velocity"], 3)
        ref = orc.nodal_grad_edge(3, 3, case.edges, f["velocity"], case.area,
                                  f["dual_nodal_volume"], case.n_nodes)
        mag = np.abs(orc.nodal_grad_edge(
            3, 3, case.edges, np.abs(f["velocity"]), np.abs(case.area),
            f["dual_nodal_volume"], case.n_nodes)) + 1e-3 * np.max(np.abs(ref))
        assert pu.scaled_err(got, ref, mag) < 1
