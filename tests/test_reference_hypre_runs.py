"""The oracle's linear-system side (graph, CoeffApplier, loadComplete layout)
against THE REFERENCE'S OWN HypreLinearSystem / HypreUVWLinearSystem, run here.

oracle/Makefile.ref compiles src/HypreLinearSystem.C and
src/HypreUVWLinearSystem.C of the reference -- unmodified -- against stand-ins
for STK / Kokkos / Realm and a RECORDING stand-in for hypre's IJ interface
(oracle/ref_shim/hypre).  What runs is the reference's:
  beginLinearSystemConstruction, buildEdgeToNodeGraph, buildDirichletNodeGraph,
  fill_owned_shared_data_structures[_1DoF], finalizeLinearSystem
  (buildCoeffApplierDevice{Owned,Shared}DataStructures, the periodic node map,
  computeRowSizes), zeroSystem, resetCoeffApplierData, the CoeffApplier's
  operator() -> sort / sum_into / sum_into_1DoF (and the UVW variant) for every
  edge, loadComplete -> hypreIJMatrixSetAddToValues / hypreIJVectorSetAddToValues
(SURVEY.md 8(a) rows a9 - a13).  One process plays one MPI rank at a time.

Compared bit for bit with the oracle: the CSR structures, the value and rhs
arrays after an assembly of blocks that the reference's own edge algorithms
produced, and the arrays the reference hands to HYPRE_IJMatrixSetValues2 /
AddToValues2 / HYPRE_IJVectorSetValues / AddToValues -- the hypre hand-off
layout that no unit test of the reference pins.  Live only (needs
oracle/_ref/libnalu_ref.so)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import oracle_py as orc  # noqa: E402
import parity_util as pu  # noqa: E402
import ref_edge as R  # noqa: E402
import test_reference_edge_runs as T  # noqa: E402

pytestmark = pytest.mark.skipif(
    not R.available(), reason="oracle/_ref not built (needs /root/reference)")


class RankCase:
    """rank `rank` of an `nranks` z-slab decomposition of a (periodic) box with
    the state of test_reference_edge_runs.State"""

    def __init__(self, dims, nranks, rank, periodic=(False, False), seed=3,
                 case=None):
        self.c = c = case if case is not None else pu.Case(
            dims=dims, warp=0.1, nranks=nranks, rank=rank, periodic=periodic)
        b = c.box
        self.b = b
        st = T.State.__new__(T.State)
        f = c.fields
        rng = np.random.default_rng(seed)
        n = c.n_nodes
        st.ndim, st.n_nodes = 3, n
        st.edges = np.ascontiguousarray(np.asarray(c.edges).reshape(-1, 2))
        st.n_edges = len(st.edges)
        st.coords = np.ascontiguousarray(b.coords.reshape(-1, 3))
        st.velocity = np.ascontiguousarray(f["velocity"].reshape(-1, 3))
        st.dudx = np.ascontiguousarray(f["dudx"].reshape(n, 9))
        st.dpdx = np.ascontiguousarray(f["dpdx"].reshape(-1, 3))
        st.dkdx = np.ascontiguousarray(f["dkdx"].reshape(-1, 3))
        st.area = np.ascontiguousarray(np.asarray(c.area).reshape(-1, 3))
        st.viscosity, st.density = f["viscosity"].copy(), f["density"].copy()
        st.pressure, st.udiag = f["pressure"].copy(), f["momentum_diag"].copy()
        st.tke, st.dflux = f["turbulent_ke"].copy(), f["effective_viscosity_tke"].copy()
        st.mask = np.ones(n)
        st.mdot = c.oracle_mdot()
        st.mvof = np.zeros(st.n_edges)
        st.pecfac = c.oracle_pecfac(orc.peclet("classic", 1.0))
        st.source, st.source_mask = np.zeros((n, 3)), np.zeros(n)
        st.efvm, st.vol = np.zeros(st.n_edges), np.ones(n)
        self.st = st
        # STK identifiers (1-based) and, for periodic slaves, the master's
        self.ident = b.gid.astype(np.int64) + 1
        self.nalu = self.ident.copy()
        slave = b.hid != b.own_hid
        if slave.any():
            of_hid = {int(h): int(i) for h, i in zip(b.own_hid[~slave], self.ident[~slave])}
            self.nalu[slave] = [of_hid[int(h)] for h in b.hid[slave]]
        self.slave = slave

    def hypre(self, uvw=False, num_dof=1, dirichlet_nodes=None):
        b = self.b
        w = self.st.world()
        return R.HypreRef(w, b.own_hid, uvw=uvw, num_dof=num_dof, rank=b.rank,
                          nranks=b.nranks, node_identifier=self.ident,
                          node_owner=b.owner, nalu_id=self.nalu,
                          offsets=b.offsets, dirichlet_nodes=dirichlet_nodes)

    def oracle_graph(self, num_dof=1, skipped_nodes=()):
        b = self.b
        sk = []
        for nd in skipped_nodes:
            sk += [int(b.hid[nd]) * num_dof + d for d in range(num_dof)]
        return self.c.oracle_graph(num_dof=num_dof, skipped=np.array(sk, dtype=np.int64))


def check_graph(h, g):
    assert (h.num_rows_owned, h.nnz_owned, h.num_rows_shared, h.nnz_shared) == (
        g.num_rows_owned, g.nnz_owned, g.num_rows_shared, g.nnz_shared)
    assert np.array_equal(h.row_start_owned, g.row_start_owned)
    assert np.array_equal(h.row_start_shared, g.row_start_shared[:len(h.row_start_shared)])
    assert np.array_equal(h.cols, g.cols)
    assert np.array_equal(h.rows, g.rows)
    assert np.array_equal(h.row_indices_shared, g.row_indices_shared)
    assert np.array_equal(np.sort(h.periodic_rows), np.sort(g.periodic_rows))


def check_handoff(h, g, vals, rhs):
    """loadComplete: SetValues2 of the owned triplets, AddToValues2 of the shared
    tail (HypreLinearSystem.C:1572-1590), then the vectors (:1662-1676)"""
    lc = h.load_complete()
    m = lc[0]
    want = [(False, 0, g.nnz_owned)]
    if g.nnz_shared:
        want.append((True, g.nnz_owned, g.nnz_owned + g.nnz_shared))
    assert len(m) == len(want)
    for (add, ncols, rows, cols, v), (wadd, lo, hi) in zip(m, want):
        assert add == wadd and np.all(ncols == 1)
        assert np.array_equal(rows, g.rows[lo:hi])
        assert np.array_equal(cols, g.cols[lo:hi])
        assert np.array_equal(v, vals[lo:hi])
    nro, nrs = g.num_rows_owned, g.num_rows_shared
    for d in range(rhs.shape[0]):
        calls = lc[1 + d]
        assert len(calls) == (2 if nrs else 1)
        add, _, rows, _, v = calls[0]
        assert not add and np.array_equal(rows, np.arange(g.i_lower, g.i_lower + nro))
        assert np.array_equal(v, rhs[d, :nro])
        if nrs:
            add, _, rows, _, v = calls[1]
            assert add and np.array_equal(v, rhs[d, nro:nro + nrs])
            # one rhs row per shared matrix row, in row_indices_shared order
            assert np.array_equal(rows, g.row_indices_shared)


CASES = [
    ("serial", (5, 4, 3), 1, (False, False)),
    ("2-ranks", (4, 3, 6), 2, (False, False)),
    ("3-ranks", (3, 3, 7), 3, (False, False)),
    ("periodic", (5, 4, 3), 1, (True, True)),
    ("periodic-2-ranks", (4, 4, 6), 2, (True, True)),
]


@pytest.mark.parametrize("tag,dims,nranks,periodic", CASES, ids=[c[0] for c in CASES])
def test_scalar_system_graph_values_handoff(tag, dims, nranks, periodic):
    for rank in range(nranks):
        rc = RankCase(dims, nranks, rank, periodic)
        lhs, rhs = T.ref_scalar(rc.st, T.SCAL_POINTS[2], T.PECLETS[2])
        h = rc.hypre()
        g = rc.oracle_graph()
        check_graph(h, g)
        if rank > 0:  # the lowest sharing rank owns an interface node
            assert g.num_rows_shared > 0
        if any(periodic):
            assert g.num_periodic > 0
        vals, r = h.assemble(lhs, rhs)
        s = orc.HypreSink(g, rc.b.hid)
        s.apply(rc.st.edges, lhs, rhs)
        ov, orh = s.get()
        assert np.array_equal(vals, ov)
        assert np.array_equal(r.ravel(), np.asarray(orh).ravel())
        check_handoff(h, g, vals, r)
        h.close()


@pytest.mark.parametrize("tag,dims,nranks,periodic", CASES[:2] + CASES[3:4],
                         ids=[c[0] for c in CASES[:2] + CASES[3:4]])
def test_dirichlet_rows_are_skipped(tag, dims, nranks, periodic):
    for rank in range(nranks):
        rc = RankCase(dims, nranks, rank, periodic)
        z = rc.st.coords[:, 2]
        wall = np.nonzero((z <= z.min() + 1e-9) & ~rc.slave)[0][::2].astype(np.int32)
        assert len(wall) > 3 or rank > 0
        lhs, rhs = T.ref_scalar(rc.st, T.SCAL_POINTS[1], T.PECLETS[0])
        h = rc.hypre(dirichlet_nodes=wall)
        g = rc.oracle_graph(skipped_nodes=wall)
        check_graph(h, g)
        vals, r = h.assemble(lhs, rhs)
        s = orc.HypreSink(g, rc.b.hid)
        s.apply(rc.st.edges, lhs, rhs)
        ov, orh = s.get()
        assert np.array_equal(vals, ov)
        assert np.array_equal(r.ravel(), np.asarray(orh).ravel())
        h.close()


@pytest.mark.parametrize("tag,dims,nranks,periodic", CASES[:2] + CASES[3:],
                         ids=[c[0] for c in CASES[:2] + CASES[3:]])
@pytest.mark.parametrize("system", ["monolithic", "uvw"])
def test_momentum_systems(system, tag, dims, nranks, periodic):
    """the 2 ndim x 2 ndim blocks of the reference's MomentumEdgeSolverAlg through
    HypreLinSysCoeffApplier::sum_into (numDof = 3, bubble sort of six ids) and
    through HypreUVWLinSysCoeffApplier::sum_into (x-x entries, three rhs)"""
    uvw = system == "uvw"
    for rank in range(nranks):
        rc = RankCase(dims, nranks, rank, periodic)
        lhs, rhs = T.ref_momentum(rc.st, T.MOM_POINTS[2])
        h = rc.hypre(uvw=uvw, num_dof=3)
        g = rc.oracle_graph(num_dof=1 if uvw else 3)
        check_graph(h, g)
        vals, r = h.assemble(lhs, rhs)
        s = orc.HypreSink(g, rc.b.hid, uvw_ndim=3 if uvw else 0)
        s.apply(rc.st.edges, lhs, rhs)
        ov, orh = s.get()
        assert np.array_equal(vals, ov)
        orh = np.asarray(orh)
        assert np.array_equal(r.reshape(orh.shape) if r.size == orh.size else r, orh)
        check_handoff(h, g, vals, r)
        h.close()


def _momentum_options(w, o):
    for k, v in (("divU", o["include_divu"]), ("alpha:velocity", o["alpha"]),
                 ("alpha_upw:velocity", o["alpha_upw"]), ("upw:velocity", o["ho_upwind"]),
                 ("relax:velocity", o["relax_fac"]),
                 ("limiter:velocity", 1.0 if o["use_limiter"] else 0.0)):
        w.option(k, v)


@pytest.mark.parametrize("system", ["monolithic", "uvw"])
@pytest.mark.parametrize("nranks", [1, 2])
def test_end_to_end_reference_assembly(system, nranks):
    """the path as the reference runs it, nothing recorded in between:
    zeroSystem, MomentumEdgeSolverAlg::execute whose loop shell hands every
    local block to the reference's CoeffApplier, loadComplete -- against the
    oracle's momentum kernel + sink, with the bench's (decks') options"""
    uvw = system == "uvw"
    o = T.MOM_POINTS[1]
    for rank in range(nranks):
        rc = RankCase((4, 3, 6), nranks, rank, (True, False))
        st = rc.st
        w = st.world()
        _momentum_options(w, o)
        b = rc.b
        h = R.HypreRef(w, b.own_hid, uvw=uvw, num_dof=3, rank=rank, nranks=nranks,
                       node_identifier=rc.ident, node_owner=b.owner,
                       nalu_id=rc.nalu, offsets=b.offsets)
        vals, r = h.sweep("momentum")
        g = rc.oracle_graph(1 if uvw else 3)
        s = orc.HypreSink(g, b.hid, uvw_ndim=3 if uvw else 0)
        orc.set_num_threads(1)
        orc.momentum_edge(3, st.edges, st.coords, st.velocity, st.dudx, st.viscosity,
                          st.density, st.mask, st.area, st.mdot, st.pecfac, s, **o)
        ov, orh = s.get()
        assert np.array_equal(vals, ov)
        assert np.array_equal(r.ravel(), np.asarray(orh).ravel())
        h.close()


@pytest.mark.parametrize("tag,dims,nranks,periodic", CASES[:2] + CASES[3:4],
                         ids=[c[0] for c in CASES[:2] + CASES[3:4]])
def test_reset_rows_and_dirichlet_bcs(tag, dims, nranks, periodic):
    """CoeffApplier::resetRows (FixPressureAtNode) and applyDirichletBCs of the
    reference on an assembled system, against the oracle's sink"""
    for rank in range(nranks):
        rc = RankCase(dims, nranks, rank, periodic)
        st, b = rc.st, rc.b
        z = st.coords[:, 2]
        owned = (b.owner == rank) & ~rc.slave
        wall = np.nonzero((z <= z.min() + 1e-9) & owned)[0][::2].astype(np.int32)
        lhs, rhs = T.ref_scalar(st, T.SCAL_POINTS[1], T.PECLETS[0])
        rng = np.random.default_rng(9)
        bc = rng.standard_normal(st.n_nodes)
        w = st.world()
        w.field("tke_bc", R.NODE, bc, 1)
        h = R.HypreRef(w, b.own_hid, rank=rank, nranks=nranks, node_identifier=rc.ident,
                       node_owner=b.owner, nalu_id=rc.nalu, offsets=b.offsets,
                       dirichlet_nodes=wall)
        g = rc.oracle_graph(skipped_nodes=wall)
        s = orc.HypreSink(g, b.hid)
        h.assemble(lhs, rhs)
        s.apply(st.edges, lhs, rhs)
        # FixPressureAtNode-style reset of a few rows: any node this rank holds,
        # owned or shared
        some = np.arange(st.n_nodes)[3::7].astype(np.int32)
        vals, r = h.reset_rows(some, 2.5, -0.75)
        s.reset_rows(some, 2.5, -0.75)
        ov, orh = s.get()
        assert np.array_equal(vals, ov) and np.array_equal(r.ravel(), np.asarray(orh).ravel())
        if len(wall):
            vals, r = h.apply_dirichlet("turbulent_ke", "tke_bc", wall)
            s.apply_dirichlet(wall, st.tke, bc)
            ov, orh = s.get()
            assert np.array_equal(vals, ov)
            assert np.array_equal(r.ravel(), np.asarray(orh).ravel())
        h.close()


def test_reference_decomposition_hybrid_g_8():
    """the reference's own 8-way STK decomposition of hybrid.g (tet / pyramid /
    hex; nodes shared by up to several ranks): every rank's graph, CoeffApplier
    sums and hand-off through the reference's HypreLinearSystem vs the oracle"""
    shared_total = 0
    for rank in range(8):
        case = pu.DecomposedRealMesh(rank=rank)
        rc = RankCase(None, 8, rank, case=case)
        lhs, rhs = T.ref_scalar(rc.st, T.SCAL_POINTS[1], T.PECLETS[2])
        h = rc.hypre()
        g = rc.oracle_graph()
        check_graph(h, g)
        shared_total += g.num_rows_shared
        vals, r = h.assemble(lhs, rhs)
        s = orc.HypreSink(g, rc.b.hid)
        s.apply(rc.st.edges, lhs, rhs)
        ov, orh = s.get()
        assert np.array_equal(vals, ov)
        assert np.array_equal(r.ravel(), np.asarray(orh).ravel())
        check_handoff(h, g, vals, r)
        h.close()
    assert shared_total > 500


@pytest.mark.parametrize("system", ["monolithic", "uvw"])
def test_extract_diagonal_side_channel(system):
    """NGPApplyCoeff::extract_diagonal (src/SolverAlgorithm.C:87-105): with
    projected_timescale_type momentum_diag_inv the momentum assembly adds
    lhs(ix, ix), the first dof's diagonal entry of each of the two nodes, into
    momentum_diag before the block goes to the CoeffApplier"""
    uvw = system == "uvw"
    o = T.MOM_POINTS[1]
    rc = RankCase((5, 4, 3), 1, 0, (True, False))
    st, b = rc.st, rc.b
    w = st.world()
    _momentum_options(w, o)
    diag = w.field("extracted_diag", R.NODE, np.zeros(st.n_nodes), 1)
    h = R.HypreRef(w, b.own_hid, uvw=uvw, num_dof=3, node_identifier=rc.ident,
                   node_owner=b.owner, nalu_id=rc.nalu, offsets=b.offsets)
    vals, r = h.sweep("momentum", diag_field="extracted_diag")
    g = rc.oracle_graph(1 if uvw else 3)
    s = orc.HypreSink(g, b.hid, uvw_ndim=3 if uvw else 0)
    od = np.zeros(st.n_nodes)
    orc.set_num_threads(1)
    orc.momentum_edge(3, st.edges, st.coords, st.velocity, st.dudx, st.viscosity,
                      st.density, st.mask, st.area, st.mdot, st.pecfac, s,
                      udiag_accum=od, **o)
    ov, orh = s.get()
    assert np.array_equal(vals, ov)
    assert np.array_equal(r.ravel(), np.asarray(orh).ravel())
    assert np.abs(od).max() > 0 and np.array_equal(diag.ravel(), od)
    h.close()
