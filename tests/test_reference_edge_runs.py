"""The oracle's edge kernels against THE REFERENCE'S OWN EDGE ALGORITHMS, run here.

oracle/Makefile.ref compiles src/edge_kernels/{Momentum,Scalar,Continuity}
EdgeSolverAlg.C and src/ngp_algorithms/MdotEdgeAlg.C of the reference --
(and NodalGradEdgeAlg.C, MomentumEdgePecletAlg.C, WallDistEdgeSolverAlg.C) --
unmodified, from where they lie -- over stand-ins for Realm / STK / Kokkos
(oracle/ref_shim; DESIGN.md section 4 says exactly what is the reference's and
what is a stand-in).  Their constructors and execute() bodies, i.e. the per-edge
lambdas SURVEY.md 8(a) a1 / a4 / a5 / a6 cite, run on the arrays below, and the
local blocks they leave in smdata.lhs / smdata.rhs are compared edge by edge,
BIT FOR BIT, with what the oracle hands to its sink.

Two forms:
  * live (needs oracle/_ref/libnalu_ref.so: the build container, and the GPU box
    through the shipped file): option matrices on warped 3-D and 2-D meshes,
    including every branch the reference's unit-test golds leave at zero --
    alpha, alpha_upw, hoUpwind, limiter, tanh blending, divU, VOF, balanced
    buoyancy, GCL;
  * fixture (tests/golden/reference_edge_runs.npz, written by
    tests/golden/extract_reference_runs.py from the same library): a small case
    with inputs and the reference's outputs committed, checked everywhere.
"""
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

live = pytest.mark.skipif(not R.available(),
                          reason="oracle/_ref not built (needs /root/reference)")

MOM_POINTS = [
    dict(include_divu=0.0, alpha=0.0, alpha_upw=0.0, ho_upwind=0.0, relax_fac=1.0, use_limiter=False),  # golds
    dict(include_divu=0.0, alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, relax_fac=0.7, use_limiter=True),   # decks, bench
    dict(include_divu=1.0, alpha=0.4, alpha_upw=0.6, ho_upwind=0.5, relax_fac=0.7, use_limiter=True),
    dict(include_divu=1.0, alpha=1.0, alpha_upw=1.0, ho_upwind=0.0, relax_fac=1.0, use_limiter=False),
    dict(include_divu=0.0, alpha=1.0, alpha_upw=0.0, ho_upwind=1.0, relax_fac=0.9, use_limiter=True),
]
SCAL_POINTS = [
    dict(alpha=0.0, alpha_upw=0.0, ho_upwind=0.0, relax_fac=1.0, use_limiter=False),
    dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, relax_fac=0.9, use_limiter=True),
    dict(alpha=0.4, alpha_upw=0.6, ho_upwind=0.5, relax_fac=0.7, use_limiter=True),
    dict(alpha=1.0, alpha_upw=1.0, ho_upwind=0.0, relax_fac=1.0, use_limiter=False),
]
PECLETS = [("classic", 1.0, 1.0), ("classic", 0.37, 1.0), ("tanh", 2.0, 1.0),
           ("tanh", 5000.0, 200.0)]


class State:
    """a mesh and a state for both sides; 2-D cases take the x-y part of a
    one-layer box"""

    def __init__(self, ndim=3, dims=(5, 4, 3), warp=0.15, seed=3, two_phase=False):
        c = pu.Case(dims=dims, warp=warp)
        f, b = c.fields, c.box
        rng = np.random.default_rng(seed)
        n = c.n_nodes
        self.ndim = ndim
        if ndim == 3:
            keep_e = np.ones(c.n_edges, dtype=bool)
        else:
            # edges lying in the plane of the first node layer
            z = b.coords.reshape(-1, 3)[:, 2]
            layer = z <= z.min() + 1e-9 + 0.3 * (z.max() - z.min()) / dims[2]
            e = np.asarray(c.edges).reshape(-1, 2)
            keep_e = layer[e[:, 0]] & layer[e[:, 1]]
        self.edges = np.ascontiguousarray(np.asarray(c.edges).reshape(-1, 2)[keep_e])
        self.n_nodes, self.n_edges = n, len(self.edges)
        d = ndim
        self.coords = np.ascontiguousarray(b.coords.reshape(-1, 3)[:, :d])
        self.velocity = np.ascontiguousarray(f["velocity"].reshape(-1, 3)[:, :d])
        self.dudx = np.ascontiguousarray(f["dudx"].reshape(-1, 3, 3)[:, :d, :d]).reshape(n, d * d)
        self.dpdx = np.ascontiguousarray(f["dpdx"].reshape(-1, 3)[:, :d])
        self.dkdx = np.ascontiguousarray(f["dkdx"].reshape(-1, 3)[:, :d])
        self.area = np.ascontiguousarray(np.asarray(c.area).reshape(-1, 3)[keep_e][:, :d])
        self.viscosity = f["viscosity"].copy()
        self.density = f["density"].copy()
        if two_phase:
            x = self.coords
            s = (x[:, d - 1] - 0.5 * x[:, d - 1].max()) / (0.25 * x[:, d - 1].max())
            self.density = (1.2 + 0.5 * (1.0 + np.tanh(s)) * 998.8) * (
                1.0 + 0.02 * rng.random(n))
        self.pressure = f["pressure"].copy()
        self.udiag = f["momentum_diag"].copy()
        self.tke = f["turbulent_ke"].copy()
        self.dflux = f["effective_viscosity_tke"].copy()
        self.mask = (rng.random(n) > 0.2).astype(np.float64)
        self.mdot = orc.mdot_edge(d, self.edges, self.coords, self.velocity, self.dpdx,
                                  self.density, self.pressure, self.udiag, self.area,
                                  1.0, 1.0)
        self.mvof = 0.3 * np.abs(self.mdot).mean() * rng.standard_normal(self.n_edges)
        self.pecfac = orc.peclet_edge(d, self.edges, self.coords, self.velocity,
                                      self.density, self.viscosity,
                                      orc.peclet("classic", 1.0))[1]
        self.source = rng.standard_normal((n, d))
        self.source_mask = (rng.random(n) > 0.3).astype(np.float64)
        self.efvm = 0.1 * rng.standard_normal(self.n_edges)
        self.vol = 0.5 + rng.random(n)

    def world(self):
        d = self.ndim
        w = R.World(d, self.n_nodes, self.edges)
        w.field("coordinates", R.NODE, self.coords, d)
        w.field("velocity", R.NODE, self.velocity, d)
        w.field("dudx", R.NODE, self.dudx, d * d)
        w.field("dpdx", R.NODE, self.dpdx, d)
        w.field("viscosity", R.NODE, self.viscosity, 1)
        w.field("density", R.NODE, self.density, 1)
        w.field("pressure", R.NODE, self.pressure, 1)
        w.field("momentum_diag", R.NODE, self.udiag, 1)
        w.field("turbulent_ke", R.NODE, self.tke, 1)
        w.field("dkdx", R.NODE, self.dkdx, d)
        w.field("effective_viscosity_tke", R.NODE, self.dflux, 1)
        w.field("abl_wall_no_slip_wall_func_node_mask", R.NODE, self.mask, 1)
        w.field("buoyancy_source", R.NODE, self.source, d)
        w.field("buoyancy_source_mask", R.NODE, self.source_mask, 1)
        w.field("edge_area_vector", R.EDGE, self.area, d)
        w.field("mass_flow_rate", R.EDGE, self.mdot, 1)
        w.field("mass_vof_balanced_flow_rate", R.EDGE, self.mvof, 1)
        w.field("peclet_factor", R.EDGE, self.pecfac, 1)
        w.field("edge_face_velocity_mag", R.EDGE, self.efvm, 1)
        w.field("dual_nodal_volume", R.NODE, self.vol, 1)
        w.field("peclet_number", R.EDGE, np.zeros(self.n_edges), 1)
        w.field("dpdx_new", R.NODE, np.zeros((self.n_nodes, d)), d)
        w.field("dudx_new", R.NODE, np.zeros((self.n_nodes, d * d)), d * d)
        return w


def ref_momentum(st, o, vof=False):
    w = st.world()
    w.option("divU", o["include_divu"])
    w.option("alpha:velocity", o["alpha"])
    w.option("alpha_upw:velocity", o["alpha_upw"])
    w.option("upw:velocity", o["ho_upwind"])
    w.option("relax:velocity", o["relax_fac"])
    w.option("limiter:velocity", 1.0 if o["use_limiter"] else 0.0)
    w.flags(has_vof=vof)
    return w.momentum()


def orc_momentum(st, o, vof=False):
    orc.set_num_threads(1)
    s = orc.RecordSink()
    orc.momentum_edge(st.ndim, st.edges, st.coords, st.velocity, st.dudx,
                      st.viscosity, st.density, st.mask, st.area, st.mdot,
                      st.pecfac, s, mass_vof=st.mvof if vof else None, **o)
    return s.get()


def ref_scalar(st, o, pec):
    w = st.world()
    w.option("alpha:turbulent_ke", o["alpha"])
    w.option("alpha_upw:turbulent_ke", o["alpha_upw"])
    w.option("upw:turbulent_ke", o["ho_upwind"])
    w.option("relax:turbulent_ke", o["relax_fac"])
    w.option("limiter:turbulent_ke", 1.0 if o["use_limiter"] else 0.0)
    w.peclet(*pec)
    return w.scalar("turbulent_ke", "dkdx", "effective_viscosity_tke")


def orc_scalar(st, o, pec):
    orc.set_num_threads(1)
    s = orc.RecordSink()
    form, a, b = pec
    orc.scalar_edge(st.ndim, st.edges, st.coords, st.velocity, st.tke, st.dkdx,
                    st.density, st.dflux, st.area, st.mdot, s,
                    pf=orc.peclet(form, a, b), **o)
    return s.get()


CONT_POINTS = [
    dict(dt=0.5, gamma1=1.5, noc_fac=1.0, interp_together=1.0, solve_incompressible=0.0),
    dict(dt=0.02, gamma1=1.0, noc_fac=0.0, interp_together=0.0, solve_incompressible=1.0),
    dict(dt=0.1, gamma1=1.5, noc_fac=1.0, interp_together=0.3, solve_incompressible=0.0),
]


def ref_cont_world(st, o, buoyancy, gcl):
    w = st.world()
    w.option("noc:pressure", o["noc_fac"])
    w.option("mdot_interp", o["interp_together"])
    w.option("solve_incompressible", o["solve_incompressible"])
    w.option("dt", o["dt"])
    w.option("gamma1", o["gamma1"])
    w.option("mesh_deformation", 1.0 if gcl else 0.0)
    w.flags(balanced_buoyancy=buoyancy, gravity=[0.3, -9.81, 0.7])
    return w


def orc_cont(st, o, buoyancy, gcl, sink):
    orc.set_num_threads(1)
    g = np.array([0.3, -9.81, 0.7])
    return orc.mdot_continuity_edge_ext(
        st.ndim, st.edges, st.coords, st.velocity, st.dpdx, st.density,
        st.pressure, st.udiag, st.area, noc_fac=o["noc_fac"],
        interp_together=o["interp_together"],
        gravity=g if buoyancy else None, source=st.source if buoyancy else None,
        source_mask=st.source_mask if buoyancy else None,
        edge_face_vel_mag=st.efvm if gcl else None, dt=o["dt"], gamma1=o["gamma1"],
        solve_incompressible=o["solve_incompressible"], sink=sink)


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a, b)


STATES = {}


def state(key):
    if key not in STATES:
        STATES[key] = {
            "3d": lambda: State(3),
            "3d-two-phase": lambda: State(3, two_phase=True, seed=5),
            "2d": lambda: State(2, dims=(7, 6, 1)),
        }[key]()
    return STATES[key]


# ------------------------------ live ---------------------------------------

@live
@pytest.mark.parametrize("key", ["3d", "2d"])
@pytest.mark.parametrize("o", MOM_POINTS, ids=lambda o: "a%(alpha)g-au%(alpha_upw)g-ho%(ho_upwind)g" % o)
def test_momentum_lambda_bitwise(key, o):
    st = state(key)
    lhs, rhs = ref_momentum(st, o)
    ol, orh = orc_momentum(st, o)
    assert st.n_edges > 50 and np.abs(lhs).max() > 0
    assert same_bits(lhs, ol) and same_bits(rhs, orh)


@live
@pytest.mark.parametrize("o", MOM_POINTS[:3], ids=["golds", "decks", "mixed"])
def test_momentum_vof_branch_bitwise(o):
    """realm_has_vof_: MomentumEdgeSolverAlg.C:88, 124-125, 174-192"""
    st = state("3d-two-phase")
    lhs, rhs = ref_momentum(st, o, vof=True)
    ol, orh = orc_momentum(st, o, vof=True)
    assert same_bits(lhs, ol) and same_bits(rhs, orh)
    # and the branch is not a no-op here
    l0, _ = ref_momentum(st, o, vof=False)
    assert not np.array_equal(l0, lhs)


@live
@pytest.mark.parametrize("key", ["3d", "2d"])
@pytest.mark.parametrize("pec", PECLETS, ids=lambda p: "%s-%g" % p[:2])
@pytest.mark.parametrize("o", SCAL_POINTS, ids=lambda o: "a%(alpha)g-au%(alpha_upw)g-ho%(ho_upwind)g" % o)
def test_scalar_lambda_bitwise(key, o, pec):
    st = state(key)
    lhs, rhs = ref_scalar(st, o, pec)
    ol, orh = orc_scalar(st, o, pec)
    assert same_bits(lhs, ol) and same_bits(rhs, orh)


@live
@pytest.mark.parametrize("key", ["3d", "2d"])
@pytest.mark.parametrize("buoyancy,gcl", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("o", CONT_POINTS, ids=["abl", "incompressible", "split"])
def test_continuity_and_mdot_lambdas_bitwise(key, o, buoyancy, gcl):
    st = state(key)
    lhs, rhs = ref_cont_world(st, o, buoyancy, gcl).continuity()
    s = orc.RecordSink()
    orc_cont(st, o, buoyancy, gcl, s)
    ol, orh = s.get()
    assert same_bits(lhs, ol) and same_bits(rhs, orh)
    # MdotEdgeAlg: the edge field it writes
    mdot = ref_cont_world(st, o, buoyancy, gcl).mdot()
    om = orc_cont(st, o, buoyancy, gcl, None)
    assert np.array_equal(mdot, om)
    assert np.abs(mdot).max() > 0


def ref_grads(st):
    return (st.world().nodal_grad("pressure", "dpdx_new"),
            st.world().nodal_grad("velocity", "dudx_new"))


def orc_grads(st):
    orc.set_num_threads(1)
    d = st.ndim
    return (orc.nodal_grad_edge(1, d, st.edges, st.pressure, st.area, st.vol, st.n_nodes),
            orc.nodal_grad_edge(d, d, st.edges, st.velocity, st.area, st.vol, st.n_nodes))


def ref_peclet_alg(st, pec):
    w = st.world()
    w.peclet(*pec)
    return w.peclet_alg()


def orc_peclet_alg(st, pec):
    orc.set_num_threads(1)
    return orc.peclet_edge(st.ndim, st.edges, st.coords, st.velocity, st.density,
                           st.viscosity, orc.peclet(*pec))


def orc_wall_dist(st):
    orc.set_num_threads(1)
    s = orc.RecordSink()
    orc.wall_dist_edge(st.ndim, st.edges, st.coords, st.area, s)
    return s.get()


@live
@pytest.mark.parametrize("key", ["3d", "2d"])
def test_nodal_grad_peclet_alg_wall_dist_bitwise(key):
    """NodalGradEdgeAlg (scalar -> vector, vector -> tensor; the serial edge
    order of the stand-in loop is the oracle's), MomentumEdgePecletAlg with all
    four blending functions, WallDistEdgeSolverAlg"""
    st = state(key)
    for got, want in zip(ref_grads(st), orc_grads(st)):
        assert np.abs(got).max() > 0 and same_bits(got, want)
    for pec in PECLETS:
        (pn, pf), (opn, opf) = ref_peclet_alg(st, pec), orc_peclet_alg(st, pec)
        assert np.array_equal(pn, opn) and np.array_equal(pf, opf), pec
    (lhs, rhs), (ol, orh) = st.world().wall_dist(), orc_wall_dist(st)
    assert same_bits(lhs, ol) and same_bits(rhs, orh)


@live
@pytest.mark.parametrize("key", ["3d", "2d"])
@pytest.mark.parametrize("states", [3, 2])
def test_node_kernels_bitwise(key, states):
    """ScalarMassBDFNodeKernel, MomentumMassBDFNodeKernel,
    ContinuityMassBDFNodeKernel (BDF2 with three states, BDF1 with two),
    WallDistNodeKernel of the reference (src/node_kernels/*.C: constructor,
    setup, execute per node) against the oracle's node kernels"""
    st = state(key)
    d, n = st.ndim, st.n_nodes
    rng = np.random.default_rng(17)
    q3 = [st.tke * (1.0 + 0.1 * rng.standard_normal(n)) for _ in range(2)] + [st.tke]
    rho3 = [st.density * (1.0 + 0.05 * rng.random(n)) for _ in range(2)] + [st.density]
    u3 = [st.velocity + 0.1 * rng.standard_normal((n, d)) for _ in range(2)] + [st.velocity]
    dnv3 = [st.vol * (1.0 + 0.02 * rng.random(n)) for _ in range(2)] + [st.vol]
    if states == 2:  # no NM1 state registered: the kernels alias it to N
        q3[0], rho3[0], u3[0] = q3[1], rho3[1], u3[1]
        dnv3[0] = dnv3[2]  # populate_dnv_states, two states: nm1 = np1
    dt, g1, g2, g3 = 0.25, 1.5, -2.0, 0.5
    w = st.world()
    for name, arrs, nc in (("turbulent_ke", q3, 1), ("density", rho3, 1),
                           ("velocity", u3, d), ("dual_nodal_volume", dnv3, 1)):
        w.field(name + "_n", R.NODE, arrs[1], nc)
        if states == 3:
            w.field(name + "_nm1", R.NODE, arrs[0], nc)
    w.option("dt", dt)
    w.option("gamma1", g1)
    w.option("gamma2", g2)
    w.option("gamma3", g3)
    nodes = np.arange(n, dtype=np.int32)
    orc.set_num_threads(1)

    def rec(f, *a):
        s = orc.RecordSink()
        f(*a, s)
        return s.get()

    lhs, rhs = w.node_kernel("scalar_mass", "turbulent_ke")
    ol, orh = rec(orc.scalar_mass_bdf_node, nodes, q3, rho3, dnv3, dt, g1, g2, g3)
    assert same_bits(lhs, ol) and same_bits(rhs, orh)
    lhs, rhs = w.node_kernel("momentum_mass")
    ol, orh = rec(orc.momentum_mass_bdf_node, d, nodes, u3, rho3, dnv3, st.dpdx,
                  dt, g1, g2, g3)
    assert same_bits(lhs, ol) and same_bits(rhs, orh)
    lhs, rhs = w.node_kernel("continuity_mass")
    ol, orh = rec(orc.continuity_mass_bdf_node, nodes, rho3, dnv3, dt, g1, g2, g3)
    assert same_bits(lhs, ol) and same_bits(rhs, orh)
    lhs, rhs = w.node_kernel("wall_dist")
    ol, orh = rec(orc.wall_dist_node, nodes, st.vol)
    assert same_bits(lhs, ol) and same_bits(rhs, orh)


@live
@pytest.mark.parametrize("seed", range(6))
def test_random_option_points_bitwise(seed):
    """30 random points of the option space per seed (momentum with and without
    VOF on the two-phase state, scalar with random blending functions)"""
    rng = np.random.default_rng(1000 + seed)
    st = state("3d-two-phase")
    for _ in range(5):
        o = dict(include_divu=float(rng.integers(0, 2)), alpha=float(rng.random()),
                 alpha_upw=float(rng.random()), ho_upwind=float(rng.random()),
                 relax_fac=float(0.3 + 0.7 * rng.random()),
                 use_limiter=bool(rng.integers(0, 2)))
        for vof in (False, True):
            lhs, rhs = ref_momentum(st, o, vof=vof)
            ol, orh = orc_momentum(st, o, vof=vof)
            assert same_bits(lhs, ol) and same_bits(rhs, orh), (o, vof)
        so = {k: o[k] for k in ("alpha", "alpha_upw", "ho_upwind", "relax_fac", "use_limiter")}
        pec = (("classic", float(rng.random()), 1.0) if rng.integers(0, 2)
               else ("tanh", float(10.0 ** rng.uniform(-1, 3)), float(10.0 ** rng.uniform(-1, 2))))
        lhs, rhs = ref_scalar(st, so, pec)
        ol, orh = orc_scalar(st, so, pec)
        assert same_bits(lhs, ol) and same_bits(rhs, orh), (so, pec)


# ------------------------------ fixture -------------------------------------

def _fixture():
    return np.load(os.path.join(HERE, "golden", "reference_edge_runs.npz"))


def _fixture_state(z):
    st = State.__new__(State)
    st.ndim = int(z["ndim"])
    for k in ("edges", "coords", "velocity", "dudx", "dpdx", "dkdx", "area",
              "viscosity", "density", "pressure", "udiag", "tke", "dflux", "mask",
              "mdot", "mvof", "pecfac", "source", "source_mask", "efvm", "vol"):
        setattr(st, k, np.ascontiguousarray(z["in_" + k]))
    st.n_nodes, st.n_edges = len(st.coords), len(st.edges)
    return st


def test_fixture_momentum_scalar_continuity_mdot():
    """inputs and the reference's outputs are committed: the oracle reproduces
    the reference's bits wherever the tests run"""
    z = _fixture()
    st = _fixture_state(z)
    n = 0
    for i, o in enumerate(MOM_POINTS):
        for vof in (False, True):
            ol, orh = orc_momentum(st, o, vof=vof)
            tag = "mom%d%s" % (i, "v" if vof else "")
            assert same_bits(ol, z[tag + "_lhs"]) and same_bits(orh, z[tag + "_rhs"]), tag
            n += 1
    for i, o in enumerate(SCAL_POINTS):
        for j, pec in enumerate(PECLETS):
            ol, orh = orc_scalar(st, o, pec)
            tag = "scal%d_%d" % (i, j)
            assert same_bits(ol, z[tag + "_lhs"]) and same_bits(orh, z[tag + "_rhs"]), tag
            n += 1
    for i, o in enumerate(CONT_POINTS):
        for j, (bu, gcl) in enumerate([(False, False), (True, True)]):
            s = orc.RecordSink()
            orc_cont(st, o, bu, gcl, s)
            ol, orh = s.get()
            tag = "cont%d_%d" % (i, j)
            assert same_bits(ol, z[tag + "_lhs"]) and same_bits(orh, z[tag + "_rhs"]), tag
            assert np.array_equal(orc_cont(st, o, bu, gcl, None), z[tag + "_mdot"]), tag
            n += 1
    assert n == 2 * len(MOM_POINTS) + len(SCAL_POINTS) * len(PECLETS) + 2 * len(CONT_POINTS)
    gs, gv = orc_grads(st)
    assert same_bits(gs, z["grad_scalar"]) and same_bits(gv, z["grad_vector"])
    for j, pec in enumerate(PECLETS):
        pn, pf = orc_peclet_alg(st, pec)
        assert np.array_equal(pn, z["pecalg%d_number" % j])
        assert np.array_equal(pf, z["pecalg%d_factor" % j])
    ol, orh = orc_wall_dist(st)
    assert same_bits(ol, z["walldist_lhs"]) and same_bits(orh, z["walldist_rhs"])


@live
def test_fixture_is_what_the_reference_build_gives_now():
    z = _fixture()
    st = _fixture_state(z)
    lhs, rhs = ref_momentum(st, MOM_POINTS[2], vof=True)
    assert same_bits(lhs, z["mom2v_lhs"]) and same_bits(rhs, z["mom2v_rhs"])
    lhs, rhs = ref_scalar(st, SCAL_POINTS[2], PECLETS[2])
    assert same_bits(lhs, z["scal2_2_lhs"]) and same_bits(rhs, z["scal2_2_rhs"])


def write_fixture(path):
    """called by tests/golden/extract_reference_runs.py"""
    st = State(3, dims=(2, 2, 2), warp=0.15, two_phase=True, seed=7)
    out = {"ndim": np.int32(3)}
    for k in ("edges", "coords", "velocity", "dudx", "dpdx", "dkdx", "area",
              "viscosity", "density", "pressure", "udiag", "tke", "dflux", "mask",
              "mdot", "mvof", "pecfac", "source", "source_mask", "efvm", "vol"):
        out["in_" + k] = getattr(st, k)
    for i, o in enumerate(MOM_POINTS):
        for vof in (False, True):
            tag = "mom%d%s" % (i, "v" if vof else "")
            out[tag + "_lhs"], out[tag + "_rhs"] = ref_momentum(st, o, vof=vof)
    for i, o in enumerate(SCAL_POINTS):
        for j, pec in enumerate(PECLETS):
            tag = "scal%d_%d" % (i, j)
            out[tag + "_lhs"], out[tag + "_rhs"] = ref_scalar(st, o, pec)
    for i, o in enumerate(CONT_POINTS):
        for j, (bu, gcl) in enumerate([(False, False), (True, True)]):
            tag = "cont%d_%d" % (i, j)
            out[tag + "_lhs"], out[tag + "_rhs"] = ref_cont_world(st, o, bu, gcl).continuity()
            out[tag + "_mdot"] = ref_cont_world(st, o, bu, gcl).mdot()
    out["grad_scalar"], out["grad_vector"] = ref_grads(st)
    for j, pec in enumerate(PECLETS):
        out["pecalg%d_number" % j], out["pecalg%d_factor" % j] = ref_peclet_alg(st, pec)
    out["walldist_lhs"], out["walldist_rhs"] = st.world().wall_dist()
    np.savez_compressed(path, **out)
    return st.n_nodes, st.n_edges
