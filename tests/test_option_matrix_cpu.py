"""The upwinding / blending / limiter / divU branches of the scalar and momentum
edge kernels (src/edge_kernels/ScalarEdgeSolverAlg.C:141-204,
src/edge_kernels/MomentumEdgeSolverAlg.C:145-310).

The reference's own golds pin only alpha = alpha_upw = upw = 0 (every
edge-kernel unit test sets alphaUpwMap_ = upwMap_ = 0:
unit_tests/edge_kernels/UnitTestMomentumAdvDiffEdge.C:243-244,
UnitTestScalarAdvDiffEdge.C:173-174); no in-tree known-answer test runs the
other branches, so they are pinned here from two sides:

 (1) properties of the scheme the reference's lines implement, checked on the
     ORACLE with an answer derived independently of it (this file's numpy
     lines, not the oracle's code):
       * Newton consistency: with hoUpwind = 0 the advective + diffusive edge
         flux is linear in the nodal unknown, the lhs the kernel assembles is
         its exact Jacobian and rhs = -residual, so  A q + rhs = 0  on an
         orthogonal mesh (no non-orthogonal correction) with relaxFac = 1 --
         for every alpha, alpha_upw and Peclet blending;
       * linear exactness: for a linear field with its exact gradient the
         limited high-order extrapolations uIpL, uIpR equal the midpoint value,
         so the rhs does not depend on alpha, alpha_upw, hoUpwind = 1 or the
         Peclet factor;
       * the van-Leer limiter's closed form on its three regimes
         (include/edge_kernels/EdgeKernelUtils.h:18-24).
 (2) the PRODUCT physics header (csrc/edge_physics.h, replayed through the tile
     plans by tests/emul) against the oracle over a matrix of option values --
     the general path the CUDA kernels take whenever a deck leaves the
     defaults.
No GPU."""
import itertools

import numpy as np
import pytest

import parity_util as pu

orc = pu.orc


def _case(dims=(7, 6, 5), **kw):
    return pu.Case(dims=dims, **kw)


def _graph_sizes(g):
    return g.nnz_owned + g.nnz_shared, g.num_rows_owned + g.num_rows_shared


def _csr_matvec(g, vals, x):
    """y = A x for the oracle's hypre-style CSR (one rank: rows == local ids)"""
    y = np.zeros(g.num_rows_owned)
    np.add.at(y, g.rows[:g.nnz_owned] - g.i_lower, vals[:g.nnz_owned] * x[g.cols[:g.nnz_owned]])
    return y


# ---------------------------------------------------------------------------
# (1) properties pinning the oracle
# ---------------------------------------------------------------------------

def test_van_leer_closed_form():
    """EdgeKernelUtils.h:18-24: 2 (a b + |a b|) / ((a + b)^2 + eps)"""
    L = orc.lib()
    import ctypes as C
    if not hasattr(L, "orc_van_leer"):
        pytest.skip("oracle exports no orc_van_leer")
    L.orc_van_leer.restype = C.c_double
    L.orc_van_leer.argtypes = [C.c_double] * 3
    eps = 1e-16
    # opposite signs: product negative -> numerator exactly 0
    assert L.orc_van_leer(1.0, -2.0, eps) == 0.0
    # equal slopes: 4 a^2 / (4 a^2 + eps)
    for a in (1e-3, 1.0, 37.5):
        assert L.orc_van_leer(a, a, eps) == (2.0 * (a * a + abs(a * a))) / ((a + a) * (a + a) + eps)
        assert abs(L.orc_van_leer(a, a, eps) - 1.0) < 1e-9
    # r = a / b: limiter = 4 r / (1 + r)^2 (harmonic form), <= 1
    for a, b in ((1.0, 3.0), (0.2, 5.0), (-4.0, -0.5)):
        r = a / b
        assert abs(L.orc_van_leer(a, b, eps) - 4.0 * r / (1.0 + r) ** 2) < 1e-14
    # both zero: 0 / eps
    assert L.orc_van_leer(0.0, 0.0, eps) == 0.0


@pytest.mark.parametrize("alpha,alpha_upw,pec", [
    (0.0, 1.0, ("classic", 1.0)), (0.0, 0.0, ("classic", 1.0)),
    (0.35, 0.6, ("tanh", 2.0, 1.0)), (1.0, 1.0, ("tanh", 0.5, 3.0)),
    (0.8, 0.1, ("classic", 0.3))])
def test_scalar_newton_consistency(alpha, alpha_upw, pec):
    """hoUpwind = 0, relaxFac = 1: the assembled system satisfies
    A q + rhs = 0 row by row"""
    c = _case()
    f, b = c.fields, c.box
    g = c.oracle_graph()
    mdot = c.oracle_mdot()
    s = orc.HypreSink(g, b.hid)
    # dqdx = 0: the non-orthogonal correction (explicit, rhs only,
    # ScalarEdgeSolverAlg.C:136-139) vanishes on any mesh
    orc.scalar_edge(3, c.edges, b.coords, f["velocity"], f["turbulent_ke"],
                    np.zeros_like(f["dkdx"]), f["density"],
                    f["effective_viscosity_tke"], c.area, mdot, s, alpha=alpha,
                    alpha_upw=alpha_upw, ho_upwind=0.0, relax_fac=1.0,
                    use_limiter=True, pf=orc.peclet(*pec))
    vals, rhs = s.get()
    av, arhs = s.get_abs()
    q = f["turbulent_ke"].ravel()
    res = _csr_matvec(g, vals, q) + rhs[0][:g.num_rows_owned]
    scale = _csr_matvec(g, np.abs(vals), np.abs(q)) + np.abs(arhs[0][:g.num_rows_owned])
    assert np.max(np.abs(res) / np.maximum(scale, 1e-12 * scale.max())) < 1e-13


@pytest.mark.parametrize("alpha,alpha_upw,inc_divu", [
    (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), (0.4, 0.7, 0.0), (1.0, 0.3, 0.0)])
def test_momentum_newton_consistency(alpha, alpha_upw, inc_divu):
    """the same for the same-component part of momentum: with hoUpwind = 0 and
    dudx = 0 the viscous stress reduces to  -mu (du_i a.a/a.dx + du.a a_i/a.dx)
    (src/edge_kernels/MomentumEdgeSolverAlg.C:213-265 with gjui = 0), whose
    Jacobian is exactly the ND x ND block the kernel assembles (:275-310), so
    the monolithic system satisfies  A u + rhs = 0"""
    c = _case()
    f, b = c.fields, c.box
    g3 = c.oracle_graph(num_dof=3)
    mdot = c.oracle_mdot()
    pec = c.oracle_pecfac(orc.peclet("tanh", 2.0, 1.0))
    s = orc.HypreSink(g3, b.hid)
    zero_dudx = np.zeros_like(f["dudx"])
    mask = np.ones_like(f["abl_wall_no_slip_wall_func_node_mask"])
    orc.momentum_edge(3, c.edges, b.coords, f["velocity"], zero_dudx,
                      f["viscosity"], f["density"], mask, c.area, mdot, pec, s,
                      include_divu=inc_divu, alpha=alpha, alpha_upw=alpha_upw,
                      ho_upwind=0.0, relax_fac=1.0, use_limiter=False)
    vals, rhs = s.get()
    av, arhs = s.get_abs()
    u = f["velocity"].ravel()
    n = g3.num_rows_owned
    res = _csr_matvec(g3, vals, u) + rhs.ravel()[:n]
    scale = _csr_matvec(g3, np.abs(vals), np.abs(u)) + np.abs(arhs.ravel()[:n])
    assert np.max(np.abs(res) / np.maximum(scale, 1e-12 * scale.max())) < 1e-13


def test_linear_field_makes_the_rhs_independent_of_the_blending():
    """q = c0 + c.x with dqdx = c: qIpL = qIpR = midpoint value whatever the
    limiter returns (it multiplies an extrapolation that is already exact only
    if it is 1: for a linear field dqML = dqMR = dq, van_leer = 1 - O(eps)), so
    the rhs is that of the central scheme for every alpha / alpha_upw / Peclet
    function"""
    c = _case()
    f, b = c.fields, c.box
    g = c.oracle_graph()
    mdot = c.oracle_mdot()
    cvec = np.array([0.3, -1.1, 0.7])
    q = 2.0 + b.coords.reshape(-1, 3) @ cvec
    dq = np.tile(cvec, (c.n_nodes, 1))

    def rhs_of(alpha, alpha_upw, ho, lim, pf):
        s = orc.HypreSink(g, b.hid)
        orc.scalar_edge(3, c.edges, b.coords, f["velocity"], q, dq,
                        f["density"], f["effective_viscosity_tke"], c.area,
                        mdot, s, alpha=alpha, alpha_upw=alpha_upw, ho_upwind=ho,
                        relax_fac=0.8, use_limiter=lim, pf=pf)
        return s.get()[1][0], s.get_abs()[1][0]

    ref, aref = rhs_of(0.0, 0.0, 0.0, False, orc.peclet("classic", 1.0))
    for alpha, au, lim, pf in ((0.0, 1.0, True, orc.peclet("classic", 1.0)),
                               (0.6, 0.4, False, orc.peclet("tanh", 2.0, 1.0)),
                               (1.0, 1.0, True, orc.peclet("tanh", 1.0, 0.2))):
        got, _ = rhs_of(alpha, au, 1.0, lim, pf)
        # the limiter is 1 - eps / (4 dq^2): a relative 1e-16 / dq^2 -- below
        # 1e-12 of the term for the |dq| > 1e-2 edges of this mesh
        assert np.max(np.abs(got - ref) / (aref + 1e-300)) < 1e-11


# ---------------------------------------------------------------------------
# (2) product physics header vs the oracle over the option matrix
# ---------------------------------------------------------------------------

_SCAL_MATRIX = [
    dict(alpha=a, alpha_upw=au, ho_upwind=ho, relax_fac=rf, use_limiter=lim)
    for (a, au, ho), rf, lim in itertools.product(
        ((0.0, 1.0, 1.0), (0.0, 0.0, 0.0), (0.4, 0.6, 0.5), (1.0, 1.0, 0.0),
         (0.0, 1.0, 0.5), (1.0, 0.0, 1.0)),
        (1.0, 0.7), (False, True))]


FMA = pytest.mark.parametrize("fma", [False, pytest.param(True, marks=pytest.mark.skipif(
    not pu.host_has_fma(), reason="host CPU without FMA"))], ids=["plain", "fma"])


@FMA
@pytest.mark.parametrize("o", _SCAL_MATRIX,
                         ids=lambda o: "a%(alpha)g-au%(alpha_upw)g-ho%(ho_upwind)g-r%(relax_fac)g-l%(use_limiter)d" % o)
def test_scalar_option_matrix_product_header_vs_oracle(o, fma):
    """fma: the header compiled with a*b+c contracted into fused multiply-adds,
    as nvcc compiles it for the device -- the rounding the CUDA kernels see,
    held to the same 1e-12 of the entry's sum of |contributions|"""
    P = pu.pkg()
    c = _case(dims=(6, 5, 4))
    f, b = c.fields, c.box
    emu = pu.Emu(c, tile_nodes=40, fma=fma)
    emu.build_linsys(0, 1)
    g = c.oracle_graph()
    nnz, rows = _graph_sizes(g)
    mdot = c.oracle_mdot()
    for pec in (("classic", 1.0), ("tanh", 2.0, 1.0)):
        s = orc.HypreSink(g, b.hid)
        orc.scalar_edge(3, c.edges, b.coords, f["velocity"], f["turbulent_ke"],
                        f["dkdx"], f["density"], f["effective_viscosity_tke"],
                        c.area, mdot, s, pf=orc.peclet(*pec), **o)
        vals, rhs = emu.assemble(1, pu.SCAL_FIELDS, P.ScalarOpts(
            o["alpha"], o["alpha_upw"], o["ho_upwind"], o["relax_fac"],
            1 if o["use_limiter"] else 0, 1e-16, P.peclet_fn(*pec)),
            nnz, rows, 1, mdot=mdot)
        ov, orhs = s.get()
        av, arhs = s.get_abs()
        assert pu.scaled_err(vals, ov, av) < 1
        assert pu.scaled_err(rhs, orhs, arhs) < 1


_MOM_MATRIX = [
    dict(include_divu=dv, alpha=a, alpha_upw=au, ho_upwind=ho, relax_fac=rf,
         use_limiter=lim)
    for (a, au, ho), dv, (rf, lim) in itertools.product(
        ((0.0, 1.0, 1.0), (0.0, 0.0, 0.0), (0.4, 0.6, 0.5), (1.0, 1.0, 0.0),
         (1.0, 0.0, 1.0)),
        (0.0, 1.0), ((1.0, False), (0.7, True)))]


@FMA
@pytest.mark.parametrize("o", _MOM_MATRIX,
                         ids=lambda o: "dv%(include_divu)g-a%(alpha)g-au%(alpha_upw)g-ho%(ho_upwind)g-r%(relax_fac)g-l%(use_limiter)d" % o)
def test_momentum_option_matrix_product_header_vs_oracle(o, fma):
    """UVW (separate and fused Peclet factor) and monolithic 3-dof"""
    P = pu.pkg()
    c = _case(dims=(6, 5, 4))
    f, b = c.fields, c.box
    emu = pu.Emu(c, tile_nodes=40, fma=fma)
    emu.build_linsys(0, 1)
    g = c.oracle_graph()
    nnz, rows = _graph_sizes(g)
    mdot = c.oracle_mdot()
    pec = c.oracle_pecfac(orc.peclet("classic", 1.0))

    def oracle(graph, uvw):
        s = orc.HypreSink(graph, b.hid, uvw_ndim=3 if uvw else 0)
        orc.momentum_edge(3, c.edges, b.coords, f["velocity"], f["dudx"],
                          f["viscosity"], f["density"],
                          f["abl_wall_no_slip_wall_func_node_mask"], c.area,
                          mdot, pec, s, **o)
        return s

    def popts(fuse):
        return P.MomentumOpts(
            o["include_divu"], o["alpha"], o["alpha_upw"], o["ho_upwind"],
            o["relax_fac"], 1 if o["use_limiter"] else 0, 1e-16, fuse,
            P.peclet_fn("classic", 1.0), 1e-16, -1)

    s = oracle(g, True)
    ov, orhs = s.get()
    av, arhs = s.get_abs()
    for fuse in (0, 1):
        vals, rhs = emu.assemble(2, pu.MOM_FIELDS, popts(fuse), nnz, rows, 3,
                                 mdot=mdot, pecfac=pec)
        assert pu.scaled_err(vals, ov, av) < 1
        assert pu.scaled_err(rhs, orhs, arhs) < 1

    g3 = c.oracle_graph(num_dof=3)
    s3 = oracle(g3, False)
    vals, rhs = emu.assemble_mono(pu.MOM_FIELDS, popts(0), *_graph_sizes(g3),
                                  mdot=mdot, pecfac=pec)
    ov, orhs = s3.get()
    av, arhs = s3.get_abs()
    assert pu.scaled_err(vals, ov, av) < 1
    assert pu.scaled_err(rhs, orhs.ravel(), arhs.ravel()) < 1


# ---------------------------------------------------------------------------
# realm_has_vof_ (src/edge_kernels/MomentumEdgeSolverAlg.C:88, 124-125, 174-192)
# ---------------------------------------------------------------------------

def _two_phase_density(c, seed=5):
    """water / air with a band of intermediate densities: the density jump
    |rhoL - rhoR| / min spans 0 ... ~1000, erf(6 x) spans 0 ... 1"""
    x = c.box.coords.reshape(-1, 3)
    rng = np.random.default_rng(seed)
    s = (x[:, 2] - 0.5 * x[:, 2].max()) / max(1.0, 0.25 * x[:, 2].max())
    rho = 1.2 + 0.5 * (1.0 + np.tanh(s)) * 998.8
    # a few nodes with tiny relative differences (erf in its linear range)
    rho *= 1.0 + 0.02 * rng.random(rho.size)
    return rho


BUILDS = pytest.mark.parametrize("build", [
    False,
    pytest.param(True, marks=pytest.mark.skipif(not pu.host_has_fma(), reason="no FMA")),
    pytest.param("erf", marks=pytest.mark.skipif(not pu.host_has_fma(), reason="no FMA"))],
    ids=["plain", "fma", "fma+erf2ulp"])


@pytest.mark.parametrize("o", [
    dict(include_divu=0.0, alpha=0.0, alpha_upw=1.0, ho_upwind=1.0,
         relax_fac=0.7, use_limiter=True),
    dict(include_divu=1.0, alpha=0.4, alpha_upw=0.6, ho_upwind=0.5,
         relax_fac=1.0, use_limiter=False),
    dict(include_divu=0.0, alpha=0.0, alpha_upw=0.0, ho_upwind=0.0,
         relax_fac=1.0, use_limiter=False)],
    ids=["deck", "mixed", "gold"])
@BUILDS
def test_vof_branch_product_header_vs_oracle(o, build):
    """plain: the header compiled like the oracle (the plain 1e-12 of the sum of
    |contributions| holds); fma: contracted multiply-adds as on the device;
    fma+erf2ulp: additionally erf moved by up to +-2 ulp (CUDA's bound).  The
    last two are held to parity_util.vof_scales."""
    P = pu.pkg()
    c = _case(dims=(6, 5, 6))
    f, b = c.fields, c.box
    f["density"] = _two_phase_density(c)
    emu = pu.Emu(c, tile_nodes=40, fma=build)
    emu.build_linsys(0, 1)
    g = c.oracle_graph()
    nnz, rows = _graph_sizes(g)
    mdot = c.oracle_mdot()
    rng = np.random.default_rng(11)
    mvof = 0.3 * np.abs(mdot).mean() * rng.standard_normal(c.n_edges)
    pec = c.oracle_pecfac(orc.peclet("classic", 1.0))

    def oracle(graph, uvw, vof=True):
        s = orc.HypreSink(graph, b.hid, uvw_ndim=3 if uvw else 0)
        orc.momentum_edge(3, c.edges, b.coords, f["velocity"], f["dudx"],
                          f["viscosity"], f["density"],
                          f["abl_wall_no_slip_wall_func_node_mask"], c.area,
                          mdot, pec, s, mass_vof=mvof if vof else None, **o)
        return s

    def popts(fuse, vof=1):
        return P.MomentumOpts(
            o["include_divu"], o["alpha"], o["alpha_upw"], o["ho_upwind"],
            o["relax_fac"], 1 if o["use_limiter"] else 0, 1e-16, fuse,
            P.peclet_fn("classic", 1.0), 1e-16, -1, vof)

    s = oracle(g, True)
    ov, orhs = s.get()
    av, arhs = s.get_abs()
    lsc, rsc = pu.vof_scales(c, g, mdot + mvof, ov, av, arhs, 1)
    # the branch must matter on this case, else the comparison says nothing
    s0 = oracle(g, True, vof=False)
    assert pu.scaled_err(s0.get()[0], ov, lsc) > 1e6
    for fuse in (0, 1):
        vals, rhs = emu.assemble(2, pu.MOM_FIELDS, popts(fuse), nnz, rows, 3,
                                 mdot=mdot + mvof, pecfac=pec)
        assert pu.scaled_err(vals, ov, lsc) < 1
        assert pu.scaled_err(rhs, orhs, rsc) < 1
        if not build:  # same arithmetic as the oracle: the plain bar holds too
            assert pu.scaled_err(vals, ov, av) < 1
            assert pu.scaled_err(rhs, orhs, arhs) < 1
    g3 = c.oracle_graph(num_dof=3)
    s3 = oracle(g3, False)
    vals, rhs = emu.assemble_mono(pu.MOM_FIELDS, popts(0), *_graph_sizes(g3),
                                  mdot=mdot + mvof, pecfac=pec)
    ov, orhs = s3.get()
    av, arhs = s3.get_abs()
    lsc, rsc = pu.vof_scales(c, g3, mdot + mvof, ov, av, arhs, 3)
    assert pu.scaled_err(vals, ov, lsc) < 1
    assert pu.scaled_err(rhs, orhs.ravel(), rsc) < 1
    if not build:
        assert pu.scaled_err(vals, ov, av) < 1
        assert pu.scaled_err(rhs, orhs.ravel(), arhs.ravel()) < 1


def test_vof_branch_is_the_identity_for_uniform_density():
    """rhoL == rhoR: the jump is 0, erf(0) = 0, the factor exactly 1, alphaUpw
    and the Peclet factor unchanged -- with a zero mass_vof_balanced_flow_rate
    the VOF branch must return the non-VOF result bit for bit (oracle and
    product header; an answer that does not depend on either's VOF lines)"""
    P = pu.pkg()
    c = _case(dims=(5, 4, 4))
    f, b = c.fields, c.box
    f["density"] = np.full(c.n_nodes, 1.178037722969475)
    g = c.oracle_graph()
    nnz, rows = _graph_sizes(g)
    mdot = c.oracle_mdot()
    pec = c.oracle_pecfac(orc.peclet("classic", 1.0))
    o = dict(include_divu=1.0, alpha=0.4, alpha_upw=0.6, ho_upwind=0.5,
             relax_fac=0.7, use_limiter=True)
    out = []
    for mv in (None, np.zeros(c.n_edges)):
        s = orc.HypreSink(g, b.hid, uvw_ndim=3)
        orc.momentum_edge(3, c.edges, b.coords, f["velocity"], f["dudx"],
                          f["viscosity"], f["density"],
                          f["abl_wall_no_slip_wall_func_node_mask"], c.area,
                          mdot, pec, s, mass_vof=mv, **o)
        out.append(s.get())
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    emu = pu.Emu(c, tile_nodes=40)
    emu.build_linsys(0, 1)
    got = []
    for vof in (0, 1):
        got.append(emu.assemble(2, pu.MOM_FIELDS, P.MomentumOpts(
            o["include_divu"], o["alpha"], o["alpha_upw"], o["ho_upwind"],
            o["relax_fac"], 1, 1e-16, 0, P.peclet_fn("classic", 1.0), 1e-16,
            -1, vof), nnz, rows, 3, mdot=mdot, pecfac=pec))
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])


@FMA
def test_wall_mask_with_zeros_product_header_vs_oracle(fma):
    """abl_wall_no_slip_wall_func_node_mask with zeros (the synthetic state sets
    it to one everywhere): the masked viscous terms of MomentumEdgeSolverAlg.C:
    267-269 through the product header's default and general paths"""
    P = pu.pkg()
    c = _case(dims=(6, 5, 6))
    f, b = c.fields, c.box
    rng = np.random.default_rng(4)
    f["abl_wall_no_slip_wall_func_node_mask"] = (rng.random(c.n_nodes) > 0.3).astype(float)
    emu = pu.Emu(c, tile_nodes=40, fma=fma)
    emu.build_linsys(0, 1)
    g = c.oracle_graph()
    nnz, rows = _graph_sizes(g)
    mdot = c.oracle_mdot()
    pec = c.oracle_pecfac(orc.peclet("classic", 1.0))
    for o in (dict(include_divu=0.0, alpha=0.0, alpha_upw=1.0, ho_upwind=1.0,
                   relax_fac=0.7, use_limiter=True),
              dict(include_divu=1.0, alpha=0.4, alpha_upw=0.6, ho_upwind=0.5,
                   relax_fac=0.7, use_limiter=True)):
        s = orc.HypreSink(g, b.hid, uvw_ndim=3)
        orc.momentum_edge(3, c.edges, b.coords, f["velocity"], f["dudx"],
                          f["viscosity"], f["density"],
                          f["abl_wall_no_slip_wall_func_node_mask"], c.area, mdot,
                          pec, s, **o)
        ov, orh = s.get()
        av, arhs = s.get_abs()
        for fuse in (0, 1):
            po = P.MomentumOpts(o["include_divu"], o["alpha"], o["alpha_upw"],
                                o["ho_upwind"], o["relax_fac"], 1, 1e-16, fuse,
                                P.peclet_fn("classic", 1.0), 1e-16, -1, 0)
            vals, rhs = emu.assemble(2, pu.MOM_FIELDS, po, nnz, rows, 3, mdot=mdot,
                                     pecfac=pec)
            assert pu.scaled_err(vals, ov, av) < 1
            assert pu.scaled_err(rhs, orh, arhs) < 1


@FMA
def test_udiag_post_processing_product_header_vs_restatement(fma):
    """udiag_post_value (the arithmetic of nw_momentum_diag_post_process,
    src/LowMachEquationSystem.C:2783-2790) against the numpy restatement on the
    oracle's extracted diagonal: bit-identical in the plain build; the
    contracted build may fuse (tmp - pts) * alphaU + pts, one rounding less"""
    c = _case(dims=(7, 6, 5))
    f, b = c.fields, c.box
    g = c.oracle_graph()
    ud = np.zeros(c.n_nodes)
    pu.oracle_momentum(c, g, c.oracle_mdot(),
                       c.oracle_pecfac(orc.peclet("classic", 1.0)), uvw=True, udiag=ud)
    for alpha_u in (0.7, 1.0):
        ref = pu.momentum_diag_post_process(
            ud, f["density"], f["dual_nodal_volume"], b.hid, b.own_hid, pu.DT,
            pu.GAMMA1, alpha_u)
        got = np.ascontiguousarray(ud.copy())
        rho = np.ascontiguousarray(f["density"], dtype=np.float64)
        vol = np.ascontiguousarray(f["dual_nodal_volume"], dtype=np.float64)
        pu.emu_lib(fma).emu_udiag_post(
            c.n_nodes, got.ctypes.data, rho.ctypes.data, vol.ctypes.data,
            pu.GAMMA1 / pu.DT, alpha_u)
        if fma:
            tmp = ud / (rho * vol)
            assert np.all(np.abs(got - ref) <= 2.3e-16 * (np.abs(tmp) + pu.GAMMA1 / pu.DT))
        else:
            assert np.array_equal(got, ref)
        # what the step undoes: udiag = rho vol ((x - pts) / alphaU + pts) gives x back
        x = 3.0 + np.arange(c.n_nodes) * 1e-3
        back = pu.momentum_diag_post_process(
            rho * vol * ((x - pu.GAMMA1 / pu.DT) / alpha_u + pu.GAMMA1 / pu.DT),
            rho, vol, b.hid, b.own_hid, pu.DT, pu.GAMMA1, alpha_u)
        assert np.allclose(back, x, rtol=1e-14, atol=0)
