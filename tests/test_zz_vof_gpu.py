"""realm_has_vof_ branch of MomentumEdgeSolverAlg
(src/edge_kernels/MomentumEdgeSolverAlg.C:88, 124-125, 174-192) through the C
ABI on the GPU: mdot = mass_flow_rate + mass_vof_balanced_flow_rate, upwinding
factors from the density jump of the edge.  All four kernels (UVW / monolithic,
tile / atomic) against the oracle on a two-phase density field.  The same
physics header is checked on the CPU in tests/test_option_matrix_cpu.py.
Needs a B200: `pytest -m gpu`."""
import numpy as np
import pytest

import oracle_py as orc
import parity_util as pu

pytestmark = pytest.mark.gpu

OPTS = [dict(include_divu=0.0, alpha=0.0, alpha_upw=1.0, ho_upwind=1.0,
             relax_fac=0.7, use_limiter=True),
        dict(include_divu=1.0, alpha=0.4, alpha_upw=0.6, ho_upwind=0.5,
             relax_fac=1.0, use_limiter=False)]


@pytest.fixture(scope="module")
def P():
    return pu.pkg()


@pytest.fixture(scope="module")
def ctx(P):
    c = P.Context(0)
    yield c
    c.close()


def _two_phase_case():
    c = pu.Case(dims=(9, 8, 8))
    x = c.box.coords.reshape(-1, 3)
    rng = np.random.default_rng(5)
    s = (x[:, 2] - 0.5 * x[:, 2].max()) / (0.25 * x[:, 2].max())
    rho = 1.2 + 0.5 * (1.0 + np.tanh(s)) * 998.8
    c.fields["density"] = rho * (1.0 + 0.02 * rng.random(rho.size))
    return c


@pytest.mark.parametrize("o", OPTS, ids=["deck", "mixed"])
@pytest.mark.parametrize("mode", ["segmented", "atomic"])
@pytest.mark.parametrize("system", ["uvw", "monolithic"])
def test_vof_momentum_vs_oracle(P, ctx, system, mode, o):
    case = _two_phase_case()
    f, b = case.fields, case.box
    mesh = b.make_mesh(ctx, tile_nodes=64)
    pu.upload_state(P, mesh, case)
    omdot = case.oracle_mdot()
    opec = case.oracle_pecfac(orc.peclet("classic", 1.0))
    rng = np.random.default_rng(11)
    mvof = 0.3 * np.abs(omdot).mean() * rng.standard_normal(case.n_edges)
    mesh.upload("mass_flow_rate", omdot)
    mesh.upload("peclet_factor", opec)
    uvw = system == "uvw"
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE_UVW if uvw else P.NW_LINSYS_HYPRE, 3)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    ls.set_scatter_mode(P.NW_SCATTER_SEGMENTED if mode == "segmented"
                        else P.NW_SCATTER_ATOMIC)
    ls.zeroSystem()
    # the edge field the branch reads must exist (get_field_ordinal would throw)
    with pytest.raises(P.NwError):
        ls.assemble_momentum_edge("viscosity", has_vof=True, **o)
    mesh.put("mass_vof_balanced_flow_rate", P.NW_EDGE, mvof)
    ls.zeroSystem()
    ls.assemble_momentum_edge("viscosity", has_vof=True, **o)
    vals, rhs = ls.values()
    g = case.oracle_graph(num_dof=1 if uvw else 3)
    s = orc.HypreSink(g, b.hid, uvw_ndim=3 if uvw else 0)
    orc.momentum_edge(3, case.edges, b.coords, f["velocity"], f["dudx"],
                      f["viscosity"], f["density"],
                      f["abl_wall_no_slip_wall_func_node_mask"], case.area,
                      omdot, opec, s, mass_vof=mvof, **o)
    ov, orhs = s.get()
    av, arhs = s.get_abs()
    # tolerance scales of the VOF branch (parity_util.vof_scales: the device's
    # erf is not bit-identical to the host's, and with alphaUpw = 1 one ulp of
    # it is the whole error of entries that are cancellation remainders; sized
    # on the CPU with the fma / +-2 ulp builds, tests/test_option_matrix_cpu.py)
    lsc, rsc = pu.vof_scales(case, g, omdot + mvof, ov, av, arhs, 1 if uvw else 3)
    assert pu.scaled_err(vals, ov, lsc) < 1
    assert pu.scaled_err(rhs.ravel() if not uvw else rhs,
                         orhs.ravel() if not uvw else orhs, rsc) < 1
    # and the branch is not a no-op on this case
    s0 = orc.HypreSink(g, b.hid, uvw_ndim=3 if uvw else 0)
    orc.momentum_edge(3, case.edges, b.coords, f["velocity"], f["dudx"],
                      f["viscosity"], f["density"],
                      f["abl_wall_no_slip_wall_func_node_mask"], case.area,
                      omdot, opec, s0, **o)
    assert pu.scaled_err(vals, s0.get()[0], lsc) > 1e6
    # has_vof off again: the ordinary kernels, the ordinary answer
    ls.zeroSystem()
    ls.assemble_momentum_edge("viscosity", **o)
    v0, r0 = ls.values()
    assert pu.scaled_err(v0, s0.get()[0], s0.get_abs()[0]) < 1
    assert pu.scaled_err(r0, s0.get()[1], s0.get_abs()[1]) < 1
    ls.close()
    mesh.close()
