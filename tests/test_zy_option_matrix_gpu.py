"""The general-option path of the scalar and momentum kernels on the GPU:
upwinding / blending / limiter / divU options away from both the reference
golds' zeros and the decks' defaults (tests/test_option_matrix_cpu.py checks
the same physics header on the CPU and pins the oracle's branches by
properties).  Through the C ABI against the oracle, 1e-12 of each entry's sum of
|contributions|.  Needs a B200: `pytest -m gpu`."""
import numpy as np
import pytest

import oracle_py as orc
import parity_util as pu

pytestmark = pytest.mark.gpu

POINTS = [
    dict(alpha=0.4, alpha_upw=0.6, ho_upwind=0.5, relax_fac=0.7, use_limiter=True),
    dict(alpha=1.0, alpha_upw=1.0, ho_upwind=0.0, relax_fac=1.0, use_limiter=False),
    dict(alpha=1.0, alpha_upw=0.0, ho_upwind=1.0, relax_fac=0.9, use_limiter=True),
    dict(alpha=0.0, alpha_upw=1.0, ho_upwind=0.5, relax_fac=0.7, use_limiter=False),
]
_ID = lambda o: "a%(alpha)g-au%(alpha_upw)g-ho%(ho_upwind)g" % o  # noqa: E731


@pytest.fixture(scope="module")
def P():
    return pu.pkg()


@pytest.fixture(scope="module")
def ctx(P):
    c = P.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def setup(P, ctx):
    case = pu.Case(dims=(8, 7, 6))
    mesh = case.box.make_mesh(ctx, tile_nodes=64)
    pu.upload_state(P, mesh, case)
    omdot = case.oracle_mdot()
    opec = case.oracle_pecfac(orc.peclet("classic", 1.0))
    mesh.upload("mass_flow_rate", omdot)
    mesh.upload("peclet_factor", opec)
    yield case, mesh, omdot, opec
    mesh.close()


@pytest.mark.parametrize("mode", ["segmented", "atomic"])
@pytest.mark.parametrize("o", POINTS, ids=_ID)
def test_scalar_options_vs_oracle(P, setup, o, mode):
    case, mesh, omdot, opec = setup
    f, b = case.fields, case.box
    g = case.oracle_graph()
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    ls.set_scatter_mode(P.NW_SCATTER_SEGMENTED if mode == "segmented"
                        else P.NW_SCATTER_ATOMIC)
    for pec in (("classic", 1.0), ("tanh", 2.0, 1.0)):
        ls.zeroSystem()
        ls.assemble_scalar_edge("turbulent_ke", "dkdx", "effective_viscosity_tke",
                                pf=P.peclet_fn(*pec), **o)
        vals, rhs = ls.values()
        s = orc.HypreSink(g, b.hid)
        orc.scalar_edge(3, case.edges, b.coords, f["velocity"], f["turbulent_ke"],
                        f["dkdx"], f["density"], f["effective_viscosity_tke"],
                        case.area, omdot, s, pf=orc.peclet(*pec), **o)
        ov, orhs = s.get()
        av, arhs = s.get_abs()
        assert pu.scaled_err(vals, ov, av) < 1
        assert pu.scaled_err(rhs, orhs, arhs) < 1
    ls.close()


@pytest.mark.parametrize("mode", ["segmented", "atomic"])
@pytest.mark.parametrize("system", ["uvw", "monolithic"])
@pytest.mark.parametrize("divu", [0.0, 1.0])
@pytest.mark.parametrize("o", POINTS, ids=_ID)
def test_momentum_options_vs_oracle(P, setup, o, divu, system, mode):
    case, mesh, omdot, opec = setup
    f, b = case.fields, case.box
    uvw = system == "uvw"
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE_UVW if uvw else P.NW_LINSYS_HYPRE, 3)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    ls.set_scatter_mode(P.NW_SCATTER_SEGMENTED if mode == "segmented"
                        else P.NW_SCATTER_ATOMIC)
    g = case.oracle_graph(num_dof=1 if uvw else 3)
    s = orc.HypreSink(g, b.hid, uvw_ndim=3 if uvw else 0)
    orc.momentum_edge(3, case.edges, b.coords, f["velocity"], f["dudx"],
                      f["viscosity"], f["density"],
                      f["abl_wall_no_slip_wall_func_node_mask"], case.area,
                      omdot, opec, s, include_divu=divu, **o)
    ov, orhs = s.get()
    av, arhs = s.get_abs()
    for fuse in ((False, True) if uvw else (False,)):
        ls.zeroSystem()
        ls.assemble_momentum_edge("viscosity", include_divu=divu,
                                  fuse_peclet=fuse,
                                  pf=P.peclet_fn("classic", 1.0), **o)
        vals, rhs = ls.values()
        assert pu.scaled_err(vals, ov, av) < 1
        assert pu.scaled_err(rhs, orhs, arhs) < 1
    ls.close()
