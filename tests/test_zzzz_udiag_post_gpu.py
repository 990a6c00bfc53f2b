"""nw_momentum_diag_post_process on the GPU: the consumer of extract_diagonal
(MomentumEquationSystem::assemble_and_solve after the solve,
src/LowMachEquationSystem.C:2759-2821) against its restatement
(parity_util.momentum_diag_post_process), on a plain and on a laterally periodic
box, fed by the momentum tile kernel's extracted diagonal.  Needs a B200:
`pytest -m gpu`.

Added after the round's GPU budget was spent: collected last, so that it cannot
shadow the tests that have already run on a B200."""
import numpy as np
import pytest

import oracle_py as orc
import parity_util as pu

pytestmark = pytest.mark.gpu
EPS = 2.2e-16


@pytest.mark.parametrize("periodic", [(False, False), (True, True)],
                         ids=["box", "periodic"])
def test_momentum_diag_post_process(periodic):
    P = pu.pkg()
    ctx = P.Context(0)
    try:
        case = pu.Case(dims=(9, 8, 6), periodic=periodic)
        f, b = case.fields, case.box
        mesh = b.make_mesh(ctx, tile_nodes=56)
        pu.upload_state(P, mesh, case)
        omdot = case.oracle_mdot()
        opec = case.oracle_pecfac(orc.peclet("classic", 1.0))
        mesh.upload("mass_flow_rate", omdot)
        mesh.upload("peclet_factor", opec)
        g = case.oracle_graph()
        ud = np.zeros(case.n_nodes)
        pu.oracle_momentum(case, g, omdot, opec, uvw=True, udiag=ud)
        mesh.register("udiag", P.NW_NODE, 1)
        mesh.fill("udiag", 0.0)   # TSCALE_UDIAGINV resets to 0 (:2741-2751)
        ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE_UVW, 3)
        ls.buildEdgeToNodeGraph()
        ls.finalizeLinearSystem()
        ls.zeroSystem()
        ls.assemble_momentum_edge("viscosity", diag_field="udiag", **pu.MOM_OPTS)
        got0 = mesh.download("udiag")
        alpha_u = pu.MOM_OPTS["relax_fac"]
        mesh.momentum_diag_post_process(pu.DT, pu.GAMMA1, alpha_u, udiag="udiag")
        got = mesh.download("udiag")
        # (1) the step itself, from the device's own input: IEEE divide and
        # separately rounded products on both sides
        ref = pu.momentum_diag_post_process(
            got0, f["density"], f["dual_nodal_volume"], b.hid, b.own_hid,
            pu.DT, pu.GAMMA1, alpha_u)
        pts = pu.GAMMA1 / pu.DT
        rv = f["density"] * f["dual_nodal_volume"]
        bar = 4 * EPS * (np.abs(got0 / rv).max() + pts)
        assert np.all(np.abs(got - ref) <= bar), float(np.max(np.abs(got - ref)) / bar)
        # (2) the chain extract_diagonal -> post-processing against the oracle's
        # diagonal sums: the assembly's bar carried through d/d(udiag)
        ref2 = pu.momentum_diag_post_process(
            ud, f["density"], f["dual_nodal_volume"], b.hid, b.own_hid,
            pu.DT, pu.GAMMA1, alpha_u)
        sl = np.nonzero(b.own_hid != b.hid)[0]
        assert (len(sl) > 0) == any(periodic)
        own_to_node = {int(h): n for n, h in enumerate(b.own_hid)}
        slope = alpha_u / (f["density"] * f["dual_nodal_volume"])
        scale = (np.abs(ud) + np.max(np.abs(ud)) * 1e-2) * slope \
            + np.abs(ref2) + pu.GAMMA1 / pu.DT
        for n in sl:   # a slave's value, and with it its bar, is its master's
            scale[n] = scale[own_to_node[int(b.hid[n])]]
        assert pu.scaled_err(got, ref2, scale) < 1
        # periodic slaves carry their master's value, bit for bit
        for n in sl:
            assert got[n] == got[own_to_node[int(b.hid[n])]]
        # a second call works on the processed field (no hidden state)
        mesh.momentum_diag_post_process(pu.DT, pu.GAMMA1, alpha_u, udiag="udiag")
        ref3 = pu.momentum_diag_post_process(
            got, f["density"], f["dual_nodal_volume"], b.hid, b.own_hid,
            pu.DT, pu.GAMMA1, alpha_u)
        got3 = mesh.download("udiag")
        assert np.all(np.abs(got3 - ref3) <= 4 * EPS * (np.abs(got / rv).max() + pts))
        ls.close()
        mesh.close()
    finally:
        ctx.close()
