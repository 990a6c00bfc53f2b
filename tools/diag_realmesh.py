"""diagnostic: momentum UVW lhs on a reference mixed-element mesh, where does the
worst scaled error sit (tile vs atomic scatter)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import parity_util as pu, oracle_py as orc
P = pu.pkg(); ctx = P.Context(0)
case = pu.RealMeshCase("multiElemTypeCylinder")
mesh = case.box.make_mesh(ctx)
pu.upload_state(P, mesh, case)
g = case.oracle_graph()
omdot = case.oracle_mdot(); opec = case.oracle_pecfac(orc.peclet("classic", 1.0))
mesh.upload("mass_flow_rate", omdot); mesh.upload("peclet_factor", opec)
o = pu.oracle_momentum(case, g, omdot, opec, uvw=True)
ov, orhs = o.get(); av, arhs = o.get_abs()
for mode in (0, 1):
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE_UVW, 3)
    ls.set_scatter_mode(mode); ls.buildEdgeToNodeGraph(); ls.finalizeLinearSystem(); ls.zeroSystem()
    ls.assemble_momentum_edge("viscosity", **pu.MOM_OPTS)
    vals, rhs = ls.values()
    s = np.maximum(np.maximum(av, np.abs(ov)), 1e-300)
    err = np.abs(vals - ov) / (1e-12 * s)
    idx = np.argsort(-err)[:6]
    gr = ls.graph()
    print("mode", mode, "worst", err[idx])
    for k in idx:
        print("   k", int(k), "row", int(gr["rows"][k]), "col", int(gr["cols"][k]), "got %.17g ref %.17g abs %.3g" % (vals[k], ov[k], av[k]),
              "diag" if gr["rows"][k] == gr["cols"][k] else "off")
    print("   rhs worst", pu.scaled_err(rhs, orhs, arhs), " n(err>1):", int((err > 1).sum()), "of", err.size)
    ls.close()
# per-edge magnitudes at the worst entry's row
k = idx[0]; r = int(gr["rows"][k])
e_at = np.flatnonzero((case.edges[:, 0] == r) | (case.edges[:, 1] == r))
print("edges at row", r, len(e_at), "mdot", omdot[e_at][:8], "pec", opec[e_at][:8])
