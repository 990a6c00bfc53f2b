"""The device geometry kernels (nw_geometry_interior_hex8 / _tet4 / _pyr5 /
_wed6 / _quad4) against outputs of the reference's own master elements
(tests/golden/reference_runs.json; see tests/test_reference_runs.py for what
the fixture is and how it is assembled).  Needs a B200: `pytest -m gpu`.

Added after the round's GPU budget was spent: collected last, so that it cannot
shadow the tests that have already run on a B200."""
import numpy as np
import pytest

import parity_util as pu
import test_reference_runs as T

pytestmark = pytest.mark.gpu
EPS = 2.2e-16


@pytest.mark.parametrize("topo", T.TOPOS_3D + ["quad"])
def test_device_geometry_vs_reference_master_elements(topo):
    P = pu.pkg()
    m, coords, conn, edges, dnv, ev, area = T._block(topo)
    nd, npe = m["ndim"], m["nodes_per_element"]
    ctx = P.Context(0)
    try:
        mesh = P.Mesh(ctx, nd, edges, np.arange(len(coords), dtype=np.int64), coords)
        mesh.register("dual_nodal_volume", P.NW_NODE, 1)
        mesh.register("edge_area_vector", P.NW_EDGE, nd)
        mesh.fill("dual_nodal_volume", 0.0)
        mesh.fill("edge_area_vector", 0.0)
        mesh.geometry_interior(conn, dnv="dual_nodal_volume", area="edge_area_vector")
        got_dnv = mesh.download("dual_nodal_volume")
        got_area = mesh.download("edge_area_vector").reshape(-1, nd)
        mesh.close()
    finally:
        ctx.close()
    # The oracle and the CPU build of the product header reproduce these values
    # bit for bit; on the device nvcc contracts a*b+c, which moves single ulps of
    # the TERMS.  The volume formulas sum products of nd absolute coordinates
    # (Grandy's hex volume: triple products), the area vectors products of
    # nd - 1, so the bar is a few dozen eps |x|_max^nd resp. eps |x|_max^(nd-1)
    # per element (measured with the contracted CPU build for Tet4 / Wed6 /
    # Pyr5: 0.13 and 0.01 of one such unit).
    xm = np.abs(coords).reshape(len(conn), npe, nd).max(axis=(1, 2))
    vtol = 48.0 * EPS * np.repeat(xm, npe) ** nd
    assert np.all(np.abs(got_dnv - dnv) <= vtol), (
        topo, float(np.max(np.abs(got_dnv - dnv) / vtol)))
    atol = 48.0 * EPS * np.repeat(xm, len(edges) // len(conn))[:, None] ** (nd - 1)
    assert np.all(np.abs(got_area - area) <= atol), (
        topo, float(np.max(np.abs(got_area - area) / atol)))
