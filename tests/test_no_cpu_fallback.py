"""The product has no CPU path: on a host-only context (device -1: plan and
graph building only) every entry of the C ABI that would compute, move or read
device data returns NW_ERR_CUDA (2) with a message, and bad arguments are
refused with NW_ERR_ARG instead of being worked around.  Runs without a GPU;
the same is checked with a device present by
tests/test_gpu_parity.py::test_no_silent_fallback."""
import ctypes as C

import numpy as np
import pytest

import parity_util as pu


@pytest.fixture(scope="module")
def env():
    P = pu.pkg()
    case = pu.Case(dims=(4, 3, 3))
    ctx = P.Context(-1)
    mesh = case.box.make_mesh(ctx, tile_nodes=16)
    for name, rank, nc in (("velocity", P.NW_NODE, 3), ("dpdx", P.NW_NODE, 3),
                           ("dudx", P.NW_NODE, 9), ("density", P.NW_NODE, 1),
                           ("pressure", P.NW_NODE, 1), ("viscosity", P.NW_NODE, 1),
                           ("momentum_diag", P.NW_NODE, 1),
                           ("dual_nodal_volume", P.NW_NODE, 1),
                           ("q", P.NW_NODE, 1), ("dqdx", P.NW_NODE, 3),
                           ("edge_area_vector", P.NW_EDGE, 3),
                           ("mass_flow_rate", P.NW_EDGE, 1),
                           ("peclet_factor", P.NW_EDGE, 1)):
        mesh.register(name, rank, nc)
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    uvw = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE_UVW, 3)
    uvw.buildEdgeToNodeGraph()
    uvw.finalizeLinearSystem()
    yield P, case, mesh, ls, uvw
    uvw.close()
    ls.close()
    mesh.close()
    ctx.close()


def _refused(P, fn, code=2, needle="no CUDA device"):
    with pytest.raises(P.NwError) as e:
        fn()
    msg = str(e.value)
    assert ("nw error %d" % code) in msg and needle in msg, msg


def test_host_side_work_is_available(env):
    """plan, graph and slot map are host products and work without a device"""
    P, case, mesh, ls, uvw = env
    assert mesh.stats()["n_edges"] == case.n_edges
    g = ls.graph()
    assert len(g["rows"]) == len(g["cols"]) > 0
    slots, rows = ls.edge_slots()
    assert slots.shape[0] == case.n_edges


def test_every_compute_entry_refuses_without_device(env):
    P, case, mesh, ls, uvw = env
    n, ne = case.n_nodes, case.n_edges
    hexes = pu.box_hex_elements(case.box)
    tets = np.ascontiguousarray(hexes[:, [0, 1, 2, 5]])
    bdf = dict(dt=0.5, gammas=(1.5, -2.0, 0.5), rho=("density",) * 3,
               dnv=("dual_nodal_volume",) * 3)
    calls = {
        "nw_field_upload": lambda: mesh.upload("density", np.ones(n)),
        "nw_field_download": lambda: mesh.download("density"),
        "nw_field_fill": lambda: mesh.fill("density", 0.0),
        "nw_field_parallel_sum": lambda: mesh.parallel_sum("dual_nodal_volume"),
        "nw_field_copy_owned_to_shared": lambda: mesh.copy_owned_to_shared("density"),
        "nw_momentum_diag_post_process": lambda: mesh.momentum_diag_post_process(
            0.5, 1.5, 0.7),
        "nw_mdot_edge": lambda: mesh.mdot_edge(1.0, 1.0),
        "nw_mdot_edge_ext": lambda: mesh.mdot_edge_ext(mesh.extra_opts(
            edge_face_vel_mag="mass_flow_rate")),
        "nw_peclet_edge": lambda: mesh.peclet_edge(),
        "nw_nodal_grad_edge": lambda: mesh.nodal_grad_edge("pressure", "dpdx"),
        "nw_geometry_interior_hex8": lambda: mesh.geometry_interior(
            hexes, dnv="dual_nodal_volume", area="edge_area_vector"),
        "nw_geometry_interior_tet4": lambda: mesh.geometry_interior(
            tets, dnv="dual_nodal_volume"),
        "nw_geometry_interior_wed6": lambda: mesh.geometry_interior(
            hexes[:, :6], dnv="dual_nodal_volume"),
        "nw_geometry_interior_pyr5": lambda: mesh.geometry_interior(
            hexes[:, :5], dnv="dual_nodal_volume"),
        "nw_linsys_zero": lambda: ls.zeroSystem(),
        "nw_assemble_continuity_edge": lambda: ls.assemble_continuity_edge(**pu.CONT_OPTS),
        "nw_assemble_continuity_edge_ext": lambda: ls.assemble_continuity_edge_ext(
            mesh.extra_opts(edge_face_vel_mag="mass_flow_rate"), **pu.CONT_OPTS),
        "nw_assemble_scalar_edge": lambda: ls.assemble_scalar_edge(
            "q", "dqdx", "viscosity", **pu.SCAL_OPTS),
        "nw_assemble_momentum_edge": lambda: uvw.assemble_momentum_edge(
            "viscosity", **pu.MOM_OPTS),
        "nw_assemble_mass_bdf_node": lambda: ls.assemble_mass_bdf_node(
            P.NW_MASS_CONTINUITY, **bdf),
        "nw_assemble_wall_dist_edge": lambda: ls.assemble_wall_dist_edge(),
        "nw_assemble_wall_dist_node": lambda: ls.assemble_wall_dist_node(),
        "nw_linsys_reset_rows": lambda: ls.resetRows([0, 1]),
        "nw_linsys_apply_dirichlet_bcs": lambda: ls.applyDirichletBCs(
            "pressure", "q", [0, 1]),
        "nw_linsys_get_values": lambda: ls.values(),
        "nw_linsys_rhs_norm2": lambda: ls.rhs_norm2(),
        "nw_linsys_load_complete": lambda: ls.loadComplete(),
        "nw_linsys_write_preassembly_files": lambda: ls.write_preassembly_files(
            "/nonexistent-dir", "ContinuityEQS"),
    }
    for name, fn in calls.items():
        with pytest.raises(P.NwError) as e:
            fn()
        msg = str(e.value)
        assert "nw error 2" in msg and "no CUDA device" in msg, (name, msg)
        assert name in msg, (name, msg)
    # raw entries the binding wraps with torch tensors
    L = P.lib()
    rc = L.nw_linsys_sum_into(ls.h, 1, 2, None, None, None)
    assert rc != 0
    view, stride = C.c_void_p(), C.c_int64()
    rc = L.nw_field_device_view(mesh.h, mesh.field_id("density"), C.byref(view),
                                C.byref(stride))
    assert rc != 0 and not view.value


def test_bad_arguments_are_refused(env):
    P, case, mesh, ls, uvw = env
    L = P.lib()
    assert L.nw_mdot_edge(None, None) == 1          # NW_ERR_ARG
    assert b"nw_mdot_edge" in L.nw_last_error()
    assert L.nw_linsys_zero(None) == 1
    fid = mesh.field_id
    assert L.nw_momentum_diag_post_process(None, 0, 0, 0, 0.5, 1.5, 0.7) == 1
    # three distinct scalar nodal fields, dt > 0
    assert L.nw_momentum_diag_post_process(
        mesh.h, fid("velocity"), fid("density"), fid("dual_nodal_volume"),
        0.5, 1.5, 0.7) == 1
    assert L.nw_momentum_diag_post_process(
        mesh.h, fid("density"), fid("density"), fid("dual_nodal_volume"),
        0.5, 1.5, 0.7) == 1
    assert L.nw_momentum_diag_post_process(
        mesh.h, fid("momentum_diag"), fid("density"), fid("dual_nodal_volume"),
        0.0, 1.5, 0.7) == 1
    assert b"nw_momentum_diag_post_process" in L.nw_last_error()
    assert L.nw_field_fill(mesh.h, 10_000, 0.0) != 0
    with pytest.raises(P.NwError):
        mesh.field_id("no_such_field")
    assert mesh.register("velocity", P.NW_NODE, 3) == mesh.field_id("velocity")
    with pytest.raises(P.NwError):
        mesh.register("velocity", P.NW_NODE, 1)     # same name, another shape
    with pytest.raises(P.NwError):
        mesh.register("too_wide", P.NW_NODE, 1000)
