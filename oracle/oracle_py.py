"""ctypes binding of the CPU oracle (oracle/libedge_oracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product
(nalu-wind_b200/, include/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)


class Peclet(C.Structure):
    _fields_ = [("form", C.c_int), ("a", C.c_double), ("b", C.c_double)]


class ContinuityOpts(C.Structure):
    _fields_ = [("dt", C.c_double), ("gamma1", C.c_double),
                ("noc_fac", C.c_double), ("interp_together", C.c_double),
                ("solve_incompressible", C.c_double)]


class ScalarOpts(C.Structure):
    _fields_ = [("alpha", C.c_double), ("alpha_upw", C.c_double),
                ("ho_upwind", C.c_double), ("relax_fac", C.c_double),
                ("use_limiter", C.c_int), ("eps", C.c_double), ("pf", Peclet)]


class MomentumOpts(C.Structure):
    _fields_ = [("include_divu", C.c_double), ("alpha", C.c_double),
                ("alpha_upw", C.c_double), ("ho_upwind", C.c_double),
                ("relax_fac", C.c_double), ("use_limiter", C.c_int),
                ("eps", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "libedge_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("edge_oracle.cpp", "edge_oracle.h")]
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp = C.c_void_p
    L.orc_set_num_threads.argtypes = [C.c_int]
    L.orc_get_num_threads.restype = C.c_int
    L.orc_peclet_eval.restype = C.c_double
    L.orc_peclet_eval.argtypes = [C.POINTER(Peclet), C.c_double]
    L.orc_applier_dense_create.restype = vp
    L.orc_applier_dense_create.argtypes = [C.c_int64, C.c_int]
    L.orc_applier_dense_get.argtypes = [vp, c_f64p, c_f64p]
    L.orc_graph_create.restype = vp
    L.orc_graph_create.argtypes = [C.c_int, C.c_int64, C.c_int64]
    L.orc_graph_destroy.argtypes = [vp]
    L.orc_graph_set_skipped.argtypes = [vp, c_i64p, C.c_int64]
    L.orc_graph_add_edges.argtypes = [vp, C.c_int64, c_i32p, c_i64p]
    L.orc_graph_add_nodes.argtypes = [vp, C.c_int64, c_i32p, c_i64p]
    L.orc_graph_finalize.argtypes = [vp]
    L.orc_graph_size.restype = C.c_int64
    L.orc_graph_size.argtypes = [vp, C.c_int]
    L.orc_graph_copy.argtypes = [vp, C.c_int, c_i64p]
    L.orc_applier_hypre_create.restype = vp
    L.orc_applier_hypre_create.argtypes = [vp, c_i64p, C.c_int64, C.c_int]
    L.orc_applier_hypre_reset.argtypes = [vp]
    L.orc_applier_hypre_track_abs.argtypes = [vp, C.c_int]
    L.orc_applier_hypre_get.argtypes = [vp, c_f64p, c_f64p]
    L.orc_applier_hypre_get_abs.argtypes = [vp, c_f64p, c_f64p]
    L.orc_applier_hypre_enable_log.argtypes = [vp, C.c_int64]
    L.orc_applier_hypre_get_log.argtypes = [vp, c_i64p, c_i64p]
    L.orc_applier_destroy.argtypes = [vp]
    L.orc_mdot_edge.argtypes = [
        C.c_int, C.c_int64, c_i32p] + [c_f64p] * 7 + [
        C.c_double, C.c_double, c_f64p]
    L.orc_peclet_edge.argtypes = [
        C.c_int, C.c_int64, c_i32p] + [c_f64p] * 4 + [
        C.POINTER(Peclet), C.c_double, c_f64p, c_f64p]
    L.orc_nodal_grad_edge.argtypes = [
        C.c_int, C.c_int, C.c_int64, c_i32p] + [c_f64p] * 4
    L.orc_continuity_edge.argtypes = [
        C.c_int, C.c_int64, c_i32p] + [c_f64p] * 7 + [
        C.POINTER(ContinuityOpts), vp]
    L.orc_scalar_edge.argtypes = [
        C.c_int, C.c_int64, c_i32p] + [c_f64p] * 8 + [
        C.POINTER(ScalarOpts), vp]
    L.orc_momentum_edge.argtypes = [
        C.c_int, C.c_int64, c_i32p] + [c_f64p] * 9 + [
        C.POINTER(MomentumOpts), vp, c_f64p]
    _LIB = L
    return L


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_f64p)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_i32p)


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(c_i64p)


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def peclet(form="classic", a=0.0, b=1.0):
    return Peclet(0 if form == "classic" else 1, float(a), float(b))


def peclet_eval(pf, pecnum):
    """PecletFunction::execute (src/PecletFunction.C:41-45, 68-71)"""
    return float(lib().orc_peclet_eval(C.byref(pf), float(pecnum)))


class DenseSink:
    """TestLinearSystem (unit_tests/UnitTestLinearSystem.h)."""

    def __init__(self, n_nodes, num_dof):
        self.n = n_nodes * num_dof
        self.h = lib().orc_applier_dense_create(n_nodes, num_dof)

    def get(self):
        lhs = np.zeros((self.n, self.n))
        rhs = np.zeros(self.n)
        lib().orc_applier_dense_get(self.h, lhs.ctypes.data_as(c_f64p),
                                    rhs.ctypes.data_as(c_f64p))
        return lhs, rhs

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_applier_destroy(self.h)
            self.h = None


class RecordSink:
    """keeps every local block (lhs n x n, rhs n) in call order; run the oracle
    on one thread so that call order == edge order"""

    def __init__(self):
        L = lib()
        L.orc_applier_record_create.restype = C.c_void_p
        L.orc_applier_record_count.restype = C.c_int64
        L.orc_applier_record_count.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.orc_applier_record_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.h = L.orc_applier_record_create()

    def get(self):
        n = C.c_int(0)
        cnt = int(lib().orc_applier_record_count(self.h, C.byref(n)))
        n = n.value
        lhs = np.zeros((cnt, n, n))
        rhs = np.zeros((cnt, n))
        if cnt:
            lib().orc_applier_record_get(self.h, lhs.ctypes.data, rhs.ctypes.data)
        return lhs, rhs

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_applier_destroy(C.c_void_p(self.h))
            self.h = None


class Graph:
    """HypreLinearSystem graph (oracle restatement)."""

    def __init__(self, num_dof, i_lower, i_upper):
        self.num_dof = num_dof
        self.i_lower, self.i_upper = int(i_lower), int(i_upper)
        self.h = lib().orc_graph_create(num_dof, int(i_lower), int(i_upper))

    def set_skipped(self, rows):
        r, p = _i64(rows)
        lib().orc_graph_set_skipped(self.h, p, len(r))

    def add_edges(self, edge_nodes, node_hid):
        en, pe = _i32(edge_nodes)
        nh, ph = _i64(node_hid)
        lib().orc_graph_add_edges(self.h, en.size // 2, pe, ph)

    def add_nodes(self, nodes, node_hid):
        nn, pn = _i32(nodes)
        nh, ph = _i64(node_hid)
        lib().orc_graph_add_nodes(self.h, nn.size, pn, ph)

    def finalize(self):
        L = lib()
        L.orc_graph_finalize(self.h)
        sz = lambda w: int(L.orc_graph_size(self.h, w))
        self.num_rows_owned = sz(0)
        self.nnz_owned = sz(1)
        self.num_rows_shared = sz(2)
        self.nnz_shared = sz(3)
        self.num_periodic = sz(4)

        def cp(what, n):
            out = np.zeros(n, dtype=np.int64)
            if n:
                L.orc_graph_copy(self.h, what, out.ctypes.data_as(c_i64p))
            return out
        self.row_start_owned = cp(0, self.num_rows_owned + 1)
        self.row_start_shared = cp(1, self.num_rows_shared + 1)
        self.cols = cp(2, self.nnz_owned + self.nnz_shared)
        self.rows = cp(3, self.nnz_owned + self.nnz_shared)
        self.row_indices_shared = cp(4, self.num_rows_shared)
        self.periodic_rows = cp(5, self.num_periodic)
        return self

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_graph_destroy(self.h)
            self.h = None


class HypreSink:
    """HypreLinSysCoeffApplier / HypreUVWLinSysCoeffApplier (oracle)."""

    def __init__(self, graph, node_hid, uvw_ndim=0):
        self.graph = graph
        self.uvw_ndim = uvw_ndim
        nh, ph = _i64(node_hid)
        self.h = lib().orc_applier_hypre_create(graph.h, ph, len(nh), uvw_ndim)
        self.nrhs = uvw_ndim if uvw_ndim > 0 else 1
        self.total_rows = graph.num_rows_owned + graph.num_rows_shared
        self.nnz = graph.nnz_owned + graph.nnz_shared
        self._log_calls = 0

    def reset(self):
        lib().orc_applier_hypre_reset(self.h)

    def track_abs(self, on):
        """|contribution| sums (tolerance scale of the tests); off for timing"""
        lib().orc_applier_hypre_track_abs(self.h, int(bool(on)))

    def enable_log(self, n_calls):
        self._log_calls = n_calls
        lib().orc_applier_hypre_enable_log(self.h, n_calls)

    def get_log(self):
        n = 2 if self.uvw_ndim > 0 else 2 * self.graph.num_dof
        slots = np.zeros((self._log_calls, n, n), dtype=np.int64)
        ridx = np.zeros((self._log_calls, n), dtype=np.int64)
        lib().orc_applier_hypre_get_log(
            self.h, slots.ctypes.data_as(c_i64p), ridx.ctypes.data_as(c_i64p))
        return slots, ridx

    def _get(self, fn):
        vals = np.zeros(self.nnz)
        rhs = np.zeros((self.nrhs, self.total_rows))  # column-major (row, d)
        fn(self.h, vals.ctypes.data_as(c_f64p), rhs.ctypes.data_as(c_f64p))
        return vals, rhs

    def get(self):
        return self._get(lib().orc_applier_hypre_get)

    def apply(self, entity_nodes, lhs, rhs):
        """CoeffApplier::operator() over a batch: entity_nodes [nEnt][npe],
        lhs [nEnt][n][n], rhs [nEnt][n]"""
        en = np.ascontiguousarray(entity_nodes, dtype=np.int32)
        L = np.ascontiguousarray(lhs, dtype=np.float64)
        R = np.ascontiguousarray(rhs, dtype=np.float64)
        n_ent, npe = en.shape
        n = R.shape[1]
        f = lib().orc_applier_apply
        f.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                      C.c_void_p, C.c_int]
        f(self.h, n_ent, npe, en.ctypes.data, L.ctypes.data, R.ctypes.data, n)

    def reset_rows(self, nodes, diag_value=0.0, rhs_residual=0.0):
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        f = lib().orc_applier_hypre_reset_rows
        f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_double]
        f(self.h, nd.size, nd.ctypes.data, diag_value, rhs_residual)

    def apply_dirichlet(self, nodes, solution, bc_values):
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        sol = np.ascontiguousarray(solution, dtype=np.float64)
        bc = np.ascontiguousarray(bc_values, dtype=np.float64)
        ncomp = 1 if sol.ndim == 1 else sol.shape[1]
        f = lib().orc_applier_hypre_dirichlet
        f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                      C.c_int]
        f(self.h, nd.size, nd.ctypes.data, sol.ctypes.data, bc.ctypes.data, ncomp)

    def get_abs(self):
        return self._get(lib().orc_applier_hypre_get_abs)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_applier_destroy(self.h)
            self.h = None


def mdot_edge(ndim, edge_nodes, coords, vel, gpdx, rho, p, udiag, area,
              noc_fac=1.0, interp_together=1.0):
    en, pe = _i32(edge_nodes)
    ne = en.size // 2
    arrs = [_f(x) for x in (coords, vel, gpdx, rho, p, udiag, area)]
    out = np.zeros(ne)
    lib().orc_mdot_edge(ndim, ne, pe, *[a[1] for a in arrs],
                        float(noc_fac), float(interp_together),
                        out.ctypes.data_as(c_f64p))
    return out


def peclet_edge(ndim, edge_nodes, coords, vrtm, rho, visc, pf, eps=1e-16):
    en, pe = _i32(edge_nodes)
    ne = en.size // 2
    arrs = [_f(x) for x in (coords, vrtm, rho, visc)]
    pn = np.zeros(ne)
    pfac = np.zeros(ne)
    lib().orc_peclet_edge(ndim, ne, pe, *[a[1] for a in arrs], C.byref(pf),
                          float(eps), pn.ctypes.data_as(c_f64p),
                          pfac.ctypes.data_as(c_f64p))
    return pn, pfac


def nodal_grad_edge(dim1, dim2, edge_nodes, phi, area, dual_vol, n_nodes):
    en, pe = _i32(edge_nodes)
    ne = en.size // 2
    arrs = [_f(x) for x in (phi, area, dual_vol)]
    grad = np.zeros((n_nodes, dim1 * dim2))
    lib().orc_nodal_grad_edge(dim1, dim2, ne, pe, *[a[1] for a in arrs],
                              grad.ctypes.data_as(c_f64p))
    return grad


def continuity_edge(ndim, edge_nodes, coords, vel, gpdx, rho, p, udiag, area,
                    sink, dt=1.0, gamma1=1.0, noc_fac=1.0,
                    interp_together=1.0, solve_incompressible=0.0):
    en, pe = _i32(edge_nodes)
    arrs = [_f(x) for x in (coords, vel, gpdx, rho, p, udiag, area)]
    o = ContinuityOpts(dt, gamma1, noc_fac, interp_together,
                       solve_incompressible)
    lib().orc_continuity_edge(ndim, en.size // 2, pe, *[a[1] for a in arrs],
                              C.byref(o), sink.h)


def scalar_edge(ndim, edge_nodes, coords, vrtm, q, dqdx, rho, dflux, area,
                mdot, sink, alpha=0.0, alpha_upw=1.0, ho_upwind=1.0,
                relax_fac=1.0, use_limiter=False, eps=1e-16, pf=None):
    en, pe = _i32(edge_nodes)
    arrs = [_f(x) for x in (coords, vrtm, q, dqdx, rho, dflux, area, mdot)]
    o = ScalarOpts(alpha, alpha_upw, ho_upwind, relax_fac,
                   1 if use_limiter else 0, eps, pf or peclet())
    lib().orc_scalar_edge(ndim, en.size // 2, pe, *[a[1] for a in arrs],
                          C.byref(o), sink.h)


def momentum_edge(ndim, edge_nodes, coords, vel, dudx, visc, rho, mask, area,
                  mdot, pecfac, sink, include_divu=0.0, alpha=0.0,
                  alpha_upw=1.0, ho_upwind=1.0, relax_fac=1.0,
                  use_limiter=False, eps=1e-16, udiag_accum=None,
                  mass_vof=None):
    """mass_vof: the mass_vof_balanced_flow_rate edge field; given = the
    realm_has_vof_ branch (MomentumEdgeSolverAlg.C:88, 124-125, 174-192)"""
    en, pe = _i32(edge_nodes)
    arrs = [_f(x) for x in (coords, vel, dudx, visc, rho, mask, area, mdot,
                            pecfac)]
    o = MomentumOpts(include_divu, alpha, alpha_upw, ho_upwind, relax_fac,
                     1 if use_limiter else 0, eps)
    ud = None
    if udiag_accum is not None:
        assert udiag_accum.dtype == np.float64 and udiag_accum.flags.c_contiguous
        ud = udiag_accum.ctypes.data_as(c_f64p)
    if mass_vof is not None:
        mv = _f(mass_vof)
        f = lib().orc_momentum_edge_vof
        f.argtypes = [C.c_int, C.c_int64, c_i32p] + [c_f64p] * 10 + [
            C.POINTER(MomentumOpts), C.c_void_p, c_f64p]
        f.restype = None
        ptrs = [a[1] for a in arrs]
        f(ndim, en.size // 2, pe, *ptrs[:8], mv[1], ptrs[8], C.byref(o), sink.h, ud)
        return
    lib().orc_momentum_edge(ndim, en.size // 2, pe, *[a[1] for a in arrs],
                            C.byref(o), sink.h, ud)


def _f64s(*arrs):
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
    return keep, [C.c_void_p(a.ctypes.data) for a in keep]


def scalar_mass_bdf_node(nodes, q3, rho3, dnv3, dt, g1, g2, g3, sink):
    """ScalarMassBDFNodeKernel over `nodes`; q3/rho3/dnv3 = (Nm1, N, Np1)"""
    nd = np.ascontiguousarray(nodes, dtype=np.int32)
    keep, ptrs = _f64s(*q3, *rho3, *dnv3)
    f = lib().orc_scalar_mass_bdf_node
    f.argtypes = [C.c_int64, C.c_void_p] + [C.c_void_p] * 9 + [C.c_double] * 4 + [C.c_void_p]
    f(nd.size, nd.ctypes.data, *ptrs, dt, g1, g2, g3, sink.h)


def momentum_mass_bdf_node(ndim, nodes, u3, rho3, dnv3, dpdx, dt, g1, g2, g3, sink):
    nd = np.ascontiguousarray(nodes, dtype=np.int32)
    keep, ptrs = _f64s(*u3, *rho3, *dnv3, dpdx)
    f = lib().orc_momentum_mass_bdf_node
    f.argtypes = ([C.c_int, C.c_int64, C.c_void_p] + [C.c_void_p] * 10 +
                  [C.c_double] * 4 + [C.c_void_p])
    f(ndim, nd.size, nd.ctypes.data, *ptrs, dt, g1, g2, g3, sink.h)


def continuity_mass_bdf_node(nodes, rho3, dnv3, dt, g1, g2, g3, sink):
    nd = np.ascontiguousarray(nodes, dtype=np.int32)
    keep, ptrs = _f64s(*rho3, *dnv3)
    f = lib().orc_continuity_mass_bdf_node
    f.argtypes = [C.c_int64, C.c_void_p] + [C.c_void_p] * 6 + [C.c_double] * 4 + [C.c_void_p]
    f(nd.size, nd.ctypes.data, *ptrs, dt, g1, g2, g3, sink.h)


def wall_dist_edge(ndim, edge_nodes, coords, area, sink):
    en = np.ascontiguousarray(edge_nodes, dtype=np.int32)
    keep, ptrs = _f64s(coords, area)
    f = lib().orc_wall_dist_edge
    f.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(ndim, en.size // 2, en.ctypes.data, *ptrs, sink.h)


def wall_dist_node(nodes, dnv, sink):
    nd = np.ascontiguousarray(nodes, dtype=np.int32)
    keep, ptrs = _f64s(dnv)
    f = lib().orc_wall_dist_node
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    f(nd.size, nd.ctypes.data, *ptrs, sink.h)


def geometry_interior_hex8(elem_nodes, coords, edge_nodes, n_nodes, elem_owned=None):
    """GeometryInteriorAlg<Hex8>: returns (dual_nodal_volume, elem_volume, edge_area)"""
    el = np.ascontiguousarray(elem_nodes, dtype=np.int32)
    en = np.ascontiguousarray(edge_nodes, dtype=np.int32)
    xyz = np.ascontiguousarray(coords, dtype=np.float64)
    own = None if elem_owned is None else np.ascontiguousarray(elem_owned, dtype=np.uint8)
    dnv = np.zeros(n_nodes)
    ev = np.zeros(len(el))
    area = np.zeros((en.size // 2, 3))
    f = lib().orc_geometry_interior_hex8
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(len(el), el.ctypes.data, None if own is None else own.ctypes.data,
      xyz.ctypes.data, en.size // 2, en.ctypes.data, dnv.ctypes.data,
      ev.ctypes.data, area.ctypes.data)
    return dnv, ev, area


def geometry_interior_quad4(elem_nodes, coords, edge_nodes, n_nodes, elem_owned=None):
    """GeometryInteriorAlg<Quad4_2D>: (dual_nodal_volume, elem_volume, edge_area[n][2])"""
    el = np.ascontiguousarray(elem_nodes, dtype=np.int32)
    en = np.ascontiguousarray(edge_nodes, dtype=np.int32)
    xy = np.ascontiguousarray(coords, dtype=np.float64)
    own = None if elem_owned is None else np.ascontiguousarray(elem_owned, dtype=np.uint8)
    dnv = np.zeros(n_nodes)
    ev = np.zeros(len(el))
    area = np.zeros((en.size // 2, 2))
    f = lib().orc_geometry_interior_quad4
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(len(el), el.ctypes.data, None if own is None else own.ctypes.data,
      xy.ctypes.data, en.size // 2, en.ctypes.data, dnv.ctypes.data,
      ev.ctypes.data, area.ctypes.data)
    return dnv, ev, area


def geometry_interior_3d(topo, elem_nodes, coords, edge_nodes, n_nodes,
                         elem_owned=None, accumulate=None):
    """GeometryInteriorAlg for one 3-D element block; topo in 'hex' 'tet' 'wed'
    'pyr'.  Returns (dual_nodal_volume, elem_volume, edge_area); `accumulate` =
    (dual_nodal_volume, edge_area) of earlier blocks of the same mesh to add to."""
    name = {"hex": "hex8", "tet": "tet4", "wed": "wed6", "pyr": "pyr5"}[topo]
    el = np.ascontiguousarray(elem_nodes, dtype=np.int32)
    en = np.ascontiguousarray(edge_nodes, dtype=np.int32)
    xyz = np.ascontiguousarray(coords, dtype=np.float64)
    own = None if elem_owned is None else np.ascontiguousarray(elem_owned, dtype=np.uint8)
    if accumulate is None:
        dnv, area = np.zeros(n_nodes), np.zeros((en.size // 2, 3))
    else:
        dnv, area = accumulate
    ev = np.zeros(len(el))
    f = getattr(lib(), "orc_geometry_interior_" + name)
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(len(el), el.ctypes.data, None if own is None else own.ctypes.data,
      xyz.ctypes.data, en.size // 2, en.ctypes.data, dnv.ctypes.data,
      ev.ctypes.data, area.ctypes.data)
    return dnv, ev, area


def mdot_continuity_edge_ext(ndim, edge_nodes, coords, vel, gpdx, rho, p, udiag,
                             area, noc_fac=1.0, interp_together=1.0,
                             gravity=None, source=None, source_mask=None,
                             edge_face_vel_mag=None, dt=1.0, gamma1=1.0,
                             solve_incompressible=0.0, sink=None):
    """MdotEdgeAlg (sink None: returns mdot) / ContinuityEdgeSolverAlg with the
    optional balanced-forcing (gravity + source + source_mask) and GCL
    (edge_face_vel_mag) terms"""
    en = np.ascontiguousarray(edge_nodes, dtype=np.int32)
    ne = en.size // 2
    bal = gravity is not None
    gcl = edge_face_vel_mag is not None
    z1 = np.zeros(1)
    keep, ptrs = _f64s(coords, vel, gpdx, rho, p, udiag, area)
    k2, p2 = _f64s(gravity if bal else np.zeros(3), source if bal else z1,
                   source_mask if bal else z1, edge_face_vel_mag if gcl else z1)
    out = np.zeros(ne)
    f = lib().orc_mdot_continuity_edge_ext
    f.argtypes = ([C.c_int, C.c_int64, C.c_void_p] + [C.c_void_p] * 7 +
                  [C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 3 +
                  [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double,
                   C.c_void_p, C.c_void_p])
    f(ndim, ne, en.ctypes.data, *ptrs, noc_fac, interp_together, int(bal),
      p2[0], p2[1], p2[2], int(gcl), p2[3], dt, gamma1, solve_incompressible,
      out.ctypes.data, None if sink is None else sink.h)
    return out
