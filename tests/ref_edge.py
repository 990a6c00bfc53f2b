"""ctypes front end of oracle/_ref/libnalu_ref.so's edge-algorithm entry points
(oracle/ref_edge_driver.cpp): the reference's own MomentumEdgeSolverAlg /
ScalarEdgeSolverAlg / ContinuityEdgeSolverAlg / MdotEdgeAlg, compiled
unmodified, run over a stand-in Realm on arrays handed in from here.  TEST
INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "..", "oracle", "_ref", "libnalu_ref.so")
NODE, EDGE = 0, 1


def available():
    return os.path.exists(SO)


_L = None


def lib():
    global _L
    if _L is None:
        L = C.CDLL(SO)
        vp = C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_world_reset.argtypes = [C.c_int, C.c_long, C.c_long, vp]
        L.ref_world_field.argtypes = [C.c_char_p, C.c_int, C.c_int, vp]
        L.ref_world_option.argtypes = [C.c_char_p, C.c_double]
        L.ref_world_gravity.argtypes = [vp]
        L.ref_world_peclet.argtypes = [C.c_int, C.c_double, C.c_double]
        L.ref_world_flags.argtypes = [C.c_int, C.c_int]
        L.ref_run_momentum.argtypes = [vp, vp]
        L.ref_run_scalar.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, vp, vp]
        L.ref_run_continuity.argtypes = [vp, vp]
        L.ref_run_wall_dist.argtypes = [vp, vp]
        L.ref_run_nodal_grad.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_run_node_kernel.argtypes = [C.c_int, C.c_char_p, vp, vp]
        _L = L
    return _L


class World:
    """one mesh + state + options handed to the stand-in Realm"""

    def __init__(self, ndim, n_nodes, edges):
        self.ndim, self.n_nodes = ndim, n_nodes
        self.edges = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 2)
        self.n_edges = len(self.edges)
        self._keep = {}
        lib().ref_world_reset(ndim, n_nodes, self.n_edges, self.edges.ctypes.data)

    def field(self, name, rank, values, ncomp):
        a = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, ncomp).copy()
        assert len(a) == (self.n_nodes if rank == NODE else self.n_edges), name
        self._keep[(name, rank)] = a
        lib().ref_world_field(name.encode(), rank, ncomp, a.ctypes.data)
        return a

    def options(self, **kv):
        for k, v in kv.items():
            lib().ref_world_option(k.replace("__", ":").encode(), float(v))

    def option(self, key, v):
        lib().ref_world_option(key.encode(), float(v))

    def peclet(self, form, a, b=1.0):
        lib().ref_world_peclet(0 if form == "classic" else 1, float(a), float(b))

    def flags(self, has_vof=False, balanced_buoyancy=False, gravity=None):
        lib().ref_world_flags(int(has_vof), int(balanced_buoyancy))
        if gravity is not None:
            g = np.ascontiguousarray(gravity, dtype=np.float64)
            lib().ref_world_gravity(g.ctypes.data)

    def _run(self, fn, n, *pre):
        lhs = np.zeros((self.n_edges, n, n))
        rhs = np.zeros((self.n_edges, n))
        rc = fn(*pre, lhs.ctypes.data, rhs.ctypes.data)
        if rc:
            raise RuntimeError(lib().ref_last_error().decode())
        return lhs, rhs

    def momentum(self):
        return self._run(lib().ref_run_momentum, 2 * self.ndim)

    def scalar(self, q, dqdx, dflux):
        return self._run(lib().ref_run_scalar, 2, q.encode(), dqdx.encode(),
                         dflux.encode())

    def continuity(self):
        return self._run(lib().ref_run_continuity, 2)

    def node_kernel(self, which, q=""):
        """{Scalar,Momentum,Continuity}MassBDFNodeKernel / WallDistNodeKernel:
        local lhs [nNodes][n][n] and rhs [nNodes][n] of every node"""
        k = {"scalar_mass": 0, "momentum_mass": 1, "continuity_mass": 2,
             "wall_dist": 3}[which]
        n = self.ndim if k == 1 else 1
        lhs = np.zeros((self.n_nodes, n, n))
        rhs = np.zeros((self.n_nodes, n))
        if lib().ref_run_node_kernel(k, q.encode(), lhs.ctypes.data, rhs.ctypes.data):
            raise RuntimeError(lib().ref_last_error().decode())
        return lhs, rhs

    def wall_dist(self):
        return self._run(lib().ref_run_wall_dist, 2)

    def nodal_grad(self, phi, grad):
        """NodalGradEdgeAlg adds into the registered field `grad`"""
        if lib().ref_run_nodal_grad(phi.encode(), grad.encode()):
            raise RuntimeError(lib().ref_last_error().decode())
        return self._keep[(grad, NODE)].copy()

    def peclet_alg(self):
        if lib().ref_run_peclet():
            raise RuntimeError(lib().ref_last_error().decode())
        return (self._keep[("peclet_number", EDGE)].ravel().copy(),
                self._keep[("peclet_factor", EDGE)].ravel().copy())

    def mdot(self):
        if lib().ref_run_mdot():
            raise RuntimeError(lib().ref_last_error().decode())
        return self._keep[("mass_flow_rate", EDGE)].ravel().copy()


# ---------------------------------------------------------------------------
# HypreLinearSystem / HypreUVWLinearSystem (oracle/ref_hypre_driver.cpp)
# ---------------------------------------------------------------------------

class HypreRef:
    """the reference's own HypreLinearSystem over the World set up before it:
    one process plays rank `rank` of `nranks`"""

    def __init__(self, world, node_hid, uvw=False, num_dof=1, rank=0, nranks=1,
                 node_identifier=None, node_owner=None, nalu_id=None,
                 offsets=None, dirichlet_nodes=None):
        L = lib()
        vp = C.c_void_p
        L.ref_hypre_last_error.restype = C.c_char_p
        L.ref_hypre_create.restype = vp
        L.ref_hypre_create.argtypes = [C.c_int, C.c_int]
        L.ref_world_parallel.argtypes = [C.c_int, C.c_int, vp, vp, C.c_long,
                                         C.c_long, C.c_long, vp]
        L.ref_world_int_field.argtypes = [C.c_char_p, C.c_int, C.c_int, vp]
        for f in ("ref_hypre_destroy", "ref_hypre_build_edge_graph_and_finalize",
                  "ref_hypre_load_complete"):
            getattr(L, f).argtypes = [vp]
        L.ref_hypre_dirichlet_nodes.argtypes = [vp, vp, C.c_int]
        L.ref_hypre_sizes.argtypes = [vp, vp]
        L.ref_hypre_graph.argtypes = [vp] * 7
        L.ref_hypre_assemble.argtypes = [vp, vp, vp, C.c_int]
        L.ref_hypre_values.argtypes = [vp, vp, vp]
        L.ref_hypre_reset_rows.argtypes = [vp, vp, C.c_int, C.c_double, C.c_double]
        L.ref_hypre_apply_dirichlet.argtypes = [vp, C.c_char_p, C.c_char_p, vp, C.c_int]
        L.ref_hypre_sweep.argtypes = [vp, C.c_int, C.c_char_p, C.c_char_p,
                                      C.c_char_p, C.c_char_p]
        L.ref_hypre_rhs_shape.argtypes = [vp, vp]
        L.ref_hypre_ij_calls.argtypes = [vp, C.c_int]
        L.ref_hypre_ij_call_sizes.argtypes = [vp, C.c_int, C.c_int, vp]
        L.ref_hypre_ij_call_get.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp]
        self.w = world
        n = world.n_nodes
        self.hid = np.ascontiguousarray(node_hid, dtype=np.int32)
        ident = (np.arange(1, n + 1) if node_identifier is None
                 else np.asarray(node_identifier))
        self.ident = np.ascontiguousarray(ident, dtype=np.int64)
        self.owner = np.ascontiguousarray(
            np.full(n, rank) if node_owner is None else node_owner, dtype=np.int32)
        self.nalu = np.ascontiguousarray(
            self.ident if nalu_id is None else nalu_id, dtype=np.int32)
        owned = self.owner == rank
        own_hid = self.hid[owned & (self.nalu == self.ident)]
        if offsets is None:
            offsets = [0, int(own_hid.max()) + 1]
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        ilo, iup = int(self.offsets[rank]), int(self.offsets[rank + 1])
        L.ref_world_parallel(rank, nranks, self.ident.ctypes.data,
                             self.owner.ctypes.data, ilo, iup,
                             int(self.offsets[-1]), self.offsets.ctypes.data)
        L.ref_world_int_field(b"hypre_global_id", NODE, 1, self.hid.ctypes.data)
        L.ref_world_int_field(b"nalu_global_id", NODE, 1, self.nalu.ctypes.data)
        self.uvw, self.num_dof = uvw, num_dof
        self.h = L.ref_hypre_create(int(uvw), num_dof)
        if not self.h:
            raise RuntimeError(L.ref_hypre_last_error().decode())
        if dirichlet_nodes is not None and len(dirichlet_nodes):
            dn = np.ascontiguousarray(dirichlet_nodes, dtype=np.int32)
            self._chk(L.ref_hypre_dirichlet_nodes(self.h, dn.ctypes.data, len(dn)))
        self._chk(L.ref_hypre_build_edge_graph_and_finalize(self.h))
        sz = (C.c_long * 5)()
        L.ref_hypre_sizes(self.h, sz)
        (self.num_rows_owned, self.nnz_owned, self.num_rows_shared,
         self.nnz_shared, self.num_periodic) = list(sz)
        i32 = lambda k: np.zeros(max(k, 1), dtype=np.int32)  # noqa: E731
        nnz = self.nnz_owned + self.nnz_shared
        rso, rss = i32(self.num_rows_owned + 1), i32(self.num_rows_shared + 1)
        cols, rows = i32(nnz), i32(nnz)
        ris, per = i32(self.num_rows_shared), i32(self.num_periodic)
        L.ref_hypre_graph(self.h, rso.ctypes.data, rss.ctypes.data, cols.ctypes.data,
                          rows.ctypes.data, ris.ctypes.data, per.ctypes.data)
        self.row_start_owned = rso[:self.num_rows_owned + 1]
        self.row_start_shared = rss[:self.num_rows_shared + 1] if self.num_rows_shared else rss[:0]
        self.cols, self.rows = cols[:nnz], rows[:nnz]
        self.row_indices_shared = ris[:self.num_rows_shared]
        self.periodic_rows = per[:self.num_periodic]

    def _chk(self, rc):
        if rc:
            raise RuntimeError(lib().ref_hypre_last_error().decode())

    def assemble(self, lhs, rhs):
        """zeroSystem + get_coeff_applier + operator() per edge; returns the
        applier's values [nnz] and rhs [nrhs][rows]"""
        lhs = np.ascontiguousarray(lhs, dtype=np.float64)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        n = rhs.shape[1]
        self._chk(lib().ref_hypre_assemble(self.h, lhs.ctypes.data, rhs.ctypes.data, n))
        sh = (C.c_long * 3)()
        lib().ref_hypre_rhs_shape(self.h, sh)
        vals = np.zeros(sh[2])
        r = np.zeros((sh[1], sh[0]))
        lib().ref_hypre_values(self.h, vals.ctypes.data, r.ctypes.data)
        return vals, r

    def sweep(self, alg, q="", dqdx="", dflux="", diag_field=""):
        """one assembly as the reference runs it (zeroSystem, the edge
        algorithm's execute() feeding the reference's CoeffApplier,
        loadComplete); alg: 'momentum' | 'continuity' | 'scalar'; diag_field:
        NGPApplyCoeff::extract_diagonal adds lhs(ix, ix) of every node into it"""
        k = {"momentum": 0, "continuity": 1, "scalar": 2}[alg]
        self._chk(lib().ref_hypre_sweep(self.h, k, q.encode(), dqdx.encode(),
                                        dflux.encode(), diag_field.encode()))
        return self.values()

    def reset_rows(self, nodes, diag_value, rhs_residual):
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        self._chk(lib().ref_hypre_reset_rows(self.h, nd.ctypes.data, len(nd),
                                             float(diag_value), float(rhs_residual)))
        return self.values()

    def apply_dirichlet(self, solution, bc_values, nodes):
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        self._chk(lib().ref_hypre_apply_dirichlet(
            self.h, solution.encode(), bc_values.encode(), nd.ctypes.data, len(nd)))
        return self.values()

    def values(self):
        sh = (C.c_long * 3)()
        lib().ref_hypre_rhs_shape(self.h, sh)
        vals = np.zeros(sh[2])
        r = np.zeros((sh[1], sh[0]))
        lib().ref_hypre_values(self.h, vals.ctypes.data, r.ctypes.data)
        return vals, r

    def load_complete(self):
        """loadComplete; returns {which: [(is_add, ncols, rows, cols, values)]}
        for the matrix (0) and each right-hand side (1..)"""
        self._chk(lib().ref_hypre_load_complete(self.h))
        out = {}
        nrhs = self.num_dof if self.uvw else 1
        for which in range(0, 1 + nrhs):
            calls = []
            for k in range(lib().ref_hypre_ij_calls(self.h, which)):
                sz = (C.c_long * 3)()
                lib().ref_hypre_ij_call_sizes(self.h, which, k, sz)
                nr, nv, add = list(sz)
                ncols, rows = np.zeros(max(nr, 1), np.int32), np.zeros(max(nr, 1), np.int32)
                cols, vals = np.zeros(max(nv, 1), np.int32), np.zeros(max(nv, 1))
                lib().ref_hypre_ij_call_get(self.h, which, k, ncols.ctypes.data,
                                            rows.ctypes.data, cols.ctypes.data,
                                            vals.ctypes.data)
                calls.append((bool(add), ncols[:nr], rows[:nr], cols[:nv], vals[:nv]))
            out[which] = calls
        return out

    def close(self):
        if self.h:
            lib().ref_hypre_destroy(self.h)
            self.h = None
