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
