"""Python host binding of the B200-native nalu-wind edge-assembly path.

Thin ctypes layer over the C ABI in include/nalu_edge_b200.h (the product is the
C ABI + CUDA kernels; this module exists so that the tests and bench.py read
like the reference's own unit tests: create a mesh, register fields by the
reference's field names, create a LinearSystem, buildEdgeToNodeGraph,
finalizeLinearSystem, zeroSystem, run an edge algorithm, loadComplete).

Load with `__graft_entry__.load_package()` (the directory name is not a valid
Python identifier).  There is no CPU fallback: every compute call raises
NwError when the CUDA library or a device is missing.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NW_LIB_PATH: load another build of the same ABI (the phase-timing build)
LIB_PATH = os.environ.get("NW_LIB_PATH") or os.path.join(
    _HERE, "libnalu_edge_b200.so")
MESHGEN_PATH = os.path.join(_HERE, "libnw_meshgen.so")

NW_NODE, NW_EDGE = 0, 1
NW_LINSYS_HYPRE, NW_LINSYS_HYPRE_UVW = 0, 1
NW_SCATTER_SEGMENTED, NW_SCATTER_ATOMIC = 0, 1
NW_PECLET_CLASSIC, NW_PECLET_TANH = 0, 1

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)


class NwError(RuntimeError):
    pass


class MeshDesc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("n_nodes", C.c_int64), ("n_edges", C.c_int64),
                ("edge_nodes", c_i32p), ("node_hypre_id", c_i64p),
                ("node_own_hypre_id", c_i64p), ("hypre_offsets", c_i64p),
                ("coords", c_f64p), ("tile_nodes", C.c_int32)]


class MeshStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "n_nodes", "n_edges", "n_tiles", "n_tile_edges", "n_halo_nodes",
        "max_tile_nodes", "max_tile_staged", "max_tile_edges",
        "max_tile_halfedges", "plan_bytes_device")]


class PecletFn(C.Structure):
    _fields_ = [("form", C.c_int32), ("a", C.c_double), ("b", C.c_double)]


class MdotOpts(C.Structure):
    _fields_ = [("noc_fac", C.c_double), ("interp_together", C.c_double)]


class PecletOpts(C.Structure):
    _fields_ = [("pf", PecletFn), ("eps", C.c_double)]


class ContinuityOpts(C.Structure):
    _fields_ = [("dt", C.c_double), ("gamma1", C.c_double),
                ("noc_fac", C.c_double), ("interp_together", C.c_double),
                ("solve_incompressible", C.c_double)]


class ScalarOpts(C.Structure):
    _fields_ = [("alpha", C.c_double), ("alpha_upw", C.c_double),
                ("ho_upwind", C.c_double), ("relax_fac", C.c_double),
                ("use_limiter", C.c_int32), ("eps", C.c_double),
                ("pf", PecletFn)]


class MomentumOpts(C.Structure):
    _fields_ = [("include_divu", C.c_double), ("alpha", C.c_double),
                ("alpha_upw", C.c_double), ("ho_upwind", C.c_double),
                ("relax_fac", C.c_double), ("use_limiter", C.c_int32),
                ("eps", C.c_double), ("fuse_peclet", C.c_int32),
                ("pf", PecletFn), ("pec_eps", C.c_double),
                ("diag_field", C.c_int32), ("has_vof", C.c_int32)]


class MdotExtraOpts(C.Structure):
    _fields_ = [("add_balanced_forcing", C.c_int32), ("gravity", C.c_double * 3),
                ("source_field", C.c_int32), ("source_mask_field", C.c_int32),
                ("needs_gcl", C.c_int32), ("edge_face_vel_mag_field", C.c_int32)]


class MassBdfOpts(C.Structure):
    _fields_ = [("dt", C.c_double), ("gamma1", C.c_double),
                ("gamma2", C.c_double), ("gamma3", C.c_double)] + [
        (n, C.c_int32) for n in (
            "q_nm1", "q_n", "q_np1", "rho_nm1", "rho_n", "rho_np1",
            "dnv_nm1", "dnv_n", "dnv_np1", "dpdx")]


NW_MASS_SCALAR, NW_MASS_MOMENTUM, NW_MASS_CONTINUITY = 0, 1, 2


class LinsysSizes(C.Structure):
    _fields_ = [("i_lower", C.c_int64), ("i_upper", C.c_int64),
                ("num_rows_owned", C.c_int64), ("num_nonzeros_owned", C.c_int64),
                ("num_rows_shared", C.c_int64),
                ("num_nonzeros_shared", C.c_int64),
                ("num_periodic_rows", C.c_int64), ("num_rhs", C.c_int32),
                ("block", C.c_int32)]


# every symbol include/nalu_edge_b200.h declares (checked by the CPU tests)
ABI_SYMBOLS = [
    "nw_last_error", "nw_version", "nw_debug_phase_times", "nw_debug_skip_exchange", "nw_ctx_create", "nw_ctx_destroy",
    "nw_ctx_sync", "nw_ctx_stream", "nw_comm_unique_id", "nw_ctx_comm_init",
    "nw_ctx_peer_memory", "nw_ctx_join_comm", "nw_mesh_halo_transport",
    "nw_linsys_halo_transport",
    "nw_mesh_create", "nw_mesh_destroy", "nw_mesh_get_stats",
    "nw_field_register", "nw_field_find", "nw_field_upload",
    "nw_field_stage", "nw_field_commit", "nw_field_download", "nw_field_fill", "nw_field_device_view",
    "nw_mesh_get_node_permutation", "nw_geometry_interior_hex8", "nw_geometry_interior_quad4",
    "nw_geometry_interior_tet4", "nw_geometry_interior_wed6", "nw_geometry_interior_pyr5", "nw_mdot_edge",
    "nw_mdot_edge_ext", "nw_assemble_continuity_edge_ext", "nw_peclet_edge",
    "nw_nodal_grad_edge", "nw_nodal_grad_edge_pair", "nw_linsys_create", "nw_linsys_destroy",
    "nw_linsys_set_skipped_rows", "nw_linsys_build_edge_to_node_graph",
    "nw_linsys_finalize", "nw_linsys_get_sizes", "nw_linsys_get_graph",
    "nw_linsys_get_edge_slots", "nw_linsys_zero",
    "nw_linsys_set_scatter_mode", "nw_linsys_set_eager_exchange",
    "nw_linsys_uses_tile_path",
    "nw_assemble_continuity_edge",
    "nw_assemble_scalar_edge", "nw_assemble_scalar_edge_pair", "nw_assemble_momentum_edge",
    "nw_assemble_mass_bdf_node", "nw_assemble_wall_dist_edge",
    "nw_assemble_wall_dist_node", "nw_linsys_write_preassembly_files", "nw_linsys_sum_into", "nw_linsys_reset_rows",
    "nw_linsys_apply_dirichlet_bcs", "nw_linsys_load_complete",
    "nw_linsys_device_arrays", "nw_linsys_get_values", "nw_linsys_rhs_norm2", "nw_linsys_rhs_norm2_global",
    "nw_mesh_halo_send_count", "nw_mesh_halo_get_send", "nw_mesh_halo_set_recv",
    "nw_mesh_halo_commit", "nw_field_parallel_sum", "nw_field_periodic_update", "nw_momentum_diag_post_process", "nw_field_copy_owned_to_shared", "nw_linsys_halo_send_info",
    "nw_linsys_halo_get_send", "nw_linsys_halo_set_recv",
    "nw_linsys_halo_commit", "nw_linsys_halo_get_recv_slots",
    "nw_linsys_get_extra",
]

_lib = None


def build(force=False):
    """compile the CUDA library and the mesh generator in-tree (make)"""
    args = ["make", "-C", _HERE, "-s", "-j4"]
    if force:
        args.append("-B")
    subprocess.check_call(args)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NwError("%s is missing: run __graft_entry__.build() "
                      "(no CPU fallback exists)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.nw_last_error.restype = C.c_char_p
    L.nw_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.nw_ctx_destroy.argtypes = [vp]
    L.nw_ctx_sync.argtypes = [vp]
    L.nw_debug_skip_exchange.argtypes = [vp, C.c_int]
    L.nw_ctx_stream.restype = vp
    L.nw_ctx_stream.argtypes = [vp]
    L.nw_comm_unique_id.argtypes = [vp]
    L.nw_ctx_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.nw_ctx_peer_memory.argtypes = [vp]
    L.nw_ctx_join_comm.argtypes = [vp]
    L.nw_mesh_halo_transport.argtypes = [vp]
    L.nw_linsys_halo_transport.argtypes = [vp]
    L.nw_mesh_create.argtypes = [vp, C.POINTER(MeshDesc), C.POINTER(vp)]
    L.nw_mesh_destroy.argtypes = [vp]
    L.nw_mesh_get_stats.argtypes = [vp, C.POINTER(MeshStats)]
    L.nw_field_register.argtypes = [vp, C.c_char_p, C.c_int, C.c_int,
                                    C.POINTER(C.c_int)]
    L.nw_field_find.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int)]
    L.nw_field_upload.argtypes = [vp, C.c_int, vp]
    L.nw_field_download.argtypes = [vp, C.c_int, vp]
    L.nw_field_stage.argtypes = [vp, C.c_int, vp]
    L.nw_field_commit.argtypes = [vp, C.c_int]
    L.nw_field_fill.argtypes = [vp, C.c_int, C.c_double]
    L.nw_field_device_view.argtypes = [vp, C.c_int, C.POINTER(vp),
                                       C.POINTER(C.c_int64)]
    L.nw_mesh_get_node_permutation.argtypes = [vp, C.POINTER(C.c_int64), c_i32p]
    L.nw_mdot_edge.argtypes = [vp, C.POINTER(MdotOpts)]
    L.nw_geometry_interior_hex8.argtypes = [vp, C.c_int64, vp, vp, C.c_int,
                                            C.c_int, C.c_int]
    L.nw_geometry_interior_quad4.argtypes = L.nw_geometry_interior_hex8.argtypes
    for _t in ("tet4", "wed6", "pyr5"):
        getattr(L, "nw_geometry_interior_" + _t).argtypes = \
            L.nw_geometry_interior_hex8.argtypes
    L.nw_mdot_edge_ext.argtypes = [vp, C.POINTER(MdotOpts), C.POINTER(MdotExtraOpts)]
    L.nw_assemble_continuity_edge_ext.argtypes = [
        vp, C.POINTER(ContinuityOpts), C.POINTER(MdotExtraOpts)]
    L.nw_peclet_edge.argtypes = [vp, C.c_int, C.POINTER(PecletOpts)]
    L.nw_nodal_grad_edge.argtypes = [vp, C.c_int, C.c_int]
    L.nw_nodal_grad_edge_pair.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.nw_linsys_create.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.nw_linsys_destroy.argtypes = [vp]
    L.nw_linsys_set_skipped_rows.argtypes = [vp, c_i64p, C.c_int64]
    L.nw_linsys_build_edge_to_node_graph.argtypes = [vp]
    L.nw_linsys_finalize.argtypes = [vp]
    L.nw_linsys_get_sizes.argtypes = [vp, C.POINTER(LinsysSizes)]
    L.nw_linsys_get_graph.argtypes = [vp] + [c_i64p] * 6
    L.nw_linsys_get_edge_slots.argtypes = [vp, c_i64p, c_i64p]
    L.nw_linsys_zero.argtypes = [vp]
    L.nw_linsys_set_scatter_mode.argtypes = [vp, C.c_int]
    L.nw_linsys_set_eager_exchange.argtypes = [vp, C.c_int]
    L.nw_linsys_uses_tile_path.argtypes = [vp]
    L.nw_assemble_continuity_edge.argtypes = [vp, C.POINTER(ContinuityOpts)]
    L.nw_assemble_scalar_edge.argtypes = [vp, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(ScalarOpts)]
    L.nw_assemble_momentum_edge.argtypes = [vp, C.c_int,
                                            C.POINTER(MomentumOpts)]
    L.nw_assemble_scalar_edge_pair.argtypes = [
        vp, C.c_int, C.c_int, C.c_int, C.POINTER(ScalarOpts),
        vp, C.c_int, C.c_int, C.c_int, C.POINTER(ScalarOpts)]
    L.nw_linsys_sum_into.argtypes = [vp, C.c_int64, C.c_int, vp, vp, vp]
    L.nw_linsys_reset_rows.argtypes = [vp, C.c_int64, vp, C.c_double, C.c_double]
    L.nw_assemble_mass_bdf_node.argtypes = [vp, C.c_int, C.POINTER(MassBdfOpts)]
    L.nw_assemble_wall_dist_edge.argtypes = [vp]
    L.nw_linsys_write_preassembly_files.argtypes = [vp, C.c_char_p, C.c_char_p,
                                                    C.c_int, C.c_int]
    L.nw_assemble_wall_dist_node.argtypes = [vp, C.c_int]
    L.nw_linsys_apply_dirichlet_bcs.argtypes = [vp, C.c_int, C.c_int, C.c_int64, vp]
    L.nw_linsys_load_complete.argtypes = [vp]
    L.nw_linsys_device_arrays.argtypes = [vp, C.POINTER(vp), C.POINTER(vp),
                                          C.POINTER(C.c_int64)]
    L.nw_linsys_get_values.argtypes = [vp, vp, vp]
    L.nw_linsys_rhs_norm2.argtypes = [vp, c_f64p]
    L.nw_linsys_rhs_norm2_global.argtypes = [vp, c_f64p]
    L.nw_mesh_halo_send_count.argtypes = [vp, C.c_int, c_i64p]
    L.nw_mesh_halo_get_send.argtypes = [vp, C.c_int, c_i64p]
    L.nw_mesh_halo_set_recv.argtypes = [vp, C.c_int, C.c_int64, c_i64p]
    L.nw_mesh_halo_commit.argtypes = [vp]
    L.nw_field_parallel_sum.argtypes = [vp, C.c_int]
    L.nw_field_periodic_update.argtypes = [vp, C.c_int]
    L.nw_momentum_diag_post_process.argtypes = [
        vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    L.nw_field_copy_owned_to_shared.argtypes = [vp, C.c_int]
    L.nw_linsys_halo_send_info.argtypes = [vp, C.c_int, c_i64p, c_i64p]
    L.nw_linsys_halo_get_send.argtypes = [vp, C.c_int, c_i64p, c_i64p, c_i64p]
    L.nw_linsys_halo_set_recv.argtypes = [vp, C.c_int, C.c_int64, c_i64p,
                                          c_i64p, c_i64p]
    L.nw_linsys_halo_commit.argtypes = [vp]
    L.nw_linsys_halo_get_recv_slots.argtypes = [vp, C.c_int, c_i64p, c_i64p,
                                                c_i64p, c_i64p]
    L.nw_linsys_get_extra.argtypes = [vp, c_i64p, c_i64p, c_i64p]
    _lib = L
    return L


def _chk(rc):
    if rc != 0:
        raise NwError("nw error %d: %s" % (rc, lib().nw_last_error().decode()))


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def peclet_fn(form="classic", a=0.0, b=1.0):
    return PecletFn(NW_PECLET_CLASSIC if form == "classic" else NW_PECLET_TANH,
                    float(a), float(b))


class Context:
    """nw_ctx: device + stream (+ NCCL communicator). device=-1: host-only."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _chk(lib().nw_ctx_create(int(device), C.byref(self.h)))
        self.device = device

    def sync(self):
        _chk(lib().nw_ctx_sync(self.h))

    def stream(self):
        return lib().nw_ctx_stream(self.h)

    def comm_init(self, unique_id_bytes, nranks, rank):
        buf = C.create_string_buffer(bytes(unique_id_bytes), 128)
        _chk(lib().nw_ctx_comm_init(self.h, buf, nranks, rank))

    def debug_skip_exchange(self, on):
        """measurement aid: halo exchanges become no-ops while on"""
        _chk(lib().nw_debug_skip_exchange(self.h, 1 if on else 0))

    def join_comm(self):
        """order the compute stream after all halo exchanges issued so far"""
        _chk(lib().nw_ctx_join_comm(self.h))

    def peer_memory(self):
        """True if the NVLink peer-memory mailbox is up (else NCCL send/recv)"""
        return bool(lib().nw_ctx_peer_memory(self.h))

    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        _chk(lib().nw_comm_unique_id(buf))
        return buf.raw

    def close(self):
        if self.h:
            lib().nw_ctx_destroy(self.h)
            self.h = C.c_void_p()


class Mesh:
    """one rank's mesh partition + its fields (Realm / BulkData stand-in)"""

    def __init__(self, ctx, ndim, edge_nodes, node_hypre_id, coords,
                 hypre_offsets=None, node_own_hypre_id=None, rank=0, nranks=1,
                 tile_nodes=0):
        self.ctx = ctx
        self.ndim = ndim
        self._edge_nodes = np.ascontiguousarray(edge_nodes, dtype=np.int32).reshape(-1)
        self._hid = np.ascontiguousarray(node_hypre_id, dtype=np.int64)
        self._coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.n_nodes = self._hid.size
        self.n_edges = self._edge_nodes.size // 2
        if hypre_offsets is None:
            hypre_offsets = [0, self.n_nodes]
        self._off = np.ascontiguousarray(hypre_offsets, dtype=np.int64)
        self._own = None
        if node_own_hypre_id is not None:
            self._own = np.ascontiguousarray(node_own_hypre_id, dtype=np.int64)
        d = MeshDesc()
        d.ndim, d.rank, d.nranks = ndim, rank, nranks
        d.n_nodes, d.n_edges = self.n_nodes, self.n_edges
        d.edge_nodes = self._edge_nodes.ctypes.data_as(c_i32p)
        d.node_hypre_id = self._hid.ctypes.data_as(c_i64p)
        d.node_own_hypre_id = (self._own.ctypes.data_as(c_i64p)
                               if self._own is not None else None)
        d.hypre_offsets = self._off.ctypes.data_as(c_i64p)
        d.coords = self._coords.ctypes.data_as(c_f64p)
        d.tile_nodes = int(tile_nodes)
        self.h = C.c_void_p()
        _chk(lib().nw_mesh_create(ctx.h, C.byref(d), C.byref(self.h)))
        self._fields = {}

    def stats(self):
        s = MeshStats()
        _chk(lib().nw_mesh_get_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in MeshStats._fields_}

    def node_permutation(self):
        n = C.c_int64()
        _chk(lib().nw_mesh_get_node_permutation(self.h, C.byref(n), None))
        perm = np.zeros(n.value, dtype=np.int32)
        _chk(lib().nw_mesh_get_node_permutation(
            self.h, C.byref(n), perm.ctypes.data_as(c_i32p)))
        return perm

    def register(self, name, rank, ncomp):
        fid = C.c_int()
        _chk(lib().nw_field_register(self.h, name.encode(), rank, ncomp,
                                     C.byref(fid)))
        self._fields[name] = (fid.value, rank, ncomp)
        return fid.value

    def field_id(self, name):
        fid = C.c_int()
        _chk(lib().nw_field_find(self.h, name.encode(), C.byref(fid)))
        return fid.value

    def upload(self, name_or_id, host):
        fid = name_or_id if isinstance(name_or_id, int) else self.field_id(name_or_id)
        a = np.ascontiguousarray(host, dtype=np.float64)
        _chk(lib().nw_field_upload(self.h, fid, _ptr(a)))
        # keep the host buffer alive until the (possibly async) copy is done
        self.ctx.sync()

    def upload_ptr(self, fid, ptr):
        """raw pointer variant (pinned host memory, asynchronous)"""
        _chk(lib().nw_field_upload(self.h, fid, C.c_void_p(ptr)))

    def stage_ptr(self, fid, ptr):
        """H2D on the copy stream into the field's staging buffer (pinned
        host memory); overlaps the compute stream until commit()"""
        _chk(lib().nw_field_stage(self.h, fid, C.c_void_p(ptr)))

    def commit(self, fid):
        """compute stream waits for the staged copy and permutes it in"""
        _chk(lib().nw_field_commit(self.h, fid))

    def put(self, name, rank, host):
        """register (shape from the array) + upload"""
        a = np.ascontiguousarray(host, dtype=np.float64)
        n_ent = self.n_nodes if rank == NW_NODE else self.n_edges
        ncomp = a.size // max(n_ent, 1)
        fid = self.register(name, rank, ncomp)
        self.upload(fid, a)
        return fid

    def download(self, name_or_id):
        fid = name_or_id if isinstance(name_or_id, int) else self.field_id(name_or_id)
        name = [k for k, v in self._fields.items() if v[0] == fid]
        _, rank, ncomp = self._fields[name[0]]
        n_ent = self.n_nodes if rank == NW_NODE else self.n_edges
        out = np.zeros((n_ent, ncomp))
        _chk(lib().nw_field_download(self.h, fid, _ptr(out)))
        return out if ncomp > 1 else out.reshape(-1)

    def fill(self, name, value):
        _chk(lib().nw_field_fill(self.h, self.field_id(name), float(value)))

    # --- edge algorithms without a linear system ---
    def mdot_edge(self, noc_fac=1.0, interp_together=1.0):
        o = MdotOpts(noc_fac, interp_together)
        _chk(lib().nw_mdot_edge(self.h, C.byref(o)))

    def peclet_edge(self, viscosity="viscosity", pf=None, eps=1e-16):
        o = PecletOpts(pf or peclet_fn(), eps)
        _chk(lib().nw_peclet_edge(self.h, self.field_id(viscosity), C.byref(o)))

    def geometry_interior_hex8(self, elem_nodes, dnv=None, area=None,
                               coords="coordinates", elem_owned=None):
        """GeometryInteriorAlg for one element block: Hex8 ([n][8]), Tet4
        ([n][4]), Wed6 ([n][6]), Pyr5 ([n][5]); on a 2-D mesh Quad4 ([n][4]).
        Accumulates dual nodal volumes / edge area vectors (zero the fields
        first)"""
        el = np.ascontiguousarray(elem_nodes, dtype=np.int32)
        ow = None if elem_owned is None else np.ascontiguousarray(
            elem_owned, dtype=np.uint8)
        name = "quad4" if self.ndim == 2 else {
            8: "hex8", 4: "tet4", 6: "wed6", 5: "pyr5"}[el.shape[1]]
        fn = getattr(lib(), "nw_geometry_interior_" + name)
        _chk(fn(
            self.h, len(el), _ptr(el), None if ow is None else _ptr(ow),
            self.field_id(coords), -1 if dnv is None else self.field_id(dnv),
            -1 if area is None else self.field_id(area)))

    geometry_interior = geometry_interior_hex8

    def extra_opts(self, gravity=None, source=None, source_mask=None,
                   edge_face_vel_mag=None):
        """nw_mdot_extra_opts: balanced forcing if gravity is given, GCL if the
        edge_face_velocity_mag field name is given"""
        x = MdotExtraOpts()
        x.add_balanced_forcing = int(gravity is not None)
        for d, gv in enumerate(gravity or ()):
            x.gravity[d] = float(gv)
        x.source_field = self.field_id(source) if source else -1
        x.source_mask_field = self.field_id(source_mask) if source_mask else -1
        x.needs_gcl = int(edge_face_vel_mag is not None)
        x.edge_face_vel_mag_field = (self.field_id(edge_face_vel_mag)
                                     if edge_face_vel_mag else -1)
        return x

    def mdot_edge_ext(self, extra, noc_fac=1.0, interp_together=1.0):
        o = MdotOpts(noc_fac, interp_together)
        _chk(lib().nw_mdot_edge_ext(self.h, C.byref(o), C.byref(extra)))

    def nodal_grad_edge(self, phi, grad):
        _chk(lib().nw_nodal_grad_edge(self.h, self.field_id(phi),
                                      self.field_id(grad)))

    def nodal_grad_edge_pair(self, phi_a, grad_a, phi_b, grad_b):
        """two scalar gradients in one launch (SST: dkdx, dwdx)"""
        _chk(lib().nw_nodal_grad_edge_pair(
            self.h, self.field_id(phi_a), self.field_id(grad_a),
            self.field_id(phi_b), self.field_id(grad_b)))

    # --- shared-node exchange lists (caller-side transport) ---
    def halo_send(self, peer):
        n = C.c_int64()
        _chk(lib().nw_mesh_halo_send_count(self.h, peer, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int64)
        _chk(lib().nw_mesh_halo_get_send(self.h, peer,
                                         out.ctypes.data_as(c_i64p)))
        return out

    def halo_set_recv(self, peer, own_hids):
        a = np.ascontiguousarray(own_hids, dtype=np.int64)
        _chk(lib().nw_mesh_halo_set_recv(self.h, peer, a.size,
                                         a.ctypes.data_as(c_i64p)))

    def halo_commit(self):
        _chk(lib().nw_mesh_halo_commit(self.h))

    def parallel_sum(self, name):
        _chk(lib().nw_field_parallel_sum(self.h, self.field_id(name)))

    def periodic_update(self, name):
        """Realm::periodic_field_update: master + slaves summed, on every copy"""
        _chk(lib().nw_field_periodic_update(self.h, self.field_id(name)))

    def copy_owned_to_shared(self, name):
        _chk(lib().nw_field_copy_owned_to_shared(self.h, self.field_id(name)))

    def momentum_diag_post_process(self, dt, gamma1, alpha_u,
                                   udiag="momentum_diag", density="density",
                                   dnv="dual_nodal_volume"):
        """udiag post-processing of MomentumEquationSystem::assemble_and_solve
        (src/LowMachEquationSystem.C:2759-2821)"""
        _chk(lib().nw_momentum_diag_post_process(
            self.h, self.field_id(udiag), self.field_id(density),
            self.field_id(dnv), dt, gamma1, alpha_u))

    def halo_transport(self):
        return ["none", "nccl", "peer_memory"][lib().nw_mesh_halo_transport(self.h)]

    def close(self):
        if self.h:
            lib().nw_mesh_destroy(self.h)
            self.h = C.c_void_p()


class LinearSystem:
    """LinearSystem / HypreLinearSystem / HypreUVWLinearSystem"""

    def __init__(self, mesh, kind=NW_LINSYS_HYPRE, num_dof=1):
        self.mesh = mesh
        self.kind = kind
        self.h = C.c_void_p()
        _chk(lib().nw_linsys_create(mesh.h, kind, num_dof, C.byref(self.h)))
        self.sizes = None

    def set_skipped_rows(self, rows):
        r = np.ascontiguousarray(rows, dtype=np.int64)
        _chk(lib().nw_linsys_set_skipped_rows(
            self.h, r.ctypes.data_as(c_i64p), r.size))

    def buildEdgeToNodeGraph(self):
        _chk(lib().nw_linsys_build_edge_to_node_graph(self.h))

    def finalizeLinearSystem(self):
        _chk(lib().nw_linsys_finalize(self.h))
        s = LinsysSizes()
        _chk(lib().nw_linsys_get_sizes(self.h, C.byref(s)))
        self.sizes = s
        return self

    def graph(self):
        s = self.sizes
        nnz = s.num_nonzeros_owned + s.num_nonzeros_shared
        out = dict(
            row_start_owned=np.zeros(s.num_rows_owned + 1, dtype=np.int64),
            row_start_shared=np.zeros(s.num_rows_shared + 1, dtype=np.int64),
            cols=np.zeros(nnz, dtype=np.int64),
            rows=np.zeros(nnz, dtype=np.int64),
            row_indices_shared=np.zeros(s.num_rows_shared, dtype=np.int64),
            periodic_rows=np.zeros(s.num_periodic_rows, dtype=np.int64))
        _chk(lib().nw_linsys_get_graph(
            self.h, *[out[k].ctypes.data_as(c_i64p) for k in (
                "row_start_owned", "row_start_shared", "cols", "rows",
                "row_indices_shared", "periodic_rows")]))
        return out

    def edge_slots(self):
        b = self.sizes.block
        ne = self.mesh.n_edges
        slots = np.zeros((ne, b, b), dtype=np.int64)
        rows = np.zeros((ne, b), dtype=np.int64)
        _chk(lib().nw_linsys_get_edge_slots(
            self.h, slots.ctypes.data_as(c_i64p), rows.ctypes.data_as(c_i64p)))
        return slots, rows

    def zeroSystem(self):
        _chk(lib().nw_linsys_zero(self.h))

    def set_scatter_mode(self, mode):
        _chk(lib().nw_linsys_set_scatter_mode(self.h, mode))

    def uses_tile_path(self):
        return bool(lib().nw_linsys_uses_tile_path(self.h))

    def set_eager_exchange(self, on=True):
        """The edge assembly is the last contribution to shared rows before
        loadComplete: the assembly kernel stores them into the owners' windows."""
        _chk(lib().nw_linsys_set_eager_exchange(self.h, 1 if on else 0))

    def assemble_continuity_edge(self, dt=1.0, gamma1=1.0, noc_fac=1.0,
                                 interp_together=1.0, solve_incompressible=0.0):
        o = ContinuityOpts(dt, gamma1, noc_fac, interp_together,
                           solve_incompressible)
        _chk(lib().nw_assemble_continuity_edge(self.h, C.byref(o)))

    def assemble_continuity_edge_ext(self, extra, dt=1.0, gamma1=1.0,
                                     noc_fac=1.0, interp_together=1.0,
                                     solve_incompressible=0.0):
        o = ContinuityOpts(dt, gamma1, noc_fac, interp_together,
                           solve_incompressible)
        _chk(lib().nw_assemble_continuity_edge_ext(self.h, C.byref(o),
                                                   C.byref(extra)))

    def assemble_scalar_edge(self, q, dqdx, dflux, alpha=0.0, alpha_upw=1.0,
                             ho_upwind=1.0, relax_fac=1.0, use_limiter=False,
                             eps=1e-16, pf=None):
        o = ScalarOpts(alpha, alpha_upw, ho_upwind, relax_fac,
                       1 if use_limiter else 0, eps, pf or peclet_fn())
        m = self.mesh
        _chk(lib().nw_assemble_scalar_edge(
            self.h, m.field_id(q), m.field_id(dqdx), m.field_id(dflux),
            C.byref(o)))

    def assemble_scalar_edge_pair(self, q, dqdx, dflux, other, q_b, dqdx_b,
                                  dflux_b, opts=None, opts_b=None):
        """this system and `other` (same graph) in one launch; opts / opts_b:
        dicts of assemble_scalar_edge keyword options"""
        def mk(o):
            o = dict(o or {})
            return ScalarOpts(o.get("alpha", 0.0), o.get("alpha_upw", 1.0),
                              o.get("ho_upwind", 1.0), o.get("relax_fac", 1.0),
                              1 if o.get("use_limiter", False) else 0,
                              o.get("eps", 1e-16), o.get("pf") or peclet_fn())
        oa, ob = mk(opts), mk(opts_b if opts_b is not None else opts)
        m = self.mesh
        _chk(lib().nw_assemble_scalar_edge_pair(
            self.h, m.field_id(q), m.field_id(dqdx), m.field_id(dflux),
            C.byref(oa), other.h, m.field_id(q_b), m.field_id(dqdx_b),
            m.field_id(dflux_b), C.byref(ob)))

    def assemble_momentum_edge(self, viscosity="viscosity", include_divu=0.0,
                               alpha=0.0, alpha_upw=1.0, ho_upwind=1.0,
                               relax_fac=1.0, use_limiter=False, eps=1e-16,
                               fuse_peclet=False, pf=None, pec_eps=1e-16,
                               diag_field=None, has_vof=False):
        m = self.mesh
        o = MomentumOpts(include_divu, alpha, alpha_upw, ho_upwind, relax_fac,
                         1 if use_limiter else 0, eps, 1 if fuse_peclet else 0,
                         pf or peclet_fn(), pec_eps,
                         m.field_id(diag_field) if diag_field else -1,
                         1 if has_vof else 0)
        _chk(lib().nw_assemble_momentum_edge(
            self.h, m.field_id(viscosity), C.byref(o)))

    def assemble_mass_bdf_node(self, kind, dt, gammas, q=None, rho=None,
                               dnv=None, dpdx=None):
        """time-derivative node kernel; q / rho / dnv = (NM1, N, NP1) field names"""
        fid = self.mesh.field_id
        o = MassBdfOpts()
        o.dt, (o.gamma1, o.gamma2, o.gamma3) = dt, gammas
        if q is not None:
            o.q_nm1, o.q_n, o.q_np1 = (fid(x) for x in q)
        o.rho_nm1, o.rho_n, o.rho_np1 = (fid(x) for x in rho)
        o.dnv_nm1, o.dnv_n, o.dnv_np1 = (fid(x) for x in dnv)
        o.dpdx = fid(dpdx) if dpdx is not None else -1
        _chk(lib().nw_assemble_mass_bdf_node(self.h, kind, C.byref(o)))

    def assemble_wall_dist_edge(self):
        _chk(lib().nw_assemble_wall_dist_edge(self.h))

    def assemble_wall_dist_node(self, dnv="dual_nodal_volume"):
        _chk(lib().nw_assemble_wall_dist_node(self.h, self.mesh.field_id(dnv)))

    def sumInto(self, entity_nodes, lhs, rhs):
        """generic CoeffApplier::operator(): entity_nodes [nEnt][npe] local node
        indices, lhs [nEnt][n][n], rhs [nEnt][n] (host arrays; copied to the
        device for the call)"""
        import torch
        en = np.ascontiguousarray(entity_nodes, dtype=np.int32)
        d_en = torch.from_numpy(en).cuda()
        d_l = torch.from_numpy(np.ascontiguousarray(lhs, dtype=np.float64)).cuda()
        d_r = torch.from_numpy(np.ascontiguousarray(rhs, dtype=np.float64)).cuda()
        torch.cuda.synchronize()
        _chk(lib().nw_linsys_sum_into(
            self.h, en.shape[0], en.shape[1], C.c_void_p(d_en.data_ptr()),
            C.c_void_p(d_l.data_ptr()), C.c_void_p(d_r.data_ptr())))
        self.mesh.ctx.sync()

    def resetRows(self, nodes, diag_value=0.0, rhs_residual=0.0):
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        _chk(lib().nw_linsys_reset_rows(self.h, nd.size, _ptr(nd), diag_value,
                                        rhs_residual))

    def applyDirichletBCs(self, solution, bc_values, nodes):
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        _chk(lib().nw_linsys_apply_dirichlet_bcs(
            self.h, self.mesh.field_id(solution), self.mesh.field_id(bc_values),
            nd.size, _ptr(nd)))

    def write_preassembly_files(self, directory, eq_sys_name, write_counter=1,
                                hypre_int_bytes=4):
        _chk(lib().nw_linsys_write_preassembly_files(
            self.h, str(directory).encode(), eq_sys_name.encode(),
            write_counter, hypre_int_bytes))

    def loadComplete(self):
        _chk(lib().nw_linsys_load_complete(self.h))

    def halo_transport(self):
        return ["none", "nccl", "peer_memory"][lib().nw_linsys_halo_transport(self.h)]

    # --- shared-row exchange structure (caller-side transport) ---
    def halo_send(self, peer):
        nr, nv = C.c_int64(), C.c_int64()
        _chk(lib().nw_linsys_halo_send_info(self.h, peer, C.byref(nr),
                                            C.byref(nv)))
        rows = np.zeros(nr.value, dtype=np.int64)
        lens = np.zeros(nr.value, dtype=np.int64)
        cols = np.zeros(nv.value, dtype=np.int64)
        _chk(lib().nw_linsys_halo_get_send(
            self.h, peer, rows.ctypes.data_as(c_i64p),
            lens.ctypes.data_as(c_i64p), cols.ctypes.data_as(c_i64p)))
        return rows, lens, cols

    def halo_set_recv(self, peer, rows, lens, cols):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        _chk(lib().nw_linsys_halo_set_recv(
            self.h, peer, rows.size, rows.ctypes.data_as(c_i64p),
            lens.ctypes.data_as(c_i64p), cols.ctypes.data_as(c_i64p)))

    def halo_commit(self):
        _chk(lib().nw_linsys_halo_commit(self.h))

    def halo_recv_slots(self, peer):
        nv, nr = C.c_int64(), C.c_int64()
        _chk(lib().nw_linsys_halo_get_recv_slots(
            self.h, peer, C.byref(nv), None, C.byref(nr), None))
        vs = np.zeros(nv.value, dtype=np.int64)
        rr = np.zeros(nr.value, dtype=np.int64)
        _chk(lib().nw_linsys_halo_get_recv_slots(
            self.h, peer, C.byref(nv), vs.ctypes.data_as(c_i64p), C.byref(nr),
            rr.ctypes.data_as(c_i64p)))
        return vs, rr

    def extra(self):
        n = C.c_int64()
        _chk(lib().nw_linsys_get_extra(self.h, C.byref(n), None, None))
        rows = np.zeros(n.value, dtype=np.int64)
        cols = np.zeros(n.value, dtype=np.int64)
        _chk(lib().nw_linsys_get_extra(self.h, C.byref(n),
                                       rows.ctypes.data_as(c_i64p),
                                       cols.ctypes.data_as(c_i64p)))
        return rows, cols

    def values(self):
        s = self.sizes
        nnz = s.num_nonzeros_owned + s.num_nonzeros_shared
        nnz += self.extra()[0].size
        rows = s.num_rows_owned + s.num_rows_shared
        vals = np.zeros(nnz)
        rhs = np.zeros((s.num_rhs, rows))
        _chk(lib().nw_linsys_get_values(self.h, _ptr(vals), _ptr(rhs)))
        return vals, rhs

    def rhs_norm2_global(self):
        """summed over all ranks (collective)"""
        out = np.zeros(self.sizes.num_rhs)
        _chk(lib().nw_linsys_rhs_norm2_global(self.h, out.ctypes.data_as(c_f64p)))
        return out

    def rhs_norm2(self):
        out = np.zeros(self.sizes.num_rhs)
        _chk(lib().nw_linsys_rhs_norm2(self.h, out.ctypes.data_as(c_f64p)))
        return out

    def close(self):
        if self.h:
            lib().nw_linsys_destroy(self.h)
            self.h = C.c_void_p()


# ---------------------------------------------------------------------------
# synthetic mesh generator (tests / bench input only)
# ---------------------------------------------------------------------------

class _MgParams(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
                ("periodic_x", C.c_int32), ("periodic_y", C.c_int32),
                ("warp", C.c_double), ("zstretch", C.c_double),
                ("nranks", C.c_int32), ("rank", C.c_int32),
                ("shuffle_bucket", C.c_int32), ("seed", C.c_uint64)]


_mg = None


def _meshgen():
    global _mg
    if _mg is None:
        if not os.path.exists(MESHGEN_PATH):
            raise NwError("%s is missing: run __graft_entry__.build()" %
                          MESHGEN_PATH)
        L = C.CDLL(MESHGEN_PATH)
        L.mg_generate.restype = C.c_void_p
        L.mg_generate.argtypes = [C.POINTER(_MgParams)]
        L.mg_free.argtypes = [C.c_void_p]
        L.mg_num_nodes.restype = C.c_int64
        L.mg_num_nodes.argtypes = [C.c_void_p]
        L.mg_num_edges.restype = C.c_int64
        L.mg_num_edges.argtypes = [C.c_void_p]
        L.mg_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        _mg = L
    return _mg


class BoxMesh:
    """one rank's part of a generated nx*ny*nz hex box (z-slab decomposition)"""

    def __init__(self, nx, ny, nz, lengths=None, periodic=(False, False),
                 warp=0.0, zstretch=1.0, nranks=1, rank=0, shuffle_bucket=0,
                 seed=20261017):
        L = _meshgen()
        lx, ly, lz = lengths if lengths else (float(nx), float(ny), float(nz))
        p = _MgParams(nx, ny, nz, lx, ly, lz, int(periodic[0]), int(periodic[1]),
                      float(warp), float(zstretch), nranks, rank,
                      int(shuffle_bucket), int(seed))
        h = L.mg_generate(C.byref(p))
        try:
            n, e = L.mg_num_nodes(h), L.mg_num_edges(h)
            self.n_nodes, self.n_edges = n, e
            self.coords = np.zeros((n, 3))
            self.gid = np.zeros(n, dtype=np.int64)
            self.hid = np.zeros(n, dtype=np.int64)
            self.own_hid = np.zeros(n, dtype=np.int64)
            self.owner = np.zeros(n, dtype=np.int32)
            self.offsets = np.zeros(nranks + 1, dtype=np.int64)
            self.edges = np.zeros((e, 2), dtype=np.int32)
            self.area = np.zeros((e, 3))
            self.vol = np.zeros(n)
            L.mg_copy(h, *[_ptr(a) for a in (
                self.coords, self.gid, self.hid, self.own_hid, self.owner,
                self.offsets, self.edges, self.area, self.vol)])
        finally:
            L.mg_free(h)
        self.rank, self.nranks = rank, nranks
        self.dims = (nx, ny, nz)
        self.periodic = tuple(periodic)

    def make_mesh(self, ctx, tile_nodes=0):
        own = self.own_hid if any(self.periodic) else None
        m = Mesh(ctx, 3, self.edges, self.hid, self.coords,
                 hypre_offsets=self.offsets, node_own_hypre_id=own,
                 rank=self.rank, nranks=self.nranks, tile_nodes=tile_nodes)
        return m
