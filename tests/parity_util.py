"""Shared helpers of the parity tests: build a synthetic case, run the CPU
oracle on it, and compare against the product (CUDA through the C ABI) or the
CPU plan walk-through (tests/emul).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_py as orc  # noqa: E402
import __graft_entry__ as graft  # noqa: E402

TOL = 1.0e-12  # north_star: every fp64 matrix / RHS entry within relative 1e-12

DT, GAMMA1 = 0.5, 1.5


def pkg():
    return graft.load_package()


def scaled_err(got, ref, scale):
    """max |got-ref| / (TOL * scale): < 1 passes.  `scale` is the oracle's sum of
    |contributions| of the entry (a rigorous cancellation-aware magnitude), with
    |ref| as a floor."""
    got, ref, scale = np.asarray(got), np.asarray(ref), np.asarray(scale)
    if got.shape != ref.shape:
        return float("inf")
    if not np.all(np.isfinite(got)):
        return float("inf")
    s = np.maximum(np.maximum(scale, np.abs(ref)), 1e-300)
    if got.size == 0:
        return 0.0
    return float(np.max(np.abs(got - ref) / (TOL * s)))


# worst plain relative error |got - ref| / |ref| (ref != 0) of the last
# run_lowmach_case, per quantity: reported beside the cancellation-aware scaled
# error; entries that are themselves the remainder of a
# cancellation make this number larger than 1e-12 without being wrong
LAST_PLAIN = {}


def plain_rel(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if got.shape != ref.shape or got.size == 0:
        return float("nan")
    nz = ref != 0.0
    if not np.any(nz):
        return 0.0
    return float(np.max(np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])))


def periodic_field_update(field, hid, own_hid):
    """Realm::periodic_field_update (src/Realm.C:3090-3100 ->
    PeriodicManager::add_slave_to_master + set_slave_to_master,
    src/PeriodicManager.C:1007-1058, 1139-1190) restated for the checker: the
    nodes sharing one resolved row id end with master + slaves (slaves added
    in ascending own id) on every copy.  field: [n_nodes][ncomp]"""
    f = np.array(field, dtype=np.float64).reshape(len(hid), -1)
    hid, own = np.asarray(hid), np.asarray(own_hid)
    slave_rows = np.unique(hid[own != hid])
    if slave_rows.size == 0:
        return f.reshape(np.shape(field))
    members = np.nonzero(np.isin(hid, slave_rows))[0]
    order = np.lexsort((own[members], own[members] != hid[members], hid[members]))
    members = members[order]
    starts = np.nonzero(np.r_[True, hid[members][1:] != hid[members][:-1]])[0]
    ends = np.r_[starts[1:], len(members)]
    for a, b in zip(starts, ends):
        grp = members[a:b]
        assert own[grp[0]] == hid[grp[0]], "master missing"
        tot = f[grp[0]].copy()
        for n in grp[1:]:
            tot = tot + f[n]
        f[grp] = tot
    return f.reshape(np.shape(field))


def momentum_diag_post_process(udiag, rho, dvol, hid, own_hid, dt, gamma1,
                               alpha_u, lo=None, hi=None):
    """The udiag post-processing of MomentumEquationSystem::assemble_and_solve
    (src/LowMachEquationSystem.C:2776-2808) restated for the checker, one rank
    (the parallel_sum / copy_owned_to_shared around it are exchanges): on the
    locally owned nodes that are not periodic slaves
        udiag = (udiag / (rho * dualVol) - gamma1/dt) * alphaU + gamma1/dt,
    then every periodic slave takes its master's value (apply_constraints with
    setSlaves only).  numpy rounds every operation on its own, like the
    reference's builds."""
    u = np.array(udiag, dtype=np.float64)
    hid, own = np.asarray(hid), np.asarray(own_hid)
    sel = own == hid
    if lo is not None:
        sel &= (own >= lo) & (own <= hi)
    pts = gamma1 / dt
    tmp = u[sel] / (np.asarray(rho)[sel] * np.asarray(dvol)[sel])
    u[sel] = (tmp - pts) * alpha_u + pts
    master = {int(hid[n]): n for n in np.nonzero(own == hid)[0]}
    for n in np.nonzero(own != hid)[0]:
        u[n] = u[master[int(hid[n])]]
    return u


def add_tet_split_edges(b, seed=20261017):
    """Turn the hex box's edge graph into that of its 6-tet (Kuhn) split: every
    cell gains its three face diagonals towards (+,+,0), (+,0,+), (0,+,+) and
    the body diagonal (+,+,+) -- the connectivity of the reference's tet /
    mixed-element meshes (BASELINE configs[4]; up to 14 neighbours per node,
    ragged rows, edges that are not grid-aligned).  Single rank, non-periodic.
    The new edges get synthetic area vectors (mostly along the edge, plus a
    seeded transverse part), which is all the edge kernels see of geometry."""
    assert b.nranks == 1 and not any(b.periodic)
    nx, ny, nz = b.dims
    # lattice index of every node from its global id (1 + i + (nx+1) j + ...)
    g0 = b.gid - 1
    i, j, k = g0 % (nx + 1), (g0 // (nx + 1)) % (ny + 1), g0 // ((nx + 1) * (ny + 1))
    node_at = -np.ones((nx + 1, ny + 1, nz + 1), dtype=np.int64)
    node_at[i, j, k] = np.arange(b.n_nodes)
    assert node_at.min() >= 0
    extra = []
    for di, dj, dk in ((1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)):
        a = node_at[:nx + 1 - di, :ny + 1 - dj, :nz + 1 - dk].ravel()
        c = node_at[di:, dj:, dk:].ravel()
        extra.append(np.stack([a, c], axis=1))
    extra = np.concatenate(extra)
    # reference orientation: L = lower global id (STK edge creation)
    swap = b.gid[extra[:, 0]] > b.gid[extra[:, 1]]
    extra[swap] = extra[swap][:, ::-1]
    rng = np.random.default_rng(seed)
    dx = b.coords[extra[:, 1]] - b.coords[extra[:, 0]]
    ln = np.linalg.norm(dx, axis=1, keepdims=True)
    area = 0.35 * dx / ln * ln + 0.05 * ln * rng.standard_normal(dx.shape)
    order = rng.permutation(len(extra))  # irregular edge ordering
    b.edges = np.ascontiguousarray(
        np.concatenate([b.edges, extra[order].astype(np.int32)]))
    b.area = np.ascontiguousarray(np.concatenate([b.area, area[order]]))
    b.n_edges = len(b.edges)


def box_hex_elements(b):
    """hex8 connectivity [n_elems][8] (local nodes, Exodus / STK node order) of a
    single-rank, non-periodic BoxMesh, from the generator's global ids"""
    assert b.nranks == 1 and not any(b.periodic)
    nx, ny, nz = b.dims
    g0 = b.gid - 1
    i, j, k = g0 % (nx + 1), (g0 // (nx + 1)) % (ny + 1), g0 // ((nx + 1) * (ny + 1))
    at = -np.ones((nx + 1, ny + 1, nz + 1), dtype=np.int64)
    at[i, j, k] = np.arange(b.n_nodes)
    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    corners = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0),
               (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    return np.stack([at[I + a, J + bb, K + c] for a, bb, c in corners],
                    axis=1).astype(np.int32)


def lhs_scale(g_rows, ov, av, floor=1e-4):
    """Tolerance scale of matrix entries on the reference's real meshes: the
    entry's own sum of |block contributions| (as everywhere else), but an entry
    below `floor` of its row's largest entry is held to that fraction of the
    largest entry instead.  A per-edge block entry can itself be the remainder
    of a cancellation inside the kernel formula (e.g. 0.5 mdot (1 - pecfac) +
    diffusion at Peclet factors of 1 - 1e-8): one ulp of an intermediate term,
    which FMA contraction legitimately moves, is then 1e-11 of the entry while
    being 1e-19 of the row.  DESIGN.md section 4: |got - ref| <= 1e-12 max(|ref|,
    row scale)."""
    rows = np.asarray(g_rows)
    rowmax = np.zeros(int(rows.max()) + 1 if rows.size else 1)
    np.maximum.at(rowmax, rows, np.abs(ov))
    return np.maximum(av, floor * rowmax[rows])


def momentum_flux_scale(edges, mdot, velocity, n_nodes):
    """per node: sum over its edges of |mdot| * max(|u_L|_inf, |u_R|_inf) -- the
    magnitude of the advective momentum flux VECTOR through the node's dual
    faces.  Floor for the rhs tolerance of the VOF branch (same idea as
    lhs_scale): a component whose own flux sum is below 1e-4 of this is held to
    that fraction of it."""
    e = np.asarray(edges).reshape(-1, 2)
    u = np.abs(np.asarray(velocity).reshape(n_nodes, -1)).max(axis=1)
    w = np.abs(np.asarray(mdot)) * np.maximum(u[e[:, 0]], u[e[:, 1]])
    out = np.zeros(n_nodes)
    np.add.at(out, e[:, 0], w)
    np.add.at(out, e[:, 1], w)
    return out


def vof_scales(c, g, mdot_total, ov, av, arhs, num_dof):
    """Tolerance scales of the VOF comparisons.  The
    branch evaluates erf, the one function of the path the device does not
    compute bit for bit like the host (CUDA: <= 2 ulp), and with the decks'
    alphaUpw = 1 it forms  1 - (f + (1 - f))  and  1 - (1 - f + f pecfac):  0 or
    +-1 ulp depending on the last bit of f.  That ulp, times mdot, is the whole
    error of an entry (or rhs component) whose own magnitude is a cancellation
    remainder -- the viscous 2e-5 beside mdot ~ 1e4, the w-momentum flux at a
    node whose upwind w is exactly 0 -- so:
      matrix entries  1e-12 max(sum |contributions|, 1e-3 row max)
      rhs components  1e-12 max(sum |contributions|, 1e-4 sum_e |mdot| |u|_inf)
    (measured with the +-2 ulp build: 0.35 and 0.01 of these; 5e4 and 9e2 of
    the plain scale)."""
    lsc = lhs_scale(g.rows - g.i_lower, ov, av, floor=1e-3)
    flux = momentum_flux_scale(c.edges, mdot_total, c.fields["velocity"],
                                  c.n_nodes)
    per_row = np.zeros(g.num_rows_owned // num_dof)
    per_row[c.box.hid - g.i_lower // num_dof] = flux
    if num_dof == 1:   # UVW: rhs[d][row]
        rsc = np.maximum(arhs, 1e-4 * per_row[None, :arhs.shape[1]])
    else:              # monolithic: rhs[row * ndof + d]
        rsc = np.maximum(arhs.ravel(), 1e-4 * np.repeat(per_row, num_dof))
    return lsc, rsc


class Case:
    """generated hex box + synthetic state (one rank)"""

    def __init__(self, dims=(12, 10, 8), lengths=None, periodic=(False, False),
                 warp=0.0, zstretch=1.0, nranks=1, rank=0, shuffle_bucket=0,
                 tet_split=False):
        P = pkg()
        synth = __import__("nalu_wind_b200.synth", fromlist=["state"])
        self.box = P.BoxMesh(*dims, lengths=lengths, periodic=periodic,
                             warp=warp, zstretch=zstretch, nranks=nranks,
                             rank=rank, shuffle_bucket=shuffle_bucket)
        b = self.box
        if tet_split:
            add_tet_split_edges(b)
        L = lengths if lengths else tuple(float(d) for d in dims)
        self.lengths = L
        pg = None
        if any(periodic):
            # periodic master's global id: own_hid differs from hid on slaves;
            # use the resolved row id as the noise key (identical on aliases)
            pg = b.hid.astype(np.int64) + 1
        self.fields = synth.state(b.coords, b.gid, L, DT, GAMMA1,
                                  periodic_gid=pg)
        self.fields["dual_nodal_volume"] = b.vol
        self.edges = b.edges
        self.area = b.area
        self.n_nodes, self.n_edges = b.n_nodes, b.n_edges

    # ---- oracle side ----
    def oracle_graph(self, num_dof=1, skipped=()):
        b = self.box
        lo = int(b.offsets[b.rank]) * num_dof
        hi = int(b.offsets[b.rank + 1]) * num_dof - 1
        g = orc.Graph(num_dof, lo, hi)
        if len(skipped):
            g.set_skipped(skipped)
        g.add_edges(self.edges, b.hid)
        return g.finalize()

    def oracle_mdot(self):
        f = self.fields
        return orc.mdot_edge(3, self.edges, self.box.coords, f["velocity"],
                             f["dpdx"], f["density"], f["pressure"],
                             f["momentum_diag"], self.area, 1.0, 1.0)

    def oracle_pecfac(self, pf):
        f = self.fields
        return orc.peclet_edge(3, self.edges, self.box.coords, f["velocity"],
                               f["density"], f["viscosity"], pf)[1]


MOM_OPTS = dict(include_divu=0.0, alpha=0.0, alpha_upw=1.0, ho_upwind=1.0,
                relax_fac=0.7, use_limiter=True)
SCAL_OPTS = dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, relax_fac=0.9,
                 use_limiter=True)
CONT_OPTS = dict(dt=DT, gamma1=GAMMA1, noc_fac=1.0, interp_together=1.0,
                 solve_incompressible=0.0)


def oracle_continuity(case, g):
    f, b = case.fields, case.box
    s = orc.HypreSink(g, b.hid)
    orc.continuity_edge(3, case.edges, b.coords, f["velocity"], f["dpdx"],
                        f["density"], f["pressure"], f["momentum_diag"],
                        case.area, s, **CONT_OPTS)
    return s


def oracle_scalar(case, g, mdot, q="turbulent_ke", dq="dkdx",
                  mu="effective_viscosity_tke", pf=None):
    f, b = case.fields, case.box
    s = orc.HypreSink(g, b.hid)
    orc.scalar_edge(3, case.edges, b.coords, f["velocity"], f[q], f[dq],
                    f["density"], f[mu], case.area, mdot, s,
                    pf=pf or orc.peclet("tanh", 2.0, 1.0), **SCAL_OPTS)
    return s


def oracle_momentum(case, g, mdot, pecfac, uvw=True, udiag=None):
    f, b = case.fields, case.box
    s = orc.HypreSink(g, b.hid, uvw_ndim=3 if uvw else 0)
    orc.momentum_edge(3, case.edges, b.coords, f["velocity"], f["dudx"],
                      f["viscosity"], f["density"],
                      f["abl_wall_no_slip_wall_func_node_mask"], case.area,
                      mdot, pecfac, s, udiag_accum=udiag, **MOM_OPTS)
    return s


# ---------------------------------------------------------------------------
# CPU plan walk-through (tests/emul)
# ---------------------------------------------------------------------------
_emu = None
_emu_variants = {}


def host_has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read().replace("\n", " ")
    except OSError:
        return False


def emu_lib(fma=False):
    """fma: the build whose a*b+c are contracted into fused multiply-adds (as
    nvcc contracts them on the device) -- a CPU stand-in for device rounding"""
    global _emu
    if fma:
        # fma == "erf": additionally erf moved by up to +-2 ulp (emul/erf_perturb.h)
        key = "erf" if fma == "erf" else "fma"
        if key not in _emu_variants:
            so = "libnw_emul_%s.so" % key
            subprocess.check_call(["make", "-C", os.path.join(HERE, "emul"), "-s", so])
            L = C.CDLL(os.path.join(HERE, "emul", so))
            base = emu_lib()
            for name in ("emu_create", "emu_error", "emu_destroy",
                         "emu_build_linsys", "emu_check_plan", "emu_assemble_mono",
                         "emu_assemble", "emu_nodal_grad", "emu_mdot",
                         "emu_udiag_post"):
                getattr(L, name).argtypes = getattr(base, name).argtypes
                getattr(L, name).restype = getattr(base, name).restype
            _emu_variants[key] = L
        return _emu_variants[key]
    if _emu is None:
        subprocess.check_call(["make", "-C", os.path.join(HERE, "emul"), "-s"])
        L = C.CDLL(os.path.join(HERE, "emul", "libnw_emul.so"))
        vp = C.c_void_p
        L.emu_create.restype = vp
        L.emu_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64,
                                 C.c_int64, vp, vp, vp, vp, vp, C.c_int]
        L.emu_error.restype = C.c_char_p
        L.emu_error.argtypes = [vp]
        L.emu_destroy.argtypes = [vp]
        L.emu_build_linsys.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int64]
        L.emu_check_plan.argtypes = [vp]
        L.emu_assemble_mono.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp,
                                        C.c_int64, vp, vp]
        L.emu_assemble.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp,
                                   vp, vp, vp]
        L.emu_nodal_grad.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.emu_mdot.argtypes = [vp, vp, vp, C.c_int, vp, C.c_double,
                               C.c_double, vp]
        L.emu_udiag_post.restype = None
        L.emu_udiag_post.argtypes = [C.c_int64, vp, vp, vp, C.c_double, C.c_double]
        _emu = L
    return _emu


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class Emu:
    def __init__(self, case, tile_nodes=64, fma=False):
        b = case.box
        self.case = case
        self.L = L = emu_lib(fma)
        self._keep = [np.ascontiguousarray(b.edges), b.hid, b.own_hid,
                      b.offsets, b.coords]
        self.h = L.emu_create(3, b.rank, b.nranks, b.n_nodes, b.n_edges,
                              *[_p(a) for a in self._keep], tile_nodes)
        err = L.emu_error(self.h).decode()
        assert err == "", err

    def build_linsys(self, kind=0, num_dof=1, skipped=()):
        sk = np.ascontiguousarray(skipped, dtype=np.int64)
        rc = self.L.emu_build_linsys(self.h, kind, num_dof, _p(sk), sk.size)
        assert rc == 0, self.L.emu_error(self.h).decode()

    def check_plan(self):
        rc = self.L.emu_check_plan(self.h)
        assert rc == 0, self.L.emu_error(self.h).decode()

    def _fields(self, names):
        arrs = [np.ascontiguousarray(self.case.fields[n] if n != "coordinates"
                                     else self.case.box.coords,
                                     dtype=np.float64) for n in names]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        nc = np.array([a.size // self.case.n_nodes for a in arrs],
                      dtype=np.int32)
        return arrs, ptrs, nc

    def assemble_mono(self, names, opts, nnz, rows, mdot, pecfac=None,
                      skipped3=()):
        """monolithic 3-dof momentum through the node graph's plan"""
        sk = np.ascontiguousarray(skipped3, dtype=np.int64)
        arrs, ptrs, nc = self._fields(names)
        vals = np.zeros(nnz)
        rhs = np.zeros(rows)
        area = np.ascontiguousarray(self.case.area)
        rc = self.L.emu_assemble_mono(
            self.h, C.cast(ptrs, C.c_void_p), _p(nc), len(arrs), _p(area),
            _p(mdot), _p(pecfac), C.cast(C.byref(opts), C.c_void_p), _p(sk),
            sk.size, _p(vals), _p(rhs))
        assert rc == 0, self.L.emu_error(self.h).decode()
        return vals, rhs

    def assemble(self, kind, names, opts, nnz, rows, nrhs, mdot=None,
                 pecfac=None):
        arrs, ptrs, nc = self._fields(names)
        vals = np.zeros(nnz)
        rhs = np.zeros((nrhs, rows))
        area = np.ascontiguousarray(self.case.area)
        rc = self.L.emu_assemble(
            self.h, kind, C.cast(ptrs, C.c_void_p), _p(nc), len(arrs),
            _p(area), _p(mdot), _p(pecfac), C.cast(C.byref(opts), C.c_void_p),
            _p(vals), _p(rhs))
        assert rc == 0, self.L.emu_error(self.h).decode()
        return vals, rhs

    def nodal_grad(self, phi, dim1):
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        out = np.zeros((self.case.n_nodes, dim1 * 3))
        area = np.ascontiguousarray(self.case.area)
        vol = np.ascontiguousarray(self.case.fields["dual_nodal_volume"])
        rc = self.L.emu_nodal_grad(self.h, dim1, _p(phi), _p(area), _p(vol),
                                      _p(out))
        assert rc == 0
        return out

    def mdot(self):
        arrs, ptrs, nc = self._fields(CONT_FIELDS)
        out = np.zeros(self.case.n_edges)
        area = np.ascontiguousarray(self.case.area)
        rc = self.L.emu_mdot(self.h, C.cast(ptrs, C.c_void_p), _p(nc),
                                len(arrs), _p(area), 1.0, 1.0, _p(out))
        assert rc == 0, self.L.emu_error(self.h).decode()
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.L.emu_destroy(self.h)
            self.h = None


CONT_FIELDS = ["coordinates", "velocity", "dpdx", "density", "pressure",
               "momentum_diag"]
SCAL_FIELDS = ["coordinates", "velocity", "dkdx", "turbulent_ke", "density",
               "effective_viscosity_tke"]
MOM_FIELDS = ["coordinates", "velocity", "dudx", "viscosity", "density",
              "abl_wall_no_slip_wall_func_node_mask"]


# ---------------------------------------------------------------------------
# product (CUDA) side
# ---------------------------------------------------------------------------

def upload_state(P, mesh, case, extra_edge=None):
    for name, arr in case.fields.items():
        mesh.put(name, P.NW_NODE, arr)
    mesh.put("edge_area_vector", P.NW_EDGE, case.area)
    mesh.register("mass_flow_rate", P.NW_EDGE, 1)
    mesh.register("peclet_factor", P.NW_EDGE, 1)
    for k, v in (extra_edge or {}).items():
        mesh.put(k, P.NW_EDGE, v)


def run_lowmach_case(P, ctx, dims=(12, 10, 8), tile_nodes=64, mode=None,
                     case=None, **case_kw):
    """The full low-Mach sweep on one rank through the C ABI, compared with the
    oracle.  Returns {name: scaled error} (every value must be < 1)."""
    real = case is not None
    if case is None:
        case = Case(dims=dims, **case_kw)
    mesh = case.box.make_mesh(ctx, tile_nodes=tile_nodes)
    upload_state(P, mesh, case)
    res = {}
    LAST_PLAIN.clear()
    pf = P.peclet_fn("classic", 1.0)
    opf = orc.peclet("classic", 1.0)

    # K1 mdot
    mesh.mdot_edge(1.0, 1.0)
    mdot = mesh.download("mass_flow_rate")
    omdot = case.oracle_mdot()
    f = case.fields
    res["mdot"] = scaled_err(mdot, omdot, np.abs(omdot) + 1e-3 * np.max(np.abs(omdot)))
    LAST_PLAIN["mdot"] = plain_rel(mdot, omdot)
    # K9 peclet
    mesh.peclet_edge("viscosity", pf)
    pec = mesh.download("peclet_factor")
    opec = case.oracle_pecfac(opf)
    res["peclet"] = scaled_err(pec, opec, np.ones_like(opec))
    LAST_PLAIN["peclet"] = plain_rel(pec, opec)
    # feed the oracle's edge fields back so later kernels see identical bits
    mesh.upload("mass_flow_rate", omdot)
    mesh.upload("peclet_factor", opec)

    # K2 gradients
    for phi, grad, d1 in (("pressure", "dpdx_out", 1), ("velocity", "dudx_out", 3)):
        mesh.register(grad, P.NW_NODE, d1 * 3)
        mesh.nodal_grad_edge(phi, grad)
        got = mesh.download(grad)
        ref = orc.nodal_grad_edge(d1, 3, case.edges, f[phi], case.area,
                                  f["dual_nodal_volume"], case.n_nodes)
        mag = orc.nodal_grad_edge(d1, 3, case.edges, np.abs(f[phi]),
                                  np.abs(case.area), f["dual_nodal_volume"],
                                  case.n_nodes)
        if any(getattr(case.box, "periodic", (False, False))):
            # NodalGradAlgDriver::post_work: periodic_field_update(gradPhi),
            # master and slave copies included in the comparison
            ref = periodic_field_update(ref, case.box.hid, case.box.own_hid)
            mag = periodic_field_update(np.abs(mag), case.box.hid, case.box.own_hid)
        # |.| version over-counts signs of R contributions: use abs of terms
        mag = np.abs(mag) + np.max(np.abs(ref)) * 1e-3
        res["grad_" + phi] = scaled_err(got.reshape(ref.shape), ref, mag)
        LAST_PLAIN["grad_" + phi] = plain_rel(got.reshape(ref.shape), ref)

    g = case.oracle_graph()
    # local row of every stored entry (owned rows; single rank on real meshes)
    lrow = np.asarray(g.rows) - int(g.rows.min()) if real else None
    lscale = (lambda ov, av: lhs_scale(lrow, ov, av)) if real else (lambda ov, av: av)
    # K4 continuity
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    if mode is not None:
        ls.set_scatter_mode(mode)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    ls.zeroSystem()
    ls.assemble_continuity_edge(**CONT_OPTS)
    ls.loadComplete()
    vals, rhs = ls.values()
    o = oracle_continuity(case, g)
    ov, orhs = o.get()
    av, arhs = o.get_abs()
    res["continuity_lhs"] = scaled_err(vals, ov, lscale(ov, av))
    res["continuity_rhs"] = scaled_err(rhs, orhs, arhs)
    LAST_PLAIN["continuity_lhs"] = plain_rel(vals, ov)
    LAST_PLAIN["continuity_rhs"] = plain_rel(rhs, orhs)
    n2 = ls.rhs_norm2()
    res["continuity_norm"] = scaled_err(
        n2, np.sum(orhs[:, :g.num_rows_owned] ** 2, axis=1),
        1e3 * np.sum(arhs[:, :g.num_rows_owned] ** 2, axis=1))
    ls.close()

    # K5 scalar
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    if mode is not None:
        ls.set_scatter_mode(mode)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    ls.zeroSystem()
    ls.assemble_scalar_edge("turbulent_ke", "dkdx", "effective_viscosity_tke",
                            pf=P.peclet_fn("tanh", 2.0, 1.0), **SCAL_OPTS)
    vals, rhs = ls.values()
    o = oracle_scalar(case, g, omdot)
    ov, orhs = o.get()
    av, arhs = o.get_abs()
    res["scalar_lhs"] = scaled_err(vals, ov, lscale(ov, av))
    res["scalar_rhs"] = scaled_err(rhs, orhs, arhs)
    LAST_PLAIN["scalar_lhs"] = plain_rel(vals, ov)
    LAST_PLAIN["scalar_rhs"] = plain_rel(rhs, orhs)
    ls.close()

    # K3 momentum, segregated UVW
    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE_UVW, 3)
    if mode is not None:
        ls.set_scatter_mode(mode)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    ls.zeroSystem()
    ls.assemble_momentum_edge("viscosity", **MOM_OPTS)
    vals, rhs = ls.values()
    o = oracle_momentum(case, g, omdot, opec, uvw=True)
    ov, orhs = o.get()
    av, arhs = o.get_abs()
    res["momentum_uvw_lhs"] = scaled_err(vals, ov, lscale(ov, av))
    res["momentum_uvw_rhs"] = scaled_err(rhs, orhs, arhs)
    LAST_PLAIN["momentum_uvw_lhs"] = plain_rel(vals, ov)
    LAST_PLAIN["momentum_uvw_rhs"] = plain_rel(rhs, orhs)
    # fused Peclet variant must give the same system
    ls.zeroSystem()
    ls.assemble_momentum_edge("viscosity", fuse_peclet=True, pf=pf, **MOM_OPTS)
    vals, rhs = ls.values()
    res["momentum_fused_lhs"] = scaled_err(vals, ov, lscale(ov, av))
    res["momentum_fused_rhs"] = scaled_err(rhs, orhs, arhs)
    LAST_PLAIN["momentum_fused_lhs"] = plain_rel(vals, ov)
    LAST_PLAIN["momentum_fused_rhs"] = plain_rel(rhs, orhs)
    ls.close()
    mesh.close()
    return res


# ---------------------------------------------------------------------------
# 2-D quad O-grid (airfoilRANSEdge-style: the reference's airfoil mesh is a 2-D
# QUAD4 mesh, ndim = 2; all kernels loop d < ndim)
# ---------------------------------------------------------------------------

class OGrid2D:
    """curvilinear O-grid of ntheta x nr quads around an ellipse, wall-normal
    stretching 1.15, edges shuffled within buckets (irregular ordering)"""

    def __init__(self, ntheta=48, nr=20, stretch=1.15, bucket=64, seed=20261017):
        pkg()
        synth = __import__("nalu_wind_b200.synth", fromlist=["state"])
        rng = np.random.default_rng(seed)
        i, j = np.meshgrid(np.arange(ntheta), np.arange(nr), indexing="xy")
        i, j = i.ravel(), j.ravel()
        th = 2.0 * np.pi * i / ntheta
        dr0 = 0.01
        r = 0.5 + dr0 * (stretch ** j - 1.0) / (stretch - 1.0)
        self.coords = np.stack([1.3 * r * np.cos(th) + 0.02 * np.sin(3 * th),
                                0.8 * r * np.sin(th)], axis=1)
        n = ntheta * nr
        self.n_nodes = n
        self.gid = np.arange(1, n + 1, dtype=np.int64)
        self.hid = np.arange(n, dtype=np.int64)
        node = lambda a, b: (a % ntheta) + ntheta * b
        ring = np.stack([node(i, j), node(i + 1, j)], axis=1)
        m = j < nr - 1
        rad = np.stack([node(i[m], j[m]), node(i[m], j[m] + 1)], axis=1)
        edges = np.concatenate([ring, rad])
        swap = edges[:, 0] > edges[:, 1]  # L = lower global id
        edges[swap] = edges[swap][:, ::-1]
        # shuffle within buckets
        order = np.arange(len(edges))
        for a in range(0, len(edges), bucket):
            rng.shuffle(order[a:a + bucket])
        edges = edges[order]
        self.edges = np.ascontiguousarray(edges.astype(np.int32))
        self.n_edges = len(edges)
        # quad4 connectivity (counter-clockwise in the (theta, r) chart flipped
        # to a positive Jacobian in (x, y))
        ei, ej = np.meshgrid(np.arange(ntheta), np.arange(nr - 1), indexing="xy")
        ei, ej = ei.ravel(), ej.ravel()
        self.elems = np.stack([node(ei, ej), node(ei, ej + 1),
                               node(ei + 1, ej + 1), node(ei + 1, ej)],
                              axis=1).astype(np.int32)
        dx = self.coords[edges[:, 1]] - self.coords[edges[:, 0]]
        ln = np.linalg.norm(dx, axis=1, keepdims=True)
        rr = np.linalg.norm(0.5 * (self.coords[edges[:, 1]] + self.coords[edges[:, 0]]),
                            axis=1, keepdims=True)
        width = np.where(ln > 0.5 * 2 * np.pi * rr / ntheta, dr0 * 4, 2 * np.pi * rr / ntheta)
        t = np.stack([-dx[:, 1], dx[:, 0]], axis=1) / ln
        self.area = np.ascontiguousarray(dx / ln * width + 0.1 * width * t *
                                         rng.standard_normal((len(edges), 1)))
        self.vol = (r * (2 * np.pi / ntheta) * dr0 * stretch ** j).astype(np.float64)
        c3 = np.concatenate([self.coords, np.zeros((n, 1))], axis=1)
        f3 = synth.state(c3 + 2.0, self.gid, (4.0, 4.0, 1.0), DT, GAMMA1)
        f = {}
        for k, v in f3.items():
            if v.ndim == 2 and v.shape[1] == 3:
                f[k] = np.ascontiguousarray(v[:, :2])
            elif v.ndim == 2 and v.shape[1] == 9:
                f[k] = np.ascontiguousarray(v[:, [0, 1, 3, 4]])
            else:
                f[k] = v
        f["dual_nodal_volume"] = self.vol
        self.fields = f


def run_quad2d_case(P, ctx, tile_nodes=64, mode=None, grid=None, **kw):
    """the edge sweep on a 2-D quad mesh (default: the O-grid), product vs
    oracle (ndim = 2)"""
    c = grid if grid is not None else OGrid2D(**kw)
    f = c.fields
    mesh = P.Mesh(ctx, 2, c.edges, c.hid, c.coords, tile_nodes=tile_nodes)
    for name, arr in f.items():
        mesh.put(name, P.NW_NODE, arr)
    mesh.put("edge_area_vector", P.NW_EDGE, c.area)
    mesh.register("mass_flow_rate", P.NW_EDGE, 1)
    mesh.register("peclet_factor", P.NW_EDGE, 1)
    res = {}
    pf, opf = P.peclet_fn("classic", 1.0), orc.peclet("classic", 1.0)
    mesh.mdot_edge(1.0, 1.0)
    mdot = mesh.download("mass_flow_rate")
    omdot = orc.mdot_edge(2, c.edges, c.coords, f["velocity"], f["dpdx"],
                          f["density"], f["pressure"], f["momentum_diag"],
                          c.area, 1.0, 1.0)
    res["mdot"] = scaled_err(mdot, omdot, np.abs(omdot) + 1e-3 * np.max(np.abs(omdot)))
    mesh.peclet_edge("viscosity", pf)
    pec = mesh.download("peclet_factor")
    opec = orc.peclet_edge(2, c.edges, c.coords, f["velocity"], f["density"],
                           f["viscosity"], opf)[1]
    res["peclet"] = scaled_err(pec, opec, np.ones_like(opec))
    mesh.upload("mass_flow_rate", omdot)
    mesh.upload("peclet_factor", opec)
    for phi, grad, d1 in (("pressure", "dpdx_out", 1), ("velocity", "dudx_out", 2)):
        mesh.register(grad, P.NW_NODE, d1 * 2)
        mesh.nodal_grad_edge(phi, grad)
        got = mesh.download(grad)
        ref = orc.nodal_grad_edge(d1, 2, c.edges, f[phi], c.area, c.vol, c.n_nodes)
        mag = np.abs(orc.nodal_grad_edge(d1, 2, c.edges, np.abs(f[phi]),
                                         np.abs(c.area), c.vol, c.n_nodes))
        mag = mag + np.max(np.abs(ref)) * 1e-3
        res["grad_" + phi] = scaled_err(got.reshape(ref.shape), ref, mag)
    g = orc.Graph(1, 0, c.n_nodes - 1)
    g.add_edges(c.edges, c.hid)
    g.finalize()

    def system(kind, nd):
        ls = P.LinearSystem(mesh, kind, nd)
        if mode is not None:
            ls.set_scatter_mode(mode)
        ls.buildEdgeToNodeGraph()
        ls.finalizeLinearSystem()
        ls.zeroSystem()
        return ls

    def compare(name, ls, sink):
        vals, rhs = ls.values()
        ov, orhs = sink.get()
        av, arhs = sink.get_abs()
        if grid is not None:  # a real mesh: see lhs_scale
            av = lhs_scale(np.asarray(g.rows), ov, av)
        res[name + "_lhs"] = scaled_err(vals, ov, av)
        res[name + "_rhs"] = scaled_err(rhs, orhs, arhs)

    ls = system(P.NW_LINSYS_HYPRE, 1)
    ls.assemble_continuity_edge(**CONT_OPTS)
    s = orc.HypreSink(g, c.hid)
    orc.continuity_edge(2, c.edges, c.coords, f["velocity"], f["dpdx"],
                        f["density"], f["pressure"], f["momentum_diag"], c.area,
                        s, **CONT_OPTS)
    compare("continuity", ls, s)
    ls.close()
    ls = system(P.NW_LINSYS_HYPRE, 1)
    ls.assemble_scalar_edge("turbulent_ke", "dkdx", "effective_viscosity_tke",
                            pf=P.peclet_fn("tanh", 2.0, 1.0), **SCAL_OPTS)
    s = orc.HypreSink(g, c.hid)
    orc.scalar_edge(2, c.edges, c.coords, f["velocity"], f["turbulent_ke"],
                    f["dkdx"], f["density"], f["effective_viscosity_tke"], c.area,
                    omdot, s, pf=orc.peclet("tanh", 2.0, 1.0), **SCAL_OPTS)
    compare("scalar", ls, s)
    ls.close()
    ls = system(P.NW_LINSYS_HYPRE_UVW, 2)
    ls.assemble_momentum_edge("viscosity", **MOM_OPTS)
    s = orc.HypreSink(g, c.hid, uvw_ndim=2)
    orc.momentum_edge(2, c.edges, c.coords, f["velocity"], f["dudx"],
                      f["viscosity"], f["density"],
                      f["abl_wall_no_slip_wall_func_node_mask"], c.area, omdot,
                      opec, s, **MOM_OPTS)
    compare("momentum_uvw", ls, s)
    ls.zeroSystem()
    ls.assemble_momentum_edge("viscosity", fuse_peclet=True, pf=pf, **MOM_OPTS)
    compare("momentum_fused", ls, s)
    ls.close()
    mesh.close()
    return res


# ---------------------------------------------------------------------------
# the reference's own regression meshes (tests/golden/mesh_*.npz, made by
# tests/golden/extract_reference_meshes.py)
# ---------------------------------------------------------------------------

def load_reference_mesh(name):
    return np.load(os.path.join(HERE, "golden", "mesh_%s.npz" % name))


# element faces, outward (stk / Exodus node order), for boundary detection and an
# independent element volume
ELEM_FACES = {
    "tet": [(0, 1, 3), (1, 2, 3), (0, 3, 2), (0, 2, 1)],
    "pyr": [(0, 1, 4), (1, 2, 4), (2, 3, 4), (0, 4, 3), (0, 3, 2, 1)],
    "wed": [(0, 1, 4, 3), (1, 2, 5, 4), (0, 3, 5, 2), (0, 2, 1), (3, 4, 5)],
    "hex": [(0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (0, 4, 7, 3),
            (0, 3, 2, 1), (4, 5, 6, 7)],
}
TOPOLOGIES = ("hex", "tet", "wed", "pyr")


def mesh_blocks(m):
    """{topology: [n][npe] connectivity} of a mesh fixture"""
    return {t: np.ascontiguousarray(m["elems_" + t]) for t in TOPOLOGIES
            if "elems_" + t in m.files}


def boundary_nodes(blocks, n_nodes):
    """mask of the nodes on faces that belong to one element only"""
    keys, nodes = [], []
    for t, conn in blocks.items():
        for f in ELEM_FACES[t]:
            fn = conn[:, list(f)].astype(np.int64)
            if fn.shape[1] == 3:
                fn = np.concatenate([fn, -np.ones((len(fn), 1), np.int64)], axis=1)
            nodes.append(fn)
            k = np.sort(fn, axis=1)
            keys.append(((k[:, 0] * (n_nodes + 1) + k[:, 1] + 1) * (n_nodes + 1)
                         + k[:, 2] + 1) * (n_nodes + 1) + k[:, 3] + 1)
    keys, nodes = np.concatenate(keys), np.concatenate(nodes)
    _, inv, cnt = np.unique(keys, return_inverse=True, return_counts=True)
    once = nodes[cnt[inv] == 1].ravel()
    mask = np.zeros(n_nodes, dtype=bool)
    mask[once[once >= 0]] = True
    return mask


def element_volumes_by_faces(topo, conn, coords):
    """divergence-theorem volume of every element, quadrilateral faces fanned
    about their centroid (independent of the sub-control-volume construction)"""
    vol = np.zeros(len(conn))
    x = coords[conn]  # [n][npe][3]
    x = x - x[:, :1]  # element-local origin: no cancellation against |x|^3
    for f in ELEM_FACES[topo]:
        p = x[:, list(f)]
        if len(f) == 3:
            tris = [(p[:, 0], p[:, 1], p[:, 2])]
        else:
            c = p.mean(axis=1)
            tris = [(p[:, k], p[:, (k + 1) % 4], c) for k in range(4)]
        for a, b, c in tris:
            vol += np.einsum("ij,ij->i", a, np.cross(b, c)) / 6.0
    return vol


def oracle_mesh_geometry(blocks, coords, edges):
    """GeometryInteriorAlg over every block (oracle): dual volumes, edge area
    vectors, {topology: element volumes}"""
    n = len(coords)
    acc = (np.zeros(n), np.zeros((len(edges), 3)))
    ev = {}
    for t, conn in blocks.items():
        _, ev[t], _ = orc.geometry_interior_3d(t, conn, coords, edges, n,
                                               accumulate=acc)
    return acc[0], acc[1], ev


class _Obj:
    pass


class RealMeshCase:
    """a 3-D reference mesh as one rank, with the attributes of Case: synthetic
    smooth + noise state, synthetic (edge-aligned + seeded transverse) area
    vectors and positive dual volumes -- all the edge kernels see of geometry --
    or the mesh's true CVFEM dual geometry"""

    def __init__(self, name, seed=20261017, geometry="synthetic"):
        """geometry="cvfem": the true dual mesh (GeometryInteriorAlg over the
        tet / pyramid / wedge / hex blocks, from the oracle) instead"""
        P = pkg()
        synth = __import__("nalu_wind_b200.synth", fromlist=["state"])
        m = load_reference_mesh(name)
        rng = np.random.default_rng(seed)
        coords = np.ascontiguousarray(m["coords"])
        edges = np.ascontiguousarray(m["edges"])
        n = len(coords)
        b = _Obj()
        b.n_nodes, b.n_edges = n, len(edges)
        b.coords, b.gid = coords, m["gid"]
        b.hid = np.arange(n, dtype=np.int64)
        b.own_hid = b.hid
        b.edges = edges
        b.offsets = np.array([0, n], dtype=np.int64)
        b.rank, b.nranks = 0, 1
        b.periodic = (False, False)
        dx = coords[edges[:, 1]] - coords[edges[:, 0]]
        ln = np.linalg.norm(dx, axis=1, keepdims=True)
        b.area = np.ascontiguousarray(
            0.3 * ln * dx + 0.05 * ln * ln * rng.standard_normal(dx.shape))
        b.vol = (0.5 + rng.random(n)) * float(np.mean(ln)) ** 3
        if geometry == "cvfem":
            self.blocks = mesh_blocks(m)
            b.vol, b.area, _ = oracle_mesh_geometry(self.blocks, coords, edges)
        b.make_mesh = lambda ctx, tile_nodes=0: P.Mesh(
            ctx, 3, b.edges, b.hid, b.coords, tile_nodes=tile_nodes)
        self.box = b
        lo, hi = coords.min(0), coords.max(0)
        self.fields = synth.state(coords - lo, b.gid, tuple(hi - lo), DT, GAMMA1)
        self.fields["dual_nodal_volume"] = b.vol
        self.edges, self.area = edges, b.area
        self.n_nodes, self.n_edges = n, len(edges)

    oracle_graph = Case.oracle_graph
    oracle_mdot = Case.oracle_mdot
    oracle_pecfac = Case.oracle_pecfac


class DecomposedRealMesh:
    """The reference's own 8-way decomposition of hybrid.g
    (reg_tests/mesh/hybrid.g.8.0 .. .7, fixture mesh_hybrid_g_8_parts.npz):
    rank=None is the serial mesh (union of the parts, nodes in ascending global
    id, every edge once); rank=r is part r as STK would hold it -- all nodes of
    the part's elements, the edges the rank OWNS (lowest holding rank), node
    ownership = lowest sharing rank, hypre row ids by Realm::set_hypre_global_id
    (src/Realm.C:3587-3679: a rank's owned nodes, sorted by global id, numbered
    from the rank's offset; shared copies carry the owner's id).  Geometry and
    state are pure functions of the global ids, so every rank sees what the
    serial mesh sees.  Attributes as Case."""

    def __init__(self, rank=None, name="hybrid_g_8_parts", seed=20261017):
        P = pkg()
        synth = __import__("nalu_wind_b200.synth", fromlist=["state"])
        m = load_reference_mesh(name)
        np_ = int(m["nparts"])
        self.nparts = np_
        # serial union: nodes by ascending global id
        gids = np.unique(np.concatenate([m["gid_%d" % r] for r in range(np_)]))
        xyz = np.zeros((len(gids), 3))
        owner = np.zeros(len(gids), dtype=np.int32)
        for r in range(np_ - 1, -1, -1):  # the lowest rank writes last (same values)
            at = np.searchsorted(gids, m["gid_%d" % r])
            xyz[at] = m["coords_%d" % r]
            owner[at] = m["node_owner_%d" % r]
        # hypre ids: per owner, owned nodes in ascending global id
        counts = np.bincount(owner, minlength=np_)
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        hyp = np.zeros(len(gids), dtype=np.int64)
        for r in range(np_):
            mine = np.nonzero(owner == r)[0]  # ascending gid already
            hyp[mine] = offsets[r] + np.arange(len(mine))
        # serial edge list: the owned edges of every rank, rank by rank
        eg = []
        for r in range(np_):
            e = m["edges_%d" % r][m["edge_owner_%d" % r] == r]
            g = m["gid_%d" % r]
            eg.append(np.stack([g[e[:, 0]], g[e[:, 1]]], axis=1))
        eg_all = np.concatenate(eg)
        key_all = eg_all[:, 0] * (gids[-1] + 1) + eg_all[:, 1]
        assert len(np.unique(key_all)) == len(key_all)
        ser_edges = np.stack([np.searchsorted(gids, eg_all[:, 0]),
                              np.searchsorted(gids, eg_all[:, 1])], axis=1)
        rng = np.random.default_rng(seed)
        dx = xyz[ser_edges[:, 1]] - xyz[ser_edges[:, 0]]
        ln = np.linalg.norm(dx, axis=1, keepdims=True)
        ser_area = np.ascontiguousarray(
            0.3 * ln * dx + 0.05 * ln * ln * rng.standard_normal(dx.shape))
        ser_vol = (0.5 + rng.random(len(gids))) * float(np.mean(ln)) ** 3
        lo, hi = xyz.min(0), xyz.max(0)
        b = _Obj()
        b.periodic = (False, False)
        if rank is None:
            sel_n = np.arange(len(gids))
            b.edges = np.ascontiguousarray(ser_edges.astype(np.int32))
            b.area = ser_area
            b.rank, b.nranks = 0, 1
            b.hid = np.arange(len(gids), dtype=np.int64)  # serial numbering
            b.offsets = np.array([0, len(gids)], dtype=np.int64)
        else:
            g = m["gid_%d" % rank]
            sel_n = np.searchsorted(gids, g)
            own = m["edge_owner_%d" % rank] == rank
            e = m["edges_%d" % rank][own]
            b.edges = np.ascontiguousarray(e.astype(np.int32))
            k = g[e[:, 0]] * (gids[-1] + 1) + g[e[:, 1]]
            order = np.argsort(key_all)
            at = order[np.searchsorted(key_all[order], k)]
            assert np.array_equal(key_all[at], k)
            b.area = np.ascontiguousarray(ser_area[at])
            b.rank, b.nranks = rank, np_
            b.hid = hyp[sel_n]
            b.offsets = offsets
        b.n_nodes, b.n_edges = len(sel_n), len(b.edges)
        b.coords = np.ascontiguousarray(xyz[sel_n])
        b.gid = gids[sel_n]
        b.own_hid = b.hid
        b.owner = owner[sel_n]
        b.vol = ser_vol[sel_n]
        b.hypre_of_serial = hyp  # serial node index -> decomposed hypre id
        b.make_mesh = lambda ctx, tile_nodes=0: P.Mesh(
            ctx, 3, b.edges, b.hid, b.coords, hypre_offsets=b.offsets,
            rank=b.rank, nranks=b.nranks, tile_nodes=tile_nodes)
        self.box = b
        self.fields = synth.state(b.coords - lo, b.gid, tuple(hi - lo), DT, GAMMA1)
        self.fields["dual_nodal_volume"] = b.vol
        self.edges, self.area = b.edges, b.area
        self.n_nodes, self.n_edges = b.n_nodes, b.n_edges

    oracle_graph = Case.oracle_graph
    oracle_mdot = Case.oracle_mdot
    oracle_pecfac = Case.oracle_pecfac


class RealMesh2D:
    """the airfoilRANSEdge mesh (2-D QUAD4) with OGrid2D's attributes; geometry
    (edge area vectors, dual volumes) is the true CVFEM dual mesh from the
    oracle's GeometryInteriorAlg<Quad4_2D>"""

    def __init__(self, name="airfoilRANSEdge"):
        pkg()
        synth = __import__("nalu_wind_b200.synth", fromlist=["state"])
        m = load_reference_mesh(name)
        self.coords = np.ascontiguousarray(m["coords"])
        self.edges = np.ascontiguousarray(m["edges"])
        self.elems = np.ascontiguousarray(m["elems_qua"])
        self.gid = m["gid"]
        n = len(self.coords)
        self.n_nodes, self.n_edges = n, len(self.edges)
        self.hid = np.arange(n, dtype=np.int64)
        self.vol, _, self.area = orc.geometry_interior_quad4(
            self.elems, self.coords, self.edges, n)
        lo, hi = self.coords.min(0), self.coords.max(0)
        c3 = np.concatenate([self.coords - lo, np.zeros((n, 1))], axis=1)
        f3 = synth.state(c3, self.gid, (hi[0] - lo[0], hi[1] - lo[1], 1.0), DT, GAMMA1)
        f = {}
        for k, v in f3.items():
            if v.ndim == 2 and v.shape[1] == 3:
                f[k] = np.ascontiguousarray(v[:, :2])
            elif v.ndim == 2 and v.shape[1] == 9:
                f[k] = np.ascontiguousarray(v[:, [0, 1, 3, 4]])
            else:
                f[k] = v
        f["dual_nodal_volume"] = self.vol
        self.fields = f
