#!/usr/bin/env python
"""bench.py -- edge-assembly throughput of the B200-native path.

One "step" = one full low-Mach edge sweep over the generated hex mesh, in the
order LowMachEquationSystem::solve_and_update calls the kernels
(src/LowMachEquationSystem.C:770-867):
    MomentumEdgePecletAlg -> momentum assembly (segregated UVW system: zeroSystem,
    MomentumEdgeSolverAlg, loadComplete) -> continuity assembly (zeroSystem,
    ContinuityEdgeSolverAlg, loadComplete) -> MdotEdgeAlg -> nodal gradient of
    pressure -> nodal gradient of velocity
(+ with --sst: TKE and SDR scalar assemblies and their gradients,
src/ShearStressTransportEquationSystem.C:247-320).
metric = mesh edges swept per second (every edge passes through all kernels of
the sweep once per step), whole job, in Medges/s.

  python bench.py --gpus 1 --steps 20 --warmup 5
  torchrun ... bench.py --gpus N ...       (one rank per GPU, z-slab partition)
  python bench.py --impl reference ...     (CPU oracle on the host cores)

At N > 1 the line carries two more things:
  * a parity gate: before anything is timed, a small box of the same shape
    (same partitioning, same kernels, same options) is assembled and its global
    residual norms (nw_linsys_rhs_norm2_global) are compared with the serial
    CPU oracle; on mismatch no line is printed and the exit code is non-zero;
  * `north_star`: BASELINE.json configs[2] -- exactly 512^3 elements
    partitioned over the N GPUs (strong scaling), SST sweep -- with ms/sweep,
    per-GPU Gedges/s, the sweep's fraction of the HBM roofline and the share of
    the sweep spent in halo exchanges.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

DT, GAMMA1 = 0.5, 1.5
MOM_OPTS = dict(include_divu=0.0, alpha=0.0, alpha_upw=1.0, ho_upwind=1.0,
                relax_fac=0.7, use_limiter=True)
SCAL_OPTS = dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, relax_fac=0.9,
                 use_limiter=True)
CONT_OPTS = dict(dt=DT, gamma1=GAMMA1, noc_fac=1.0, interp_together=1.0,
                 solve_incompressible=0.0)

# ALGORITHMIC bytes per edge on a hex mesh (r = 1/3, z = 7)
# momentum_uvw_fused: the Peclet factor is computed in the kernel, so its 8 B /
# edge read (and the whole K9 pass) drop out of the algorithmic traffic
# grad_scalar_pair: two scalar gradients in one launch; charged as two (the
# shared area vectors / dual volumes are read once, the figure is conservative
# for the roofline fraction only in the sense that it is the un-fused one)
ALG_BYTES = {"peclet": 69.3333, "momentum_uvw": 122.6667,
             "momentum_uvw_fused": 114.6667, "continuity": 85.3333,
             "mdot": 72.0, "grad_scalar": 45.3333, "grad_vector": 66.6667,
             "scalar": 93.3333, "grad_scalar_pair": 2 * 45.3333,
             "scalar_pair": 2 * 93.3333}
# what a solver changes between two sweeps of one nonlinear iteration and has
# to hand over again (the solves update velocity and pressure, the momentum
# system's diagonal gives momentum_diag); density / viscosity and the
# coordinates stay on the device
E2E_FIELDS = [("velocity", 3), ("pressure", 1), ("momentum_diag", 1)]
SST_SCALARS = (("tke", "turbulent_ke", "dkdx", "effective_viscosity_tke", "dkdx_new"),
               ("sdr", "specific_dissipation_rate", "dwdx",
                "effective_viscosity_sdr", "dwdx_new"))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("NW_BENCH_N", "128")),
                    help="elements per box side per GPU (BASELINE config: 128)")
    ap.add_argument("--mesh", default="box", choices=["box", "warped", "mixed"],
                    help="box: BASELINE configs[1] (default); warped: configs[3]-"
                         "style curvilinear (warped + stretched) hex box whose "
                         "edges arrive in bucket-shuffled order; mixed: "
                         "configs[4], the reference's multiElemTypeCylinder.g "
                         "(tet / hex / wedge / pyramid) tiled to >= 10^7 edges "
                         "(one GPU)")
    ap.add_argument("--tile", type=int,
                    default=int(os.environ.get("NW_TILE_NODES", "0")))
    ap.add_argument("--sst", action="store_true",
                    help="add the k and omega scalar assemblies + gradients")
    ap.add_argument("--mode", default="segmented", choices=["segmented", "atomic"])
    ap.add_argument("--fuse-peclet", dest="fuse_peclet", action="store_true",
                    default=True,
                    help="fold MomentumEdgePecletAlg into the momentum kernel "
                         "(SURVEY 8f-1; default: r01d measured +6%% sweep throughput)")
    ap.add_argument("--no-fuse-peclet", dest="fuse_peclet", action="store_false",
                    help="launch MomentumEdgePecletAlg as its own kernel")
    ap.add_argument("--serial-upload", action="store_true",
                    help="e2e leg: nw_field_upload on the compute stream instead "
                         "of the pipelined nw_field_stage / nw_field_commit")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="eager", choices=["eager", "plain"],
                    help="N > 1: 'eager' (default) lets every edge assembly "
                         "store its shared rows straight into the owners' windows "
                         "(nw_linsys_set_eager_exchange; the sweep has no other "
                         "contribution before loadComplete); 'plain' sends them "
                         "in loadComplete")
    ap.add_argument("--fuse-scalars", dest="no_fuse_scalars", action="store_false",
                    default=True,
                    help="--sst: assemble the TKE and SDR systems through "
                         "nw_assemble_scalar_edge_pair (one launch with "
                         "NW_SCALAR_PAIR_FUSED=1; measured slower, see DESIGN.md)")
    ap.add_argument("--no-fuse-scalars", dest="no_fuse_scalars", action="store_true",
                    help="(default) two scalar assembly launches")
    ap.add_argument("--north-star", default=os.environ.get("NW_BENCH_NORTH_STAR", "auto"),
                    choices=["auto", "on", "off"],
                    help="the 512^3 SST strong-scaling record (auto: when N > 1)")
    ap.add_argument("--ns-n", type=int, default=int(os.environ.get("NW_BENCH_NS_N", "512")))
    ap.add_argument("--sustain-s", type=float, default=1.0,
                    help="length of the sustained-throughput leg in seconds")
    ap.add_argument("--detail", action="store_true",
                    help="also print per-kernel timings (stderr)")
    return ap.parse_args()


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_numa_node(local):
    """pin this rank (and the pinned host buffers it allocates afterwards) to the
    NUMA node its GPU hangs off: r01 SCALE showed all ranks on node 0 and the
    per-GPU H2D rate halving at N = 8"""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "cpus": len(ids)}
    except Exception as e:  # no sysfs / no permission: leave the affinity alone
        return {"numa_node": None, "error": str(e)[:80]}


class TiledMixedMesh:
    """BASELINE configs[4]: the reference's multiElemTypeCylinder.g (TETRA4 /
    HEX8 / WEDGE6 / PYRAMID5; fixture tests/golden/mesh_multiElemTypeCylinder.npz)
    repeated on a lattice of translated copies until it has >= min_edges edges.
    Connectivity, valence distribution (tets: ~14 neighbours per row) and edge
    order are the mesh's own; the dual geometry the edge kernels read is
    synthetic (edge-aligned area vectors + seeded transverse part, positive
    volumes), as in the parity tests.  One rank."""

    def __init__(self, P, min_edges=10_000_000, seed=20261017):
        m = np.load(os.path.join(ROOT, "tests", "golden",
                                 "mesh_multiElemTypeCylinder.npz"))
        c0, e0 = m["coords"], m["edges"].astype(np.int64)
        n0, ne0 = len(c0), len(e0)
        k = int(math.ceil(min_edges / ne0))
        side = int(math.ceil(k ** (1.0 / 3.0)))
        ext = (c0.max(0) - c0.min(0)) * 1.05
        self.copies = k
        shifts = np.array([[i, j, l] for l in range(side) for j in range(side)
                           for i in range(side)][:k], dtype=np.float64) * ext
        self.coords = np.ascontiguousarray(
            (c0[None, :, :] + shifts[:, None, :]).reshape(-1, 3))
        self.edges = np.ascontiguousarray(
            (e0[None, :, :] + (np.arange(k) * n0)[:, None, None]).reshape(-1, 2)
            .astype(np.int32))
        n = self.n_nodes = k * n0
        self.n_edges = k * ne0
        self.gid = np.arange(1, n + 1, dtype=np.int64)
        self.hid = np.arange(n, dtype=np.int64)
        self.own_hid = self.hid
        self.offsets = np.array([0, n], dtype=np.int64)
        rng = np.random.default_rng(seed)
        dx = self.coords[self.edges[:, 1]] - self.coords[self.edges[:, 0]]
        ln = np.linalg.norm(dx, axis=1, keepdims=True)
        self.area = np.ascontiguousarray(
            0.3 * ln * dx + 0.05 * ln * ln * rng.standard_normal(dx.shape))
        self.vol = (0.5 + rng.random(n)) * float(np.mean(ln)) ** 3
        self.lengths = tuple((self.coords.max(0) - self.coords.min(0)).tolist())
        self.coords0 = self.coords.min(0)
        self._P = P

    def make_mesh(self, ctx, tile_nodes=0):
        return self._P.Mesh(ctx, 3, self.edges, self.hid, self.coords,
                            tile_nodes=tile_nodes)


def build_case(P, dims, nranks, rank, kind="box"):
    """per-rank part of an nx x ny x nz box, z-slab decomposition (kind "box" /
    "warped"), or the tiled mixed-element mesh (kind "mixed", one rank)"""
    synth = __import__("nalu_wind_b200.synth", fromlist=["state"])
    if kind == "mixed":
        assert nranks == 1, "--mesh mixed runs on one GPU"
        box = TiledMixedMesh(P)
        fields = synth.state_chunked(
            box.coords - box.coords0, box.gid, box.lengths, DT, GAMMA1,
            threads=int(os.environ.get("NW_HOST_THREADS", "0")) or None)
        fields["dual_nodal_volume"] = box.vol
        return box, fields
    nx, ny, nz = dims
    if kind == "warped":
        box = P.BoxMesh(nx, ny, nz, nranks=nranks, rank=rank, warp=0.15,
                        zstretch=1.1, shuffle_bucket=512)
    else:
        box = P.BoxMesh(nx, ny, nz, nranks=nranks, rank=rank)
    fields = synth.state_chunked(box.coords, box.gid,
                                 (float(nx), float(ny), float(nz)), DT, GAMMA1,
                                 threads=int(os.environ.get("NW_HOST_THREADS", "0")) or None)
    fields["dual_nodal_volume"] = box.vol
    return box, fields


def cpu_sweep(orc, box, fields, g, sst, norms=None):
    """the same sweep on the CPU oracle; returns seconds.  norms: dict that
    receives the sums of squares of the rhs of every system"""
    f = fields
    pf = orc.peclet("classic", 1.0)
    t0 = time.perf_counter()
    pec = orc.peclet_edge(3, box.edges, box.coords, f["velocity"], f["density"],
                          f["viscosity"], pf)[1]
    mdot0 = orc.mdot_edge(3, box.edges, box.coords, f["velocity"], f["dpdx"],
                          f["density"], f["pressure"], f["momentum_diag"],
                          box.area, 1.0, 1.0)
    s = orc.HypreSink(g, box.hid, uvw_ndim=3)
    s.track_abs(False)  # test-only bookkeeping the reference does not have
    orc.momentum_edge(3, box.edges, box.coords, f["velocity"], f["dudx"],
                      f["viscosity"], f["density"],
                      f["abl_wall_no_slip_wall_func_node_mask"], box.area, mdot0,
                      pec, s, **MOM_OPTS)
    s2 = orc.HypreSink(g, box.hid)
    s2.track_abs(False)
    orc.continuity_edge(3, box.edges, box.coords, f["velocity"], f["dpdx"],
                        f["density"], f["pressure"], f["momentum_diag"],
                        box.area, s2, **CONT_OPTS)
    orc.nodal_grad_edge(1, 3, box.edges, f["pressure"], box.area,
                        f["dual_nodal_volume"], box.n_nodes)
    orc.nodal_grad_edge(3, 3, box.edges, f["velocity"], box.area,
                        f["dual_nodal_volume"], box.n_nodes)
    if norms is not None:
        norms["momentum"] = np.sum(s.get()[1] ** 2, axis=1)
        norms["continuity"] = np.sum(s2.get()[1] ** 2, axis=1)
    if sst:
        for nm, q, dq, mu, _ in SST_SCALARS:
            s3 = orc.HypreSink(g, box.hid)
            s3.track_abs(False)
            orc.scalar_edge(3, box.edges, box.coords, f["velocity"], f[q], f[dq],
                            f["density"], f[mu], box.area, mdot0, s3,
                            pf=orc.peclet("tanh", 2.0, 1.0), **SCAL_OPTS)
            orc.nodal_grad_edge(1, 3, box.edges, f[q], box.area,
                                f["dual_nodal_volume"], box.n_nodes)
            if norms is not None:
                norms[nm] = np.sum(s3.get()[1] ** 2, axis=1)
    return time.perf_counter() - t0


def oracle_mod():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    return orc


def host_threads():
    """threads this process may use (the affinity mask, not the machine)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


class CpuSample:
    """Bounded sample of the workload for the CPU legs: a box of n_s^3 elements
    with the same fields (same edges-per-node ratio; throughput is per edge),
    timed the two ways the reference uses host cores:

      * rank per core -- how nalu-wind is run on CPUs: one MPI rank per core,
        Kokkos Serial inside (plain adds, no atomics).  The box is cut into one
        z-slab per thread with STK ownership semantics (the same partitioner the
        GPU arm uses at N > 1); every thread sweeps its own part with the
        serial oracle into its own owned + shared rows.  The MPI exchanges
        (loadComplete, the shared-node gradient sum) are NOT timed, which
        favours the CPU;
      * OpenMP threads over one rank's edges with an atomic on every add (the
        reference's Kokkos OpenMP build).
    """

    def __init__(self, P, orc, n, threads):
        from concurrent.futures import ThreadPoolExecutor
        self.orc = orc
        self.ns = ns = min(n, 96)
        self.threads = threads
        self.nparts = max(1, min(threads, 64))
        # slabs at least 6 element layers thick, so that a part is mostly
        # interior like a production rank (throughput is per edge)
        self.nz = nz = max(ns, 6 * self.nparts)
        self.parts = []
        for r in range(self.nparts):
            box, fields = build_case(P, (ns, ns, nz), self.nparts, r)
            g = orc.Graph(1, int(box.offsets[r]), int(box.offsets[r + 1]) - 1)
            g.add_edges(box.edges, box.hid)
            g.finalize()
            self.parts.append((box, fields, g))
        self.n_edges = sum(p[0].n_edges for p in self.parts)
        self.pool = ThreadPoolExecutor(self.nparts)
        self._serial = None

    def sweep_rank_per_core(self, sst):
        """one sweep of every part, all parts at once; wall seconds (= the
        slowest rank, as a solver step would see it)"""
        self.orc.set_num_threads(1)
        t0 = time.perf_counter()
        list(self.pool.map(
            lambda p: cpu_sweep(self.orc, p[0], p[1], p[2], sst), self.parts))
        return time.perf_counter() - t0

    def serial_case(self, P):
        if self._serial is None:
            box, fields = build_case(P, (self.ns, self.ns, self.nz), 1, 0)
            g = self.orc.Graph(1, 0, box.n_nodes - 1)
            g.add_edges(box.edges, box.hid)
            g.finalize()
            self._serial = (box, fields, g)
        return self._serial

    def sweep_one_rank(self, P, sst, threads):
        """the whole sample as one rank on `threads` OpenMP threads (atomics
        when threads > 1); seconds"""
        box, fields, g = self.serial_case(P)
        self.orc.set_num_threads(threads)
        t = cpu_sweep(self.orc, box, fields, g, sst)
        self.orc.set_num_threads(1)
        return t

    def describe(self):
        return ("%dx%dx%d-element box (%d edges) cut into %d z-slabs, one serial "
                "oracle sweep per slab and host thread at once (the reference's "
                "rank-per-core CPU deployment, Kokkos Serial; MPI exchanges not "
                "timed)" % (self.ns, self.ns, self.nz, self.n_edges, self.nparts))

    def extras(self, P, sst):
        """the two other figures SURVEY 8(d) asks for: one thread, and all
        threads under OpenMP with atomic adds"""
        box = self.serial_case(P)[0]
        self.sweep_one_rank(P, sst, self.threads)  # warm-up
        ta = min(self.sweep_one_rank(P, sst, self.threads) for _ in range(2))
        t1 = min(self.sweep_one_rank(P, sst, 1) for _ in range(2))
        return {"single_thread_value": box.n_edges / t1 / 1e6,
                "openmp_atomic_value": box.n_edges / ta / 1e6,
                "openmp_atomic_threads": self.threads,
                "reference_code": reference_code_leg()}


def reference_code_leg(n=40):
    """Calibration of the port, beside it: one momentum (UVW) and one continuity
    assembly through the REFERENCE'S OWN edge algorithm + HypreLinearSystem code
    (oracle/_ref: compiled unmodified against stand-in Kokkos / STK / hypre
    headers) and through the oracle port on the same small
    box, one thread each, results compared bit for bit.  The stand-in Realm
    reads flat arrays where nalu-wind reads STK buckets, so this is not the
    baseline (`value` stays the port, the faster of the two); never fatal."""
    try:
        so = os.path.join(ROOT, "oracle", "_ref", "libnalu_ref.so")
        if not os.path.exists(so):
            return {"unavailable": "oracle/_ref/libnalu_ref.so not built"}
        for d in ("tools", "tests", "oracle"):
            p = os.path.join(ROOT, d)
            if p not in sys.path:
                sys.path.insert(0, p)
        import cpu_reference_code_timing as crt
        out = crt.measure(n, reps=2)
        out["what"] = ("Medges/s of one assembly, one thread: the reference's own "
                       "MomentumEdgeSolverAlg / ContinuityEdgeSolverAlg + "
                       "HypreLinearSystem code run over stand-in infrastructure "
                       "(oracle/_ref) vs the oracle port")
        return out
    except Exception as e:  # calibration only
        return {"unavailable": str(e)[:200]}


def run_cpu_baseline(P, n, sst):
    """CPU oracle timed on the host cores on a bounded sample of the workload"""
    orc = oracle_mod()
    cs = CpuSample(P, orc, n, host_threads())
    cs.sweep_rank_per_core(sst)  # warm-up (page faults, caches)
    reps, tot = 0, 0.0
    while tot < 8.0 and reps < 40:
        tot += cs.sweep_rank_per_core(sst)
        reps += 1
    out = {"value": cs.n_edges * reps / tot / 1e6, "unit": "Medges/s",
           "cores": cs.nparts, "kind": "port",
           "sample": "%s: %d sweeps in %.1f s; oracle/edge_oracle.cpp (the "
                     "reference cannot be compiled here: no Kokkos/STK/hypre)" % (
                         cs.describe(), reps, tot)}
    out.update(cs.extras(P, sst))
    return out


def main_reference(args):
    """--impl reference: the reference's CPU implementation of the path (here
    the oracle port: the reference itself needs Trilinos/Kokkos/hypre, absent)
    on all host cores, one rank per core; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P = graft.load_package()
    orc = oracle_mod()
    cs = CpuSample(P, orc, args.n, host_threads())
    for _ in range(max(args.warmup, 1)):
        cs.sweep_rank_per_core(args.sst)
    t = 0.0
    for _ in range(args.steps):
        t += cs.sweep_rank_per_core(args.sst)
    val = cs.n_edges * args.steps / t / 1e6
    sample = ("each step = one full sweep over a " + cs.describe() +
              " -- a bounded sample of the workload, same fields / options as "
              "the GPU arm; the metric is per edge")
    cfg = workload_config(args, args.gpus)
    cfg["reference_sample"] = sample
    cpu = {"value": val, "unit": "Medges/s", "cores": cs.nparts, "kind": "port",
           "sample": sample}
    cpu.update(cs.extras(P, args.sst))
    line = {
        "impl": "reference", "metric": "edge_assembly_throughput",
        "value": val, "unit": "Medges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": cpu,
        "e2e": {"value": val, "unit": "Medges/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n_gpus):
    n = args.n
    mesh = {"box": "generated %dx%dx%d hex box (BASELINE configs[1]: %d^3 per "
                   "GPU)" % (n, n, n * n_gpus, n),
            "warped": "generated %dx%dx%d curvilinear hex box (warp 0.15, "
                      "z-stretch 1.1, edges delivered in bucket-shuffled order; "
                      "BASELINE configs[3]-style)" % (n, n, n * n_gpus),
            "mixed": "reg_tests/mesh/multiElemTypeCylinder.g (tet / hex / wedge "
                     "/ pyramid) tiled to >= 10^7 edges, synthetic dual geometry "
                     "(BASELINE configs[4])"}[getattr(args, "mesh", "box")]
    return {
        "workload": "%s, low-Mach edge sweep: Peclet + momentum(UVW) + "
                    "continuity + mdot + grad(p) + grad(u)%s; fp64; z-slab "
                    "partition over %d GPU(s)" % (
                        mesh,
                        " + k/omega scalar assemblies + gradients" if args.sst else "",
                        n_gpus),
        "scatter": args.mode,
        "peclet": "fused into momentum" if args.fuse_peclet else "separate kernel",
        "l2": "inputs larger than L2 (sweep working set ~%.1f GB per GPU vs "
              "126 MB L2), no explicit flush" % (
                  ((n + 1) ** 3 * 8 * 60) / 1e9),
        "tile_nodes": args.tile if args.tile else "default",
        "exchange": ("n/a (one GPU)" if n_gpus == 1 else
                     "fused: the tile kernels store shared rows / nodal partial "
                     "sums straight into the neighbours' windows over NVLink, "
                     "one pull kernel per exchange on a communication stream"
                     if getattr(args, "exchange", "eager") == "eager" and
                     os.environ.get("NW_HALO_OVERLAP", "1") != "0" else
                     "after the full kernel (loadComplete / post_work)"),
    }


class Sweep:
    """one rank's mesh, fields, linear systems and the sweep over them"""

    def __init__(self, P, ctx, args, dims, world, rank, sst, torch, kind="box"):
        self.P, self.ctx, self.args, self.sst, self.torch = P, ctx, args, sst, torch
        self.world = world
        t0 = time.time()
        self.box, self.fields = build_case(P, dims, world, rank, kind)
        t1 = time.time()
        box = self.box
        self.mesh = mesh = box.make_mesh(ctx, tile_nodes=args.tile)
        t2 = time.time()
        for name, arr in self.fields.items():
            if not sst and name in ("turbulent_ke", "dkdx", "specific_dissipation_rate",
                                    "dwdx", "effective_viscosity_tke",
                                    "effective_viscosity_sdr"):
                continue
            mesh.put(name, P.NW_NODE, arr)
        # host copies are needed only for the fields the e2e leg re-uploads
        keep = {k for k, _ in E2E_FIELDS}
        for name in list(self.fields):
            if name not in keep:
                del self.fields[name]
        mesh.put("edge_area_vector", P.NW_EDGE, box.area)
        mesh.register("mass_flow_rate", P.NW_EDGE, 1)
        mesh.register("peclet_factor", P.NW_EDGE, 1)
        mesh.register("dpdx_new", P.NW_NODE, 3)
        mesh.register("dudx_new", P.NW_NODE, 9)
        if sst:
            mesh.register("dkdx_new", P.NW_NODE, 3)
            mesh.register("dwdx_new", P.NW_NODE, 3)
        mode = (P.NW_SCATTER_SEGMENTED if args.mode == "segmented"
                else P.NW_SCATTER_ATOMIC)
        self.systems = {}
        for name, kind, nd in (("momentum", P.NW_LINSYS_HYPRE_UVW, 3),
                               ("continuity", P.NW_LINSYS_HYPRE, 1)) + (
                (("tke", P.NW_LINSYS_HYPRE, 1), ("sdr", P.NW_LINSYS_HYPRE, 1))
                if sst else ()):
            ls = P.LinearSystem(mesh, kind, nd)
            ls.set_scatter_mode(mode)
            ls.buildEdgeToNodeGraph()
            ls.finalizeLinearSystem()
            if world > 1:
                ls.set_eager_exchange(args.exchange == "eager")
            self.systems[name] = ls
        t3 = time.time()
        # r = nodes per edge, z = non-zeros per row: the mesh's own figures for
        # the general byte formulas of BASELINE.md section 4
        sz = self.systems["continuity"].sizes
        self.r = box.n_nodes / max(1, box.n_edges)
        self.z = sz.num_nonzeros_owned / max(1, sz.num_rows_owned)
        self.setup_s = {"mesh_generation_and_state": t1 - t0, "mesh_plan": t2 - t1,
                        "fields_and_linear_systems": t3 - t2}
        self.pf = P.peclet_fn("classic", 1.0)
        self.pf_scalar = P.peclet_fn("tanh", 2.0, 1.0)
        self.stream = torch.cuda.ExternalStream(ctx.stream())
        self.kernel_events = {}
        mesh.mdot_edge()  # initial mdot (LowMachEquationSystem.C:710-721)
        ctx.sync()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def _timed(self, name, fn, record):
        if record:
            a, b = self.ev(), self.ev()
            a.record(self.stream)
            fn()
            b.record(self.stream)
            self.kernel_events.setdefault(name, []).append((a, b))
        else:
            fn()

    def run(self, record=False, detail=False):
        args, mesh, systems = self.args, self.mesh, self.systems
        timed = self._timed
        rec = record or detail
        if not args.fuse_peclet:
            timed("peclet", lambda: mesh.peclet_edge("viscosity", self.pf), detail)
        m = systems["momentum"]

        def mom():
            m.zeroSystem()
            if args.fuse_peclet:
                m.assemble_momentum_edge("viscosity", fuse_peclet=True, pf=self.pf,
                                         **MOM_OPTS)
            else:
                m.assemble_momentum_edge("viscosity", **MOM_OPTS)
        timed("momentum_uvw", mom, rec)
        timed("load_complete", m.loadComplete, detail)
        c = systems["continuity"]

        def cont():
            c.zeroSystem()
            c.assemble_continuity_edge(**CONT_OPTS)
        timed("continuity", cont, detail)
        timed("load_complete", c.loadComplete, detail)
        timed("mdot", lambda: mesh.mdot_edge(), detail)
        timed("grad_scalar", lambda: mesh.nodal_grad_edge("pressure", "dpdx_new"), detail)
        timed("grad_vector", lambda: mesh.nodal_grad_edge("velocity", "dudx_new"), detail)
        if self.sst and not args.no_fuse_scalars:
            # TKE + SDR assembled in one launch (same graph, same state)
            (na, qa_, dqa, mua, _), (nb_, qb_, dqb, mub, _) = SST_SCALARS
            sa, sb = systems[na], systems[nb_]
            so = dict(SCAL_OPTS, pf=self.pf_scalar)

            def pair():
                sa.zeroSystem()
                sb.zeroSystem()
                sa.assemble_scalar_edge_pair(qa_, dqa, mua, sb, qb_, dqb, mub, opts=so)
            timed("scalar_pair", pair, detail)
            timed("load_complete", sa.loadComplete, detail)
            timed("load_complete", sb.loadComplete, detail)
        elif self.sst:
            for nm, q, dq, mu, go in SST_SCALARS:
                s = systems[nm]

                def sc(s=s, q=q, dq=dq, mu=mu):
                    s.zeroSystem()
                    s.assemble_scalar_edge(q, dq, mu, pf=self.pf_scalar, **SCAL_OPTS)
                timed("scalar", sc, detail)
                timed("load_complete", s.loadComplete, detail)
            # dkdx and dwdx: one launch (both NodalGradEdgeAlg instances of the
            # SST system see the same, unchanged inputs)
            (_, qa, _, _, ga), (_, qb, _, _, gb) = SST_SCALARS
            timed("grad_scalar_pair",
                  lambda: mesh.nodal_grad_edge_pair(qa, ga, qb, gb), detail)

    def alg_bytes(self):
        """ALGORITHMIC bytes per edge of every kernel (BASELINE.md section 4,
        general formulas with this mesh's r and z; a hex box gives ALG_BYTES)"""
        r, z = self.r, self.z
        b = {"mdot": 96 * r + 40, "grad_scalar": 40 * r + 32,
             "grad_vector": 104 * r + 32, "peclet": 64 * r + 48,
             "continuity": (96 + 8 * (z + 1)) * r + 32,
             "scalar": (96 + 8 * (z + 1)) * r + 40,
             "momentum_uvw": (144 + 8 * (z + 3)) * r + 48}
        b["momentum_uvw_fused"] = b["momentum_uvw"] - 8.0
        b["grad_scalar_pair"] = 2 * b["grad_scalar"]
        b["scalar_pair"] = 2 * b["scalar"]
        return b

    def sweep_bytes(self):
        """algorithmic bytes per edge of one sweep as it runs here"""
        a = self.args
        B = self.alg_bytes()
        b = (B["momentum_uvw_fused" if a.fuse_peclet else "momentum_uvw"] +
             B["continuity"] + B["mdot"] + B["grad_scalar"] + B["grad_vector"])
        if not a.fuse_peclet:
            b += B["peclet"]
        if self.sst:
            b += 2 * (B["scalar"] + B["grad_scalar"])
        return b

    def launches_per_step(self):
        n = 5 if self.args.fuse_peclet else 6
        if self.sst:
            # scalar assemblies (one fused launch or two) + the paired gradient
            n += (2 if self.args.no_fuse_scalars else 1) + 1
        if self.world > 1:
            # halo exchange kernels (peer-memory transport).  Fused: the
            # producing kernel stores into the neighbours' windows, one pull
            # kernel per exchange (the pair gradient is one exchange).  Plain:
            # a push and a pull kernel per exchange, one exchange per field.
            fused = (self.args.exchange == "eager" and
                     os.environ.get("NW_HALO_OVERLAP", "1") != "0")
            n_ls = 2 + (2 if self.sst else 0)
            n_grad_calls = 2 + (1 if self.sst else 0)
            n_sums = 2 + (2 if self.sst else 0)
            n += (n_ls + n_grad_calls) if fused else 2 * (n_ls + n_sums)
        return n

    def norms(self, glob):
        out = {}
        for nm, ls in self.systems.items():
            out[nm] = [float(x) for x in (ls.rhs_norm2_global() if glob
                                          else ls.rhs_norm2())]
        return out

    def close(self):
        for ls in self.systems.values():
            ls.close()
        self.mesh.close()
        self.systems, self.mesh, self.box, self.fields = {}, None, None, None


def main():
    args = parse()
    if args.impl == "reference":
        return main_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU "
                         "fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    # host-side set-up (mesh generation, plan builder): share the cores among
    # the ranks of this node (torchrun exports OMP_NUM_THREADS=1)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    os.environ.setdefault("NW_HOST_THREADS", str(max(
        1, len(os.sched_getaffinity(0)) // max(1, local_world))))
    P = graft.load_package()
    ctx = P.Context(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(P.Context.comm_unique_id()),
                               dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().tolist()), world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.sync()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def time_steps(sw, steps, record=False, detail=False):
        """K sweeps bracketed by barrier + synchronize, CUDA events on the
        launching stream, max over ranks; returns ms for the K sweeps"""
        barrier()
        e0, e1 = sw.ev(), sw.ev()
        e0.record(sw.stream)
        for _ in range(steps):
            sw.run(record=record, detail=detail)
        ctx.join_comm()  # the last step's halo adds (communication stream) count
        e1.record(sw.stream)
        barrier()
        return allmax(e0.elapsed_time(e1))

    # ------------- parity gate of a multi-rank run (before any timing) ------
    gate = None
    if world > 1:
        try:
            gate = parity_gate(P, ctx, args, world, rank, torch, dist)
        except P.NwError:
            raise  # the product refused: that is a failure, not a gate problem
        except Exception as e:  # a bug in the checker must not pose as a mismatch
            gate = {"ok": None, "checker_error": str(e)[:300]}
        if gate["ok"] is False:
            if rank == 0:
                log("PARITY GATE FAILED, no bench line:", json.dumps(gate))
            dist.destroy_process_group()
            raise SystemExit(3)

    n = args.n
    if args.mesh == "mixed" and world > 1:
        raise SystemExit("bench.py: --mesh mixed runs on one GPU")
    sw = Sweep(P, ctx, args, (n, n, n * world), world, rank, args.sst, torch,
               kind=args.mesh)
    box, mesh, systems = sw.box, sw.mesh, sw.systems

    # ---------------- device-resident throughput ("value") ----------------
    # Warm-up in two parts around the start of the clock sampler (nvidia-smi
    # needs ~0.3 s to spawn): the last three warm-up sweeps run immediately
    # ahead of the timed region, so that it does not begin on a GPU that has
    # just idled for 0.3 s.  W (>= 3) sweeps in total.
    warm = max(args.warmup, 3)
    for _ in range(warm - 3):
        sw.run()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    for _ in range(3):
        sw.run()
    ms_max = time_steps(sw, args.steps, record=True, detail=args.detail)
    clocks = sampler.stop() if rank == 0 else None
    edges_local = box.n_edges
    edges_total = allsum(edges_local)
    value = edges_total * args.steps / (ms_max * 1e-3) / 1e6

    # sustained: the same sweep back to back for >= sustain_s seconds
    per_step = ms_max / args.steps
    ksus = max(args.steps, int(math.ceil(args.sustain_s * 1e3 / per_step)))
    ms_sus = time_steps(sw, ksus)
    sustained = {"value": edges_total * ksus / (ms_sus * 1e-3) / 1e6,
                 "unit": "Medges/s", "steps": ksus, "seconds": ms_sus * 1e-3}

    # N > 1: the same sweep with the halo exchanges switched off (the results
    # are then wrong and are not used): what the exchanges cost in this line
    exchange = None
    if world > 1:
        ctx.debug_skip_exchange(True)
        for _ in range(2):
            sw.run()
        ms_skip = time_steps(sw, args.steps)
        ctx.debug_skip_exchange(False)
        sw.run()  # leave consistent state behind
        exchange = {"ms_per_step_without_exchanges": ms_skip / args.steps,
                    "share_of_step": 1.0 - ms_skip / ms_max}

    # dominant kernel: momentum UVW assembly, timed live inside the region
    kev = sw.kernel_events
    mom_ms = float(np.mean([a.elapsed_time(b) for a, b in kev["momentum_uvw"]]))
    peak, peak_src = measured_peak()
    mom_key = "momentum_uvw_fused" if args.fuse_peclet else "momentum_uvw"
    AB = sw.alg_bytes()
    ALG_BYTES_RUN = dict(AB, momentum_uvw=AB[mom_key])
    ach = ALG_BYTES_RUN["momentum_uvw"] * edges_local / (mom_ms * 1e-3) / 1e9
    sweep_bytes = sw.sweep_bytes()
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel
    # from the committed `ncu --set full` capture (same mesh: 128^3 per GPU,
    # default tile); null for any other configuration
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_momentum_uvw.json")
    if os.path.exists(tp) and args.n == 128 and not args.tile and args.mesh == "box":
        try:
            traffic = json.load(open(tp)).get(mom_key, {}).get(
                "dram_bytes_per_launch")
        except Exception:
            traffic = None
    sweep_gbs = sweep_bytes * edges_local / (ms_max * 1e-3 / args.steps) / 1e9
    roofline = {"bound": "hbm", "kernel": "ls_tile_kernel<MomentumUvwP<3>> "
                "(momentum UVW edge assembly incl. row init%s)" % (
                    ", Peclet factor fused" if args.fuse_peclet else ""),
                "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_edge": ALG_BYTES_RUN["momentum_uvw"],
                "kernel_ms": mom_ms,
                "sweep_bytes_per_edge": sweep_bytes,
                "sweep_achieved_gbs": sweep_gbs,
                "sweep_frac": sweep_gbs / peak,
                "frac_of_nominal_8TBs": ach / 8000.0}
    per_kernel = {}
    for k, v in kev.items():
        if k == "load_complete":
            if args.detail and rank == 0 and v:
                print("  %-14s %8.3f ms x%.0f  (shared-row exchange)" % (
                    k, float(np.mean([a.elapsed_time(b) for a, b in v])),
                    len(v) / args.steps), file=sys.stderr)
            continue
        tms = float(np.mean([a.elapsed_time(b) for a, b in v]))
        gbs = ALG_BYTES_RUN[k] * edges_local / (tms * 1e-3) / 1e9
        per_kernel[k] = {"ms": tms, "calls_per_step": len(v) / args.steps,
                         "algorithmic_gbs": gbs, "frac": gbs / peak}
        if args.detail and rank == 0:
            print("  %-14s %8.3f ms x%.0f  %8.1f GB/s algorithmic  %5.1f%% of peak"
                  % (k, tms, len(v) / args.steps, gbs, 100 * gbs / peak),
                  file=sys.stderr)
    if args.detail:
        # the monolithic 3-dof momentum system (HypreLinearSystem, 6x6 blocks;
        # not part of the sweep: every deck of SURVEY 8 runs the segregated UVW
        # system), timed on its own: the tile kernel (node graph's plan, three
        # rows per node, no atomics) and the fp64-atomic scatter through the
        # slot map (zero fill + atomics)
        try:
            mono = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 3)
            mono.buildEdgeToNodeGraph()
            mono.finalizeLinearSystem()
            mesh.peclet_edge("viscosity", sw.pf)
            bpe = (144 + 8 * (9 * sw.z + 3)) * sw.r + 48
            for key, smode, note in (
                    ("momentum_monolithic_tile", P.NW_SCATTER_SEGMENTED,
                     "ls_tile_kernel<MomentumMonoP>: rows written once, no zero fill"),
                    ("momentum_monolithic_atomic", P.NW_SCATTER_ATOMIC,
                     "zero fill (memset of 9 z N values) + atomic scatter")):
                if smode == P.NW_SCATTER_SEGMENTED and not mono.uses_tile_path():
                    continue
                mono.set_scatter_mode(smode)

                def mono_asm():
                    mono.zeroSystem()
                    mono.assemble_momentum_edge("viscosity", **MOM_OPTS)
                for _ in range(2):
                    mono_asm()
                barrier()
                a, b = sw.ev(), sw.ev()
                a.record(sw.stream)
                for _ in range(5):
                    mono_asm()
                b.record(sw.stream)
                barrier()
                tms = a.elapsed_time(b) / 5
                gbs = bpe * edges_local / (tms * 1e-3) / 1e9
                per_kernel[key] = {
                    "ms": tms, "calls_per_step": 0, "algorithmic_gbs": gbs,
                    "frac": gbs / peak, "algorithmic_bytes_per_edge": bpe,
                    "note": note}
                if rank == 0:
                    print("  %-26s %8.3f ms (not in the sweep) %8.1f GB/s algorithmic  "
                          "%5.1f%% of peak" % (key, tms, gbs, 100 * gbs / peak),
                          file=sys.stderr)
            mono.close()
        except P.NwError as e:  # e.g. > 2^31 non-zeros on a large partition
            per_kernel["momentum_monolithic"] = {"error": str(e)[:160]}
        roofline["per_kernel"] = per_kernel

    # ---------------- end to end through the C ABI with host buffers -------
    pinned = {}
    for name, nc in E2E_FIELDS:
        tns = torch.from_numpy(np.ascontiguousarray(sw.fields[name])).pin_memory()
        pinned[name] = (mesh.field_id(name), tns)
    h2d = sum(tns.numel() * 8 for _, tns in pinned.values())
    d2h = 8 * (3 + 1)

    def stage_all():
        for fid, tns in pinned.values():
            mesh.stage_ptr(fid, tns.data_ptr())

    def e2e_step():
        """every step copies the nodal state a solver changes per iteration
        host -> device and reads its residual norms back.  Pipelined form
        (default): the state of step i+1 is staged on the copy stream while
        step i computes."""
        if args.serial_upload:
            for fid, tns in pinned.values():
                mesh.upload_ptr(fid, tns.data_ptr())
        else:
            for fid, _ in pinned.values():
                mesh.commit(fid)
            stage_all()
        sw.run()
        n_m = systems["momentum"].rhs_norm2()
        n_c = systems["continuity"].rhs_norm2()
        return n_m, n_c
    if not args.serial_upload:
        stage_all()
    for _ in range(3):
        e2e_step()
    barrier()
    e0, e1 = sw.ev(), sw.ev()
    e0.record(sw.stream)
    for _ in range(args.steps):
        norms = e2e_step()
    ctx.join_comm()
    e1.record(sw.stream)
    if not args.serial_upload:
        for fid, _ in pinned.values():  # drain the look-ahead copy
            mesh.commit(fid)
    barrier()
    e2e_ms = allmax(e0.elapsed_time(e1))
    e2e_val = edges_total * args.steps / (e2e_ms * 1e-3) / 1e6

    line = {
        "metric": "edge_assembly_throughput", "value": value, "unit": "Medges/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, world),
        "edges_total": edges_total, "edges_per_gpu": edges_local,
        "roofline": roofline,
        "sustained": sustained,
        "e2e": {"value": e2e_val, "unit": "Medges/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "per step and per GPU: H2D of the nodal state a solver "
                        "changes between sweeps (velocity, pressure, "
                        "momentum_diag: 40 B per node) from pinned host memory "
                        "(%s), the sweep, D2H of the rhs norms of both systems "
                        "(nw_linsys_rhs_norm2); density, viscosity, coordinates "
                        "and the gradients stay resident" % (
                            "nw_field_upload on the compute stream"
                            if args.serial_upload else
                            "nw_field_stage on the copy stream one step ahead + "
                            "nw_field_commit"),
                "numa_binding": numa},
        # edge kernels (+ row-init launches only when a system has rows no tile
        # owns: none on this mesh); at N > 1 every halo exchange adds a push and
        # one or two pull kernels of this library
        "gpu_launches": sw.launches_per_step() * args.steps,
        "clocks": clocks,
        "halo_transport": {"nodal_sum": mesh.halo_transport(),
                           "load_complete": systems["momentum"].halo_transport()},
        "mesh_stats": mesh.stats() if rank == 0 else None,
        "setup_seconds": sw.setup_s,
        "residual_norms": {"momentum": [float(x) for x in norms[0]],
                           "continuity": [float(x) for x in norms[1]]},
    }
    if gate is not None:
        line["parity_gate"] = gate
    if exchange is not None:
        line["halo_exchange"] = exchange

    # ---------------- BASELINE configs[2]: 512^3 SST over the N GPUs --------
    want_ns = args.north_star == "on" or (args.north_star == "auto" and world > 1)
    if want_ns:
        sw.close()
        del pinned
        try:
            line["north_star"] = north_star(P, ctx, args, world, rank, torch,
                                            time_steps, allsum, allmax, barrier,
                                            peak)
        except Exception as e:  # never lose the headline line
            line["north_star"] = {"error": str(e)[:300]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = run_cpu_baseline(P, args.n, args.sst)
        except Exception as e:  # the GPU line must not be lost to the CPU leg
            line["cpu_baseline"] = {"error": str(e)[:300]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def parity_gate(P, ctx, args, world, rank, torch, dist):
    """a small box of the same shape and partitioning through
    the same kernels; global rhs norms (sum over the owned rows of all ranks,
    one ncclAllReduce) against the serial CPU oracle (test infrastructure:
    only the checker) before anything is timed."""
    orc = oracle_mod()
    g_n = 20
    dims = (g_n, g_n, 6 * world)
    sw = Sweep(P, ctx, args, dims, world, rank, True, torch)
    sw.run()
    got = sw.norms(glob=True)
    transports = {"nodal_sum": sw.mesh.halo_transport(),
                  "load_complete": sw.systems["momentum"].halo_transport()}
    # nodal gradient after the shared-node sum: compare the owned nodes' sum
    # of squares as well (exercises the second kind of exchange)
    gsum = 0.0
    gp = sw.mesh.download("dpdx_new").reshape(-1, 3)
    lo, hi = int(sw.box.offsets[rank]), int(sw.box.offsets[rank + 1])
    own = (sw.box.own_hid >= lo) & (sw.box.own_hid < hi)
    gsum = float(np.sum(gp[own] ** 2))
    t = torch.tensor([gsum], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    gsum = float(t.item())
    sw.close()
    box, fields = build_case(P, dims, 1, 0)
    g = orc.Graph(1, 0, box.n_nodes - 1)
    g.add_edges(box.edges, box.hid)
    g.finalize()
    ref = {}
    cpu_sweep(orc, box, fields, g, True, norms=ref)
    gref = orc.nodal_grad_edge(1, 3, box.edges, fields["pressure"], box.area,
                               fields["dual_nodal_volume"], box.n_nodes)
    worst = abs(gsum - float(np.sum(gref ** 2))) / float(np.sum(gref ** 2))
    detail = {"grad_pressure": worst}
    for nm, r in ref.items():
        r = np.asarray(r, dtype=np.float64)
        e = float(np.max(np.abs(np.asarray(got[nm]) - r) / r))
        detail[nm] = e
        worst = max(worst, e)
    tol = 1e-10  # a norm of ~1e5 terms each good to 1e-12 relative
    return {"ok": bool(worst < tol), "worst_relative_error": worst,
            "tolerance": tol, "mesh": "%dx%dx%d box over %d ranks, SST sweep" % (
                dims + (world,)), "checked": detail, "transports": transports}


def north_star(P, ctx, args, world, rank, torch, time_steps, allsum, allmax,
               barrier, peak):
    nsn = args.ns_n
    if nsn % world:
        return {"skipped": "%d^3 does not split into %d z-slabs" % (nsn, world)}
    # ~1.1 GB of host memory per million nodes while the plan is built
    try:
        import psutil
        need = 1.1e3 * (nsn + 1) ** 2 * (nsn // world + 1) * world
        avail = psutil.virtual_memory().available
        if need > 0.8 * avail:
            return {"skipped": "host memory: ~%.0f GB needed to build the plans "
                               "of all ranks, %.0f GB available" % (
                                   need / 1e9, avail / 1e9)}
    except ImportError:
        pass
    t0 = time.time()
    sw = Sweep(P, ctx, args, (nsn, nsn, nsn), world, rank, True, torch)
    setup = time.time() - t0
    steps = max(5, min(args.steps, 10))
    for _ in range(3):
        sw.run()
    ms = time_steps(sw, steps, record=True, detail=True)
    edges_local = sw.box.n_edges
    edges_max = allmax(edges_local)
    edges_total = allsum(edges_local)
    ms_sweep = ms / steps
    # the same sweep without its halo exchanges (results incomplete; timing only)
    ctx.debug_skip_exchange(True)
    for _ in range(2):
        sw.run()
    ms_nox = time_steps(sw, steps) / steps
    ctx.debug_skip_exchange(False)
    sw.run()  # a complete sweep again before the norms are read
    sweep_bytes = sw.sweep_bytes()
    gbs = sweep_bytes * edges_max / (ms_sweep * 1e-3) / 1e9
    kern = {}
    for k, v in sw.kernel_events.items():
        kern[k] = {"ms": float(np.mean([a.elapsed_time(b) for a, b in v])),
                   "calls_per_sweep": len(v) / steps}
    out = {
        "config": "BASELINE configs[2]: generated %d^3 hex mesh (%d nodes, %d "
                  "edges), SST k-omega edge sweep (momentum UVW with fused "
                  "Peclet + continuity + mdot + grad p + grad u + k, omega "
                  "assemblies + their gradients), z-slab partition over %d "
                  "B200" % (nsn, (nsn + 1) ** 3, int(edges_total), world),
        "scaling": "strong", "n_gpus": world, "steps": steps,
        "ms_per_sweep": ms_sweep, "goal_ms_per_sweep_at_8_gpus": 9.5,
        "value_medges_per_s": edges_total / (ms_sweep * 1e-3) / 1e6,
        "gedges_per_s_per_gpu": edges_max / (ms_sweep * 1e-3) / 1e9,
        "edges_per_gpu_max": edges_max,
        "sweep_bytes_per_edge": sweep_bytes,
        "sweep_achieved_gbs_per_gpu": gbs,
        "sweep_frac": gbs / peak,
        "sweep_frac_vs_738_7_bytes": 738.6667 * edges_max /
        (ms_sweep * 1e-3) / 1e9 / peak,
        "note": "sweep_frac charges the sweep as it runs (%.1f B/edge: the fused "
                "momentum kernel does not read or write peclet_factor and K9 "
                "is not launched); sweep_frac_vs_738_7_bytes uses SURVEY 8(d)'s "
                "un-fused figure" % sweep_bytes,
        "ms_per_sweep_without_exchanges": ms_nox,
        "exchange_share": max(0.0, 1.0 - ms_nox / ms_sweep),
        "kernels": kern,
        "halo_transport": {"nodal_sum": sw.mesh.halo_transport(),
                           "load_complete": sw.systems["momentum"].halo_transport()},
        "setup_seconds": dict(sw.setup_s, total=setup),
        "residual_norms_global": sw.norms(glob=True),
        "mesh_stats": sw.mesh.stats() if rank == 0 else None,
    }
    sw.close()
    return out


if __name__ == "__main__":
    main()
