#!/usr/bin/env python
"""bench.py -- edge-assembly throughput of the B200-native path.

One "step" = one full low-Mach edge sweep over the generated hex mesh, in the
order LowMachEquationSystem::solve_and_update calls the kernels
(src/LowMachEquationSystem.C:770-867):
    MomentumEdgePecletAlg -> momentum assembly (segregated UVW system: zeroSystem,
    MomentumEdgeSolverAlg, loadComplete) -> continuity assembly (zeroSystem,
    ContinuityEdgeSolverAlg, loadComplete) -> MdotEdgeAlg -> nodal gradient of
    pressure -> nodal gradient of velocity
(+ with --sst: TKE and SDR scalar assemblies and their gradients).
metric = mesh edges swept per second (every edge passes through all kernels of
the sweep once per step), whole job, in Medges/s.

  python bench.py --gpus 1 --steps 20 --warmup 5
  torchrun ... bench.py --gpus N ...       (one rank per GPU, z-slab partition)
  python bench.py --impl reference ...     (CPU oracle on the host cores)
"""
import argparse
import json
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
ALG_BYTES = {"peclet": 69.3333, "momentum_uvw": 122.6667,
             "momentum_uvw_fused": 114.6667, "continuity": 85.3333,
             "mdot": 72.0, "grad_scalar": 45.3333, "grad_vector": 66.6667,
             "scalar": 93.3333}
STATE_FIELDS = [("velocity", 3), ("pressure", 1), ("density", 1),
                ("viscosity", 1), ("momentum_diag", 1)]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("NW_BENCH_N", "128")),
                    help="elements per box side per GPU (BASELINE config: 128)")
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
    ap.add_argument("--detail", action="store_true",
                    help="also print per-kernel timings (stderr)")
    return ap.parse_args()


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


def build_case(P, n, nranks, rank):
    """per-rank part of an n x n x (n*nranks) box, z-slab decomposition"""
    synth = __import__("nalu_wind_b200.synth", fromlist=["state"])
    nz = n * nranks
    box = P.BoxMesh(n, n, nz, nranks=nranks, rank=rank)
    fields = synth.state(box.coords, box.gid, (float(n), float(n), float(nz)),
                         DT, GAMMA1)
    fields["dual_nodal_volume"] = box.vol
    return box, fields


def cpu_sweep(orc, box, fields, g, sst):
    """the same sweep on the CPU oracle; returns seconds"""
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
    if sst:
        for q, dq, mu in (("turbulent_ke", "dkdx", "effective_viscosity_tke"),
                          ("specific_dissipation_rate", "dwdx",
                           "effective_viscosity_sdr")):
            s3 = orc.HypreSink(g, box.hid)
            s3.track_abs(False)
            orc.scalar_edge(3, box.edges, box.coords, f["velocity"], f[q], f[dq],
                            f["density"], f[mu], box.area, mdot0, s3,
                            pf=orc.peclet("tanh", 2.0, 1.0), **SCAL_OPTS)
            orc.nodal_grad_edge(1, 3, box.edges, f[q], box.area,
                                f["dual_nodal_volume"], box.n_nodes)
    return time.perf_counter() - t0


def oracle_mod():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    return orc


def run_cpu_baseline(P, n, sst, threads):
    """CPU oracle timed on the host cores on a bounded sample of the workload:
    a box of n_s^3 elements with the same fields (same edges-per-node ratio)."""
    orc = oracle_mod()
    ns = min(n, 96)
    box, fields = build_case(P, ns, 1, 0)
    g = orc.Graph(1, 0, box.n_nodes - 1)
    g.add_edges(box.edges, box.hid)
    g.finalize()
    orc.set_num_threads(threads)
    cpu_sweep(orc, box, fields, g, sst)  # warm-up (page faults, caches)
    reps, tot = 0, 0.0
    while tot < 8.0 and reps < 20:
        tot += cpu_sweep(orc, box, fields, g, sst)
        reps += 1
    orc.set_num_threads(1)
    t1 = cpu_sweep(orc, box, fields, g, sst)  # also single-thread
    return {"value": box.n_edges * reps / tot / 1e6, "unit": "Medges/s",
            "cores": threads, "kind": "port",
            "single_thread_value": box.n_edges / t1 / 1e6,
            "sample": "%d^3-element box (%d edges), %d full sweeps, %.1f s; "
                      "oracle/edge_oracle.cpp (reference cannot be compiled: no "
                      "Kokkos/STK/hypre)" % (ns, box.n_edges, reps, tot)}


def main_reference(args):
    """--impl reference: the reference's CPU implementation of the path (here
    the oracle port: the reference itself needs Trilinos/Kokkos/hypre, absent)
    on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P = graft.load_package()
    threads = os.cpu_count() or 1
    orc = oracle_mod()
    ns = min(args.n, 96)
    box, fields = build_case(P, ns, 1, 0)
    g = orc.Graph(1, 0, box.n_nodes - 1)
    g.add_edges(box.edges, box.hid)
    g.finalize()
    orc.set_num_threads(threads)
    for _ in range(max(args.warmup, 1)):
        cpu_sweep(orc, box, fields, g, args.sst)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_sweep(orc, box, fields, g, args.sst)
    val = box.n_edges * args.steps / t / 1e6
    orc.set_num_threads(1)
    t1 = cpu_sweep(orc, box, fields, g, args.sst)  # also single-thread
    sample = ("each step = one full sweep over a %d^3-element box (%d edges), "
              "same fields/options as the GPU arm" % (ns, box.n_edges))
    line = {
        "impl": "reference", "metric": "edge_assembly_throughput",
        "value": val, "unit": "Medges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": val, "unit": "Medges/s", "cores": threads,
                         "kind": "port", "single_thread_value": box.n_edges / t1 / 1e6,
                         "sample": sample},
        "e2e": {"value": val, "unit": "Medges/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n_gpus):
    n = args.n
    return {
        "workload": "generated %dx%dx%d hex box (BASELINE configs[1]: %d^3 per "
                    "GPU), low-Mach edge sweep: Peclet + momentum(UVW) + "
                    "continuity + mdot + grad(p) + grad(u)%s; fp64; z-slab "
                    "partition over %d GPU(s)" % (
                        n, n, n * n_gpus, n,
                        " + k/omega scalar assemblies + gradients" if args.sst else "",
                        n_gpus),
        "scatter": args.mode,
        "peclet": "fused into momentum" if args.fuse_peclet else "separate kernel",
        "l2": "inputs larger than L2 (sweep working set ~%.1f GB per GPU vs "
              "126 MB L2), no explicit flush" % (
                  ((n + 1) ** 3 * 8 * 60) / 1e9),
        "tile_nodes": args.tile if args.tile else "default",
    }


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

    box, fields = build_case(P, args.n, world, rank)
    mesh = box.make_mesh(ctx, tile_nodes=args.tile)
    for name, arr in fields.items():
        mesh.put(name, P.NW_NODE, arr)
    mesh.put("edge_area_vector", P.NW_EDGE, box.area)
    mesh.register("mass_flow_rate", P.NW_EDGE, 1)
    mesh.register("peclet_factor", P.NW_EDGE, 1)
    mesh.register("dpdx_new", P.NW_NODE, 3)
    mesh.register("dudx_new", P.NW_NODE, 9)
    if args.sst:
        mesh.register("dkdx_new", P.NW_NODE, 3)
        mesh.register("dwdx_new", P.NW_NODE, 3)
    mode = P.NW_SCATTER_SEGMENTED if args.mode == "segmented" else P.NW_SCATTER_ATOMIC
    systems = {}
    for name, kind, nd in (("momentum", P.NW_LINSYS_HYPRE_UVW, 3),
                           ("continuity", P.NW_LINSYS_HYPRE, 1)) + (
            (("tke", P.NW_LINSYS_HYPRE, 1), ("sdr", P.NW_LINSYS_HYPRE, 1))
            if args.sst else ()):
        ls = P.LinearSystem(mesh, kind, nd)
        ls.set_scatter_mode(mode)
        ls.buildEdgeToNodeGraph()
        ls.finalizeLinearSystem()
        systems[name] = ls
    pf = P.peclet_fn("classic", 1.0)
    stream = torch.cuda.ExternalStream(ctx.stream())
    mesh.mdot_edge()  # initial mdot (LowMachEquationSystem.C:710-721)
    ctx.sync()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    mom_events = []
    kernel_events = {}

    def timed(name, fn, record):
        if record:
            a, b = ev(), ev()
            a.record(stream)
            fn()
            b.record(stream)
            kernel_events.setdefault(name, []).append((a, b))
        else:
            fn()

    def sweep(record=False, detail=False):
        rec = record or detail
        if not args.fuse_peclet:
            timed("peclet", lambda: mesh.peclet_edge("viscosity", pf), detail)
        m = systems["momentum"]

        def mom():
            m.zeroSystem()
            if args.fuse_peclet:
                m.assemble_momentum_edge("viscosity", fuse_peclet=True, pf=pf,
                                         **MOM_OPTS)
            else:
                m.assemble_momentum_edge("viscosity", **MOM_OPTS)
        timed("momentum_uvw", mom, rec)
        m.loadComplete()
        c = systems["continuity"]

        def cont():
            c.zeroSystem()
            c.assemble_continuity_edge(**CONT_OPTS)
        timed("continuity", cont, detail)
        c.loadComplete()
        timed("mdot", lambda: mesh.mdot_edge(), detail)
        timed("grad_scalar", lambda: mesh.nodal_grad_edge("pressure", "dpdx_new"), detail)
        timed("grad_vector", lambda: mesh.nodal_grad_edge("velocity", "dudx_new"), detail)
        if args.sst:
            for nm, q, dq, mu, go in (
                    ("tke", "turbulent_ke", "dkdx", "effective_viscosity_tke", "dkdx_new"),
                    ("sdr", "specific_dissipation_rate", "dwdx",
                     "effective_viscosity_sdr", "dwdx_new")):
                s = systems[nm]

                def sc(s=s, q=q, dq=dq, mu=mu):
                    s.zeroSystem()
                    s.assemble_scalar_edge(q, dq, mu, pf=P.peclet_fn("tanh", 2.0, 1.0),
                                           **SCAL_OPTS)
                timed("scalar", sc, detail)
                s.loadComplete()
                timed("grad_scalar", lambda q=q, go=go: mesh.nodal_grad_edge(q, go), detail)

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.sync()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value") ----------------
    for _ in range(max(args.warmup, 3)):
        sweep()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(args.steps):
        sweep(record=True, detail=args.detail)
    ctx.join_comm()  # the last step's halo adds (communication stream) count
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    edges_local = box.n_edges
    te = torch.tensor([edges_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.SUM)
    edges_total = float(te.item())
    value = edges_total * args.steps / (ms_max * 1e-3) / 1e6

    # dominant kernel: momentum UVW assembly, timed live inside the region
    mom_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events["momentum_uvw"]]))
    peak, peak_src = measured_peak()
    mom_key = "momentum_uvw_fused" if args.fuse_peclet else "momentum_uvw"
    ALG_BYTES_RUN = dict(ALG_BYTES, momentum_uvw=ALG_BYTES[mom_key])
    ach = ALG_BYTES_RUN["momentum_uvw"] * edges_local / (mom_ms * 1e-3) / 1e9
    sweep_bytes = sum(ALG_BYTES_RUN[k] for k in (
        "momentum_uvw", "continuity", "mdot", "grad_scalar", "grad_vector"))
    if not args.fuse_peclet:
        sweep_bytes += ALG_BYTES["peclet"]
    if args.sst:
        sweep_bytes += 2 * (ALG_BYTES["scalar"] + ALG_BYTES["grad_scalar"])
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel
    # from the committed `ncu --set full` capture (same mesh: 128^3 per GPU,
    # default tile); null for any other configuration
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_momentum_uvw.json")
    if os.path.exists(tp) and args.n == 128 and not args.tile:
        try:
            traffic = json.load(open(tp)).get(mom_key, {}).get(
                "dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "ls_tile_kernel<MomentumUvwP<3>> "
                "(momentum UVW edge assembly incl. row init%s)" % (
                    ", Peclet factor fused" if args.fuse_peclet else ""),
                "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_edge": ALG_BYTES_RUN["momentum_uvw"],
                "kernel_ms": mom_ms,
                "sweep_achieved_gbs": sweep_bytes * edges_local /
                (ms_max * 1e-3 / args.steps) / 1e9,
                "sweep_frac": sweep_bytes * edges_local /
                (ms_max * 1e-3 / args.steps) / 1e9 / peak,
                "frac_of_nominal_8TBs": ach / 8000.0}
    if args.detail and rank == 0:
        for k, v in kernel_events.items():
            tms = float(np.mean([a.elapsed_time(b) for a, b in v]))
            calls = len(v) / args.steps
            gbs = ALG_BYTES_RUN[k] * edges_local / (tms * 1e-3) / 1e9
            print("  %-14s %8.3f ms x%.0f  %8.1f GB/s algorithmic  %5.1f%% of peak"
                  % (k, tms, calls, gbs, 100 * gbs / peak), file=sys.stderr)

    # ---------------- end to end through the C ABI with host buffers -------
    pinned = {}
    for name, nc in STATE_FIELDS:
        tns = torch.from_numpy(np.ascontiguousarray(fields[name])).pin_memory()
        pinned[name] = (mesh.field_id(name), tns)
    h2d = sum(tns.numel() * 8 for _, tns in pinned.values())
    d2h = 8 * (3 + 1)

    def stage_all():
        for fid, tns in pinned.values():
            mesh.stage_ptr(fid, tns.data_ptr())

    def e2e_step():
        """every step copies its nodal state host -> device and reads its
        residual norms back.  Pipelined form (default): the state of step i+1
        is staged on the copy stream while step i computes."""
        if args.serial_upload:
            for fid, tns in pinned.values():
                mesh.upload_ptr(fid, tns.data_ptr())
        else:
            for fid, _ in pinned.values():
                mesh.commit(fid)
            stage_all()
        sweep()
        n_m = systems["momentum"].rhs_norm2()
        n_c = systems["continuity"].rhs_norm2()
        return n_m, n_c
    if not args.serial_upload:
        stage_all()
    for _ in range(3):
        e2e_step()
    barrier()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(args.steps):
        norms = e2e_step()
    ctx.join_comm()
    e1.record(stream)
    if not args.serial_upload:
        for fid, _ in pinned.values():  # drain the look-ahead copy
            mesh.commit(fid)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = edges_total * args.steps / (float(t.item()) * 1e-3) / 1e6

    # edge kernels (+ row-init launches only when a system has rows no tile
    # owns: none on this mesh)
    launches_per_step = 5 if args.fuse_peclet else 6
    if args.sst:
        launches_per_step += 2 * 3
    line = {
        "metric": "edge_assembly_throughput", "value": value, "unit": "Medges/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, world),
        "edges_total": edges_total, "edges_per_gpu": edges_local,
        "roofline": roofline,
        "e2e": {"value": e2e_val, "unit": "Medges/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "per step: H2D of the nodal state (velocity, pressure, "
                        "density, viscosity, momentum_diag) from pinned host "
                        "memory (%s), the sweep, D2H of the rhs norms of both "
                        "systems (nw_linsys_rhs_norm2)" % (
                            "nw_field_upload on the compute stream"
                            if args.serial_upload else
                            "nw_field_stage on the copy stream one step ahead + "
                            "nw_field_commit")},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "halo_transport": {"nodal_sum": mesh.halo_transport(),
                           "load_complete": systems["momentum"].halo_transport()},
        "mesh_stats": mesh.stats() if rank == 0 else None,
        "residual_norms": {"momentum": [float(x) for x in norms[0]],
                           "continuity": [float(x) for x in norms[1]]},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = run_cpu_baseline(P, args.n, args.sst, 1)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
