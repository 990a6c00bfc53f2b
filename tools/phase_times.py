#!/usr/bin/env python
"""Per-phase cycle accounting of the tile kernels (diagnostic).

Runs the bench sweep with the NW_PHASE_TIMING build of the library
(`make -C nalu-wind_b200 prof`) and prints, per kernel, the average cycles
thread 0 of a CTA spends between consecutive marks.

  NW_LIB_PATH=nalu-wind_b200/libnalu_edge_b200_prof.so python tools/phase_times.py [--n 128] [--tile T]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NW_LIB_PATH", os.path.join(
    ROOT, "nalu-wind_b200", "libnalu_edge_b200_prof.so"))
import bench  # noqa: E402
import __graft_entry__ as graft  # noqa: E402

KERNELS = ["continuity", "scalar", "momentum", "mdot", "peclet", "grad_scalar",
           "grad_vector", "-",
           "continuity/stream", "scalar/stream", "momentum/stream", "mdot/stream",
           "peclet/stream", "grad_scalar/stream", "grad_vector/stream", "-"]
STREAM = ["loop", "cp.async wait", "tma wait", "top barrier", "issue next",
          "zero+phase 1", "p1 barrier", "phase 2", "p2 barrier", "phase 3"]
LS = ["hdr+init", "tma issue", "halo gather", "stage wait", "phase 1",
      "p1 barrier", "phase 2+3"]
EDGE = ["hdr+init", "tma issue", "halo gather", "stage wait", "compute"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    P = graft.load_package()
    ctx = P.Context(0)
    box, fields = bench.build_case(P, (a.n, a.n, a.n), 1, 0)
    mesh = box.make_mesh(ctx, tile_nodes=a.tile)
    for name, arr in fields.items():
        mesh.put(name, P.NW_NODE, arr)
    mesh.put("edge_area_vector", P.NW_EDGE, box.area)
    mesh.register("mass_flow_rate", P.NW_EDGE, 1)
    mesh.register("peclet_factor", P.NW_EDGE, 1)
    mesh.register("dpdx_new", P.NW_NODE, 3)
    mesh.register("dudx_new", P.NW_NODE, 9)
    mom = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE_UVW, 3)
    con = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    for ls in (mom, con):
        ls.buildEdgeToNodeGraph()
        ls.finalizeLinearSystem()
    pf = P.peclet_fn("classic", 1.0)

    def sweep():
        mesh.peclet_edge("viscosity", pf)
        mom.zeroSystem()
        mom.assemble_momentum_edge("viscosity", fuse_peclet=True, pf=pf,
                                   **bench.MOM_OPTS)  # the bench default
        mom.loadComplete()
        con.zeroSystem()
        con.assemble_continuity_edge(**bench.CONT_OPTS)
        con.loadComplete()
        mesh.mdot_edge()
        mesh.nodal_grad_edge("pressure", "dpdx_new")
        mesh.nodal_grad_edge("velocity", "dudx_new")

    L = P.lib()
    buf = (C.c_ulonglong * (16 * 12))()
    for _ in range(3):
        sweep()
    ctx.sync()
    L.nw_debug_phase_times(buf, 192, 1)
    for _ in range(a.reps):
        sweep()
    ctx.sync()
    L.nw_debug_phase_times(buf, 192, 1)
    print("stats", mesh.stats())
    for k in range(15):
        row = [buf[k * 12 + s] for s in range(12)]
        n = row[11]
        if not n:
            continue
        names = STREAM if k >= 8 else (LS if k < 3 else EDGE)
        tot = sum(row[:len(names)])
        print("%-18s tiles/launch %d, total %.0f cycles per tile (one observer thread)" %
              (KERNELS[k], n // a.reps, tot / n))
        for nm, v in zip(names, row):
            print("    %-12s %9.0f cycles  %5.1f%%" % (nm, v / n, 100.0 * v / tot))


if __name__ == "__main__":
    main()
