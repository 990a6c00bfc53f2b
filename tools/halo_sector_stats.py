#!/usr/bin/env python
"""Design aid (CPU only): 32-byte sectors the tile kernels' halo gather touches
under three nodal layouts -- SoA per component (today), per-field AoS (coords[3],
velocity[3], dudx[9], scalars; the reference's own field layout) and one packed
18-double record per node -- from the real tile plans, per warp-wide group of 32
halo entries.  Usage: python tools/halo_sector_stats.py [n] [tile_nodes]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import parity_util as pu  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    tile = int(sys.argv[2]) if len(sys.argv) > 2 else 192
    L = pu.emu_lib()
    L.emu_plan_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.emu_halo_sectors_aos.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    case = pu.Case(dims=(n, n, n))
    emu = pu.Emu(case, tile_nodes=tile)
    out = np.zeros(8, dtype=np.int64)
    L.emu_plan_stats(emu.h, out.ctypes.data)
    halo, E, N = int(out[1]), case.n_edges, case.n_nodes
    sec = {}
    for k in (1, 3, 9, 18):
        o = np.zeros(2, dtype=np.int64)
        L.emu_halo_sectors_aos(emu.h, k, 32, o.ctypes.data)
        sec[k] = int(o[1])
    print("box %d^3, tile %d: %d nodes, %d edges, %d halo entries (%.2f per node), "
          "%.3f tile-edge records per edge" % (n, tile, N, E, halo, halo / N, out[2] / E))
    rows = [("SoA per component (18 x 8 B requests per halo node)", 18 * sec[1], 18),
            ("per-field AoS (2 x 24 B, 72 B, 3 x 8 B)", 2 * sec[3] + sec[9] + 3 * sec[1], 2 * 2 + 5 + 3),
            ("one 144 B record per node (9 x 16 B)", sec[18], 9)]
    print("momentum halo gather, per edge (useful bytes %.1f):" % (halo * 144.0 / E))
    for name, s, req in rows:
        print("  %-52s %6.1f B in sectors, %5.2f requests" % (name, s * 32.0 / E, req * halo / E))


if __name__ == "__main__":
    main()
