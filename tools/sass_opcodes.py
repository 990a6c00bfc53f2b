#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the in-tree library:
  python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier, LDGSTS = cp.async,
DFMA / DADD / DMUL / DSETP = the FP64 pipe."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nalu-wind_b200", "libnalu_edge_b200.so")
KEEP = ("ls_tile_kernel", "ls_pipe_kernel", "scalar_pair_tile_kernel", "mdot_tile_kernel",
        "peclet_tile_kernel", "grad_tile_kernel", "p2p_", "periodic_update",
        "momentum_mono_atomic", "ls_atomic_kernel")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), text=True,
                         capture_output=True).stdout.splitlines()
    return [re.sub(r"\(anonymous namespace\)::|nw::|\(.*$", "", o) for o in out]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], text=True,
                         capture_output=True).stdout
    funcs = re.split(r"\n\s+Function : ", txt)[1:]
    names = [f.split("\n")[0].strip() for f in funcs]
    nice = demangle(names)
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], text=True,
                         capture_output=True).stdout
    regs = dict(re.findall(r"Function (\S+):\n\s+REG:(\d+)", res))
    print("# SASS opcode histogram, sm_100a, %s" % os.path.relpath(LIB, ROOT))
    print("# kernel | instructions | registers | selected opcodes | FP64 "
          "(DFMA+DADD+DMUL+DSETP) | top opcodes")
    for f, nm, nc in zip(funcs, names, nice):
        if not any(k in nc for k in KEEP) or "stream" in nc:
            continue
        ops = collections.Counter()
        for line in f.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops[m.group(1)] += 1
        n = sum(ops.values())
        sel = {k: ops.get(k, 0) for k in ("UBLKCP", "UTMALDG", "SYNCS", "LDGSTS",
                                          "LDS", "STS", "LDG", "STG", "ATOMG", "RED")}
        fp64 = sum(ops.get(k, 0) for k in ("DFMA", "DADD", "DMUL", "DSETP"))
        top = ", ".join("%s %d" % kv for kv in ops.most_common(8))
        print("%s | %d | %s | %s | %d | %s" % (
            nc, n, regs.get(nm, "?"),
            " ".join("%s=%d" % kv for kv in sel.items() if kv[1]), fp64, top))


if __name__ == "__main__":
    sys.exit(main())
