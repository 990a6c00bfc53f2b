"""The ctypes mirror of every option / descriptor struct of
include/nalu_edge_b200.h has the C compiler's size and field offsets (a field
added on one side only would shift every later option silently).  No GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

PAIRS = [("nw_mesh_desc", "MeshDesc"), ("nw_mesh_stats", "MeshStats"),
         ("nw_peclet_fn", "PecletFn"), ("nw_mdot_opts", "MdotOpts"),
         ("nw_peclet_opts", "PecletOpts"), ("nw_linsys_sizes", "LinsysSizes"),
         ("nw_continuity_opts", "ContinuityOpts"),
         ("nw_mdot_extra_opts", "MdotExtraOpts"),
         ("nw_scalar_opts", "ScalarOpts"), ("nw_momentum_opts", "MomentumOpts"),
         ("nw_mass_bdf_opts", "MassBdfOpts")]


def test_ctypes_structs_match_the_header(tmp_path):
    import __graft_entry__ as graft
    P = graft.load_package()
    lines = ['#include <stdio.h>', '#include <stddef.h>',
             '#include "nalu_edge_b200.h"', 'int main(void) {']
    for cname, pyname in PAIRS:
        st = getattr(P, pyname)
        lines.append('printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in st._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));'
                         % (cname, fname, cname, fname))
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True,
                         check=True).stdout
    got = {}
    for m in re.finditer(r"^(\w+) (\w+) (\d+)$", out, re.M):
        got[(m.group(1), m.group(2))] = int(m.group(3))
    for cname, pyname in PAIRS:
        st = getattr(P, pyname)
        assert got[(cname, "size")] == C.sizeof(st), cname
        for fname, _ in st._fields_:
            assert got[(cname, fname)] == getattr(st, fname).offset, (cname, fname)
    # nw_momentum_opts.has_vof sits in what used to be tail padding: kernels
    # that take the struct by value keep their parameter layout
    assert got[("nw_momentum_opts", "size")] == 104
    assert got[("nw_momentum_opts", "has_vof")] == 100
