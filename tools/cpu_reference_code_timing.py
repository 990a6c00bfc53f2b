#!/usr/bin/env python
"""How fast is the reference's OWN code on one host core, next to the oracle port?

bench.py's CPU legs time the oracle port (cpu_baseline.kind = "port"): the
reference as a program cannot be built here.  Its edge algorithms and
HypreLinearSystem do compile against stand-in headers (oracle/_ref, DESIGN.md
section 4), so this script times ONE assembly of each linear system the way the
reference runs it -- zeroSystem, the edge algorithm's execute() feeding the
reference's CoeffApplier, loadComplete (tests/ref_edge.py: HypreRef.sweep) --
beside the oracle's kernel + sink on the same mesh, one thread each.  The
stand-in Realm reads flat arrays where the reference reads STK buckets, so this
is an upper bound on the reference's per-core speed, not a measurement of
nalu-wind; it calibrates the port.   python tools/cpu_reference_code_timing.py [n]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as orc  # noqa: E402
import ref_edge as R  # noqa: E402
import test_reference_edge_runs as T  # noqa: E402
import test_reference_hypre_runs as H  # noqa: E402


def best(f, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def measure(n=40, reps=3):
    """one momentum (UVW) and one continuity assembly through the reference's
    own code and through the oracle port, one thread each; Medges/s"""
    rc = H.RankCase((n, n, n), 1, 0)
    st, b = rc.st, rc.b
    orc.set_num_threads(1)
    o = T.MOM_POINTS[1]
    out = {"box": "%d^3 elements, %d edges" % (n, st.n_edges), "cores": 1,
           "unit": "Medges/s", "identical_results": True}
    # momentum (UVW)
    w = st.world()
    H._momentum_options(w, o)
    h = R.HypreRef(w, b.own_hid, uvw=True, num_dof=3, node_identifier=rc.ident,
                   node_owner=b.owner, nalu_id=rc.nalu, offsets=b.offsets)
    t_ref = best(lambda: h.sweep("momentum"), reps)
    g = rc.oracle_graph(1)
    s = orc.HypreSink(g, b.hid, uvw_ndim=3)

    def orc_mom():
        s.reset()
        orc.momentum_edge(3, st.edges, st.coords, st.velocity, st.dudx, st.viscosity,
                          st.density, st.mask, st.area, st.mdot, st.pecfac, s, **o)
    t_orc = best(orc_mom, reps)
    vals, r = h.values()
    ov, orh = s.get()
    same = np.array_equal(vals, ov) and np.array_equal(r.ravel(), np.asarray(orh).ravel())
    out["momentum_uvw"] = {"reference_code": st.n_edges / t_ref / 1e6,
                           "port": st.n_edges / t_orc / 1e6}
    h.close()
    # continuity
    c = T.CONT_POINTS[0]
    w = T.ref_cont_world(st, c, False, False)
    h = R.HypreRef(w, b.own_hid, num_dof=1, node_identifier=rc.ident,
                   node_owner=b.owner, nalu_id=rc.nalu, offsets=b.offsets)
    t_ref = best(lambda: h.sweep("continuity"), reps)
    s1 = orc.HypreSink(g, b.hid)

    def orc_cont():
        s1.reset()
        T.orc_cont(st, c, False, False, s1)
    t_orc = best(orc_cont, reps)
    vals, r = h.values()
    ov, orh = s1.get()
    same = same and np.array_equal(vals, ov) and np.array_equal(
        r.ravel(), np.asarray(orh).ravel())
    out["continuity"] = {"reference_code": st.n_edges / t_ref / 1e6,
                         "port": st.n_edges / t_orc / 1e6}
    h.close()
    out["identical_results"] = bool(same)
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    m = measure(n)
    print("box %s; one thread" % m["box"])
    print("%-16s %24s %24s" % ("assembly", "reference's own code", "oracle port"))
    for name in ("momentum_uvw", "continuity"):
        print("%-16s %14.2f Medges/s %14.2f Medges/s" % (
            name, m[name]["reference_code"], m[name]["port"]))
    print("results identical bit for bit:", "yes" if m["identical_results"] else "NO")


if __name__ == "__main__":
    main()
