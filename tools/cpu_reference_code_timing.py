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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rc = H.RankCase((n, n, n), 1, 0)
    st, b = rc.st, rc.b
    orc.set_num_threads(1)
    o = T.MOM_POINTS[1]
    print("box %d^3: %d nodes, %d edges; one thread" % (n, st.n_nodes, st.n_edges))
    rows = []
    # momentum (UVW)
    w = st.world()
    H._momentum_options(w, o)
    h = R.HypreRef(w, b.own_hid, uvw=True, num_dof=3, node_identifier=rc.ident,
                   node_owner=b.owner, nalu_id=rc.nalu, offsets=b.offsets)
    t_ref = best(lambda: h.sweep("momentum"))
    g = rc.oracle_graph(1)
    s = orc.HypreSink(g, b.hid, uvw_ndim=3)

    def orc_mom():
        s.reset()
        orc.momentum_edge(3, st.edges, st.coords, st.velocity, st.dudx, st.viscosity,
                          st.density, st.mask, st.area, st.mdot, st.pecfac, s, **o)
    t_orc = best(orc_mom)
    vals, r = h.values()
    ov, orh = s.get()
    assert np.array_equal(vals, ov) and np.array_equal(r.ravel(), np.asarray(orh).ravel())
    rows.append(("momentum (UVW)", t_ref, t_orc))
    h.close()
    # continuity
    c = T.CONT_POINTS[0]
    w = T.ref_cont_world(st, c, False, False)
    h = R.HypreRef(w, b.own_hid, num_dof=1, node_identifier=rc.ident,
                   node_owner=b.owner, nalu_id=rc.nalu, offsets=b.offsets)
    t_ref = best(lambda: h.sweep("continuity"))
    s1 = orc.HypreSink(g, b.hid)

    def orc_cont():
        s1.reset()
        T.orc_cont(st, c, False, False, s1)
    t_orc = best(orc_cont)
    vals, r = h.values()
    ov, orh = s1.get()
    assert np.array_equal(vals, ov) and np.array_equal(r.ravel(), np.asarray(orh).ravel())
    rows.append(("continuity", t_ref, t_orc))
    h.close()
    print("%-16s %28s %28s" % ("assembly", "reference's own code", "oracle port"))
    for name, a, p in rows:
        print("%-16s %10.3f s %9.2f Medges/s %10.3f s %9.2f Medges/s" % (
            name, a, st.n_edges / a / 1e6, p, st.n_edges / p / 1e6))
    print("results identical bit for bit: yes")


if __name__ == "__main__":
    main()
