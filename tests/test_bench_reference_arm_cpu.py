"""bench.py --impl reference (the CPU arm the driver times beside the GPU
arm): runs without a GPU, prints one JSON line with the keys the bench
contract names, and its rank-per-core sample covers every edge of the sample
box exactly once.  No GPU."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_reference_arm_line():
    env = dict(os.environ, NW_BENCH_N="12")
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
         "--steps", "2", "--warmup", "1"], env=env, capture_output=True,
        text=True, timeout=600, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"] == "edge_assembly_throughput" and line["unit"] == "Medges/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert cb["single_thread_value"] > 0 and cb["openmp_atomic_value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "Medges/s",
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "12x12x%d-element box" % max(12, 6 * cb["cores"]) in cb["sample"]  # the box it ran, not the label


def test_rank_per_core_sample_covers_the_box_once():
    sys.path.insert(0, ROOT)
    import bench
    import __graft_entry__ as graft
    P = graft.load_package()
    orc = bench.oracle_mod()
    cs = bench.CpuSample(P, orc, 10, 3)
    assert cs.nparts == 3
    serial = cs.serial_case(P)[0]
    assert cs.nz == 18 and cs.n_edges == serial.n_edges
    # owned rows of the parts tile the serial row range
    rows = sorted((int(p[0].offsets[r]), int(p[0].offsets[r + 1]))
                  for r, p in enumerate(cs.parts))
    assert rows[0][0] == 0 and rows[-1][1] == serial.n_nodes
    assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    assert cs.sweep_rank_per_core(False) > 0
    # a norm of the partitioned sweep: rhs rows summed over owners + sharers
    # reproduce the serial oracle's rhs (the shared tail is what MPI would add)
    import numpy as np
    f = cs.serial_case(P)[1]
    box, g = serial, cs.serial_case(P)[2]
    s = orc.HypreSink(g, box.hid)
    orc.continuity_edge(3, box.edges, box.coords, f["velocity"], f["dpdx"],
                        f["density"], f["pressure"], f["momentum_diag"],
                        box.area, s, **bench.CONT_OPTS)
    ref = s.get()[1][0][:box.n_nodes]
    acc = np.zeros(box.n_nodes)
    for r, (b, fp, gp) in enumerate(cs.parts):
        sp = orc.HypreSink(gp, b.hid)
        orc.continuity_edge(3, b.edges, b.coords, fp["velocity"], fp["dpdx"],
                            fp["density"], fp["pressure"], fp["momentum_diag"],
                            b.area, sp, **bench.CONT_OPTS)
        rhs = sp.get()[1][0]
        lo, hi = int(b.offsets[r]), int(b.offsets[r + 1])
        acc[lo:hi] += rhs[:hi - lo]
        acc[gp.row_indices_shared] += rhs[hi - lo:hi - lo + gp.num_rows_shared]
    assert np.allclose(acc, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
