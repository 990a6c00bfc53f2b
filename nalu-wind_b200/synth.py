"""Synthetic nodal state for tests and bench: smooth trig fields
shaped like the reference's TrigFieldFunction
(unit_tests/kernels/UnitTestKernelUtils.C:31-330) plus 1 % uniform noise.  The
noise is a counter-based hash of (global node id, field index), so every rank of
a partitioned run sees exactly the same value for a shared node as the serial
run does.  Test / bench input only."""
import numpy as np

SEED = 20261017


def _hash01(gid, k):
    """splitmix64 of (gid, k) -> uniform [0,1)"""
    x = (gid.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) +
         np.uint64((k * 0xBF58476D1CE4E5B9 + SEED) & 0xFFFFFFFFFFFFFFFF))
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def _noise(gid, k, amp=0.01):
    return 1.0 + amp * (2.0 * _hash01(gid, k) - 1.0)


def state(coords, gid, lengths, dt=0.5, gamma1=1.5, periodic_gid=None):
    """dict of nodal fields in the reference's layout.  `periodic_gid` (global
    id of the periodic master for each node) makes slave copies carry their
    master's state, as the periodic manager guarantees in the reference."""
    with np.errstate(over="ignore"):
        g = gid if periodic_gid is None else periodic_gid
        L = np.asarray(lengths, dtype=np.float64)
        x = coords / L  # in [0,1]
        if periodic_gid is not None:
            x = x % 1.0
        tp = 2.0 * np.pi
        sx, cx = np.sin(tp * x[:, 0]), np.cos(tp * x[:, 0])
        sy, cy = np.sin(tp * x[:, 1]), np.cos(tp * x[:, 1])
        sz, cz = np.sin(tp * x[:, 2]), np.cos(tp * x[:, 2])
        n = len(coords)
        f = {}
        u = np.zeros((n, 3))
        u[:, 0] = (7.25 + 1.5 * cx * sy * cz) * _noise(g, 1)
        u[:, 1] = (3.38 + 1.5 * sx * cy * cz) * _noise(g, 2)
        u[:, 2] = (0.5 * sx * sy * sz) * _noise(g, 3)
        f["velocity"] = u
        f["pressure"] = (-0.25 * (np.cos(2 * tp * x[:, 0]) +
                                  np.cos(2 * tp * x[:, 1])) * _noise(g, 4))
        f["density"] = (1.178 + 0.05 * cx * cy * cz) * _noise(g, 5)
        f["viscosity"] = (1.2e-5 * (2.0 + cx * cy * cz)) * _noise(g, 6)
        f["momentum_diag"] = (gamma1 / dt) * (1.25 + 0.75 * sx * cy) * _noise(g, 7)
        dp = np.zeros((n, 3))
        dp[:, 0] = 0.5 * np.sin(2 * tp * x[:, 0]) * _noise(g, 8)
        dp[:, 1] = 0.5 * np.sin(2 * tp * x[:, 1]) * _noise(g, 9)
        dp[:, 2] = 0.05 * sz * _noise(g, 10)
        f["dpdx"] = dp
        du = np.zeros((n, 9))
        for k in range(9):
            du[:, k] = 0.3 * np.sin(tp * x[:, k % 3] + 0.7 * k) * _noise(g, 11 + k)
        f["dudx"] = du
        f["turbulent_ke"] = (2.0 + cx * sy * cz) * _noise(g, 21)
        dk = np.zeros((n, 3))
        dk[:, 0] = -sx * sy * cz * _noise(g, 22)
        dk[:, 1] = cx * cy * cz * _noise(g, 23)
        dk[:, 2] = -cx * sy * sz * _noise(g, 24)
        f["dkdx"] = dk
        f["specific_dissipation_rate"] = (2.0 + cx * sy * sz) * _noise(g, 25)
        dw = np.zeros((n, 3))
        dw[:, 0] = -sx * sy * sz * _noise(g, 26)
        dw[:, 1] = cx * cy * sz * _noise(g, 27)
        dw[:, 2] = cx * sy * cz * _noise(g, 28)
        f["dwdx"] = dw
        f["effective_viscosity_tke"] = (3.0e-5 * (2.0 + sx * cy * cz)) * _noise(g, 29)
        f["effective_viscosity_sdr"] = (4.0e-5 * (2.0 + cx * sy * cz)) * _noise(g, 30)
        f["abl_wall_no_slip_wall_func_node_mask"] = np.ones(n)
    return f


def state_chunked(coords, gid, lengths, dt=0.5, gamma1=1.5, periodic_gid=None,
                  threads=None, chunk=1 << 20):
    """state() evaluated on node chunks in a thread pool (numpy releases the
    GIL inside its loops): every field is a pure function of the node, so the
    result is identical to state(); used by bench.py for the 10^7..10^8-node
    partitions of the 512^3 configuration."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    n = len(coords)
    if n <= chunk:
        return state(coords, gid, lengths, dt, gamma1, periodic_gid)
    threads = threads or min(16, os.cpu_count() or 1)
    cuts = list(range(0, n, chunk)) + [n]

    def part(i):
        a, b = cuts[i], cuts[i + 1]
        return state(coords[a:b], gid[a:b], lengths, dt, gamma1,
                     None if periodic_gid is None else periodic_gid[a:b])
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(part, range(len(cuts) - 1)))
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}


NODE_FIELDS = {
    "velocity": 3, "pressure": 1, "density": 1, "viscosity": 1,
    "momentum_diag": 1, "dpdx": 3, "dudx": 9, "turbulent_ke": 1, "dkdx": 3,
    "specific_dissipation_rate": 1, "dwdx": 3, "effective_viscosity_tke": 1,
    "effective_viscosity_sdr": 1, "abl_wall_no_slip_wall_func_node_mask": 1,
    "dual_nodal_volume": 1,
}
