"""The reference's one-element hex8 fixtures, restated from numbers only.

unit_tests/kernels/UnitTestKernelUtils.h:268-595, 1357-1447 and
unit_tests/kernels/UnitTestKernelUtils.C:31-330, 667-768: STK `generated:1x1xN`
mesh, node id-1 = i + 2j + 4k, edges stored in hex edge-ordinal order with each
edge's nodes ordered by ascending global id, edge_area_vector = 0.25 * e_axis,
dual_nodal_volume = 0.125, trig fields with a = 0.3.
"""
import numpy as np

A = 0.3
PI = np.arccos(-1.0)

# hex8 local edge ordinals (local node pairs) and local->(id-1) node map
HEX_EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4),
             (0, 4), (1, 5), (2, 6), (3, 7)]
HEX_LOCAL_TO_ID = [0, 1, 3, 2, 4, 5, 7, 6]


def mesh(nz=1):
    """generated:1x1x<nz>: nodes, edges (element traversal, first visit)"""
    nn = 4 * (nz + 1)
    coords = np.zeros((nn, 3))
    for k in range(nz + 1):
        for j in range(2):
            for i in range(2):
                coords[i + 2 * j + 4 * k] = (i, j, k)
    edges = []
    seen = set()
    for el in range(nz):
        for (a, b) in HEX_EDGES:
            na = HEX_LOCAL_TO_ID[a] + 4 * el
            nb = HEX_LOCAL_TO_ID[b] + 4 * el
            e = (min(na, nb), max(na, nb))
            if e not in seen:
                seen.add(e)
                edges.append(e)
    edges = np.array(edges, dtype=np.int32)
    return coords, edges


def edge_area(coords, edges, nz=1):
    """calc_edge_area_vec (UnitTestKernelUtils.C:943-1015): each rank sums the
    SCS areas of ITS OWN element only (no parallel sum in the fixture), so every
    edge -- including the interface edges of generated:1x1x2 -- carries one
    quarter face, 0.25, directed L->R."""
    d = coords[edges[:, 1]] - coords[edges[:, 0]]
    av = np.zeros_like(d)
    for e in range(len(edges)):
        axis = int(np.argmax(np.abs(d[e])))
        av[e, axis] = 0.25 * np.sign(d[e, axis])
    return av


def velocity(c):
    x, y = c[:, 0], c[:, 1]
    u = np.zeros((len(c), 3))
    u[:, 0] = -np.cos(A * PI * x) * np.sin(A * PI * y)
    u[:, 1] = +np.sin(A * PI * x) * np.cos(A * PI * y)
    return u


def dudx(c):
    x, y = c[:, 0], c[:, 1]
    ap = A * PI
    cx, sx, cy, sy = np.cos(ap * x), np.sin(ap * x), np.cos(ap * y), np.sin(ap * y)
    g = np.zeros((len(c), 9))
    g[:, 0] = ap * sx * sy
    g[:, 1] = -ap * cx * cy
    g[:, 3] = ap * cx * cy
    g[:, 4] = -ap * sx * sy
    return g


def pressure(c):
    x, y = c[:, 0], c[:, 1]
    return -1.0 / 4.0 * (np.cos(2.0 * A * PI * x) + np.cos(2.0 * A * PI * y))


def dpdx(c):
    x, y = c[:, 0], c[:, 1]
    g = np.zeros((len(c), 3))
    g[:, 0] = 0.5 * A * PI * np.sin(2.0 * A * PI * x)
    g[:, 1] = 0.5 * A * PI * np.sin(2.0 * A * PI * y)
    return g


def true_mdot(edges, vel, rho, av):
    """calc_mass_flow_rate, UnitTestKernelUtils.C:725-768"""
    l, r = edges[:, 0], edges[:, 1]
    return np.sum(0.5 * (rho[l, None] * vel[l] + rho[r, None] * vel[r]) * av,
                  axis=1)


def fixture_mdot(edges, vel, rho, av):
    """What the edge kernels actually read in the fixtures: `mass_flow_rate` is
    registered with 3 components per edge (UnitTestKernelUtils.h:569-570,
    1408-1409) but filled as a packed 1-per-edge array, and the kernels read
    component 0 => effective mdot[e] = packed[3e] (0 beyond the bucket)."""
    t = true_mdot(edges, vel, rho, av)
    buf = np.zeros(3 * len(edges))
    buf[:len(edges)] = t
    return buf[0::3].copy()


def mixture_fraction_fields(c):
    """MixtureFractionKernelHex8Mesh (UnitTestKernelUtils.h:1357-1447): the
    fixture passes (amf_, znot_) into (znot, amf) => Z = 2 cos(pi x)cos(pi y)cos(pi z)"""
    znot, amf = 2.0, 1.0
    z = znot * np.cos(amf * PI * c[:, 0]) * np.cos(amf * PI * c[:, 1]) * \
        np.cos(amf * PI * c[:, 2])
    rho_p, rho_s = 0.163, 1.18
    mu_p, mu_s = 1.967e-5, 1.85e-5
    rho = 1.0 / (z / rho_p + (1.0 - z) / rho_s)
    visc = mu_p * z + mu_s * (1.0 - z)
    return z, rho, visc
