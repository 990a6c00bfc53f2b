/*
 * meshgen.cpp -- synthetic structured hex mesh generator (TEST / BENCH INPUT
 * ONLY; not part of the product path).
 *
 * Produces, for one rank of a z-slab decomposition of an nx*ny*nz element box,
 * exactly the inputs the edge assembly path consumes in nalu-wind:
 *   - local nodes (owned + shared) with STK-`generated:`-style global ids
 *     (id-1 = i + (nx+1) j + (nx+1)(ny+1) k), coordinates, owner rank,
 *   - hypre row ids numbered as Realm::set_hypre_global_id does
 *     (src/Realm.C:3587-3679: rank-contiguous, ascending global id inside a
 *     rank), with lateral-periodic slaves resolved to their master's id
 *     (HypreLinearSystem::get_entity_hypre_id, src/HypreLinearSystem.C:2460-2470),
 *   - the locally-owned edges, nodes ordered by ascending global id (the
 *     orientation the reference's golds pin),
 *   - edge_area_vector and dual_nodal_volume from the CVFEM dual mesh of the
 *     trilinear hexes (the quantities GeometryInteriorAlg produces,
 *     src/ngp_algorithms/GeometryInteriorAlg.C:165-225: sub-control-surface
 *     area vectors summed per edge with the L->R sign rule, sub-control-volume
 *     volumes summed per node; already summed over ranks).
 * Ownership follows STK's convention: a node / edge shared by several ranks is
 * owned by the lowest rank.
 */
#if defined(_OPENMP)
#include <omp.h>
#endif
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

namespace {

struct Gen
{
  int nx, ny, nz;
  double Lx, Ly, Lz;
  int perX, perY;
  double warp;   /* curvilinear amplitude in units of the cell size */
  double zstretch; /* geometric growth factor in z (1 = uniform) */
  int nranks, rank;
  int shuffleBucket; /* >0: shuffle edges inside buckets of this size */
  uint64_t seed;

  /* slab ownership */
  int k0, k1;             /* element k range [k0,k1) */
  int nodeK0, nodeK1;     /* local node planes [nodeK0, nodeK1] */
  int ownK0;              /* first owned node plane */

  std::vector<double> zc; /* stretched z coordinates of planes */

  /* results */
  std::vector<double> coords;
  std::vector<int64_t> gid, hid, ownHid;
  std::vector<int32_t> owner;
  std::vector<int64_t> offsets;
  std::vector<int32_t> edges;
  std::vector<double> area, vol;
};

inline int64_t
gid_of(const Gen& g, int i, int j, int k)
{
  return 1 + i + int64_t(g.nx + 1) * j + int64_t(g.nx + 1) * (g.ny + 1) * k;
}

void
node_xyz(const Gen& g, int i, int j, int k, double* x)
{
  const double hx = g.Lx / g.nx, hy = g.Ly / g.ny;
  const double px = i * hx, py = j * hy, pz = g.zc[k];
  const double hz = g.Lz / g.nz;
  x[0] = px;
  x[1] = py;
  x[2] = pz;
  if (g.warp != 0.0) {
    const double tp = 2.0 * M_PI;
    /* smooth, periodic in x and y, vanishing on the z boundaries */
    const double sz = std::sin(M_PI * pz / g.Lz);
    x[0] += g.warp * hx * std::sin(tp * py / g.Ly) * sz;
    x[1] += g.warp * hy * std::sin(tp * px / g.Lx) * sz;
    x[2] += g.warp * hz * std::sin(tp * px / g.Lx) * std::sin(tp * py / g.Ly) *
            sz;
  }
}

/* trilinear map of the hex with corner coordinates X[8][3] (standard order) */
inline void
tri(const double X[8][3], double a, double b, double c, double* p)
{
  const double w[8] = {(1 - a) * (1 - b) * (1 - c), a * (1 - b) * (1 - c),
                       a * b * (1 - c),             (1 - a) * b * (1 - c),
                       (1 - a) * (1 - b) * c,       a * (1 - b) * c,
                       a * b * c,                   (1 - a) * b * c};
  for (int d = 0; d < 3; ++d) {
    p[d] = 0;
    for (int n = 0; n < 8; ++n)
      p[d] += w[n] * X[n][d];
  }
}

/* area vector of a quad q[4][3] by a 4-triangle fan about its mid point */
void
quad_area(const double q[4][3], double* a)
{
  double m[3];
  for (int d = 0; d < 3; ++d)
    m[d] = 0.25 * (q[0][d] + q[1][d] + q[2][d] + q[3][d]);
  a[0] = a[1] = a[2] = 0;
  for (int t = 0; t < 4; ++t) {
    const double* p1 = q[t];
    const double* p2 = q[(t + 1) % 4];
    const double r1[3] = {p1[0] - m[0], p1[1] - m[1], p1[2] - m[2]};
    const double r2[3] = {p2[0] - m[0], p2[1] - m[1], p2[2] - m[2]};
    a[0] += 0.5 * (r1[1] * r2[2] - r2[1] * r1[2]);
    a[1] += 0.5 * (r1[2] * r2[0] - r2[2] * r1[0]);
    a[2] += 0.5 * (r1[0] * r2[1] - r2[0] * r1[1]);
  }
}

/* volume of a hexahedron H[8][3] (standard order): 6 faces x 4 triangles, each
 * forming a tet with the cell centroid */
double
hex_volume(const double H[8][3])
{
  static const int F[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4},
                              {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}};
  double c[3] = {0, 0, 0};
  for (int n = 0; n < 8; ++n)
    for (int d = 0; d < 3; ++d)
      c[d] += 0.125 * H[n][d];
  double vol = 0;
  for (int f = 0; f < 6; ++f) {
    double m[3] = {0, 0, 0};
    for (int v = 0; v < 4; ++v)
      for (int d = 0; d < 3; ++d)
        m[d] += 0.25 * H[F[f][v]][d];
    for (int t = 0; t < 4; ++t) {
      const double* p1 = H[F[f][t]];
      const double* p2 = H[F[f][(t + 1) % 4]];
      const double a[3] = {m[0] - c[0], m[1] - c[1], m[2] - c[2]};
      const double b[3] = {p1[0] - c[0], p1[1] - c[1], p1[2] - c[2]};
      const double e[3] = {p2[0] - c[0], p2[1] - c[1], p2[2] - c[2]};
      vol += (a[0] * (b[1] * e[2] - b[2] * e[1]) -
              a[1] * (b[0] * e[2] - b[2] * e[0]) +
              a[2] * (b[0] * e[1] - b[1] * e[0])) /
             6.0;
    }
  }
  return std::fabs(vol);
}

void
generate(Gen& g)
{
  const int nx = g.nx, ny = g.ny, nz = g.nz;
  /* z planes */
  g.zc.resize(nz + 1);
  if (g.zstretch == 1.0) {
    for (int k = 0; k <= nz; ++k)
      g.zc[k] = g.Lz * k / nz;
  } else {
    double s = 0, h = 1;
    std::vector<double> hh(nz);
    for (int k = 0; k < nz; ++k) {
      hh[k] = h;
      s += h;
      h *= g.zstretch;
    }
    g.zc[0] = 0;
    for (int k = 0; k < nz; ++k)
      g.zc[k + 1] = g.zc[k] + g.Lz * hh[k] / s;
  }
  /* slabs */
  auto kBegin = [&](int r) { return int(int64_t(nz) * r / g.nranks); };
  g.k0 = kBegin(g.rank);
  g.k1 = kBegin(g.rank + 1);
  g.nodeK0 = g.k0;
  g.nodeK1 = g.k1;
  g.ownK0 = g.rank == 0 ? 0 : g.k0 + 1;
  const int64_t plane = int64_t(nx + 1) * (ny + 1);
  const int nkLocal = g.nodeK1 - g.nodeK0 + 1;
  const int64_t N = plane * nkLocal;

  /* hypre offsets: owned nodes per rank */
  g.offsets.assign(g.nranks + 1, 0);
  for (int r = 0; r < g.nranks; ++r) {
    const int a = r == 0 ? 0 : kBegin(r) + 1, b = kBegin(r + 1);
    g.offsets[r + 1] = g.offsets[r] + plane * (b - a + 1);
  }
  auto ownerOfPlane = [&](int k) {
    /* plane k is owned by the lowest rank that has it */
    for (int r = 0; r < g.nranks; ++r)
      if (k <= kBegin(r + 1))
        return r;
    return g.nranks - 1;
  };
  auto hidOf = [&](int i, int j, int k) -> int64_t {
    const int r = ownerOfPlane(k);
    const int a = r == 0 ? 0 : kBegin(r) + 1;
    return g.offsets[r] + int64_t(k - a) * plane + int64_t(nx + 1) * j + i;
  };
  auto lidOf = [&](int i, int j, int k) -> int32_t {
    return int32_t(int64_t(k - g.nodeK0) * plane + int64_t(nx + 1) * j + i);
  };

  g.coords.resize(size_t(N) * 3);
  g.gid.resize(N);
  g.hid.resize(N);
  g.ownHid.resize(N);
  g.owner.resize(N);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = g.nodeK0; k <= g.nodeK1; ++k)
    for (int j = 0; j <= ny; ++j)
      for (int i = 0; i <= nx; ++i) {
        const int32_t l = lidOf(i, j, k);
        node_xyz(g, i, j, k, &g.coords[size_t(l) * 3]);
        g.gid[l] = gid_of(g, i, j, k);
        g.owner[l] = ownerOfPlane(k);
        g.ownHid[l] = hidOf(i, j, k);
        const int im = (g.perX && i == nx) ? 0 : i;
        const int jm = (g.perY && j == ny) ? 0 : j;
        g.hid[l] = hidOf(im, jm, k);
      }

  /* owned edges: all edges of my elements except those lying in the bottom
   * plane when that plane belongs to the lower rank */
  const int kLo = g.rank == 0 ? g.k0 : g.k0 + 1; /* in-plane edges from here */
  std::vector<int32_t>& E = g.edges;
  E.clear();
  for (int k = g.nodeK0; k <= g.nodeK1; ++k)
    for (int j = 0; j <= ny; ++j)
      for (int i = 0; i <= nx; ++i) {
        if (k >= kLo) {
          if (i < nx) {
            E.push_back(lidOf(i, j, k));
            E.push_back(lidOf(i + 1, j, k));
          }
          if (j < ny) {
            E.push_back(lidOf(i, j, k));
            E.push_back(lidOf(i, j + 1, k));
          }
        }
        if (k < g.nodeK1) {
          E.push_back(lidOf(i, j, k));
          E.push_back(lidOf(i, j, k + 1));
        }
      }
  int64_t nE = (int64_t)E.size() / 2;
  if (g.shuffleBucket > 1) {
    std::mt19937_64 rng(g.seed);
    for (int64_t b = 0; b < nE; b += g.shuffleBucket) {
      const int64_t e = std::min<int64_t>(nE, b + g.shuffleBucket);
      for (int64_t q = e - 1; q > b; --q) {
        const int64_t r = b + int64_t(rng() % uint64_t(q - b + 1));
        std::swap(E[2 * q], E[2 * r]);
        std::swap(E[2 * q + 1], E[2 * r + 1]);
      }
    }
  }

  /* geometry: accumulate SCS areas on edges and SCV volumes on nodes from all
   * elements touching my local nodes (including the neighbour ranks' layer) */
  const bool uniform = (g.warp == 0.0);
  g.vol.assign(N, 0.0);
  /* edge accumulators indexed by (node, direction) */
  std::vector<double> acc(size_t(N) * 9, 0.0);
  const int ek0 = std::max(0, g.k0 - 1), ek1 = std::min(nz, g.k1 + 1);
  /* reference-space description of the 12 SCS: axis, (b,c) in other axes */
  for (int k = ek0; k < ek1; ++k) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        double X[8][3];
        const int ci[8] = {0, 1, 1, 0, 0, 1, 1, 0};
        const int cj[8] = {0, 0, 1, 1, 0, 0, 1, 1};
        const int ck[8] = {0, 0, 0, 0, 1, 1, 1, 1};
        for (int n = 0; n < 8; ++n)
          node_xyz(g, i + ci[n], j + cj[n], k + ck[n], X[n]);
        /* SCV volumes */
        for (int n = 0; n < 8; ++n) {
          const int kk = k + ck[n];
          if (kk < g.nodeK0 || kk > g.nodeK1)
            continue;
          double v;
          if (uniform) {
            v = 0.125 * (g.Lx / nx) * (g.Ly / ny) * (g.zc[k + 1] - g.zc[k]);
          } else {
            double H[8][3];
            const double a0 = 0.5 * ci[n], b0 = 0.5 * cj[n], c0 = 0.5 * ck[n];
            for (int m = 0; m < 8; ++m)
              tri(X, a0 + 0.5 * ci[m], b0 + 0.5 * cj[m], c0 + 0.5 * ck[m], H[m]);
            v = hex_volume(H);
          }
#pragma omp atomic
          g.vol[lidOf(i + ci[n], j + cj[n], kk)] += v;
        }
        /* SCS area vectors: for axis ax and the 4 edges parallel to it */
        for (int ax = 0; ax < 3; ++ax)
          for (int b = 0; b < 2; ++b)
            for (int c = 0; c < 2; ++c) {
              /* the edge starts at corner with coordinate 0 on axis ax */
              int o[3];
              o[ax] = 0;
              o[(ax + 1) % 3] = b;
              o[(ax + 2) % 3] = c;
              const int kk = k + o[2];
              if (kk < g.nodeK0 || kk > g.nodeK1)
                continue;
              if (ax == 2 && kk + 1 > g.nodeK1)
                continue;
              double q[4][3];
              const double bb = b, cc = c;
              const double pts[4][2] = {{bb, cc}, {0.5, cc}, {0.5, 0.5}, {bb, 0.5}};
              for (int v = 0; v < 4; ++v) {
                double abc[3];
                abc[ax] = 0.5;
                abc[(ax + 1) % 3] = pts[v][0];
                abc[(ax + 2) % 3] = pts[v][1];
                tri(X, abc[0], abc[1], abc[2], q[v]);
              }
              double a[3];
              quad_area(q, a);
              /* orient L -> R (ascending global id == +axis direction) */
              double xl[3], xr[3];
              {
                double abc[3];
                abc[ax] = 0;
                abc[(ax + 1) % 3] = bb;
                abc[(ax + 2) % 3] = cc;
                tri(X, abc[0], abc[1], abc[2], xl);
                abc[ax] = 1;
                tri(X, abc[0], abc[1], abc[2], xr);
              }
              const double dot = a[0] * (xr[0] - xl[0]) + a[1] * (xr[1] - xl[1]) +
                                 a[2] * (xr[2] - xl[2]);
              const double sg = dot < 0 ? -1.0 : 1.0;
              const int32_t nl = lidOf(i + o[0], j + o[1], kk);
              for (int d = 0; d < 3; ++d) {
#pragma omp atomic
                acc[size_t(nl) * 9 + ax * 3 + d] += sg * a[d];
              }
            }
      }
  }
  /* periodic sum of the dual volume (the realm's periodic manager sums
   * dual_nodal_volume over master/slave copies) */
  if (g.perX || g.perY) {
    std::vector<double> tot(g.vol);
    for (int k = g.nodeK0; k <= g.nodeK1; ++k)
      for (int j = 0; j <= ny; ++j)
        for (int i = 0; i <= nx; ++i) {
          const int im = (g.perX && i == nx) ? 0 : i;
          const int jm = (g.perY && j == ny) ? 0 : j;
          if (im != i || jm != j)
            tot[lidOf(im, jm, k)] += g.vol[lidOf(i, j, k)];
        }
    for (int k = g.nodeK0; k <= g.nodeK1; ++k)
      for (int j = 0; j <= ny; ++j)
        for (int i = 0; i <= nx; ++i) {
          const int im = (g.perX && i == nx) ? 0 : i;
          const int jm = (g.perY && j == ny) ? 0 : j;
          g.vol[lidOf(i, j, k)] = tot[lidOf(im, jm, k)];
        }
  }
  nE = (int64_t)E.size() / 2;
  g.area.resize(size_t(nE) * 3);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nE; ++e) {
    const int32_t l = E[2 * e], r = E[2 * e + 1];
    const int64_t diff = int64_t(r) - l;
    const int ax = diff == 1 ? 0 : (diff == nx + 1 ? 1 : 2);
    for (int d = 0; d < 3; ++d)
      g.area[size_t(e) * 3 + d] = acc[size_t(l) * 9 + ax * 3 + d];
  }
}

} // namespace

extern "C" {

struct mg_params
{
  int32_t nx, ny, nz;
  double Lx, Ly, Lz;
  int32_t periodic_x, periodic_y;
  double warp, zstretch;
  int32_t nranks, rank;
  int32_t shuffle_bucket;
  uint64_t seed;
};

void*
mg_generate(const mg_params* p)
{
  Gen* g = new Gen;
  g->nx = p->nx;
  g->ny = p->ny;
  g->nz = p->nz;
  g->Lx = p->Lx;
  g->Ly = p->Ly;
  g->Lz = p->Lz;
  g->perX = p->periodic_x;
  g->perY = p->periodic_y;
  g->warp = p->warp;
  g->zstretch = p->zstretch > 0 ? p->zstretch : 1.0;
  g->nranks = p->nranks;
  g->rank = p->rank;
  g->shuffleBucket = p->shuffle_bucket;
  g->seed = p->seed;
#if defined(_OPENMP)
  if (const char* e = std::getenv("NW_HOST_THREADS")) /* see plan.cpp */
    if (std::atoi(e) > 0)
      omp_set_num_threads(std::atoi(e));
#endif
  generate(*g);
  return g;
}

void
mg_free(void* h)
{
  delete static_cast<Gen*>(h);
}

int64_t
mg_num_nodes(void* h)
{
  return (int64_t) static_cast<Gen*>(h)->gid.size();
}
int64_t
mg_num_edges(void* h)
{
  return (int64_t) static_cast<Gen*>(h)->edges.size() / 2;
}

void
mg_copy(
  void* h,
  double* coords,
  int64_t* gid,
  int64_t* hid,
  int64_t* own_hid,
  int32_t* owner,
  int64_t* offsets,
  int32_t* edges,
  double* area,
  double* vol)
{
  Gen& g = *static_cast<Gen*>(h);
  auto cp = [](auto* dst, const auto& v) {
    if (dst && !v.empty())
      std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
  };
  cp(coords, g.coords);
  cp(gid, g.gid);
  cp(hid, g.hid);
  cp(own_hid, g.ownHid);
  cp(owner, g.owner);
  cp(offsets, g.offsets);
  cp(edges, g.edges);
  cp(area, g.area);
  cp(vol, g.vol);
}

} // extern "C"
