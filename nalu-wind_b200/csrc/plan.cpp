/*
 * plan.cpp -- host-side plan builder (see plan.h).  Pure C++17 + OpenMP.
 */
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <stdexcept>
#if defined(_OPENMP)
#include <omp.h>
#endif
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace nw {

namespace {

[[noreturn]] void
fail(const std::string& m)
{
  throw std::runtime_error(m);
}

/* NW_PLAN_TIMING=1: wall-clock of the plan-builder stages on stderr */
struct StageTimer
{
  bool on;
  std::chrono::steady_clock::time_point t0;
  StageTimer() : on(std::getenv("NW_PLAN_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* what)
  {
    if (!on)
      return;
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[nw plan] %-28s %8.3f s\n", what,
                 std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};

/* NW_HOST_THREADS=n: OpenMP threads of the plan builder.  Launchers such as
 * torchrun export OMP_NUM_THREADS=1 to every rank; a caller that knows how many
 * cores one rank may use (bench.py: cores / ranks) says so here. */
void
apply_host_threads()
{
#if defined(_OPENMP)
  if (const char* e = std::getenv("NW_HOST_THREADS")) {
    const int n = std::atoi(e);
    if (n > 0)
      omp_set_num_threads(n);
  }
#endif
}

inline int64_t
even_up(int64_t v)
{
  return (v + 1) & ~int64_t(1);
}

/* ---- recursive coordinate bisection over row groups ---- */

/* items: group ids; pts: representative coordinates [G][3].  Emits leaves in
 * depth-first (left first) order, which keeps spatially adjacent tiles close
 * in launch order (L2 reuse of halo nodes). */
/* one bisection step of items[begin, end) into `leaves` leaves; returns the
 * split position (the same arithmetic rule rcb_leaf_bounds uses) */
int64_t
rcb_split(
  std::vector<int32_t>& items,
  const std::vector<double>& pts,
  int64_t begin,
  int64_t end,
  int64_t leaves)
{
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t i = begin; i < end; ++i) {
    const double* p = &pts[size_t(items[i]) * 3];
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(lo[d], p[d]);
      hi[d] = std::max(hi[d], p[d]);
    }
  }
  int dim = 0;
  for (int d = 1; d < 3; ++d)
    if (hi[d] - lo[d] > hi[dim] - lo[dim])
      dim = d;
  const int d1 = (dim + 1) % 3, d2 = (dim + 2) % 3;
  const int64_t l1 = leaves / 2;
  const int64_t mid = begin + (end - begin) * l1 / leaves;
  auto cmp = [&](int32_t a, int32_t b) {
    const double* pa = &pts[size_t(a) * 3];
    const double* pb = &pts[size_t(b) * 3];
    if (pa[dim] != pb[dim])
      return pa[dim] < pb[dim];
    if (pa[d1] != pb[d1])
      return pa[d1] < pb[d1];
    if (pa[d2] != pb[d2])
      return pa[d2] < pb[d2];
    return a < b;
  };
  std::nth_element(
    items.begin() + begin, items.begin() + mid, items.begin() + end, cmp);
  return mid;
}

void
rcb_part(
  std::vector<int32_t>& items,
  const std::vector<double>& pts,
  int64_t begin,
  int64_t end,
  int64_t leaves)
{
  if (leaves <= 1 || end - begin <= 1)
    return;
  const int64_t mid = rcb_split(items, pts, begin, end, leaves);
  const int64_t l1 = leaves / 2;
  /* the two halves are independent: sub-trees above a size threshold become
   * OpenMP tasks (the result does not depend on the execution order) */
  if (end - begin > 200000) {
#pragma omp task shared(items, pts)
    rcb_part(items, pts, begin, mid, l1);
#pragma omp task shared(items, pts)
    rcb_part(items, pts, mid, end, leaves - l1);
#pragma omp taskwait
  } else {
    rcb_part(items, pts, begin, mid, l1);
    rcb_part(items, pts, mid, end, leaves - l1);
  }
}

/* the leaf boundaries follow from the split rule alone (no coordinates) */
void
rcb_leaf_bounds(
  int64_t begin, int64_t end, int64_t leaves, std::vector<int64_t>& leafBegin)
{
  if (leaves <= 1 || end - begin <= 1) {
    leafBegin.push_back(begin);
    return;
  }
  const int64_t l1 = leaves / 2;
  const int64_t mid = begin + (end - begin) * l1 / leaves;
  rcb_leaf_bounds(begin, mid, l1, leafBegin);
  rcb_leaf_bounds(mid, end, leaves - l1, leafBegin);
}

/* items: group ids; pts: representative coordinates [G][3].  Leaves come out
 * in depth-first (left first) order, which keeps spatially adjacent tiles
 * close in launch order (L2 reuse of halo nodes). */
void
rcb(
  std::vector<int32_t>& items,
  const std::vector<double>& pts,
  int64_t nLeaves,
  std::vector<int64_t>& leafBegin)
{
  leafBegin.clear();
#pragma omp parallel
#pragma omp single
  rcb_part(items, pts, 0, (int64_t)items.size(), nLeaves);
  rcb_leaf_bounds(0, (int64_t)items.size(), nLeaves, leafBegin);
  leafBegin.push_back((int64_t)items.size());
}

} // namespace

void
split_half_edges(const uint32_t* he, int n, int nWarps, int32_t* split)
{
  split[0] = 0;
  for (int w = 1; w < nWarps; ++w) {
    int pos = int((int64_t(n) * w) / nWarps);
    if (pos < split[w - 1])
      pos = split[w - 1];
    /* move forward to an entity boundary */
    while (pos > 0 && pos < n && he_ent(he[pos]) == he_ent(he[pos - 1]))
      ++pos;
    split[w] = pos;
  }
  split[nWarps] = n;
}

int64_t
append_sliced_ell(
  const uint32_t* he,
  int n,
  int nEnts,
  std::vector<uint32_t>& ell,
  std::vector<int32_t>& sliceOff)
{
  const int nSlices = (nEnts + 31) / 32;
  std::vector<int32_t> first(nEnts + 1, 0);
  for (int q = 0; q < n; ++q)
    first[he_ent(he[q]) + 1]++;
  for (int i = 0; i < nEnts; ++i)
    first[i + 1] += first[i];
  const size_t base = ell.size();
  int64_t off = 0;
  for (int s = 0; s < nSlices; ++s) {
    sliceOff.push_back((int32_t)off);
    int w = 0;
    for (int i = 32 * s; i < std::min(nEnts, 32 * s + 32); ++i)
      w = std::max(w, first[i + 1] - first[i]);
    ell.resize(base + off + size_t(w) * 32, 0u);
    for (int i = 32 * s; i < std::min(nEnts, 32 * s + 32); ++i)
      for (int k = 0; k < first[i + 1] - first[i]; ++k)
        ell[base + off + size_t(k) * 32 + (i - 32 * s)] = he[first[i] + k];
    off += int64_t(w) * 32;
  }
  sliceOff.push_back((int32_t)off);
  return off;
}

void
build_mesh_plan(const MeshInput& in, MeshPlan& mp)
{
  if (in.ndim != 2 && in.ndim != 3)
    fail("nw_mesh_create: ndim must be 2 or 3");
  if (in.nNodes <= 0 || in.nEdges < 0)
    fail("nw_mesh_create: empty mesh");
  if (in.nNodes >= (int64_t(1) << 31) - 16 || in.nEdges >= (int64_t(1) << 30))
    fail("nw_mesh_create: partition too large for 32-bit local indices");
  if (!in.edgeNodes && in.nEdges > 0)
    fail("nw_mesh_create: edge_nodes is NULL");
  if (!in.nodeHid || !in.coords || !in.hypreOffsets)
    fail("nw_mesh_create: node_hypre_id / coords / hypre_offsets is NULL");
  if (in.rank < 0 || in.rank >= in.nranks)
    fail("nw_mesh_create: bad rank");

  const int64_t N = in.nNodes, E = in.nEdges;
  const int nd = in.ndim;
  apply_host_threads();
  StageTimer tm;
  mp = MeshPlan();
  mp.ndim = nd;
  mp.rank = in.rank;
  mp.nranks = in.nranks;
  mp.nNodes = N;
  mp.nEdges = E;
  mp.hypreOffsets.assign(in.hypreOffsets, in.hypreOffsets + in.nranks + 1);
  mp.iLowerNode = mp.hypreOffsets[in.rank];
  mp.iUpperNode = mp.hypreOffsets[in.rank + 1] - 1;
  mp.edgeNodes.assign(in.edgeNodes, in.edgeNodes + 2 * E);
  mp.nodeHid.assign(in.nodeHid, in.nodeHid + N);
  for (int64_t e = 0; e < 2 * E; ++e)
    if (mp.edgeNodes[e] < 0 || mp.edgeNodes[e] >= N)
      fail("nw_mesh_create: edge node index out of range");

  /* default 192: momentum then fits two CTAs per SM (<= 113 KB of shared
   * memory each) and the ~700 tile-edges fill three rounds of 256 threads;
   * measured best of 128..256 on a 128^3 hex box (profiles/r01b_*) */
  int T = in.tileNodes > 0 ? in.tileNodes : 192;
  if (T > kMaxTileEnts)
    T = kMaxTileEnts;
  if (T < 8)
    T = 8;

  /* ---- row groups: nodes sharing one (periodic-resolved) row id ---- */
  std::vector<int32_t> order(N);
  std::iota(order.begin(), order.end(), 0);
  auto ownHid = [&](int32_t n) {
    return in.nodeOwnHid ? in.nodeOwnHid[n] : in.nodeHid[n];
  };
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    if (in.nodeHid[a] != in.nodeHid[b])
      return in.nodeHid[a] < in.nodeHid[b];
    const int64_t oa = ownHid(a), ob = ownHid(b);
    if (oa != ob)
      return oa < ob;
    return a < b;
  });
  std::vector<int64_t> gStart; /* into order[] */
  for (int64_t i = 0; i < N; ++i)
    if (i == 0 || in.nodeHid[order[i]] != in.nodeHid[order[i - 1]])
      gStart.push_back(i);
  const int64_t G = (int64_t)gStart.size();
  gStart.push_back(N);
  std::vector<double> pts(size_t(G) * 3, 0.0);
  for (int64_t g = 0; g < G; ++g) {
    /* representative: the node that owns the row id, else the first alias */
    int32_t rep = order[gStart[g]];
    for (int64_t i = gStart[g]; i < gStart[g + 1]; ++i)
      if (ownHid(order[i]) == in.nodeHid[order[i]]) {
        rep = order[i];
        break;
      }
    for (int d = 0; d < nd; ++d)
      pts[size_t(g) * 3 + d] = in.coords[size_t(rep) * nd + d];
  }

  tm.mark("mesh: copy + row groups");
  /* ---- tiles ---- */
  std::vector<int32_t> items(G);
  std::iota(items.begin(), items.end(), 0);
  const int64_t nLeaves = std::max<int64_t>(1, (N + T - 1) / T);
  std::vector<int64_t> leafBegin;
  rcb(items, pts, nLeaves, leafBegin);
  tm.mark("mesh: rcb");
  /* ---- cut refinement ----
   * Coordinate bisection cuts a curvilinear or unstructured mesh obliquely:
   * ragged tile boundaries, i.e. more cut edges (each is evaluated in both
   * tiles) and more halo nodes (profiles/r02f_bench_warped: 1.8 halo nodes
   * per node against 1.06 on the regular box).  When the cut is above what a
   * lattice gives, a few greedy passes move a row group to the neighbouring
   * tile that holds more of its edges than its own tile does (every move
   * lowers the cut; tiles may grow by ~8 % and never run empty).  Serial and
   * in group order, so the result is deterministic.
   * Opt-in (NW_TILE_REFINE=1).  Measured (profiles/r02s_bench_*): the tet /
   * wedge / pyramid mesh gains 4 % (halo nodes per node 1.49 -> 1.39, largest
   * staged tile 428 -> 363 nodes), the warped hex box LOSES 15 %: its cut
   * barely moves (1.392 -> 1.389: an oblique cut through a lattice is a
   * staircase no local move removes) while the grown tiles (208 nodes, 824
   * edges) cost the scalar kernel a resident CTA. */
  {
    const int64_t nT = (int64_t)leafBegin.size() - 1;
    const char* env = std::getenv("NW_TILE_REFINE");
    const bool wanted = env && env[0] == '1';
    if (wanted && nT > 1 && E > 0) {
      std::vector<int32_t> tileOfGroup(G), groupOfNode(N);
      for (int64_t t = 0; t < nT; ++t)
        for (int64_t gi = leafBegin[t]; gi < leafBegin[t + 1]; ++gi)
          tileOfGroup[items[gi]] = (int32_t)t;
      for (int64_t g = 0; g < G; ++g)
        for (int64_t i = gStart[g]; i < gStart[g + 1]; ++i)
          groupOfNode[order[i]] = (int32_t)g;
      int64_t cut = 0;
#pragma omp parallel for reduction(+ : cut) schedule(static)
      for (int64_t e = 0; e < E; ++e) {
        const int32_t ga = groupOfNode[mp.edgeNodes[2 * e]];
        const int32_t gb = groupOfNode[mp.edgeNodes[2 * e + 1]];
        cut += tileOfGroup[ga] != tileOfGroup[gb];
      }
      /* a hex lattice in tiles of T nodes cuts ~ (2 / cbrt(T)) / 2 of its
       * edges (0.19 at T = 192); start refining a fifth above that */
      const double lattice = 1.0 / std::cbrt((double)std::max(T, 8));
      if ((double)cut > 1.2 * lattice * (double)E) {
        std::vector<int64_t> aptr(G + 1, 0);
        for (int64_t e = 0; e < E; ++e) {
          const int32_t ga = groupOfNode[mp.edgeNodes[2 * e]];
          const int32_t gb = groupOfNode[mp.edgeNodes[2 * e + 1]];
          if (ga != gb) {
            aptr[ga + 1]++;
            aptr[gb + 1]++;
          }
        }
        for (int64_t g = 0; g < G; ++g)
          aptr[g + 1] += aptr[g];
        std::vector<int32_t> adj(aptr[G]);
        {
          std::vector<int64_t> fill(aptr.begin(), aptr.end() - 1);
          for (int64_t e = 0; e < E; ++e) {
            const int32_t ga = groupOfNode[mp.edgeNodes[2 * e]];
            const int32_t gb = groupOfNode[mp.edgeNodes[2 * e + 1]];
            if (ga != gb) {
              adj[fill[ga]++] = gb;
              adj[fill[gb]++] = ga;
            }
          }
        }
        std::vector<int32_t> tileSize(nT, 0);
        for (int64_t g = 0; g < G; ++g)
          tileSize[tileOfGroup[g]] += (int32_t)(gStart[g + 1] - gStart[g]);
        const int32_t tMax =
          (int32_t)std::min<int64_t>(kMaxTileEnts, T + std::max(2, T / 12));
        std::vector<std::pair<int32_t, int32_t>> cnt; /* (tile, edges) */
        for (int pass = 0; pass < 4; ++pass) {
          int64_t moved = 0;
          for (int64_t g = 0; g < G; ++g) {
            const int32_t t = tileOfGroup[g];
            cnt.clear();
            int32_t own = 0;
            for (int64_t q = aptr[g]; q < aptr[g + 1]; ++q) {
              const int32_t tn = tileOfGroup[adj[q]];
              if (tn == t) {
                ++own;
                continue;
              }
              bool found = false;
              for (auto& c : cnt)
                if (c.first == tn) {
                  ++c.second;
                  found = true;
                  break;
                }
              if (!found)
                cnt.push_back({tn, 1});
            }
            int32_t bestT = -1, bestC = own;
            for (const auto& c : cnt)
              if (c.second > bestC || (c.second == bestC && bestT >= 0 && c.first < bestT)) {
                bestT = c.first;
                bestC = c.second;
              }
            const int32_t gs = (int32_t)(gStart[g + 1] - gStart[g]);
            if (bestT >= 0 && bestC > own && tileSize[bestT] + gs <= tMax &&
                tileSize[t] - gs >= 1) {
              tileOfGroup[g] = bestT;
              tileSize[bestT] += gs;
              tileSize[t] -= gs;
              ++moved;
            }
          }
          if (moved == 0)
            break;
        }
        /* back to the leaf lists: groups of a tile in ascending group id */
        std::vector<int64_t> lb(nT + 1, 0);
        for (int64_t g = 0; g < G; ++g)
          lb[tileOfGroup[g] + 1]++;
        for (int64_t t = 0; t < nT; ++t)
          lb[t + 1] += lb[t];
        std::vector<int64_t> fill(lb.begin(), lb.end() - 1);
        for (int64_t g = 0; g < G; ++g)
          items[fill[tileOfGroup[g]]++] = (int32_t)g;
        leafBegin = lb;
      }
    }
  }
  tm.mark("mesh: cut refinement");
  const int64_t nTiles = (int64_t)leafBegin.size() - 1;
  mp.nTiles = nTiles;
  mp.tiles.assign(nTiles, TileHdr());
  mp.tileOfNode.assign(N, -1);
  mp.slotOfNode.assign(N, -1);

  /* nodes of each tile ordered by (resolved row id, own id): the groups of a
   * leaf sorted by group id are already in that order */
  int64_t slot = 0;
  std::vector<int64_t> tileNodeBegin(nTiles + 1, 0);
  for (int64_t t = 0; t < nTiles; ++t) {
    std::sort(items.begin() + leafBegin[t], items.begin() + leafBegin[t + 1]);
    slot = even_up(slot);
    TileHdr& h = mp.tiles[t];
    h.node0 = (int32_t)slot;
    int64_t cnt = 0;
    for (int64_t gi = leafBegin[t]; gi < leafBegin[t + 1]; ++gi) {
      const int32_t g = items[gi];
      for (int64_t i = gStart[g]; i < gStart[g + 1]; ++i) {
        const int32_t n = order[i];
        mp.tileOfNode[n] = (int32_t)t;
        mp.slotOfNode[n] = (int32_t)(slot + cnt);
        ++cnt;
      }
    }
    if (cnt > kMaxTileEnts)
      fail("nw_mesh_create: tile exceeds the per-tile node limit; lower "
           "tile_nodes");
    h.nOwn = (int32_t)cnt;
    h.nOwnPad = (int32_t)even_up(cnt);
    slot += cnt;
    mp.maxTileNodes = std::max<int64_t>(mp.maxTileNodes, cnt);
  }
  /* room so that a padded TMA copy of the last tile stays in bounds */
  mp.nSlots = even_up(slot) + 2;
  mp.nodeOfSlot.assign(mp.nSlots, -1);
  for (int64_t n = 0; n < N; ++n)
    mp.nodeOfSlot[mp.slotOfNode[n]] = (int32_t)n;

  tm.mark("mesh: tile numbering");
  /* ---- tile-edges ---- */
  std::vector<int64_t> cnt(nTiles + 1, 0);
  for (int64_t e = 0; e < E; ++e) {
    const int32_t tL = mp.tileOfNode[mp.edgeNodes[2 * e]];
    const int32_t tR = mp.tileOfNode[mp.edgeNodes[2 * e + 1]];
    cnt[tL]++;
    if (tR != tL)
      cnt[tR]++;
  }
  std::vector<int64_t> edge0(nTiles + 1, 0);
  for (int64_t t = 0; t < nTiles; ++t) {
    if (cnt[t] > kMaxTileEdges)
      fail("nw_mesh_create: tile exceeds the per-tile edge limit; lower "
           "tile_nodes");
    /* tile starts on a multiple of 4: 16-byte aligned runs of 4-byte records */
    edge0[t + 1] = (edge0[t] + cnt[t] + 3) & ~int64_t(3);
    mp.tiles[t].edge0 = (int32_t)edge0[t];
    mp.tiles[t].nEdges = (int32_t)cnt[t];
    mp.maxTileEdges = std::max(mp.maxTileEdges, cnt[t]);
  }
  if (edge0[nTiles] + 2 >= (int64_t(1) << 31))
    fail("nw_mesh_create: too many tile-edges for 32-bit indices");
  mp.nTileEdgeSlots = edge0[nTiles] + 4;
  mp.tileEdgeSrc.assign(mp.nTileEdgeSlots, -1);
  mp.lr.assign(mp.nTileEdgeSlots, 0u);
  mp.tileEdgePrimary.assign(mp.nTileEdgeSlots, 0);
  mp.primarySlotOfEdge.assign(E, -1);
  mp.secondSlotOfEdge.assign(E, -1);
  {
    std::vector<int64_t> fill(edge0.begin(), edge0.end() - 1);
    for (int64_t e = 0; e < E; ++e) {
      const int32_t tL = mp.tileOfNode[mp.edgeNodes[2 * e]];
      const int32_t tR = mp.tileOfNode[mp.edgeNodes[2 * e + 1]];
      mp.tileEdgeSrc[fill[tL]++] = (int32_t)e;
      if (tR != tL)
        mp.tileEdgeSrc[fill[tR]++] = (int32_t)e;
    }
  }

  tm.mark("mesh: tile-edge bins");
  std::vector<std::vector<int32_t>> haloPer(nTiles);
  std::vector<std::vector<uint32_t>> hePer(nTiles);
  std::string err;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t t = 0; t < nTiles; ++t) {
    TileHdr& h = mp.tiles[t];
    const int64_t e0 = h.edge0, ne = h.nEdges;
    /* halo = referenced nodes owned by other tiles, ascending slot */
    std::vector<int32_t>& halo = haloPer[t];
    for (int64_t j = 0; j < ne; ++j) {
      const int32_t e = mp.tileEdgeSrc[e0 + j];
      for (int s = 0; s < 2; ++s) {
        const int32_t n = mp.edgeNodes[2 * e + s];
        if (mp.tileOfNode[n] != t)
          halo.push_back(mp.slotOfNode[n]);
      }
    }
    std::sort(halo.begin(), halo.end());
    halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
    h.nHalo = (int32_t)halo.size();
    if (h.nOwnPad + (int64_t)halo.size() > kMaxTileStaged) {
#pragma omp critical
      err = "nw_mesh_create: tile stages too many nodes";
      continue;
    }
    auto local = [&](int32_t n) -> uint32_t {
      const int32_t s = mp.slotOfNode[n];
      if (mp.tileOfNode[n] == t)
        return uint32_t(s - h.node0);
      return uint32_t(
        h.nOwnPad +
        (std::lower_bound(halo.begin(), halo.end(), s) - halo.begin()));
    };
    struct TE
    {
      uint32_t l, r;
      int32_t e;
    };
    std::vector<TE> te(ne);
    for (int64_t j = 0; j < ne; ++j) {
      const int32_t e = mp.tileEdgeSrc[e0 + j];
      te[j] = {local(mp.edgeNodes[2 * e]), local(mp.edgeNodes[2 * e + 1]), e};
    }
    std::sort(te.begin(), te.end(), [](const TE& a, const TE& b) {
      if (a.l != b.l)
        return a.l < b.l;
      if (a.r != b.r)
        return a.r < b.r;
      return a.e < b.e;
    });
    std::vector<uint32_t>& he = hePer[t];
    for (int64_t j = 0; j < ne; ++j) {
      const int32_t e = te[j].e;
      mp.tileEdgeSrc[e0 + j] = e;
      mp.lr[e0 + j] = te[j].l | (te[j].r << 16);
      const bool primary = mp.tileOfNode[mp.edgeNodes[2 * e]] == t;
      mp.tileEdgePrimary[e0 + j] = primary ? 1 : 0;
      if (primary)
        mp.primarySlotOfEdge[e] = (int32_t)(e0 + j);
      else
        mp.secondSlotOfEdge[e] = (int32_t)(e0 + j);
      for (int s = 0; s < 2; ++s) {
        const int32_t n = mp.edgeNodes[2 * e + s];
        if (mp.tileOfNode[n] == t)
          he.push_back(he_pack(
            (uint32_t)j, (uint32_t)s, 0u,
            uint32_t(mp.slotOfNode[n] - h.node0), false));
      }
    }
    std::sort(he.begin(), he.end(), [](uint32_t a, uint32_t b) {
      if (he_ent(a) != he_ent(b))
        return he_ent(a) < he_ent(b);
      return he_edge(a) < he_edge(b);
    });
    h.nHalfNode = (int32_t)he.size();
  }
  if (!err.empty())
    fail(err);

  tm.mark("mesh: per-tile lists");
  /* flatten halo + half-edge lists */
  mp.warpSplitNode.assign(size_t(nTiles) * (kMaxWarps + 1), 0);
  int64_t hp = 0, qp = 0;
  for (int64_t t = 0; t < nTiles; ++t) {
    TileHdr& h = mp.tiles[t];
    h.haloPtr = (int32_t)hp;
    h.hePtrNode = (int32_t)qp;
    h.warpPtrNode = (int32_t)(t * (kMaxWarps + 1));
    hp += (int64_t)haloPer[t].size();
    qp += (int64_t)hePer[t].size();
    /* keep every tile's half-edge list 4-aligned (16-byte vector loads) */
    qp = (qp + 3) & ~int64_t(3);
    mp.maxTileStaged =
      std::max<int64_t>(mp.maxTileStaged, h.nOwnPad + h.nHalo);
    mp.maxTileHalf = std::max<int64_t>(mp.maxTileHalf, h.nHalfNode);
  }
  if (hp >= (int64_t(1) << 31) || qp >= (int64_t(1) << 31))
    fail("nw_mesh_create: plan arrays exceed 32-bit offsets");
  mp.totalHalo = hp;
  mp.haloNodes.assign(hp + 4, 0);
  mp.haloBlock.assign(size_t(nTiles) * kHaloBlock, -1);
  mp.heNode.assign(qp + 4, 0u);
  for (int64_t t = 0; t < nTiles; ++t) {
    const TileHdr& h = mp.tiles[t];
    std::copy(haloPer[t].begin(), haloPer[t].end(),
              mp.haloNodes.begin() + h.haloPtr);
    std::copy(
      haloPer[t].begin(),
      haloPer[t].begin() + std::min<size_t>(haloPer[t].size(), kHaloBlock),
      mp.haloBlock.begin() + size_t(t) * kHaloBlock);
    std::copy(hePer[t].begin(), hePer[t].end(), mp.heNode.begin() + h.hePtrNode);
    split_half_edges(
      mp.heNode.data() + h.hePtrNode, h.nHalfNode, kMaxWarps,
      mp.warpSplitNode.data() + h.warpPtrNode);
  }
  for (int64_t t = 0; t < nTiles; ++t) {
    TileHdr& h = mp.tiles[t];
    h.ellPtrNode = (int32_t)mp.heNodeEll.size();
    h.slicePtrNode = (int32_t)mp.sliceOffNode.size();
    if (mp.heNodeEll.size() >= (size_t(1) << 31) - 65536)
      fail("nw_mesh_create: plan arrays exceed 32-bit offsets");
    h.ellLenNode = (int32_t)append_sliced_ell(
      mp.heNode.data() + h.hePtrNode, h.nHalfNode, h.nOwn, mp.heNodeEll,
      mp.sliceOffNode);
    mp.maxTileEllNode = std::max<int64_t>(mp.maxTileEllNode, h.ellLenNode);
  }
  mp.heNodeEll.resize(mp.heNodeEll.size() + 32, 0u);
  mp.sliceOffNode.push_back(0);
  tm.mark("mesh: flatten + sliced ELL");
}

/* ------------------------------------------------------------------ */
/*  graph                                                              */
/* ------------------------------------------------------------------ */

int64_t
Graph::localRow(int64_t hid) const
{
  if (hid >= iLower && hid <= iUpper)
    return hid - iLower;
  auto it =
    std::lower_bound(rowIndicesShared.begin(), rowIndicesShared.end(), hid);
  if (it == rowIndicesShared.end() || *it != hid)
    return -1;
  return numRowsOwned + (it - rowIndicesShared.begin());
}

void
build_graph(
  const MeshPlan& mp,
  int kind,
  int numDof,
  const std::vector<int64_t>& skippedIn,
  Graph& g)
{
  apply_host_threads();
  StageTimer tm;
  g = Graph();
  if (kind == NW_LINSYS_HYPRE_UVW)
    numDof = 1; /* HypreUVWLinearSystem builds its base with numDof = 1
                   (src/HypreUVWLinearSystem.C:15-32) */
  if (numDof < 1 || numDof > 3)
    fail("nw_linsys_create: num_dof must be 1..3");
  g.kind = kind;
  g.numDof = numDof;
  g.ndim = mp.ndim;
  g.iLower = mp.iLowerNode * numDof;
  g.iUpper = (mp.iUpperNode + 1) * numDof - 1;
  g.block = 2 * numDof;
  g.skippedRows = skippedIn;
  std::sort(g.skippedRows.begin(), g.skippedRows.end());
  g.skippedRows.erase(
    std::unique(g.skippedRows.begin(), g.skippedRows.end()),
    g.skippedRows.end());
  auto isSkipped = [&](int64_t row) {
    return std::binary_search(g.skippedRows.begin(), g.skippedRows.end(), row);
  };

  const int64_t E = mp.nEdges;
  const int64_t nOwnedNodes = mp.iUpperNode - mp.iLowerNode + 1;
  const int64_t* hid = mp.nodeHid.data();

  /* node-level adjacency.  Owned node rows are indexed directly; non-owned
   * (shared) node ids are collected, sorted and uniqued -- the order of the
   * reference's std::map<HypreIntType,...> (src/HypreLinearSystem.C:1152). */
  std::vector<int64_t> sharedIds;
  for (int64_t e = 0; e < 2 * E; ++e) {
    const int64_t h = hid[mp.edgeNodes[e]];
    if (h < mp.iLowerNode || h > mp.iUpperNode)
      sharedIds.push_back(h);
  }
  std::sort(sharedIds.begin(), sharedIds.end());
  sharedIds.erase(
    std::unique(sharedIds.begin(), sharedIds.end()), sharedIds.end());
  const int64_t nSharedNodes = (int64_t)sharedIds.size();
  auto nodeRowU = [&](int64_t h) -> int64_t {
    if (h >= mp.iLowerNode && h <= mp.iUpperNode)
      return h - mp.iLowerNode;
    return nOwnedNodes + (std::lower_bound(
                            sharedIds.begin(), sharedIds.end(), h) -
                          sharedIds.begin());
  };
  const int64_t U = nOwnedNodes + nSharedNodes;
  std::vector<int64_t> ptr(U + 1, 0);
  std::vector<int64_t> uL(E), uR(E);
  for (int64_t e = 0; e < E; ++e) {
    uL[e] = nodeRowU(hid[mp.edgeNodes[2 * e]]);
    uR[e] = nodeRowU(hid[mp.edgeNodes[2 * e + 1]]);
    ptr[uL[e] + 1] += 2;
    ptr[uR[e] + 1] += 2;
  }
  for (int64_t u = 0; u < U; ++u)
    ptr[u + 1] += ptr[u];
  std::vector<int64_t> adj(ptr[U]);
  {
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < E; ++e) {
      const int64_t hL = hid[mp.edgeNodes[2 * e]];
      const int64_t hR = hid[mp.edgeNodes[2 * e + 1]];
      adj[fill[uL[e]]++] = hL;
      adj[fill[uL[e]]++] = hR;
      adj[fill[uR[e]]++] = hL;
      adj[fill[uR[e]]++] = hR;
    }
  }
  std::vector<int64_t> ucount(U, 0);
#pragma omp parallel for schedule(static)
  for (int64_t u = 0; u < U; ++u) {
    int64_t* b = adj.data() + ptr[u];
    int64_t* e = adj.data() + ptr[u + 1];
    std::sort(b, e);
    ucount[u] = std::unique(b, e) - b;
  }

  tm.mark("graph: adjacency");
  /* ---- owned rows: src/HypreLinearSystem.C:999-1066 ---- */
  g.numRowsOwned = nOwnedNodes * numDof;
  g.rowStartOwned.assign(g.numRowsOwned + 1, 0);
  for (int64_t u = 0; u < nOwnedNodes; ++u)
    for (int d = 0; d < numDof; ++d) {
      const int64_t row = g.iLower + u * numDof + d;
      int64_t len;
      if (isSkipped(row))
        len = 1; /* Dirichlet row: diagonal only */
      else if (ucount[u] == 0) {
        len = 1; /* untouched row: periodic slave */
        g.periodicRowsOwned.push_back(row);
      } else
        len = ucount[u] * numDof;
      g.rowStartOwned[u * numDof + d + 1] = len;
    }
  for (int64_t r = 0; r < g.numRowsOwned; ++r)
    g.rowStartOwned[r + 1] += g.rowStartOwned[r];
  g.nnzOwned = g.rowStartOwned[g.numRowsOwned];

  /* ---- shared rows: :1143-1236 (skipped rows are dropped) ---- */
  std::vector<int64_t> sharedNodeOfRow; /* index into sharedIds per kept row */
  std::vector<int> sharedDofOfRow;
  for (int64_t s = 0; s < nSharedNodes; ++s)
    for (int d = 0; d < numDof; ++d) {
      const int64_t row = sharedIds[s] * numDof + d;
      if (isSkipped(row))
        continue;
      g.rowIndicesShared.push_back(row);
      sharedNodeOfRow.push_back(s);
      sharedDofOfRow.push_back(d);
    }
  g.numRowsShared = (int64_t)g.rowIndicesShared.size();
  g.rowStartShared.assign(g.numRowsShared + 1, 0);
  for (int64_t i = 0; i < g.numRowsShared; ++i)
    g.rowStartShared[i + 1] =
      g.rowStartShared[i] + ucount[nOwnedNodes + sharedNodeOfRow[i]] * numDof;
  g.nnzShared = g.rowStartShared[g.numRowsShared];

  /* ---- cols / rows arrays: :956-993 ---- */
  const int64_t nnz = g.nnzOwned + g.nnzShared;
  g.cols.assign(nnz, 0);
  g.rows.assign(nnz, 0);
  auto fillRow = [&](int64_t at, int64_t len, int64_t row, int64_t u) {
    if (len != ucount[u] * numDof) {
      /* diagonal-only row: Dirichlet (skipped) or untouched (periodic slave) */
      g.cols[at] = row;
      g.rows[at] = row;
      return;
    }
    const int64_t* a = adj.data() + ptr[u];
    int64_t k = at;
    for (int64_t c = 0; c < ucount[u]; ++c)
      for (int dd = 0; dd < numDof; ++dd) {
        g.cols[k] = a[c] * numDof + dd;
        g.rows[k] = row;
        ++k;
      }
  };
#pragma omp parallel for schedule(static)
  for (int64_t u = 0; u < nOwnedNodes; ++u)
    for (int d = 0; d < numDof; ++d) {
      const int64_t r = u * numDof + d;
      fillRow(
        g.rowStartOwned[r], g.rowStartOwned[r + 1] - g.rowStartOwned[r],
        g.iLower + r, u);
    }
  for (int64_t i = 0; i < g.numRowsShared; ++i)
    fillRow(
      g.nnzOwned + g.rowStartShared[i],
      g.rowStartShared[i + 1] - g.rowStartShared[i], g.rowIndicesShared[i],
      nOwnedNodes + sharedNodeOfRow[i]);

  tm.mark("graph: rows + cols");
  /* ---- edge -> slot map ---- */
  const int nb = (kind == NW_LINSYS_HYPRE_UVW) ? 2 : g.block;
  g.block = nb;
  g.edgeSlots.assign(size_t(E) * nb * nb, -1);
  g.edgeRhsRows.assign(size_t(E) * nb, -1);
  std::string err;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < E; ++e) {
    const int64_t hL = hid[mp.edgeNodes[2 * e]];
    const int64_t hR = hid[mp.edgeNodes[2 * e + 1]];
    if (hL == hR) {
#pragma omp critical
      err = "edge connects two nodes that resolve to the same row";
      continue;
    }
    const int64_t hh[2] = {hL, hR};
    for (int i = 0; i < 2; ++i) {
      /* sum_into: the skip test uses the node's first row id (:2095-2099) */
      const int64_t first = hh[i] * numDof;
      if (isSkipped(first))
        continue;
      for (int d = 0; d < numDof; ++d) {
        const int64_t row = first + d;
        if (numDof > 1 && isSkipped(row)) {
#pragma omp critical
          err = "num_dof > 1: skipped rows must cover all dofs of a node";
          continue;
        }
        const int64_t lrow = g.localRow(row);
        if (lrow < 0)
          continue; /* not in map_shared_: silently skipped (:2137) */
        const int64_t base = g.rowPtr(lrow), len = g.rowLen(lrow);
        const int ii = i * numDof + d;
        g.edgeRhsRows[size_t(e) * nb + ii] = lrow;
        const int64_t* rc = g.cols.data() + base;
        for (int k = 0; k < 2; ++k)
          for (int dd = 0; dd < numDof; ++dd) {
            const int64_t col = hh[k] * numDof + dd;
            const int64_t pos = std::lower_bound(rc, rc + len, col) - rc;
            if (pos >= len || rc[pos] != col) {
#pragma omp critical
              err = "internal: column missing from graph row";
              continue;
            }
            g.edgeSlots[(size_t(e) * nb + ii) * nb + (k * numDof + dd)] =
              base + pos;
          }
      }
    }
  }
  if (!err.empty())
    fail("nw_linsys_finalize: " + err);
  tm.mark("graph: edge slots");
}

/* ------------------------------------------------------------------ */
/*  per-tile linear-system plan                                        */
/* ------------------------------------------------------------------ */

void
build_ls_plan(const MeshPlan& mp, const Graph& g, LsPlan& lp)
{
  apply_host_threads();
  StageTimer tm;
  lp = LsPlan();
  if (g.numDof != 1) {
    lp.usable = false;
    lp.whyNot = "tile path implemented for 1-dof graphs only";
    return;
  }
  const int64_t nTiles = mp.nTiles;
  const int64_t N = mp.nNodes;
  std::vector<int64_t> nodeRow(N);
  auto isSkipped = [&](int64_t row) {
    return std::binary_search(g.skippedRows.begin(), g.skippedRows.end(), row);
  };
  /* a row nothing assembles into (no columns: a periodic slave's own row, or a
   * node no edge touches) is one of the reference's periodic_bc_rows_owned_
   * (src/HypreLinearSystem.C:1036-1039): diagonal 1, rhs 0 after every reset
   * (:1420-1428).  No tile may own it -- it stays with row_init. */
  auto isDiagOnlyRow = [&](int64_t row) {
    return std::binary_search(
      g.periodicRowsOwned.begin(), g.periodicRowsOwned.end(), row);
  };
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const int64_t h = mp.nodeHid[n];
    nodeRow[n] = (isSkipped(h) || isDiagOnlyRow(h)) ? -1 : g.localRow(h);
  }

  lp.tiles.assign(nTiles, LsTileHdr());
  std::vector<std::vector<EntInfo>> entPer(nTiles);
  std::vector<std::vector<int32_t>> rowPer(nTiles);
  std::vector<std::vector<uint32_t>> hePer(nTiles);
  std::vector<std::vector<Run>> runPer(nTiles);
  std::string err;

#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t t = 0; t < nTiles; ++t) {
    const TileHdr& h = mp.tiles[t];
    /* rows of the tile: distinct rows of its nodes, ascending */
    std::vector<int64_t> rows;
    for (int32_t s = h.node0; s < h.node0 + h.nOwn; ++s) {
      const int64_t r = nodeRow[mp.nodeOfSlot[s]];
      if (r >= 0)
        rows.push_back(r);
    }
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    if ((int64_t)rows.size() > kMaxTileEnts) {
#pragma omp critical
      err = "tile has too many rows";
      continue;
    }
    std::vector<EntInfo>& ents = entPer[t];
    std::vector<Run>& runs = runPer[t];
    int64_t so = 0;
    bool bad = false;
    for (size_t i = 0; i < rows.size(); ++i) {
      const int64_t r = rows[i];
      const int64_t len = g.rowLen(r), go = g.rowPtr(r);
      if (len > kMaxRowNnz || so + len > 65535) {
        bad = true;
        break;
      }
      const int64_t rowId =
        r < g.numRowsOwned ? g.iLower + r
                           : g.rowIndicesShared[r - g.numRowsOwned];
      const int64_t* rc = g.cols.data() + go;
      const int64_t dk = std::lower_bound(rc, rc + len, rowId) - rc;
      EntInfo ei;
      ei.base = (uint16_t)so;
      ei.diagK = (uint8_t)dk;
      ei.nnz = (uint8_t)len;
      ents.push_back(ei);
      rowPer[t].push_back((int32_t)r);
      if (!runs.empty() && runs.back().go + runs.back().len == go)
        runs.back().len += (int32_t)len;
      else
        runs.push_back(Run{go, (int32_t)so, (int32_t)len});
      so += len;
    }
    if (bad) {
#pragma omp critical
      err = "tile row staging exceeds limits (row too long or tile too big)";
      continue;
    }
    LsTileHdr& lh = lp.tiles[t];
    lh.nEnts = (int32_t)rows.size();
    lh.hasShared = 0;
    for (const auto r : rows)
      if ((int64_t)r >= g.numRowsOwned)
        lh.hasShared = 1;
    lh.nnz = (int32_t)so;
    lh.nRuns = (int32_t)runs.size();

    struct HE
    {
      uint32_t ent, k, j, side;
    };
    std::vector<HE> hes;
    for (int32_t j = 0; j < h.nEdges; ++j) {
      const int32_t e = mp.tileEdgeSrc[h.edge0 + j];
      for (uint32_t s = 0; s < 2; ++s) {
        const int32_t n = mp.edgeNodes[2 * e + s];
        if (mp.tileOfNode[n] != t || nodeRow[n] < 0)
          continue;
        const int32_t o = mp.edgeNodes[2 * e + 1 - s];
        const int64_t r = nodeRow[n];
        const uint32_t ent =
          uint32_t(std::lower_bound(rows.begin(), rows.end(), r) - rows.begin());
        const int64_t go = g.rowPtr(r), len = g.rowLen(r);
        const int64_t* rc = g.cols.data() + go;
        const int64_t col = mp.nodeHid[o];
        const int64_t k = std::lower_bound(rc, rc + len, col) - rc;
        if (k >= len || rc[k] != col || k == ents[ent].diagK) {
          bad = true;
          break;
        }
        hes.push_back(HE{ent, (uint32_t)k, (uint32_t)j, s});
      }
      if (bad)
        break;
    }
    if (bad) {
#pragma omp critical
      err = "edge column missing from row or edge joins two nodes of one row";
      continue;
    }
    std::sort(hes.begin(), hes.end(), [](const HE& a, const HE& b) {
      if (a.ent != b.ent)
        return a.ent < b.ent;
      if (a.k != b.k)
        return a.k < b.k;
      return a.j < b.j;
    });
    std::vector<uint32_t>& he = hePer[t];
    he.resize(hes.size());
    for (size_t i = 0; i < hes.size(); ++i) {
      /* only the later members of a (row, k) group are marked */
      const bool dup =
        i > 0 && hes[i - 1].ent == hes[i].ent && hes[i - 1].k == hes[i].k;
      he[i] = he_pack(hes[i].j, hes[i].side, hes[i].k, hes[i].ent, dup);
    }
    lh.nHalf = (int32_t)he.size();
  }
  if (!err.empty()) {
    lp.usable = false;
    lp.whyNot = err;
    return;
  }

  tm.mark("lsplan: per-tile lists");
  lp.warpSplit.assign(size_t(nTiles) * (kMaxWarps + 1), 0);
  int64_t ep = 0, hp = 0, rp = 0;
  for (int64_t t = 0; t < nTiles; ++t) {
    LsTileHdr& lh = lp.tiles[t];
    lh.entPtr = (int32_t)ep;
    lh.hePtr = (int32_t)hp;
    lh.runPtr = (int32_t)rp;
    lh.warpPtr = (int32_t)(t * (kMaxWarps + 1));
    /* 4-aligned: the per-tile runs of 4-byte records are bulk-copied */
    ep = (ep + lh.nEnts + 3) & ~int64_t(3);
    hp = (hp + lh.nHalf + 3) & ~int64_t(3);
    rp += lh.nRuns;
    lp.maxTileRuns = std::max<int64_t>(lp.maxTileRuns, lh.nRuns);
    lp.maxTileNnz = std::max<int64_t>(lp.maxTileNnz, lh.nnz);
    lp.maxTileEnts = std::max<int64_t>(lp.maxTileEnts, lh.nEnts);
    lp.maxTileHalf = std::max<int64_t>(lp.maxTileHalf, lh.nHalf);
  }
  if (hp >= (int64_t(1) << 31)) {
    lp.usable = false;
    lp.whyNot = "half-edge list exceeds 32-bit offsets";
    return;
  }
  lp.entInfo.resize(ep + 4);
  lp.entRhsRow.assign(ep + 4, -1);
  lp.entGo.assign(ep + 4, 0);
  lp.he.assign(hp + 4, 0u);
  lp.runs.resize(rp + 1);
  for (int64_t t = 0; t < nTiles; ++t) {
    const LsTileHdr& lh = lp.tiles[t];
    std::copy(entPer[t].begin(), entPer[t].end(), lp.entInfo.begin() + lh.entPtr);
    std::copy(rowPer[t].begin(), rowPer[t].end(),
              lp.entRhsRow.begin() + lh.entPtr);
    for (size_t i = 0; i < rowPer[t].size(); ++i)
      lp.entGo[lh.entPtr + i] = (int32_t)g.rowPtr(rowPer[t][i]);
    std::copy(hePer[t].begin(), hePer[t].end(), lp.he.begin() + lh.hePtr);
    std::copy(runPer[t].begin(), runPer[t].end(), lp.runs.begin() + lh.runPtr);
    split_half_edges(
      lp.he.data() + lh.hePtr, lh.nHalf, kMaxWarps,
      lp.warpSplit.data() + lh.warpPtr);
  }
  for (int64_t t = 0; t < nTiles; ++t) {
    LsTileHdr& lh = lp.tiles[t];
    lh.ellPtr = (int32_t)lp.heEll.size();
    lh.slicePtr = (int32_t)lp.sliceOff.size();
    if (lp.heEll.size() >= (size_t(1) << 31) - 65536) {
      lp.usable = false;
      lp.whyNot = "half-edge list exceeds 32-bit offsets";
      return;
    }
    lh.ellLen = (int32_t)append_sliced_ell(
      lp.he.data() + lh.hePtr, lh.nHalf, lh.nEnts, lp.heEll, lp.sliceOff);
    lp.maxTileEll = std::max<int64_t>(lp.maxTileEll, lh.ellLen);
  }
  lp.heEll.resize(lp.heEll.size() + 32, 0u);
  lp.sliceOff.push_back(0);
  std::vector<uint8_t> covered(g.numRowsLocal(), 0);
  for (int64_t i = 0; i < ep; ++i)
    if (lp.entRhsRow[i] >= 0)
      covered[lp.entRhsRow[i]] = 1;
  for (int64_t r = 0; r < g.numRowsLocal(); ++r)
    if (!covered[r])
      lp.uncoveredRows.push_back((int32_t)r);
  tm.mark("lsplan: flatten + sliced ELL");
}

} // namespace nw
