/*
 * nw_emul.cpp -- CPU walk-through of the tile plans (TEST INFRASTRUCTURE ONLY;
 * lives under tests/, is never linked into the product library).
 *
 * The CUDA tile kernels cannot run in the build container (no GPU).  This file
 * replays their phases sequentially on the host -- same plan arrays, same
 * record decoding, same physics header (csrc/edge_physics.h) -- so that the
 * CPU test-suite can check the plan builder and the restated arithmetic against
 * the oracle before any GPU time is spent.  It compiles plan.cpp directly.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <map>

#include "edge_physics.h"
#include "geometry_cvfem.h"
#include "plan.h"

using namespace nw;

namespace {

/* visit the half-edges of one tile in the device's order: entity by entity,
 * each entity's records from its sliced-ELL column (w ascending) */
template <class F>
bool
walk_ell(
  const uint32_t* ell, const int32_t* sliceOff, int nEnts, int ellLen, F&& f)
{
  const int nSlices = (nEnts + 31) / 32;
  if (sliceOff[0] != 0 || sliceOff[nSlices] != ellLen)
    return false;
  for (int row = 0; row < nEnts; ++row) {
    const int sl = row >> 5;
    const int o0 = sliceOff[sl], o1 = sliceOff[sl + 1];
    if (o1 < o0 || ((o1 - o0) & 31))
      return false;
    const int W = (o1 - o0) >> 5;
    bool ended = false;
    for (int w = 0; w < W; ++w) {
      const uint32_t hv = ell[o0 + w * 32 + (row & 31)];
      if (!(hv & kHeValid)) {
        ended = true;
        continue;
      }
      if (ended || (int)he_ent(hv) != row)
        return false;
      f(row, hv);
    }
  }
  return true;
}

struct Emu
{
  MeshPlan mp;
  Graph g;
  LsPlan lp;
  bool hasLs = false;
  std::string err;
};

/* SoA internal storage of a node field given in caller AoS layout */
std::vector<double>
to_slots(const MeshPlan& mp, const double* aos, int ncomp)
{
  std::vector<double> s(size_t(mp.nSlots) * ncomp, 0.0);
  for (int64_t i = 0; i < mp.nSlots; ++i) {
    const int32_t n = mp.nodeOfSlot[i];
    if (n < 0)
      continue;
    for (int c = 0; c < ncomp; ++c)
      s[size_t(c) * mp.nSlots + i] = aos[size_t(n) * ncomp + c];
  }
  return s;
}

std::vector<double>
edge_to_slots(const MeshPlan& mp, const double* aos, int ncomp)
{
  std::vector<double> s(size_t(mp.nTileEdgeSlots) * ncomp, 0.0);
  for (int64_t i = 0; i < mp.nTileEdgeSlots; ++i) {
    const int32_t e = mp.tileEdgeSrc[i];
    if (e < 0)
      continue;
    for (int c = 0; c < ncomp; ++c)
      s[size_t(c) * mp.nTileEdgeSlots + i] = aos[size_t(e) * ncomp + c];
  }
  return s;
}

/* stage one tile: returns staged[c*stride + local] */
struct Staged
{
  std::vector<double> s;
  int stride;
  double operator()(int c, int i) const { return s[size_t(c) * stride + i]; }
};

Staged
stage(const MeshPlan& mp, const TileHdr& h, const std::vector<const double*>& comps)
{
  Staged st;
  st.stride = (h.nOwnPad + h.nHalo + 1) & ~1;
  st.s.assign(size_t(st.stride) * comps.size(), 0.0);
  for (size_t c = 0; c < comps.size(); ++c) {
    for (int i = 0; i < h.nOwnPad; ++i)
      st.s[c * st.stride + i] = comps[c][h.node0 + i];
    for (int k = 0; k < h.nHalo; ++k)
      st.s[c * st.stride + h.nOwnPad + k] =
        comps[c][mp.haloNodes[h.haloPtr + k]];
  }
  return st;
}

} // namespace

/* GeometryInteriorAlg<Tet4 / Wed6 / Pyr5> with the product's per-element
 * arithmetic (csrc/geometry_cvfem.h) replayed element by element: the host
 * side of nw_geometry_interior_* (edge look-up by the surface's node pair,
 * sign from the edge's first node) + the body of geometry_cvfem_kernel.
 * topo: 0 tet4, 1 wed6, 2 pyr5; accumulates into dnv [n_nodes], area [n_edges][3] */
template <int T>
static void
emu_geo_block(
  int64_t nElems, const int32_t* elemNodes, const double* coords, int64_t nEdges,
  const int32_t* edgeNodes, double* dnv, double* area)
{
  constexpr int npe = geo::Traits<T>::npe;
  std::map<std::pair<int32_t, int32_t>, int64_t> edgeOf;
  for (int64_t e = 0; e < nEdges; ++e) {
    const int32_t a = edgeNodes[2 * e], b = edgeNodes[2 * e + 1];
    edgeOf[{std::min(a, b), std::max(a, b)}] = e;
  }
  for (int64_t el = 0; el < nElems; ++el) {
    const int32_t* en = elemNodes + (int64_t)npe * el;
    double c[npe][3], v[geo::Traits<T>::nSub][3];
    for (int n = 0; n < npe; ++n)
      for (int d = 0; d < 3; ++d)
        c[n][d] = coords[(int64_t)en[n] * 3 + d];
    geo::sub_points<T>(c, v);
    if (dnv)
      for (int ip = 0; ip < geo::Traits<T>::nScv; ++ip)
        dnv[en[ip]] += geo::scv_volume<T>(ip, v);
    if (!area)
      continue;
    for (int ip = 0; ip < geo::Traits<T>::nScs; ++ip) {
      int l, r;
      geo::scs_nodes<T>(ip, &l, &r);
      auto it = edgeOf.find({std::min(en[l], en[r]), std::max(en[l], en[r])});
      if (it == edgeOf.end())
        continue;
      double a[3];
      geo::scs_area<T>(ip, v, a);
      const double sg = en[l] == edgeNodes[2 * it->second] ? 1.0 : -1.0;
      for (int d = 0; d < 3; ++d)
        area[it->second * 3 + d] += a[d] * sg;
    }
  }
}

extern "C" {

void*
emu_create(
  int ndim,
  int rank,
  int nranks,
  int64_t nNodes,
  int64_t nEdges,
  const int32_t* edgeNodes,
  const int64_t* hid,
  const int64_t* ownHid,
  const int64_t* offsets,
  const double* coords,
  int tileNodes)
{
  Emu* e = new Emu;
  try {
    MeshInput in;
    in.ndim = ndim;
    in.rank = rank;
    in.nranks = nranks;
    in.nNodes = nNodes;
    in.nEdges = nEdges;
    in.edgeNodes = edgeNodes;
    in.nodeHid = hid;
    in.nodeOwnHid = ownHid;
    in.hypreOffsets = offsets;
    in.coords = coords;
    in.tileNodes = tileNodes;
    build_mesh_plan(in, e->mp);
  } catch (const std::exception& ex) {
    e->err = ex.what();
  }
  return e;
}

const char*
emu_error(void* h)
{
  return static_cast<Emu*>(h)->err.c_str();
}

void
emu_destroy(void* h)
{
  delete static_cast<Emu*>(h);
}

int
emu_build_linsys(
  void* h, int kind, int numDof, const int64_t* skipped, int64_t nSkipped)
{
  Emu* e = static_cast<Emu*>(h);
  try {
    std::vector<int64_t> sk(skipped, skipped + nSkipped);
    build_graph(e->mp, kind, numDof, sk, e->g);
    build_ls_plan(e->mp, e->g, e->lp);
    e->hasLs = true;
  } catch (const std::exception& ex) {
    e->err = ex.what();
    return 1;
  }
  if (!e->lp.usable) {
    e->err = "tile plan unusable: " + e->lp.whyNot;
    return 2;
  }
  return 0;
}

/* plan invariants; returns 0 or writes a message */
int
emu_check_plan(void* h)
{
  Emu* e = static_cast<Emu*>(h);
  const MeshPlan& mp = e->mp;
  auto bad = [&](const std::string& m) {
    e->err = m;
    return 1;
  };
  std::vector<int> seen(mp.nNodes, 0);
  for (int64_t t = 0; t < mp.nTiles; ++t) {
    const TileHdr& hd = mp.tiles[t];
    if (hd.node0 & 1)
      return bad("tile node0 not even");
    if (hd.edge0 & 1)
      return bad("tile edge0 not even");
    for (int i = 0; i < hd.nOwn; ++i) {
      const int32_t n = mp.nodeOfSlot[hd.node0 + i];
      if (n < 0 || mp.tileOfNode[n] != t)
        return bad("slot/tile mismatch");
      seen[n]++;
    }
    for (int k = 1; k < hd.nHalo; ++k)
      if (mp.haloNodes[hd.haloPtr + k] <= mp.haloNodes[hd.haloPtr + k - 1])
        return bad("halo list not strictly ascending");
    /* every tile-edge decodes back to its source edge */
    for (int j = 0; j < hd.nEdges; ++j) {
      const int32_t src = mp.tileEdgeSrc[hd.edge0 + j];
      const uint32_t v = mp.lr[hd.edge0 + j];
      const int l = v & 0xffff, r = v >> 16;
      auto gslot = [&](int loc) {
        return loc < hd.nOwnPad ? hd.node0 + loc
                                : mp.haloNodes[hd.haloPtr + loc - hd.nOwnPad];
      };
      if (mp.nodeOfSlot[gslot(l)] != mp.edgeNodes[2 * src] ||
          mp.nodeOfSlot[gslot(r)] != mp.edgeNodes[2 * src + 1])
        return bad("lr record does not decode to the edge's nodes");
    }
    /* half-edges sorted by entity, warp split at entity boundaries */
    const uint32_t* he = mp.heNode.data() + hd.hePtrNode;
    for (int q = 1; q < hd.nHalfNode; ++q)
      if (he_ent(he[q]) < he_ent(he[q - 1]))
        return bad("node half-edges not sorted");
    const int32_t* sp = mp.warpSplitNode.data() + hd.warpPtrNode;
    if (sp[0] != 0 || sp[kMaxWarps] != hd.nHalfNode)
      return bad("warp split endpoints");
    for (int w = 1; w < kMaxWarps; ++w) {
      if (sp[w] < sp[w - 1])
        return bad("warp split not monotone");
      if (sp[w] > 0 && sp[w] < hd.nHalfNode &&
          he_ent(he[sp[w]]) == he_ent(he[sp[w] - 1]))
        return bad("warp split inside an entity");
    }
    /* the sliced-ELL form must replay the flat list exactly */
    if ((hd.ellPtrNode & 31) || (hd.ellLenNode & 31) || (hd.edge0 & 3))
      return bad("node ELL block / tile-edge run not aligned");
    int q = 0;
    bool same = true;
    if (!walk_ell(
          mp.heNodeEll.data() + hd.ellPtrNode,
          mp.sliceOffNode.data() + hd.slicePtrNode, hd.nOwn, hd.ellLenNode,
          [&](int, uint32_t hv) {
            same = same && q < hd.nHalfNode && he[q] == hv;
            ++q;
          }) ||
        !same || q != hd.nHalfNode)
      return bad("node ELL list differs from the flat list");
  }
  for (int64_t n = 0; n < mp.nNodes; ++n)
    if (seen[n] != 1)
      return bad("node not owned by exactly one tile");
  for (int64_t ed = 0; ed < mp.nEdges; ++ed) {
    const int32_t p = mp.primarySlotOfEdge[ed];
    if (p < 0 || mp.tileEdgeSrc[p] != ed || !mp.tileEdgePrimary[p])
      return bad("primary slot map broken");
    const int32_t tL = mp.tileOfNode[mp.edgeNodes[2 * ed]];
    const int32_t tR = mp.tileOfNode[mp.edgeNodes[2 * ed + 1]];
    if ((tL != tR) != (mp.secondSlotOfEdge[ed] >= 0))
      return bad("cut edge must have exactly two copies");
  }
  if (e->hasLs) {
    const LsPlan& lp = e->lp;
    const Graph& g = e->g;
    std::vector<int> cov(g.numRowsLocal(), 0);
    for (int64_t t = 0; t < mp.nTiles; ++t) {
      const LsTileHdr& lh = lp.tiles[t];
      int64_t so = 0;
      for (int i = 0; i < lh.nEnts; ++i) {
        const EntInfo& ei = lp.entInfo[lh.entPtr + i];
        const int64_t r = lp.entRhsRow[lh.entPtr + i];
        cov[r]++;
        if (ei.base != so || ei.nnz != g.rowLen(r))
          return bad("ent staging layout");
        so += ei.nnz;
      }
      if (so != lh.nnz)
        return bad("tile nnz");
      int64_t covered = 0;
      for (int q = 0; q < lh.nRuns; ++q) {
        const Run& rn = lp.runs[lh.runPtr + q];
        if (rn.so != covered)
          return bad("runs not contiguous in staging");
        covered += rn.len;
      }
      if (covered != lh.nnz)
        return bad("runs do not cover the staging");
      const uint32_t* he = lp.he.data() + lh.hePtr;
      for (int q = 1; q < lh.nHalf; ++q)
        if (he_ent(he[q]) < he_ent(he[q - 1]))
          return bad("row half-edges not sorted");
      const int32_t* sp = lp.warpSplit.data() + lh.warpPtr;
      for (int w = 1; w < kMaxWarps; ++w)
        if (sp[w] > 0 && sp[w] < lh.nHalf &&
            he_ent(he[sp[w]]) == he_ent(he[sp[w] - 1]))
          return bad("ls warp split inside a row");
      if ((lh.ellPtr & 31) || (lh.ellLen & 31) || (lh.entPtr & 3))
        return bad("row ELL block / ent run not aligned");
      int q = 0;
      bool same = true;
      if (!walk_ell(
            lp.heEll.data() + lh.ellPtr, lp.sliceOff.data() + lh.slicePtr,
            lh.nEnts, lh.ellLen,
            [&](int, uint32_t hv) {
              same = same && q < lh.nHalf && he[q] == hv;
              ++q;
            }) ||
          !same || q != lh.nHalf)
        return bad("row ELL list differs from the flat list");
    }
    for (int32_t r : lp.uncoveredRows)
      cov[r]++;
    for (int64_t r = 0; r < g.numRowsLocal(); ++r)
      if (cov[r] != 1)
        return bad("row not covered exactly once");
  }
  return 0;
}

/* ---- emulated kernels: all fields in caller AoS layout ---- */

/* kind: 0 continuity, 1 scalar, 2 momentum-UVW.  values/rhs sized like the
 * product's arrays; emulates a tile assembly on a lazily zeroed system. */
int
emu_assemble(
  void* h,
  int kind,
  const double* const* nodeFields, /* per policy, AoS */
  const int* nodeNcomp,
  int nNodeFields,
  const double* area,
  const double* mdot,
  const double* pecfac,
  const void* opts,
  double* values,
  double* rhs)
{
  Emu* e = static_cast<Emu*>(h);
  if (!e->hasLs || !e->lp.usable) {
    e->err = "no usable linear-system plan";
    return 1;
  }
  const MeshPlan& mp = e->mp;
  const Graph& g = e->g;
  const LsPlan& lp = e->lp;
  constexpr int ND = 3;
  if (mp.ndim != 3) {
    e->err = "emulator handles ndim == 3";
    return 1;
  }
  std::vector<std::vector<double>> store;
  std::vector<const double*> comps;
  for (int f = 0; f < nNodeFields; ++f) {
    store.push_back(to_slots(mp, nodeFields[f], nodeNcomp[f]));
  }
  for (int f = 0; f < nNodeFields; ++f)
    for (int c = 0; c < nodeNcomp[f]; ++c)
      comps.push_back(store[f].data() + size_t(c) * mp.nSlots);
  std::vector<double> sArea = edge_to_slots(mp, area, ND);
  std::vector<double> sMdot, sPec;
  if (mdot)
    sMdot = edge_to_slots(mp, mdot, 1);
  if (pecfac)
    sPec = edge_to_slots(mp, pecfac, 1);
  const int NR = kind == 2 ? ND : 1;
  const int NRES = kind == 0 ? 2 : (kind == 1 ? 5 : 4 + ND);
  const int64_t rows = g.numRowsLocal();
  const int64_t S = mp.nTileEdgeSlots;

  /* lazy zero semantics: nothing is pre-zeroed; poison to catch gaps */
  const int64_t nnz = g.nnzOwned + g.nnzShared;
  for (int64_t i = 0; i < nnz; ++i)
    values[i] = NAN;
  for (int64_t i = 0; i < rows * NR; ++i)
    rhs[i] = NAN;

  for (int64_t t = 0; t < mp.nTiles; ++t) {
    const TileHdr& hd = mp.tiles[t];
    const LsTileHdr& lh = lp.tiles[t];
    Staged st = stage(mp, hd, comps);
    std::vector<double> res(size_t(NRES) * hd.nEdges);
    /* phase 1 */
    for (int j = 0; j < hd.nEdges; ++j) {
      const uint32_t v = mp.lr[hd.edge0 + j];
      const int l = v & 0xffff, r = v >> 16;
      double av[ND];
      for (int d = 0; d < ND; ++d)
        av[d] = sArea[size_t(d) * S + hd.edge0 + j];
      double out[8];
      if (kind == 0) {
        ContNode<ND> L, R;
        auto ld = [&](int i, ContNode<ND>& n) {
          for (int d = 0; d < ND; ++d) {
            n.x[d] = st(d, i);
            n.u[d] = st(ND + d, i);
            n.g[d] = st(2 * ND + d, i);
          }
          n.rho = st(3 * ND, i);
          n.p = st(3 * ND + 1, i);
          n.ud = st(3 * ND + 2, i);
        };
        ld(l, L);
        ld(r, R);
        continuity_edge<ND>(
          L, R, av, *static_cast<const nw_continuity_opts*>(opts), out[0],
          out[1]);
      } else if (kind == 1) {
        ScalNode<ND> L, R;
        auto ld = [&](int i, ScalNode<ND>& n) {
          for (int d = 0; d < ND; ++d) {
            n.x[d] = st(d, i);
            n.v[d] = st(ND + d, i);
            n.dq[d] = st(2 * ND + d, i);
          }
          n.q = st(3 * ND, i);
          n.rho = st(3 * ND + 1, i);
          n.mu = st(3 * ND + 2, i);
        };
        ld(l, L);
        ld(r, R);
        scalar_edge<ND>(
          L, R, av, sMdot[hd.edge0 + j],
          *static_cast<const nw_scalar_opts*>(opts), out, out[4]);
      } else {
        const nw_momentum_opts& o = *static_cast<const nw_momentum_opts*>(opts);
        MomNode<ND> L, R;
        auto ld = [&](int i, MomNode<ND>& n) {
          for (int d = 0; d < ND; ++d) {
            n.x[d] = st(d, i);
            n.u[d] = st(ND + d, i);
          }
          for (int d = 0; d < ND * ND; ++d)
            n.g[d] = st(2 * ND + d, i);
          n.mu = st(2 * ND + ND * ND, i);
          n.rho = st(2 * ND + ND * ND + 1, i);
          n.mask = st(2 * ND + ND * ND + 2, i);
        };
        ld(l, L);
        ld(r, R);
        double pf = pecfac ? sPec[hd.edge0 + j] : 0.0;
        if (o.fuse_peclet) {
          PecNode<ND> pl, pr;
          for (int d = 0; d < ND; ++d) {
            pl.x[d] = L.x[d];
            pr.x[d] = R.x[d];
            pl.v[d] = L.u[d];
            pr.v[d] = R.u[d];
          }
          pl.rho = L.rho;
          pr.rho = R.rho;
          pl.mu = L.mu;
          pr.mu = R.mu;
          pf = peclet_eval(o.pf, peclet_number<ND>(pl, pr, o.pec_eps));
        }
        MomResult<ND> m;
        /* has_vof: the caller hands mass_flow_rate + mass_vof_balanced_flow_rate
         * as the mdot stream, as nw_assemble_momentum_edge's edge_sum_kernel does */
        if (o.has_vof)
          momentum_edge_vof<ND>(L, R, av, sMdot[hd.edge0 + j], pf, o, m);
        else
          momentum_edge<ND>(L, R, av, sMdot[hd.edge0 + j], pf, o, m);
        momentum_block_entry<ND>(
          m, av, o.relax_fac, 0, 0, out[0], out[1], out[2], out[3]);
        for (int d = 0; d < ND; ++d)
          out[4 + d] = m.flux[d];
      }
      for (int k = 0; k < NRES; ++k)
        res[size_t(k) * hd.nEdges + j] = out[k];
    }
    /* phase 2: one row at a time through the sliced-ELL list, exactly the
     * device's order; the staging is poisoned so an unwritten slot shows */
    std::vector<double> sVals(lh.nnz, NAN), sRhs(size_t(NR) * lh.nEnts, NAN);
    std::vector<double> diagAcc(lh.nEnts, 0.0), rhsAcc(size_t(NR) * lh.nEnts, 0.0);
    bool slotBad = false;
    const bool okWalk = walk_ell(
      lp.heEll.data() + lh.ellPtr, lp.sliceOff.data() + lh.slicePtr, lh.nEnts,
      lh.ellLen, [&](int ent, uint32_t hv) {
        const int j = he_edge(hv), side = he_side(hv);
        const EntInfo& ei = lp.entInfo[lh.entPtr + ent];
        double r_[8];
        for (int k = 0; k < NRES; ++k)
          r_[k] = res[size_t(k) * hd.nEdges + j];
        double diag, off, rr[3];
        if (kind == 0) {
          diag = -r_[0];
          off = r_[0];
          rr[0] = side ? r_[1] : -r_[1];
        } else {
          diag = side ? r_[3] : r_[0];
          off = side ? r_[2] : r_[1];
          for (int d = 0; d < NR; ++d)
            rr[d] = side ? r_[4 + d] : -r_[4 + d];
        }
        if (he_k(hv) >= ei.nnz || he_k(hv) == ei.diagK) {
          slotBad = true;
          return;
        }
        double& dst = sVals[ei.base + he_k(hv)];
        if (hv & kHeDup)
          off += dst;
        dst = off;
        diagAcc[ent] += diag;
        for (int d = 0; d < NR; ++d)
          rhsAcc[size_t(d) * lh.nEnts + ent] += rr[d];
      });
    if (!okWalk || slotBad) {
      e->err = "row ELL list malformed or half-edge slot out of row";
      return 1;
    }
    for (int i = 0; i < lh.nEnts; ++i) {
      const EntInfo& ei = lp.entInfo[lh.entPtr + i];
      sVals[ei.base + ei.diagK] = diagAcc[i];
      for (int d = 0; d < NR; ++d)
        sRhs[size_t(d) * lh.nEnts + i] = rhsAcc[size_t(d) * lh.nEnts + i];
    }
    /* phase 3: element-wise copy-out through the per-row value offsets */
    for (int i = 0; i < lh.nEnts; ++i) {
      const EntInfo& ei = lp.entInfo[lh.entPtr + i];
      const int64_t go = lp.entGo[lh.entPtr + i];
      if (go != g.rowPtr(lp.entRhsRow[lh.entPtr + i])) {
        e->err = "entGo does not match the row's value offset";
        return 1;
      }
      for (int k = 0; k < ei.nnz; ++k)
        values[go + k] = sVals[ei.base + k];
      for (int d = 0; d < NR; ++d)
        rhs[size_t(d) * rows + lp.entRhsRow[lh.entPtr + i]] =
          sRhs[size_t(d) * lh.nEnts + i];
    }
  }
  /* row init of the rows no tile writes */
  for (int32_t r : lp.uncoveredRows) {
    const int64_t a = g.rowPtr(r), len = g.rowLen(r);
    for (int64_t k = 0; k < len; ++k)
      values[a + k] = 0.0;
    if (r < g.numRowsOwned &&
        std::binary_search(
          g.periodicRowsOwned.begin(), g.periodicRowsOwned.end(),
          g.iLower + r))
      values[a] = 1.0;
    for (int d = 0; d < NR; ++d)
      rhs[size_t(d) * rows + r] = 0.0;
  }
  return 0;
}

/* Monolithic ndim-dof momentum system on the tile path (product:
 * build_mono_twin in nw_api.cu + the MomentumMonoP branch of ls_tile_kernel):
 * the emulator's linear system is the NODE graph (numDof 1, the skipped nodes
 * of the 3-dof system as its skipped rows);
 * the 3-dof graph is built here, row nd r + i of node row r.  values / rhs
 * sized for the 3-dof graph (rhs: one column).  Node fields: x, u, dudx, visc,
 * rho, mask as for emu_assemble kind 2. */
int
emu_assemble_mono(
  void* h,
  const double* const* nodeFields,
  const int* nodeNcomp,
  int nNodeFields,
  const double* area,
  const double* mdot,
  const double* pecfac,
  const void* opts,
  const int64_t* skipped3, /* skipped rows of the 3-dof system (whole nodes) */
  int64_t nSkipped3,
  double* values,
  double* rhs)
{
  Emu* e = static_cast<Emu*>(h);
  if (!e->hasLs || !e->lp.usable || e->g.numDof != 1) {
    e->err = "emu_assemble_mono needs the node graph's plan (numDof 1)";
    return 1;
  }
  const MeshPlan& mp = e->mp;
  const Graph& g1 = e->g;
  const LsPlan& lp = e->lp;
  constexpr int ND = 3;
  Graph g3;
  try {
    build_graph(
      mp, NW_LINSYS_HYPRE, ND, std::vector<int64_t>(skipped3, skipped3 + nSkipped3), g3);
  } catch (const std::exception& ex) {
    e->err = ex.what();
    return 1;
  }
  auto row3 = [&](int64_t r1) {
    return r1 < g1.numRowsOwned ? ND * r1
                                : g3.numRowsOwned + ND * (r1 - g1.numRowsOwned);
  };
  const nw_momentum_opts& o = *static_cast<const nw_momentum_opts*>(opts);
  std::vector<std::vector<double>> store;
  std::vector<const double*> comps;
  for (int f = 0; f < nNodeFields; ++f)
    store.push_back(to_slots(mp, nodeFields[f], nodeNcomp[f]));
  for (int f = 0; f < nNodeFields; ++f)
    for (int c = 0; c < nodeNcomp[f]; ++c)
      comps.push_back(store[f].data() + size_t(c) * mp.nSlots);
  const std::vector<double> sArea = edge_to_slots(mp, area, ND);
  const std::vector<double> sMdot = edge_to_slots(mp, mdot, 1);
  std::vector<double> sPec;
  if (pecfac)
    sPec = edge_to_slots(mp, pecfac, 1);
  const int64_t S = mp.nTileEdgeSlots;
  const int64_t nnz = g3.nnzOwned + g3.nnzShared, rows = g3.numRowsLocal();
  for (int64_t i = 0; i < nnz; ++i)
    values[i] = NAN;
  for (int64_t i = 0; i < rows; ++i)
    rhs[i] = NAN;
  const double invRelax = 1.0 / o.relax_fac;
  for (int64_t t = 0; t < mp.nTiles; ++t) {
    const TileHdr& hd = mp.tiles[t];
    const LsTileHdr& lh = lp.tiles[t];
    Staged st = stage(mp, hd, comps);
    std::vector<MomResult<ND>> res(hd.nEdges);
    std::vector<double> av(size_t(ND) * hd.nEdges);
    for (int j = 0; j < hd.nEdges; ++j) {
      const uint32_t v = mp.lr[hd.edge0 + j];
      const int l = v & 0xffff, r = v >> 16;
      double a[ND];
      for (int d = 0; d < ND; ++d)
        a[d] = av[size_t(d) * hd.nEdges + j] = sArea[size_t(d) * S + hd.edge0 + j];
      MomNode<ND> L, R;
      auto ld = [&](int i, MomNode<ND>& n) {
        for (int d = 0; d < ND; ++d) {
          n.x[d] = st(d, i);
          n.u[d] = st(ND + d, i);
        }
        for (int d = 0; d < ND * ND; ++d)
          n.g[d] = st(2 * ND + d, i);
        n.mu = st(2 * ND + ND * ND, i);
        n.rho = st(2 * ND + ND * ND + 1, i);
        n.mask = st(2 * ND + ND * ND + 2, i);
      };
      ld(l, L);
      ld(r, R);
      const double pf = pecfac ? sPec[hd.edge0 + j] : 0.0;
      if (o.has_vof)
        momentum_edge_vof<ND>(L, R, a, sMdot[hd.edge0 + j], pf, o, res[j]);
      else
        momentum_edge<ND>(L, R, a, sMdot[hd.edge0 + j], pf, o, res[j]);
    }
    /* row walk: ND rows per node, blocks written in place (the device stages
     * them and copies out; same values, same order of additions) */
    bool bad = false;
    std::vector<double> dg(size_t(ND) * ND * lh.nEnts, 0.0), rr(size_t(ND) * lh.nEnts, 0.0);
    const bool okWalk = walk_ell(
      lp.heEll.data() + lh.ellPtr, lp.sliceOff.data() + lh.slicePtr, lh.nEnts,
      lh.ellLen, [&](int ent, uint32_t hv) {
        const int j = he_edge(hv), side = he_side(hv), k = he_k(hv);
        const EntInfo& ei = lp.entInfo[lh.entPtr + ent];
        if (k >= ei.nnz || k == ei.diagK) {
          bad = true;
          return;
        }
        const MomResult<ND>& m = res[j];
        const int64_t r3 = row3(lp.entRhsRow[lh.entPtr + ent]);
        const int64_t rowLen = ND * (int64_t)ei.nnz;
        if (g3.rowLen(r3) != rowLen) {
          bad = true;
          return;
        }
        const double sXX = side ? m.sRR : m.sLL, sXY = side ? m.sRL : m.sLR;
        for (int i = 0; i < ND; ++i) {
          rr[size_t(ent) * ND + i] += side ? m.flux[i] : -m.flux[i];
          for (int c = 0; c < ND; ++c) {
            const double ns = -m.viscIp * av[size_t(i) * hd.nEdges + j] *
                              av[size_t(c) * hd.nEdges + j] * m.inv_axdx;
            const double s_ = (i == c) ? 1.0 : 0.0;
            dg[(size_t(ent) * ND + i) * ND + c] += s_ * sXX - ns * invRelax;
            double off = s_ * sXY + ns;
            double& dst = values[g3.rowPtr(r3 + i) + ND * k + c];
            if (hv & kHeDup)
              off += dst;
            dst = off;
          }
        }
      });
    if (!okWalk || bad) {
      e->err = "row ELL list malformed, slot out of row, or 3-dof row is not "
               "the blow-up of the node row";
      return 1;
    }
    for (int ent = 0; ent < lh.nEnts; ++ent) {
      const EntInfo& ei = lp.entInfo[lh.entPtr + ent];
      const int64_t r3 = row3(lp.entRhsRow[lh.entPtr + ent]);
      for (int i = 0; i < ND; ++i) {
        for (int c = 0; c < ND; ++c)
          values[g3.rowPtr(r3 + i) + ND * ei.diagK + c] =
            dg[(size_t(ent) * ND + i) * ND + c];
        rhs[r3 + i] = rr[size_t(ent) * ND + i];
      }
    }
  }
  for (int32_t r1 : lp.uncoveredRows)
    for (int i = 0; i < ND; ++i) {
      const int64_t r = row3(r1) + i;
      const int64_t a = g3.rowPtr(r), len = g3.rowLen(r);
      for (int64_t k = 0; k < len; ++k)
        values[a + k] = 0.0;
      if (r < g3.numRowsOwned &&
          std::binary_search(
            g3.periodicRowsOwned.begin(), g3.periodicRowsOwned.end(), g3.iLower + r))
        values[a] = 1.0;
      rhs[r] = 0.0;
    }
  return 0;
}

/* nodal gradient through the node-keyed half-edge lists; grad AoS out */
int
emu_nodal_grad(
  void* h, int dim1, const double* phi, const double* area, const double* vol,
  double* grad)
{
  Emu* e = static_cast<Emu*>(h);
  const MeshPlan& mp = e->mp;
  constexpr int ND = 3;
  std::vector<double> sPhi = to_slots(mp, phi, dim1);
  std::vector<double> sVol = to_slots(mp, vol, 1);
  std::vector<double> sArea = edge_to_slots(mp, area, ND);
  const int64_t S = mp.nTileEdgeSlots;
  const int NV = dim1 * ND;
  for (int64_t t = 0; t < mp.nTiles; ++t) {
    const TileHdr& hd = mp.tiles[t];
    std::vector<const double*> comps;
    for (int c = 0; c < dim1; ++c)
      comps.push_back(sPhi.data() + size_t(c) * mp.nSlots);
    Staged st = stage(mp, hd, comps);
    std::vector<double> out(size_t(NV) * hd.nOwn, 0.0);
    if (!walk_ell(
          mp.heNodeEll.data() + hd.ellPtrNode,
          mp.sliceOffNode.data() + hd.slicePtrNode, hd.nOwn, hd.ellLenNode,
          [&](int ent, uint32_t hv) {
            const int j = he_edge(hv);
            const uint32_t v = mp.lr[hd.edge0 + j];
            const int l = v & 0xffff, r = v >> 16;
            const double sgn = he_side(hv) ? -1.0 : 1.0;
            for (int i = 0; i < dim1; ++i) {
              const double phiIp = 0.5 * (st(i, l) + st(i, r));
              for (int d = 0; d < ND; ++d)
                out[size_t(i * ND + d) * hd.nOwn + ent] +=
                  (sgn * sArea[size_t(d) * S + hd.edge0 + j]) * phiIp;
            }
          })) {
      e->err = "node ELL list malformed";
      return 1;
    }
    for (int i = 0; i < hd.nOwn; ++i) {
      const double invVol = 1.0 / sVol[hd.node0 + i];
      for (int k = 0; k < NV; ++k)
        out[size_t(k) * hd.nOwn + i] *= invVol;
    }
    for (int i = 0; i < hd.nOwn; ++i) {
      const int32_t n = mp.nodeOfSlot[hd.node0 + i];
      for (int k = 0; k < NV; ++k)
        grad[size_t(n) * NV + k] = out[size_t(k) * hd.nOwn + i];
    }
  }
  return 0;
}

/* mdot per edge (caller order) through the tile-edge lists; both copies of a
 * cut edge must agree bit for bit */
int
emu_mdot(
  void* h, const double* const* nodeFields, const int* nodeNcomp,
  int nNodeFields, const double* area, double nocFac, double interp,
  double* mdotOut)
{
  Emu* e = static_cast<Emu*>(h);
  const MeshPlan& mp = e->mp;
  constexpr int ND = 3;
  std::vector<std::vector<double>> store;
  std::vector<const double*> comps;
  for (int f = 0; f < nNodeFields; ++f)
    store.push_back(to_slots(mp, nodeFields[f], nodeNcomp[f]));
  for (int f = 0; f < nNodeFields; ++f)
    for (int c = 0; c < nodeNcomp[f]; ++c)
      comps.push_back(store[f].data() + size_t(c) * mp.nSlots);
  std::vector<double> sArea = edge_to_slots(mp, area, ND);
  const int64_t S = mp.nTileEdgeSlots;
  std::vector<double> slots(S, 0.0);
  for (int64_t t = 0; t < mp.nTiles; ++t) {
    const TileHdr& hd = mp.tiles[t];
    Staged st = stage(mp, hd, comps);
    for (int j = 0; j < hd.nEdges; ++j) {
      const uint32_t v = mp.lr[hd.edge0 + j];
      const int l = v & 0xffff, r = v >> 16;
      double av[ND];
      for (int d = 0; d < ND; ++d)
        av[d] = sArea[size_t(d) * S + hd.edge0 + j];
      ContNode<ND> L, R;
      auto ld = [&](int i, ContNode<ND>& n) {
        for (int d = 0; d < ND; ++d) {
          n.x[d] = st(d, i);
          n.u[d] = st(ND + d, i);
          n.g[d] = st(2 * ND + d, i);
        }
        n.rho = st(3 * ND, i);
        n.p = st(3 * ND + 1, i);
        n.ud = st(3 * ND + 2, i);
      };
      ld(l, L);
      ld(r, R);
      slots[hd.edge0 + j] = mdot_core<ND>(L, R, av, nocFac, interp).tmdot;
    }
  }
  for (int64_t ed = 0; ed < mp.nEdges; ++ed) {
    const int32_t p = mp.primarySlotOfEdge[ed], s2 = mp.secondSlotOfEdge[ed];
    if (s2 >= 0 && std::memcmp(&slots[p], &slots[s2], sizeof(double)) != 0) {
      e->err = "the two copies of a cut edge disagree";
      return 1;
    }
    mdotOut[ed] = slots[p];
  }
  return 0;
}

/* plan statistics for design work: 32-byte sectors a warp-wide halo gather of
 * one field component touches, summed over tiles (out[0]), halo entries
 * (out[1]), tile-edge records (out[2]), ELL records row/node (out[3], out[4]) */
int
emu_plan_stats(void* h, int64_t* out)
{
  Emu* e = static_cast<Emu*>(h);
  const MeshPlan& mp = e->mp;
  int64_t sectors = 0, halo = 0, te = 0;
  for (int64_t t = 0; t < mp.nTiles; ++t) {
    const TileHdr& hd = mp.tiles[t];
    halo += hd.nHalo;
    te += hd.nEdges;
    for (int k0 = 0; k0 < hd.nHalo; k0 += 32) {
      std::vector<int64_t> sec;
      for (int k = k0; k < std::min(hd.nHalo, k0 + 32); ++k)
        sec.push_back(mp.haloNodes[hd.haloPtr + k] / 4);
      std::sort(sec.begin(), sec.end());
      sectors += std::unique(sec.begin(), sec.end()) - sec.begin();
    }
  }
  out[0] = sectors;
  out[1] = halo;
  out[2] = te;
  out[3] = e->hasLs ? (int64_t)e->lp.heEll.size() : 0;
  out[4] = (int64_t)mp.heNodeEll.size();
  return 0;
}

/* the same sector count for a field stored node-major with k components
 * (per-field AoS, 8k bytes per node): distinct 32-byte sectors touched by each
 * tile's halo list (whole tile, not per warp: an upper bound on reuse) and per
 * group of `group` consecutive halo entries (out[0], out[1]) */
int
emu_halo_sectors_aos(void* h, int k, int group, int64_t* out)
{
  Emu* e = static_cast<Emu*>(h);
  const MeshPlan& mp = e->mp;
  int64_t perTile = 0, perGroup = 0;
  for (int64_t t = 0; t < mp.nTiles; ++t) {
    const TileHdr& hd = mp.tiles[t];
    std::vector<int64_t> all;
    for (int k0 = 0; k0 < hd.nHalo; k0 += group) {
      std::vector<int64_t> sec;
      for (int q = k0; q < std::min(hd.nHalo, k0 + group); ++q) {
        const int64_t n = mp.haloNodes[hd.haloPtr + q];
        for (int64_t b = n * 8 * k / 32; b <= ((n + 1) * 8 * k - 1) / 32; ++b)
          sec.push_back(b);
      }
      std::sort(sec.begin(), sec.end());
      sec.erase(std::unique(sec.begin(), sec.end()), sec.end());
      perGroup += (int64_t)sec.size();
      all.insert(all.end(), sec.begin(), sec.end());
    }
    std::sort(all.begin(), all.end());
    perTile += std::unique(all.begin(), all.end()) - all.begin();
  }
  out[0] = perTile;
  out[1] = perGroup;
  return 0;
}

int
emu_geometry_cvfem(
  int topo, int64_t nElems, const int32_t* elemNodes, const double* coords,
  int64_t nEdges, const int32_t* edgeNodes, double* dnv, double* area)
{
  switch (topo) {
  case geo::TET4:
    emu_geo_block<geo::TET4>(nElems, elemNodes, coords, nEdges, edgeNodes, dnv, area);
    return 0;
  case geo::WED6:
    emu_geo_block<geo::WED6>(nElems, elemNodes, coords, nEdges, edgeNodes, dnv, area);
    return 0;
  case geo::PYR5:
    emu_geo_block<geo::PYR5>(nElems, elemNodes, coords, nEdges, edgeNodes, dnv, area);
    return 0;
  }
  return 1;
}

/* the product's peclet_eval (edge_physics.h), host build */
double
emu_peclet_eval(int form, double a, double b, double pecnum)
{
  nw_peclet_fn f;
  f.form = form;
  f.a = a;
  f.b = b;
  return peclet_eval(f, pecnum);
}

/* the product's van_leer (edge_physics.h), host build */
double
emu_van_leer(double dqm, double dqp, double eps)
{
  return van_leer(dqm, dqp, eps);
}

/* The default-option paths of momentum_edge / scalar_edge (edge_physics.h,
 * template DEF) against the general paths on n seeded random edges with
 * alpha = 0, alpha_upw = 1, hoUpwind = 1: returns the number of result doubles
 * whose bits differ (exact zeros of either sign count as equal). */
int64_t
emu_default_path_mismatches(int64_t n, uint64_t seed)
{
  uint64_t st = seed * 0x9E3779B97F4A7C15ull + 1;
  auto rnd = [&]() { /* splitmix64 -> (-1, 1) */
    st += 0x9E3779B97F4A7C15ull;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return 2.0 * (double(z >> 11) / 9007199254740992.0) - 1.0;
  };
  auto same = [](double a, double b) {
    if (a == 0.0 && b == 0.0)
      return true;
    return std::memcmp(&a, &b, sizeof(double)) == 0;
  };
  int64_t bad = 0;
  for (int64_t q = 0; q < n; ++q) {
    MomNode<3> L, R;
    ScalNode<3> sl, sr;
    double av[3];
    for (int d = 0; d < 3; ++d) {
      L.x[d] = rnd();
      R.x[d] = L.x[d] + 0.2 * rnd() + (d == int(q % 3) ? 0.5 : 0.0);
      L.u[d] = 5.0 * rnd();
      R.u[d] = L.u[d] + rnd();
      av[d] = (R.x[d] - L.x[d]) * (0.5 + 0.4 * rnd());
      sl.x[d] = L.x[d];
      sr.x[d] = R.x[d];
      sl.v[d] = L.u[d];
      sr.v[d] = R.u[d];
      sl.dq[d] = rnd();
      sr.dq[d] = rnd();
    }
    for (int d = 0; d < 9; ++d) {
      L.g[d] = rnd();
      R.g[d] = rnd();
    }
    L.mu = 1e-5 * (2.0 + rnd());
    R.mu = 1e-5 * (2.0 + rnd());
    L.rho = 1.2 + 0.1 * rnd();
    R.rho = 1.2 + 0.1 * rnd();
    L.mask = R.mask = 1.0;
    sl.q = 2.0 + rnd();
    sr.q = q % 7 == 0 ? sl.q : 2.0 + rnd(); /* flat field: limiter corner */
    sl.rho = L.rho;
    sr.rho = R.rho;
    sl.mu = 3.0 * L.mu;
    sr.mu = 3.0 * R.mu;
    const double mdot = q % 5 == 0 ? 0.0 : rnd();
    const double pecfac = q % 3 == 0 ? 1.0 - 1e-11 * (1.0 + rnd()) : 0.5 * (1.0 + rnd());
    nw_momentum_opts mo{};
    mo.include_divu = 0.0;
    mo.alpha = 0.0;
    mo.alpha_upw = 1.0;
    mo.ho_upwind = 1.0;
    mo.relax_fac = 0.7;
    mo.use_limiter = 1;
    mo.eps = 1e-16;
    MomResult<3> a, b;
    momentum_edge_t<3, true>(L, R, av, mdot, pecfac, mo, a);
    momentum_edge_t<3, false>(L, R, av, mdot, pecfac, mo, b);
    bad += !same(a.sLL, b.sLL) + !same(a.sLR, b.sLR) + !same(a.sRL, b.sRL) +
           !same(a.sRR, b.sRR);
    for (int d = 0; d < 3; ++d)
      bad += !same(a.flux[d], b.flux[d]);
    nw_scalar_opts so{};
    so.alpha = 0.0;
    so.alpha_upw = 1.0;
    so.ho_upwind = 1.0;
    so.relax_fac = 0.9;
    so.use_limiter = 1;
    so.eps = 1e-16;
    so.pf.form = NW_PECLET_TANH;
    so.pf.a = 2.0;
    so.pf.b = 1.0;
    double ca[4], cb[4], fa, fb;
    scalar_edge_t<3, true>(sl, sr, av, mdot, so, ca, fa);
    scalar_edge_t<3, false>(sl, sr, av, mdot, so, cb, fb);
    for (int k = 0; k < 4; ++k)
      bad += !same(ca[k], cb[k]);
    bad += !same(fa, fb);
  }
  return bad;
}

/* udiag_post_kernel's arithmetic over a list of nodes (csrc/edge_physics.h:
 * udiag_post_value) */
void
emu_udiag_post(
  int64_t n, double* udiag, const double* rho, const double* dvol,
  double projTimeScale, double alphaU)
{
  for (int64_t i = 0; i < n; ++i)
    udiag[i] = udiag_post_value(udiag[i], rho[i], dvol[i], projTimeScale, alphaU);
}

} // extern "C"
