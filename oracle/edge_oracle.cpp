/*
 * edge_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE ONLY, see edge_oracle.h).
 *
 * Restates, with the reference's per-edge operation order, the hot path of
 * Exawind/nalu-wind's edge-based CVFEM assembly.  Reference citations are
 * relative to /root/reference.  Build: `make -C oracle` (g++ -O2
 * -ffp-contract=off so results are plain IEEE double, reproducible across
 * hosts).
 *
 * Pinned (a) to every golden vector the reference's unit tests hold for the
 * path (tests/test_oracle_golds.py) and (b) to the reference's own edge
 * algorithm, master-element and Peclet sources, compiled unmodified against
 * stand-in Kokkos / STK / Realm headers and run in the build container
 * (oracle/Makefile.ref -> oracle/_ref; tests/test_reference_edge_runs.py,
 * tests/test_reference_runs.py): the kernels below reproduce the reference's
 * per-edge blocks, edge fields, gradients and geometry bit for bit.
 */
#include "edge_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <unordered_map>
#include <vector>

/* Threading.  Default 1 thread: the plain serial loop ("Kokkos Serial"), which is
 * what the golden fixtures pin.  With orc_set_num_threads(n > 1) the edge loops
 * run under OpenMP and every scatter add becomes an `omp atomic`, the analogue
 * of the reference's OpenMP build (Kokkos::atomic_add); used only to time the
 * CPU baseline on all host cores. */
static int g_threads = 1;
extern "C" void
orc_set_num_threads(int n)
{
  g_threads = n < 1 ? 1 : n;
}
extern "C" int
orc_get_num_threads(void)
{
  return g_threads;
}
static inline void
add_to(double& dst, double v)
{
  if (g_threads > 1) {
#pragma omp atomic
    dst += v;
  } else
    dst += v;
}
#define ORC_EDGE_LOOP \
  _Pragma("omp parallel for schedule(static) if (g_threads > 1) num_threads(g_threads)")

namespace {

constexpr int kMaxDim = 3;
constexpr int kMaxRhs = 2 * kMaxDim;

/* include/edge_kernels/EdgeKernelUtils.h:18-24 */
inline double
van_leer(double dqm, double dqp, double eps)
{
  return (2.0 * (dqm * dqp + std::fabs(dqm * dqp))) /
         ((dqm + dqp) * (dqm + dqp) + eps);
}

} // namespace

/* the limiter by itself, for tests/test_option_matrix_cpu.py */
extern "C" double
orc_van_leer(double dqm, double dqp, double eps)
{
  return van_leer(dqm, dqp, eps);
}

/* ------------------------------------------------------------------ */
/*  sinks                                                              */
/* ------------------------------------------------------------------ */

struct orc_applier
{
  virtual ~orc_applier() {}
  /* lhs: row-major n x n with n = nEnt*ldDof, rhs: n */
  virtual void apply(
    int nEnt, const int32_t* nodes, const double* rhs, const double* lhs,
    int n) = 0;
};

namespace {

/* unit_tests/UnitTestLinearSystem.h:42-72 */
struct DenseApplier : orc_applier
{
  int64_t nNodes;
  int numDof;
  std::vector<double> lhs_, rhs_;
  DenseApplier(int64_t nn, int nd) : nNodes(nn), numDof(nd)
  {
    const size_t n = size_t(nn) * nd;
    lhs_.assign(n * n, 0.0);
    rhs_.assign(n, 0.0);
  }
  void apply(
    int nEnt, const int32_t* nodes, const double* rhs, const double* lhs,
    int n) override
  {
    const size_t N = size_t(nNodes) * numDof;
    for (int i = 0; i < nEnt; ++i) {
      const size_t ioff = size_t(nodes[i]) * numDof;
      for (int d = 0; d < numDof; ++d)
        add_to(rhs_[ioff + d], rhs[i * numDof + d]);
    }
    for (int i = 0; i < nEnt; ++i) {
      const size_t ioff = size_t(nodes[i]) * numDof;
      for (int j = 0; j < nEnt; ++j) {
        const size_t joff = size_t(nodes[j]) * numDof;
        for (int d = 0; d < numDof; ++d) {
          const int ii = i * numDof + d;
          const int jj = j * numDof + d;
          add_to(lhs_[(ioff + d) * N + (joff + d)], lhs[ii * n + jj]);
        }
      }
    }
  }
};

} // namespace

/* ------------------------------------------------------------------ */
/*  hypre-IJ style graph                                               */
/* ------------------------------------------------------------------ */

struct orc_graph
{
  int numDof;
  int64_t iLower, iUpper; /* inclusive, already scaled by numDof */
  int64_t numRows;

  /* host graph containers, as in HypreLinearSystem.h */
  std::vector<std::vector<int64_t>> columnsOwned;
  std::vector<unsigned> rowCountOwned;
  std::map<int64_t, std::vector<int64_t>> columnsShared;
  std::map<int64_t, unsigned> rowCountShared;
  std::set<int64_t> skippedRows;

  /* finalized */
  int64_t numRowsOwned = 0, nnzOwned = 0, numRowsShared = 0, nnzShared = 0;
  std::vector<int64_t> rowStartOwned, rowStartShared;
  std::vector<int64_t> cols, rows;
  std::vector<int64_t> rowIndicesShared;
  std::vector<int64_t> periodicRowsOwned;
  std::unordered_map<int64_t, int64_t> mapShared;
};

extern "C" orc_graph*
orc_graph_create(int num_dof, int64_t i_lower, int64_t i_upper)
{
  /* beginLinearSystemConstruction: src/HypreLinearSystem.C:97-208.
   * iLower_/iUpper_ are node offsets * numDof, iUpper inclusive. */
  auto* g = new orc_graph;
  g->numDof = num_dof;
  g->iLower = i_lower;
  g->iUpper = i_upper;
  g->numRows = i_upper - i_lower + 1;
  g->columnsOwned.resize(g->numRows);
  g->rowCountOwned.assign(g->numRows, 0u);
  return g;
}

extern "C" void
orc_graph_destroy(orc_graph* g)
{
  delete g;
}

extern "C" void
orc_graph_set_skipped(orc_graph* g, const int64_t* rows, int64_t n)
{
  for (int64_t i = 0; i < n; ++i)
    g->skippedRows.insert(rows[i]);
}

namespace {

/* fill_owned_shared_data_structures_1DoF: src/HypreLinearSystem.C:211-236 */
void
fill_1dof(orc_graph* g, unsigned numNodes, const std::vector<int64_t>& hids)
{
  for (unsigned i = 0; i < numNodes; ++i) {
    const int64_t hid = hids[i];
    if (hid >= g->iLower && hid <= g->iUpper) {
      const int64_t lid = hid - g->iLower;
      g->rowCountOwned[lid]++;
      auto& c = g->columnsOwned[lid];
      c.insert(c.end(), hids.begin(), hids.end());
    } else {
      auto it = g->rowCountShared.find(hid);
      if (it != g->rowCountShared.end()) {
        it->second++;
        auto& c = g->columnsShared.at(hid);
        c.insert(c.end(), hids.begin(), hids.end());
      } else {
        g->rowCountShared.insert(std::make_pair(hid, 1u));
        g->columnsShared.insert(std::make_pair(hid, hids));
      }
    }
  }
}

/* fill_owned_shared_data_structures: src/HypreLinearSystem.C:238-269 */
void
fill_ndof(
  orc_graph* g,
  unsigned numNodes,
  const std::vector<int64_t>& hids,
  const std::vector<int64_t>& columns)
{
  for (unsigned i = 0; i < numNodes; ++i) {
    const int64_t hid = hids[i];
    for (int d = 0; d < g->numDof; ++d) {
      const int64_t HID = hid * g->numDof + d;
      if (HID >= g->iLower && HID <= g->iUpper) {
        const int64_t lid = HID - g->iLower;
        g->rowCountOwned[lid]++;
        auto& c = g->columnsOwned[lid];
        c.insert(c.end(), columns.begin(), columns.end());
      } else {
        auto it = g->rowCountShared.find(HID);
        if (it != g->rowCountShared.end()) {
          it->second++;
          auto& c = g->columnsShared.at(HID);
          c.insert(c.end(), columns.begin(), columns.end());
        } else {
          g->rowCountShared.insert(std::make_pair(HID, 1u));
          g->columnsShared.insert(std::make_pair(HID, columns));
        }
      }
    }
  }
}

void
graph_add_entities(
  orc_graph* g,
  int64_t nEnt,
  unsigned nodesPerEnt,
  const int32_t* entNodes,
  const int64_t* node_hid)
{
  std::vector<int64_t> hids(nodesPerEnt);
  std::vector<int64_t> columns(nodesPerEnt * g->numDof);
  for (int64_t k = 0; k < nEnt; ++k) {
    for (unsigned i = 0; i < nodesPerEnt; ++i) {
      hids[i] = node_hid[entNodes[k * nodesPerEnt + i]];
      /* fill_hids_columns: :271-283 */
      for (int d = 0; d < g->numDof; ++d)
        columns[i * g->numDof + d] = hids[i] * g->numDof + d;
    }
    if (g->numDof == 1)
      fill_1dof(g, nodesPerEnt, hids);
    else
      fill_ndof(g, nodesPerEnt, hids, columns);
  }
}

/* sort + scan-unique as in :1044-1054 */
int64_t
push_sorted_unique(std::vector<int64_t> columns, std::vector<int64_t>& out)
{
  std::sort(columns.begin(), columns.end());
  int64_t count = 1;
  int64_t col = columns[0];
  for (size_t i = 1; i < columns.size(); ++i) {
    if (columns[i] != col) {
      out.push_back(col);
      col = columns[i];
      count++;
    }
  }
  out.push_back(col);
  return count;
}

} // namespace

extern "C" void
orc_graph_add_edges(
  orc_graph* g,
  int64_t n_edges,
  const int32_t* edge_nodes,
  const int64_t* node_hid)
{
  /* buildEdgeToNodeGraph: src/HypreLinearSystem.C:412-478 */
  graph_add_entities(g, n_edges, 2, edge_nodes, node_hid);
}

extern "C" void
orc_graph_add_nodes(
  orc_graph* g, int64_t n_nodes, const int32_t* nodes, const int64_t* node_hid)
{
  /* buildNodeGraph: src/HypreLinearSystem.C:286-340 */
  graph_add_entities(g, n_nodes, 1, nodes, node_hid);
}

extern "C" void
orc_graph_finalize(orc_graph* g)
{
  /* buildCoeffApplierDeviceOwnedDataStructures: :999-1137 (no overset) */
  std::vector<int64_t> colsOwned, countOwned;
  g->periodicRowsOwned.clear();
  for (int64_t j = g->iLower; j <= g->iUpper; ++j) {
    const int64_t jShift = j - g->iLower;
    int64_t cnt = 1;
    const auto& columns = g->columnsOwned[jShift];
    if (g->skippedRows.find(j) != g->skippedRows.end()) {
      colsOwned.push_back(j); /* Dirichlet row: diagonal only */
    } else if (columns.size() == 0) {
      colsOwned.push_back(j); /* untouched row == periodic slave */
      g->periodicRowsOwned.push_back(j);
    } else if (columns.size() == 1) {
      colsOwned.push_back(j);
    } else {
      cnt = push_sorted_unique(columns, colsOwned);
    }
    countOwned.push_back(cnt);
  }
  g->numRowsOwned = (int64_t)countOwned.size();
  g->nnzOwned = (int64_t)colsOwned.size();
  g->rowStartOwned.assign(g->numRowsOwned + 1, 0);
  for (int64_t i = 0; i < g->numRowsOwned; ++i)
    g->rowStartOwned[i + 1] = g->rowStartOwned[i] + countOwned[i];

  /* buildCoeffApplierDeviceSharedDataStructures: :1143-1236 */
  std::vector<int64_t> colsShared, countShared;
  g->rowIndicesShared.clear();
  for (auto it = g->rowCountShared.begin(); it != g->rowCountShared.end();
       ++it) {
    const int64_t hid = it->first;
    const auto& columns = g->columnsShared[hid];
    int64_t cnt = 1;
    if (g->skippedRows.find(hid) != g->skippedRows.end()) {
      continue;
    } else if (columns.size() == 1) {
      colsShared.push_back(hid);
    } else if (columns.size() > 1) {
      cnt = push_sorted_unique(columns, colsShared);
    } else
      continue;
    g->rowIndicesShared.push_back(hid);
    countShared.push_back(cnt);
  }
  g->numRowsShared = (int64_t)g->rowIndicesShared.size();
  g->nnzShared = (int64_t)colsShared.size();
  g->rowStartShared.assign(g->numRowsShared + 1, 0);
  g->mapShared.clear();
  for (int64_t i = 0; i < g->numRowsShared; ++i) {
    g->rowStartShared[i + 1] = g->rowStartShared[i] + countShared[i];
    g->mapShared[g->rowIndicesShared[i]] = i; /* init_shared_map :1233 */
  }

  /* computeRowSizes: :956-993 -- monolithic cols / rows arrays, owned first */
  g->cols.clear();
  g->cols.insert(g->cols.end(), colsOwned.begin(), colsOwned.end());
  g->cols.insert(g->cols.end(), colsShared.begin(), colsShared.end());
  g->rows.assign(g->cols.size(), 0);
  int64_t k = 0;
  for (int64_t i = 0; i < g->numRowsOwned; ++i)
    for (int64_t j = 0; j < countOwned[i]; ++j)
      g->rows[k++] = g->iLower + i;
  k = g->nnzOwned;
  for (int64_t i = 0; i < g->numRowsShared; ++i)
    for (int64_t j = 0; j < countShared[i]; ++j)
      g->rows[k++] = g->rowIndicesShared[i];

  /* :1319-1330 the host containers are cleared for the next build */
  for (auto& c : g->columnsOwned)
    c.clear();
  std::fill(g->rowCountOwned.begin(), g->rowCountOwned.end(), 0u);
  g->columnsShared.clear();
  g->rowCountShared.clear();
}

extern "C" int64_t
orc_graph_size(const orc_graph* g, int what)
{
  switch (what) {
  case ORC_G_NUM_ROWS_OWNED:
    return g->numRowsOwned;
  case ORC_G_NNZ_OWNED:
    return g->nnzOwned;
  case ORC_G_NUM_ROWS_SHARED:
    return g->numRowsShared;
  case ORC_G_NNZ_SHARED:
    return g->nnzShared;
  case ORC_G_NUM_PERIODIC_ROWS:
    return (int64_t)g->periodicRowsOwned.size();
  }
  return -1;
}

extern "C" void
orc_graph_copy(const orc_graph* g, int what, int64_t* out)
{
  const std::vector<int64_t>* v = nullptr;
  switch (what) {
  case ORC_G_ROW_START_OWNED:
    v = &g->rowStartOwned;
    break;
  case ORC_G_ROW_START_SHARED:
    v = &g->rowStartShared;
    break;
  case ORC_G_COLS:
    v = &g->cols;
    break;
  case ORC_G_ROWS:
    v = &g->rows;
    break;
  case ORC_G_ROW_INDICES_SHARED:
    v = &g->rowIndicesShared;
    break;
  case ORC_G_PERIODIC_ROWS:
    v = &g->periodicRowsOwned;
    break;
  }
  if (v && !v->empty())
    std::memcpy(out, v->data(), v->size() * sizeof(int64_t));
}

/* ------------------------------------------------------------------ */
/*  hypre coefficient applier                                          */
/* ------------------------------------------------------------------ */

namespace {

struct HypreApplier : orc_applier
{
  const orc_graph* g;
  std::vector<int64_t> nodeHid;
  int uvwDim; /* 0: HypreLinSysCoeffApplier; >0: UVW with that many rhs */
  int nRhs;
  int64_t totalRows;
  std::vector<double> values, rhs;       /* rhs column-major [totalRows][nRhs] */
  std::vector<double> absValues, absRhs; /* |.| scatter, tolerance scale */
  bool trackAbs = true; /* off: timing runs (the reference has no such sums) */
  bool logOn = false;
  int64_t logCalls = 0, callNo = 0;
  int logN = 0;
  std::vector<int64_t> logSlots, logRhs;

  HypreApplier(const orc_graph* gg, const int64_t* hid, int64_t nn, int uvw)
    : g(gg), nodeHid(hid, hid + nn), uvwDim(uvw)
  {
    nRhs = uvw > 0 ? uvw : 1;
    totalRows = g->numRowsOwned + g->numRowsShared;
    reset();
  }

  /* resetCoeffApplierData: src/HypreLinearSystem.C:1386-1430 */
  void reset()
  {
    values.assign(g->cols.size(), 0.0);
    rhs.assign(size_t(totalRows) * nRhs, 0.0);
    absValues.assign(g->cols.size(), 0.0);
    absRhs.assign(size_t(totalRows) * nRhs, 0.0);
    for (int64_t hid : g->periodicRowsOwned) {
      const int64_t matIndex = g->rowStartOwned[hid - g->iLower];
      values[matIndex] = 1.0;
      absValues[matIndex] = 1.0;
      for (int d = 0; d < nRhs; ++d)
        rhs[size_t(d) * totalRows + (hid - g->iLower)] = 0.0;
    }
    callNo = 0;
  }

  /* HypreLinSysCoeffApplier::sort, :1960-2055: for N==2 a single
   * compare-exchange, N<=4 fixed networks, else bubble sort; all are stable
   * orderings of distinct ids, restated here as one stable insertion sort
   * (ids within one entity set are distinct unless periodic aliasing, where
   * the relative order of equal ids follows the strict '>' comparisons). */
  static void sort_ids(int64_t* ids, int* perm, unsigned N)
  {
    if (N == 2) {
      if (ids[0] > ids[1]) {
        std::swap(ids[0], ids[1]);
        std::swap(perm[0], perm[1]);
      }
      return;
    }
    for (unsigned i = 0; i + 1 < N; ++i)
      for (unsigned j = 0; j + 1 < N - i; ++j)
        if (ids[j] > ids[j + 1]) {
          std::swap(ids[j], ids[j + 1]);
          std::swap(perm[j], perm[j + 1]);
        }
  }

  void log_slot(int n, int ii, int kk, int64_t idx)
  {
    if (logOn && g_threads == 1 && callNo < logCalls)
      logSlots[(size_t(callNo) * n + ii) * n + kk] = idx;
  }
  void log_rhs(int n, int ii, int64_t idx)
  {
    if (logOn && g_threads == 1 && callNo < logCalls)
      logRhs[size_t(callNo) * n + ii] = idx;
  }

  void add_rhs(int64_t row, int d, double v)
  {
    add_to(rhs[size_t(d) * totalRows + row], v);
    if (trackAbs)
      add_to(absRhs[size_t(d) * totalRows + row], std::fabs(v));
  }
  void add_val(int64_t idx, double v)
  {
    add_to(values[idx], v);
    if (trackAbs)
      add_to(absValues[idx], std::fabs(v));
  }

  /* sum_into_1DoF: src/HypreLinearSystem.C:2165-2239 */
  void sum_into_1dof(
    unsigned nEnt, const int32_t* nodes, const double* r, const double* lhs,
    int n)
  {
    int64_t localIds[8];
    int perm[8];
    for (unsigned i = 0; i < nEnt; ++i) {
      localIds[i] = nodeHid[nodes[i]];
      perm[i] = int(i);
    }
    sort_ids(localIds, perm, nEnt);
    const int64_t memShift = g->nnzOwned;
    for (unsigned i = 0; i < nEnt; ++i) {
      const int64_t hid = localIds[i];
      if (g->skippedRows.count(hid))
        continue;
      const int ii = perm[i];
      const double* cur = &lhs[ii * n];
      if (hid >= g->iLower && hid <= g->iUpper) {
        const int64_t index = hid - g->iLower;
        int64_t matIndex = g->rowStartOwned[index];
        for (unsigned k = 0; k < nEnt; ++k) {
          const int64_t col = localIds[k];
          while (g->cols[matIndex] < col)
            matIndex++;
          const int kk = perm[k];
          add_val(matIndex, cur[kk]);
          log_slot(n, ii, kk, matIndex);
          matIndex++;
        }
        add_rhs(index, 0, r[ii]);
        log_rhs(n, ii, index);
      } else {
        auto it = g->mapShared.find(hid);
        if (it == g->mapShared.end())
          continue;
        const int64_t index = it->second;
        int64_t matIndex = g->rowStartShared[index] + memShift;
        for (unsigned k = 0; k < nEnt; ++k) {
          const int64_t col = localIds[k];
          while (g->cols[matIndex] < col)
            matIndex++;
          const int kk = perm[k];
          add_val(matIndex, cur[kk]);
          log_slot(n, ii, kk, matIndex);
          matIndex++;
        }
        /* rhs_row_start_shared_(index) == index (one rhs slot per row) */
        const int64_t rhsIndex = index + (g->iUpper - g->iLower + 1);
        add_rhs(rhsIndex, 0, r[ii]);
        log_rhs(n, ii, rhsIndex);
      }
    }
  }

  /* sum_into: src/HypreLinearSystem.C:2059-2161 */
  void sum_into_ndof(
    unsigned nEnt, const int32_t* nodes, const double* r, const double* lhs,
    int n)
  {
    const unsigned numDof = unsigned(g->numDof);
    const unsigned numRows = nEnt * numDof;
    int64_t localIds[4 * kMaxDim];
    int perm[4 * kMaxDim];
    for (unsigned i = 0; i < nEnt; ++i) {
      const int64_t hid = nodeHid[nodes[i]];
      for (unsigned d = 0; d < numDof; ++d) {
        const unsigned lid = i * numDof + d;
        localIds[lid] = hid * numDof + d;
        perm[lid] = int(lid);
      }
    }
    sort_ids(localIds, perm, numRows);
    const int64_t memShift = g->nnzOwned;
    for (unsigned i = 0; i < nEnt; ++i) {
      const unsigned ix = i * numDof;
      int64_t hid = localIds[ix];
      /* quirk: only the first dof's row id is tested (:2095-2099) */
      if (g->skippedRows.count(hid))
        continue;
      if (hid >= g->iLower && hid <= g->iUpper) {
        for (unsigned d = 0; d < numDof; ++d) {
          const unsigned ir = ix + d;
          hid = localIds[ir];
          const int ii = perm[ir];
          const double* cur = &lhs[ii * n];
          const int64_t index = hid - g->iLower;
          int64_t matIndex = g->rowStartOwned[index];
          for (unsigned k = 0; k < numRows; ++k) {
            const int64_t col = localIds[k];
            while (g->cols[matIndex] < col)
              matIndex++;
            const int kk = perm[k];
            add_val(matIndex, cur[kk]);
            log_slot(n, ii, kk, matIndex);
          }
          add_rhs(index, 0, r[ii]);
          log_rhs(n, ii, index);
        }
      } else {
        for (unsigned d = 0; d < numDof; ++d) {
          const unsigned ir = ix + d;
          hid = localIds[ir];
          const int ii = perm[ir];
          const double* cur = &lhs[ii * n];
          auto it = g->mapShared.find(hid);
          if (it == g->mapShared.end())
            continue;
          const int64_t index = it->second;
          int64_t matIndex = g->rowStartShared[index] + memShift;
          for (unsigned k = 0; k < numRows; ++k) {
            const int64_t col = localIds[k];
            while (g->cols[matIndex] < col)
              matIndex++;
            const int kk = perm[k];
            add_val(matIndex, cur[kk]);
            log_slot(n, ii, kk, matIndex);
          }
          const int64_t rhsIndex = index + (g->iUpper - g->iLower + 1);
          add_rhs(rhsIndex, 0, r[ii]);
          log_rhs(n, ii, rhsIndex);
        }
      }
    }
  }

  /* HypreUVWLinSysCoeffApplier::sum_into:
   * src/HypreUVWLinearSystem.C:695-767.  lhs is the full
   * (nEnt*nDim)^2 block; only entries (i*nDim, k*nDim) reach the matrix. */
  void sum_into_uvw(
    unsigned nEnt, const int32_t* nodes, const double* r, const double* lhs,
    int n)
  {
    const unsigned nDim = unsigned(uvwDim);
    int64_t localIds[8];
    int perm[8];
    for (unsigned i = 0; i < nEnt; ++i) {
      localIds[i] = nodeHid[nodes[i]];
      perm[i] = int(i * nDim);
    }
    sort_ids(localIds, perm, nEnt);
    const int64_t memShift = g->nnzOwned;
    for (unsigned i = 0; i < nEnt; ++i) {
      const int ix = perm[i];
      const int64_t hid = localIds[i];
      if (g->skippedRows.count(hid))
        continue;
      int64_t index, matIndex, rhsIndex;
      if (hid >= g->iLower && hid <= g->iUpper) {
        index = hid - g->iLower;
        matIndex = g->rowStartOwned[index];
        rhsIndex = index;
      } else {
        auto it = g->mapShared.find(hid);
        if (it == g->mapShared.end())
          continue;
        index = it->second;
        matIndex = g->rowStartShared[index] + memShift;
        rhsIndex = index + (g->iUpper - g->iLower + 1);
      }
      for (unsigned k = 0; k < nEnt; ++k) {
        const int64_t col = localIds[k];
        while (g->cols[matIndex] < col)
          matIndex++;
        add_val(matIndex, lhs[ix * n + perm[k]]);
        log_slot(int(nEnt), ix / int(nDim), perm[k] / int(nDim), matIndex);
        matIndex++;
      }
      for (unsigned d = 0; d < nDim; ++d)
        add_rhs(rhsIndex, int(d), r[ix + int(d)]);
      log_rhs(int(nEnt), ix / int(nDim), rhsIndex);
    }
  }

  /* HypreLinSysCoeffApplier::operator(): :2241-2260 */
  void apply(
    int nEnt, const int32_t* nodes, const double* r, const double* lhs,
    int n) override
  {
    if (uvwDim > 0)
      sum_into_uvw(unsigned(nEnt), nodes, r, lhs, n);
    else if (g->numDof == 1)
      sum_into_1dof(unsigned(nEnt), nodes, r, lhs, n);
    else
      sum_into_ndof(unsigned(nEnt), nodes, r, lhs, n);
    if (g_threads == 1)
      callNo++;
  }
};

} // namespace

extern "C" orc_applier*
orc_applier_dense_create(int64_t n_nodes, int num_dof)
{
  return new DenseApplier(n_nodes, num_dof);
}

extern "C" void
orc_applier_dense_get(const orc_applier* a, double* lhs, double* rhs)
{
  auto* d = static_cast<const DenseApplier*>(a);
  std::memcpy(lhs, d->lhs_.data(), d->lhs_.size() * sizeof(double));
  std::memcpy(rhs, d->rhs_.data(), d->rhs_.size() * sizeof(double));
}

/* records the local blocks in call order (one thread: edge order) -- for the
 * edge-by-edge comparison with the reference's own lambdas, oracle/_ref */
namespace {
struct RecordApplier : orc_applier
{
  int n_ = 0;
  std::vector<double> lhs_, rhs_;
  void apply(
    int, const int32_t*, const double* rhs, const double* lhs, int n) override
  {
    n_ = n;
    lhs_.insert(lhs_.end(), lhs, lhs + size_t(n) * n);
    rhs_.insert(rhs_.end(), rhs, rhs + n);
  }
};
} // namespace

extern "C" orc_applier*
orc_applier_record_create()
{
  return new RecordApplier();
}

/* number of recorded calls; *n = block size of the last call */
extern "C" int64_t
orc_applier_record_count(const orc_applier* a, int* n)
{
  auto* r = static_cast<const RecordApplier*>(a);
  *n = r->n_;
  return r->n_ ? int64_t(r->rhs_.size()) / r->n_ : 0;
}

extern "C" void
orc_applier_record_get(const orc_applier* a, double* lhs, double* rhs)
{
  auto* r = static_cast<const RecordApplier*>(a);
  std::memcpy(lhs, r->lhs_.data(), r->lhs_.size() * sizeof(double));
  std::memcpy(rhs, r->rhs_.data(), r->rhs_.size() * sizeof(double));
}

extern "C" orc_applier*
orc_applier_hypre_create(
  const orc_graph* g, const int64_t* node_hid, int64_t n_nodes, int uvw_ndim)
{
  return new HypreApplier(g, node_hid, n_nodes, uvw_ndim);
}

/* CoeffApplier::operator() (include/LinearSystem.h:62-70) for a batch of
 * entities: lhs [nEnt][n][n] row-major, rhs [nEnt][n] */
extern "C" void
orc_applier_apply(
  orc_applier* a, int64_t n_entities, int nodes_per_entity,
  const int32_t* entity_nodes, const double* lhs, const double* rhs, int n)
{
  for (int64_t e = 0; e < n_entities; ++e)
    a->apply(
      nodes_per_entity, entity_nodes + e * nodes_per_entity, rhs + e * n,
      lhs + e * n * n, n);
}

/* ---- node kernels (AssembleNGPNodeSolverAlgorithm::execute,
 * src/AssembleNGPNodeSolverAlgorithm.C:85-146: per selected node zero the
 * 1-node block, run the kernel, hand it to the CoeffApplier) ---- */

/* ScalarMassBDFNodeKernel::execute, src/node_kernels/ScalarMassBDFNodeKernel.C:72-96 */
extern "C" void
orc_scalar_mass_bdf_node(
  int64_t n_sel, const int32_t* nodes, const double* qNm1, const double* qN,
  const double* qNp1, const double* rhoNm1, const double* rhoN,
  const double* rhoNp1, const double* dnvNm1, const double* dnvN,
  const double* dnvNp1, double dt, double gamma1, double gamma2, double gamma3,
  orc_applier* a)
{
  for (int64_t i = 0; i < n_sel; ++i) {
    const int32_t n = nodes[i];
    double lhs = 0.0, rhs = 0.0;
    const double lhsTime = gamma1 * rhoNp1[n] * dnvNp1[n] / dt;
    rhs -= (gamma1 * rhoNp1[n] * qNp1[n] * dnvNp1[n] +
            gamma2 * qN[n] * rhoN[n] * dnvN[n] +
            gamma3 * qNm1[n] * rhoNm1[n] * dnvNm1[n]) /
           dt;
    lhs += lhsTime;
    a->apply(1, &n, &rhs, &lhs, 1);
  }
}

/* MomentumMassBDFNodeKernel::execute, src/node_kernels/MomentumMassBDFNodeKernel.C:80-107 */
extern "C" void
orc_momentum_mass_bdf_node(
  int ndim, int64_t n_sel, const int32_t* nodes, const double* uNm1,
  const double* uN, const double* uNp1, const double* rhoNm1,
  const double* rhoN, const double* rhoNp1, const double* dnvNm1,
  const double* dnvN, const double* dnvNp1, const double* dpdx, double dt,
  double gamma1, double gamma2, double gamma3, orc_applier* a)
{
  for (int64_t q = 0; q < n_sel; ++q) {
    const int32_t n = nodes[q];
    double lhs[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, rhs[3] = {0, 0, 0};
    const double lhsfac = gamma1 * rhoNp1[n] * dnvNp1[n] / dt;
    for (int i = 0; i < ndim; ++i) {
      rhs[i] += -(gamma1 * rhoNp1[n] * uNp1[n * ndim + i] * dnvNp1[n] +
                  gamma2 * rhoN[n] * uN[n * ndim + i] * dnvN[n] +
                  gamma3 * rhoNm1[n] * uNm1[n * ndim + i] * dnvNm1[n]) /
                  dt -
                dpdx[n * ndim + i] * dnvNp1[n];
      lhs[i * ndim + i] += lhsfac;
    }
    a->apply(1, &n, rhs, lhs, ndim);
  }
}

/* ContinuityMassBDFNodeKernel::execute, src/node_kernels/ContinuityMassBDFNodeKernel.C:63-84 */
extern "C" void
orc_continuity_mass_bdf_node(
  int64_t n_sel, const int32_t* nodes, const double* rhoNm1, const double* rhoN,
  const double* rhoNp1, const double* dnvNm1, const double* dnvN,
  const double* dnvNp1, double dt, double gamma1, double gamma2, double gamma3,
  orc_applier* a)
{
  for (int64_t i = 0; i < n_sel; ++i) {
    const int32_t n = nodes[i];
    double lhs = 0.0, rhs = 0.0;
    rhs -= (gamma1 * rhoNp1[n] * dnvNp1[n] + gamma2 * rhoN[n] * dnvN[n] +
            gamma3 * rhoNm1[n] * dnvNm1[n]) /
           dt * (gamma1 / dt);
    a->apply(1, &n, &rhs, &lhs, 1);
  }
}

/* HypreLinSysCoeffApplier::reset_rows (src/HypreLinearSystem.C:2262-2315) /
 * HypreUVWLinSysCoeffApplier::reset_rows (src/HypreUVWLinearSystem.C:787-835):
 * zero the rows of the given nodes, diagonal = diag_value, rhs = rhs_residual */
extern "C" void
orc_applier_hypre_reset_rows(
  orc_applier* a, int64_t n_nodes, const int32_t* nodes, double diag_value,
  double rhs_residual)
{
  auto* h = static_cast<HypreApplier*>(a);
  const orc_graph* g = h->g;
  const int numDof = h->uvwDim > 0 ? 1 : g->numDof;
  const int64_t memShift = g->nnzOwned;
  for (int64_t i = 0; i < n_nodes; ++i) {
    const int64_t lid = h->nodeHid[nodes[i]];
    for (int d = 0; d < numDof; ++d) {
      const int64_t hid = lid * numDof + d;
      int64_t lower, upper, rhsIndex;
      if (hid >= g->iLower && hid <= g->iUpper) {
        const int64_t index = hid - g->iLower;
        lower = g->rowStartOwned[index];
        upper = g->rowStartOwned[index + 1];
        rhsIndex = index;
      } else {
        auto it = g->mapShared.find(hid);
        if (it == g->mapShared.end())
          continue;
        const int64_t index = it->second;
        lower = g->rowStartShared[index] + memShift;
        upper = g->rowStartShared[index + 1] + memShift;
        rhsIndex = index + (g->iUpper - g->iLower + 1);
      }
      for (int64_t k = lower; k < upper; ++k) {
        h->values[k] = 0.0;
        h->absValues[k] = 0.0;
        if (g->cols[k] == hid) {
          h->values[k] = diag_value;
          h->absValues[k] = std::fabs(diag_value);
        }
      }
      for (int q = 0; q < h->nRhs; ++q) {
        h->rhs[size_t(q) * h->totalRows + rhsIndex] = rhs_residual;
        h->absRhs[size_t(q) * h->totalRows + rhsIndex] = std::fabs(rhs_residual);
      }
    }
  }
}

/* HypreLinearSystem::applyDirichletBCs (src/HypreLinearSystem.C:2407-2457) /
 * HypreUVWLinearSystem::applyDirichletBCs (src/HypreUVWLinearSystem.C:377-427),
 * locally-owned nodes only: first entry of the row = 1, rhs = bc - solution.
 * solution / bc: [n_local_nodes][ncomp] in the reference layout. */
extern "C" void
orc_applier_hypre_dirichlet(
  orc_applier* a, int64_t n_nodes, const int32_t* nodes, const double* solution,
  const double* bc_values, int ncomp)
{
  auto* h = static_cast<HypreApplier*>(a);
  const orc_graph* g = h->g;
  for (int64_t i = 0; i < n_nodes; ++i) {
    const int64_t hid = h->nodeHid[nodes[i]];
    if (h->uvwDim > 0) {
      if (hid < g->iLower || hid > g->iUpper)
        continue;
      const int64_t matIndex = g->rowStartOwned[hid - g->iLower];
      h->values[matIndex] = 1.0;
      h->absValues[matIndex] = 1.0;
      for (int d = 0; d < h->uvwDim; ++d) {
        const double v = bc_values[size_t(nodes[i]) * ncomp + d] -
                         solution[size_t(nodes[i]) * ncomp + d];
        h->rhs[size_t(d) * h->totalRows + (hid - g->iLower)] = v;
        h->absRhs[size_t(d) * h->totalRows + (hid - g->iLower)] = std::fabs(v);
      }
    } else {
      for (int d = 0; d < g->numDof; ++d) {
        const int64_t lid = hid * g->numDof + d;
        if (lid < g->iLower || lid > g->iUpper)
          continue;
        const int64_t matIndex = g->rowStartOwned[lid - g->iLower];
        h->values[matIndex] = 1.0;
        h->absValues[matIndex] = 1.0;
        const double v = bc_values[size_t(nodes[i]) * ncomp + d] -
                         solution[size_t(nodes[i]) * ncomp + d];
        h->rhs[lid - g->iLower] = v;
        h->absRhs[lid - g->iLower] = std::fabs(v);
      }
    }
  }
}

extern "C" void
orc_applier_hypre_reset(orc_applier* a)
{
  static_cast<HypreApplier*>(a)->reset();
}

/* the |contribution| sums are the tests' tolerance scale; a timing run of the
 * restated reference path switches them off */
extern "C" void
orc_applier_hypre_track_abs(orc_applier* a, int on)
{
  static_cast<HypreApplier*>(a)->trackAbs = on != 0;
}

extern "C" void
orc_applier_hypre_get(const orc_applier* a, double* values, double* rhs)
{
  auto* h = static_cast<const HypreApplier*>(a);
  if (values && !h->values.empty())
    std::memcpy(values, h->values.data(), h->values.size() * sizeof(double));
  if (rhs && !h->rhs.empty())
    std::memcpy(rhs, h->rhs.data(), h->rhs.size() * sizeof(double));
}

extern "C" void
orc_applier_hypre_get_abs(const orc_applier* a, double* values, double* rhs)
{
  auto* h = static_cast<const HypreApplier*>(a);
  if (values && !h->absValues.empty())
    std::memcpy(
      values, h->absValues.data(), h->absValues.size() * sizeof(double));
  if (rhs && !h->absRhs.empty())
    std::memcpy(rhs, h->absRhs.data(), h->absRhs.size() * sizeof(double));
}

extern "C" void
orc_applier_hypre_enable_log(orc_applier* a, int64_t n_calls)
{
  auto* h = static_cast<HypreApplier*>(a);
  h->logOn = true;
  h->logCalls = n_calls;
  const int n = h->uvwDim > 0 ? 2 : 2 * h->g->numDof;
  h->logN = n;
  h->logSlots.assign(size_t(n_calls) * n * n, -1);
  h->logRhs.assign(size_t(n_calls) * n, -1);
}

extern "C" void
orc_applier_hypre_get_log(
  const orc_applier* a, int64_t* slots, int64_t* rhs_index)
{
  auto* h = static_cast<const HypreApplier*>(a);
  std::memcpy(slots, h->logSlots.data(), h->logSlots.size() * sizeof(int64_t));
  std::memcpy(
    rhs_index, h->logRhs.data(), h->logRhs.size() * sizeof(int64_t));
}

extern "C" void
orc_applier_destroy(orc_applier* a)
{
  delete a;
}

/* ------------------------------------------------------------------ */
/*  Peclet function                                                    */
/* ------------------------------------------------------------------ */

extern "C" double
orc_peclet_eval(const orc_peclet* f, double pecnum)
{
  if (f->form == ORC_PECLET_CLASSIC) {
    /* ClassicPecletFunction::execute, src/PecletFunction.C:41-45 (A_ unused) */
    const double modPeclet = f->a * pecnum;
    return modPeclet * modPeclet / (5.0 + modPeclet * modPeclet);
  }
  /* TanhFunction::execute, src/PecletFunction.C:68-71 */
  return 0.50 * (1.0 + std::tanh((pecnum - f->a) / f->b));
}

/* ------------------------------------------------------------------ */
/*  MdotEdgeAlg                                                        */
/* ------------------------------------------------------------------ */

extern "C" void
orc_mdot_edge(
  int ndim,
  int64_t n_edges,
  const int32_t* edge_nodes,
  const double* coords,
  const double* velocity,
  const double* gpdx,
  const double* density,
  const double* pressure,
  const double* udiag,
  const double* edge_area,
  double noc_fac,
  double interp_together,
  double* mdot)
{
  /* src/ngp_algorithms/MdotEdgeAlg.C:117-190 */
  const double om_interp = 1.0 - interp_together;
  ORC_EDGE_LOOP
  for (int64_t e = 0; e < n_edges; ++e) {
    double av[kMaxDim];
    for (int d = 0; d < ndim; ++d)
      av[d] = edge_area[e * ndim + d];
    const int64_t nL = edge_nodes[2 * e], nR = edge_nodes[2 * e + 1];
    const double pressureL = pressure[nL], pressureR = pressure[nR];
    const double densityL = density[nL], densityR = density[nR];
    const double udiagL = udiag[nL], udiagR = udiag[nR];
    const double projTimeScale = 0.5 * (1.0 / udiagL + 1.0 / udiagR);
    const double rhoIp = 0.5 * (densityL + densityR);
    double axdx = 0.0, asq = 0.0;
    for (int d = 0; d < ndim; ++d) {
      const double dxj = coords[nR * ndim + d] - coords[nL * ndim + d];
      asq += av[d] * av[d];
      axdx += av[d] * dxj;
    }
    const double inv_axdx = 1.0 / axdx;
    double tmdot = -projTimeScale * (pressureR - pressureL) * asq * inv_axdx;
    for (int d = 0; d < ndim; ++d) {
      const double dxj = coords[nR * ndim + d] - coords[nL * ndim + d];
      const double kxj = av[d] - asq * inv_axdx * dxj;
      const double rhoUjIp = 0.5 * (densityR * velocity[nR * ndim + d] +
                                    densityL * velocity[nL * ndim + d]);
      const double ujIp =
        0.5 * (velocity[nR * ndim + d] + velocity[nL * ndim + d]);
      const double GjIp =
        0.5 * (gpdx[nR * ndim + d] / udiagR + gpdx[nL * ndim + d] / udiagL);
      tmdot +=
        (interp_together * rhoUjIp + om_interp * rhoIp * ujIp + GjIp) * av[d] -
        kxj * GjIp * noc_fac;
    }
    mdot[e] = tmdot;
  }
}

/* ------------------------------------------------------------------ */
/*  MomentumEdgePecletAlg                                              */
/* ------------------------------------------------------------------ */

extern "C" void
orc_peclet_edge(
  int ndim,
  int64_t n_edges,
  const int32_t* edge_nodes,
  const double* coords,
  const double* vrtm,
  const double* density,
  const double* viscosity,
  const orc_peclet* pf,
  double eps,
  double* pecnum_out,
  double* pecfac)
{
  /* src/edge_kernels/MomentumEdgePecletAlg.C:74-101 */
  ORC_EDGE_LOOP
  for (int64_t e = 0; e < n_edges; ++e) {
    double udotx = 0.0;
    const int64_t nL = edge_nodes[2 * e], nR = edge_nodes[2 * e + 1];
    const double diffIp =
      0.5 * (viscosity[nL] / density[nL] + viscosity[nR] / density[nR]);
    for (int d = 0; d < ndim; ++d) {
      const double dxj = coords[nR * ndim + d] - coords[nL * ndim + d];
      udotx += 0.5 * dxj * (vrtm[nR * ndim + d] + vrtm[nL * ndim + d]);
    }
    const double pecnum = std::fabs(udotx) / (diffIp + eps);
    if (pecnum_out)
      pecnum_out[e] = pecnum;
    pecfac[e] = orc_peclet_eval(pf, pecnum);
  }
}

/* ------------------------------------------------------------------ */
/*  NodalGradEdgeAlg                                                   */
/* ------------------------------------------------------------------ */

extern "C" void
orc_nodal_grad_edge(
  int dim1,
  int dim2,
  int64_t n_edges,
  const int32_t* edge_nodes,
  const double* phi,
  const double* edge_area,
  const double* dual_vol,
  double* grad)
{
  /* src/ngp_algorithms/NodalGradEdgeAlg.C:85-109; the atomic adds of
   * include/ngp_utils/NgpFieldOps.h:84 become serial adds in edge order. */
  const int gsz = dim1 * dim2;
  ORC_EDGE_LOOP
  for (int64_t e = 0; e < n_edges; ++e) {
    double av[kMaxDim];
    for (int d = 0; d < dim2; ++d)
      av[d] = edge_area[e * dim2 + d];
    const int64_t nL = edge_nodes[2 * e], nR = edge_nodes[2 * e + 1];
    const double invVolL = 1.0 / dual_vol[nL];
    const double invVolR = 1.0 / dual_vol[nR];
    int counter = 0;
    for (int i = 0; i < dim1; ++i) {
      const double phiIp = 0.5 * (phi[nL * dim1 + i] + phi[nR * dim1 + i]);
      for (int j = 0; j < dim2; ++j) {
        const double ajPhiIp = av[j] * phiIp;
        add_to(grad[nL * gsz + counter], ajPhiIp * invVolL);
        add_to(grad[nR * gsz + counter], -(ajPhiIp * invVolR));
        counter++;
      }
    }
  }
}

/* ------------------------------------------------------------------ */
/*  ContinuityEdgeSolverAlg                                            */
/* ------------------------------------------------------------------ */

/* ------------------------------------------------------------------ */
/*  GeometryInteriorAlg<Hex8>                                          */
/* ------------------------------------------------------------------ */
namespace {

/* subdivide_hex_8, include/master_element/Hex8GeometryFunctions.h:258-325 */
void
geo_subdivide_hex8(const double c[8][3], double v[27][3])
{
  for (int n = 0; n < 8; ++n)
    for (int d = 0; d < 3; ++d)
      v[n][d] = c[n][d];
  for (int d = 0; d < 3; ++d) {
    v[8][d] = 0.5 * (c[0][d] + c[1][d]);
    v[9][d] = 0.5 * (c[1][d] + c[2][d]);
    v[10][d] = 0.5 * (c[2][d] + c[3][d]);
    v[11][d] = 0.5 * (c[3][d] + c[0][d]);
    v[12][d] = 0.25 * (c[0][d] + c[1][d] + c[2][d] + c[3][d]);
    v[13][d] = 0.5 * (c[4][d] + c[5][d]);
    v[14][d] = 0.5 * (c[5][d] + c[6][d]);
    v[15][d] = 0.5 * (c[6][d] + c[7][d]);
    v[16][d] = 0.5 * (c[7][d] + c[4][d]);
    v[17][d] = 0.25 * (c[4][d] + c[5][d] + c[6][d] + c[7][d]);
    v[18][d] = 0.5 * (c[1][d] + c[5][d]);
    v[19][d] = 0.5 * (c[0][d] + c[4][d]);
    v[20][d] = 0.25 * (c[0][d] + c[1][d] + c[4][d] + c[5][d]);
    v[21][d] = 0.5 * (c[3][d] + c[7][d]);
    v[22][d] = 0.5 * (c[2][d] + c[6][d]);
    v[23][d] = 0.25 * (c[2][d] + c[3][d] + c[6][d] + c[7][d]);
    v[24][d] = 0.25 * (c[1][d] + c[2][d] + c[5][d] + c[6][d]);
    v[25][d] = 0.25 * (c[0][d] + c[3][d] + c[4][d] + c[7][d]);
    v[26][d] = 0.;
    for (int n = 0; n < 8; ++n)
      v[26][d] += c[n][d];
    v[26][d] *= 0.125;
  }
}

/* hex_volume_grandy, Hex8GeometryFunctions.h:83-158 */
double
geo_hex_volume_grandy(const double sc[8][3])
{
  double cv[14][3];
  for (int n = 0; n < 8; ++n)
    for (int d = 0; d < 3; ++d)
      cv[n][d] = sc[n][d];
  static const int face_nodes[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4},
                                       {2, 3, 7, 6}, {1, 2, 6, 5}, {0, 4, 3, 7}};
  for (int k = 0; k < 6; ++k)
    for (int d = 0; d < 3; ++d)
      cv[k + 8][d] =
        0.25 * (cv[face_nodes[k][0]][d] + cv[face_nodes[k][1]][d] +
                cv[face_nodes[k][2]][d] + cv[face_nodes[k][3]][d]);
  static const int tf[24][3] = {
    {0, 8, 1},  {8, 2, 1},  {3, 2, 8},  {3, 8, 0},  {6, 9, 5},  {7, 9, 6},
    {4, 9, 7},  {4, 5, 9},  {10, 0, 1}, {5, 10, 1}, {4, 10, 5}, {4, 0, 10},
    {7, 6, 11}, {6, 2, 11}, {2, 3, 11}, {3, 7, 11}, {6, 12, 2}, {5, 12, 6},
    {5, 1, 12}, {1, 2, 12}, {0, 4, 13}, {4, 7, 13}, {7, 3, 13}, {3, 0, 13}};
  double volume = 0.0;
  for (int k = 0; k < 24; ++k) {
    const int p = tf[k][0], q = tf[k][1], r = tf[k][2];
    const double mid[3] = {cv[p][0] + cv[q][0] + cv[r][0],
                           cv[p][1] + cv[q][1] + cv[r][1],
                           cv[p][2] + cv[q][2] + cv[r][2]};
    double dxv[3];
    dxv[0] = (cv[q][1] - cv[p][1]) * (cv[r][2] - cv[p][2]) -
             (cv[r][1] - cv[p][1]) * (cv[q][2] - cv[p][2]);
    dxv[1] = (cv[r][0] - cv[p][0]) * (cv[q][2] - cv[p][2]) -
             (cv[q][0] - cv[p][0]) * (cv[r][2] - cv[p][2]);
    dxv[2] = (cv[q][0] - cv[p][0]) * (cv[r][1] - cv[p][1]) -
             (cv[r][0] - cv[p][0]) * (cv[q][1] - cv[p][1]);
    volume += mid[0] * dxv[0] + mid[1] * dxv[1] + mid[2] * dxv[2];
  }
  volume /= 18.0;
  return volume;
}

/* quad_area_by_triangulation, Hex8GeometryFunctions.h:33-81 */
void
geo_quad_area(const double ac[4][3], double* area)
{
  area[0] = area[1] = area[2] = 0.0;
  const double xmid[3] = {0.25 * (ac[0][0] + ac[1][0] + ac[2][0] + ac[3][0]),
                          0.25 * (ac[0][1] + ac[1][1] + ac[2][1] + ac[3][1]),
                          0.25 * (ac[0][2] + ac[1][2] + ac[2][2] + ac[3][2])};
  double r1[3] = {ac[0][0] - xmid[0], ac[0][1] - xmid[1], ac[0][2] - xmid[2]};
  for (int it = 0; it < 4; ++it) {
    const int t = (it + 1) % 4;
    const double r2[3] = {ac[t][0] - xmid[0], ac[t][1] - xmid[1],
                          ac[t][2] - xmid[2]};
    area[0] += r1[1] * r2[2] - r2[1] * r1[2];
    area[1] += r1[2] * r2[0] - r2[2] * r1[0];
    area[2] += r1[0] * r2[1] - r2[0] * r1[1];
    r1[0] = r2[0];
    r1[1] = r2[1];
    r1[2] = r2[2];
  }
  area[0] *= 0.5;
  area[1] *= 0.5;
  area[2] *= 0.5;
}

} // namespace

/* GeometryInteriorAlg<AlgTraitsHex8>: impl_compute_dual_nodal_volume
 * (src/ngp_algorithms/GeometryInteriorAlg.C:72-112) and
 * impl_compute_edge_area_vector (:165-225), with HexSCV::determinant_scv
 * (src/master_element/Hex8CVFEM.C:365-390) and HexSCS::determinant_scs
 * (:567-592).  elem_nodes: [n_elems][8] local nodes in the topology's node
 * order; elem_owned (may be NULL): volumes are accumulated from locally-owned
 * elements only (the reference's selector; the shared-node sum follows), edge
 * area vectors from every element given, into the edges of `edge_nodes` (an
 * element edge that is not in the list is skipped).  Accumulates (the driver's
 * pre_work zero-fill is the caller's). */
extern "C" void
orc_geometry_interior_hex8(
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  const double* coords, int64_t n_edges, const int32_t* edge_nodes,
  double* dual_nodal_volume, double* elem_volume, double* edge_area)
{
  static const int subDivisionTable[8][8] = {
    {0, 8, 12, 11, 19, 20, 26, 25},  {8, 1, 9, 12, 20, 18, 24, 26},
    {12, 9, 2, 10, 26, 24, 22, 23},  {11, 12, 10, 3, 25, 26, 23, 21},
    {19, 20, 26, 25, 4, 13, 17, 16}, {20, 18, 24, 26, 13, 5, 14, 17},
    {26, 24, 22, 23, 17, 14, 6, 15}, {25, 26, 23, 21, 16, 17, 15, 7}};
  static const int hex_edge_facet_table[12][4] = {
    {20, 8, 12, 26},  {24, 9, 12, 26},  {10, 12, 26, 23}, {11, 25, 26, 12},
    {13, 20, 26, 17}, {17, 14, 24, 26}, {17, 15, 23, 26}, {16, 17, 26, 25},
    {19, 20, 26, 25}, {20, 18, 24, 26}, {22, 23, 26, 24}, {21, 25, 26, 23}};
  static const int lrscv[24] = {0, 1, 1, 2, 2, 3, 0, 3, 4, 5, 5, 6,
                                6, 7, 4, 7, 0, 4, 1, 5, 2, 6, 3, 7};
  static const int ipNodeMap[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  std::map<std::pair<int32_t, int32_t>, int64_t> edgeOf;
  for (int64_t e = 0; e < n_edges; ++e) {
    const int32_t a = edge_nodes[2 * e], b = edge_nodes[2 * e + 1];
    edgeOf[{std::min(a, b), std::max(a, b)}] = e;
  }
  for (int64_t el = 0; el < n_elems; ++el) {
    const int32_t* en = elem_nodes + 8 * el;
    double c[8][3], v[27][3];
    for (int n = 0; n < 8; ++n)
      for (int d = 0; d < 3; ++d)
        c[n][d] = coords[size_t(en[n]) * 3 + d];
    geo_subdivide_hex8(c, v);
    if (!elem_owned || elem_owned[el]) {
      double ev = 0.0;
      for (int ip = 0; ip < 8; ++ip) {
        double sc[8][3];
        for (int n = 0; n < 8; ++n)
          for (int d = 0; d < 3; ++d)
            sc[n][d] = v[subDivisionTable[ip][n]][d];
        const double vol = geo_hex_volume_grandy(sc);
        dual_nodal_volume[en[ipNodeMap[ip]]] += vol;
        ev += vol;
      }
      if (elem_volume)
        elem_volume[el] = ev;
    }
    if (!edge_area)
      continue;
    for (int ip = 0; ip < 12; ++ip) {
      double sc[4][3], av[3];
      for (int n = 0; n < 4; ++n)
        for (int d = 0; d < 3; ++d)
          sc[n][d] = v[hex_edge_facet_table[ip][n]][d];
      geo_quad_area(sc, av);
      /* scsIpEdgeOrd is the identity for Hex8: edge `ip` joins lrscv pair ip */
      const int32_t nl = en[lrscv[2 * ip]], nr = en[lrscv[2 * ip + 1]];
      auto it = edgeOf.find({std::min(nl, nr), std::max(nl, nr)});
      if (it == edgeOf.end())
        continue;
      const int64_t e = it->second;
      const double sign = (nl == edge_nodes[2 * e]) ? 1.0 : -1.0;
      for (int d = 0; d < 3; ++d)
        edge_area[e * 3 + d] += av[d] * sign;
    }
  }
}

/* GeometryInteriorAlg<AlgTraitsQuad4_2D>: Quad42DSCV::determinant_scv
 * (src/master_element/Quad42DCVFEM.C:139-200) and Quad42DSCS::determinant_scs
 * (:384-445), lrscv / scsIpEdgeOrd include/master_element/Quad42DCVFEM.h:250-253.
 * Same conventions as orc_geometry_interior_hex8; coords / edge_area have 2
 * components. */
extern "C" void
orc_geometry_interior_quad4(
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  const double* coords, int64_t n_edges, const int32_t* edge_nodes,
  double* dual_nodal_volume, double* elem_volume, double* edge_area)
{
  static const int lrscv[8] = {0, 1, 1, 2, 2, 3, 0, 3};
  std::map<std::pair<int32_t, int32_t>, int64_t> edgeOf;
  for (int64_t e = 0; e < n_edges; ++e) {
    const int32_t a = edge_nodes[2 * e], b = edge_nodes[2 * e + 1];
    edgeOf[{std::min(a, b), std::max(a, b)}] = e;
  }
  for (int64_t el = 0; el < n_elems; ++el) {
    const int32_t* en = elem_nodes + 4 * el;
    double c[4][2];
    for (int n = 0; n < 4; ++n)
      for (int d = 0; d < 2; ++d)
        c[n][d] = coords[size_t(en[n]) * 2 + d];
    if (!elem_owned || elem_owned[el]) {
      const double gpp = 0.144337567, gpm = -0.144337567;
      const double cvm = -0.25, cvp = 0.25;
      const double half = 0.5, zero = 0.0, one16th = 0.0625;
      const double xi[2][4] = {{cvm, cvp, cvp, cvm}, {cvm, cvm, cvp, cvp}};
      const double xigp[2][4] = {{gpm, gpp, gpp, gpm}, {gpm, gpm, gpp, gpp}};
      double ev = 0.0;
      for (int ki = 0; ki < 4; ++ki) {
        double vol = zero;
        for (int kq = 0; kq < 4; ++kq) {
          double dx_ds1 = zero, dx_ds2 = zero, dy_ds1 = zero, dy_ds2 = zero;
          const double ximod = xi[0][ki] + xigp[0][kq];
          const double etamod = xi[1][ki] + xigp[1][kq];
          double deriv[2][4];
          deriv[0][0] = -(half - etamod);
          deriv[0][1] = (half - etamod);
          deriv[0][2] = (half + etamod);
          deriv[0][3] = -(half + etamod);
          deriv[1][0] = -(half - ximod);
          deriv[1][1] = -(half + ximod);
          deriv[1][2] = (half + ximod);
          deriv[1][3] = (half - ximod);
          for (int kn = 0; kn < 4; ++kn) {
            dx_ds1 += deriv[0][kn] * c[kn][0];
            dx_ds2 += deriv[1][kn] * c[kn][0];
            dy_ds1 += deriv[0][kn] * c[kn][1];
            dy_ds2 += deriv[1][kn] * c[kn][1];
          }
          const double det_j = (dx_ds1 * dy_ds2 - dy_ds1 * dx_ds2);
          vol += det_j * one16th;
        }
        dual_nodal_volume[en[ki]] += vol; /* ipNodeMap is the identity */
        ev += vol;
      }
      if (elem_volume)
        elem_volume[el] = ev;
    }
    if (!edge_area)
      continue;
    const double x1 = (c[0][0] + c[1][0] + c[2][0] + c[3][0]) * 0.25;
    const double y1 = (c[0][1] + c[1][1] + c[2][1] + c[3][1]) * 0.25;
    double areav[4][2];
    for (int f = 0; f < 4; ++f) {
      const int a = f, b = (f + 1) % 4;
      /* mid-face f joins nodes (f, f+1); face 3 is written (3, 0) */
      const double x2 = (c[a][0] + c[b][0]) * 0.5, y2 = (c[a][1] + c[b][1]) * 0.5;
      const double rr = 1.0;
      if (f < 3) {
        areav[f][0] = -(y2 - y1) * rr;
        areav[f][1] = (x2 - x1) * rr;
      } else {
        areav[f][0] = (y2 - y1) * rr;
        areav[f][1] = -(x2 - x1) * rr;
      }
    }
    for (int ip = 0; ip < 4; ++ip) {
      const int32_t nl = en[lrscv[2 * ip]], nr = en[lrscv[2 * ip + 1]];
      auto it = edgeOf.find({std::min(nl, nr), std::max(nl, nr)});
      if (it == edgeOf.end())
        continue;
      const int64_t e = it->second;
      const double sign = (nl == edge_nodes[2 * e]) ? 1.0 : -1.0;
      for (int d = 0; d < 2; ++d)
        edge_area[e * 2 + d] += areav[ip][d] * sign;
    }
  }
}

/* ------------------------------------------------------------------ */
/*  GeometryInteriorAlg<Tet4 / Wed6 / Pyr5>                            */
/* ------------------------------------------------------------------ */
namespace {

/* bhex_volume_grandy, Hex8GeometryFunctions.h:161-252: Grandy's formula with
 * the top face (nodes 4..7) split along its 5-7 diagonal */
double
geo_bhex_volume_grandy(const double sc[8][3])
{
  double cv[14][3];
  for (int n = 0; n < 8; ++n)
    for (int d = 0; d < 3; ++d)
      cv[n][d] = sc[n][d];
  for (int d = 0; d < 3; ++d) {
    cv[8][d] = 0.25 * (sc[0][d] + sc[1][d] + sc[2][d] + sc[3][d]);
    cv[9][d] = 0.5 * (sc[5][d] + sc[7][d]);
    cv[10][d] = 0.25 * (sc[0][d] + sc[1][d] + sc[5][d] + sc[4][d]);
    cv[11][d] = 0.25 * (sc[3][d] + sc[2][d] + sc[6][d] + sc[7][d]);
    cv[12][d] = 0.25 * (sc[1][d] + sc[2][d] + sc[6][d] + sc[5][d]);
    cv[13][d] = 0.25 * (sc[0][d] + sc[3][d] + sc[7][d] + sc[4][d]);
  }
  static const int tf[24][3] = {
    {0, 8, 1},  {8, 2, 1},  {3, 2, 8},  {3, 8, 0},  {6, 9, 5},  {7, 9, 6},
    {4, 9, 7},  {4, 5, 9},  {10, 0, 1}, {5, 10, 1}, {4, 10, 5}, {4, 0, 10},
    {7, 6, 11}, {6, 2, 11}, {2, 3, 11}, {3, 7, 11}, {6, 12, 2}, {5, 12, 6},
    {5, 1, 12}, {1, 2, 12}, {0, 4, 13}, {4, 7, 13}, {7, 3, 13}, {3, 0, 13}};
  double volume = 0.0;
  for (int k = 0; k < 24; ++k) {
    const int p = tf[k][0], q = tf[k][1], r = tf[k][2];
    const double mid[3] = {cv[p][0] + cv[q][0] + cv[r][0],
                           cv[p][1] + cv[q][1] + cv[r][1],
                           cv[p][2] + cv[q][2] + cv[r][2]};
    double dxv[3];
    dxv[0] = (cv[q][1] - cv[p][1]) * (cv[r][2] - cv[p][2]) -
             (cv[r][1] - cv[p][1]) * (cv[q][2] - cv[p][2]);
    dxv[1] = (cv[r][0] - cv[p][0]) * (cv[q][2] - cv[p][2]) -
             (cv[q][0] - cv[p][0]) * (cv[r][2] - cv[p][2]);
    dxv[2] = (cv[q][0] - cv[p][0]) * (cv[r][1] - cv[p][1]) -
             (cv[r][0] - cv[p][0]) * (cv[q][1] - cv[p][1]);
    volume += mid[0] * dxv[0] + mid[1] * dxv[1] + mid[2] * dxv[2];
  }
  volume /= 18.0;
  return volume;
}

/* octohedron_volume_by_triangle_facets + polyhedral_volume_by_faces,
 * src/master_element/Pyr5CVFEM.C:348-432 (the apex control volume) */
double
geo_octohedron_volume(const double vc[10][3])
{
  double c[14][3];
  for (int j = 0; j < 10; ++j)
    for (int k = 0; k < 3; ++k)
      c[j][k] = vc[j][k];
  for (int k = 0; k < 3; ++k) {
    c[10][k] = 0.50 * (vc[3][k] + vc[9][k]);
    c[11][k] = 0.50 * (vc[3][k] + vc[5][k]);
    c[12][k] = 0.50 * (vc[5][k] + vc[7][k]);
    c[13][k] = 0.50 * (vc[7][k] + vc[9][k]);
  }
  static const int tf[24][3] = {
    {1, 3, 10}, {2, 10, 3}, {2, 9, 10}, {10, 9, 1}, {4, 3, 11}, {3, 1, 11},
    {11, 1, 5}, {4, 11, 5}, {1, 12, 5}, {1, 7, 12}, {12, 7, 6}, {5, 12, 6},
    {9, 8, 13}, {13, 8, 7}, {13, 7, 1}, {9, 13, 1}, {4, 5, 0},  {5, 6, 0},
    {6, 7, 0},  {7, 8, 0},  {0, 8, 9},  {0, 9, 2},  {0, 2, 3},  {0, 3, 4}};
  double volume = 0.0;
  for (int t = 0; t < 24; ++t) {
    const int ip = tf[t][0], iq = tf[t][1], ir = tf[t][2];
    double xf[3];
    for (int k = 0; k < 3; ++k)
      xf[k] = c[ip][k] + c[iq][k] + c[ir][k];
    volume = volume +
             xf[0] * ((c[iq][1] - c[ip][1]) * (c[ir][2] - c[ip][2]) -
                      (c[ir][1] - c[ip][1]) * (c[iq][2] - c[ip][2])) -
             xf[1] * ((c[iq][0] - c[ip][0]) * (c[ir][2] - c[ip][2]) -
                      (c[ir][0] - c[ip][0]) * (c[iq][2] - c[ip][2])) +
             xf[2] * ((c[iq][0] - c[ip][0]) * (c[ir][1] - c[ip][1]) -
                      (c[ir][0] - c[ip][0]) * (c[iq][1] - c[ip][1]));
  }
  volume = volume / 18.0;
  return volume;
}

/* the 15 points of TetSCV/TetSCS::determinant_*,
 * src/master_element/Tet4CVFEM.C:255-322, 536-603 */
void
geo_subdivide_tet4(const double c[][3], double v[][3])
{
  const double half = 0.5, one3rd = 1.0 / 3.0;
  for (int j = 0; j < 4; ++j)
    for (int k = 0; k < 3; ++k)
      v[j][k] = c[j][k];
  for (int k = 0; k < 3; ++k) {
    v[4][k] = half * (c[0][k] + c[1][k]);
    v[5][k] = half * (c[1][k] + c[2][k]);
    v[6][k] = half * (c[2][k] + c[0][k]);
    v[7][k] = one3rd * (c[0][k] + c[1][k] + c[2][k]);
    v[8][k] = half * (c[2][k] + c[3][k]);
    v[9][k] = half * (c[3][k] + c[1][k]);
    v[10][k] = one3rd * (c[1][k] + c[2][k] + c[3][k]);
    v[11][k] = half * (c[0][k] + c[3][k]);
    v[12][k] = one3rd * (c[0][k] + c[2][k] + c[3][k]);
    v[13][k] = one3rd * (c[0][k] + c[1][k] + c[3][k]);
    v[14][k] = 0.0;
    for (int j = 0; j < 4; ++j)
      v[14][k] = v[14][k] + 0.25 * c[j][k];
  }
}

/* the 21 points of WedSCV/WedSCS::determinant_*,
 * src/master_element/Wed6CVFEM.C:286-358, 565-636 */
void
geo_subdivide_wed6(const double c[][3], double v[][3])
{
  const double half = 0.5, one3rd = 1.0 / 3.0, one6th = 1.0 / 6.0;
  for (int j = 0; j < 6; ++j)
    for (int k = 0; k < 3; ++k)
      v[j][k] = c[j][k];
  for (int k = 0; k < 3; ++k) {
    v[6][k] = half * (c[0][k] + c[1][k]);
    v[7][k] = half * (c[1][k] + c[2][k]);
    v[8][k] = half * (c[2][k] + c[0][k]);
    v[9][k] = one3rd * (c[0][k] + c[1][k] + c[2][k]);
    v[10][k] = half * (c[3][k] + c[4][k]);
    v[11][k] = half * (c[4][k] + c[5][k]);
    v[12][k] = half * (c[5][k] + c[3][k]);
    v[13][k] = one3rd * (c[3][k] + c[4][k] + c[5][k]);
    v[14][k] = half * (c[1][k] + c[4][k]);
    v[15][k] = half * (c[0][k] + c[3][k]);
    v[16][k] = 0.25 * (c[0][k] + c[1][k] + c[4][k] + c[3][k]);
    v[17][k] = half * (c[2][k] + c[5][k]);
    v[18][k] = 0.25 * (c[1][k] + c[4][k] + c[5][k] + c[2][k]);
    v[19][k] = 0.25 * (c[5][k] + c[3][k] + c[0][k] + c[2][k]);
    v[20][k] = 0.0;
    for (int j = 0; j < 6; ++j)
      v[20][k] += one6th * c[j][k];
  }
}

/* the 19 points of PyrSCV/PyrSCS::determinant_*,
 * src/master_element/Pyr5CVFEM.C:459-543, 798-883 */
void
geo_subdivide_pyr5(const double c[][3], double v[][3])
{
  const double one3rd = 1.0 / 3.0;
  for (int j = 0; j < 5; ++j)
    for (int k = 0; k < 3; ++k)
      v[j][k] = c[j][k];
  for (int k = 0; k < 3; ++k) {
    v[5][k] = 0.5 * (c[0][k] + c[1][k]);
    v[6][k] = 0.5 * (c[1][k] + c[2][k]);
    v[7][k] = 0.5 * (c[2][k] + c[3][k]);
    v[8][k] = 0.5 * (c[3][k] + c[0][k]);
    v[9][k] = 0.25 * (c[0][k] + c[1][k] + c[2][k] + c[3][k]);
    v[10][k] = 0.5 * (c[1][k] + c[4][k]);
    v[11][k] = 0.5 * (c[4][k] + c[0][k]);
    v[12][k] = one3rd * (c[0][k] + c[1][k] + c[4][k]);
    v[13][k] = 0.5 * (c[2][k] + c[4][k]);
    v[14][k] = one3rd * (c[1][k] + c[2][k] + c[4][k]);
    v[15][k] = 0.5 * (c[3][k] + c[4][k]);
    v[16][k] = one3rd * (c[3][k] + c[4][k] + c[2][k]);
    v[17][k] = one3rd * (c[0][k] + c[4][k] + c[3][k]);
    v[18][k] = 0.0;
    for (int j = 0; j < 5; ++j)
      v[18][k] += 0.2 * c[j][k];
  }
}

struct GeoTopo3
{
  int npe, nScv, nScs;
  void (*subdivide)(const double c[][3], double v[][3]);
  int scvHex[6][10]; /* sub-control volume -> sub-points (8, apex of Pyr5: 10) */
  int scsQuad[12][4];
  int lrscv[24];
};

/* tables: Tet4CVFEM.C:247-251, 526-528 + Tet4CVFEM.h:266;
 * Wed6CVFEM.C:271-275, 545-555 + Wed6CVFEM.h:268;
 * Pyr5CVFEM.C:448-453, 776-789 + Pyr5CVFEM.h:300-301 */
const GeoTopo3 kGeoTet4 = {
  4, 4, 6, geo_subdivide_tet4,
  {{0, 4, 7, 6, 11, 13, 14, 12}, {1, 5, 7, 4, 9, 10, 14, 13},
   {2, 6, 7, 5, 8, 12, 14, 10}, {3, 9, 13, 11, 8, 10, 14, 12}},
  {{4, 7, 14, 13}, {7, 14, 10, 5}, {6, 12, 14, 7}, {11, 13, 14, 12},
   {13, 9, 10, 14}, {10, 8, 12, 14}},
  {0, 1, 1, 2, 0, 2, 0, 3, 1, 3, 2, 3}};
const GeoTopo3 kGeoWed6 = {
  6, 6, 9, geo_subdivide_wed6,
  {{0, 15, 16, 6, 8, 19, 20, 9}, {9, 6, 1, 7, 20, 16, 14, 18},
   {8, 9, 7, 2, 19, 20, 18, 17}, {19, 15, 16, 20, 12, 3, 10, 13},
   {20, 16, 14, 18, 13, 10, 4, 11}, {19, 20, 18, 17, 12, 13, 11, 5}},
  {{6, 9, 20, 16}, {7, 9, 20, 18}, {9, 8, 19, 20}, {10, 16, 20, 13},
   {13, 11, 18, 20}, {12, 13, 20, 19}, {15, 16, 20, 19}, {16, 14, 18, 20},
   {19, 20, 18, 17}},
  {0, 1, 1, 2, 0, 2, 3, 4, 4, 5, 3, 5, 0, 3, 1, 4, 2, 5}};
const GeoTopo3 kGeoPyr5 = {
  5, 5, 12, geo_subdivide_pyr5,
  {{0, 5, 9, 8, 11, 12, 18, 17, -1, -1}, {1, 6, 9, 5, 10, 14, 18, 12, -1, -1},
   {2, 7, 9, 6, 13, 16, 18, 14, -1, -1}, {3, 8, 9, 7, 15, 17, 18, 16, -1, -1},
   {4, 18, 15, 17, 11, 12, 10, 14, 13, 16}},
  {{5, 9, 18, 12}, {6, 9, 18, 14}, {7, 9, 18, 16}, {8, 17, 18, 9},
   {12, 12, 18, 17}, {11, 12, 12, 17}, {14, 14, 18, 12}, {10, 14, 14, 12},
   {16, 16, 18, 14}, {13, 16, 16, 14}, {17, 17, 18, 16}, {15, 17, 17, 16}},
  {0, 1, 1, 2, 2, 3, 0, 3, 0, 4, 0, 4, 1, 4, 1, 4, 2, 4, 2, 4, 3, 4, 3, 4}};

void
geo_interior_3d(
  const GeoTopo3& T, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, const double* coords, int64_t n_edges,
  const int32_t* edge_nodes, double* dual_nodal_volume, double* elem_volume,
  double* edge_area)
{
  std::map<std::pair<int32_t, int32_t>, int64_t> edgeOf;
  for (int64_t e = 0; e < n_edges; ++e) {
    const int32_t a = edge_nodes[2 * e], b = edge_nodes[2 * e + 1];
    edgeOf[{std::min(a, b), std::max(a, b)}] = e;
  }
  for (int64_t el = 0; el < n_elems; ++el) {
    const int32_t* en = elem_nodes + T.npe * el;
    double c[6][3], v[21][3];
    for (int n = 0; n < T.npe; ++n)
      for (int d = 0; d < 3; ++d)
        c[n][d] = coords[size_t(en[n]) * 3 + d];
    T.subdivide(c, v);
    if (!elem_owned || elem_owned[el]) {
      double ev = 0.0;
      for (int ip = 0; ip < T.nScv; ++ip) {
        double vol;
        if (T.npe == 5 && ip == 4) {
          double oc[10][3];
          for (int n = 0; n < 10; ++n)
            for (int d = 0; d < 3; ++d)
              oc[n][d] = v[T.scvHex[ip][n]][d];
          vol = geo_octohedron_volume(oc);
        } else {
          double sc[8][3];
          for (int n = 0; n < 8; ++n)
            for (int d = 0; d < 3; ++d)
              sc[n][d] = v[T.scvHex[ip][n]][d];
          /* the pyramid's base volumes have a bent top face */
          vol = T.npe == 5 ? geo_bhex_volume_grandy(sc) : geo_hex_volume_grandy(sc);
        }
        dual_nodal_volume[en[ip]] += vol; /* ipNodeMap is the identity */
        ev += vol;
      }
      if (elem_volume)
        elem_volume[el] = ev;
    }
    if (!edge_area)
      continue;
    for (int ip = 0; ip < T.nScs; ++ip) {
      double sc[4][3], av[3];
      for (int n = 0; n < 4; ++n)
        for (int d = 0; d < 3; ++d)
          sc[n][d] = v[T.scsQuad[ip][n]][d];
      geo_quad_area(sc, av);
      /* scsIpEdgeOrd (Pyr5: {0,1,2,3,4,4,5,5,6,6,7,7}) names the element edge
       * that joins the ip's lrscv pair: look it up by its two nodes */
      const int32_t nl = en[T.lrscv[2 * ip]], nr = en[T.lrscv[2 * ip + 1]];
      auto it = edgeOf.find({std::min(nl, nr), std::max(nl, nr)});
      if (it == edgeOf.end())
        continue;
      const int64_t e = it->second;
      const double sign = (nl == edge_nodes[2 * e]) ? 1.0 : -1.0;
      for (int d = 0; d < 3; ++d)
        edge_area[e * 3 + d] += av[d] * sign;
    }
  }
}

} // namespace

/* GeometryInteriorAlg<AlgTraitsTet4 / Wed6 / Pyr5>
 * (src/ngp_algorithms/GeometryInteriorAlg.C:72-112, 165-225) with
 * TetSCV/TetSCS (src/master_element/Tet4CVFEM.C:243-343, 522-619),
 * WedSCV/WedSCS (Wed6CVFEM.C:267-369, 541-647) and PyrSCV/PyrSCS
 * (Pyr5CVFEM.C:438-572, 772-900).  Conventions of orc_geometry_interior_hex8.
 * No in-tree known-answer test holds values for these three topologies: the
 * restatement is pinned by properties (tests/test_geometry_topologies_cpu.py). */
extern "C" void
orc_geometry_interior_tet4(
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  const double* coords, int64_t n_edges, const int32_t* edge_nodes,
  double* dual_nodal_volume, double* elem_volume, double* edge_area)
{
  geo_interior_3d(kGeoTet4, n_elems, elem_nodes, elem_owned, coords, n_edges,
                  edge_nodes, dual_nodal_volume, elem_volume, edge_area);
}

extern "C" void
orc_geometry_interior_wed6(
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  const double* coords, int64_t n_edges, const int32_t* edge_nodes,
  double* dual_nodal_volume, double* elem_volume, double* edge_area)
{
  geo_interior_3d(kGeoWed6, n_elems, elem_nodes, elem_owned, coords, n_edges,
                  edge_nodes, dual_nodal_volume, elem_volume, edge_area);
}

extern "C" void
orc_geometry_interior_pyr5(
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  const double* coords, int64_t n_edges, const int32_t* edge_nodes,
  double* dual_nodal_volume, double* elem_volume, double* edge_area)
{
  geo_interior_3d(kGeoPyr5, n_elems, elem_nodes, elem_owned, coords, n_edges,
                  edge_nodes, dual_nodal_volume, elem_volume, edge_area);
}

/* MdotEdgeAlg / ContinuityEdgeSolverAlg with the optional terms:
 * balanced buoyancy forcing (use_balanced_buoyancy_force; gravity, fields
 * buoyancy_source / buoyancy_source_mask) and the GCL term of deforming meshes
 * (edge_face_velocity_mag).  src/ngp_algorithms/MdotEdgeAlg.C:117-190 (terms
 * :153-163, :175-180); src/edge_kernels/ContinuityEdgeSolverAlg.C:109-194
 * (terms :147-158, :172-177).  sink == NULL: MdotEdgeAlg (writes mdot);
 * else ContinuityEdgeSolverAlg (scaled tmdot, 2x2 block into the sink). */
extern "C" void
orc_mdot_continuity_edge_ext(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* velocity, const double* gpdx, const double* density,
  const double* pressure, const double* udiag, const double* edge_area,
  double noc_fac, double interp_together, int add_balanced_forcing,
  const double* gravity, const double* source, const double* source_mask,
  int needs_gcl, const double* edge_face_vel_mag,
  /* continuity only: */ double dt, double gamma1, double solve_incompressible,
  double* mdot, orc_applier* sink)
{
  const double om_interp = 1.0 - interp_together;
  const double tauScale = dt / gamma1;
  const double om_solveInc = 1.0 - solve_incompressible;
  for (int64_t e = 0; e < n_edges; ++e) {
    double av[kMaxDim];
    for (int d = 0; d < ndim; ++d)
      av[d] = edge_area[e * ndim + d];
    const int32_t nodes[2] = {edge_nodes[2 * e], edge_nodes[2 * e + 1]};
    const int64_t nL = nodes[0], nR = nodes[1];
    const double pressureL = pressure[nL], pressureR = pressure[nR];
    const double densityL = density[nL], densityR = density[nR];
    const double udiagL = udiag[nL], udiagR = udiag[nR];
    const double projTimeScale = 0.5 * (1.0 / udiagL + 1.0 / udiagR);
    const double rhoIp = 0.5 * (densityL + densityR);
    const double denScale = (1.0 / rhoIp) * solve_incompressible + om_solveInc;
    double axdx = 0.0, asq = 0.0;
    for (int d = 0; d < ndim; ++d) {
      const double dxj = coords[nR * ndim + d] - coords[nL * ndim + d];
      asq += av[d] * av[d];
      axdx += av[d] * dxj;
    }
    const double inv_axdx = 1.0 / axdx;
    double tmdot = -projTimeScale * (pressureR - pressureL) * asq * inv_axdx;
    if (add_balanced_forcing) {
      const double masked_weights = 0.5 * (source_mask[nL] + source_mask[nR]);
      for (int d = 0; d < ndim; ++d)
        tmdot += projTimeScale * av[d] * gravity[d] * rhoIp * masked_weights;
    }
    if (needs_gcl)
      tmdot -= rhoIp * edge_face_vel_mag[e];
    for (int d = 0; d < ndim; ++d) {
      const double dxj = coords[nR * ndim + d] - coords[nL * ndim + d];
      const double kxj = av[d] - asq * inv_axdx * dxj;
      const double rhoUjIp = 0.5 * (densityR * velocity[nR * ndim + d] +
                                    densityL * velocity[nL * ndim + d]);
      const double ujIp =
        0.5 * (velocity[nR * ndim + d] + velocity[nL * ndim + d]);
      double GjIp =
        0.5 * (gpdx[nR * ndim + d] / udiagR + gpdx[nL * ndim + d] / udiagL);
      if (add_balanced_forcing)
        GjIp -= 0.5 * ((source_mask[nR] * source[nR * ndim + d]) / (udiagR) +
                       (source_mask[nL] * source[nL * ndim + d]) / (udiagL));
      tmdot +=
        (interp_together * rhoUjIp + om_interp * rhoIp * ujIp + GjIp) * av[d] -
        kxj * GjIp * noc_fac;
    }
    if (!sink) {
      mdot[e] = tmdot;
      continue;
    }
    tmdot /= tauScale;
    tmdot *= denScale;
    const double lhsfac = -asq * inv_axdx * projTimeScale * denScale / tauScale;
    const double lhs[4] = {-lhsfac, +lhsfac, +lhsfac, -lhsfac};
    const double rhs[2] = {-tmdot, tmdot};
    sink->apply(2, nodes, rhs, lhs, 2);
  }
}

/* WallDistEdgeSolverAlg::execute, src/edge_kernels/WallDistEdgeSolverAlg.C:28-66 */
extern "C" void
orc_wall_dist_edge(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* edge_area, orc_applier* sink)
{
  ORC_EDGE_LOOP
  for (int64_t e = 0; e < n_edges; ++e) {
    const int32_t nodes[2] = {edge_nodes[2 * e], edge_nodes[2 * e + 1]};
    const int64_t nL = nodes[0], nR = nodes[1];
    double asq = 0.0, axdx = 0.0;
    for (int d = 0; d < ndim; d++) {
      const double axj = edge_area[e * ndim + d];
      const double dxj = coords[nR * ndim + d] - coords[nL * ndim + d];
      asq += axj * axj;
      axdx += axj * dxj;
    }
    const double pfac = 1.0;
    const double lhsfac = pfac * asq / axdx;
    const double lhs[4] = {+lhsfac, -lhsfac, -lhsfac, +lhsfac};
    const double rhs[2] = {0.0, 0.0};
    sink->apply(2, nodes, rhs, lhs, 2);
  }
}

/* WallDistNodeKernel::execute, src/node_kernels/WallDistNodeKernel.C:34-43 */
extern "C" void
orc_wall_dist_node(
  int64_t n_sel, const int32_t* nodes, const double* dual_nodal_volume,
  orc_applier* a)
{
  for (int64_t i = 0; i < n_sel; ++i) {
    const int32_t n = nodes[i];
    double lhs = 0.0, rhs = 0.0;
    rhs += dual_nodal_volume[n];
    a->apply(1, &n, &rhs, &lhs, 1);
  }
}

extern "C" void
orc_continuity_edge(
  int ndim,
  int64_t n_edges,
  const int32_t* edge_nodes,
  const double* coords,
  const double* velocity,
  const double* gpdx,
  const double* density,
  const double* pressure,
  const double* udiag,
  const double* edge_area,
  const orc_continuity_opts* o,
  orc_applier* sink)
{
  /* shell: include/AssembleEdgeSolverAlgorithm.h:72-98
   * body:  src/edge_kernels/ContinuityEdgeSolverAlg.C:37-60, 109-194 */
  const double nocFac = o->noc_fac;
  const double tauScale = o->dt / o->gamma1;
  const double interpTogether = o->interp_together;
  const double om_interpTogether = 1.0 - interpTogether;
  const double solveInc = o->solve_incompressible;
  const double om_solveInc = 1.0 - solveInc;

  ORC_EDGE_LOOP
  for (int64_t e = 0; e < n_edges; ++e) {
    double lhs[4] = {0.0, 0.0, 0.0, 0.0};
    double rhs[2] = {0.0, 0.0};
    double av[kMaxDim];
    for (int d = 0; d < ndim; ++d)
      av[d] = edge_area[e * ndim + d];
    const int32_t nodes[2] = {edge_nodes[2 * e], edge_nodes[2 * e + 1]};
    const int64_t nL = nodes[0], nR = nodes[1];

    const double pressureL = pressure[nL], pressureR = pressure[nR];
    const double densityL = density[nL], densityR = density[nR];
    const double udiagL = udiag[nL], udiagR = udiag[nR];
    const double projTimeScale = 0.5 * (1.0 / udiagL + 1.0 / udiagR);
    const double rhoIp = 0.5 * (densityL + densityR);
    const double denScale = (1.0 / rhoIp) * solveInc + om_solveInc;

    double axdx = 0.0, asq = 0.0;
    for (int d = 0; d < ndim; ++d) {
      const double dxj = coords[nR * ndim + d] - coords[nL * ndim + d];
      asq += av[d] * av[d];
      axdx += av[d] * dxj;
    }
    const double inv_axdx = 1.0 / axdx;

    double tmdot = -projTimeScale * (pressureR - pressureL) * asq * inv_axdx;
    for (int d = 0; d < ndim; ++d) {
      const double dxj = coords[nR * ndim + d] - coords[nL * ndim + d];
      const double kxj = av[d] - asq * inv_axdx * dxj;
      const double rhoUjIp = 0.5 * (densityR * velocity[nR * ndim + d] +
                                    densityL * velocity[nL * ndim + d]);
      const double ujIp =
        0.5 * (velocity[nR * ndim + d] + velocity[nL * ndim + d]);
      const double GjIp = 0.5 * (gpdx[nR * ndim + d] / (udiagR) +
                                 gpdx[nL * ndim + d] / (udiagL));
      tmdot += (interpTogether * rhoUjIp + om_interpTogether * rhoIp * ujIp +
                GjIp) *
                 av[d] -
               kxj * GjIp * nocFac;
    }
    tmdot /= tauScale;
    tmdot *= denScale;
    const double lhsfac = -asq * inv_axdx * projTimeScale * denScale / tauScale;

    lhs[0] = -lhsfac;
    lhs[1] = +lhsfac;
    rhs[0] = -tmdot;
    lhs[2] = +lhsfac;
    lhs[3] = -lhsfac;
    rhs[1] = tmdot;

    sink->apply(2, nodes, rhs, lhs, 2);
  }
}

/* ------------------------------------------------------------------ */
/*  ScalarEdgeSolverAlg                                                */
/* ------------------------------------------------------------------ */

extern "C" void
orc_scalar_edge(
  int ndim,
  int64_t n_edges,
  const int32_t* edge_nodes,
  const double* coords,
  const double* vrtm,
  const double* q,
  const double* dqdx,
  const double* density,
  const double* diff_flux_coeff,
  const double* edge_area,
  const double* mdot_f,
  const orc_scalar_opts* o,
  orc_applier* sink)
{
  /* src/edge_kernels/ScalarEdgeSolverAlg.C:55-70, 85-205 */
  const double eps = o->eps;
  const double alpha = o->alpha;
  const double alphaUpw = o->alpha_upw;
  const double hoUpwind = o->ho_upwind;
  const double relaxFac = o->relax_fac;
  const bool useLimiter = o->use_limiter != 0;
  const double om_alpha = 1.0 - alpha;
  const double om_alphaUpw = 1.0 - alphaUpw;

  ORC_EDGE_LOOP
  for (int64_t e = 0; e < n_edges; ++e) {
    double lhs[4] = {0.0, 0.0, 0.0, 0.0};
    double rhs[2] = {0.0, 0.0};
    double av[kMaxDim];
    for (int d = 0; d < ndim; ++d)
      av[d] = edge_area[e * ndim + d];
    const int32_t nodes[2] = {edge_nodes[2 * e], edge_nodes[2 * e + 1]};
    const int64_t nL = nodes[0], nR = nodes[1];

    const double mdot = mdot_f[e];
    const double densityL = density[nL], densityR = density[nR];
    const double qNp1L = q[nL], qNp1R = q[nR];
    const double viscosityL = diff_flux_coeff[nL];
    const double viscosityR = diff_flux_coeff[nR];
    const double viscIp = 0.5 * (viscosityL + viscosityR);
    const double diffIp =
      0.5 * (viscosityL / densityL + viscosityR / densityR);

    double axdx = 0.0, asq = 0.0, udotx = 0.0;
    for (int d = 0; d < ndim; ++d) {
      const double dxj = coords[nR * ndim + d] - coords[nL * ndim + d];
      asq += av[d] * av[d];
      axdx += av[d] * dxj;
      udotx += 0.5 * dxj * (vrtm[nR * ndim + d] + vrtm[nL * ndim + d]);
    }
    const double inv_axdx = 1.0 / axdx;

    double dqL = 0.0, dqR = 0.0, nonOrth = 0.0;
    for (int d = 0; d < ndim; ++d) {
      const double dxj = (coords[nR * ndim + d] - coords[nL * ndim + d]);
      dqL += 0.5 * dxj * dqdx[nL * ndim + d];
      dqR += 0.5 * dxj * dqdx[nR * ndim + d];
      const double kxj = av[d] - asq * inv_axdx * dxj;
      nonOrth +=
        -viscIp * kxj * 0.5 * (dqdx[nR * ndim + d] + dqdx[nL * ndim + d]);
    }

    const double pecnum = std::fabs(udotx) / (diffIp + eps);
    const double pecfac = orc_peclet_eval(&o->pf, pecnum);
    const double om_pecfac = 1.0 - pecfac;

    double limitL = 1.0, limitR = 1.0;
    if (useLimiter) {
      const double dq = qNp1R - qNp1L;
      const double dqML = 4.0 * dqL - dq;
      const double dqMR = 4.0 * dqR - dq;
      limitL = van_leer(dqML, dq, eps);
      limitR = van_leer(dqMR, dq, eps);
    }

    const double qIpL = qNp1L + dqL * hoUpwind * limitL;
    const double qIpR = qNp1R - dqR * hoUpwind * limitR;

    const double lhsfac = -viscIp * asq * inv_axdx;
    const double diffFlux = lhsfac * (qNp1R - qNp1L) + nonOrth;

    lhs[0] = -lhsfac / relaxFac;
    lhs[1] = lhsfac;
    rhs[0] = -diffFlux;
    lhs[2] = lhsfac;
    lhs[3] = -lhsfac / relaxFac;
    rhs[1] = diffFlux;

    const double qIp = 0.5 * (qNp1R + qNp1L);
    const double qUpw = (mdot > 0) ? (alphaUpw * qIpL + om_alphaUpw * qIp)
                                   : (alphaUpw * qIpR + om_alphaUpw * qIp);
    const double qHatL = (alpha * qIpL + om_alpha * qIp);
    const double qHatR = (alpha * qIpR + om_alpha * qIp);
    const double qCds = 0.5 * (qHatL + qHatR);

    const double adv_flux = mdot * (pecfac * qUpw + om_pecfac * qCds);
    rhs[0] -= adv_flux;
    rhs[1] += adv_flux;

    double alhsfac = 0.5 * (mdot + std::fabs(mdot)) * pecfac * alphaUpw +
                     0.5 * alpha * om_pecfac * mdot;
    lhs[0] += alhsfac / relaxFac;
    lhs[2] -= alhsfac;

    alhsfac = 0.5 * (mdot - std::fabs(mdot)) * pecfac * alphaUpw +
              0.5 * alpha * om_pecfac * mdot;
    lhs[3] -= alhsfac / relaxFac;
    lhs[1] += alhsfac;

    alhsfac = 0.5 * mdot * (pecfac * om_alphaUpw + om_pecfac * om_alpha);
    lhs[0] += alhsfac / relaxFac;
    lhs[1] += alhsfac;
    lhs[2] -= alhsfac;
    lhs[3] -= alhsfac / relaxFac;

    sink->apply(2, nodes, rhs, lhs, 2);
  }
}

/* ------------------------------------------------------------------ */
/*  MomentumEdgeSolverAlg                                              */
/* ------------------------------------------------------------------ */

/* mass_vof_f: massVofBalancedFlowRate; the reference aliases it to
 * massFlowRate when realm_has_vof_ is off (:52-56) and multiplies it by
 * has_vof = 0.0 -- here a null pointer. */
static void
momentum_edge_impl(
  int ndim,
  int64_t n_edges,
  const int32_t* edge_nodes,
  const double* coords,
  const double* vel,
  const double* dudx,
  const double* viscosity,
  const double* density,
  const double* node_mask,
  const double* edge_area,
  const double* mdot_f,
  const double* mass_vof_f,
  const double* pecfac_f,
  const orc_momentum_opts* o,
  orc_applier* sink,
  double* udiag_accum)
{
  /* src/edge_kernels/MomentumEdgeSolverAlg.C:70-88, 105-312 */
  const double eps = o->eps;
  const double includeDivU = o->include_divu;
  const double alpha = o->alpha;
  const double alphaUpw_input = o->alpha_upw;
  const double hoUpwind = o->ho_upwind;
  const double relaxFacU = o->relax_fac;
  const bool useLimiter = o->use_limiter != 0;
  const double om_alpha = 1.0 - alpha;
  const double om_alphaUpw_input = 1.0 - alphaUpw_input;
  const double has_vof = mass_vof_f ? 1.0 : 0.0; /* (double)realm_has_vof_, :88 */
  const int n = 2 * ndim;

  ORC_EDGE_LOOP
  for (int64_t e = 0; e < n_edges; ++e) {
    double lhs[kMaxRhs * kMaxRhs];
    double rhs[kMaxRhs];
    for (int i = 0; i < n * n; ++i)
      lhs[i] = 0.0;
    for (int i = 0; i < n; ++i)
      rhs[i] = 0.0;
#define LHS(r, c) lhs[(r) * n + (c)]

    double av[kMaxDim];
    for (int d = 0; d < ndim; ++d)
      av[d] = edge_area[e * ndim + d];
    const int32_t nodes[2] = {edge_nodes[2 * e], edge_nodes[2 * e + 1]};
    const int64_t nL = nodes[0], nR = nodes[1];
    const double* xL = &coords[nL * ndim];
    const double* xR = &coords[nR * ndim];
    const double* uL = &vel[nL * ndim];
    const double* uR = &vel[nR * ndim];
    const double* gL = &dudx[nL * ndim * ndim];
    const double* gR = &dudx[nR * ndim * ndim];

    /* :110-111: per-edge copies the VOF branch modifies */
    double alphaUpw = alphaUpw_input;
    double om_alphaUpw = om_alphaUpw_input;
    /* :124-125 */
    const double mdot =
      mdot_f[e] + has_vof * (mass_vof_f ? mass_vof_f[e] : mdot_f[e]);
    const double densityL = density[nL], densityR = density[nR];
    const double viscosityL = viscosity[nL], viscosityR = viscosity[nR];
    const double viscIp = 0.5 * (viscosityL + viscosityR);

    double axdx = 0.0, asq = 0.0;
    for (int d = 0; d < ndim; ++d) {
      const double dxj = xR[d] - xL[d];
      asq += av[d] * av[d];
      axdx += av[d] * dxj;
    }
    const double inv_axdx = 1.0 / axdx;

    double duL[kMaxDim], duR[kMaxDim];
    for (int i = 0; i < ndim; ++i) {
      const int offset = i * ndim;
      duL[i] = 0.0;
      duR[i] = 0.0;
      for (int j = 0; j < ndim; ++j) {
        const double dxj = 0.5 * (xR[j] - xL[j]);
        duL[i] += dxj * gL[offset + j];
        duR[i] += dxj * gR[offset + j];
      }
    }

    double limitL[kMaxDim] = {1.0, 1.0, 1.0};
    double limitR[kMaxDim] = {1.0, 1.0, 1.0};
    if (useLimiter) {
      for (int d = 0; d < ndim; ++d) {
        const double du = uR[d] - uL[d];
        const double duML = 4.0 * duL[d] - du;
        const double duMR = 4.0 * duR[d] - du;
        limitL[d] = van_leer(duML, du, eps);
        limitR[d] = van_leer(duMR, du, eps);
      }
    }

    /* :174-192 upwinding switch for multiphase cases */
    double pecfac = pecfac_f[e];
    double om_pecfac = 1.0 - pecfac;
    double density_upwinding_factor = 1.0;
    if (has_vof > 0.5) {
      const double min_density = std::fmin(densityL, densityR);
      const double density_differential =
        std::fabs(densityL - densityR) / min_density;
      density_upwinding_factor = 1.0 - std::erf(6.0 * density_differential);

      alphaUpw = density_upwinding_factor * alphaUpw +
                 (1.0 - density_upwinding_factor);
      om_alphaUpw = 1.0 - alphaUpw;
      pecfac =
        1.0 - density_upwinding_factor + density_upwinding_factor * pecfac;
      om_pecfac = 1.0 - pecfac;
    }

    double uIpL[kMaxDim], uIpR[kMaxDim];
    for (int d = 0; d < ndim; ++d) {
      uIpL[d] = uL[d] + duL[d] * hoUpwind * limitL[d] * density_upwinding_factor;
      uIpR[d] = uR[d] - duR[d] * hoUpwind * limitR[d] * density_upwinding_factor;
    }

    double duidxj[kMaxDim][kMaxDim];
    for (int i = 0; i < ndim; ++i) {
      const double dui = uR[i] - uL[i];
      const int offset = i * ndim;
      double gjuidx = 0.0;
      for (int j = 0; j < ndim; ++j) {
        const double dxj = xR[j] - xL[j];
        const double gjui = 0.5 * (gR[offset + j] + gL[offset + j]);
        gjuidx += gjui * dxj;
      }
      for (int j = 0; j < ndim; ++j) {
        const double gjui = 0.5 * (gR[offset + j] + gL[offset + j]);
        duidxj[i][j] = gjui + (dui - gjuidx) * av[j] * inv_axdx;
      }
    }

    const double dlhsfac = -viscIp * asq * inv_axdx;

    for (int i = 0; i < ndim; ++i) {
      const int rowL = i;
      const int rowR = i + ndim;

      const double uiIp = 0.5 * (uR[i] + uL[i]);
      const double uiUpw = (mdot > 0.0)
                             ? (alphaUpw * uIpL[i] + om_alphaUpw * uiIp)
                             : (alphaUpw * uIpR[i] + om_alphaUpw * uiIp);
      const double uiHatL = (alpha * uIpL[i] + om_alpha * uiIp);
      const double uiHatR = (alpha * uIpR[i] + om_alpha * uiIp);
      const double uiCds = 0.5 * (uiHatL + uiHatR);

      const double adv_flux = mdot * (pecfac * uiUpw + om_pecfac * uiCds);

      double diff_flux = 0.0;
      for (int j = 0; j < ndim; ++j)
        diff_flux += duidxj[j][j];
      diff_flux *= 2.0 / 3.0 * viscIp * av[i] * includeDivU;
      for (int j = 0; j < ndim; ++j)
        diff_flux += -viscIp * (duidxj[i][j] + duidxj[j][i]) * av[j];

      const double maskNode = std::fmin(node_mask[nL], node_mask[nR]);
      const double total_flux = adv_flux + diff_flux * maskNode;

      rhs[rowL] -= total_flux;
      rhs[rowR] += total_flux;

      double alhsfac = 0.5 * (mdot + std::fabs(mdot)) * pecfac * alphaUpw +
                       0.5 * alpha * om_pecfac * mdot;
      LHS(rowL, rowL) += alhsfac / relaxFacU;
      LHS(rowR, rowL) -= alhsfac;

      alhsfac = 0.5 * (mdot - std::fabs(mdot)) * pecfac * alphaUpw +
                0.5 * alpha * om_pecfac * mdot;
      LHS(rowR, rowR) -= alhsfac / relaxFacU;
      LHS(rowL, rowR) += alhsfac;

      alhsfac = 0.5 * mdot * (pecfac * om_alphaUpw + om_pecfac * om_alpha);
      LHS(rowL, rowL) += alhsfac / relaxFacU;
      LHS(rowL, rowR) += alhsfac;
      LHS(rowR, rowL) -= alhsfac;
      LHS(rowR, rowR) -= alhsfac / relaxFacU;

      LHS(rowL, rowL) -= dlhsfac / relaxFacU;
      LHS(rowL, rowR) += dlhsfac;
      LHS(rowR, rowL) += dlhsfac;
      LHS(rowR, rowR) -= dlhsfac / relaxFacU;

      for (int j = 0; j < ndim; ++j) {
        const double lhsfacNS = -viscIp * av[i] * av[j] * inv_axdx;
        const int colL = j;
        const int colR = j + ndim;
        LHS(rowL, colL) -= lhsfacNS / relaxFacU;
        LHS(rowL, colR) += lhsfacNS;
        LHS(rowR, colL) += lhsfacNS;
        LHS(rowR, colR) -= lhsfacNS / relaxFacU;
      }
    }

    /* NGPApplyCoeff::operator(): src/SolverAlgorithm.C:132-151 */
    if (udiag_accum) {
      for (int i = 0; i < 2; ++i) {
        const int ix = i * ndim;
        add_to(udiag_accum[nodes[i]], LHS(ix, ix));
      }
    }
    sink->apply(2, nodes, rhs, lhs, n);
#undef LHS
  }
}

extern "C" void
orc_momentum_edge(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* vel, const double* dudx, const double* viscosity,
  const double* density, const double* node_mask, const double* edge_area,
  const double* mdot_f, const double* pecfac_f, const orc_momentum_opts* o,
  orc_applier* sink, double* udiag_accum)
{
  momentum_edge_impl(
    ndim, n_edges, edge_nodes, coords, vel, dudx, viscosity, density, node_mask,
    edge_area, mdot_f, nullptr, pecfac_f, o, sink, udiag_accum);
}

/* realm_has_vof_ on: mass_vof = the mass_vof_balanced_flow_rate edge field */
extern "C" void
orc_momentum_edge_vof(
  int ndim, int64_t n_edges, const int32_t* edge_nodes, const double* coords,
  const double* vel, const double* dudx, const double* viscosity,
  const double* density, const double* node_mask, const double* edge_area,
  const double* mdot_f, const double* mass_vof, const double* pecfac_f,
  const orc_momentum_opts* o, orc_applier* sink, double* udiag_accum)
{
  momentum_edge_impl(
    ndim, n_edges, edge_nodes, coords, vel, dudx, viscosity, density, node_mask,
    edge_area, mdot_f, mass_vof, pecfac_f, o, sink, udiag_accum);
}
