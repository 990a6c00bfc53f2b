/*
 * oracle/ref_shim/hypre/HypreHarness.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Stand-ins that let the reference's src/HypreLinearSystem.C (graph build,
 * CoeffApplier::sum_into / sum_into_1DoF, resetCoeffApplierData, loadComplete)
 * compile unmodified and run on flat arrays: the hypre IJ interface as a
 * recorder (what is handed to HYPRE_IJMatrixSetValues2 / AddToValues2 and the
 * vector calls is kept for inspection), the solver / config classes, MPI.
 * Written for this repo; not hypre code.
 */
#ifndef NW_HYPRE_HARNESS_H
#define NW_HYPRE_HARNESS_H
#include <RefHarness.h>
#include <Kokkos_UnorderedMap.hpp>

#include <iomanip>
#include <iostream>
#include <set>
#include <unordered_set>

/* ---------------- MPI (one process; rank / size from the World) ------------ */
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_LONG 3
#define MPI_LONG_LONG 4
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define HYPRE_MPI_INT MPI_INT
inline int
nwref_mpi_copy(const void* s, void* r, int n, MPI_Datatype t)
{
  const size_t sz = (t == MPI_INT) ? sizeof(int) : 8;
  std::memcpy(r, s, sz * (size_t)n);
  return 0;
}
inline int MPI_Barrier(MPI_Comm) { return 0; }
inline int
MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm)
{
  return nwref_mpi_copy(s, r, n, t);
}
inline int
MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{
  return nwref_mpi_copy(s, r, n, t);
}

/* ---------------- hypre IJ interface: a recorder --------------------------- */
typedef int HYPRE_Int;
typedef int HYPRE_BigInt;
typedef double HYPRE_Real;
typedef double HYPRE_Complex;
#define HYPRE_PARCSR 5555

namespace nwref {
struct IJCall
{
  std::string what; /* "Set" / "AddTo" */
  std::vector<HYPRE_Int> ncols, rows, row_indexes, cols;
  std::vector<double> values;
};
struct IJMatrix
{
  HYPRE_BigInt ilower, iupper, jlower, jupper;
  std::vector<IJCall> calls;
  bool assembled = false;
};
struct IJVector
{
  HYPRE_BigInt jlower, jupper;
  std::vector<IJCall> calls;
  bool assembled = false;
  double constant = 0.0;
};
struct Recorder
{
  std::vector<IJMatrix*> matrices;
  std::vector<IJVector*> vectors;
  static Recorder& self()
  {
    static Recorder r;
    return r;
  }
};
} // namespace nwref
typedef nwref::IJMatrix* HYPRE_IJMatrix;
typedef nwref::IJVector* HYPRE_IJVector;
typedef nwref::IJMatrix* HYPRE_ParCSRMatrix;
typedef nwref::IJVector* HYPRE_ParVector;

inline HYPRE_Int
HYPRE_IJMatrixCreate(
  MPI_Comm, HYPRE_BigInt il, HYPRE_BigInt iu, HYPRE_BigInt jl, HYPRE_BigInt ju,
  HYPRE_IJMatrix* m)
{
  *m = new nwref::IJMatrix{il, iu, jl, ju, {}, false};
  nwref::Recorder::self().matrices.push_back(*m);
  return 0;
}
inline HYPRE_Int HYPRE_IJMatrixSetObjectType(HYPRE_IJMatrix, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixInitialize(HYPRE_IJMatrix) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixInitialize_v2(HYPRE_IJMatrix, int) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixSetRowSizes(HYPRE_IJMatrix, const HYPRE_Int*) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixSetDiagOffdSizes(HYPRE_IJMatrix, const HYPRE_Int*, const HYPRE_Int*) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixSetMaxOffProcElmts(HYPRE_IJMatrix, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixSetMaxOnProcElmts(HYPRE_IJMatrix, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixSetOffProcSendElmts(HYPRE_IJMatrix, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixSetOffProcRecvElmts(HYPRE_IJMatrix, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixSetOMPFlag(HYPRE_IJMatrix, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixSetConstantValues(HYPRE_IJMatrix, double) { return 0; }
inline HYPRE_Int HYPRE_IJMatrixPrint(HYPRE_IJMatrix, const char*) { return 0; }
inline HYPRE_Int
HYPRE_IJMatrixAssemble(HYPRE_IJMatrix m)
{
  m->assembled = true;
  return 0;
}
inline HYPRE_Int
HYPRE_IJMatrixDestroy(HYPRE_IJMatrix)
{
  return 0; /* kept alive for inspection; freed by the driver */
}
inline HYPRE_Int
HYPRE_IJMatrixGetObject(HYPRE_IJMatrix m, void** o)
{
  *o = m;
  return 0;
}
inline HYPRE_Int
nwref_ij_record(
  std::vector<nwref::IJCall>& calls, const char* what, HYPRE_Int nrows,
  const HYPRE_Int* ncols, const HYPRE_BigInt* rows, const HYPRE_Int* row_indexes,
  const HYPRE_BigInt* cols, const double* values)
{
  nwref::IJCall c;
  c.what = what;
  HYPRE_Int nnz = 0;
  for (HYPRE_Int i = 0; i < nrows; ++i) {
    const HYPRE_Int n = ncols ? ncols[i] : 1;
    c.ncols.push_back(n);
    c.rows.push_back(rows[i]);
    const HYPRE_Int at = row_indexes ? row_indexes[i] : nnz;
    c.row_indexes.push_back(at);
    for (HYPRE_Int k = 0; k < n; ++k) {
      c.cols.push_back(cols ? cols[at + k] : rows[i]);
      c.values.push_back(values[at + k]);
    }
    nnz += n;
  }
  calls.push_back(std::move(c));
  return 0;
}
inline HYPRE_Int
HYPRE_IJMatrixSetValues2(
  HYPRE_IJMatrix m, HYPRE_Int nrows, HYPRE_Int* ncols, const HYPRE_BigInt* rows,
  const HYPRE_Int* row_indexes, const HYPRE_BigInt* cols, const double* values)
{
  return nwref_ij_record(m->calls, "Set", nrows, ncols, rows, row_indexes, cols, values);
}
inline HYPRE_Int
HYPRE_IJMatrixAddToValues2(
  HYPRE_IJMatrix m, HYPRE_Int nrows, HYPRE_Int* ncols, const HYPRE_BigInt* rows,
  const HYPRE_Int* row_indexes, const HYPRE_BigInt* cols, const double* values)
{
  return nwref_ij_record(m->calls, "AddTo", nrows, ncols, rows, row_indexes, cols, values);
}
inline HYPRE_Int
HYPRE_IJMatrixSetValues(
  HYPRE_IJMatrix m, HYPRE_Int nrows, HYPRE_Int* ncols, const HYPRE_BigInt* rows,
  const HYPRE_BigInt* cols, const double* values)
{
  return nwref_ij_record(m->calls, "Set", nrows, ncols, rows, nullptr, cols, values);
}
inline HYPRE_Int
HYPRE_IJMatrixAddToValues(
  HYPRE_IJMatrix m, HYPRE_Int nrows, HYPRE_Int* ncols, const HYPRE_BigInt* rows,
  const HYPRE_BigInt* cols, const double* values)
{
  return nwref_ij_record(m->calls, "AddTo", nrows, ncols, rows, nullptr, cols, values);
}

inline HYPRE_Int
HYPRE_IJVectorCreate(MPI_Comm, HYPRE_BigInt jl, HYPRE_BigInt ju, HYPRE_IJVector* v)
{
  *v = new nwref::IJVector{jl, ju, {}, false, 0.0};
  nwref::Recorder::self().vectors.push_back(*v);
  return 0;
}
inline HYPRE_Int HYPRE_IJVectorSetObjectType(HYPRE_IJVector, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJVectorInitialize(HYPRE_IJVector) { return 0; }
inline HYPRE_Int HYPRE_IJVectorInitialize_v2(HYPRE_IJVector, int) { return 0; }
inline HYPRE_Int HYPRE_IJVectorSetMaxOffProcElmts(HYPRE_IJVector, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJVectorSetMaxOnProcElmts(HYPRE_IJVector, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJVectorSetOffProcSendElmts(HYPRE_IJVector, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJVectorSetOffProcRecvElmts(HYPRE_IJVector, HYPRE_Int) { return 0; }
inline HYPRE_Int HYPRE_IJVectorPrint(HYPRE_IJVector, const char*) { return 0; }
inline HYPRE_Int HYPRE_IJVectorDestroy(HYPRE_IJVector) { return 0; }
inline HYPRE_Int
HYPRE_IJVectorAssemble(HYPRE_IJVector v)
{
  v->assembled = true;
  return 0;
}
inline HYPRE_Int
HYPRE_IJVectorGetObject(HYPRE_IJVector v, void** o)
{
  *o = v;
  return 0;
}
inline HYPRE_Int
HYPRE_IJVectorSetValues(
  HYPRE_IJVector v, HYPRE_Int n, const HYPRE_BigInt* idx, const double* values)
{
  return nwref_ij_record(v->calls, "Set", n, nullptr, idx, nullptr, nullptr, values);
}
inline HYPRE_Int
HYPRE_IJVectorAddToValues(
  HYPRE_IJVector v, HYPRE_Int n, const HYPRE_BigInt* idx, const double* values)
{
  return nwref_ij_record(v->calls, "AddTo", n, nullptr, idx, nullptr, nullptr, values);
}
inline HYPRE_Int
HYPRE_ParVectorSetConstantValues(HYPRE_ParVector v, double c)
{
  v->constant = c;
  return 0;
}

inline HYPRE_Int
HYPRE_IJMatrixGetRowCounts(HYPRE_IJMatrix, HYPRE_Int n, HYPRE_BigInt*, HYPRE_Int* c)
{
  for (HYPRE_Int i = 0; i < n; ++i)
    c[i] = 0;
  return 0;
}
/* internal hypre accessors copy_hypre_to_stk uses (solution hand-back; not on
 * the assembly path): an empty local vector */
typedef nwref::IJVector hypre_ParVector;
inline hypre_ParVector* hypre_IJVectorObject(HYPRE_IJVector v) { return v; }
inline hypre_ParVector* hypre_ParVectorLocalVector(hypre_ParVector* v) { return v; }
inline double*
hypre_VectorData(hypre_ParVector*)
{
  static std::vector<double> z(1, 0.0);
  return z.data();
}
inline double hypre_ParVectorInnerProd(hypre_ParVector*, hypre_ParVector*) { return 0.0; }

namespace sierra {
namespace nalu {

/* LinearSolverConfig.h / LinearSolver.h / HypreDirectSolver.h */
class LinearSolverConfig
{
public:
  virtual ~LinearSolverConfig() {}
  bool getWriteMatrixFiles() const { return false; }
  bool recomputePreconditioner() const { return true; }
  bool reusePreconditioner() const { return false; }
  bool useSegregatedSolver() const { return false; }
  std::string name() const { return "hypre"; }
};
class HypreLinearSolverConfig : public LinearSolverConfig
{
public:
  bool simpleHypreMatrixAssemble() const
  {
    return nwref::World::self().get("simple_hypre_matrix_assemble", 0.0) != 0.0;
  }
  bool getWritePreassemblyMatrixFiles() const { return false; }
  bool dumpHypreMatrixStats() const { return false; }
};
class LinearSolver
{
public:
  virtual ~LinearSolver() {}
  LinearSolverConfig* getConfig() { return &config_; }
  HypreLinearSolverConfig config_;
  bool& activeMueLu() { return mueLu_; }
  bool mueLu_ = false;
};
class HypreDirectSolver : public LinearSolver
{
public:
  HYPRE_ParCSRMatrix parMat_ = nullptr;
  HYPRE_ParVector parRhs_ = nullptr;
  HYPRE_ParVector parSln_ = nullptr;
  MPI_Comm comm_ = 0;
  int solve(int& iters, double& norm, bool)
  {
    iters = 0;
    norm = 0.0;
    return 0;
  }
  void set_initialize_solver_flag() {}
};
/* HypreUVWSolver.h */
class HypreUVWSolver : public HypreDirectSolver
{
public:
  using HypreDirectSolver::solve;
  int solve(int, int& iters, double& norm, bool)
  {
    iters = 0;
    norm = 0.0;
    return 0;
  }
  /* src/HypreUVWSolver.C:21 sizes both to 3 */
  mutable std::vector<HYPRE_ParVector> parRhsU_ = std::vector<HYPRE_ParVector>(3);
  mutable std::vector<HYPRE_ParVector> parSlnU_ = std::vector<HYPRE_ParVector>(3);
};

} // namespace nalu
} // namespace sierra
#endif
