/*
 * nw_api.cu -- implementation of the C ABI declared in include/nalu_edge_b200.h.
 * Host logic only; the kernels are in nw_kernels.cu, the plan builder in
 * plan.cpp, the communicator in nw_comm.cpp.
 */
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "geometry_cvfem.h"
#include "nw_comm.h"
#include "nw_internal.h"

namespace nw {

static thread_local std::string g_err;

void
set_error(const std::string& m)
{
  g_err = m;
}

static int
fail(int code, const std::string& m)
{
  g_err = m;
  return code;
}

#define NW_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess)                                                   \
      return fail(                                                           \
        NW_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
  } while (0)

#define NW_TRY(body)                                                         \
  try {                                                                      \
    body                                                                     \
  } catch (const std::exception& ex) {                                       \
    return fail(NW_ERR_ARG, ex.what());                                      \
  }

static int
need_device(const nw_ctx* ctx, const char* what)
{
  if (!ctx || ctx->device < 0)
    return fail(
      NW_ERR_CUDA, std::string(what) +
                     ": no CUDA device (host-only context); this library has "
                     "no CPU fallback");
  return NW_OK;
}

template <class T>
static int
upload(DevBuf& b, const std::vector<T>& v, cudaStream_t s, int64_t* acc)
{
  NW_CUDA(b.alloc(v.size() * sizeof(T)));
  if (!v.empty())
    NW_CUDA(cudaMemcpyAsync(
      b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  if (acc)
    *acc += (int64_t)(v.size() * sizeof(T));
  return NW_OK;
}

static inline int
even_up(int v)
{
  return (v + 1) & ~1;
}

} // namespace nw

using namespace nw;

extern "C" const char*
nw_last_error(void)
{
  return g_err.c_str();
}

extern "C" int
nw_version(void)
{
  return 1000;
}

extern "C" int
nw_debug_phase_times(unsigned long long* out, int n, int reset)
{
  if (!out || n < kPhaseKernels * kPhaseSlots)
    return fail(NW_ERR_ARG, "nw_debug_phase_times: buffer too small");
  if (phase_times_read(out, reset != 0) != cudaSuccess)
    return fail(NW_ERR_CUDA, "nw_debug_phase_times: device read failed");
  return NW_OK;
}

/* ------------------------------------------------------------------ */
/*  context                                                            */
/* ------------------------------------------------------------------ */

extern "C" int
nw_ctx_create(int cuda_device, nw_ctx** out)
{
  if (!out)
    return fail(NW_ERR_ARG, "nw_ctx_create: out is NULL");
  auto* c = new nw_ctx;
  c->device = cuda_device;
  if (cuda_device >= 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || cuda_device >= n) {
      delete c;
      return fail(
        NW_ERR_CUDA,
        std::string("nw_ctx_create: CUDA device not available: ") +
          (e != cudaSuccess ? cudaGetErrorString(e) : "index out of range"));
    }
    e = cudaSetDevice(cuda_device);
    if (e == cudaSuccess)
      e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete c;
      return fail(
        NW_ERR_CUDA,
        std::string("nw_ctx_create: ") + cudaGetErrorString(e));
    }
  }
  *out = c;
  return NW_OK;
}

static int p2p_init(nw_ctx* ctx);
static int p2p_complete_pending(nw_ctx* ctx);
static int p2p_check_error(nw_ctx* ctx);
static void p2p_queue_error_read(nw_ctx* ctx, cudaStream_t s);
static int p2p_error_after_sync(nw_ctx* ctx);
static void p2p_close(nw_ctx* ctx);

extern "C" int
nw_ctx_destroy(nw_ctx* ctx)
{
  if (!ctx)
    return NW_OK;
  p2p_close(ctx);
  comm_destroy(ctx->comm);
  if (ctx->stream)
    cudaStreamDestroy(ctx->stream);
  if (ctx->copyStream)
    cudaStreamDestroy(ctx->copyStream);
  delete ctx;
  return NW_OK;
}

extern "C" int
nw_ctx_sync(nw_ctx* ctx)
{
  if (int rc = need_device(ctx, "nw_ctx_sync"))
    return rc;
  NW_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->p2p.commStream)
    NW_CUDA(cudaStreamSynchronize(ctx->p2p.commStream));
  /* a pull kernel that gave up waiting for a peer leaves a mark */
  return p2p_check_error(ctx);
}

extern "C" int
nw_debug_skip_exchange(nw_ctx* ctx, int on)
{
  if (!ctx)
    return fail(NW_ERR_ARG, "nw_debug_skip_exchange: NULL context");
  ctx->skipExchange = on != 0;
  return NW_OK;
}

extern "C" void*
nw_ctx_stream(nw_ctx* ctx)
{
  return ctx ? (void*)ctx->stream : nullptr;
}

extern "C" int
nw_comm_unique_id(void* unique_id_out)
{
  std::string err;
  if (!comm_unique_id(unique_id_out, err))
    return fail(NW_ERR_COMM, "nw_comm_unique_id: " + err);
  return NW_OK;
}

extern "C" int
nw_ctx_comm_init(nw_ctx* ctx, const void* unique_id, int nranks, int rank)
{
  if (int rc = need_device(ctx, "nw_ctx_comm_init"))
    return rc;
  NW_CUDA(cudaSetDevice(ctx->device));
  std::string err;
  if (!comm_init(ctx->comm, unique_id, nranks, rank, err))
    return fail(NW_ERR_COMM, "nw_ctx_comm_init: " + err);
  /* peer-memory mailbox for the halo exchanges (falls back to NCCL) */
  return p2p_init(ctx);
}

/* ------------------------------------------------------------------ */
/*  mesh                                                               */
/* ------------------------------------------------------------------ */

static int mesh_halo_prepare(nw_mesh* mesh, const int64_t* ownHid);
static int mesh_halo_auto(nw_mesh* mesh);

static int
mesh_upload_plan(nw_mesh* m)
{
  cudaStream_t s = m->ctx->stream;
  MeshPlan& p = m->plan;
  int64_t& acc = m->planBytes;
  int rc;
  if ((rc = upload(m->dTiles, p.tiles, s, &acc)) ||
      (rc = upload(m->dHalo, p.haloNodes, s, &acc)) ||
      (rc = upload(m->dHaloBlock, p.haloBlock, s, &acc)) ||
      (rc = upload(m->dLr, p.lr, s, &acc)) ||
      (rc = upload(m->dHeNode, p.heNodeEll, s, &acc)) ||
      (rc = upload(m->dWarpNode, p.sliceOffNode, s, &acc)) ||
      (rc = upload(m->dPrimary, p.tileEdgePrimary, s, &acc)) ||
      (rc = upload(m->dNodeOfSlot, p.nodeOfSlot, s, &acc)) ||
      (rc = upload(m->dTileEdgeSrc, p.tileEdgeSrc, s, &acc)) ||
      (rc = upload(m->dPrimarySlot, p.primarySlotOfEdge, s, &acc)) ||
      (rc = upload(m->dSecondSlot, p.secondSlotOfEdge, s, &acc)))
    return rc;
  NW_CUDA(cudaStreamSynchronize(s));
  MeshPlanDev& d = m->dev;
  d.tiles = m->dTiles.as<TileHdr>();
  d.haloNodes = m->dHalo.as<int32_t>();
  d.haloBlock = m->dHaloBlock.as<int32_t>();
  d.lr = m->dLr.as<uint32_t>();
  d.heNodeEll = m->dHeNode.as<uint32_t>();
  d.sliceOffNode = m->dWarpNode.as<int32_t>();
  d.primary = m->dPrimary.as<uint8_t>();
  return NW_OK;
}

extern "C" int
nw_mesh_create(nw_ctx* ctx, const nw_mesh_desc* desc, nw_mesh** out)
{
  if (!ctx || !desc || !out)
    return fail(NW_ERR_ARG, "nw_mesh_create: NULL argument");
  auto m = std::make_unique<nw_mesh>();
  m->ctx = ctx;
  MeshInput in;
  in.ndim = desc->ndim;
  in.rank = desc->rank;
  in.nranks = desc->nranks;
  in.nNodes = desc->n_nodes;
  in.nEdges = desc->n_edges;
  in.edgeNodes = desc->edge_nodes;
  in.nodeHid = desc->node_hypre_id;
  in.nodeOwnHid = desc->node_own_hypre_id;
  in.hypreOffsets = desc->hypre_offsets;
  in.coords = desc->coords;
  in.tileNodes = desc->tile_nodes;
  NW_TRY(build_mesh_plan(in, m->plan);)
  if (desc->tile_nodes <= 0) {
    /* default tile size: meshes with many edges per node (tets: ~7 against a
     * hex mesh's 3) stage more per tile; halve the tile until the largest
     * shared-memory layout (momentum: 18 node components, 7 doubles per edge,
     * row staging, reduction plan) fits one CTA with room to spare */
    int T = 192;
    auto need = [&]() {
      const MeshPlan& q = m->plan;
      const size_t staged = (size_t)q.maxTileStaged + 2, te = (size_t)q.maxTileEdges + 4;
      return 8 * (18 * staged + 7 * te) + 4 * te + 6 * (size_t)q.maxTileHalf +
             (size_t)16 * 192 * 8;
    };
    while (need() > 200 * 1024 && T > 24) {
      T /= 2;
      in.tileNodes = T;
      NW_TRY(build_mesh_plan(in, m->plan);)
    }
  }
  MeshPlanDev& d = m->dev;
  d.nTiles = (int)m->plan.nTiles;
  d.ndim = m->plan.ndim;
  d.maxStaged = even_up((int)m->plan.maxTileStaged);
  d.maxTileEdges = (int)m->plan.maxTileEdges;
  d.maxTileNodes = (int)m->plan.maxTileNodes;
  d.maxTileEllNode = (int)m->plan.maxTileEllNode;
  if (ctx->device >= 0) {
    NW_CUDA(cudaSetDevice(ctx->device));
    if (int rc = mesh_upload_plan(m.get()))
      return rc;
  }
  {
    /* node-kernel selector (src/AssembleNGPNodeSolverAlgorithm.C:108-111):
     * locally owned & !periodic-slave.  A slave carries its master's row id
     * in node_hypre_id and its own in node_own_hypre_id. */
    const int64_t* own =
      desc->node_own_hypre_id ? desc->node_own_hypre_id : desc->node_hypre_id;
    m->nodeKernelActive.assign(desc->n_nodes, 0);
    for (int64_t n = 0; n < desc->n_nodes; ++n)
      m->nodeKernelActive[n] =
        own[n] >= m->plan.iLowerNode && own[n] <= m->plan.iUpperNode &&
        own[n] == desc->node_hypre_id[n];
  }
  if (desc->node_own_hypre_id) {
    /* periodic row groups: nodes whose resolved row id coincides */
    const int64_t* hid = desc->node_hypre_id;
    const int64_t* own = desc->node_own_hypre_id;
    std::vector<int32_t> alias; /* nodes of groups with a slave */
    std::vector<int64_t> slaveRows;
    for (int64_t n = 0; n < desc->n_nodes; ++n)
      if (own[n] != hid[n])
        slaveRows.push_back(hid[n]);
    if (!slaveRows.empty()) {
      std::sort(slaveRows.begin(), slaveRows.end());
      slaveRows.erase(
        std::unique(slaveRows.begin(), slaveRows.end()), slaveRows.end());
      for (int64_t n = 0; n < desc->n_nodes; ++n)
        if (std::binary_search(slaveRows.begin(), slaveRows.end(), hid[n]))
          alias.push_back((int32_t)n);
      std::sort(alias.begin(), alias.end(), [&](int32_t x, int32_t y) {
        if (hid[x] != hid[y])
          return hid[x] < hid[y];
        const bool mx = own[x] == hid[x], my = own[y] == hid[y];
        if (mx != my)
          return mx; /* the master leads its group */
        if (own[x] != own[y])
          return own[x] < own[y];
        return x < y;
      });
      for (size_t q = 0; q < alias.size(); ++q) {
        const int32_t n = alias[q];
        if (q == 0 || hid[n] != hid[alias[q - 1]]) {
          m->perPtr.push_back((int32_t)m->perSlots.size());
          if (own[n] != hid[n])
            m->perMasterMissing = true;
        }
        m->perSlots.push_back(m->plan.slotOfNode[n]);
      }
      m->perPtr.push_back((int32_t)m->perSlots.size());
    }
  }
  if (desc->nranks > 1) {
    const int64_t* own =
      desc->node_own_hypre_id ? desc->node_own_hypre_id : desc->node_hypre_id;
    m->ownedNodeOfHid.assign(
      (size_t)std::max<int64_t>(0, m->plan.iUpperNode - m->plan.iLowerNode + 1),
      -1);
    for (int64_t n = 0; n < desc->n_nodes; ++n)
      if (own[n] >= m->plan.iLowerNode && own[n] <= m->plan.iUpperNode)
        m->ownedNodeOfHid[own[n] - m->plan.iLowerNode] = (int32_t)n;
    if (int rc = mesh_halo_prepare(m.get(), own))
      return rc;
  }
  nw_mesh* raw = m.release();
  /* the coordinates field comes with the mesh (realm.get_coordinates_name()) */
  int fid = -1;
  int rc = nw_field_register(raw, "coordinates", NW_NODE, desc->ndim, &fid);
  if (rc == NW_OK && ctx->device >= 0)
    rc = nw_field_upload(raw, fid, desc->coords);
  if (rc != NW_OK) {
    delete raw;
    return rc;
  }
  *out = raw;
  return NW_OK;
}

extern "C" int
nw_mesh_destroy(nw_mesh* mesh)
{
  if (mesh)
    for (auto& f : mesh->fields)
      if (f->pullDone) {
        cudaEventSynchronize(f->pullDone);
        if (mesh->ctx->p2p.lastPull == f->pullDone)
          mesh->ctx->p2p.lastPull = nullptr;
      }
  delete mesh;
  return NW_OK;
}

extern "C" int
nw_mesh_get_stats(const nw_mesh* mesh, nw_mesh_stats* out)
{
  if (!mesh || !out)
    return fail(NW_ERR_ARG, "nw_mesh_get_stats: NULL argument");
  const MeshPlan& p = mesh->plan;
  out->n_nodes = p.nNodes;
  out->n_edges = p.nEdges;
  out->n_tiles = p.nTiles;
  int64_t te = 0;
  for (const TileHdr& h : p.tiles)
    te += h.nEdges;
  out->n_tile_edges = te;
  out->n_halo_nodes = p.totalHalo;
  out->max_tile_nodes = p.maxTileNodes;
  out->max_tile_staged = p.maxTileStaged;
  out->max_tile_edges = p.maxTileEdges;
  out->max_tile_halfedges = p.maxTileHalf;
  out->plan_bytes_device = mesh->planBytes;
  return NW_OK;
}

extern "C" int
nw_mesh_get_node_permutation(
  const nw_mesh* mesh, int64_t* n_slots, int32_t* perm)
{
  if (!mesh || !n_slots)
    return fail(NW_ERR_ARG, "nw_mesh_get_node_permutation: NULL argument");
  *n_slots = mesh->plan.nSlots;
  if (perm)
    std::memcpy(
      perm, mesh->plan.nodeOfSlot.data(), mesh->plan.nSlots * sizeof(int32_t));
  return NW_OK;
}

/* ---- fields ---- */

extern "C" int
nw_field_register(
  nw_mesh* mesh, const char* name, int entity_rank, int ncomp, int* field_id)
{
  if (!mesh || !name || !field_id)
    return fail(NW_ERR_ARG, "nw_field_register: NULL argument");
  if (ncomp < 1 || ncomp > 9)
    return fail(NW_ERR_ARG, "nw_field_register: ncomp must be 1..9");
  if (entity_rank != NW_NODE && entity_rank != NW_EDGE)
    return fail(NW_ERR_ARG, "nw_field_register: bad entity rank");
  auto it = mesh->fieldByName.find(name);
  if (it != mesh->fieldByName.end()) {
    nw_field_t& f = *mesh->fields[it->second];
    if (f.rank != entity_rank || f.ncomp != ncomp)
      return fail(
        NW_ERR_ARG, std::string("nw_field_register: field '") + name +
                      "' already registered with another shape");
    *field_id = it->second;
    return NW_OK;
  }
  auto f = std::make_unique<nw_field_t>();
  f->name = name;
  f->rank = entity_rank;
  f->ncomp = ncomp;
  f->stride =
    entity_rank == NW_NODE ? mesh->plan.nSlots : mesh->plan.nTileEdgeSlots;
  if (mesh->ctx->device >= 0) {
    NW_CUDA(cudaSetDevice(mesh->ctx->device));
    NW_CUDA(f->buf.alloc(sizeof(double) * f->stride * ncomp));
    NW_CUDA(cudaMemsetAsync(
      f->buf.p, 0, sizeof(double) * f->stride * ncomp, mesh->ctx->stream));
  }
  *field_id = (int)mesh->fields.size();
  mesh->fieldByName[name] = *field_id;
  mesh->fields.push_back(std::move(f));
  return NW_OK;
}

extern "C" int
nw_field_find(const nw_mesh* mesh, const char* name, int* field_id)
{
  if (!mesh || !name || !field_id)
    return fail(NW_ERR_ARG, "nw_field_find: NULL argument");
  auto it = mesh->fieldByName.find(name);
  if (it == mesh->fieldByName.end())
    return fail(
      NW_ERR_ARG, std::string("field '") + name + "' is not registered");
  *field_id = it->second;
  return NW_OK;
}

/* a halo sum of this field may still be running on the communication stream:
 * order the compute stream (and with it every later use) after it */
static void
field_wait_pull(nw_mesh* mesh, nw_field_t* f)
{
  if (f->pullPending) {
    cudaStreamWaitEvent(mesh->ctx->stream, f->pullDone, 0);
    f->pullPending = false;
  }
}

static nw_field_t*
get_field(nw_mesh* mesh, int id)
{
  if (!mesh || id < 0 || id >= (int)mesh->fields.size())
    return nullptr;
  nw_field_t* f = mesh->fields[id].get();
  field_wait_pull(mesh, f);
  return f;
}

static int
ensure_scratch(nw_mesh* mesh, size_t bytes)
{
  if (mesh->scratch.bytes < bytes) {
    /* wait for earlier users of the old scratch before freeing it */
    NW_CUDA(cudaStreamSynchronize(mesh->ctx->stream));
    NW_CUDA(mesh->scratch.alloc(bytes));
  }
  return NW_OK;
}

extern "C" int
nw_field_upload(nw_mesh* mesh, int field_id, const double* host)
{
  nw_field_t* f = get_field(mesh, field_id);
  if (!f || !host)
    return fail(NW_ERR_ARG, "nw_field_upload: bad field id or NULL buffer");
  if (int rc = need_device(mesh->ctx, "nw_field_upload"))
    return rc;
  cudaStream_t s = mesh->ctx->stream;
  const int64_t nEnt = f->rank == NW_NODE ? mesh->plan.nNodes : mesh->plan.nEdges;
  const size_t bytes = sizeof(double) * nEnt * f->ncomp;
  if (int rc = ensure_scratch(mesh, bytes))
    return rc;
  NW_CUDA(cudaMemcpyAsync(mesh->scratch.p, host, bytes, cudaMemcpyHostToDevice, s));
  if (f->rank == NW_NODE)
    NW_CUDA(launch_node_gather(
      mesh->scratch.as<double>(), f->ncomp, mesh->dNodeOfSlot.as<int32_t>(),
      f->stride, f->buf.as<double>(), s));
  else
    NW_CUDA(launch_edge_gather(
      mesh->scratch.as<double>(), f->ncomp, mesh->dTileEdgeSrc.as<int32_t>(),
      f->stride, f->buf.as<double>(), s));
  return NW_OK;
}

extern "C" int
nw_field_stage(nw_mesh* mesh, int field_id, const double* host)
{
  nw_field_t* f = get_field(mesh, field_id);
  if (!f || !host)
    return fail(NW_ERR_ARG, "nw_field_stage: bad field id or NULL buffer");
  if (int rc = need_device(mesh->ctx, "nw_field_stage"))
    return rc;
  nw_ctx* ctx = mesh->ctx;
  if (!ctx->copyStream)
    NW_CUDA(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
  const int64_t nEnt = f->rank == NW_NODE ? mesh->plan.nNodes : mesh->plan.nEdges;
  const size_t bytes = sizeof(double) * nEnt * f->ncomp;
  if (f->staging.bytes < bytes) {
    if (f->stagePending)
      return fail(NW_ERR_STATE, "nw_field_stage: previous stage not committed");
    NW_CUDA(f->staging.alloc(bytes));
  }
  if (!f->staged) {
    NW_CUDA(cudaEventCreateWithFlags(&f->staged, cudaEventDisableTiming));
    NW_CUDA(cudaEventCreateWithFlags(&f->consumed, cudaEventDisableTiming));
  } else {
    /* the previous commit's permute kernel must have read the buffer */
    NW_CUDA(cudaStreamWaitEvent(ctx->copyStream, f->consumed, 0));
  }
  NW_CUDA(cudaMemcpyAsync(
    f->staging.p, host, bytes, cudaMemcpyHostToDevice, ctx->copyStream));
  NW_CUDA(cudaEventRecord(f->staged, ctx->copyStream));
  f->stagePending = true;
  return NW_OK;
}

extern "C" int
nw_field_commit(nw_mesh* mesh, int field_id)
{
  nw_field_t* f = get_field(mesh, field_id);
  if (!f)
    return fail(NW_ERR_ARG, "nw_field_commit: bad field id");
  if (!f->stagePending)
    return fail(NW_ERR_STATE, "nw_field_commit: nothing staged for this field");
  cudaStream_t s = mesh->ctx->stream;
  NW_CUDA(cudaStreamWaitEvent(s, f->staged, 0));
  if (f->rank == NW_NODE)
    NW_CUDA(launch_node_gather(
      f->staging.as<double>(), f->ncomp, mesh->dNodeOfSlot.as<int32_t>(),
      f->stride, f->buf.as<double>(), s));
  else
    NW_CUDA(launch_edge_gather(
      f->staging.as<double>(), f->ncomp, mesh->dTileEdgeSrc.as<int32_t>(),
      f->stride, f->buf.as<double>(), s));
  NW_CUDA(cudaEventRecord(f->consumed, s));
  f->stagePending = false;
  return NW_OK;
}

extern "C" int
nw_field_download(nw_mesh* mesh, int field_id, double* host)
{
  nw_field_t* f = get_field(mesh, field_id);
  if (!f || !host)
    return fail(NW_ERR_ARG, "nw_field_download: bad field id or NULL buffer");
  if (int rc = need_device(mesh->ctx, "nw_field_download"))
    return rc;
  cudaStream_t s = mesh->ctx->stream;
  const int64_t nEnt = f->rank == NW_NODE ? mesh->plan.nNodes : mesh->plan.nEdges;
  const size_t bytes = sizeof(double) * nEnt * f->ncomp;
  if (int rc = ensure_scratch(mesh, bytes))
    return rc;
  if (f->rank == NW_NODE)
    NW_CUDA(launch_node_scatter(
      f->buf.as<double>(), f->ncomp, mesh->dNodeOfSlot.as<int32_t>(), f->stride,
      mesh->scratch.as<double>(), s));
  else
    NW_CUDA(launch_edge_scatter(
      f->buf.as<double>(), f->ncomp, mesh->dPrimarySlot.as<int32_t>(),
      mesh->plan.nEdges, f->stride, mesh->scratch.as<double>(), s));
  NW_CUDA(cudaMemcpyAsync(host, mesh->scratch.p, bytes, cudaMemcpyDeviceToHost, s));
  p2p_queue_error_read(mesh->ctx, s);
  NW_CUDA(cudaStreamSynchronize(s));
  return p2p_error_after_sync(mesh->ctx);
}

extern "C" int
nw_field_fill(nw_mesh* mesh, int field_id, double value)
{
  nw_field_t* f = get_field(mesh, field_id);
  if (!f)
    return fail(NW_ERR_ARG, "nw_field_fill: bad field id");
  if (int rc = need_device(mesh->ctx, "nw_field_fill"))
    return rc;
  NW_CUDA(launch_fill(
    f->buf.as<double>(), f->stride * f->ncomp, value, mesh->ctx->stream));
  return NW_OK;
}

extern "C" int
nw_field_device_view(
  nw_mesh* mesh, int field_id, double** base, int64_t* stride)
{
  nw_field_t* f = get_field(mesh, field_id);
  if (!f || !base || !stride)
    return fail(NW_ERR_ARG, "nw_field_device_view: bad argument");
  if (int rc = need_device(mesh->ctx, "nw_field_device_view"))
    return rc;
  *base = f->buf.as<double>();
  *stride = f->stride;
  return NW_OK;
}

/* ---- helpers to bind fields by the reference's names ---- */

static int
bind(
  nw_mesh* mesh, const char* name, int rank, int ncomp, const double** comps)
{
  auto it = mesh->fieldByName.find(name);
  if (it == mesh->fieldByName.end())
    return fail(
      NW_ERR_STATE, std::string("required field '") + name +
                      "' is not registered (get_field_ordinal would throw)");
  nw_field_t& f = *mesh->fields[it->second];
  field_wait_pull(mesh, &f);
  if (f.rank != rank || f.ncomp != ncomp)
    return fail(
      NW_ERR_STATE, std::string("field '") + name + "' has the wrong shape");
  for (int c = 0; c < ncomp; ++c)
    comps[c] = f.buf.as<double>() + (int64_t)c * f.stride;
  return NW_OK;
}

static int
bind_id(nw_mesh* mesh, int id, int rank, int ncomp, const double** comps)
{
  nw_field_t* f = get_field(mesh, id);
  if (!f)
    return fail(NW_ERR_ARG, "bad field id");
  if (f->rank != rank || f->ncomp != ncomp)
    return fail(
      NW_ERR_ARG, std::string("field '") + f->name + "' has the wrong shape");
  for (int c = 0; c < ncomp; ++c)
    comps[c] = f->buf.as<double>() + (int64_t)c * f->stride;
  return NW_OK;
}

static int
bind_edge_common(nw_mesh* mesh, EdgeComps& ec, bool mdot, bool pec)
{
  const int nd = mesh->plan.ndim;
  ec = EdgeComps();
  const double* a[3] = {nullptr, nullptr, nullptr};
  if (int rc = bind(mesh, "edge_area_vector", NW_EDGE, nd, a))
    return rc;
  for (int d = 0; d < 3; ++d)
    ec.area[d] = a[d];
  if (mdot)
    if (int rc = bind(mesh, "mass_flow_rate", NW_EDGE, 1, &ec.mdot))
      return rc;
  if (pec)
    if (int rc = bind(mesh, "peclet_factor", NW_EDGE, 1, &ec.pecfac))
      return rc;
  return NW_OK;
}

/* continuity / mdot node bundle: x, u, dpdx, rho, p, udiag */
static int
bind_cont_nodes(nw_mesh* mesh, NodeComps& nc)
{
  const int nd = mesh->plan.ndim;
  int rc;
  if ((rc = bind(mesh, "coordinates", NW_NODE, nd, &nc.c[0])) ||
      (rc = bind(mesh, "velocity", NW_NODE, nd, &nc.c[nd])) ||
      (rc = bind(mesh, "dpdx", NW_NODE, nd, &nc.c[2 * nd])) ||
      (rc = bind(mesh, "density", NW_NODE, 1, &nc.c[3 * nd])) ||
      (rc = bind(mesh, "pressure", NW_NODE, 1, &nc.c[3 * nd + 1])) ||
      (rc = bind(mesh, "momentum_diag", NW_NODE, 1, &nc.c[3 * nd + 2])))
    return rc;
  return NW_OK;
}

/* ------------------------------------------------------------------ */
/*  geometry producers                                                 */
/* ------------------------------------------------------------------ */

/* GeometryInteriorAlg for one element block of `npe`-node elements whose
 * sub-control surfaces pair local nodes lr[2 ip], lr[2 ip + 1] (scsIpEdgeOrd
 * is the identity for Hex8 and Quad4) */
enum { NW_GEO_HEX8 = 0, NW_GEO_QUAD4 = 1, NW_GEO_TET4 = 2, NW_GEO_WED6 = 3, NW_GEO_PYR5 = 4 };

static int
geometry_interior(
  nw_mesh* mesh, const char* what, int topo, int ndim, int npe, int nScs, const int* lr,
  int64_t n_elems, const int32_t* elem_nodes, const unsigned char* elem_owned,
  int coordinates_field, int dual_nodal_volume_field, int edge_area_vector_field)
{
  if (!mesh || n_elems < 0 || (n_elems > 0 && !elem_nodes))
    return fail(NW_ERR_ARG, std::string(what) + ": bad argument");
  if (int rc = need_device(mesh->ctx, what))
    return rc;
  const MeshPlan& mp = mesh->plan;
  if (mp.ndim != ndim)
    return fail(NW_ERR_ARG, std::string(what) + ": wrong spatial dimension");
  nw_field_t* xf = get_field(mesh, coordinates_field);
  nw_field_t* vf =
    dual_nodal_volume_field >= 0 ? get_field(mesh, dual_nodal_volume_field) : nullptr;
  nw_field_t* af =
    edge_area_vector_field >= 0 ? get_field(mesh, edge_area_vector_field) : nullptr;
  if (!xf || xf->rank != NW_NODE || xf->ncomp != ndim ||
      (dual_nodal_volume_field >= 0 &&
       (!vf || vf->rank != NW_NODE || vf->ncomp != 1)) ||
      (edge_area_vector_field >= 0 &&
       (!af || af->rank != NW_EDGE || af->ncomp != ndim)))
    return fail(NW_ERR_ARG, std::string(what) + ": bad field id or shape");
  cudaStream_t s = mesh->ctx->stream;
  /* FNV-1a over the connectivity (+ ownership flags): rebuild on change */
  uint64_t h = 1469598103934665603ull ^ (uint64_t)npe;
  auto mix = [&](const unsigned char* p, size_t n) {
    for (size_t i = 0; i < n; ++i) {
      h ^= p[i];
      h *= 1099511628211ull;
    }
  };
  mix(reinterpret_cast<const unsigned char*>(elem_nodes),
      sizeof(int32_t) * (size_t)npe * n_elems);
  if (elem_owned)
    mix(elem_owned, (size_t)n_elems);
  nw_mesh::GeoCache& gc = mesh->geo[topo];
  if (gc.nElems != n_elems || gc.hash != h) {
    std::map<std::pair<int32_t, int32_t>, int64_t> edgeOf;
    for (int64_t e = 0; e < mp.nEdges; ++e) {
      const int32_t a = mp.edgeNodes[2 * e], b = mp.edgeNodes[2 * e + 1];
      edgeOf[{std::min(a, b), std::max(a, b)}] = e;
    }
    std::vector<int32_t> slots((size_t)n_elems * npe),
      edges((size_t)n_elems * nScs, -1);
    for (int64_t el = 0; el < n_elems; ++el) {
      const int32_t* en = elem_nodes + (size_t)npe * el;
      for (int n = 0; n < npe; ++n) {
        if (en[n] < 0 || en[n] >= mp.nNodes)
          return fail(NW_ERR_ARG, std::string(what) + ": node index out of range");
        slots[(size_t)npe * el + n] = mp.slotOfNode[en[n]];
      }
      for (int ip = 0; ip < nScs; ++ip) {
        const int32_t nl = en[lr[2 * ip]], nr = en[lr[2 * ip + 1]];
        auto it = edgeOf.find({std::min(nl, nr), std::max(nl, nr)});
        if (it == edgeOf.end())
          continue; /* edge owned by another rank */
        const int64_t e = it->second;
        const int negate = nl == mp.edgeNodes[2 * e] ? 0 : 1;
        edges[(size_t)nScs * el + ip] = 2 * mp.primarySlotOfEdge[e] + negate;
      }
    }
    int rc;
    if ((rc = upload(gc.dElemSlots, slots, s, nullptr)) ||
        (rc = upload(gc.dElemEdges, edges, s, nullptr)))
      return rc;
    gc.hasOwned = elem_owned != nullptr;
    if (elem_owned) {
      std::vector<unsigned char> ow(elem_owned, elem_owned + n_elems);
      if ((rc = upload(gc.dOwned, ow, s, nullptr)))
        return rc;
    }
    NW_CUDA(cudaStreamSynchronize(s));
    gc.nElems = n_elems;
    gc.hash = h;
  }
  const unsigned char* ow = gc.hasOwned ? gc.dOwned.as<unsigned char>() : nullptr;
  double* vol = vf ? vf->buf.as<double>() : nullptr;
  double* area = af ? af->buf.as<double>() : nullptr;
  if (topo == NW_GEO_HEX8)
    NW_CUDA(launch_geometry_hex8(
      n_elems, gc.dElemSlots.as<int32_t>(), gc.dElemEdges.as<int32_t>(), ow,
      xf->buf.as<double>(), xf->stride, vol, area, af ? af->stride : 0, s));
  else if (topo == NW_GEO_QUAD4)
    NW_CUDA(launch_geometry_quad4(
      n_elems, gc.dElemSlots.as<int32_t>(), gc.dElemEdges.as<int32_t>(), ow,
      xf->buf.as<double>(), xf->stride, vol, area, af ? af->stride : 0, s));
  else
    NW_CUDA(launch_geometry_cvfem(
      topo == NW_GEO_TET4 ? geo::TET4 : (topo == NW_GEO_WED6 ? geo::WED6 : geo::PYR5),
      n_elems, gc.dElemSlots.as<int32_t>(), gc.dElemEdges.as<int32_t>(), ow,
      xf->buf.as<double>(), xf->stride, vol, area, af ? af->stride : 0, s));
  if (af)
    NW_CUDA(launch_edge_mirror(
      mesh->dPrimarySlot.as<int32_t>(), mesh->dSecondSlot.as<int32_t>(),
      mp.nEdges, ndim, af->stride, af->buf.as<double>(), s));
  return NW_OK;
}

extern "C" int
nw_geometry_interior_hex8(
  nw_mesh* mesh, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, int coordinates_field,
  int dual_nodal_volume_field, int edge_area_vector_field)
{
  /* HexSCS::lrscv_, include/master_element/Hex8CVFEM.h:292-293 */
  static const int lr[24] = {0, 1, 1, 2, 2, 3, 0, 3, 4, 5, 5, 6,
                             6, 7, 4, 7, 0, 4, 1, 5, 2, 6, 3, 7};
  return geometry_interior(
    mesh, "nw_geometry_interior_hex8", NW_GEO_HEX8, 3, 8, 12, lr, n_elems, elem_nodes,
    elem_owned, coordinates_field, dual_nodal_volume_field,
    edge_area_vector_field);
}

extern "C" int
nw_geometry_interior_quad4(
  nw_mesh* mesh, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, int coordinates_field,
  int dual_nodal_volume_field, int edge_area_vector_field)
{
  /* Quad42DSCS::lrscv_, include/master_element/Quad42DCVFEM.h:250 */
  static const int lr[8] = {0, 1, 1, 2, 2, 3, 0, 3};
  return geometry_interior(
    mesh, "nw_geometry_interior_quad4", NW_GEO_QUAD4, 2, 4, 4, lr, n_elems, elem_nodes,
    elem_owned, coordinates_field, dual_nodal_volume_field,
    edge_area_vector_field);
}

/* Tet4 / Wed6 / Pyr5 blocks: the (left, right) node pair of every sub-control
 * surface comes from the same tables the kernel uses (geometry_cvfem.h) */
template <int T>
static int
geometry_interior_cvfem(
  nw_mesh* mesh, const char* what, int topo, int64_t n_elems,
  const int32_t* elem_nodes, const unsigned char* elem_owned,
  int coordinates_field, int dual_nodal_volume_field, int edge_area_vector_field)
{
  constexpr int nScs = geo::Traits<T>::nScs;
  int lr[2 * nScs];
  for (int ip = 0; ip < nScs; ++ip)
    geo::scs_nodes<T>(ip, &lr[2 * ip], &lr[2 * ip + 1]);
  return geometry_interior(
    mesh, what, topo, 3, geo::Traits<T>::npe, nScs, lr, n_elems, elem_nodes,
    elem_owned, coordinates_field, dual_nodal_volume_field,
    edge_area_vector_field);
}

extern "C" int
nw_geometry_interior_tet4(
  nw_mesh* mesh, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, int coordinates_field,
  int dual_nodal_volume_field, int edge_area_vector_field)
{
  return geometry_interior_cvfem<geo::TET4>(
    mesh, "nw_geometry_interior_tet4", NW_GEO_TET4, n_elems, elem_nodes,
    elem_owned, coordinates_field, dual_nodal_volume_field,
    edge_area_vector_field);
}

extern "C" int
nw_geometry_interior_wed6(
  nw_mesh* mesh, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, int coordinates_field,
  int dual_nodal_volume_field, int edge_area_vector_field)
{
  return geometry_interior_cvfem<geo::WED6>(
    mesh, "nw_geometry_interior_wed6", NW_GEO_WED6, n_elems, elem_nodes,
    elem_owned, coordinates_field, dual_nodal_volume_field,
    edge_area_vector_field);
}

extern "C" int
nw_geometry_interior_pyr5(
  nw_mesh* mesh, int64_t n_elems, const int32_t* elem_nodes,
  const unsigned char* elem_owned, int coordinates_field,
  int dual_nodal_volume_field, int edge_area_vector_field)
{
  return geometry_interior_cvfem<geo::PYR5>(
    mesh, "nw_geometry_interior_pyr5", NW_GEO_PYR5, n_elems, elem_nodes,
    elem_owned, coordinates_field, dual_nodal_volume_field,
    edge_area_vector_field);
}

/* ------------------------------------------------------------------ */
/*  edge algorithms without a linear system                            */
/* ------------------------------------------------------------------ */

extern "C" int
nw_mdot_edge(nw_mesh* mesh, const nw_mdot_opts* opts)
{
  if (!mesh || !opts)
    return fail(NW_ERR_ARG, "nw_mdot_edge: NULL argument");
  if (int rc = need_device(mesh->ctx, "nw_mdot_edge"))
    return rc;
  NodeComps nc;
  EdgeComps ec;
  if (int rc = bind_cont_nodes(mesh, nc))
    return rc;
  if (int rc = bind_edge_common(mesh, ec, false, false))
    return rc;
  const double* out = nullptr;
  if (int rc = bind(mesh, "mass_flow_rate", NW_EDGE, 1, &out))
    return rc;
  NW_CUDA(launch_mdot_tile(
    mesh->dev, nc, ec, const_cast<double*>(out), *opts, mesh->ctx->stream));
  return NW_OK;
}

/* device view of the optional mdot / continuity terms */
static int
bind_cont_extra(nw_mesh* mesh, const nw_mdot_extra_opts* x, ContExtraDev& ex)
{
  ex = ContExtraDev();
  if (!x)
    return fail(NW_ERR_ARG, "extra options: NULL");
  const int nd = mesh->plan.ndim;
  ex.balanced = x->add_balanced_forcing != 0;
  ex.gcl = x->needs_gcl != 0;
  for (int d = 0; d < nd; ++d)
    ex.gravity[d] = x->gravity[d];
  int rc;
  if (ex.balanced)
    if ((rc = bind_id(mesh, x->source_mask_field, NW_NODE, 1, &ex.smask)) ||
        (rc = bind_id(mesh, x->source_field, NW_NODE, nd, ex.src)))
      return rc;
  if (ex.gcl)
    if ((rc = bind_id(mesh, x->edge_face_vel_mag_field, NW_EDGE, 1, &ex.faceVelMag)))
      return rc;
  return NW_OK;
}

extern "C" int
nw_mdot_edge_ext(
  nw_mesh* mesh, const nw_mdot_opts* opts, const nw_mdot_extra_opts* extra)
{
  if (!mesh || !opts)
    return fail(NW_ERR_ARG, "nw_mdot_edge_ext: NULL argument");
  if (int rc = need_device(mesh->ctx, "nw_mdot_edge_ext"))
    return rc;
  NodeComps nc;
  EdgeComps ec;
  ContExtraDev ex;
  int rc;
  if ((rc = bind_cont_nodes(mesh, nc)) ||
      (rc = bind_edge_common(mesh, ec, false, false)) ||
      (rc = bind_cont_extra(mesh, extra, ex)))
    return rc;
  const double* out = nullptr;
  if ((rc = bind(mesh, "mass_flow_rate", NW_EDGE, 1, &out)))
    return rc;
  NW_CUDA(launch_mdot_ext(
    mesh->dev, nc, ec, ex, const_cast<double*>(out), *opts, mesh->ctx->stream));
  return NW_OK;
}

extern "C" int
nw_peclet_edge(nw_mesh* mesh, int viscosity_field, const nw_peclet_opts* opts)
{
  if (!mesh || !opts)
    return fail(NW_ERR_ARG, "nw_peclet_edge: NULL argument");
  if (int rc = need_device(mesh->ctx, "nw_peclet_edge"))
    return rc;
  const int nd = mesh->plan.ndim;
  NodeComps nc;
  int rc;
  if ((rc = bind(mesh, "coordinates", NW_NODE, nd, &nc.c[0])) ||
      (rc = bind(mesh, "velocity", NW_NODE, nd, &nc.c[nd])) ||
      (rc = bind(mesh, "density", NW_NODE, 1, &nc.c[2 * nd])) ||
      (rc = bind_id(mesh, viscosity_field, NW_NODE, 1, &nc.c[2 * nd + 1])))
    return rc;
  const double* out = nullptr;
  if ((rc = bind(mesh, "peclet_factor", NW_EDGE, 1, &out)))
    return rc;
  NW_CUDA(launch_peclet_tile(
    mesh->dev, nc, const_cast<double*>(out), *opts, mesh->ctx->stream));
  return NW_OK;
}

static int node_halo_sum(nw_mesh* mesh, nw_field_t* f);
static P2pDev p2p_next(nw_ctx* ctx, bool fusedPush);
static int p2p_pull_begin(nw_ctx* ctx, cudaStream_t* out);
static int p2p_pull_end(nw_ctx* ctx, cudaEvent_t* objEvent, bool* objPending);
static int node_halo_sum_end(NodeHaloSum* st);
static bool node_halo_overlap_applicable(nw_mesh* mesh);
static int periodic_update(nw_mesh* mesh, nw_field_t* f);

/* NodalGradEdgeAlg kernel + NodalGradAlgDriver::post_work
 * (NodalGradAlgDriver.C:41-71: parallel sum, periodic update).  On several
 * ranks over peer memory the exchange is fused into the kernel: the tiles
 * that own shared nodes run first and store their partial sums straight into
 * the other sharers' windows (the last of them publishes the epoch), the
 * interior tiles are computed while the data travels, and one small kernel
 * adds what has arrived.  The two fields of a pair travel as ONE exchange of
 * 2 x ndim components. */
static int
grad_with_post_work(
  nw_mesh* mesh, int dim1, const NodeComps& nc, const double* vol,
  const EdgeComps& ec, double* const* out, nw_field_t* const* grads, int nGrads)
{
  nw_ctx* ctx = mesh->ctx;
  cudaStream_t s = ctx->stream;
  const bool multi = mesh->plan.nranks > 1;
  int first = 0;
  bool launched = false;
  if (multi) {
    if (int rc = mesh_halo_auto(mesh))
      return rc;
    nw_node_halo& H = mesh->halo;
    const int ncomp = dim1 * mesh->plan.ndim;
    if (node_halo_overlap_applicable(mesh) && ncomp <= 9 &&
        (H.p2pMode == 1 || nGrads == 1)) {
      NodePushDev pd;
      pd.tilePtr = H.dPushTilePtr.as<int32_t>();
      pd.slot = H.dPushSlot.as<int32_t>();
      pd.peer = H.dPushPeer.as<int32_t>();
      pd.dst = H.dPushDst.as<int64_t>();
      pd.nc = ncomp;
      pd.nSendTiles = H.nPushTiles;
      pd.pp = p2p_next(ctx, true);
      NW_CUDA(launch_grad_tile(mesh->dev, dim1, nc, vol, ec, out, s, &pd));
      launched = true;
      if (H.p2pMode == 1) {
        CompPtrs comps;
        for (int c = 0; c < ncomp; ++c)
          comps.c[c] = out[c];
        cudaStream_t ps;
        if (int rc = p2p_pull_begin(ctx, &ps))
          return rc;
        NW_CUDA(launch_p2p_pull_nodal(
          comps, ncomp, H.dRecvIdx.as<int64_t>(), H.nRecvP2p, pd.pp,
          ctx->p2p.async, ps, true));
        for (int k = 0; k < nGrads; ++k)
          if (int rc = p2p_pull_end(ctx, &grads[k]->pullDone, &grads[k]->pullPending))
            return rc;
        first = nGrads; /* every field of the call is summed */
        for (int k = 0; k < nGrads; ++k)
          if (int rc = periodic_update(mesh, grads[k]))
            return rc;
      } else {
        NodeHaloSum st;
        st.mesh = mesh;
        st.f = grads[0];
        st.pp = pd.pp;
        st.mode = 2;
        st.signal = true;
        if (int rc = node_halo_sum_end(&st))
          return rc;
        if (int rc = periodic_update(mesh, grads[0]))
          return rc;
        first = 1;
      }
    }
  }
  if (!launched)
    NW_CUDA(launch_grad_tile(mesh->dev, dim1, nc, vol, ec, out, s));
  for (int k = first; k < nGrads; ++k) {
    if (multi)
      if (int rc = node_halo_sum(mesh, grads[k]))
        return rc;
    if (int rc = periodic_update(mesh, grads[k]))
      return rc;
  }
  return NW_OK;
}

/* Realm::periodic_field_update on one nodal field (compute stream) */
static int
periodic_update(nw_mesh* mesh, nw_field_t* f)
{
  if (mesh->perPtr.size() < 2)
    return NW_OK;
  if (mesh->perMasterMissing)
    return fail(
      NW_ERR_LIMIT, "periodic_field_update: a periodic master lives on another "
                    "rank than its slaves (not supported)");
  cudaStream_t s = mesh->ctx->stream;
  if (!mesh->perUploaded) {
    int rc;
    if ((rc = upload(mesh->dPerPtr, mesh->perPtr, s, &mesh->planBytes)) ||
        (rc = upload(mesh->dPerSlots, mesh->perSlots, s, &mesh->planBytes)))
      return rc;
    mesh->perUploaded = true;
  }
  field_wait_pull(mesh, f); /* after the shared-node sum, as the reference */
  NW_CUDA(launch_periodic_update(
    f->buf.as<double>(), f->stride, f->ncomp, mesh->dPerPtr.as<int32_t>(),
    mesh->dPerSlots.as<int32_t>(), (int)mesh->perPtr.size() - 1, s));
  return NW_OK;
}

extern "C" int
nw_field_periodic_update(nw_mesh* mesh, int field_id)
{
  nw_field_t* f = get_field(mesh, field_id);
  if (!f || f->rank != NW_NODE)
    return fail(NW_ERR_ARG, "nw_field_periodic_update: bad nodal field id");
  if (int rc = need_device(mesh->ctx, "nw_field_periodic_update"))
    return rc;
  return periodic_update(mesh, f);
}

extern "C" int
nw_nodal_grad_edge(nw_mesh* mesh, int phi_field, int grad_field)
{
  if (!mesh)
    return fail(NW_ERR_ARG, "nw_nodal_grad_edge: NULL mesh");
  if (int rc = need_device(mesh->ctx, "nw_nodal_grad_edge"))
    return rc;
  nw_field_t* phi = get_field(mesh, phi_field);
  nw_field_t* grad = get_field(mesh, grad_field);
  const int nd = mesh->plan.ndim;
  if (!phi || !grad || phi->rank != NW_NODE || grad->rank != NW_NODE)
    return fail(NW_ERR_ARG, "nw_nodal_grad_edge: bad field id");
  /* NodalGradEdgeAlg constructor checks, src/ngp_algorithms/NodalGradEdgeAlg.C:24-55 */
  if (!(phi->ncomp == 1 || phi->ncomp == nd) ||
      grad->ncomp != phi->ncomp * nd)
    return fail(
      NW_ERR_ARG,
      "nw_nodal_grad_edge: phi must be a scalar or vector field and grad "
      "must have dim1*ndim components");
  NodeComps nc;
  for (int c = 0; c < phi->ncomp; ++c)
    nc.c[c] = phi->buf.as<double>() + (int64_t)c * phi->stride;
  const double* vol = nullptr;
  EdgeComps ec;
  int rc;
  if ((rc = bind(mesh, "dual_nodal_volume", NW_NODE, 1, &vol)) ||
      (rc = bind_edge_common(mesh, ec, false, false)))
    return rc;
  double* out[9];
  for (int c = 0; c < grad->ncomp; ++c)
    out[c] = grad->buf.as<double>() + (int64_t)c * grad->stride;
  nw_field_t* grads[1] = {grad};
  return grad_with_post_work(mesh, phi->ncomp, nc, vol, ec, out, grads, 1);
}

extern "C" int
nw_nodal_grad_edge_pair(
  nw_mesh* mesh, int phi_a, int grad_a, int phi_b, int grad_b)
{
  if (!mesh)
    return fail(NW_ERR_ARG, "nw_nodal_grad_edge_pair: NULL mesh");
  if (int rc = need_device(mesh->ctx, "nw_nodal_grad_edge_pair"))
    return rc;
  nw_field_t* phi[2] = {get_field(mesh, phi_a), get_field(mesh, phi_b)};
  nw_field_t* grad[2] = {get_field(mesh, grad_a), get_field(mesh, grad_b)};
  const int nd = mesh->plan.ndim;
  for (int k = 0; k < 2; ++k)
    if (!phi[k] || !grad[k] || phi[k]->rank != NW_NODE ||
        grad[k]->rank != NW_NODE || phi[k]->ncomp != 1 || grad[k]->ncomp != nd)
      return fail(
        NW_ERR_ARG, "nw_nodal_grad_edge_pair: phi must be scalar nodal fields "
                    "and grad nodal fields of ndim components");
  if (grad[0] == grad[1])
    return fail(NW_ERR_ARG, "nw_nodal_grad_edge_pair: the two outputs coincide");
  NodeComps nc;
  nc.c[0] = phi[0]->buf.as<double>();
  nc.c[1] = phi[1]->buf.as<double>();
  const double* vol = nullptr;
  EdgeComps ec;
  int rc;
  if ((rc = bind(mesh, "dual_nodal_volume", NW_NODE, 1, &vol)) ||
      (rc = bind_edge_common(mesh, ec, false, false)))
    return rc;
  double* out[9];
  for (int k = 0; k < 2; ++k)
    for (int c = 0; c < nd; ++c)
      out[k * nd + c] = grad[k]->buf.as<double>() + (int64_t)c * grad[k]->stride;
  return grad_with_post_work(mesh, 2, nc, vol, ec, out, grad, 2);
}

/* ------------------------------------------------------------------ */
/*  linear system                                                      */
/* ------------------------------------------------------------------ */

extern "C" int
nw_linsys_create(nw_mesh* mesh, int kind, int num_dof, nw_linsys** out)
{
  if (!mesh || !out)
    return fail(NW_ERR_ARG, "nw_linsys_create: NULL argument");
  if (kind != NW_LINSYS_HYPRE && kind != NW_LINSYS_HYPRE_UVW)
    return fail(NW_ERR_ARG, "nw_linsys_create: unknown kind");
  if (kind == NW_LINSYS_HYPRE && !(num_dof == 1 || num_dof == mesh->plan.ndim))
    return fail(NW_ERR_ARG, "nw_linsys_create: num_dof must be 1 or ndim");
  auto* ls = new nw_linsys;
  ls->sh = std::make_shared<nw_ls_shared>();
  ls->mesh = mesh;
  ls->kind = kind;
  ls->numDof = kind == NW_LINSYS_HYPRE_UVW ? 1 : num_dof;
  ls->nRhs = kind == NW_LINSYS_HYPRE_UVW ? mesh->plan.ndim : 1;
  *out = ls;
  return NW_OK;
}

extern "C" int
nw_linsys_destroy(nw_linsys* ls)
{
  if (ls && ls->mesh->ctx->p2p.pendingEager == ls)
    p2p_complete_pending(ls->mesh->ctx); /* keep the neighbours' protocol intact */
  if (ls && ls->pullDone) {
    cudaEventSynchronize(ls->pullDone);
    if (ls->mesh->ctx->p2p.lastPull == ls->pullDone)
      ls->mesh->ctx->p2p.lastPull = nullptr;
    cudaEventDestroy(ls->pullDone);
    ls->pullDone = nullptr;
  }
  delete ls;
  return NW_OK;
}

extern "C" int
nw_linsys_set_skipped_rows(nw_linsys* ls, const int64_t* rows, int64_t n)
{
  if (!ls || (n > 0 && !rows))
    return fail(NW_ERR_ARG, "nw_linsys_set_skipped_rows: NULL argument");
  if (ls->finalized)
    return fail(NW_ERR_STATE, "nw_linsys_set_skipped_rows: already finalized");
  ls->skipped.assign(rows, rows + n);
  return NW_OK;
}

extern "C" int
nw_linsys_build_edge_to_node_graph(nw_linsys* ls)
{
  if (!ls)
    return fail(NW_ERR_ARG, "nw_linsys_build_edge_to_node_graph: NULL");
  ls->graphBuilt = true;
  return NW_OK;
}

/* the integer plan of a shared graph goes to the device once */
static int
shared_upload(nw_ls_shared& sh, cudaStream_t s)
{
  const Graph& g = sh.g;
  const int64_t rows = g.numRowsLocal();
  int rc;
  if (!sh.uploaded) {
    if (sh.lp.usable) {
      if ((rc = upload(sh.dLsTiles, sh.lp.tiles, s, nullptr)) ||
          (rc = upload(sh.dEntInfo, sh.lp.entInfo, s, nullptr)) ||
          (rc = upload(sh.dEntRhsRow, sh.lp.entRhsRow, s, nullptr)) ||
          (rc = upload(sh.dHe, sh.lp.heEll, s, nullptr)) ||
          (rc = upload(sh.dWarp, sh.lp.sliceOff, s, nullptr)) ||
          (rc = upload(sh.dRuns, sh.lp.entGo, s, nullptr)))
        return rc;
      /* rows the tiles do not write */
      std::vector<uint8_t> isPer(sh.lp.uncoveredRows.size(), 0);
      for (size_t i = 0; i < isPer.size(); ++i) {
        const int64_t r = sh.lp.uncoveredRows[i];
        if (r < g.numRowsOwned)
          isPer[i] = std::binary_search(
            g.periodicRowsOwned.begin(), g.periodicRowsOwned.end(),
            g.iLower + r);
      }
      if ((rc = upload(sh.dUncovered, sh.lp.uncoveredRows, s, nullptr)) ||
          (rc = upload(sh.dUncoveredPeriodic, isPer, s, nullptr)))
        return rc;
    }
    std::vector<int64_t> rowPtr(rows + 1);
    for (int64_t r = 0; r < rows; ++r)
      rowPtr[r] = g.rowPtr(r);
    rowPtr[rows] = g.nnzOwned + g.nnzShared;
    if ((rc = upload(sh.dRowPtr, rowPtr, s, nullptr)))
      return rc;
    std::vector<int64_t> per(g.periodicRowsOwned.size());
    for (size_t i = 0; i < per.size(); ++i)
      per[i] = g.rowStartOwned[g.periodicRowsOwned[i] - g.iLower];
    if ((rc = upload(sh.dPeriodicRows, per, s, nullptr)))
      return rc;
    sh.uploaded = true;
  }
  return NW_OK;
}

static int
linsys_upload(nw_linsys* ls)
{
  nw_mesh* m = ls->mesh;
  cudaStream_t s = m->ctx->stream;
  NW_CUDA(cudaSetDevice(m->ctx->device));
  nw_ls_shared& sh = *ls->sh;
  const Graph& g = sh.g;
  const int64_t nnz = g.nnzOwned + g.nnzShared + ls->nExtra;
  const int64_t rows = g.numRowsLocal();
  NW_CUDA(ls->dValues.alloc(sizeof(double) * (nnz + 2)));
  NW_CUDA(ls->dRhs.alloc(sizeof(double) * (rows * ls->nRhs + 2)));
  ls->dev.values = ls->dValues.as<double>();
  ls->dev.rhs = ls->dRhs.as<double>();
  ls->dev.rhsStride = rows;
  int rc;
  if ((rc = shared_upload(sh, s)))
    return rc;
  if (ls->monoOk) {
    if ((rc = shared_upload(*ls->twin, s)) ||
        (rc = upload(ls->dMonoGo, ls->monoGo, s, nullptr)) ||
        (rc = upload(ls->dMonoRow, ls->monoRow, s, nullptr)) ||
        (rc = upload(ls->dMonoUncovered, ls->monoUncovered, s, nullptr)) ||
        (rc = upload(ls->dMonoUncoveredPer, ls->monoUncoveredPer, s, nullptr)))
      return rc;
  }
  if (sh.lp.usable) {
    ls->dev.tiles = sh.dLsTiles.as<LsTileHdr>();
    ls->dev.entInfo = sh.dEntInfo.as<EntInfo>();
    ls->dev.entRhsRow = sh.dEntRhsRow.as<int32_t>();
    ls->dev.heEll = sh.dHe.as<uint32_t>();
    ls->dev.sliceOff = sh.dWarp.as<int32_t>();
    ls->dev.entGo = sh.dRuns.as<int32_t>();
    ls->dev.maxTileNnz = (int)sh.lp.maxTileNnz;
    ls->dev.maxTileEnts = (int)sh.lp.maxTileEnts;
    ls->dev.maxTileEll = (int)sh.lp.maxTileEll;
  }
  const int nPartial = 296;
  NW_CUDA(ls->dNormPartial.alloc(sizeof(double) * nPartial * ls->nRhs));
  NW_CUDA(ls->dNormOut.alloc(sizeof(double) * 8));
  NW_CUDA(cudaStreamSynchronize(s));
  return NW_OK;
}

/* Monolithic system on the tile path: find (or build) the node graph of this
 * mesh and check that the ndim-dof graph is its exact blow-up -- row nd r + i
 * of node row r has nd entries per node-row entry, in the same column order,
 * and the nd rows of a node are contiguous.  Anything else (skipped rows that
 * do not cover whole nodes, a graph the builder laid out differently) keeps
 * the atomic kernel. */
static int
build_mono_twin(nw_linsys* ls)
{
  ls->monoOk = false;
  ls->twin.reset();
  const int nd = ls->numDof;
  if (ls->kind != NW_LINSYS_HYPRE || nd < 2)
    return NW_OK;
  /* skipped rows (Dirichlet nodes) become skipped NODE rows of the twin when
   * they cover all dofs of their nodes -- sum_into tests the first dof's row
   * id only (src/HypreLinearSystem.C:2095-2099), applyDirichletBCs lists all
   * of them; a partial list keeps the atomic kernel */
  std::vector<int64_t> key1;
  {
    std::vector<int64_t> sk = ls->skipped;
    std::sort(sk.begin(), sk.end());
    sk.erase(std::unique(sk.begin(), sk.end()), sk.end());
    if (sk.size() % (size_t)nd != 0)
      return NW_OK;
    for (size_t i = 0; i < sk.size(); i += (size_t)nd) {
      if (sk[i] % nd != 0)
        return NW_OK;
      for (int d = 1; d < nd; ++d)
        if (sk[i + d] != sk[i] + d)
          return NW_OK;
      key1.push_back(sk[i] / nd);
    }
  }
  std::shared_ptr<nw_ls_shared> tw;
  for (auto& c : ls->mesh->lsCache)
    if (c->numDof == 1 && c->skipped == key1)
      tw = c;
  if (!tw) {
    tw = std::make_shared<nw_ls_shared>();
    NW_TRY(build_graph(ls->mesh->plan, NW_LINSYS_HYPRE, 1, key1, tw->g);
           build_ls_plan(ls->mesh->plan, tw->g, tw->lp);)
    tw->numDof = 1;
    tw->skipped = key1;
    ls->mesh->lsCache.push_back(tw);
  }
  const Graph& g1 = tw->g;
  const Graph& g3 = ls->sh->g;
  const LsPlan& lp = tw->lp;
  if (!lp.usable || g3.numRowsOwned != nd * g1.numRowsOwned ||
      g3.numRowsShared != nd * g1.numRowsShared)
    return NW_OK;
  auto row3 = [&](int64_t r1) {
    return r1 < g1.numRowsOwned
             ? nd * r1
             : g3.numRowsOwned + nd * (r1 - g1.numRowsOwned);
  };
  /* checked for the rows the tiles write (a periodic slave row is a lone
   * diagonal in both graphs and is left to the row initialisation) */
  /* (the per-tile lists are padded: only the first nEnts of a tile count) */
  std::vector<uint8_t> live(lp.entRhsRow.size(), 0);
  for (const LsTileHdr& lh : lp.tiles)
    for (int32_t e = lh.entPtr; e < lh.entPtr + lh.nEnts; ++e)
      live[(size_t)e] = 1;
  const int64_t nE = (int64_t)lp.entRhsRow.size();
  bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok)
  for (int64_t e = 0; e < nE; ++e) {
    if (!live[(size_t)e])
      continue;
    const int64_t r1 = lp.entRhsRow[(size_t)e];
    const int64_t len1 = g1.rowLen(r1), p1 = g1.rowPtr(r1);
    const int64_t r3 = row3(r1);
    for (int i = 0; i < nd && ok; ++i) {
      if (g3.rowLen(r3 + i) != nd * len1 ||
          g3.rowPtr(r3 + i) != g3.rowPtr(r3) + (int64_t)i * nd * len1) {
        ok = false;
        break;
      }
      const int64_t p3 = g3.rowPtr(r3 + i);
      for (int64_t k = 0; k < len1 && ok; ++k)
        for (int c = 0; c < nd; ++c)
          if (g3.cols[p3 + nd * k + c] != nd * g1.cols[p1 + k] + c)
            ok = false;
    }
  }
  if (!ok)
    return NW_OK;
  const size_t nEnt = lp.entRhsRow.size();
  ls->monoGo.resize(nEnt);
  ls->monoRow.resize(nEnt);
  for (size_t e = 0; e < nEnt; ++e) {
    if (!live[e]) {
      ls->monoRow[e] = 0;
      ls->monoGo[e] = 0;
      continue;
    }
    const int64_t r3 = row3(lp.entRhsRow[e]);
    ls->monoRow[e] = (int32_t)r3;
    ls->monoGo[e] = (int32_t)g3.rowPtr(r3);
  }
  ls->monoUncovered.clear();
  ls->monoUncoveredPer.clear();
  for (int32_t r1 : lp.uncoveredRows)
    for (int i = 0; i < nd; ++i) {
      const int64_t r3 = row3(r1) + i;
      ls->monoUncovered.push_back((int32_t)r3);
      ls->monoUncoveredPer.push_back(
        r3 < g3.numRowsOwned &&
        std::binary_search(
          g3.periodicRowsOwned.begin(), g3.periodicRowsOwned.end(), g3.iLower + r3));
    }
  ls->twin = tw;
  ls->monoOk = true;
  return NW_OK;
}

static int linsys_build_halo(nw_linsys* ls);

extern "C" int
nw_linsys_finalize(nw_linsys* ls)
{
  if (!ls)
    return fail(NW_ERR_ARG, "nw_linsys_finalize: NULL");
  if (!ls->graphBuilt)
    return fail(
      NW_ERR_STATE, "nw_linsys_finalize: buildEdgeToNodeGraph not called");
  if (ls->finalized)
    return NW_OK;
  {
    /* same dofs per node + same skipped rows => same graph, slot map and
     * reduction plan: reuse the instance another system of this mesh built */
    std::vector<int64_t> key = ls->skipped;
    std::sort(key.begin(), key.end());
    key.erase(std::unique(key.begin(), key.end()), key.end());
    std::shared_ptr<nw_ls_shared> hit;
    for (auto& c : ls->mesh->lsCache)
      if (c->numDof == ls->numDof && c->skipped == key)
        hit = c;
    if (hit) {
      ls->sh = hit;
    } else {
      NW_TRY(build_graph(
               ls->mesh->plan, ls->kind, ls->numDof, ls->skipped, ls->sh->g);
             build_ls_plan(ls->mesh->plan, ls->sh->g, ls->sh->lp);)
      ls->sh->numDof = ls->numDof;
      ls->sh->skipped = key;
      ls->mesh->lsCache.push_back(ls->sh);
    }
  }
  if (ls->sh->g.nnzOwned + ls->sh->g.nnzShared >= (int64_t(1) << 31) - 8)
    return fail(
      NW_ERR_LIMIT, "nw_linsys_finalize: more than 2^31 nonzeros per rank");
  if (int rc = build_mono_twin(ls))
    return rc;
  if (ls->mesh->plan.nranks > 1)
    if (int rc = linsys_build_halo(ls))
      return rc;
  if (ls->mesh->ctx->device >= 0)
    if (int rc = linsys_upload(ls))
      return rc;
  ls->finalized = true;
  ls->state = NW_LS_UNSET;
  return NW_OK;
}

extern "C" int
nw_linsys_get_sizes(const nw_linsys* ls, nw_linsys_sizes* out)
{
  if (!ls || !out)
    return fail(NW_ERR_ARG, "nw_linsys_get_sizes: NULL argument");
  if (!ls->finalized)
    return fail(NW_ERR_STATE, "nw_linsys_get_sizes: not finalized");
  const Graph& g = ls->sh->g;
  out->i_lower = g.iLower;
  out->i_upper = g.iUpper;
  out->num_rows_owned = g.numRowsOwned;
  out->num_nonzeros_owned = g.nnzOwned;
  out->num_rows_shared = g.numRowsShared;
  out->num_nonzeros_shared = g.nnzShared;
  out->num_periodic_rows = (int64_t)g.periodicRowsOwned.size();
  out->num_rhs = ls->nRhs;
  out->block = g.block;
  return NW_OK;
}

template <class T>
static void
copy_out(T* dst, const std::vector<T>& v)
{
  if (dst && !v.empty())
    std::memcpy(dst, v.data(), v.size() * sizeof(T));
}

extern "C" int
nw_linsys_get_graph(
  const nw_linsys* ls,
  int64_t* mat_row_start_owned,
  int64_t* mat_row_start_shared,
  int64_t* cols,
  int64_t* rows,
  int64_t* row_indices_shared,
  int64_t* periodic_rows_owned)
{
  if (!ls)
    return fail(NW_ERR_ARG, "nw_linsys_get_graph: NULL");
  if (!ls->finalized)
    return fail(NW_ERR_STATE, "nw_linsys_get_graph: not finalized");
  const Graph& g = ls->sh->g;
  copy_out(mat_row_start_owned, g.rowStartOwned);
  copy_out(mat_row_start_shared, g.rowStartShared);
  copy_out(cols, g.cols);
  copy_out(rows, g.rows);
  copy_out(row_indices_shared, g.rowIndicesShared);
  copy_out(periodic_rows_owned, g.periodicRowsOwned);
  return NW_OK;
}

extern "C" int
nw_linsys_get_edge_slots(
  const nw_linsys* ls, int64_t* slots, int64_t* rhs_rows)
{
  if (!ls)
    return fail(NW_ERR_ARG, "nw_linsys_get_edge_slots: NULL");
  if (!ls->finalized)
    return fail(NW_ERR_STATE, "nw_linsys_get_edge_slots: not finalized");
  copy_out(slots, ls->sh->g.edgeSlots);
  copy_out(rhs_rows, ls->sh->g.edgeRhsRows);
  return NW_OK;
}

/* kind: 0 reads the system (an outstanding eager exchange is completed
 * first), 1 adds to it (an error while the shared rows are on their way: the
 * contribution could not reach their owners any more), 2 load_complete */
static int
ls_ready(nw_linsys* ls, const char* what, int kind = 0)
{
  if (!ls)
    return fail(NW_ERR_ARG, std::string(what) + ": NULL linear system");
  if (int rc = need_device(ls->mesh->ctx, what))
    return rc;
  if (!ls->finalized)
    return fail(
      NW_ERR_STATE,
      std::string(what) + ": finalizeLinearSystem has not been called");
  if (ls->eagerState != 0 && kind == 1)
    return fail(
      NW_ERR_STATE,
      std::string(what) + ": the shared rows of this system have already been "
                          "sent (eager exchange); call nw_linsys_load_complete "
                          "before adding to it");
  if (ls->eagerState == 1 && kind == 0)
    if (int rc = p2p_complete_pending(ls->mesh->ctx))
      return rc;
  if (ls->pullPending) {
    /* the shared-row add of the last loadComplete may still be running on the
     * communication stream: every later use of the system comes after it */
    cudaStreamWaitEvent(ls->mesh->ctx->stream, ls->pullDone, 0);
    ls->pullPending = false;
  }
  return NW_OK;
}

/* turn a lazy zero into real zeros (atomic kernels and sum_into accumulate) */
static int
materialize_zero(nw_linsys* ls)
{
  cudaStream_t s = ls->mesh->ctx->stream;
  const Graph& g = ls->sh->g;
  const int64_t nnz = g.nnzOwned + g.nnzShared + ls->nExtra;
  NW_CUDA(cudaMemsetAsync(ls->dValues.p, 0, sizeof(double) * nnz, s));
  NW_CUDA(cudaMemsetAsync(
    ls->dRhs.p, 0, sizeof(double) * g.numRowsLocal() * ls->nRhs, s));
  /* periodic-slave rows: diagonal 1 (src/HypreLinearSystem.C:1420-1428) */
  const int64_t np = (int64_t)g.periodicRowsOwned.size();
  if (np > 0) {
    /* dPeriodicRows holds the value offsets of those diagonals */
    std::vector<double> ones(np, 1.0);
    if (int rc = ensure_scratch(ls->mesh, sizeof(double) * np))
      return rc;
    NW_CUDA(cudaMemcpyAsync(
      ls->mesh->scratch.p, ones.data(), sizeof(double) * np,
      cudaMemcpyHostToDevice, s));
    NW_CUDA(cudaStreamSynchronize(s));
    /* scatter: dst[idx[i]] += src[i] on zeroed memory */
    NW_CUDA(launch_unpack_add(
      ls->mesh->scratch.as<double>(), ls->sh->dPeriodicRows.as<int64_t>(), np,
      ls->dValues.as<double>(), s));
  }
  ls->state = NW_LS_ACCUM;
  return NW_OK;
}

extern "C" int
nw_linsys_zero(nw_linsys* ls)
{
  if (int rc = ls_ready(ls, "nw_linsys_zero"))
    return rc;
  /* The zero-fill is deferred: the tile kernels write every row exactly once,
   * so a following segmented assembly needs no memset at all. */
  ls->eagerState = 0; /* ls_ready completed an outstanding eager exchange */
  ls->state = NW_LS_LAZY_ZERO;
  return NW_OK;
}

extern "C" int
nw_linsys_set_scatter_mode(nw_linsys* ls, int mode)
{
  if (!ls || (mode != NW_SCATTER_SEGMENTED && mode != NW_SCATTER_ATOMIC))
    return fail(NW_ERR_ARG, "nw_linsys_set_scatter_mode: bad argument");
  ls->mode = mode;
  return NW_OK;
}

extern "C" int
nw_linsys_uses_tile_path(const nw_linsys* ls)
{
  if (!ls || !ls->finalized)
    return 0;
  if (ls->kind == NW_LINSYS_HYPRE && ls->numDof > 1) {
    if (!ls->monoOk)
      return 0;
    if (ls->mesh->ctx->device < 0)
      return 1;
    LsPlanDev ld = ls->dev; /* shared-memory need of the monolithic policy */
    ld.maxTileNnz = (int)ls->twin->lp.maxTileNnz;
    ld.maxTileEnts = (int)ls->twin->lp.maxTileEnts;
    ld.maxTileEll = (int)ls->twin->lp.maxTileEll;
    return ls_tile_fits(ls->mesh->dev, ld, 3) ? 1 : 0;
  }
  return ls->sh->lp.usable ? 1 : 0;
}

static int
build_atomic_map(nw_linsys* ls)
{
  if (ls->atomicBuilt)
    return NW_OK;
  const MeshPlan& mp = ls->mesh->plan;
  const Graph& g = ls->sh->g;
  const int nb = g.block;
  const int64_t S = mp.nTileEdgeSlots;
  std::vector<int32_t> slots(size_t(S) * nb * nb, -1);
  std::vector<int32_t> rows(size_t(S) * nb, -1);
  for (int64_t e = 0; e < mp.nEdges; ++e) {
    const int64_t s = mp.primarySlotOfEdge[e];
    for (int i = 0; i < nb * nb; ++i)
      slots[size_t(s) * nb * nb + i] =
        (int32_t)g.edgeSlots[size_t(e) * nb * nb + i];
    for (int i = 0; i < nb; ++i)
      rows[size_t(s) * nb + i] = (int32_t)g.edgeRhsRows[size_t(e) * nb + i];
  }
  cudaStream_t st = ls->mesh->ctx->stream;
  int rc;
  if ((rc = upload(ls->dASlots, slots, st, nullptr)) ||
      (rc = upload(ls->dARhsRows, rows, st, nullptr)))
    return rc;
  NW_CUDA(cudaStreamSynchronize(st));
  ls->atomicBuilt = true;
  return NW_OK;
}

/* after a tile assembly on a lazily-zeroed system: initialise the rows no tile
 * owns (Dirichlet rows, periodic-slave rows) */
static int
finish_tile_assembly(nw_linsys* ls)
{
  cudaStream_t s = ls->mesh->ctx->stream;
  NW_CUDA(launch_row_init(
    ls->sh->dUncovered.as<int32_t>(), (int)ls->sh->lp.uncoveredRows.size(),
    ls->sh->dRowPtr.as<int64_t>(), ls->sh->dUncoveredPeriodic.as<uint8_t>(),
    ls->dev.values, ls->dev.rhs, ls->dev.rhsStride, ls->nRhs, s));
  if (ls->nExtra > 0)
    NW_CUDA(cudaMemsetAsync(
      ls->dev.values + ls->sh->g.nnzOwned + ls->sh->g.nnzShared, 0,
      sizeof(double) * ls->nExtra, s));
  ls->state = NW_LS_ACCUM;
  return NW_OK;
}

static bool ls_eager_applicable(const nw_linsys* ls);
static int ls_eager_begin(nw_linsys* ls, LsPushDev* out);
static bool ls_fused_push_possible(const nw_linsys* ls);
static int ls_eager_push_separately(nw_linsys* ls);

/* One tile assembly: launch(MeshPlanDev, LsPlanDev) runs the policy's tile
 * kernel.  Eager exchange (several ranks, peer memory,
 * nw_linsys_set_eager_exchange): the tiles that own shared or receiving rows
 * come first in the launch and store the shared tail straight into the
 * owners' windows; the neighbours' rows travel while the interior tiles are
 * assembled and nw_linsys_load_complete only adds them. */
template <class Launch>
static int
tile_assembly(nw_linsys* ls, Launch&& launch)
{
  nw_mesh* mesh = ls->mesh;
  if (!ls_eager_applicable(ls)) {
    NW_CUDA(launch(mesh->dev, ls->dev));
    return finish_tile_assembly(ls);
  }
  if (!ls_fused_push_possible(ls)) {
    NW_CUDA(launch(mesh->dev, ls->dev));
    if (int rc = finish_tile_assembly(ls))
      return rc;
    return ls_eager_push_separately(ls);
  }
  LsPlanDev lpd = ls->dev;
  if (int rc = ls_eager_begin(ls, &lpd.push))
    return rc;
  NW_CUDA(launch(mesh->dev, lpd));
  return finish_tile_assembly(ls);
}

/* decide the path for an edge assembly; returns 1 for the tile kernel.
 * policy: the kernel's policy id (ls_tile_fits): a user-chosen tile size or a
 * high-valence mesh whose tile does not fit one CTA's shared memory takes the
 * atomic path instead of failing at launch */
static int
use_tile_path(nw_linsys* ls, bool needsDiagExtract, int* rcOut, int policy)
{
  *rcOut = NW_OK;
  const bool tile = ls->mode == NW_SCATTER_SEGMENTED && ls->sh->lp.usable &&
                    ls->state == NW_LS_LAZY_ZERO && !needsDiagExtract &&
                    ls_tile_fits(ls->mesh->dev, ls->dev, policy);
  if (tile)
    return 1;
  if (ls->state != NW_LS_ACCUM)
    if ((*rcOut = materialize_zero(ls)))
      return 0;
  *rcOut = build_atomic_map(ls);
  return 0;
}

extern "C" int
nw_assemble_continuity_edge(nw_linsys* ls, const nw_continuity_opts* opts)
{
  if (int rc = ls_ready(ls, "nw_assemble_continuity_edge", 1))
    return rc;
  if (!opts)
    return fail(NW_ERR_ARG, "nw_assemble_continuity_edge: NULL options");
  if (ls->numDof != 1 || ls->kind != NW_LINSYS_HYPRE)
    return fail(
      NW_ERR_ARG, "nw_assemble_continuity_edge: needs a 1-dof hypre system");
  nw_mesh* mesh = ls->mesh;
  NodeComps nc;
  EdgeComps ec;
  int rc;
  if ((rc = bind_cont_nodes(mesh, nc)) ||
      (rc = bind_edge_common(mesh, ec, false, false)))
    return rc;
  cudaStream_t s = mesh->ctx->stream;
  if (use_tile_path(ls, false, &rc, 0)) {
    return tile_assembly(ls, [&](const MeshPlanDev& md, const LsPlanDev& ld) {
      return launch_continuity_tile(md, ld, nc, ec, *opts, s);
    });
  }
  if (rc)
    return rc;
  AtomicMapDev am{ls->dASlots.as<int32_t>(), ls->dARhsRows.as<int32_t>()};
  NW_CUDA(launch_continuity_atomic(mesh->dev, ls->dev, am, nc, ec, *opts, s));
  return NW_OK;
}

extern "C" int
nw_assemble_continuity_edge_ext(
  nw_linsys* ls, const nw_continuity_opts* opts, const nw_mdot_extra_opts* extra)
{
  if (int rc = ls_ready(ls, "nw_assemble_continuity_edge_ext", 1))
    return rc;
  if (!opts)
    return fail(NW_ERR_ARG, "nw_assemble_continuity_edge_ext: NULL options");
  if (ls->numDof != 1 || ls->kind != NW_LINSYS_HYPRE)
    return fail(
      NW_ERR_ARG, "nw_assemble_continuity_edge_ext: needs a 1-dof hypre system");
  nw_mesh* mesh = ls->mesh;
  NodeComps nc;
  EdgeComps ec;
  ContExtraDev ex;
  int rc;
  if ((rc = bind_cont_nodes(mesh, nc)) ||
      (rc = bind_edge_common(mesh, ec, false, false)) ||
      (rc = bind_cont_extra(mesh, extra, ex)))
    return rc;
  if (ls->state != NW_LS_ACCUM)
    if ((rc = materialize_zero(ls)))
      return rc;
  if ((rc = build_atomic_map(ls)))
    return rc;
  AtomicMapDev am{ls->dASlots.as<int32_t>(), ls->dARhsRows.as<int32_t>()};
  NW_CUDA(launch_continuity_ext_atomic(
    mesh->dev, ls->dev, am, nc, ec, ex, *opts, mesh->ctx->stream));
  return NW_OK;
}

extern "C" int
nw_assemble_scalar_edge(
  nw_linsys* ls,
  int q_field,
  int dqdx_field,
  int diff_flux_coeff_field,
  const nw_scalar_opts* opts)
{
  if (int rc = ls_ready(ls, "nw_assemble_scalar_edge", 1))
    return rc;
  if (!opts)
    return fail(NW_ERR_ARG, "nw_assemble_scalar_edge: NULL options");
  if (ls->numDof != 1 || ls->kind != NW_LINSYS_HYPRE)
    return fail(
      NW_ERR_ARG, "nw_assemble_scalar_edge: needs a 1-dof hypre system");
  nw_mesh* mesh = ls->mesh;
  const int nd = mesh->plan.ndim;
  NodeComps nc;
  EdgeComps ec;
  int rc;
  /* x, vrtm, dqdx, q, rho, dflux */
  if ((rc = bind(mesh, "coordinates", NW_NODE, nd, &nc.c[0])) ||
      (rc = bind(mesh, "velocity", NW_NODE, nd, &nc.c[nd])) ||
      (rc = bind_id(mesh, dqdx_field, NW_NODE, nd, &nc.c[2 * nd])) ||
      (rc = bind_id(mesh, q_field, NW_NODE, 1, &nc.c[3 * nd])) ||
      (rc = bind(mesh, "density", NW_NODE, 1, &nc.c[3 * nd + 1])) ||
      (rc = bind_id(mesh, diff_flux_coeff_field, NW_NODE, 1, &nc.c[3 * nd + 2])) ||
      (rc = bind_edge_common(mesh, ec, true, false)))
    return rc;
  cudaStream_t s = mesh->ctx->stream;
  if (use_tile_path(ls, false, &rc, 1)) {
    return tile_assembly(ls, [&](const MeshPlanDev& md, const LsPlanDev& ld) {
      return launch_scalar_tile(md, ld, nc, ec, *opts, s);
    });
  }
  if (rc)
    return rc;
  AtomicMapDev am{ls->dASlots.as<int32_t>(), ls->dARhsRows.as<int32_t>()};
  NW_CUDA(launch_scalar_atomic(mesh->dev, ls->dev, am, nc, ec, *opts, s));
  return NW_OK;
}

extern "C" int
nw_assemble_scalar_edge_pair(
  nw_linsys* la, int q_a, int dqdx_a, int dflux_a, const nw_scalar_opts* oa,
  nw_linsys* lb, int q_b, int dqdx_b, int dflux_b, const nw_scalar_opts* ob)
{
  if (int rc = ls_ready(la, "nw_assemble_scalar_edge_pair", 1))
    return rc;
  if (int rc = ls_ready(lb, "nw_assemble_scalar_edge_pair", 1))
    return rc;
  if (!oa || !ob)
    return fail(NW_ERR_ARG, "nw_assemble_scalar_edge_pair: NULL options");
  if (la == lb || la->mesh != lb->mesh)
    return fail(
      NW_ERR_ARG, "nw_assemble_scalar_edge_pair: needs two distinct systems of "
                  "one mesh");
  for (nw_linsys* ls : {la, lb})
    if (ls->numDof != 1 || ls->kind != NW_LINSYS_HYPRE)
      return fail(
        NW_ERR_ARG, "nw_assemble_scalar_edge_pair: needs 1-dof hypre systems");
  nw_mesh* mesh = la->mesh;
  const int nd = mesh->plan.ndim;
  /* The fused kernel stages 17 node components + both systems' results for
   * one tile: ~150 KB of shared memory, one 512-thread CTA per SM.  Measured
   * (profiles/r02f_bench_sst*.detail.txt, 128^3): 0.691 ms against 2 x 0.312
   * ms for the two single launches -- with one CTA per SM nothing covers the
   * staging latency of a tile, which costs more than the shared staging
   * saves.  It therefore runs only on request (NW_SCALAR_PAIR_FUSED=1); the
   * default is the two assemblies in turn. */
  static const bool wantFused = [] {
    const char* e = getenv("NW_SCALAR_PAIR_FUSED");
    return e && e[0] == '1';
  }();
  const bool fused = wantFused && la->sh == lb->sh && la->sh->lp.usable &&
                     la->mode == NW_SCATTER_SEGMENTED &&
                     lb->mode == NW_SCATTER_SEGMENTED &&
                     la->state == NW_LS_LAZY_ZERO && lb->state == NW_LS_LAZY_ZERO;
  if (fused) {
    NodeComps nc;
    EdgeComps ec;
    int rc;
    /* x, vrtm, rho, then per system q, dqdx, dflux */
    const int ca = 2 * nd + 1, cb = ca + nd + 2;
    if ((rc = bind(mesh, "coordinates", NW_NODE, nd, &nc.c[0])) ||
        (rc = bind(mesh, "velocity", NW_NODE, nd, &nc.c[nd])) ||
        (rc = bind(mesh, "density", NW_NODE, 1, &nc.c[2 * nd])) ||
        (rc = bind_id(mesh, q_a, NW_NODE, 1, &nc.c[ca])) ||
        (rc = bind_id(mesh, dqdx_a, NW_NODE, nd, &nc.c[ca + 1])) ||
        (rc = bind_id(mesh, dflux_a, NW_NODE, 1, &nc.c[ca + 1 + nd])) ||
        (rc = bind_id(mesh, q_b, NW_NODE, 1, &nc.c[cb])) ||
        (rc = bind_id(mesh, dqdx_b, NW_NODE, nd, &nc.c[cb + 1])) ||
        (rc = bind_id(mesh, dflux_b, NW_NODE, 1, &nc.c[cb + 1 + nd])) ||
        (rc = bind_edge_common(mesh, ec, true, false)))
      return rc;
    bool launched = false;
    NW_CUDA(launch_scalar_pair_tile(
      mesh->dev, la->dev, lb->dev.values, lb->dev.rhs, nc, ec, *oa, *ob,
      &launched, mesh->ctx->stream));
    if (launched) {
      if (int rc2 = finish_tile_assembly(la))
        return rc2;
      return finish_tile_assembly(lb);
    }
  }
  if (int rc = nw_assemble_scalar_edge(la, q_a, dqdx_a, dflux_a, oa))
    return rc;
  return nw_assemble_scalar_edge(lb, q_b, dqdx_b, dflux_b, ob);
}

extern "C" int
nw_assemble_momentum_edge(
  nw_linsys* ls, int viscosity_field, const nw_momentum_opts* opts)
{
  if (int rc = ls_ready(ls, "nw_assemble_momentum_edge", 1))
    return rc;
  if (!opts)
    return fail(NW_ERR_ARG, "nw_assemble_momentum_edge: NULL options");
  nw_mesh* mesh = ls->mesh;
  const int nd = mesh->plan.ndim;
  const bool uvw = ls->kind == NW_LINSYS_HYPRE_UVW;
  if (!uvw && ls->numDof != nd)
    return fail(
      NW_ERR_ARG,
      "nw_assemble_momentum_edge: needs a UVW system or numDof == ndim");
  NodeComps nc;
  EdgeComps ec;
  int rc;
  /* x, u, dudx, visc, rho, mask */
  if ((rc = bind(mesh, "coordinates", NW_NODE, nd, &nc.c[0])) ||
      (rc = bind(mesh, "velocity", NW_NODE, nd, &nc.c[nd])) ||
      (rc = bind(mesh, "dudx", NW_NODE, nd * nd, &nc.c[2 * nd])) ||
      (rc = bind_id(mesh, viscosity_field, NW_NODE, 1, &nc.c[2 * nd + nd * nd])) ||
      (rc = bind(mesh, "density", NW_NODE, 1, &nc.c[2 * nd + nd * nd + 1])) ||
      (rc = bind(
         mesh, "abl_wall_no_slip_wall_func_node_mask", NW_NODE, 1,
         &nc.c[2 * nd + nd * nd + 2])) ||
      (rc = bind_edge_common(mesh, ec, true, !opts->fuse_peclet)))
    return rc;
  double* diagOut = nullptr;
  if (opts->diag_field >= 0) {
    const double* dp = nullptr;
    if ((rc = bind_id(mesh, opts->diag_field, NW_NODE, 1, &dp)))
      return rc;
    diagOut = const_cast<double*>(dp);
  }
  cudaStream_t s = mesh->ctx->stream;
  if (opts->has_vof) {
    /* mdot = massFlowRate + has_vof * massVofBalancedFlowRate
     * (src/edge_kernels/MomentumEdgeSolverAlg.C:124-125), summed once per
     * assembly over the tile-edge slots; the VOF kernels read the sum */
    const double* mv = nullptr;
    if ((rc = bind(mesh, "mass_vof_balanced_flow_rate", NW_EDGE, 1, &mv)))
      return rc;
    const int64_t n = mesh->plan.nTileEdgeSlots;
    if (mesh->dVofMdot.bytes < sizeof(double) * (size_t)n)
      NW_CUDA(mesh->dVofMdot.alloc(sizeof(double) * (size_t)n));
    NW_CUDA(launch_edge_sum(ec.mdot, mv, n, mesh->dVofMdot.as<double>(), s));
    ec.mdot = mesh->dVofMdot.as<double>();
  }
  if (uvw) {
    /* extract_diagonal rides on the tile kernel (node-keyed pass) */
    if (use_tile_path(ls, false, &rc, 2)) {
      return tile_assembly(ls, [&](const MeshPlanDev& md, const LsPlanDev& ld) {
        return launch_momentum_uvw_tile(md, ld, nc, ec, *opts, diagOut, s);
      });
    }
    if (rc)
      return rc;
    AtomicMapDev am{ls->dASlots.as<int32_t>(), ls->dARhsRows.as<int32_t>()};
    NW_CUDA(launch_momentum_uvw_atomic(
      mesh->dev, ls->dev, am, nc, ec, *opts, diagOut, s));
    return NW_OK;
  }
  /* monolithic on the tile path: the node graph's plan, ndim rows per node */
  if (ls->monoOk && ls->mode == NW_SCATTER_SEGMENTED &&
      ls->state == NW_LS_LAZY_ZERO) {
    const nw_ls_shared& tw = *ls->twin;
    LsPlanDev ld = ls->dev; /* values, rhs of THIS system */
    ld.tiles = tw.dLsTiles.as<LsTileHdr>();
    ld.entInfo = tw.dEntInfo.as<EntInfo>();
    ld.heEll = tw.dHe.as<uint32_t>();
    ld.sliceOff = tw.dWarp.as<int32_t>();
    ld.entRhsRow = ls->dMonoRow.as<int32_t>();
    ld.entGo = ls->dMonoGo.as<int32_t>();
    ld.maxTileNnz = (int)tw.lp.maxTileNnz;
    ld.maxTileEnts = (int)tw.lp.maxTileEnts;
    ld.maxTileEll = (int)tw.lp.maxTileEll;
    if (ls_tile_fits(mesh->dev, ld, 3)) {
      NW_CUDA(launch_momentum_mono_tile(mesh->dev, ld, nc, ec, *opts, diagOut, s));
      /* rows no tile writes (periodic slaves: diag 1), the COO extension */
      NW_CUDA(launch_row_init(
        ls->dMonoUncovered.as<int32_t>(), (int)ls->monoUncovered.size(),
        ls->sh->dRowPtr.as<int64_t>(), ls->dMonoUncoveredPer.as<uint8_t>(),
        ls->dev.values, ls->dev.rhs, ls->dev.rhsStride, ls->nRhs, s));
      if (ls->nExtra > 0)
        NW_CUDA(cudaMemsetAsync(
          ls->dev.values + ls->sh->g.nnzOwned + ls->sh->g.nnzShared, 0,
          sizeof(double) * ls->nExtra, s));
      ls->state = NW_LS_ACCUM;
      return NW_OK;
    }
  }
  /* otherwise: atomic scatter of the full block */
  if (ls->state != NW_LS_ACCUM)
    if ((rc = materialize_zero(ls)))
      return rc;
  if ((rc = build_atomic_map(ls)))
    return rc;
  NW_CUDA(launch_momentum_mono_atomic(
    mesh->dev, ls->dASlots.as<int32_t>(), ls->dARhsRows.as<int32_t>(),
    ls->dev.values, ls->dev.rhs, nc, ec, *opts, diagOut, s));
  return NW_OK;
}

/* local row (owned, then shared tail) of a global row id, -1 if this rank
 * holds no such row (non-owned rows absent from map_shared_ are skipped,
 * src/HypreLinearSystem.C:2301-2302) */
static int64_t
local_row_of(const Graph& g, int64_t hid)
{
  if (hid >= g.iLower && hid <= g.iUpper)
    return hid - g.iLower;
  auto it = std::lower_bound(
    g.rowIndicesShared.begin(), g.rowIndicesShared.end(), hid);
  if (it == g.rowIndicesShared.end() || *it != hid)
    return -1;
  return g.numRowsOwned + (it - g.rowIndicesShared.begin());
}

/* scatter table of the node algorithms (AssembleNGPNodeSolverAlgorithm's
 * selector + the applier's skipped-row test), built once per system */
static int
build_node_rows(nw_linsys* ls)
{
  if (ls->nodeRowsBuilt)
    return NW_OK;
  nw_mesh* mesh = ls->mesh;
  const MeshPlan& mp = mesh->plan;
  const Graph& g = ls->sh->g;
  const bool uvw = ls->kind == NW_LINSYS_HYPRE_UVW;
  cudaStream_t s = mesh->ctx->stream;
  std::vector<int64_t> rows;
  const int ndof = uvw ? 1 : ls->numDof;
  for (int64_t n = 0; n < mp.nNodes; ++n) {
    if (!mesh->nodeKernelActive[n])
      continue;
    /* the applier tests the skipped-row map with the first dof's row id
     * (src/HypreLinearSystem.C:2095-2099) */
    if (std::binary_search(
          g.skippedRows.begin(), g.skippedRows.end(), mp.nodeHid[n] * ndof))
      continue;
    for (int d = 0; d < ndof; ++d) {
      const int64_t hid = mp.nodeHid[n] * ndof + d;
      const int64_t lr = hid - g.iLower; /* owned by the selector */
      const int64_t a = g.rowStartOwned[lr], len = g.rowStartOwned[lr + 1] - a;
      const int64_t* rc2 = g.cols.data() + a;
      const int64_t* p = std::lower_bound(rc2, rc2 + len, hid);
      if (p == rc2 + len || *p != hid)
        return fail(NW_ERR_STATE, "node algorithm: row without a diagonal entry");
      rows.push_back(mp.slotOfNode[n]);
      rows.push_back(a + (p - rc2));
      rows.push_back(lr);
      rows.push_back(uvw ? -1 : (ls->numDof > 1 ? d : 0));
    }
  }
  ls->nNodeRows = (int64_t)rows.size() / 4;
  if (int rc = upload(ls->dNodeRows, rows, s, nullptr))
    return rc;
  NW_CUDA(cudaStreamSynchronize(s));
  ls->nodeRowsBuilt = true;
  return NW_OK;
}

extern "C" int
nw_assemble_mass_bdf_node(nw_linsys* ls, int kind, const nw_mass_bdf_opts* opts)
{
  if (int rc = ls_ready(ls, "nw_assemble_mass_bdf_node", 1))
    return rc;
  if (!opts || kind < NW_MASS_SCALAR || kind > NW_MASS_CONTINUITY)
    return fail(NW_ERR_ARG, "nw_assemble_mass_bdf_node: bad argument");
  nw_mesh* mesh = ls->mesh;
  const MeshPlan& mp = mesh->plan;
  const Graph& g = ls->sh->g;
  const bool uvw = ls->kind == NW_LINSYS_HYPRE_UVW;
  const int nd = mp.ndim;
  if (kind == NW_MASS_MOMENTUM ? !(uvw || ls->numDof == nd)
                               : (uvw || ls->numDof != 1))
    return fail(
      NW_ERR_ARG, "nw_assemble_mass_bdf_node: kernel / system dof mismatch");
  MassBdfFields F{};
  F.fieldStride = mp.nSlots;
  auto nodal = [&](int id, int ncomp, const double** out) -> int {
    nw_field_t* f = get_field(mesh, id);
    if (!f || f->rank != NW_NODE || f->ncomp != ncomp)
      return fail(
        NW_ERR_ARG, "nw_assemble_mass_bdf_node: bad field id or shape");
    *out = f->buf.as<double>();
    F.fieldStride = f->stride; /* the same for every nodal field of a mesh */
    return NW_OK;
  };
  int rc;
  const int rid[3] = {opts->rho_nm1, opts->rho_n, opts->rho_np1};
  const int vid[3] = {opts->dnv_nm1, opts->dnv_n, opts->dnv_np1};
  const int qid[3] = {opts->q_nm1, opts->q_n, opts->q_np1};
  for (int k = 0; k < 3; ++k) {
    if ((rc = nodal(rid[k], 1, &F.rho[k])) || (rc = nodal(vid[k], 1, &F.dnv[k])))
      return rc;
    if (kind != NW_MASS_CONTINUITY)
      if ((rc = nodal(qid[k], kind == NW_MASS_MOMENTUM ? nd : 1, &F.q[k])))
        return rc;
  }
  if (kind == NW_MASS_MOMENTUM)
    if ((rc = nodal(opts->dpdx, nd, &F.dpdx)))
      return rc;
  cudaStream_t s = mesh->ctx->stream;
  if ((rc = build_node_rows(ls)))
    return rc;
  if (ls->state != NW_LS_ACCUM)
    if ((rc = materialize_zero(ls)))
      return rc;
  NW_CUDA(launch_mass_bdf_node(
    kind, nd, ls->dNodeRows.as<int64_t>(), ls->nNodeRows, F, opts->dt,
    opts->gamma1, opts->gamma2, opts->gamma3, ls->dev.values, ls->dev.rhs,
    ls->dev.rhsStride, s));
  return NW_OK;
}

extern "C" int
nw_assemble_wall_dist_edge(nw_linsys* ls)
{
  if (int rc = ls_ready(ls, "nw_assemble_wall_dist_edge", 1))
    return rc;
  if (ls->numDof != 1 || ls->kind != NW_LINSYS_HYPRE)
    return fail(
      NW_ERR_ARG, "nw_assemble_wall_dist_edge: needs a 1-dof hypre system");
  nw_mesh* mesh = ls->mesh;
  NodeComps nc;
  EdgeComps ec;
  int rc;
  if ((rc = bind(mesh, "coordinates", NW_NODE, mesh->plan.ndim, &nc.c[0])) ||
      (rc = bind_edge_common(mesh, ec, false, false)))
    return rc;
  cudaStream_t s = mesh->ctx->stream;
  if (use_tile_path(ls, false, &rc, 7)) {
    return tile_assembly(ls, [&](const MeshPlanDev& md, const LsPlanDev& ld) {
      return launch_wall_dist_tile(md, ld, nc, ec, s);
    });
  }
  if (rc)
    return rc;
  AtomicMapDev am{ls->dASlots.as<int32_t>(), ls->dARhsRows.as<int32_t>()};
  NW_CUDA(launch_wall_dist_atomic(mesh->dev, ls->dev, am, nc, ec, s));
  return NW_OK;
}

extern "C" int
nw_assemble_wall_dist_node(nw_linsys* ls, int dual_nodal_volume_field)
{
  if (int rc = ls_ready(ls, "nw_assemble_wall_dist_node", 1))
    return rc;
  if (ls->numDof != 1 || ls->kind != NW_LINSYS_HYPRE)
    return fail(
      NW_ERR_ARG, "nw_assemble_wall_dist_node: needs a 1-dof hypre system");
  nw_field_t* f = get_field(ls->mesh, dual_nodal_volume_field);
  if (!f || f->rank != NW_NODE || f->ncomp != 1)
    return fail(NW_ERR_ARG, "nw_assemble_wall_dist_node: bad field");
  int rc;
  if ((rc = build_node_rows(ls)))
    return rc;
  if (ls->state != NW_LS_ACCUM)
    if ((rc = materialize_zero(ls)))
      return rc;
  NW_CUDA(launch_wall_dist_node(
    ls->dNodeRows.as<int64_t>(), ls->nNodeRows, f->buf.as<double>(),
    ls->dev.rhs, ls->mesh->ctx->stream));
  return NW_OK;
}

extern "C" int
nw_linsys_reset_rows(
  nw_linsys* ls, int64_t n_nodes, const int32_t* nodes, double diag_value,
  double rhs_residual)
{
  if (int rc = ls_ready(ls, "nw_linsys_reset_rows"))
    return rc;
  if (n_nodes < 0 || (n_nodes > 0 && !nodes))
    return fail(NW_ERR_ARG, "nw_linsys_reset_rows: bad node list");
  const MeshPlan& mp = ls->mesh->plan;
  const Graph& g = ls->sh->g;
  /* UVW: one matrix row per node, all rhs columns; else numDof rows per node */
  const int nd = ls->kind == NW_LINSYS_HYPRE_UVW ? 1 : ls->numDof;
  std::vector<int64_t> rows;
  for (int64_t i = 0; i < n_nodes; ++i) {
    if (nodes[i] < 0 || nodes[i] >= mp.nNodes)
      return fail(NW_ERR_ARG, "nw_linsys_reset_rows: node index out of range");
    for (int d = 0; d < nd; ++d) {
      const int64_t hid = mp.nodeHid[nodes[i]] * nd + d;
      const int64_t lr = local_row_of(g, hid);
      if (lr < 0)
        continue;
      const int64_t a = g.rowPtr(lr), len = g.rowLen(lr);
      const int64_t* rc = g.cols.data() + a;
      const int64_t* p = std::lower_bound(rc, rc + len, hid);
      rows.push_back(a);
      rows.push_back(len);
      rows.push_back((p != rc + len && *p == hid) ? (p - rc) : -1);
      rows.push_back(lr);
    }
  }
  if (ls->state != NW_LS_ACCUM)
    if (int rc = materialize_zero(ls))
      return rc;
  if (rows.empty())
    return NW_OK;
  cudaStream_t s = ls->mesh->ctx->stream;
  DevBuf d;
  if (int rc = upload(d, rows, s, nullptr))
    return rc;
  NW_CUDA(launch_reset_rows(
    d.as<int64_t>(), (int64_t)rows.size() / 4, diag_value, rhs_residual,
    ls->dev.values, ls->dev.rhs, ls->dev.rhsStride, ls->nRhs, s));
  NW_CUDA(cudaStreamSynchronize(s)); /* the row table is freed on return */
  return NW_OK;
}

extern "C" int
nw_linsys_apply_dirichlet_bcs(
  nw_linsys* ls, int solution_field, int bc_values_field, int64_t n_nodes,
  const int32_t* nodes)
{
  if (int rc = ls_ready(ls, "nw_linsys_apply_dirichlet_bcs"))
    return rc;
  if (n_nodes < 0 || (n_nodes > 0 && !nodes))
    return fail(NW_ERR_ARG, "nw_linsys_apply_dirichlet_bcs: bad node list");
  nw_mesh* mesh = ls->mesh;
  const MeshPlan& mp = mesh->plan;
  const Graph& g = ls->sh->g;
  const bool uvw = ls->kind == NW_LINSYS_HYPRE_UVW;
  const int ncomp = uvw ? ls->nRhs : ls->numDof;
  nw_field_t* sol = get_field(mesh, solution_field);
  nw_field_t* bc = get_field(mesh, bc_values_field);
  if (!sol || !bc || sol->rank != NW_NODE || bc->rank != NW_NODE ||
      sol->ncomp != ncomp || bc->ncomp != ncomp)
    return fail(
      NW_ERR_ARG, "nw_linsys_apply_dirichlet_bcs: solution / bc fields must be "
                  "nodal with one component per dof");
  std::vector<int64_t> rows;
  for (int64_t i = 0; i < n_nodes; ++i) {
    if (nodes[i] < 0 || nodes[i] >= mp.nNodes)
      return fail(
        NW_ERR_ARG, "nw_linsys_apply_dirichlet_bcs: node index out of range");
    const int64_t hid = mp.nodeHid[nodes[i]];
    const int64_t slot = mp.slotOfNode[nodes[i]];
    for (int d = 0; d < ncomp; ++d) {
      /* locally-owned rows only (the reference's selector) */
      const int64_t row = uvw ? hid : hid * ls->numDof + d;
      if (row < g.iLower || row > g.iUpper)
        continue;
      const int64_t lr = row - g.iLower;
      rows.push_back(g.rowStartOwned[lr]);
      rows.push_back(lr);
      rows.push_back(uvw ? d : 0);
      rows.push_back(slot);
      rows.push_back(d);
    }
  }
  if (ls->state != NW_LS_ACCUM)
    if (int rc = materialize_zero(ls))
      return rc;
  if (rows.empty())
    return NW_OK;
  cudaStream_t s = mesh->ctx->stream;
  DevBuf d;
  if (int rc = upload(d, rows, s, nullptr))
    return rc;
  NW_CUDA(launch_dirichlet_rows(
    d.as<int64_t>(), (int64_t)rows.size() / 5, sol->buf.as<double>(),
    bc->buf.as<double>(), sol->stride, ls->dev.values, ls->dev.rhs,
    ls->dev.rhsStride, s));
  NW_CUDA(cudaStreamSynchronize(s));
  return NW_OK;
}

static int
build_dev_graph(nw_linsys* ls)
{
  if (ls->devGraphBuilt)
    return NW_OK;
  cudaStream_t s = ls->mesh->ctx->stream;
  const Graph& g = ls->sh->g;
  int rc;
  if ((rc = upload(ls->dRowStartOwned, g.rowStartOwned, s, nullptr)) ||
      (rc = upload(ls->dRowStartShared, g.rowStartShared, s, nullptr)) ||
      (rc = upload(ls->dRowIndicesShared, g.rowIndicesShared, s, nullptr)) ||
      (rc = upload(ls->dCols, g.cols, s, nullptr)) ||
      (rc = upload(ls->dSkipped, g.skippedRows, s, nullptr)) ||
      (rc = upload(ls->dNodeHid, ls->mesh->plan.nodeHid, s, nullptr)))
    return rc;
  NW_CUDA(cudaStreamSynchronize(s));
  ls->devGraphBuilt = true;
  return NW_OK;
}

extern "C" int
nw_linsys_sum_into(
  nw_linsys* ls,
  int64_t n_entities,
  int nodes_per_entity,
  const int32_t* d_entity_nodes,
  const double* d_lhs,
  const double* d_rhs)
{
  if (int rc = ls_ready(ls, "nw_linsys_sum_into", 1))
    return rc;
  if (n_entities < 0 || nodes_per_entity < 1 || nodes_per_entity > 8 ||
      (n_entities > 0 && (!d_entity_nodes || !d_lhs || !d_rhs)))
    return fail(NW_ERR_ARG, "nw_linsys_sum_into: bad argument");
  int rc;
  if (ls->state != NW_LS_ACCUM)
    if ((rc = materialize_zero(ls)))
      return rc;
  if ((rc = build_dev_graph(ls)))
    return rc;
  const Graph& g = ls->sh->g;
  NW_CUDA(launch_sum_into(
    n_entities, nodes_per_entity, ls->numDof, d_entity_nodes,
    ls->dNodeHid.as<int64_t>(), d_lhs, d_rhs, g.iLower, g.iUpper,
    g.numRowsOwned, g.nnzOwned, ls->dRowStartOwned.as<int64_t>(),
    ls->dRowStartShared.as<int64_t>(), ls->dRowIndicesShared.as<int64_t>(),
    g.numRowsShared, ls->dCols.as<int64_t>(), ls->dSkipped.as<int64_t>(),
    (int64_t)g.skippedRows.size(),
    ls->kind == NW_LINSYS_HYPRE_UVW ? ls->mesh->plan.ndim : 0, ls->dev.values,
    ls->dev.rhs, ls->dev.rhsStride, ls->mesh->ctx->stream));
  return NW_OK;
}

static bool
write_ints(const std::string& path, const std::vector<int64_t>& v, int bytes)
{
  FILE* f = fopen(path.c_str(), "wb");
  if (!f)
    return false;
  bool ok = true;
  if (bytes == 8) {
    ok = v.empty() || fwrite(v.data(), 8, v.size(), f) == v.size();
  } else {
    std::vector<int32_t> w(v.begin(), v.end());
    ok = w.empty() || fwrite(w.data(), 4, w.size(), f) == w.size();
  }
  return fclose(f) == 0 && ok;
}

static bool
write_doubles(const std::string& path, const double* v, size_t n)
{
  FILE* f = fopen(path.c_str(), "wb");
  if (!f)
    return false;
  const bool ok = n == 0 || fwrite(v, 8, n, f) == n;
  return fclose(f) == 0 && ok;
}

extern "C" int
nw_linsys_write_preassembly_files(
  nw_linsys* ls, const char* directory, const char* eq_sys_name,
  int write_counter, int hypre_int_bytes)
{
  if (int rc = ls_ready(ls, "nw_linsys_write_preassembly_files"))
    return rc;
  if (!eq_sys_name || (hypre_int_bytes != 4 && hypre_int_bytes != 8))
    return fail(NW_ERR_ARG, "nw_linsys_write_preassembly_files: bad argument");
  if (ls->state == NW_LS_LAZY_ZERO)
    if (int rc = materialize_zero(ls))
      return rc;
  const Graph& g = ls->sh->g;
  const MeshPlan& mp = ls->mesh->plan;
  cudaStream_t s = ls->mesh->ctx->stream;
  const int64_t nnz = g.nnzOwned + g.nnzShared;
  const int64_t nrows = g.numRowsOwned + g.numRowsShared;
  std::vector<double> vals(nnz), rhs((size_t)nrows * ls->nRhs);
  if (nnz)
    NW_CUDA(cudaMemcpyAsync(
      vals.data(), ls->dev.values, sizeof(double) * nnz, cudaMemcpyDeviceToHost, s));
  for (int d = 0; d < ls->nRhs; ++d)
    if (nrows)
      NW_CUDA(cudaMemcpyAsync(
        rhs.data() + (size_t)d * nrows, ls->dev.rhs + (int64_t)d * ls->dev.rhsStride,
        sizeof(double) * nrows, cudaMemcpyDeviceToHost, s));
  NW_CUDA(cudaStreamSynchronize(s));
  char rank_str[16];
  snprintf(rank_str, sizeof(rank_str), "%05d", mp.rank);
  const std::string dir =
    (directory && *directory) ? std::string(directory) + "/" : std::string();
  const std::string cnt = std::to_string(write_counter);
  const std::string mat =
    dir + eq_sys_name + ".IJM." + cnt + ".mat." + rank_str + ".preassem.";
  const int ib = hypre_int_bytes;
  const std::vector<int64_t> meta = {
    mp.hypreOffsets.back() * g.numDof, g.iLower, g.iUpper, g.nnzOwned,
    g.nnzShared, nnz};
  bool ok = write_ints(mat + "i", g.rows, ib) && write_ints(mat + "j", g.cols, ib) &&
            write_doubles(mat + "v", vals.data(), (size_t)nnz) &&
            write_ints(mat + "meta", meta, ib);
  /* rhs rows: owned rows ascending, then the shared tail (:963-984) */
  std::vector<int64_t> rrows(nrows);
  for (int64_t i = 0; i < g.numRowsOwned; ++i)
    rrows[i] = g.iLower + i;
  for (int64_t i = 0; i < g.numRowsShared; ++i)
    rrows[g.numRowsOwned + i] = g.rowIndicesShared[i];
  const std::vector<int64_t> rmeta = {g.numRowsOwned, g.numRowsShared, nrows};
  const bool uvw = ls->kind == NW_LINSYS_HYPRE_UVW;
  for (int d = 0; d < ls->nRhs && ok; ++d) {
    const std::string vec = dir + eq_sys_name + (uvw ? std::to_string(d) : "") +
                            ".IJV." + cnt + ".rhs." + rank_str + ".preassem.";
    ok = write_ints(vec + "i", rrows, ib) &&
         write_doubles(vec + "v", rhs.data() + (size_t)d * nrows, (size_t)nrows) &&
         write_ints(vec + "meta", rmeta, ib);
  }
  if (!ok)
    return fail(
      NW_ERR_ARG, "nw_linsys_write_preassembly_files: cannot write under '" +
                    dir + "'");
  return NW_OK;
}

extern "C" int
nw_linsys_device_arrays(
  nw_linsys* ls, double** values, double** rhs, int64_t* rhs_stride)
{
  if (int rc = ls_ready(ls, "nw_linsys_device_arrays"))
    return rc;
  if (ls->state == NW_LS_LAZY_ZERO)
    if (int rc = materialize_zero(ls))
      return rc;
  if (values)
    *values = ls->dev.values;
  if (rhs)
    *rhs = ls->dev.rhs;
  if (rhs_stride)
    *rhs_stride = ls->dev.rhsStride;
  return NW_OK;
}

extern "C" int
nw_linsys_get_values(nw_linsys* ls, double* values, double* rhs)
{
  if (int rc = ls_ready(ls, "nw_linsys_get_values"))
    return rc;
  if (ls->state == NW_LS_LAZY_ZERO)
    if (int rc = materialize_zero(ls))
      return rc;
  cudaStream_t s = ls->mesh->ctx->stream;
  const Graph& g = ls->sh->g;
  if (values)
    NW_CUDA(cudaMemcpyAsync(
      values, ls->dev.values,
      sizeof(double) * (g.nnzOwned + g.nnzShared + ls->nExtra),
      cudaMemcpyDeviceToHost, s));
  if (rhs)
    NW_CUDA(cudaMemcpyAsync(
      rhs, ls->dev.rhs, sizeof(double) * g.numRowsLocal() * ls->nRhs,
      cudaMemcpyDeviceToHost, s));
  p2p_queue_error_read(ls->mesh->ctx, s);
  NW_CUDA(cudaStreamSynchronize(s));
  return p2p_error_after_sync(ls->mesh->ctx);
}

extern "C" int
nw_linsys_rhs_norm2(nw_linsys* ls, double* out)
{
  if (int rc = ls_ready(ls, "nw_linsys_rhs_norm2"))
    return rc;
  if (!out)
    return fail(NW_ERR_ARG, "nw_linsys_rhs_norm2: NULL output");
  if (ls->state == NW_LS_LAZY_ZERO)
    if (int rc = materialize_zero(ls))
      return rc;
  cudaStream_t s = ls->mesh->ctx->stream;
  NW_CUDA(launch_norm2(
    ls->dev.rhs, ls->sh->g.numRowsOwned, ls->dev.rhsStride, ls->nRhs,
    ls->dNormPartial.as<double>(), 296, ls->dNormOut.as<double>(), s));
  NW_CUDA(cudaMemcpyAsync(
    out, ls->dNormOut.p, sizeof(double) * ls->nRhs, cudaMemcpyDeviceToHost, s));
  p2p_queue_error_read(ls->mesh->ctx, s);
  NW_CUDA(cudaStreamSynchronize(s));
  return p2p_error_after_sync(ls->mesh->ctx);
}

/* the same summed over all ranks (the nonlinear residual norm the reference
 * prints and reg_test check_norms compares): local deterministic reduction,
 * then one ncclAllReduce of num_rhs doubles */
extern "C" int
nw_linsys_rhs_norm2_global(nw_linsys* ls, double* out)
{
  if (int rc = ls_ready(ls, "nw_linsys_rhs_norm2_global"))
    return rc;
  if (!out)
    return fail(NW_ERR_ARG, "nw_linsys_rhs_norm2_global: NULL output");
  if (ls->state == NW_LS_LAZY_ZERO)
    if (int rc = materialize_zero(ls))
      return rc;
  nw_ctx* ctx = ls->mesh->ctx;
  cudaStream_t s = ctx->stream;
  NW_CUDA(launch_norm2(
    ls->dev.rhs, ls->sh->g.numRowsOwned, ls->dev.rhsStride, ls->nRhs,
    ls->dNormPartial.as<double>(), 296, ls->dNormOut.as<double>(), s));
  if (ls->mesh->plan.nranks > 1) {
    if (!ctx->comm.comm)
      return fail(NW_ERR_COMM, "nw_linsys_rhs_norm2_global: no communicator");
    std::string err;
    if (!comm_allreduce_sum_f64(
          ctx->comm, ls->dNormOut.as<double>(), ls->nRhs, s, err))
      return fail(NW_ERR_COMM, "nw_linsys_rhs_norm2_global: " + err);
  }
  NW_CUDA(cudaMemcpyAsync(
    out, ls->dNormOut.p, sizeof(double) * ls->nRhs, cudaMemcpyDeviceToHost, s));
  p2p_queue_error_read(ls->mesh->ctx, s);
  NW_CUDA(cudaStreamSynchronize(s));
  return p2p_error_after_sync(ls->mesh->ctx);
}

/* ------------------------------------------------------------------ */
/*  multi-rank halo (see nw_halo.cu)                                   */
/* ------------------------------------------------------------------ */

#include "nw_halo.inc"
