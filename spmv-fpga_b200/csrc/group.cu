// Multi-GPU behind the C ABI (include/spmvb.h, spmvb_group_*): the reference's compute-unit dimension mapped to the
// GPUs of one box.  The reference runs all CUs inside one spmv() call, every CU with a private copy of x
// (src/spmv.cpp:249-294; dispatch src/csr_hw_wrapper.cpp:3-80, 202-271); here every GPU owns a contiguous range of
// rows balanced by non-zero count (the S1/S2/S3 rule of csr_hw.cpp:459-468 applied to whole rows: SURVEY 8e mapping
// A), builds the hw_matrix layout of its own rows, keeps all of x and produces its slice of y.  A single SpMV needs no
// collective.  The iterated caller (power iteration, BASELINE configs[4]) exchanges the y slices into every GPU's x
// once per iteration: ONE grouped NCCL call (a broadcast per row owner inside ncclGroupStart/End - the slices are
// balanced by non-zeros, so their lengths differ and an all-gather would have to pad) plus a one-scalar all-reduce
// for the norm; the scale kernel writes the normalised slice straight into the owner's place in x, so the exchange is
// in place.
//
// Two ways to form a group, same code underneath: spmvb_group_create drives n GPUs from ONE process (what a -DCU=8
// program of the reference becomes: include/spmv_fpga_compat.h with SPMVB_DEVICES), spmvb_group_create_rank makes this
// process one rank of a multi-process group (one process per GPU, e.g. under torchrun; the NCCL unique id travels
// through the launcher).  NCCL is loaded with dlopen on first use: libspmvb.so itself does not depend on it.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <dlfcn.h>
#include <nccl.h>
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spmvb.h"
#include "errors.h"

namespace spmvb {

// ---- the exchange of the iterated caller over peer memory (NVLink): the normalisation kernel IS the first half of
// the collective.  Every GPU scales its rows of y by 1 / ||y|| (the all-reduced sum of squares is on the device) and
// stores them straight into x of a peer instead of into its own memory first:
//   mode 1 "peer all":    into every GPU's x (one kernel, W stores per element; the biggest row owner's NVLink egress
//                         is the bound: its rows x (W - 1) peers)
//   mode 2 "peer + gather": into x of the ONE GPU that forwards that part of the vector - x is also cut into W equal
//                         chunks, GPU f forwards chunk f - and an all-gather of the equal chunks, in place in x, does
//                         the rest (NCCL, NVLS multicast on NVSwitch).  Every row crosses NVLink twice, but both steps
//                         are balanced whatever the row ownership looks like (nnz-balanced R-MAT: one GPU owns 43 %
//                         of the rows).
// peers[g] = base of GPU g's x (peer-mapped or CUDA-IPC-opened pointers); r0 = first row of this GPU.
template <typename VT, int MODE>
__global__ void __launch_bounds__(256) scale_store_peers_kernel(const VT *__restrict__ src, uint32_t n, uint32_t r0,
                                                                const double *__restrict__ sumsq, VT *const *__restrict__ peers,
                                                                int world, uint32_t chunk) {
  const double ss = *sumsq;
  const VT s = (VT)(ss > 0.0 ? 1.0 / sqrt(ss) : 0.0);
  constexpr int V = 16 / sizeof(VT);  // elements per 16-byte store; r0 and chunk are multiples of V
  const uint32_t nv = n / V;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
    uint4 raw = reinterpret_cast<const uint4 *>(src)[i];
    VT *e = reinterpret_cast<VT *>(&raw);
#pragma unroll
    for (int k = 0; k < V; k++) e[k] *= s;
    const size_t g = (size_t)r0 + i * V;  // global row = element of x
    if (MODE == 1) {
      for (int p = 0; p < world; p++) reinterpret_cast<uint4 *>(peers[p] + g)[0] = raw;
    } else {
      reinterpret_cast<uint4 *>(peers[g / chunk] + g)[0] = raw;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < n - nv * V) {  // tail shorter than one vector
    const size_t i = (size_t)nv * V + threadIdx.x;
    const VT v = src[i] * s;
    const size_t g = (size_t)r0 + i;
    if (MODE == 1) {
      for (int p = 0; p < world; p++) peers[p][g] = v;
    } else {
      peers[g / chunk][g] = v;
    }
  }
}

namespace {

struct Nccl {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

Nccl *nccl() {
  static Nccl n;
  static bool tried = false;
  if (tried) return &n;
  tried = true;
  // a process that already holds NCCL (torch brings its own) gets that copy back: same soname
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (n.handle) break;
  }
  if (!n.handle) { n.error = std::string("NCCL is not available: ") + dlerror(); return &n; }
  auto sym = [&](const char *s) -> void * {
    void *p = dlsym(n.handle, s);
    if (!p && n.error.empty()) n.error = std::string("NCCL symbol missing: ") + s;
    return p;
  };
  n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
  n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
  n.CommInitAll = (decltype(n.CommInitAll))sym("ncclCommInitAll");
  n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
  n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
  n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
  n.AllReduce = (decltype(n.AllReduce))sym("ncclAllReduce");
  n.Broadcast = (decltype(n.Broadcast))sym("ncclBroadcast");
  n.AllGather = (decltype(n.AllGather))sym("ncclAllGather");
  n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
  return &n;
}

struct Member {  // one GPU of this process
  int rank = 0, device = 0;
  spmvb_layout *layout = nullptr;
  spmvb_engine *engine = nullptr;
  bool engine_borrowed = false;  // spmvb_group_adopt_engine: the caller keeps the engine
  ncclComm_t comm = nullptr;
  double *d_scalar = nullptr;
  double *d_token = nullptr;   // 8 bytes all-reduced as a barrier between the peer stores and whoever reads x next
  void **d_peers = nullptr;    // [world] base of every GPU's x as seen from this GPU
  std::vector<void *> ipc_opened;  // pointers this process opened with cudaIpcOpenMemHandle
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ph[5] = {};      // phase marks of the last iteration of a power_iter call
};

}  // namespace

struct Group {
  int world = 1, is_double = 1, vb = 8;
  uint32_t rows = 0, cols = 0;           // of the whole matrix
  std::vector<uint32_t> bounds;          // [world + 1] row ownership
  std::vector<Member> local;             // the members this process drives
  double *h_scalar = nullptr;            // pinned
  float last_iter_ms = 0.f;
  float phase_ms[4] = {0.f, 0.f, 0.f, 0.f};  // last iteration: SpMV + sum of squares, norm all-reduce, scale / stores + barrier, gather
  uint64_t nnz_local = 0;
  int exchange = 0;            // 0 NCCL grouped broadcasts, 1 peer stores to all, 2 peer stores to the forwarder + all-gather
  uint32_t chunk = 0;          // mode 2: elements per forwarded chunk of x (world * chunk <= x length)
  // spmv_host: x over the links.  -1 = not decided yet; 1 = every GPU uploads 1/world of x over its own PCIe link and the
  // chunks are all-gathered in place over NVLink (NCCL); 0 = every GPU uploads what its rows read (banded matrices)
  int x_links = -1;
  uint32_t x_chunk = 0;        // elements per uploaded chunk of x (world * x_chunk <= x length of every engine)
};

#define G_CUDA(expr)                                                                       \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) return fail(SPMVB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)
#define G_NCCL(expr)                                                                       \
  do {                                                                                     \
    ncclResult_t _r = (expr);                                                              \
    if (_r != ncclSuccess) return fail(SPMVB_E_CUDA, std::string(#expr) + ": " + nccl()->GetErrorString(_r)); \
  } while (0)

static int need_nccl() {
  Nccl *n = nccl();
  if (!n->handle || !n->error.empty()) return fail(SPMVB_E_CUDA, n->error.empty() ? "NCCL is not available" : n->error);
  return SPMVB_OK;
}

static int member_resources(Group *G, Member &m);

// layout + engine of one member from its row slice (row_ptr rebased to 0)
static int member_build(Group *G, Member &m, uint32_t n_rows, const uint64_t *row_ptr, const uint32_t *col_ind,
                        const void *values, int variant) {
  int rc = spmvb_layout_build(n_rows, G->cols, row_ptr, col_ind, values, 1, 1, G->is_double, 0, &m.layout);
  if (rc) return rc;
  rc = spmvb_engine_create(m.layout, m.device, variant, &m.engine);
  if (rc) return rc;
  return member_resources(G, m);
}

// what a member needs next to its engine: scalars, events
static int member_resources(Group *G, Member &m) {
  G_CUDA(cudaSetDevice(m.device));
  G_CUDA(cudaMalloc((void **)&m.d_scalar, 64));
  G_CUDA(cudaMalloc((void **)&m.d_token, 64));
  G_CUDA(cudaMemset(m.d_token, 0, 64));
  G_CUDA(cudaEventCreate(&m.ev0));
  G_CUDA(cudaEventCreate(&m.ev1));
  for (auto &e : m.ph) G_CUDA(cudaEventCreate(&e));
  return SPMVB_OK;
}

// installs the table of peer x pointers on one member and picks the exchange mode
static int member_set_peers(Group *G, Member &m, const std::vector<void *> &x_of) {
  G_CUDA(cudaSetDevice(m.device));
  if (!m.d_peers) G_CUDA(cudaMalloc((void **)&m.d_peers, sizeof(void *) * G->world));
  G_CUDA(cudaMemcpy(m.d_peers, x_of.data(), sizeof(void *) * G->world, cudaMemcpyHostToDevice));
  return SPMVB_OK;
}
static void group_pick_exchange(Group *G, int requested) {
  const uint32_t V = 16u / (uint32_t)G->vb;
  bool aligned = true;  // 16-byte stores: every owner's first row must be a multiple of V
  for (int r = 0; r < G->world; r++) aligned = aligned && G->bounds[r] % V == 0;
  const uint64_t chunk = (((uint64_t)G->rows + G->world - 1) / G->world + V - 1) / V * V;
  G->chunk = (uint32_t)chunk;
  G->exchange = aligned ? requested : 0;
  // the all-gather of mode 2 fills world * chunk elements of every x
  for (Member &m : G->local)
    if (G->exchange == 2 && (uint64_t)G->world * chunk > spmvb_engine_x_len(m.engine)) G->exchange = 1;
}

static void group_destroy(Group *G) {
  if (!G) return;
  for (Member &m : G->local) {
    if (!m.engine && !m.comm && !m.d_scalar && !m.d_peers) {  // never got as far as touching its device (e.g. a bad device index)
      if (m.layout) spmvb_layout_free(m.layout);
      continue;
    }
    cudaSetDevice(m.device);
    if (m.comm && nccl()->CommDestroy) nccl()->CommDestroy(m.comm);
    if (m.engine && !m.engine_borrowed) spmvb_engine_free(m.engine);
    if (m.layout) spmvb_layout_free(m.layout);
    cudaFree(m.d_scalar); cudaFree(m.d_token); cudaFree(m.d_peers);
    for (void *p : m.ipc_opened) cudaIpcCloseMemHandle(p);
    if (m.ev0) cudaEventDestroy(m.ev0);
    if (m.ev1) cudaEventDestroy(m.ev1);
    for (auto &e : m.ph) if (e) cudaEventDestroy(e);
  }
  if (G->h_scalar) cudaFreeHost(G->h_scalar);
  cudaGetLastError();  // a failed creation must not leave its error behind for the next CUDA call of the process
  delete G;
}

}  // namespace spmvb

using namespace spmvb;

extern "C" {

int spmvb_group_unique_id(uint8_t *out128) {
  if (!out128) return fail(SPMVB_E_ARG, "group_unique_id");
  int rc = need_nccl();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  G_NCCL(nccl()->GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return SPMVB_OK;
}

int spmvb_group_create(uint32_t rows, uint32_t cols, const uint64_t *row_ptr, const uint32_t *col_ind,
                       const void *values, int is_double, int n_devices, const int *devices, int variant,
                       spmvb_group **out) {
  if (!out || !row_ptr || n_devices < 1 || rows == 0 || cols == 0) return fail(SPMVB_E_ARG, "group_create");
  *out = nullptr;
  Group *G = new Group();
  G->world = n_devices; G->is_double = is_double ? 1 : 0; G->vb = is_double ? 8 : 4;
  G->rows = rows; G->cols = cols;
  G->bounds.assign((size_t)n_devices + 1, 0);
  int rc = spmvb_partition_rows(rows, row_ptr, n_devices, is_double ? 2 : 4, G->bounds.data());
  // a split that did not fire leaves an owner without rows: fall back to equal row counts
  bool all = rc == SPMVB_OK;
  for (int k = 0; k < n_devices && all; k++) all = G->bounds[k + 1] > G->bounds[k];
  if (!all)
    for (int k = 0; k <= n_devices; k++) G->bounds[k] = (uint32_t)((uint64_t)rows * k / n_devices);
  G->local.resize(n_devices);
  auto build = [&]() -> int {
    for (int k = 0; k < n_devices; k++) {
      Member &m = G->local[k];
      m.rank = k; m.device = devices ? devices[k] : k;
      const uint32_t r0 = G->bounds[k], r1 = G->bounds[k + 1];
      std::vector<uint64_t> rp((size_t)(r1 - r0) + 1);
      for (uint32_t r = r0; r <= r1; r++) rp[r - r0] = row_ptr[r] - row_ptr[r0];
      const size_t j0 = (size_t)row_ptr[r0];
      int r = member_build(G, m, r1 - r0, rp.data(), col_ind ? col_ind + j0 : nullptr,
                           values ? (const uint8_t *)values + j0 * G->vb : nullptr, variant);
      if (r) return r;
      G->nnz_local += row_ptr[r1] - row_ptr[r0];
    }
    G_CUDA(cudaMallocHost((void **)&G->h_scalar, 64));
    if (n_devices > 1) {
      int r = need_nccl();
      if (r) return r;
      std::vector<ncclComm_t> comms(n_devices);
      std::vector<int> devs(n_devices);
      for (int k = 0; k < n_devices; k++) devs[k] = G->local[k].device;
      G_NCCL(nccl()->CommInitAll(comms.data(), n_devices, devs.data()));
      for (int k = 0; k < n_devices; k++) G->local[k].comm = comms[k];
      // peer access between all GPUs of the group: the exchange of the iterated caller stores into the peers' x
      bool p2p = true;
      for (int a = 0; a < n_devices && p2p; a++)
        for (int b = 0; b < n_devices && p2p; b++) {
          if (a == b) continue;
          int can = 0;
          G_CUDA(cudaDeviceCanAccessPeer(&can, devs[a], devs[b]));
          if (!can) { p2p = false; break; }
          G_CUDA(cudaSetDevice(devs[a]));
          cudaError_t e = cudaDeviceEnablePeerAccess(devs[b], 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
          else if (e != cudaSuccess) p2p = false;
        }
      if (p2p) {
        std::vector<void *> x_of(n_devices);
        for (int k = 0; k < n_devices; k++) x_of[k] = spmvb_engine_x_dev(G->local[k].engine);
        for (int k = 0; k < n_devices; k++) {
          int r = member_set_peers(G, G->local[k], x_of);
          if (r) return r;
        }
        group_pick_exchange(G, 2);
      }
    }
    return SPMVB_OK;
  };
  rc = build();
  if (rc) { group_destroy(G); return rc; }
  *out = (spmvb_group *)G;
  return SPMVB_OK;
}

// Multi-process groups: every rank publishes the CUDA IPC handle of its x (64 bytes), the launcher hands all of them
// to every rank, which maps the peers' vectors and switches the exchange to peer stores.
int spmvb_group_ipc_handle(spmvb_group *g, uint8_t *out64) {
  Group *G = (Group *)g;
  if (!G || !out64 || G->local.size() != 1) return fail(SPMVB_E_ARG, "group_ipc_handle: one local GPU per process");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  Member &m = G->local[0];
  G_CUDA(cudaSetDevice(m.device));
  cudaIpcMemHandle_t h;
  G_CUDA(cudaIpcGetMemHandle(&h, spmvb_engine_x_dev(m.engine)));
  memcpy(out64, &h, 64);
  return SPMVB_OK;
}

int spmvb_group_set_peer_handles(spmvb_group *g, const uint8_t *handles, int mode) {
  Group *G = (Group *)g;
  if (!G || !handles || G->local.size() != 1 || mode < 0 || mode > 2) return fail(SPMVB_E_ARG, "group_set_peer_handles");
  Member &m = G->local[0];
  G_CUDA(cudaSetDevice(m.device));
  std::vector<void *> x_of(G->world, nullptr);
  for (int r = 0; r < G->world; r++) {
    if (r == m.rank) { x_of[r] = spmvb_engine_x_dev(m.engine); continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    void *p = nullptr;
    G_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    m.ipc_opened.push_back(p);
    x_of[r] = p;
  }
  int rc = member_set_peers(G, m, x_of);
  if (rc) return rc;
  group_pick_exchange(G, mode);
  return SPMVB_OK;
}

int spmvb_group_set_exchange(spmvb_group *g, int mode) {
  Group *G = (Group *)g;
  if (!G || mode < 0 || mode > 2) return fail(SPMVB_E_ARG, "group_set_exchange");
  if (mode && (G->world < 2 || !G->local[0].d_peers)) return fail(SPMVB_E_ARG, "group_set_exchange: the peers' x vectors are not mapped");
  group_pick_exchange(G, mode);
  return SPMVB_OK;
}
int spmvb_group_exchange(const spmvb_group *g) { return g ? ((const Group *)g)->exchange : -1; }

int spmvb_group_create_rank(uint32_t global_rows, uint32_t cols, const uint32_t *bounds, const uint64_t *row_ptr_local,
                            const uint32_t *col_ind, const void *values, int is_double, int device, int variant,
                            const uint8_t *unique_id128, int rank, int world, spmvb_group **out) {
  if (!out || !bounds || !row_ptr_local || world < 1 || rank < 0 || rank >= world || (world > 1 && !unique_id128))
    return fail(SPMVB_E_ARG, "group_create_rank");
  *out = nullptr;
  if (bounds[0] != 0 || bounds[world] != global_rows) return fail(SPMVB_E_ARG, "group_create_rank: bounds must cover the rows");
  for (int k = 0; k < world; k++)
    if (bounds[k + 1] < bounds[k]) return fail(SPMVB_E_ARG, "group_create_rank: bounds must ascend");
  if (bounds[rank + 1] == bounds[rank]) return fail(SPMVB_E_ARG, "group_create_rank: this rank owns no rows");
  Group *G = new Group();
  G->world = world; G->is_double = is_double ? 1 : 0; G->vb = is_double ? 8 : 4;
  G->rows = global_rows; G->cols = cols;
  G->bounds.assign(bounds, bounds + world + 1);
  G->local.resize(1);
  Member &m = G->local[0];
  m.rank = rank; m.device = device;
  auto build = [&]() -> int {
    const uint32_t n_rows = bounds[rank + 1] - bounds[rank];
    int r = member_build(G, m, n_rows, row_ptr_local, col_ind, values, variant);
    if (r) return r;
    G->nnz_local = row_ptr_local[n_rows];
    G_CUDA(cudaMallocHost((void **)&G->h_scalar, 64));
    if (world > 1) {
      r = need_nccl();
      if (r) return r;
      ncclUniqueId id;
      memcpy(&id, unique_id128, 128);
      G_CUDA(cudaSetDevice(device));
      G_NCCL(nccl()->CommInitRank(&m.comm, world, id, rank));
    }
    return SPMVB_OK;
  };
  int rc = build();
  if (rc) { group_destroy(G); return rc; }
  *out = (spmvb_group *)G;
  return SPMVB_OK;
}

// One rank of a multi-process group around an engine the caller has created (and keeps): for callers that build their
// layout / engine themselves and want the group's collectives (x over the links in spmv_host).  No iterated caller on
// such a group unless the peers' handles are installed as well.
int spmvb_group_adopt_engine(uint32_t global_rows, uint32_t cols, const uint32_t *bounds, spmvb_engine *engine, int is_double,
                             int device, const uint8_t *unique_id128, int rank, int world, spmvb_group **out) {
  if (!out || !bounds || !engine || world < 1 || rank < 0 || rank >= world || (world > 1 && !unique_id128))
    return fail(SPMVB_E_ARG, "group_adopt_engine");
  *out = nullptr;
  if (bounds[0] != 0 || bounds[world] != global_rows) return fail(SPMVB_E_ARG, "group_adopt_engine: bounds must cover the rows");
  for (int k = 0; k < world; k++)
    if (bounds[k + 1] < bounds[k]) return fail(SPMVB_E_ARG, "group_adopt_engine: bounds must ascend");
  Group *G = new Group();
  G->world = world; G->is_double = is_double ? 1 : 0; G->vb = is_double ? 8 : 4;
  G->rows = global_rows; G->cols = cols;
  G->bounds.assign(bounds, bounds + world + 1);
  G->local.resize(1);
  Member &m = G->local[0];
  m.rank = rank; m.device = device; m.engine = engine; m.engine_borrowed = true;
  auto build = [&]() -> int {
    int r = member_resources(G, m);
    if (r) return r;
    G_CUDA(cudaMallocHost((void **)&G->h_scalar, 64));
    if (world > 1) {
      r = need_nccl();
      if (r) return r;
      ncclUniqueId id;
      memcpy(&id, unique_id128, 128);
      G_CUDA(cudaSetDevice(device));
      G_NCCL(nccl()->CommInitRank(&m.comm, world, id, rank));
    }
    return SPMVB_OK;
  };
  int rc = build();
  if (rc) { group_destroy(G); return rc; }
  *out = (spmvb_group *)G;
  return SPMVB_OK;
}

void spmvb_group_free(spmvb_group *g) { group_destroy((Group *)g); }

int spmvb_group_world(const spmvb_group *g) { return g ? ((const Group *)g)->world : 0; }
int spmvb_group_local_count(const spmvb_group *g) { return g ? (int)((const Group *)g)->local.size() : 0; }
int spmvb_group_bounds(const spmvb_group *g, uint32_t *out) {
  const Group *G = (const Group *)g;
  if (!G || !out) return fail(SPMVB_E_ARG, "group_bounds");
  memcpy(out, G->bounds.data(), G->bounds.size() * 4);
  return SPMVB_OK;
}
spmvb_engine *spmvb_group_engine(spmvb_group *g, int local_index) {
  Group *G = (Group *)g;
  if (!G || local_index < 0 || local_index >= (int)G->local.size()) return nullptr;
  return G->local[local_index].engine;
}
int spmvb_group_rank(const spmvb_group *g, int local_index) {
  const Group *G = (const Group *)g;
  if (!G || local_index < 0 || local_index >= (int)G->local.size()) return -1;
  return G->local[local_index].rank;
}
float spmvb_group_last_iter_ms(const spmvb_group *g) { return g ? ((const Group *)g)->last_iter_ms : 0.f; }
int spmvb_group_phase_ms(const spmvb_group *g, float *out4) {
  const Group *G = (const Group *)g;
  if (!G || !out4) return fail(SPMVB_E_ARG, "group_phase_ms");
  for (int k = 0; k < 4; k++) out4[k] = G->phase_ms[k];
  return SPMVB_OK;
}

int spmvb_group_set_x(spmvb_group *g, const void *x_host, uint32_t n) {
  Group *G = (Group *)g;
  if (!G || !x_host) return fail(SPMVB_E_ARG, "group_set_x");
  for (Member &m : G->local) {  // replicated: every GPU gets what its rows can read of x
    int rc = spmvb_engine_set_x(m.engine, x_host, n);
    if (rc) return rc;
  }
  for (Member &m : G->local) {
    int rc = spmvb_engine_sync(m.engine);
    if (rc) return rc;
  }
  return SPMVB_OK;
}

int spmvb_group_get_x(spmvb_group *g, void *x_host, uint32_t n) {
  Group *G = (Group *)g;
  if (!G || !x_host) return fail(SPMVB_E_ARG, "group_get_x");
  Member &m = G->local[0];
  G_CUDA(cudaSetDevice(m.device));
  const uint32_t k = std::min(n, G->cols);
  G_CUDA(cudaMemcpyAsync(x_host, spmvb_engine_x_dev(m.engine), (size_t)k * G->vb, cudaMemcpyDeviceToHost,
                         (cudaStream_t)spmvb_engine_stream(m.engine)));
  return spmvb_engine_sync(m.engine);
}

// Collective, once per group: does every GPU read (almost) all of x?  Then replicating x is an exchange that belongs on
// the NVLink fabric: each GPU uploads 1/world of x over its own PCIe link and ONE in-place ncclAllGather fills every
// GPU's x - 8 x fewer host bytes than 8 full uploads, which is what bounds the end-to-end call on 8 GPUs (round 1:
// the host side of the links carries ~64-75 GB/s whatever the number of GPUs).  A banded matrix, whose row shards read a
// band of x each, keeps the per-GPU uploads of exactly that band.
static int group_decide_x_links(Group *G) {
  if (G->x_links >= 0) return SPMVB_OK;
  const uint32_t V = 16u / (uint32_t)G->vb;
  const uint64_t chunk = (((uint64_t)G->cols + G->world - 1) / G->world + V - 1) / V * V;
  int want = G->world > 1 ? 1 : 0;
  for (Member &m : G->local) {
    if (!m.comm) want = 0;
    if (spmvb_engine_x_upload_bytes(m.engine) * 2 < (uint64_t)G->cols * G->vb) want = 0;       // reads a band only
    if ((uint64_t)G->world * chunk > spmvb_engine_x_len(m.engine)) want = 0;                   // the gather would not fit
  }
  if (G->world > 1 && G->local[0].comm) {  // all ranks must take the same path: minimum over the group
    Nccl *N = nccl();
    for (Member &m : G->local) {
      G_CUDA(cudaSetDevice(m.device));
      const double v = (double)want;
      G_CUDA(cudaMemcpyAsync(m.d_scalar, &v, 8, cudaMemcpyHostToDevice, (cudaStream_t)spmvb_engine_stream(m.engine)));
    }
    G_NCCL(N->GroupStart());
    for (Member &m : G->local)
      G_NCCL(N->AllReduce(m.d_scalar, m.d_scalar, 1, ncclDouble, ncclMin, m.comm, (cudaStream_t)spmvb_engine_stream(m.engine)));
    G_NCCL(N->GroupEnd());
    Member &m0 = G->local[0];
    G_CUDA(cudaSetDevice(m0.device));
    G_CUDA(cudaMemcpyAsync(G->h_scalar, m0.d_scalar, 8, cudaMemcpyDeviceToHost, (cudaStream_t)spmvb_engine_stream(m0.engine)));
    for (Member &m : G->local) {
      int rc = spmvb_engine_sync(m.engine);
      if (rc) return rc;
    }
    want = G->h_scalar[0] > 0.5 ? 1 : 0;
  }
  G->x_chunk = (uint32_t)chunk;
  G->x_links = want;
  if (want)  // nothing ever writes x behind the gathered chunks: clear it once
    for (Member &m : G->local) {
      G_CUDA(cudaSetDevice(m.device));
      const uint64_t end = (uint64_t)G->world * chunk, len = spmvb_engine_x_len(m.engine);
      if (len > end)
        G_CUDA(cudaMemsetAsync((uint8_t *)spmvb_engine_x_dev(m.engine) + end * G->vb, 0, (len - end) * G->vb,
                               (cudaStream_t)spmvb_engine_stream(m.engine)));
    }
  return SPMVB_OK;
}

// spmv_hw over the group: x reaches every GPU (see group_decide_x_links), all kernels run concurrently (one stream per
// GPU), every GPU's slice of y comes back into its place of y_host.  The local members' rows only: in a multi-process
// group every rank fills its own slice of its own y_host.  y_base_row = global row that y_host[0] stands for.
static int group_spmv_host(Group *G, const void *x_host, uint32_t n, void *y_host, uint32_t y_base_row, int accumulate) {
  int rc = group_decide_x_links(G);
  if (rc) return rc;
  // option build_trace: wall-clock time of every stage (with a synchronisation after each: diagnostics only)
  const bool trace = spmvb_get_option("build_trace") > 0;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_mark = now();
  auto stage = [&](const char *what) {
    if (!trace) return;
    for (Member &m : G->local) spmvb_engine_sync(m.engine);
    const double t = now();
    fprintf(stderr, "[group spmv_host rank %d] %-22s %8.3f ms\n", G->local[0].rank, what, t - t_mark);
    t_mark = t;
  };
  if (G->x_links == 1) {
    Nccl *N = nccl();
    const ncclDataType_t dt = G->is_double ? ncclDouble : ncclFloat;
    const uint64_t chunk = G->x_chunk, m_cols = std::min<uint64_t>(n, G->cols);
    for (Member &m : G->local) {  // this GPU's chunk of x over this GPU's link
      G_CUDA(cudaSetDevice(m.device));
      cudaStream_t st = (cudaStream_t)spmvb_engine_stream(m.engine);
      uint8_t *x = (uint8_t *)spmvb_engine_x_dev(m.engine);
      const uint64_t first = (uint64_t)m.rank * chunk, end = first + chunk, up = std::min(end, m_cols);
      if (up > first)
        G_CUDA(cudaMemcpyAsync(x + first * G->vb, (const uint8_t *)x_host + first * G->vb, (size_t)(up - first) * G->vb,
                               cudaMemcpyHostToDevice, st));
      if (up < end)  // zero padding behind a short x (csr_hw.cpp:1478-1481)
        G_CUDA(cudaMemsetAsync(x + std::max(up, first) * G->vb, 0, (size_t)(end - std::max(up, first)) * G->vb, st));
    }
    stage("x chunk upload");
    G_NCCL(N->GroupStart());
    for (Member &m : G->local) {
      uint8_t *x = (uint8_t *)spmvb_engine_x_dev(m.engine);
      G_NCCL(N->AllGather(x + (size_t)m.rank * chunk * G->vb, x, chunk, dt, m.comm, (cudaStream_t)spmvb_engine_stream(m.engine)));
    }
    G_NCCL(N->GroupEnd());
    stage("all-gather of x");
    if (G->local.size() == 1) {  // one GPU per process: kernel(s) + y down through the engine's row-tile pipeline
      Member &m = G->local[0];
      const uint32_t r0 = G->bounds[m.rank];
      if (r0 < y_base_row) return fail(SPMVB_E_ARG, "group_spmv_host: y does not start at the first local row");
      rc = spmvb_engine_spmv_host_x_resident(m.engine, (uint8_t *)y_host + (size_t)(r0 - y_base_row) * G->vb, accumulate);
      stage("SpMV + y down");
      return rc;
    }
    for (Member &m : G->local) {
      rc = spmvb_engine_spmv_dev(m.engine, nullptr, nullptr, 0, nullptr);
      if (rc) return rc;
    }
    stage("SpMV");
  } else {
    for (Member &m : G->local) {  // uploads and kernels of all GPUs are queued before anything is waited for
      rc = spmvb_engine_set_x(m.engine, x_host, n);
      if (rc) return rc;
      rc = spmvb_engine_spmv_dev(m.engine, nullptr, nullptr, 0, nullptr);
      if (rc) return rc;
    }
  }
  for (Member &m : G->local) {
    const uint32_t r0 = G->bounds[m.rank], r1 = G->bounds[m.rank + 1];
    if (r0 < y_base_row) return fail(SPMVB_E_ARG, "group_spmv_host: y does not start at the first local row");
    rc = spmvb_engine_get_y(m.engine, (uint8_t *)y_host + (size_t)(r0 - y_base_row) * G->vb, r1 - r0, accumulate);
    if (rc) return rc;
  }
  stage("y down (+ host add)");
  return SPMVB_OK;
}

int spmvb_group_spmv_host(spmvb_group *g, const void *x_host, uint32_t n, void *y_host, int accumulate) {
  Group *G = (Group *)g;
  if (!G || !x_host || !y_host) return fail(SPMVB_E_ARG, "group_spmv_host");
  return group_spmv_host(G, x_host, n, y_host, 0, accumulate);
}

// the same with y_rows = the rows of this process's GPUs only (y_rows[0] = the first local GPU's first row)
int spmvb_group_spmv_host_rows(spmvb_group *g, const void *x_host, uint32_t n, void *y_rows, int accumulate) {
  Group *G = (Group *)g;
  if (!G || !x_host || !y_rows || G->local.empty()) return fail(SPMVB_E_ARG, "group_spmv_host_rows");
  return group_spmv_host(G, x_host, n, y_rows, G->bounds[G->local[0].rank], accumulate);
}

int spmvb_group_x_over_links(const spmvb_group *g) { return g ? ((const Group *)g)->x_links : -1; }

// y slices of the local members after the last SpMV / iteration, into their places of y_host (diagnostics, tests)
int spmvb_group_get_y(spmvb_group *g, void *y_host) {
  Group *G = (Group *)g;
  if (!G || !y_host) return fail(SPMVB_E_ARG, "group_get_y");
  for (Member &m : G->local) {
    const uint32_t r0 = G->bounds[m.rank], r1 = G->bounds[m.rank + 1];
    int rc = spmvb_engine_get_y(m.engine, (uint8_t *)y_host + (size_t)r0 * G->vb, r1 - r0, 0);
    if (rc) return rc;
  }
  return SPMVB_OK;
}

// x <- A x / ||A x||_2, `iters` times, x replicated on every GPU.  Collective: every rank of a multi-process group
// calls it with the same `iters`.  Per iteration and GPU: clear rows + SpMV kernel, sum of squares, [all-reduce of
// one double], scale kernel that writes the normalised slice into its place of x, [one grouped NCCL exchange].
int spmvb_group_power_iter(spmvb_group *g, int iters, double *norm_out) {
  Group *G = (Group *)g;
  if (!G || iters < 1) return fail(SPMVB_E_ARG, "group_power_iter");
  if (G->rows != G->cols) return fail(SPMVB_E_ARG, "power_iter needs a square matrix");
  Nccl *N = nccl();
  const bool multi = G->world > 1;
  const ncclDataType_t dt = G->is_double ? ncclDouble : ncclFloat;
  for (Member &m : G->local) {
    G_CUDA(cudaSetDevice(m.device));
    G_CUDA(cudaEventRecord(m.ev0, (cudaStream_t)spmvb_engine_stream(m.engine)));
  }
  auto mark = [&](int it, int k) -> int {  // phase marks, last iteration only
    if (it != iters - 1) return SPMVB_OK;
    for (Member &m : G->local) {
      G_CUDA(cudaSetDevice(m.device));
      G_CUDA(cudaEventRecord(m.ph[k], (cudaStream_t)spmvb_engine_stream(m.engine)));
    }
    return SPMVB_OK;
  };
  for (int it = 0; it < iters; it++) {
    if (int rc = mark(it, 0)) return rc;
    for (Member &m : G->local) {
      const uint32_t n_local = G->bounds[m.rank + 1] - G->bounds[m.rank];
      int rc = spmvb_engine_spmv_dev(m.engine, nullptr, nullptr, 0, nullptr);
      if (rc) return rc;
      rc = spmvb_engine_sumsq(m.engine, spmvb_engine_y_dev(m.engine), n_local, m.d_scalar, nullptr);
      if (rc) return rc;
    }
    if (int rc = mark(it, 1)) return rc;
    if (multi) {
      G_NCCL(N->GroupStart());
      for (Member &m : G->local)
        G_NCCL(N->AllReduce(m.d_scalar, m.d_scalar, 1, ncclDouble, ncclSum, m.comm, (cudaStream_t)spmvb_engine_stream(m.engine)));
      G_NCCL(N->GroupEnd());
    }
    if (int rc = mark(it, 2)) return rc;
    if (!multi || G->exchange == 0) {
      for (Member &m : G->local) {
        const uint32_t r0 = G->bounds[m.rank], n_local = G->bounds[m.rank + 1] - r0;
        uint8_t *x = (uint8_t *)spmvb_engine_x_dev(m.engine);
        int rc = spmvb_engine_scale_rsqrt(m.engine, spmvb_engine_y_dev(m.engine), x + (size_t)r0 * G->vb, n_local, m.d_scalar, nullptr);
        if (rc) return rc;
      }
      if (int rc = mark(it, 3)) return rc;
      if (multi) {  // every owner's slice into every GPU's x, in place: one NCCL launch per GPU
        G_NCCL(N->GroupStart());
        for (Member &m : G->local) {
          uint8_t *x = (uint8_t *)spmvb_engine_x_dev(m.engine);
          for (int r = 0; r < G->world; r++) {
            const uint32_t b0 = G->bounds[r], len = G->bounds[r + 1] - b0;
            if (!len) continue;
            void *p = x + (size_t)b0 * G->vb;
            G_NCCL(N->Broadcast(p, p, len, dt, r, m.comm, (cudaStream_t)spmvb_engine_stream(m.engine)));
          }
        }
        G_NCCL(N->GroupEnd());
      }
    } else {
      // normalise + store into the peers' x (NVLink), then a barrier: an 8-byte all-reduce that no rank leaves before
      // every rank's store kernel has finished
      for (Member &m : G->local) {
        const uint32_t r0 = G->bounds[m.rank], n_local = G->bounds[m.rank + 1] - r0;
        G_CUDA(cudaSetDevice(m.device));
        cudaStream_t st = (cudaStream_t)spmvb_engine_stream(m.engine);
        const int grid = 148 * 4;
        if (G->is_double) {
          auto k1 = scale_store_peers_kernel<double, 1>;
          auto k2 = scale_store_peers_kernel<double, 2>;
          (G->exchange == 1 ? k1 : k2)<<<grid, 256, 0, st>>>((const double *)spmvb_engine_y_dev(m.engine), n_local, r0, m.d_scalar,
                                                             (double *const *)m.d_peers, G->world, G->chunk);
        } else {
          auto k1 = scale_store_peers_kernel<float, 1>;
          auto k2 = scale_store_peers_kernel<float, 2>;
          (G->exchange == 1 ? k1 : k2)<<<grid, 256, 0, st>>>((const float *)spmvb_engine_y_dev(m.engine), n_local, r0, m.d_scalar,
                                                             (float *const *)m.d_peers, G->world, G->chunk);
        }
        G_CUDA(cudaGetLastError());
      }
      G_NCCL(N->GroupStart());
      for (Member &m : G->local)
        G_NCCL(N->AllReduce(m.d_token, m.d_token, 1, ncclDouble, ncclSum, m.comm, (cudaStream_t)spmvb_engine_stream(m.engine)));
      G_NCCL(N->GroupEnd());
      if (int rc = mark(it, 3)) return rc;
      if (G->exchange == 2) {  // every GPU forwards its equal chunk of x to all: all-gather in place
        G_NCCL(N->GroupStart());
        for (Member &m : G->local) {
          uint8_t *x = (uint8_t *)spmvb_engine_x_dev(m.engine);
          G_NCCL(N->AllGather(x + (size_t)m.rank * G->chunk * G->vb, x, G->chunk, dt, m.comm, (cudaStream_t)spmvb_engine_stream(m.engine)));
        }
        G_NCCL(N->GroupEnd());
      }
    }
    if (int rc = mark(it, 4)) return rc;
  }
  for (Member &m : G->local) {
    G_CUDA(cudaSetDevice(m.device));
    G_CUDA(cudaEventRecord(m.ev1, (cudaStream_t)spmvb_engine_stream(m.engine)));
  }
  Member &m0 = G->local[0];
  G_CUDA(cudaSetDevice(m0.device));
  G_CUDA(cudaMemcpyAsync(G->h_scalar, m0.d_scalar, sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)spmvb_engine_stream(m0.engine)));
  float worst = 0.f;
  for (Member &m : G->local) {
    int rc = spmvb_engine_sync(m.engine);
    if (rc) return rc;
    float ms = 0.f;
    G_CUDA(cudaSetDevice(m.device));
    if (cudaEventElapsedTime(&ms, m.ev0, m.ev1) == cudaSuccess) worst = std::max(worst, ms);
  }
  G->last_iter_ms = worst / (float)iters;
  for (int k = 0; k < 4; k++) {
    G->phase_ms[k] = 0.f;
    for (Member &m : G->local) {
      float ms = 0.f;
      G_CUDA(cudaSetDevice(m.device));
      if (cudaEventElapsedTime(&ms, m.ph[k], m.ph[k + 1]) == cudaSuccess) G->phase_ms[k] = std::max(G->phase_ms[k], ms);
    }
  }
  if (norm_out) *norm_out = std::sqrt(G->h_scalar[0]);
  return SPMVB_OK;
}

}  // extern "C"
