/* group.cu -- several device contexts driven as ONE mesh from a single host thread (hexed_b200_group_*, include/hexed_b200.h).
 *
 * The reference is a single process that hands its kernels ONE Kernel_mesh (include/Kernel_mesh.hpp:14-25), so a drop-in that
 * wants more than one GPU has to split that mesh itself: the C++ adapter (hexed_b200/host/) partitions the flattened tables,
 * creates one context per device and a group over them; this file is the device side of that: the halo exchange of the cut
 * faces with NCCL send/recv over NVLink, overlapped with the interior flux work, and the global time step as an
 * ncclAllReduce(min).
 *
 * Per stage and rank r (all asynchronous; the host thread only enqueues):
 *     ctx stream r : gather cut faces -> send buffers | Neighbor on interior connections ........... | scatter -> halo slots | rest of the stage
 *     comm stream r:          (waits for the gather)  ncclGroup{ncclRecv + ncclSend per peer}  (event)^
 * The comm stream exists so that the interior Neighbor kernels enqueued on the context's stream AFTER the gather run WHILE the
 * transfer is in flight. One ncclGroupStart/End spans every rank's sends and receives (single-thread multi-device NCCL usage).
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2) so that the single-GPU library has no hard dependency on it; a group cannot be
 * created without it -- there is no fallback transport. The host-thread emulation build (tests only) copies between the "devices"
 * with memcpy at the point where the real build posts the NCCL group.
 */
#include "common.cuh"
#include "group_workers.hpp"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#ifndef HB_EMULATE
#include <dlfcn.h>
#include <nccl.h>
#endif

namespace
{

#ifndef HB_EMULATE
struct Nccl
{
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool load()
  {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) if ((lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!lib) { error = std::string("cannot load NCCL: ") + dlerror(); return false; }
    auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) error = std::string("NCCL symbol missing: ") + n; return p; };
    CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll"); CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart"); GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    Send = (decltype(Send))sym("ncclSend"); Recv = (decltype(Recv))sym("ncclRecv"); AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
    GetVersion = (decltype(GetVersion))sym("ncclGetVersion"); GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!error.empty()) { lib = nullptr; return false; }
    return true;
  }
};
Nccl g_nccl;
#endif

struct Link
{
  int peer = -1;
  int send_list = -1, recv_list = -1; // face lists of the owning context
  int n_send = 0, n_recv = 0;
  double* d_send = nullptr; double* d_recv = nullptr; // [n][widest face kind]
};

} // namespace

struct hexed_b200_group
{
  hb::Workers workers;
  std::mutex err_mutex;
  int n = 0;
  std::vector<hexed_b200_ctx*> ctx;
  std::vector<std::vector<Link>> links;
  std::vector<cudaStream_t> comm_stream;
  std::vector<cudaEvent_t> ev_ready, ev_done;
  std::vector<double*> d_dt; // per rank: local minimum, reduced in place
#ifndef HB_EMULATE
  std::vector<ncclComm_t> comm;
#endif
  int nccl_version = 0;
  long long exchanges = 0, bytes_sent = 0;
  std::string err;
};

namespace
{

std::string g_group_create_error;

int gfail(hexed_b200_group* g, int code, const std::string& msg)
{
  if (g) { std::lock_guard<std::mutex> lk(g->err_mutex); g->err = msg; } else g_group_create_error = msg;
  return code;
}

#define HG_CUDA(g, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return gfail(g, HEXED_B200_CUDA_ERROR, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define HG_CTX(g, r, call) do { int rc_ = (call); if (rc_) return gfail(g, rc_, "rank " + std::to_string(r) + ": " + hexed_b200_last_error((g)->ctx[r])); } while (0)
#ifndef HB_EMULATE
#define HG_NCCL(g, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return gfail(g, HEXED_B200_CUDA_ERROR, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); } while (0)
#endif

int face_width(hexed_b200_ctx* c, int kind) { return (kind == 2 ? c->nd + c->rs : c->nv)*c->nfq; }

const Link* find_link(const hexed_b200_group* g, int rank, int peer)
{
  for (const Link& l : g->links[rank]) if (l.peer == peer) return &l;
  return nullptr;
}

/* gather on every context stream, then one NCCL group on the comm streams (ordered after the gathers by events) */
int exchange_start(hexed_b200_group* g, int kind)
{
  int rc = g->workers.run([&](int r) -> int {
    hexed_b200_ctx* c = g->ctx[r];
    for (const Link& l : g->links[r]) if (l.n_send) HG_CTX(g, r, hexed_b200_face_list_gather(c, l.send_list, kind, l.d_send));
    HG_CUDA(g, cudaSetDevice(c->device));
    HG_CUDA(g, cudaEventRecord(g->ev_ready[r], c->stream));
    HG_CUDA(g, cudaStreamWaitEvent(g->comm_stream[r], g->ev_ready[r], 0));
    return 0;
  });
  if (rc) return rc;
#ifndef HB_EMULATE
  HG_NCCL(g, g_nccl.GroupStart());
  for (int r = 0; r < g->n; ++r) {
    const size_t w = face_width(g->ctx[r], kind);
    for (const Link& l : g->links[r]) {
      if (l.n_recv) HG_NCCL(g, g_nccl.Recv(l.d_recv, l.n_recv*w, ncclDouble, l.peer, g->comm[r], g->comm_stream[r]));
      if (l.n_send) HG_NCCL(g, g_nccl.Send(l.d_send, l.n_send*w, ncclDouble, l.peer, g->comm[r], g->comm_stream[r]));
    }
  }
  HG_NCCL(g, g_nccl.GroupEnd());
#else
  for (int r = 0; r < g->n; ++r) { // emulation (tests only): the "devices" share the host's memory
    const size_t w = face_width(g->ctx[r], kind);
    for (const Link& l : g->links[r]) if (l.n_send) {
      const Link* back = find_link(g, l.peer, r);
      if (!back || back->n_recv != l.n_send) return gfail(g, HEXED_B200_BAD_ARGUMENT, "halo lists of two ranks do not match");
      std::memcpy(back->d_recv, l.d_send, sizeof(double)*l.n_send*w);
    }
  }
#endif
  for (int r = 0; r < g->n; ++r) {
    const size_t w = face_width(g->ctx[r], kind);
    for (const Link& l : g->links[r]) g->bytes_sent += (long long)(l.n_send*w*sizeof(double));
    HG_CUDA(g, cudaSetDevice(g->ctx[r]->device));
    HG_CUDA(g, cudaEventRecord(g->ev_done[r], g->comm_stream[r]));
  }
  ++g->exchanges;
  return 0;
}

/* the context streams wait for the transfer (the host does not) and unpack into the halo slots */
int exchange_finish(hexed_b200_group* g, int kind)
{
  return g->workers.run([&](int r) -> int {
    hexed_b200_ctx* c = g->ctx[r];
    HG_CUDA(g, cudaSetDevice(c->device));
    HG_CUDA(g, cudaStreamWaitEvent(c->stream, g->ev_done[r], 0));
    for (const Link& l : g->links[r]) if (l.n_recv) HG_CTX(g, r, hexed_b200_face_list_scatter(c, l.recv_list, kind, l.d_recv));
    return 0;
  });
}

void free_links(hexed_b200_group* g, int r)
{
  cudaSetDevice(g->ctx[r]->device);
  for (Link& l : g->links[r]) { if (l.d_send) cudaFree(l.d_send); if (l.d_recv) cudaFree(l.d_recv); }
  g->links[r].clear();
}

} // namespace

extern "C" {

const char* hexed_b200_group_last_error(const hexed_b200_group* g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }

int hexed_b200_group_create(hexed_b200_group** out, int n, hexed_b200_ctx* const* ctxs)
{
  *out = nullptr;
  if (n < 1) return gfail(nullptr, HEXED_B200_BAD_ARGUMENT, "a group needs at least one context");
  for (int r = 0; r < n; ++r) {
    if (!ctxs[r]) return gfail(nullptr, HEXED_B200_BAD_ARGUMENT, "null context");
    if (ctxs[r]->nd != ctxs[0]->nd || ctxs[r]->rs != ctxs[0]->rs) return gfail(nullptr, HEXED_B200_BAD_ARGUMENT, "contexts of one group must share n_dim and row_size");
#ifndef HB_EMULATE
    for (int q = 0; q < r; ++q) if (ctxs[q]->device == ctxs[r]->device) return gfail(nullptr, HEXED_B200_BAD_ARGUMENT, "two contexts of a group on the same device");
#endif
  }
  hexed_b200_group* g = new hexed_b200_group();
  g->n = n;
  g->ctx.assign(ctxs, ctxs + n);
  g->workers.start(n);
  g->links.resize(n); g->comm_stream.assign(n, nullptr); g->ev_ready.assign(n, nullptr); g->ev_done.assign(n, nullptr); g->d_dt.assign(n, nullptr);
  auto bail = [&](int code, const std::string& msg) { g_group_create_error = msg; hexed_b200_group_destroy(g); return code; };
  for (int r = 0; r < n; ++r) {
    if (cudaSetDevice(ctxs[r]->device) != cudaSuccess) return bail(HEXED_B200_NO_DEVICE, "cudaSetDevice failed");
    if (cudaStreamCreateWithFlags(&g->comm_stream[r], cudaStreamNonBlocking) != cudaSuccess) return bail(HEXED_B200_CUDA_ERROR, "cannot create the communication stream");
    if (cudaEventCreateWithFlags(&g->ev_ready[r], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&g->ev_done[r], cudaEventDisableTiming) != cudaSuccess)
      return bail(HEXED_B200_CUDA_ERROR, "cannot create events");
    if (cudaMalloc(&g->d_dt[r], sizeof(double)) != cudaSuccess) return bail(HEXED_B200_CUDA_ERROR, "cudaMalloc failed");
  }
#ifndef HB_EMULATE
  if (!g_nccl.load()) return bail(HEXED_B200_NO_DEVICE, g_nccl.error + " (hexed_b200 has no other multi-GPU transport)");
  std::vector<int> devs(n);
  for (int r = 0; r < n; ++r) devs[r] = ctxs[r]->device;
  g->comm.assign(n, nullptr);
  ncclResult_t res = g_nccl.CommInitAll(g->comm.data(), n, devs.data());
  if (res != ncclSuccess) { g->comm.clear(); return bail(HEXED_B200_CUDA_ERROR, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(res)); }
  g_nccl.GetVersion(&g->nccl_version);
#endif
  *out = g;
  return 0;
}

int hexed_b200_group_destroy(hexed_b200_group* g)
{
  if (!g) return 0;
  for (int r = 0; r < g->n; ++r) {
    cudaSetDevice(g->ctx[r]->device);
    cudaStreamSynchronize(g->ctx[r]->stream);
    if (g->comm_stream[r]) cudaStreamSynchronize(g->comm_stream[r]);
  }
#ifndef HB_EMULATE
  for (ncclComm_t c : g->comm) if (c) g_nccl.CommDestroy(c);
#endif
  for (int r = 0; r < g->n; ++r) {
    free_links(g, r);
    if (g->d_dt[r]) cudaFree(g->d_dt[r]);
    if (g->ev_ready[r]) cudaEventDestroy(g->ev_ready[r]);
    if (g->ev_done[r]) cudaEventDestroy(g->ev_done[r]);
    if (g->comm_stream[r]) cudaStreamDestroy(g->comm_stream[r]);
  }
  delete g;
  return 0;
}

int hexed_b200_group_size(const hexed_b200_group* g) { return g->n; }
hexed_b200_ctx* hexed_b200_group_ctx(hexed_b200_group* g, int rank) { return rank >= 0 && rank < g->n ? g->ctx[rank] : nullptr; }

int hexed_b200_group_info(const hexed_b200_group* g, int* nccl_version, long long* exchanges, long long* bytes_sent)
{
  if (nccl_version) *nccl_version = g->nccl_version;
  if (exchanges) *exchanges = g->exchanges;
  if (bytes_sent) *bytes_sent = g->bytes_sent;
  return 0;
}

int hexed_b200_group_set_halo(hexed_b200_group* g, int rank, int n_peers, const int* peers, const int* n_send, const int* send_slots,
                              const int* n_recv, const int* recv_slots)
{
  if (rank < 0 || rank >= g->n) return gfail(g, HEXED_B200_BAD_ARGUMENT, "rank out of range");
  hexed_b200_ctx* c = g->ctx[rank];
  if (!c->have_mesh) return gfail(g, HEXED_B200_NO_MESH, "upload the rank's mesh before its halo");
  free_links(g, rank);
  const size_t widest = (size_t)(c->nd + c->rs)*c->nfq;
  HG_CUDA(g, cudaSetDevice(c->device));
  for (int i = 0; i < n_peers; ++i) {
    if (peers[i] < 0 || peers[i] >= g->n || peers[i] == rank) return gfail(g, HEXED_B200_BAD_ARGUMENT, "bad peer");
    Link l; l.peer = peers[i]; l.n_send = n_send[i]; l.n_recv = n_recv[i];
    if (l.n_send) {
      HG_CTX(g, rank, hexed_b200_face_list_create(c, send_slots, l.n_send, &l.send_list));
      HG_CUDA(g, cudaMalloc(&l.d_send, sizeof(double)*l.n_send*widest));
    }
    if (l.n_recv) {
      HG_CTX(g, rank, hexed_b200_face_list_create(c, recv_slots, l.n_recv, &l.recv_list));
      HG_CUDA(g, cudaMalloc(&l.d_recv, sizeof(double)*l.n_recv*widest));
    }
    send_slots += l.n_send; recv_slots += l.n_recv;
    g->links[rank].push_back(l);
  }
  return 0;
}

/* halo exchange of one face kind on its own (e.g. after compute_write_face at initialisation); ends stream-ordered, not synchronised */
int hexed_b200_group_exchange(hexed_b200_group* g, int kind)
{
  if (kind < 0 || kind > 2) return gfail(g, HEXED_B200_BAD_ARGUMENT, "unknown face kind");
  int rc = exchange_start(g, kind); if (rc) return rc;
  return exchange_finish(g, kind);
}

int hexed_b200_group_synchronize(hexed_b200_group* g)
{
  for (int r = 0; r < g->n; ++r) {
    HG_CUDA(g, cudaSetDevice(g->ctx[r]->device));
    HG_CUDA(g, cudaStreamSynchronize(g->comm_stream[r]));
    HG_CUDA(g, cudaStreamSynchronize(g->ctx[r]->stream));
  }
  return 0;
}

/* void compute_euler(Kernel_mesh, Kernel_options) on the partitioned mesh                 include/kernels.hpp:22, src/kernels_convective.cpp:18 */
int hexed_b200_group_compute_euler(hexed_b200_group* g, hexed_b200_options o)
{
  int rc = exchange_start(g, 0); if (rc) return rc;
  // (one dispatch for the rest of the stage: interior flux, wait for the transfer, unpack, everything else)
  return g->workers.run([&](int r) -> int {
    hexed_b200_ctx* c = g->ctx[r];
    HG_CTX(g, r, hexed_b200_compute_euler_begin(c));
    HG_CUDA(g, cudaSetDevice(c->device));
    HG_CUDA(g, cudaStreamWaitEvent(c->stream, g->ev_done[r], 0));
    for (const Link& l : g->links[r]) if (l.n_recv) HG_CTX(g, r, hexed_b200_face_list_scatter(c, l.recv_list, 0, l.d_recv));
    HG_CTX(g, r, hexed_b200_compute_euler_finish(c, o));
    return 0;
  });
}

/* void compute_navier_stokes(Kernel_mesh, Kernel_options, flux_bc, visc, therm_cond)     include/kernels.hpp:24-25, src/kernels_diffusive.cpp:28-29
 * two exchanges (state faces, then the LDG viscous-flux faces); `flux_bc` is called ONCE, when every rank has its Prolong enqueued */
int hexed_b200_group_compute_navier_stokes(hexed_b200_group* g, hexed_b200_options o, hexed_b200_callback flux_bc, void* user,
                                           hexed_b200_transport visc, hexed_b200_transport therm_cond)
{
  int rc = exchange_start(g, 0); if (rc) return rc;
  if ((rc = g->workers.run([&](int r) -> int { HG_CTX(g, r, hexed_b200_compute_navier_stokes_begin(g->ctx[r], o, visc, therm_cond)); return 0; }))) return rc;
  if ((rc = exchange_finish(g, 0))) return rc;
  if ((rc = g->workers.run([&](int r) -> int { HG_CTX(g, r, hexed_b200_compute_navier_stokes_middle_local(g->ctx[r], o, visc, therm_cond)); return 0; }))) return rc;
  if (!o.i_stage) {
    if (flux_bc) flux_bc(user);
    if ((rc = exchange_start(g, 1))) return rc;
    if ((rc = g->workers.run([&](int r) -> int { HG_CTX(g, r, hexed_b200_compute_navier_stokes_middle_reconcile(g->ctx[r], o, visc, therm_cond)); return 0; }))) return rc;
    if ((rc = exchange_finish(g, 1))) return rc;
  }
  if ((rc = g->workers.run([&](int r) -> int { HG_CTX(g, r, hexed_b200_compute_navier_stokes_finish(g->ctx[r], o, visc, therm_cond)); return 0; }))) return rc;
  return 0;
}

/* double max_dt_*(Kernel_mesh, ...) on the partitioned mesh (src/kernels_max_dt.cpp:14-21): every rank leaves its minimum on its
 * device, ncclAllReduce(min) makes it global, ONE 8-byte read-back. pde: 0 Euler, 1 Navier-Stokes, 2 advection, 3 smooth AV, 4 FTA */
int hexed_b200_group_max_dt(hexed_b200_group* g, int pde, double convective_safety, double diffusive_safety, int local_time,
                            hexed_b200_transport visc, hexed_b200_transport therm_cond, double advect_length, double* dt)
{
  int rc = g->workers.run([&](int r) -> int {
    HG_CTX(g, r, hexed_b200_max_dt_device(g->ctx[r], pde, convective_safety, diffusive_safety, local_time, visc, therm_cond, advect_length, g->d_dt[r]));
    return 0;
  });
  if (rc) return rc;
  if (local_time) { *dt = 1.; return 0; }
#ifndef HB_EMULATE
  if (g->n > 1) {
    for (int r = 0; r < g->n; ++r) {
      HG_CUDA(g, cudaSetDevice(g->ctx[r]->device));
      HG_CUDA(g, cudaEventRecord(g->ev_ready[r], g->ctx[r]->stream));
      HG_CUDA(g, cudaStreamWaitEvent(g->comm_stream[r], g->ev_ready[r], 0));
    }
    HG_NCCL(g, g_nccl.GroupStart());
    for (int r = 0; r < g->n; ++r) HG_NCCL(g, g_nccl.AllReduce(g->d_dt[r], g->d_dt[r], 1, ncclDouble, ncclMin, g->comm[r], g->comm_stream[r]));
    HG_NCCL(g, g_nccl.GroupEnd());
  }
  double result = 0.;
  HG_CUDA(g, cudaSetDevice(g->ctx[0]->device));
  HG_CUDA(g, cudaMemcpyAsync(&result, g->d_dt[0], sizeof(double), cudaMemcpyDeviceToHost, g->n > 1 ? g->comm_stream[0] : g->ctx[0]->stream));
  HG_CUDA(g, cudaStreamSynchronize(g->n > 1 ? g->comm_stream[0] : g->ctx[0]->stream));
  *dt = result;
#else
  double result = *g->d_dt[0];
  for (int r = 1; r < g->n; ++r) result = std::min(result, *g->d_dt[r]);
  *dt = result;
#endif
  return 0;
}

} // extern "C"
