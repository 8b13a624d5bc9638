// dzb_comm.inl — the communicator of a bar split across GPUs and the collective block solve (part of dzb_api.cu's
// extern "C" section).  One process per GPU.  NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host such as
// PyTorch has already loaded, else the system one), so the library has no link-time dependency on it and single-GPU users
// never load it.  Reference semantics: ONE solve_by_thomas over the whole system (matrix_solver/mod.rs:17-48); here every
// rank eliminates its own row block and only the 2 interface rows per block (8 doubles) cross GPUs.

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) return a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.handle, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.handle, "ncclCommInitRank"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.handle, "ncclAllGather"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.handle, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.handle, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllGather && a.CommDestroy && a.GetErrorString;
    return a;
  }();
  return api;
}
#define NCCLCK(call)                                                                                     \
  do {                                                                                                   \
    ncclResult_t r_ = (call);                                                                            \
    if (r_ != ncclSuccess) return fail(DZB_CUDA, std::string(#call) + ": " + nccl_api().GetErrorString(r_)); \
  } while (0)

}  // namespace

struct dzb_comm {
  int rank = 0, world = 1, device = 0;
  ncclComm_t nccl = nullptr;
  // peer mailboxes (comm_kernels.cuh)
  double* mbox = nullptr;                  // this rank's
  double* peer[dzb::MB_MAX] = {};          // every rank's as mapped here; peer[rank] == mbox
  bool peer_ok = false;                    // every rank mapped every other rank's mailbox
  unsigned long long epoch = 0;            // exchanges made so far (the same on every rank)
  int* d_err = nullptr;                    // set by a mailbox wait that timed out
  unsigned long long timeout_ns = 30000000000ULL;  // DZB_COMM_TIMEOUT_S: how far apart the ranks may make the same call
  // scratch of the materialised path
  double *d_top8 = nullptr, *d_rows = nullptr, *d_ends = nullptr, *d_work = nullptr, *d_ends2 = nullptr, *d_edge8 = nullptr,
         *d_edges = nullptr;
};

namespace {

dzb::PeerBox comm_peerbox(dzb_comm* c) {  // the next exchange
  dzb::PeerBox px{};
  for (int r = 0; r < c->world; ++r) px.box[r] = c->peer[r];
  px.rank = c->rank, px.world = c->world, px.epoch = ++c->epoch, px.err = c->d_err, px.timeout_ns = c->timeout_ns;
  return px;
}
__global__ void k_block_edges(const double* __restrict__ u, long long n, double* __restrict__ edge8) {
  if (threadIdx.x < 8) edge8[threadIdx.x] = threadIdx.x == 0 ? u[0] : (threadIdx.x == 1 ? u[n - 1] : 0.0);
}
// all ranks' 8 doubles -> all[world][8]; with `solve`, also this block's (u_s, u_e) of the interface system -> c->d_ends2
int comm_exchange8(dzb_comm* c, const double* d_src8, double* d_all, bool solve) {
  if (c->peer_ok) {
    const dzb::PeerBox px = comm_peerbox(c);
    if (solve) dzb::k_peer_exchange<true><<<1, 32, 0, g_stream>>>(px, d_src8, d_all, c->d_ends2);
    else dzb::k_peer_exchange<false><<<1, 32, 0, g_stream>>>(px, d_src8, d_all, nullptr);
    LAUNCHED();
    return DZB_OK;
  }
  if (c->world == 1) CK(cudaMemcpyAsync(d_all, d_src8, 8 * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
  else NCCLCK(nccl_api().AllGather(d_src8, d_all, 8, ncclDouble, c->nccl, g_stream));
  if (solve) {
    dzb::k_interface_solve<<<1, 32, 0, g_stream>>>(d_all, 2 * c->world, c->d_ends, c->d_work);
    LAUNCHED();
    CK(cudaMemcpyAsync(c->d_ends2, c->d_ends + 2 * c->rank, 2 * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
  }
  return DZB_OK;
}

}  // namespace

extern "C" {

/* ---- communicator ------------------------------------------------------------------------------------------------ */
int dzb_comm_unique_id(unsigned char id[DZB_COMM_ID_BYTES]) {
  if (!id) return fail(DZB_INVALID, "null pointer");
  static_assert(sizeof(ncclUniqueId) == DZB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  if (!nccl_api().ok) return fail(DZB_CUDA, "libnccl.so.2 could not be loaded");
  ncclUniqueId u;
  NCCLCK(nccl_api().GetUniqueId(&u));
  std::memcpy(id, &u, sizeof u);
  return DZB_OK;
}
void dzb_comm_destroy(dzb_comm* c) {
  if (!c) return;
  cudaStreamSynchronize(g_stream);
  for (int r = 0; r < c->world && r < dzb::MB_MAX; ++r)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  if (c->nccl) nccl_api().CommDestroy(c->nccl);
  for (double** p : {&c->mbox, &c->d_top8, &c->d_rows, &c->d_ends, &c->d_work, &c->d_ends2, &c->d_edge8, &c->d_edges}) dfree(*p);
  if (c->d_err) cudaFree(c->d_err);
  delete c;
}
int dzb_comm_init(const unsigned char id[DZB_COMM_ID_BYTES], int rank, int world, dzb_comm** out) {
  if (!out) return fail(DZB_INVALID, "null pointer");
  *out = nullptr;
  if (world < 1 || world > dzb::MB_MAX || rank < 0 || rank >= world) return fail(DZB_INVALID, "rank / world (at most 8 ranks: one box)");
  if (world > 1 && !id) return fail(DZB_INVALID, "null unique id");
  dzb_comm* c = new dzb_comm;
  c->rank = rank, c->world = world;
  if (const char* t = std::getenv("DZB_COMM_TIMEOUT_S")) c->timeout_ns = (unsigned long long)(std::max(0.001, std::atof(t)) * 1e9);
  auto bail = [&](int code) {
    const std::string msg = g_err;
    dzb_comm_destroy(c);
    g_err = msg;
    return code;
  };
  if (cudaGetDevice(&c->device) != cudaSuccess) return bail(fail(DZB_CUDA, "no CUDA device"));
  int rc = DZB_OK;
  const size_t w = (size_t)world;
  if ((rc = dalloc(&c->mbox, dzb::MB_DOUBLES)) || (rc = dalloc(&c->d_top8, 8)) || (rc = dalloc(&c->d_rows, 8 * w)) ||
      (rc = dalloc(&c->d_ends, 2 * w)) || (rc = dalloc(&c->d_work, 4 * w)) || (rc = dalloc(&c->d_ends2, 2)) ||
      (rc = dalloc(&c->d_edge8, 8)) || (rc = dalloc(&c->d_edges, 8 * w)))
    return bail(rc);
  if (cudaMalloc((void**)&c->d_err, sizeof(int)) != cudaSuccess) return bail(fail(DZB_CUDA, "cudaMalloc"));
  if (cudaMemsetAsync(c->mbox, 0, dzb::MB_DOUBLES * sizeof(double), g_stream) != cudaSuccess ||
      cudaMemsetAsync(c->d_err, 0, sizeof(int), g_stream) != cudaSuccess)
    return bail(fail(DZB_CUDA, "cudaMemsetAsync"));
  c->peer[rank] = c->mbox;
  if (world == 1) {
    c->peer_ok = !env_set("DZB_COMM_NO_PEER");
    if (cudaStreamSynchronize(g_stream) != cudaSuccess) return bail(fail(DZB_CUDA, "cudaStreamSynchronize"));
    *out = c;
    return DZB_OK;
  }
  if (!nccl_api().ok) return bail(fail(DZB_CUDA, "libnccl.so.2 could not be loaded"));
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof u);
  {
    ncclResult_t r = nccl_api().CommInitRank(&c->nccl, world, u, rank);
    if (r != ncclSuccess) return bail(fail(DZB_CUDA, std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(r)));
  }
  // Map every rank's mailbox: the 64-byte IPC handles travel through one ncclAllGather (the mailboxes were zeroed on this
  // stream before it, so nobody can post into a mailbox that is not ready), then a second one agrees on the outcome.
  struct Rec {
    cudaIpcMemHandle_t h;
    long long ok;
  };
  static_assert(sizeof(Rec) == 72, "IPC handle record");
  Rec mine{}, *d_mine = nullptr, *d_all = nullptr;
  std::vector<Rec> all(w);
  mine.ok = !env_set("DZB_COMM_NO_PEER") && cudaIpcGetMemHandle(&mine.h, c->mbox) == cudaSuccess;
  cudaGetLastError();
  if (cudaMalloc((void**)&d_mine, sizeof(Rec)) != cudaSuccess || cudaMalloc((void**)&d_all, w * sizeof(Rec)) != cudaSuccess)
    return bail(fail(DZB_CUDA, "cudaMalloc"));
  auto gather = [&]() -> int {
    CK(cudaMemcpyAsync(d_mine, &mine, sizeof(Rec), cudaMemcpyHostToDevice, g_stream));
    NCCLCK(nccl_api().AllGather(d_mine, d_all, sizeof(Rec), ncclChar, c->nccl, g_stream));
    CK(cudaMemcpyAsync(all.data(), d_all, w * sizeof(Rec), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return DZB_OK;
  };
  rc = gather();
  if (!rc) {
    bool ok = true;
    for (int r = 0; r < world; ++r) ok = ok && all[(size_t)r].ok;
    for (int r = 0; ok && r < world; ++r) {
      if (r == rank) continue;
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        break;
      }
      c->peer[r] = static_cast<double*>(p);
    }
    mine.ok = ok;
    rc = gather();  // everybody uses the mailboxes, or nobody does
    if (!rc) {
      c->peer_ok = true;
      for (int r = 0; r < world; ++r) c->peer_ok = c->peer_ok && all[(size_t)r].ok;
    }
  }
  cudaFree(d_mine), cudaFree(d_all);
  if (rc) return bail(rc);
  *out = c;
  return DZB_OK;
}
int dzb_comm_info(const dzb_comm* c, int* rank, int* world, int* peer_mailboxes) {
  if (!c) return fail(DZB_INVALID, "null pointer");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  if (peer_mailboxes) *peer_mailboxes = c->peer_ok ? 1 : 0;
  return DZB_OK;
}
/* ncclAllGather of `count` doubles per rank on the library's stream (device pointers). */
int dzb_comm_all_gather(dzb_comm* c, const double* d_send, double* d_recv, size_t count) {
  if (!c || !d_send || !d_recv) return fail(DZB_INVALID, "null pointer");
  if (c->world == 1) {
    CK(cudaMemcpyAsync(d_recv, d_send, count * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    return DZB_OK;
  }
  NCCLCK(nccl_api().AllGather(d_send, d_recv, count, ncclDouble, c->nccl, g_stream));
  return DZB_OK;
}

/* ---- the collective solve of a split bar -------------------------------------------------------------------------- */
int dzb_solver_block_solve_device(dzb_solver* s, dzb_comm* c, const double** device_ptr) {
  if (!s || !c) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  const int want = (c->rank > 0 ? dzb::BS_OPEN_LEFT : 0) | (c->rank + 1 < c->world ? dzb::BS_OPEN_RIGHT : 0);
  if (s->block_flags != want) return fail(DZB_INVALID, "the block's open ends do not match its rank in the communicator");
  int rc = DZB_OK;
  if (s->fused && c->peer_ok) {  // assembly, block elimination, exchange, back-substitution: ONE launch per rank
    bool done = false;
    const dzb::PeerBox px = comm_peerbox(c);
    rc = static_solve_onelaunch(s, &done, &px);
    if (rc) return rc;
    if (done) {
      s->last_path = 1;
      if (device_ptr) *device_ptr = s->out;
      return DZB_OK;
    }
    --c->epoch;  // nothing was launched (no cooperative launch on this device): every rank takes the same decision
  }
  if ((rc = block_fresh_bands(s))) return rc;
  s->last_path = 0;
  auto one_solve = [&](int which_rhs, bool accumulate) -> int {
    const double* f = which_rhs ? s->big.resid : s->rhs;
    if (dzb::bigsys_block_reduce(&s->big, s->lo, s->di, s->up, f, s->block_flags, c->d_top8, g_stream, &g_launches))
      return fail(DZB_CUDA, dzb::bigsys_error());
    int r = comm_exchange8(c, c->d_top8, c->d_rows, true);
    if (r) return r;
    if (dzb::bigsys_block_backsolve(&s->big, s->lo, s->di, s->up, f, s->block_flags, 0.0, 0.0, c->d_ends2, s->out, accumulate, g_stream,
                                    &g_launches))
      return fail(DZB_CUDA, dzb::bigsys_error());
    return DZB_OK;
  };
  if ((rc = one_solve(0, false))) return rc;
  for (int k = 0; k < s->refine_steps; ++k) {  // r = b - K u in double-double needs the neighbours' edge values
    k_block_edges<<<1, 32, 0, g_stream>>>(s->out, (long long)s->n, c->d_edge8);
    LAUNCHED();
    if ((rc = comm_exchange8(c, c->d_edge8, c->d_edges, false))) return rc;
    const double* left = c->rank > 0 ? c->d_edges + 8 * (c->rank - 1) + 1 : nullptr;
    const double* right = c->rank + 1 < c->world ? c->d_edges + 8 * (c->rank + 1) : nullptr;
    dzb::k_residual_dd<<<dzb::bigsys_grid(s->n), 256, 0, g_stream>>>(s->lo, s->di, s->up, s->rhs, s->out, s->big.resid, (long long)s->n,
                                                                      s->block_flags, 0.0, 0.0, left, right);
    LAUNCHED();
    if ((rc = one_solve(1, true))) return rc;
  }
  if (device_ptr) *device_ptr = s->out;
  return DZB_OK;
}
/* Synchronises and reports what the collective solves since the last call left behind: DZB_CUDA when a rank's record did
 * not arrive (some rank never made the call), DZB_INVALID when the block's new coordinates gave a matrix the partitioned
 * elimination is not reliable on or a row the closed form does not cover. */
int dzb_solver_block_status(dzb_solver* s, dzb_comm* c) {
  if (!s) return fail(DZB_INVALID, "null pointer");
  if (!s->is_block) return fail(DZB_INVALID, "not a block solver");
  int err = 0, cls[2] = {0, 0};
  if (c) CK(cudaMemcpyAsync(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  if (s->cls_pending) CK(cudaMemcpyAsync(cls, s->d_cls, sizeof cls, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  s->cls_pending = false;
  if (err) return fail(DZB_CUDA, "a rank's interface rows did not arrive within the time-out (did every rank call the collective solve?)");
  if (cls[1]) return fail(DZB_INVALID, "a row of the block is outside the closed-form assembly (degenerate spacing): use DZB_ASSEMBLY_FAITHFUL");
  if (cls[0] == 3) return fail(DZB_INVALID, "the block's matrix is neither an M-matrix nor strictly dominant");
  return DZB_OK;
}

}  // extern "C"
