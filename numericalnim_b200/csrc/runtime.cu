// runtime.cu — context life cycle, lazy NCCL binding, peer mailboxes, the vector pool, knobs and the profiler.
#include "internal.hpp"

static thread_local std::string g_thread_err;
std::string& thread_error() { return g_thread_err; }

int fail(const b200rk_ctx* ctx, int code, const std::string& msg) {
  g_thread_err = msg;
  if (ctx) ctx->err = msg;
  return code;
}

NcclApi g_nccl;

int nccl_bind(const b200rk_ctx* ctx) {
  if (g_nccl.handle) return B200RK_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  std::string where = "already loaded in the process";
  if (!h) {
    if (const char* p = getenv("B200RK_NCCL_LIB")) { h = dlopen(p, RTLD_NOW | RTLD_GLOBAL); where = p; }
  }
  if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); where = "libnccl.so.2 (default search path)"; }
  if (!h) return fail(ctx, B200RK_ENCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
  NcclApi a;
  a.handle = h; a.where = where;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
  a.Send = (decltype(a.Send))dlsym(h, "ncclSend");
  a.Recv = (decltype(a.Recv))dlsym(h, "ncclRecv");
  a.GroupStart = (decltype(a.GroupStart))dlsym(h, "ncclGroupStart");
  a.GroupEnd = (decltype(a.GroupEnd))dlsym(h, "ncclGroupEnd");
  a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString || !a.Send || !a.Recv ||
      !a.GroupStart || !a.GroupEnd || !a.AllGather)
    return fail(ctx, B200RK_ENCCL, "libnccl.so.2 lacks a required symbol");
  g_nccl = a;
  return B200RK_OK;
}

int ensure_partials(b200rk_ctx* c, size_t blocks) {
  if (blocks <= c->partials_cap) return B200RK_OK;
  if (c->d_partials) CUDA_TRY(c, cudaFree(c->d_partials));
  size_t cap = std::max(blocks, (size_t)1 << 16);
  CUDA_TRY(c, cudaMalloc(&c->d_partials, cap * sizeof(double)));
  c->partials_cap = cap;
  return B200RK_OK;
}

ReduceScratch reduce_scratch(b200rk_ctx* c) {
  ReduceScratch rs;
  rs.partials = c->d_partials;
  rs.ticket = c->d_ticket;
  rs.result = c->d_result;
  const bool in_kernel_collective = c->world > 1 && c->p2p;
  rs.result_host = (c->world == 1 || in_kernel_collective) ? c->h_result_dev : nullptr;
  rs.seq_host = c->h_seq_dev;
  rs.seq = ++c->seq;
  rs.mail.world = in_kernel_collective ? c->world : 1;
  rs.mail.rank = c->rank;
  rs.mail.timeout_cycles = c->peer_timeout_cycles;
  for (int p = 0; p < kMaxPeers; ++p) rs.mail.box[p] = (in_kernel_collective && p < c->world) ? c->peer_mail[p] : nullptr;
  if (in_kernel_collective) c->collectives++;
  return rs;
}

// Peer mailboxes for the in-kernel all-reduce of the error norm (kernels.cuh: peer_allreduce). Every rank
// cudaMalloc's a 512-byte mailbox, the CUDA-IPC handles travel through one ncclAllGather on the
// communicator we already have, and each rank maps its peers' mailboxes (NVLink/NVSwitch peer access). The
// outcome is agreed by an ncclAllReduce(min): either every rank uses the mailboxes or every rank falls back to
// ncclAllReduce per attempt.
int setup_p2p(b200rk_ctx* c) {
  c->p2p = false;
  if (const char* e = getenv("B200RK_P2P")) if (atoi(e) == 0) { c->p2p_note = "disabled by B200RK_P2P=0"; return B200RK_OK; }
  if (c->world > kMaxPeers) { c->p2p_note = "world larger than kMaxPeers"; return B200RK_OK; }
  const size_t mail_bytes = 2 * kMaxPeers * 2 * sizeof(unsigned long long);
  CUDA_TRY(c, cudaMalloc(&c->d_mail, mail_bytes));
  CUDA_TRY(c, cudaMemset(c->d_mail, 0, mail_bytes));
  cudaIpcMemHandle_t mine;
  int ok = 1;
  if (cudaIpcGetMemHandle(&mine, c->d_mail) != cudaSuccess) { ok = 0; cudaGetLastError(); std::memset(&mine, 0, sizeof(mine)); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  char* d_handles = nullptr;
  int* d_flag = nullptr;
  CUDA_TRY(c, cudaMalloc(&d_handles, 64 * (size_t)c->world));
  CUDA_TRY(c, cudaMalloc(&d_flag, sizeof(int)));
  CUDA_TRY(c, cudaMemcpyAsync(d_handles + 64 * (size_t)c->rank, &mine, 64, cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(c, g_nccl.AllGather(d_handles + 64 * (size_t)c->rank, d_handles, 64, ncclChar, c->comm, c->stream));
  std::vector<cudaIpcMemHandle_t> all(c->world);
  CUDA_TRY(c, cudaMemcpyAsync(all.data(), d_handles, 64 * (size_t)c->world, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int p = 0; p < c->world && ok; ++p) {
    if (p == c->rank) { c->peer_mail[p] = c->d_mail; continue; }
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      c->p2p_note = std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError());
      ok = 0;
    } else {
      c->peer_mail[p] = static_cast<unsigned long long*>(ptr);
      c->peer_opened[p] = true;
    }
  }
  CUDA_TRY(c, cudaMemcpyAsync(d_flag, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(c, g_nccl.AllReduce(d_flag, d_flag, 1, ncclInt, ncclMin, c->comm, c->stream));
  int agreed = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&agreed, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_handles);
  cudaFree(d_flag);
  c->p2p = agreed != 0;
  if (c->p2p) c->p2p_note = "peer mailboxes mapped over CUDA IPC";
  else if (c->p2p_note.empty()) c->p2p_note = "a peer could not map the mailboxes";
  return B200RK_OK;
}

// This GPU's own mailbox (setup_p2p allocates it for sharded contexts; a single-GPU context gets it on first use by the
// device-resident loop, which sends its grid-wide sum through it).
int ensure_local_mailbox(b200rk_ctx* c) {
  if (c->d_mail) return B200RK_OK;
  const size_t mail_bytes = 2 * kMaxPeers * 2 * sizeof(unsigned long long);
  CUDA_TRY(c, cudaMalloc(&c->d_mail, mail_bytes));
  CUDA_TRY(c, cudaMemset(c->d_mail, 0, mail_bytes));
  return B200RK_OK;
}

int stream_barrier(b200rk_ctx* c) {
  if (c->world < 2) return B200RK_OK;
  if (!c->d_barrier) {
    CUDA_TRY(c, cudaMalloc(&c->d_barrier, sizeof(int)));
    CUDA_TRY(c, cudaMemset(c->d_barrier, 0, sizeof(int)));
  }
  NCCL_TRY(c, g_nccl.AllReduce(c->d_barrier, c->d_barrier, 1, ncclInt, ncclSum, c->comm, c->stream));   // zeros: only the ordering matters
  c->collectives++;
  return B200RK_OK;
}

// Same scheme as setup_p2p: handles travel through one ncclAllGather, each rank maps its two ring neighbours' vectors,
// the outcome is agreed by an ncclAllReduce(min) — either every rank reads its halo in place or none does.
int peer_view_open(b200rk_ctx* c, b200rk_vec* const* vecs, int count, PeerVecView* out) {
  *out = PeerVecView{};
  if (c->world < 2 || !c->p2p || count < 1 || count > PeerVecView::kMax) return B200RK_OK;
  const size_t per_rank = 64 * (size_t)PeerVecView::kMax;
  std::vector<cudaIpcMemHandle_t> mine(PeerVecView::kMax);
  std::memset(mine.data(), 0, per_rank);
  int ok = 1;
  for (int i = 0; i < count; ++i)
    if (cudaIpcGetMemHandle(&mine[i], vecs[i]->d) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  char* d_handles = nullptr;
  int* d_flag = nullptr;
  CUDA_TRY(c, cudaMalloc(&d_handles, per_rank * (size_t)c->world));
  CUDA_TRY(c, cudaMalloc(&d_flag, sizeof(int)));
  CUDA_TRY(c, cudaMemcpyAsync(d_handles + per_rank * (size_t)c->rank, mine.data(), per_rank, cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(c, g_nccl.AllGather(d_handles + per_rank * (size_t)c->rank, d_handles, per_rank, ncclChar, c->comm, c->stream));
  std::vector<cudaIpcMemHandle_t> all((size_t)PeerVecView::kMax * c->world);
  CUDA_TRY(c, cudaMemcpyAsync(all.data(), d_handles, per_rank * (size_t)c->world, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  const int left = (c->rank + c->world - 1) % c->world, right = (c->rank + 1) % c->world;
  PeerVecView v;
  for (int i = 0; i < count && ok; ++i) {
    void* pl = nullptr;
    if (cudaIpcOpenMemHandle(&pl, all[(size_t)left * PeerVecView::kMax + i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    v.opened[v.n_opened++] = pl;
    v.left[i] = static_cast<const double*>(pl);
    if (right == left) { v.right[i] = v.left[i]; continue; }   // two ranks: both neighbours are the same peer (one mapping per allocation)
    void* pr = nullptr;
    if (cudaIpcOpenMemHandle(&pr, all[(size_t)right * PeerVecView::kMax + i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    v.opened[v.n_opened++] = pr;
    v.right[i] = static_cast<const double*>(pr);
  }
  CUDA_TRY(c, cudaMemcpyAsync(d_flag, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(c, g_nccl.AllReduce(d_flag, d_flag, 1, ncclInt, ncclMin, c->comm, c->stream));
  int agreed = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&agreed, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_handles);
  cudaFree(d_flag);
  c->collectives += 2;
  if (!agreed) {
    for (int i = 0; i < v.n_opened; ++i) cudaIpcCloseMemHandle(v.opened[i]);
    cudaGetLastError();
    return B200RK_OK;
  }
  size_t off = 0;
  shard_range(vecs[0]->n_global, left, c->world, &off, &v.n_left);
  shard_range(vecs[0]->n_global, right, c->world, &off, &v.n_right);
  for (int i = 0; i < count; ++i) v.local[i] = vecs[i]->d;
  v.count = count;
  *out = v;
  return B200RK_OK;
}

int peer_view_close(b200rk_ctx* c, PeerVecView* v) {
  if (!v->count) return B200RK_OK;
  // the neighbours may still be reading this rank's vectors in their last kernel, and exported memory must outlive every
  // mapping of it: all streams drain, all ranks unmap, and only then does anyone go on to release its vectors
  TRY(stream_barrier(c));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < v->n_opened; ++i) cudaIpcCloseMemHandle(v->opened[i]);
  cudaGetLastError();
  *v = PeerVecView{};
  TRY(stream_barrier(c));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200RK_OK;
}

// =====================================================================================================
// vectors
// =====================================================================================================
void shard_range(size_t n, int rank, int world, size_t* off, size_t* len) {
  size_t chunk = (n + world - 1) / world;
  chunk = (chunk + 3) / 4 * 4;  // keep shard boundaries 32-byte aligned in the global index space
  size_t lo = std::min(n, (size_t)rank * chunk), hi = std::min(n, (size_t)(rank + 1) * chunk);
  *off = lo;
  *len = hi - lo;
}

int vec_alloc(b200rk_ctx* c, size_t n_global, b200rk_vec** out) {
  for (size_t i = 0; i < c->pool.size(); ++i) {
    if (c->pool[i]->n_global == n_global) {
      *out = c->pool[i];
      c->pool.erase(c->pool.begin() + i);
      return B200RK_OK;
    }
  }
  b200rk_vec* v = new b200rk_vec{c, n_global, 0, 0, nullptr};
  shard_range(n_global, c->rank, c->world, &v->offset, &v->n_local);
  cudaError_t e = cudaMalloc(&v->d, std::max<size_t>(v->n_local, 4) * sizeof(double));
  if (e == cudaErrorMemoryAllocation && !c->pool.empty()) {
    // cached vectors of other lengths are holding the memory: give them back to the driver and try once more
    cudaGetLastError();
    pool_trim(c, 0);
    e = cudaMalloc(&v->d, std::max<size_t>(v->n_local, 4) * sizeof(double));
  }
  if (e != cudaSuccess) {
    delete v;
    return fail(c, e == cudaErrorMemoryAllocation ? B200RK_ENOMEM : B200RK_ECUDA,
                std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  *out = v;
  return B200RK_OK;
}
// Free pooled vectors until at most `keep_bytes` stay cached (0: empty the pool).
void pool_trim(b200rk_ctx* c, size_t keep_bytes) {
  size_t held = 0;
  for (auto* p : c->pool) held += p->n_local * sizeof(double);
  if (held <= keep_bytes) return;
  cudaStreamSynchronize(c->stream);  // enqueued work may still use them
  while (!c->pool.empty() && held > keep_bytes) {
    b200rk_vec* p = c->pool.back();
    c->pool.pop_back();
    held -= p->n_local * sizeof(double);
    cudaFree(p->d);
    delete p;
  }
}
// Back to the pool (stream-ordered reuse is safe: one stream per context), bounded by count AND by the byte budget — the
// same rule for the library's internal releases (solver work vectors, solve_host / quadrature outputs) and the public
// b200rk_vec_free: a call that held 10^5 vectors, or 512 x 64 MiB, must not leave them all cached.
bool pool_put(b200rk_vec* v) {
  b200rk_ctx* c = v->ctx;
  size_t held = 0;
  for (auto* p : c->pool) held += p->n_local * sizeof(double);
  if (c->pool.size() < 512 && held + v->n_local * sizeof(double) <= c->pool_budget_bytes) { c->pool.push_back(v); return true; }
  return false;
}
void vec_release(b200rk_vec* v) {
  if (!v) return;
  if (pool_put(v)) return;
  cudaStreamSynchronize(v->ctx->stream);  // enqueued work may still use it
  cudaFree(v->d);
  delete v;
}
int check_same(const b200rk_ctx* c, const b200rk_vec* a, const b200rk_vec* b) {
  if (!a || !b) return fail(c, B200RK_EINVAL, "null vector");
  if (a->ctx != b->ctx) return fail(c, B200RK_EINVAL, "vectors belong to different contexts");
  if (a->n_global != b->n_global) return fail(c, B200RK_EINVAL, "Vectors must have the same size.");  // utils.nim:26
  return B200RK_OK;
}

int vec_copy_raw(b200rk_ctx* c, b200rk_vec* dst, const b200rk_vec* src) {
  CUDA_TRY(c, cudaMemcpyAsync(dst->d, src->d, src->n_local * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return B200RK_OK;
}

extern "C" {

const char* b200rk_last_error(const b200rk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_thread_err.c_str(); }

static int ctx_common_init(b200rk_ctx* c) {
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CUDA_TRY(c, cudaGetDeviceProperties(&prop, c->device));
  c->sm_count = prop.multiProcessorCount;
  c->clock_khz = prop.clockRate > 0 ? prop.clockRate : 1965000;
  if (const char* e = getenv("B200RK_PEER_TIMEOUT_S")) c->peer_timeout_s = std::max(1, atoi(e));
  c->peer_timeout_cycles = (long long)c->peer_timeout_s * (long long)c->clock_khz * 1000ll;
  if (prop.l2CacheSize > 0) c->l2_bytes = (size_t)prop.l2CacheSize;
  CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {  // stream-ordered scratch (quadrature.cu's per-call tables): keep freed blocks cached instead of returning them to the OS at every sync
    cudaMemPool_t pool = nullptr;
    unsigned long long keep = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    cudaGetLastError();
  }
  CUDA_TRY(c, cudaMalloc(&c->d_ticket, sizeof(unsigned int)));
  CUDA_TRY(c, cudaMemset(c->d_ticket, 0, sizeof(unsigned int)));
  CUDA_TRY(c, cudaMalloc(&c->d_result, sizeof(double)));
  CUDA_TRY(c, cudaMalloc(&c->d_halo, 4 * sizeof(double)));
  CUDA_TRY(c, cudaHostAlloc(&c->h_result, sizeof(double), cudaHostAllocMapped));
  CUDA_TRY(c, cudaHostGetDevicePointer(&c->h_result_dev, c->h_result, 0));
  CUDA_TRY(c, cudaHostAlloc(&c->h_seq, sizeof(unsigned long long), cudaHostAllocMapped));
  *c->h_seq = 0;
  CUDA_TRY(c, cudaHostGetDevicePointer(&c->h_seq_dev, c->h_seq, 0));
  if (const char* e = getenv("B200RK_SPIN_READBACK")) c->spin_readback = atoi(e) != 0;
  if (const char* e = getenv("B200RK_DEVICE_LOOP")) c->device_loop = atoi(e) < 0 ? -1 : (atoi(e) != 0);
  TRY(ensure_partials(c, 1));
  if (const char* e = getenv("B200RK_VEC_WIDTH")) c->vec_width = (atoi(e) == 2) ? 2 : 4;
  if (const char* e = getenv("B200RK_CTAS_PER_SM")) c->ctas_per_sm = std::max(0, atoi(e));
  if (const char* e = getenv("B200RK_FINISH_CTAS_PER_SM")) c->finish_ctas_per_sm = std::max(0, atoi(e));
  if (const char* e = getenv("B200RK_STRICT_ZEROS")) c->strict_zeros = atoi(e) != 0;
  if (const char* e = getenv("B200RK_FUSE_POINTWISE")) c->fuse_pointwise = atoi(e) != 0;
  if (const char* e = getenv("B200RK_FUSE_STENCIL")) c->fuse_stencil = atoi(e) != 0;
  if (const char* e = getenv("B200RK_TSTART_COPY")) c->tstart_copy = atoi(e) != 0;
  if (const char* e = getenv("B200RK_L96_ATTEMPT_THREADS")) c->l96_attempt_threads = atoi(e) == 128 ? 128 : 256;
  if (const char* e = getenv("B200RK_L96_WARP_TILES")) c->l96_warp_tiles = (atoi(e) == 4) ? 4 : (atoi(e) != 0 ? 8 : 0);
  if (const char* e = getenv("B200RK_L96_CTAS_PER_SM")) c->l96_ctas_per_sm = std::max(0, std::min(32, atoi(e)));
  if (const char* e = getenv("B200RK_FUSE_STENCIL_ATTEMPT")) c->fuse_stencil_attempt = atoi(e) != 0;
  if (const char* e = getenv("B200RK_L96_PEER_HALO")) c->l96_peer_halo = atoi(e) != 0;
  if (const char* e = getenv("B200RK_L96_ATTEMPT_PAIRS")) c->l96_attempt_pairs = (atoi(e) == 1) ? 1 : 2;
  if (const char* e = getenv("B200RK_FUSE_SIMPSON")) c->fuse_simpson = atoi(e) != 0;
  if (const char* e = getenv("B200RK_FINISH_PREFETCH")) c->finish_prefetch = atoi(e) != 0;
  if (const char* e = getenv("B200RK_L2_HINTS")) c->l2_hints = atoi(e) < 0 ? -1 : (atoi(e) != 0);
  CUDA_TRY(c, cudaDeviceSynchronize());
  return B200RK_OK;
}

int b200rk_init(b200rk_ctx** out, int device) {
  if (!out) return fail(nullptr, B200RK_EINVAL, "null out");
  b200rk_ctx* c = new b200rk_ctx;
  c->device = device;
  int rc = ctx_common_init(c);
  if (rc != B200RK_OK) { thread_error() = c->err; b200rk_destroy(c); cudaGetLastError(); return rc; }  // releases whatever was created
  *out = c;
  return B200RK_OK;
}

int b200rk_nccl_unique_id(void* out128) {
  if (!out128) return fail(nullptr, B200RK_EINVAL, "null argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  TRY(nccl_bind(nullptr));
  ncclUniqueId id;
  NCCL_TRY(nullptr, g_nccl.GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return B200RK_OK;
}

int b200rk_init_distributed(b200rk_ctx** out, int device, int rank, int world, const void* id128) {
  if (!out || world < 1 || rank < 0 || rank >= world) return fail(nullptr, B200RK_EINVAL, "bad rank/world");
  if (world > 1 && !id128) return fail(nullptr, B200RK_EINVAL, "null NCCL id");
  b200rk_ctx* c = new b200rk_ctx;
  c->device = device; c->rank = rank; c->world = world;
  int rc = ctx_common_init(c);
  if (rc == B200RK_OK && world > 1) {
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    rc = nccl_bind(c);
    if (rc == B200RK_OK) {
      ncclResult_t e = g_nccl.CommInitRank(&c->comm, world, id, rank);
      if (e != ncclSuccess) rc = fail(c, B200RK_ENCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(e));
    }
    if (rc == B200RK_OK) rc = setup_p2p(c);
  }
  if (rc != B200RK_OK) { thread_error() = c->err; b200rk_destroy(c); cudaGetLastError(); return rc; }  // releases whatever was created
  *out = c;
  return B200RK_OK;
}

void b200rk_destroy(b200rk_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (auto* v : c->pool) { cudaFree(v->d); delete v; }
  for (auto& r : c->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : c->ev_free) cudaEventDestroy(e);
  for (int p = 0; p < kMaxPeers; ++p) if (c->peer_opened[p]) cudaIpcCloseMemHandle(c->peer_mail[p]);
  if (c->d_mail) cudaFree(c->d_mail);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  cudaFree(c->d_partials); cudaFree(c->d_ticket); cudaFree(c->d_result); cudaFree(c->d_halo); cudaFreeHost(c->h_result); cudaFreeHost(c->h_seq);
  if (c->d_halo_attempt) cudaFree(c->d_halo_attempt);
  if (c->d_halo_stencil) cudaFree(c->d_halo_stencil);
  if (c->d_barrier) cudaFree(c->d_barrier);
  if (c->d_run_state) cudaFree(c->d_run_state);
  if (c->h_run_state) cudaFreeHost(c->h_run_state);
  if (c->copy_event) cudaEventDestroy(c->copy_event);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

void* b200rk_stream(const b200rk_ctx* c) { return c ? (void*)c->stream : nullptr; }
int b200rk_synchronize(b200rk_ctx* c) {
  if (!c) return fail(nullptr, B200RK_EINVAL, "null context");
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200RK_OK;
}
int b200rk_rank(const b200rk_ctx* c) { return c ? c->rank : 0; }
int b200rk_world(const b200rk_ctx* c) { return c ? c->world : 0; }

int b200rk_set(b200rk_ctx* c, const char* key, int64_t v) {
  if (!c) return fail(nullptr, B200RK_EINVAL, "null context");
  std::string k = key ? key : "";
  if (k == "strict_zeros") c->strict_zeros = v != 0;
  else if (k == "vec_width") { if (v != 2 && v != 4) return fail(c, B200RK_EINVAL, "vec_width must be 2 or 4"); c->vec_width = (int)v; }
  else if (k == "ctas_per_sm") { if (v < 0) return fail(c, B200RK_EINVAL, "ctas_per_sm must be >= 0"); c->ctas_per_sm = (int)v; }
  else if (k == "finish_ctas_per_sm") { if (v < 0) return fail(c, B200RK_EINVAL, "finish_ctas_per_sm must be >= 0"); c->finish_ctas_per_sm = (int)v; }
  else if (k == "profile") c->profile = v != 0;
  else if (k == "fuse_pointwise") c->fuse_pointwise = v != 0;
  else if (k == "spin_readback") c->spin_readback = v != 0;
  else if (k == "device_loop") c->device_loop = v < 0 ? -1 : (v != 0);
  else if (k == "fuse_stencil") c->fuse_stencil = v != 0;
  else if (k == "fuse_stencil_attempt") c->fuse_stencil_attempt = v != 0;
  else if (k == "l96_peer_halo") c->l96_peer_halo = v != 0;
  else if (k == "tstart_copy") c->tstart_copy = v != 0;
  else if (k == "l96_warp_tiles") c->l96_warp_tiles = (v == 4) ? 4 : (v != 0 ? 8 : 0);
  else if (k == "l96_attempt_threads") { if (v != 128 && v != 256) return fail(c, B200RK_EINVAL, "l96_attempt_threads must be 128 or 256"); c->l96_attempt_threads = (int)v; }
  else if (k == "l96_ctas_per_sm") { if (v < 0 || v > 32) return fail(c, B200RK_EINVAL, "l96_ctas_per_sm must be in 0..32"); c->l96_ctas_per_sm = (int)v; }
  else if (k == "l96_attempt_pairs") { if (v != 1 && v != 2) return fail(c, B200RK_EINVAL, "l96_attempt_pairs must be 1 or 2"); c->l96_attempt_pairs = (int)v; }
  else if (k == "fuse_simpson") c->fuse_simpson = v != 0;
  else if (k == "finish_prefetch") c->finish_prefetch = v != 0;
  else if (k == "stream_simpson") c->stream_simpson = v < 0 ? -1 : (v != 0);
  else if (k == "l2_hints") c->l2_hints = v < 0 ? -1 : (v != 0);
  else if (k == "fused_ctas_per_sm") { if (v < 0) return fail(c, B200RK_EINVAL, "fused_ctas_per_sm must be >= 0"); c->fused_ctas_per_sm = (int)v; }
  else if (k == "peer_timeout_s") {
    if (v < 1) return fail(c, B200RK_EINVAL, "peer_timeout_s must be >= 1");
    c->peer_timeout_s = (int)std::min<int64_t>(v, 86400);
    c->peer_timeout_cycles = (long long)c->peer_timeout_s * (long long)c->clock_khz * 1000ll;
  }
  else if (k == "pool_budget_mb") {
    c->pool_budget_bytes = (size_t)std::max<int64_t>(0, v) << 20;
    if (v == 0) {  // trim now
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      for (auto* p : c->pool) { cudaFree(p->d); delete p; }
      c->pool.clear();
    }
  }
  else return fail(c, B200RK_EINVAL, "unknown knob " + k);
  return B200RK_OK;
}
int b200rk_get(const b200rk_ctx* c, const char* key, int64_t* v) {
  if (!c || !v) return fail(c, B200RK_EINVAL, "null argument");
  std::string k = key ? key : "";
  if (k == "strict_zeros") *v = c->strict_zeros;
  else if (k == "vec_width") *v = c->vec_width;
  else if (k == "ctas_per_sm") *v = c->ctas_per_sm;
  else if (k == "finish_ctas_per_sm") *v = c->finish_ctas_per_sm;
  else if (k == "fuse_pointwise") *v = c->fuse_pointwise;
  else if (k == "spin_readback") *v = c->spin_readback;
  else if (k == "device_loop") *v = c->device_loop;
  else if (k == "p2p") *v = c->p2p;
  else if (k == "fuse_stencil") *v = c->fuse_stencil;
  else if (k == "fuse_stencil_attempt") *v = c->fuse_stencil_attempt;
  else if (k == "l96_peer_halo") *v = c->l96_peer_halo;
  else if (k == "l96_attempt_pairs") *v = c->l96_attempt_pairs;
  else if (k == "fuse_simpson") *v = c->fuse_simpson;
  else if (k == "finish_prefetch") *v = c->finish_prefetch;
  else if (k == "stream_simpson") *v = c->stream_simpson;
  else if (k == "l2_hints") *v = c->l2_hints;
  else if (k == "fused_ctas_per_sm") *v = c->fused_ctas_per_sm;
  else if (k == "profile") *v = c->profile;
  else if (k == "sm_count") *v = c->sm_count;
  else if (k == "pool_budget_mb") *v = (int64_t)(c->pool_budget_bytes >> 20);
  else if (k == "l96_ctas_per_sm") *v = c->l96_ctas_per_sm;
  else if (k == "tstart_copy") *v = c->tstart_copy;
  else if (k == "l96_warp_tiles") *v = c->l96_warp_tiles;
  else if (k == "l96_attempt_threads") *v = c->l96_attempt_threads;
  else if (k == "peer_timeout_s") *v = c->peer_timeout_s;
  else return fail(c, B200RK_EINVAL, "unknown knob " + k);
  return B200RK_OK;
}

int b200rk_profile_reset(b200rk_ctx* c) {
  if (!c) return fail(nullptr, B200RK_EINVAL, "null context");
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (auto& r : c->prof) { c->ev_free.push_back(r.a); c->ev_free.push_back(r.b); }
  c->prof.clear();
  return B200RK_OK;
}
int b200rk_profile_read(b200rk_ctx* c, b200rk_profile* out) {
  if (!c || !out) return fail(c, B200RK_EINVAL, "null argument");
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  std::memset(out, 0, sizeof(*out));
  for (auto& r : c->prof) {
    float ms = 0;
    CUDA_TRY(c, cudaEventElapsedTime(&ms, r.a, r.b));
    out->launches[r.cls]++;
    out->ms[r.cls] += ms;
    out->algorithmic_bytes[r.cls] += r.bytes;
  }
  return B200RK_OK;
}
int b200rk_ctx_stats(const b200rk_ctx* c, b200rk_stats* out) {
  if (!c || !out) return fail(c, B200RK_EINVAL, "null argument");
  std::memset(out, 0, sizeof(*out));
  out->launches = c->launches;
  out->collectives = c->collectives;
  return B200RK_OK;
}

// ---- vectors ----------------------------------------------------------------------------------------
int b200rk_vec_new(b200rk_ctx* c, size_t n_global, b200rk_vec** out) {
  if (!c || !out) return fail(c, B200RK_EINVAL, "null argument");
  return vec_alloc(c, n_global, out);
}
int b200rk_vec_free(b200rk_vec* v) {
  // Freed vectors go back to the context's pool (stream-ordered reuse is safe: one stream per context);
  // beyond the pool budget they are released to the driver.
  if (!v) return B200RK_OK;
  b200rk_ctx* c = v->ctx;
  if (pool_put(v)) return B200RK_OK;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaFree(v->d));
  delete v;
  return B200RK_OK;
}
size_t b200rk_vec_len(const b200rk_vec* v) { return v ? v->n_global : 0; }
size_t b200rk_vec_local_len(const b200rk_vec* v) { return v ? v->n_local : 0; }
size_t b200rk_vec_local_offset(const b200rk_vec* v) { return v ? v->offset : 0; }
double* b200rk_vec_data(const b200rk_vec* v) { return v ? v->d : nullptr; }

int b200rk_vec_upload_local(b200rk_vec* v, const double* h) {
  if (!v || (!h && v->n_local)) return fail(v ? v->ctx : nullptr, B200RK_EINVAL, "null argument");
  b200rk_ctx* c = v->ctx;
  CUDA_TRY(c, cudaMemcpyAsync(v->d, h, v->n_local * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200RK_OK;
}
int b200rk_vec_upload_local_async(b200rk_vec* v, const double* h) {
  if (!v || (!h && v->n_local)) return fail(v ? v->ctx : nullptr, B200RK_EINVAL, "null argument");
  b200rk_ctx* c = v->ctx;
  CUDA_TRY(c, cudaMemcpyAsync(v->d, h, v->n_local * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return B200RK_OK;
}
int b200rk_vec_download_local(const b200rk_vec* v, double* h) {
  if (!v || (!h && v->n_local)) return fail(v ? v->ctx : nullptr, B200RK_EINVAL, "null argument");
  b200rk_ctx* c = v->ctx;
  CUDA_TRY(c, cudaMemcpyAsync(h, v->d, v->n_local * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200RK_OK;
}
int b200rk_vec_upload(b200rk_vec* v, const double* hg) {
  if (!v || !hg) return fail(v ? v->ctx : nullptr, B200RK_EINVAL, "null argument");
  return b200rk_vec_upload_local(v, hg + v->offset);
}
int b200rk_vec_download(const b200rk_vec* v, double* hg) {
  if (!v || !hg) return fail(v ? v->ctx : nullptr, B200RK_EINVAL, "null argument");
  return b200rk_vec_download_local(v, hg + v->offset);
}
int b200rk_vec_copy(b200rk_vec* dst, const b200rk_vec* src) {
  TRY(check_same(dst ? dst->ctx : nullptr, dst, src));
  return vec_copy_raw(dst->ctx, dst, src);
}
}  // extern "C"
