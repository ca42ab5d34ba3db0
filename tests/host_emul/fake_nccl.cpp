// fake_nccl.cpp -> tests/host_emul/_build/libnccl_emul.so — TEST INFRASTRUCTURE ONLY.
// An in-process stand-in for the handful of NCCL entry points libb200rk binds with dlopen (runtime.cu: nccl_bind), so
// that the host-emulated library can run with world > 1 on the CPU: every rank is a THREAD of one process holding its
// own b200rk context; "device" memory is host memory and the emulated streams are synchronous, so a collective is a
// blocking rendezvous of the rank threads. Selected with B200RK_NCCL_LIB by tests/host_emul/multi_rank_emul.py.
// Semantics kept from NCCL: collectives match by call order per communicator; ncclSend / ncclRecv between one pair of
// ranks match in order; inside ncclGroupStart/End nothing blocks until the group ends (sends are buffered).
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {
struct World {
  int n = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  std::vector<const void*> send;
  std::map<std::pair<int, int>, std::deque<std::vector<char>>> mail;   // (src, dst) -> messages in order
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const long g = generation;
    if (++arrived == n) { arrived = 0; ++generation; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != g; });
  }
};
struct Comm { World* w; int rank; };
std::mutex g_registry_mutex;
std::map<std::string, World*> g_registry;
int g_next_id = 1;

size_t type_size(int t) { return t == 0 || t == 1 ? 1 : (t == 2 || t == 3 || t == 7) ? 4 : (t == 6 || t == 9) ? 2 : 8; }

struct P2pOp { bool is_send; void* buf; size_t bytes; int peer; Comm* c; };
thread_local int t_group_depth = 0;
thread_local std::vector<P2pOp> t_group_ops;

void do_send(const P2pOp& op) {
  World* w = op.c->w;
  std::lock_guard<std::mutex> lk(w->m);
  const char* b = static_cast<const char*>(op.buf);
  w->mail[{op.c->rank, op.peer}].emplace_back(b, b + op.bytes);
  w->cv.notify_all();
}
int do_recv(const P2pOp& op) {
  World* w = op.c->w;
  std::unique_lock<std::mutex> lk(w->m);
  auto& q = w->mail[{op.peer, op.c->rank}];
  w->cv.wait(lk, [&] { return !q.empty(); });
  std::vector<char> msg = std::move(q.front());
  q.pop_front();
  if (msg.size() != op.bytes) return 3;   // ncclInvalidArgument: mismatched send / recv sizes
  std::memcpy(op.buf, msg.data(), op.bytes);
  return 0;
}
int run_ops(const std::vector<P2pOp>& ops) {
  for (auto& o : ops) if (o.is_send) do_send(o);
  int rc = 0;
  for (auto& o : ops) if (!o.is_send) { const int r = do_recv(o); if (r) rc = r; }
  return rc;
}

template <class T>
void reduce_into(std::vector<char>& out, const std::vector<const void*>& src, size_t count, int op) {
  T* o = reinterpret_cast<T*>(out.data());
  for (size_t i = 0; i < count; ++i) {
    T acc = static_cast<const T*>(src[0])[i];
    for (size_t r = 1; r < src.size(); ++r) {   // rank order, like a ring all-reduce is NOT — good enough for a test double
      const T v = static_cast<const T*>(src[r])[i];
      acc = op == 0 ? T(acc + v) : op == 1 ? T(acc * v) : op == 2 ? (v > acc ? v : acc) : (v < acc ? v : acc);
    }
    o[i] = acc;
  }
}
}  // namespace

extern "C" {
typedef struct { char internal[128]; } ncclUniqueId;

int ncclGetVersion(int* v) { *v = 22809; return 0; }
const char* ncclGetErrorString(int r) { return r == 0 ? "no error" : r == 3 ? "invalid argument (fake NCCL)" : "fake NCCL error"; }

int ncclGetUniqueId(ncclUniqueId* id) {
  std::lock_guard<std::mutex> lk(g_registry_mutex);
  std::memset(id, 0, sizeof(*id));
  std::snprintf(id->internal, sizeof(id->internal), "emul-world-%d", g_next_id++);
  return 0;
}
int ncclCommInitRank(void** comm, int nranks, ncclUniqueId id, int rank) {
  World* w = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_registry_mutex);
    auto& slot = g_registry[std::string(id.internal)];
    if (!slot) { slot = new World; slot->n = nranks; slot->send.assign(nranks, nullptr); }
    w = slot;
  }
  if (w->n != nranks || rank < 0 || rank >= nranks) return 3;
  *comm = new Comm{w, rank};
  w->barrier();
  return 0;
}
int ncclCommDestroy(void* comm) { delete static_cast<Comm*>(comm); return 0; }

int ncclAllReduce(const void* send, void* recv, size_t count, int type, int op, void* comm, void* /*stream*/) {
  Comm* c = static_cast<Comm*>(comm);
  World* w = c->w;
  { std::lock_guard<std::mutex> lk(w->m); w->send[c->rank] = send; }
  w->barrier();
  std::vector<char> out(count * type_size(type));
  if (type == 8) reduce_into<double>(out, w->send, count, op);
  else if (type == 2) reduce_into<int>(out, w->send, count, op);
  else if (type == 0) reduce_into<signed char>(out, w->send, count, op);
  else return 3;
  w->barrier();   // everybody has read every send buffer (in-place calls overwrite theirs next)
  std::memcpy(recv, out.data(), out.size());
  return 0;
}
int ncclAllGather(const void* send, void* recv, size_t count, int type, void* comm, void* /*stream*/) {
  Comm* c = static_cast<Comm*>(comm);
  World* w = c->w;
  const size_t bytes = count * type_size(type);
  { std::lock_guard<std::mutex> lk(w->m); w->send[c->rank] = send; }
  w->barrier();
  for (int r = 0; r < w->n; ++r) {
    char* dst = static_cast<char*>(recv) + (size_t)r * bytes;
    if (dst != w->send[r]) std::memcpy(dst, w->send[r], bytes);   // in place: a rank's own slot already holds its data
  }
  w->barrier();
  return 0;
}
int ncclGroupStart() { ++t_group_depth; return 0; }
int ncclGroupEnd() {
  if (--t_group_depth > 0) return 0;
  std::vector<P2pOp> ops;
  ops.swap(t_group_ops);
  return run_ops(ops);
}
int ncclSend(const void* buf, size_t count, int type, int peer, void* comm, void* /*stream*/) {
  P2pOp op{true, const_cast<void*>(buf), count * type_size(type), peer, static_cast<Comm*>(comm)};
  if (t_group_depth > 0) { t_group_ops.push_back(op); return 0; }
  return run_ops({op});
}
int ncclRecv(void* buf, size_t count, int type, int peer, void* comm, void* /*stream*/) {
  P2pOp op{false, buf, count * type_size(type), peer, static_cast<Comm*>(comm)};
  if (t_group_depth > 0) { t_group_ops.push_back(op); return 0; }
  return run_ops({op});
}
}  // extern "C"
