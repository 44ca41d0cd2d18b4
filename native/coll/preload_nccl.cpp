// libshipyard_preload — LD_PRELOAD shim giving the NCCL collective entry points to the shipyard kernels.
//
// The task runner preloads this library into every rank of a multi-instance task.
// ncclAllReduce / ncclReduceScatter / ncclAllGather / ncclBroadcast / ncclReduce then resolve here:
// supported calls run on libshipyard_coll's sm_100a kernels ON THE CALLER'S STREAM (no host sync, the
// 1/N of ncclAvg fused into the reduction); anything else is forwarded to the real library found with
// dlsym(RTLD_NEXT).  Communicator management is forwarded too; when no real NCCL is loaded at all the
// shim implements the handful of management calls itself, so plain C programs written against nccl.h
// run with no NCCL installed.
//
// Communicator matching is by CONTENT, never by call order: on the first collective of an NCCL communicator rank 0 draws a
// random session token and the shim broadcasts it with the REAL ncclBroadcast on that very communicator, so exactly the ranks
// of that communicator (however it was created: InitRank, InitRankConfig, Split, ...) attach to one shipyard communicator.
// Grouped point-to-point calls (ncclGroupStart; ncclSend/ncclRecv to every peer; ncclGroupEnd — how PyTorch spells
// all_to_all) are collected and, when they form a dense equal-sized exchange, run as ONE all-to-all kernel; any other group
// is replayed to the real library unchanged.
//
//   SHIPYARD_COLL_DISABLE=1     pass everything through
//   SHIPYARD_PRELOAD_STATS=1    print intercepted / forwarded call counts at exit
//   SHIPYARD_PRELOAD_MIN_BYTES  forward calls smaller than this (default 0)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "sy_coll.h"

namespace {

struct Native { int rank, world, device; std::string session; };   // comm created by the shim itself (no real NCCL)
struct Entry { sy_comm* sy = nullptr; bool tried = false; int world = 0; };

std::mutex g_mu;
std::map<void*, Entry> g_comms;
std::atomic<unsigned long long> g_hit{0}, g_fwd{0};
enum { OP_ALLREDUCE, OP_REDUCESCATTER, OP_ALLGATHER, OP_BROADCAST, OP_REDUCE, OP_ALLTOALL, OP_SENDRECV, OP_COUNT };
const char* const kOpName[OP_COUNT] = {"allreduce", "reduce_scatter", "allgather", "broadcast", "reduce", "alltoall", "send/recv"};
std::atomic<unsigned long long> g_op_hit[OP_COUNT], g_op_fwd[OP_COUNT], g_op_bytes[OP_COUNT];
inline void hit(int op, size_t bytes) { g_hit.fetch_add(1); g_op_hit[op].fetch_add(1); g_op_bytes[op].fetch_add(bytes); }
inline void fwd(int op) { g_fwd.fetch_add(1); g_op_fwd[op].fetch_add(1); }

// ncclGroupStart/End nesting and the point-to-point calls queued inside the outermost group (NCCL groups are per thread)
struct P2P { bool send; const void* sbuf; void* rbuf; size_t count; ncclDataType_t dt; int peer; ncclComm_t comm; cudaStream_t stream; };
thread_local int t_group_depth = 0;
thread_local std::vector<P2P> t_p2p;
bool g_disable = false, g_stats = false;
size_t g_min_bytes = 0, g_heap_bytes_init = 0;

// The real NCCL may sit in the global scope (linked / RTLD_GLOBAL: found by RTLD_NEXT) or only in a local scope
// (a DT_NEEDED of a dlopen'ed extension module, as with PyTorch wheels): then RTLD_NEXT cannot see it and we ask the
// loader for the already-loaded object by soname (RTLD_NOLOAD never loads anything new).
void* real_handle() {
  static std::atomic<void*> h{nullptr};
  void* cur = h.load();
  if (cur) return cur;
  for (const char* so : {"libnccl.so.2", "libnccl.so"}) {
    if (void* x = dlopen(so, RTLD_NOLOAD | RTLD_NOW | RTLD_LOCAL)) {
      if (dlsym(x, "shipyard_preload_hits") == nullptr) { h.store(x); return x; }   // not ourselves under an alias
    }
  }
  return nullptr;
}
void* real_sym(const char* name) {
  if (void* f = dlsym(RTLD_NEXT, name)) return f;
  if (void* h = real_handle()) return dlsym(h, name);
  return nullptr;
}
template <typename F> F real(const char* name) { return reinterpret_cast<F>(real_sym(name)); }
bool have_real() { return real_sym("ncclCommInitRank") != nullptr; }
std::map<void*, bool> g_native;          // communicators created by the shim itself (guarded by g_mu)

struct Init {
  Init() {
    g_disable = getenv("SHIPYARD_COLL_DISABLE") != nullptr;
    g_stats = getenv("SHIPYARD_PRELOAD_STATS") != nullptr;
    if (const char* m = getenv("SHIPYARD_PRELOAD_MIN_BYTES")) g_min_bytes = strtoull(m, nullptr, 10);
    if (const char* m = getenv("SHIPYARD_PRELOAD_HEAP")) { const size_t v = strtoull(m, nullptr, 10); if (v >= (64ul << 20)) g_heap_bytes_init = v; }
  }
  ~Init() {
    if (!g_stats) return;
    std::string d;
    for (int i = 0; i < OP_COUNT; ++i)
      if (g_op_hit[i].load() || g_op_fwd[i].load()) {
        char b[160];
        snprintf(b, sizeof b, " %s=%llu/%llu(%.1fMB)", kOpName[i], g_op_hit[i].load(), g_op_fwd[i].load(), g_op_bytes[i].load() / 1048576.0);
        d += b;
      }
    fprintf(stderr, "[shipyard-preload pid %d] collectives on shipyard kernels: %llu, forwarded to NCCL: %llu  [ours/forwarded(bytes ours):%s]\n",
            getpid(), g_hit.load(), g_fwd.load(), d.c_str());
  }
} g_init;

int map_dtype(ncclDataType_t t) {
  switch ((int)t) {
    case ncclFloat32: return SY_F32; case ncclFloat64: return SY_F64; case ncclFloat16: return SY_F16;
    case ncclBfloat16: return SY_BF16; case ncclInt32: return SY_I32; case ncclInt64: return SY_I64;
    case ncclUint8: case ncclInt8: return SY_U8;
  }
  return -1;
}
size_t dtype_bytes(ncclDataType_t t) {
  switch ((int)t) {
    case ncclInt8: case ncclUint8: return 1; case ncclFloat16: case ncclBfloat16: return 2;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4; case ncclInt64: case ncclUint64: case ncclFloat64: return 8;
  }
  return 1;
}

// Every rank of `comm` receives the same random token from rank 0 THROUGH THE REAL LIBRARY on that communicator.
bool agree_token(ncclComm_t comm, int rank, char* token, size_t cap) {
  auto bc = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)>("ncclBroadcast");
  if (!bc) return false;
  memset(token, 0, cap);
  if (rank == 0) {
    unsigned long long r[2] = {0, 0};
    FILE* f = fopen("/dev/urandom", "rb");
    if (f) { size_t n = fread(r, sizeof r, 1, f); (void)n; fclose(f); }
    snprintf(token, cap, "nccl-%d-%016llx%016llx", getpid(), r[0], r[1]);
  }
  void* d = nullptr; cudaStream_t st = nullptr;
  bool ok = cudaMalloc(&d, cap) == cudaSuccess && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMemcpyAsync(d, token, cap, cudaMemcpyHostToDevice, st) == cudaSuccess;
  ok = ok && bc(d, d, cap, ncclChar, 0, comm, st) == ncclSuccess;
  ok = ok && cudaMemcpyAsync(token, d, cap, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
  if (st) cudaStreamDestroy(st);
  if (d) cudaFree(d);
  token[cap - 1] = 0;
  return ok && token[0] != 0;
}

// Attach a shipyard communicator to an NCCL communicator.  COLLECTIVE over the communicator (token broadcast through the real
// library, then the rendezvous of sy_comm_init), so it runs where NCCL itself requires every rank to be: right after communicator
// creation (the ncclCommInit* / ncclCommSplit hooks below).  Communicators whose creation the shim did not see are attached on
// their first ungrouped collective instead.  Caller holds g_mu.
constexpr int kMaxAttached = 8;
size_t g_heap_bytes = 384ul << 20;        // symmetric heap per attached communicator: 128 MB staging + mailboxes (SHIPYARD_PRELOAD_HEAP)
sy_comm* attach_locked(ncclComm_t comm) {
  Entry& e = g_comms[(void*)comm];
  if (e.tried) return e.sy;
  e.tried = true;
  int world = 0, rank = 0, dev = 0;
  std::string session;
  if (!g_native.count((void*)comm)) {
    if (!have_real()) return nullptr;                                // not ours and no NCCL to ask: never guess
    auto cnt = real<ncclResult_t (*)(const ncclComm_t, int*)>("ncclCommCount");
    auto urk = real<ncclResult_t (*)(const ncclComm_t, int*)>("ncclCommUserRank");
    auto cud = real<ncclResult_t (*)(const ncclComm_t, int*)>("ncclCommCuDevice");
    if (!cnt || !urk || !cud || cnt(comm, &world) != ncclSuccess || urk(comm, &rank) != ncclSuccess || cud(comm, &dev) != ncclSuccess) return nullptr;
    e.world = world;
    if (world < 2 || world > 8) return nullptr;                      // single rank / multi-box communicators stay on NCCL
    int attached = 0;
    for (auto& kv : g_comms) attached += kv.second.sy != nullptr;
    if (attached >= kMaxAttached) return nullptr;                    // (same decision on every rank: same creation history)
    int cur = 0; cudaGetDevice(&cur);
    if (cur != dev) cudaSetDevice(dev);
    char token[64];
    const bool ok = agree_token(comm, rank, token, sizeof token);
    if (cur != dev) cudaSetDevice(cur);
    if (!ok) { fprintf(stderr, "[shipyard-preload] session token exchange failed; communicator stays on NCCL\n"); return nullptr; }
    session = token;
  } else {
    Native* n = reinterpret_cast<Native*>(comm);
    world = n->world; rank = n->rank; dev = n->device;
    session = n->session + "-w" + std::to_string(world);            // the unique id string: identical on every rank by construction
    e.world = world;
    if (world < 2 || world > 8) return nullptr;
  }
  sy_comm* c = nullptr;
  if (sy_comm_init(&c, rank, world, session.c_str(), dev, g_heap_bytes_init ? g_heap_bytes_init : g_heap_bytes, SY_TRANSPORT_AUTO) != SY_OK) {
    fprintf(stderr, "[shipyard-preload] cannot attach to communicator (%s); falling back to NCCL\n", sy_last_error());
    return nullptr;
  }
  e.sy = c;
  if (g_stats) fprintf(stderr, "[shipyard-preload pid %d] rank %d/%d attached (session %s, transport %d, multicast %d)\n", getpid(), rank, world,
                       session.c_str(), sy_comm_transport(c), sy_comm_has_multicast(c));
  return c;
}

// creation hook: attach eagerly (every rank of the new communicator is inside the same creation call)
void on_created(ncclComm_t comm) {
  if (g_disable || !comm || t_group_depth > 0 || getenv("SHIPYARD_PRELOAD_LAZY")) return;   // (grouped creation completes at GroupEnd)
  auto aerr = real<ncclResult_t (*)(ncclComm_t, ncclResult_t*)>("ncclCommGetAsyncError");
  ncclResult_t st = ncclSuccess;
  if (aerr && (aerr(comm, &st) != ncclSuccess || st != ncclSuccess)) return;      // non-blocking init still in progress: attach lazily
  std::lock_guard<std::mutex> lk(g_mu);
  attach_locked(comm);
}

sy_comm* bind(ncclComm_t comm) {
  if (g_disable) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_comms.find((void*)comm);
  if (it != g_comms.end() && it->second.tried) return it->second.sy;
  // first sight of a communicator created before the shim could see it: the token exchange is a real NCCL call followed by a
  // stream sync, which inside an open group would never run — such a call is forwarded and a later ungrouped one attaches
  if (t_group_depth > 0 && !g_native.count((void*)comm)) return nullptr;
  return attach_locked(comm);
}

bool is_native(const void* comm) { std::lock_guard<std::mutex> lk(g_mu); return g_native.count(const_cast<void*>(comm)) > 0; }

}  // namespace

extern "C" {

// ---- collectives ------------------------------------------------------------------------------------
ncclResult_t ncclAllReduce(const void* sendbuff, void* recvbuff, size_t count, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t comm,
                           cudaStream_t stream) {
  const int sdt = map_dtype(dt);
  int sop = -1; float scale = 1.0f;
  sy_comm* c = (sdt >= 0 && sdt != SY_U8 && count * dtype_bytes(dt) >= g_min_bytes) ? bind(comm) : nullptr;
  if (c) {
    switch ((int)op) {
      case ncclSum: sop = SY_SUM; break; case ncclProd: sop = SY_PROD; break; case ncclMax: sop = SY_MAX; break; case ncclMin: sop = SY_MIN; break;
      case ncclAvg: sop = SY_SUM; scale = 1.0f / (float)sy_comm_world(c); break;     // the averaging is fused into the reduction kernel
    }
    if (sop >= 0 && !(scale != 1.0f && (sdt == SY_I32 || sdt == SY_I64)) &&
        sy_allreduce(c, sendbuff, recvbuff, count, sdt, sdt, scale, sop, SY_ALGO_AUTO, stream) == SY_OK) {
      hit(OP_ALLREDUCE, count * dtype_bytes(dt));
      return ncclSuccess;
    }
  }
  fwd(OP_ALLREDUCE);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)>("ncclAllReduce");
  return f ? f(sendbuff, recvbuff, count, dt, op, comm, stream) : ncclInvalidUsage;
}

ncclResult_t ncclReduceScatter(const void* sendbuff, void* recvbuff, size_t recvcount, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t comm,
                               cudaStream_t stream) {
  const int sdt = map_dtype(dt);
  sy_comm* c = (sdt >= 0 && sdt != SY_U8 && (op == ncclSum || op == ncclAvg || op == ncclMax || op == ncclMin)) ? bind(comm) : nullptr;
  if (c) {
    const int sop = op == ncclMax ? SY_MAX : op == ncclMin ? SY_MIN : SY_SUM;
    const float scale = op == ncclAvg ? 1.0f / (float)sy_comm_world(c) : 1.0f;
    if (sy_reduce_scatter(c, sendbuff, recvbuff, recvcount, sdt, sdt, scale, sop, stream) == SY_OK) { hit(OP_REDUCESCATTER, recvcount * dtype_bytes(dt)); return ncclSuccess; }
  }
  fwd(OP_REDUCESCATTER);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)>("ncclReduceScatter");
  return f ? f(sendbuff, recvbuff, recvcount, dt, op, comm, stream) : ncclInvalidUsage;
}

ncclResult_t ncclAllGather(const void* sendbuff, void* recvbuff, size_t sendcount, ncclDataType_t dt, ncclComm_t comm, cudaStream_t stream) {
  sy_comm* c = bind(comm);
  if (c && sy_allgather(c, sendbuff, recvbuff, sendcount * dtype_bytes(dt), SY_U8, stream) == SY_OK) { hit(OP_ALLGATHER, sendcount * dtype_bytes(dt)); return ncclSuccess; }
  fwd(OP_ALLGATHER);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)>("ncclAllGather");
  return f ? f(sendbuff, recvbuff, sendcount, dt, comm, stream) : ncclInvalidUsage;
}

ncclResult_t ncclBroadcast(const void* sendbuff, void* recvbuff, size_t count, ncclDataType_t dt, int root, ncclComm_t comm, cudaStream_t stream) {
  sy_comm* c = bind(comm);
  if (c) {
    const void* in = sy_comm_rank(c) == root ? sendbuff : recvbuff;
    if (sy_broadcast(c, in, recvbuff, count * dtype_bytes(dt), SY_U8, root, stream) == SY_OK) { hit(OP_BROADCAST, count * dtype_bytes(dt)); return ncclSuccess; }
  }
  fwd(OP_BROADCAST);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)>("ncclBroadcast");
  return f ? f(sendbuff, recvbuff, count, dt, root, comm, stream) : ncclInvalidUsage;
}

ncclResult_t ncclBcast(void* buff, size_t count, ncclDataType_t dt, int root, ncclComm_t comm, cudaStream_t stream) {
  return ncclBroadcast(buff, buff, count, dt, root, comm, stream);
}

ncclResult_t ncclReduce(const void* sendbuff, void* recvbuff, size_t count, ncclDataType_t dt, ncclRedOp_t op, int root, ncclComm_t comm,
                        cudaStream_t stream) {
  const int sdt = map_dtype(dt);
  sy_comm* c = (sdt >= 0 && sdt != SY_U8 && (op == ncclSum || op == ncclMax || op == ncclMin || op == ncclProd)) ? bind(comm) : nullptr;
  if (c) {
    const int sop = op == ncclMax ? SY_MAX : op == ncclMin ? SY_MIN : op == ncclProd ? SY_PROD : SY_SUM;
    if (sy_reduce(c, sendbuff, recvbuff, count, sdt, sop, root, stream) == SY_OK) { hit(OP_REDUCE, count * dtype_bytes(dt)); return ncclSuccess; }
  }
  fwd(OP_REDUCE);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t)>("ncclReduce");
  return f ? f(sendbuff, recvbuff, count, dt, op, root, comm, stream) : ncclInvalidUsage;
}

// ---- NCCL >= 2.28 single-call all-to-all / gather / scatter (what torch.distributed uses when the library has them) ----------
ncclResult_t ncclAlltoAll(const void* sendbuff, void* recvbuff, size_t count, ncclDataType_t dt, ncclComm_t comm, cudaStream_t stream) {
  sy_comm* c = bind(comm);
  const size_t bytes = count * dtype_bytes(dt);
  if (c && sy_alltoall(c, sendbuff, recvbuff, bytes, SY_U8, stream) == SY_OK) { hit(OP_ALLTOALL, bytes * (size_t)sy_comm_world(c)); return ncclSuccess; }
  fwd(OP_ALLTOALL);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)>("ncclAlltoAll");
  return f ? f(sendbuff, recvbuff, count, dt, comm, stream) : ncclInvalidUsage;
}
ncclResult_t ncclGather(const void* sendbuff, void* recvbuff, size_t count, ncclDataType_t dt, int root, ncclComm_t comm, cudaStream_t stream) {
  sy_comm* c = bind(comm);
  if (c && sy_gather(c, sendbuff, recvbuff, count * dtype_bytes(dt), SY_U8, root, stream) == SY_OK) { hit(OP_REDUCE, count * dtype_bytes(dt)); return ncclSuccess; }
  fwd(OP_REDUCE);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)>("ncclGather");
  return f ? f(sendbuff, recvbuff, count, dt, root, comm, stream) : ncclInvalidUsage;
}
ncclResult_t ncclScatter(const void* sendbuff, void* recvbuff, size_t count, ncclDataType_t dt, int root, ncclComm_t comm, cudaStream_t stream) {
  sy_comm* c = bind(comm);
  if (c && sy_scatter(c, sendbuff, recvbuff, count * dtype_bytes(dt), SY_U8, root, stream) == SY_OK) { hit(OP_BROADCAST, count * dtype_bytes(dt)); return ncclSuccess; }
  fwd(OP_BROADCAST);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)>("ncclScatter");
  return f ? f(sendbuff, recvbuff, count, dt, root, comm, stream) : ncclInvalidUsage;
}

// ---- communicator management: forwarded when a real NCCL exists, native otherwise ----------------------
ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  if (auto f = real<ncclResult_t (*)(ncclUniqueId*)>("ncclGetUniqueId")) return f(id);
  memset(id, 0, sizeof *id);
  snprintf(id->internal, sizeof id->internal, "sy-%d-%ld", getpid(), (long)random());
  return ncclSuccess;
}

ncclResult_t ncclCommInitRankConfig(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank, ncclConfig_t* config) {
  auto f = real<ncclResult_t (*)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*)>("ncclCommInitRankConfig");
  if (!f) return ncclCommInitRank(comm, nranks, id, rank);
  const ncclResult_t r = f(comm, nranks, id, rank, config);
  if (r == ncclSuccess && comm) on_created(*comm);
  return r;
}
ncclResult_t ncclCommInitRankScalable(ncclComm_t* comm, int nranks, int rank, int nid, ncclUniqueId* ids, ncclConfig_t* config) {
  auto f = real<ncclResult_t (*)(ncclComm_t*, int, int, int, ncclUniqueId*, ncclConfig_t*)>("ncclCommInitRankScalable");
  if (!f) return nid > 0 && ids ? ncclCommInitRankConfig(comm, nranks, ids[0], rank, config) : ncclInvalidUsage;
  const ncclResult_t r = f(comm, nranks, rank, nid, ids, config);
  if (r == ncclSuccess && comm) on_created(*comm);
  return r;
}
ncclResult_t ncclCommSplit(ncclComm_t comm, int color, int key, ncclComm_t* newcomm, ncclConfig_t* config) {
  auto f = real<ncclResult_t (*)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*)>("ncclCommSplit");
  if (!f) return ncclInvalidUsage;
  const ncclResult_t r = f(comm, color, key, newcomm, config);
  if (r == ncclSuccess && newcomm && *newcomm) on_created(*newcomm);
  return r;
}

ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) {
  if (auto f = real<ncclResult_t (*)(ncclComm_t*, int, ncclUniqueId, int)>("ncclCommInitRank")) {
    const ncclResult_t r = f(comm, nranks, id, rank);
    if (r == ncclSuccess && comm) on_created(*comm);
    return r;
  }
  Native* n = new Native();
  n->rank = rank; n->world = nranks; n->session = std::string(id.internal, strnlen(id.internal, sizeof id.internal));
  cudaGetDevice(&n->device);
  *comm = reinterpret_cast<ncclComm_t>(n);
  { std::lock_guard<std::mutex> lk(g_mu); g_native[(void*)n] = true; }
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_comms.find((void*)comm);
    if (it != g_comms.end()) { if (it->second.sy) sy_comm_destroy(it->second.sy); g_comms.erase(it); }
  }
  bool native = false;
  { std::lock_guard<std::mutex> lk(g_mu); native = g_native.erase((void*)comm) > 0; }
  if (native) { delete reinterpret_cast<Native*>(comm); return ncclSuccess; }
  if (auto f = real<ncclResult_t (*)(ncclComm_t)>("ncclCommDestroy")) return f(comm);
  return ncclInvalidArgument;
}

ncclResult_t ncclCommCount(const ncclComm_t comm, int* count) {
  if (!is_native(comm)) { auto f = real<ncclResult_t (*)(const ncclComm_t, int*)>("ncclCommCount"); return f ? f(comm, count) : ncclInvalidArgument; }
  *count = reinterpret_cast<Native*>(comm)->world; return ncclSuccess;
}
ncclResult_t ncclCommUserRank(const ncclComm_t comm, int* rank) {
  if (!is_native(comm)) { auto f = real<ncclResult_t (*)(const ncclComm_t, int*)>("ncclCommUserRank"); return f ? f(comm, rank) : ncclInvalidArgument; }
  *rank = reinterpret_cast<Native*>(comm)->rank; return ncclSuccess;
}
ncclResult_t ncclCommCuDevice(const ncclComm_t comm, int* dev) {
  if (!is_native(comm)) { auto f = real<ncclResult_t (*)(const ncclComm_t, int*)>("ncclCommCuDevice"); return f ? f(comm, dev) : ncclInvalidArgument; }
  *dev = reinterpret_cast<Native*>(comm)->device; return ncclSuccess;
}
// ---- groups and point-to-point ---------------------------------------------------------------------------
// Point-to-point calls inside a group are queued; at the outermost ncclGroupEnd a dense equal-sized exchange on one communicator
// and one stream (every peer exactly one send and one recv, block r of one contiguous send / recv buffer for peer r: torch's
// all_to_all_single) runs as ONE shipyard all-to-all kernel.  Everything else is replayed to the real library in call order.
static ncclResult_t flush_p2p() {
  std::vector<P2P> q; q.swap(t_p2p);
  if (q.empty()) return ncclSuccess;
  auto rsend = real<ncclResult_t (*)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)>("ncclSend");
  auto rrecv = real<ncclResult_t (*)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)>("ncclRecv");
  bool dense = !g_disable;
  const ncclComm_t comm = q[0].comm; const cudaStream_t st = q[0].stream;
  const size_t bytes = q[0].count * dtype_bytes(q[0].dt);
  for (auto& o : q) dense = dense && o.comm == comm && o.stream == st && o.count * dtype_bytes(o.dt) == bytes;
  sy_comm* c = nullptr;
  if (dense && bytes > 0) {
    Entry e; { std::lock_guard<std::mutex> lk(g_mu); auto it = g_comms.find((void*)comm); if (it != g_comms.end()) e = it->second; }
    c = e.sy;                                          // binding needs a real collective outside a group: only already-bound communicators
  }
  if (c) {
    const int world = sy_comm_world(c);
    std::vector<const void*> sp(world, nullptr); std::vector<void*> rp(world, nullptr);
    dense = (int)q.size() == 2 * world;
    for (auto& o : q) {
      if (!dense || o.peer < 0 || o.peer >= world) { dense = false; break; }
      if (o.send) { if (sp[o.peer]) dense = false; sp[o.peer] = o.sbuf; } else { if (rp[o.peer]) dense = false; rp[o.peer] = o.rbuf; }
    }
    for (int r = 0; dense && r < world; ++r)
      dense = sp[r] && rp[r] && (const char*)sp[r] == (const char*)sp[0] + (size_t)r * bytes && (char*)rp[r] == (char*)rp[0] + (size_t)r * bytes;
    if (dense && sy_alltoall(c, sp[0], rp[0], bytes, SY_U8, st) == SY_OK) { hit(OP_ALLTOALL, bytes * (size_t)world); return ncclSuccess; }
  }
  if (!rsend || !rrecv) return ncclInvalidUsage;
  fwd(q.size() > 2 ? OP_ALLTOALL : OP_SENDRECV);
  for (auto& o : q) {
    const ncclResult_t r = o.send ? rsend(o.sbuf, o.count, o.dt, o.peer, o.comm, o.stream) : rrecv(o.rbuf, o.count, o.dt, o.peer, o.comm, o.stream);
    if (r != ncclSuccess) return r;
  }
  return ncclSuccess;
}

ncclResult_t ncclGroupStart() {
  ++t_group_depth;
  if (auto f = real<ncclResult_t (*)()>("ncclGroupStart")) return f();
  return ncclSuccess;
}
ncclResult_t ncclGroupEnd() {
  ncclResult_t rc = ncclSuccess;
  if (t_group_depth > 0 && --t_group_depth == 0) rc = flush_p2p();      // still inside the real group: replayed calls join it
  if (auto f = real<ncclResult_t (*)()>("ncclGroupEnd")) { const ncclResult_t r2 = f(); return rc != ncclSuccess ? rc : r2; }
  return rc;
}
ncclResult_t ncclSend(const void* sendbuff, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t stream) {
  if (t_group_depth > 0 && !is_native(comm)) { t_p2p.push_back(P2P{true, sendbuff, nullptr, count, dt, peer, comm, stream}); return ncclSuccess; }
  fwd(OP_SENDRECV);
  auto f = real<ncclResult_t (*)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)>("ncclSend");
  return f ? f(sendbuff, count, dt, peer, comm, stream) : ncclInvalidUsage;
}
ncclResult_t ncclRecv(void* recvbuff, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t stream) {
  if (t_group_depth > 0 && !is_native(comm)) { t_p2p.push_back(P2P{false, nullptr, recvbuff, count, dt, peer, comm, stream}); return ncclSuccess; }
  fwd(OP_SENDRECV);
  auto f = real<ncclResult_t (*)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)>("ncclRecv");
  return f ? f(recvbuff, count, dt, peer, comm, stream) : ncclInvalidUsage;
}
// communicator teardown paths other than ncclCommDestroy: drop our attachment first
static void detach(ncclComm_t comm) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_comms.find((void*)comm);
  if (it != g_comms.end()) { if (it->second.sy) sy_comm_destroy(it->second.sy); g_comms.erase(it); }
}
ncclResult_t ncclCommAbort(ncclComm_t comm) {
  detach(comm);
  if (is_native(comm)) return ncclCommDestroy(comm);
  auto f = real<ncclResult_t (*)(ncclComm_t)>("ncclCommAbort");
  return f ? f(comm) : ncclInvalidArgument;
}
const char* ncclGetErrorString(ncclResult_t r) {
  if (auto f = real<const char* (*)(ncclResult_t)>("ncclGetErrorString")) return f(r);
  return r == ncclSuccess ? "no error" : "shipyard-preload: error";
}
ncclResult_t ncclGetVersion(int* v) {
  if (auto f = real<ncclResult_t (*)(int*)>("ncclGetVersion")) return f(v);
  *v = 22809; return ncclSuccess;
}

unsigned long long shipyard_preload_hits() { return g_hit.load(); }
unsigned long long shipyard_preload_op_hits(int op) { return op >= 0 && op < OP_COUNT ? g_op_hit[op].load() : 0; }
unsigned long long shipyard_preload_op_forwards(int op) { return op >= 0 && op < OP_COUNT ? g_op_fwd[op].load() : 0; }
unsigned long long shipyard_preload_forwards() { return g_fwd.load(); }

}  // extern "C"
