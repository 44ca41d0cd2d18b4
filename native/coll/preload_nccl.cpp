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
#include "sy_coll.h"

namespace {

struct Native { int rank, world, device; std::string session; };   // comm created by the shim itself (no real NCCL)
struct Entry { sy_comm* sy = nullptr; bool tried = false; int world = 0; };

std::mutex g_mu;
std::map<void*, Entry> g_comms;
std::map<int, int> g_ordinal;            // world size -> number of communicators seen (agreed across ranks by call order)
std::atomic<unsigned long long> g_hit{0}, g_fwd{0};
bool g_disable = false, g_stats = false;
size_t g_min_bytes = 0;

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
  }
  ~Init() {
    if (g_stats) fprintf(stderr, "[shipyard-preload pid %d] collectives on shipyard kernels: %llu, forwarded to NCCL: %llu\n", getpid(),
                         g_hit.load(), g_fwd.load());
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

// lazily bind a shipyard communicator to an NCCL communicator (collective: every rank gets here on its first call)
sy_comm* bind(ncclComm_t comm) {
  if (g_disable) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  Entry& e = g_comms[(void*)comm];
  if (e.tried) return e.sy;
  e.tried = true;
  int world = 0, rank = 0, dev = 0;
  std::string base;
  if (!g_native.count((void*)comm) && have_real()) {
    auto cnt = real<ncclResult_t (*)(const ncclComm_t, int*)>("ncclCommCount");
    auto urk = real<ncclResult_t (*)(const ncclComm_t, int*)>("ncclCommUserRank");
    auto cud = real<ncclResult_t (*)(const ncclComm_t, int*)>("ncclCommCuDevice");
    if (!cnt || !urk || !cud || cnt(comm, &world) != ncclSuccess || urk(comm, &rank) != ncclSuccess || cud(comm, &dev) != ncclSuccess) return nullptr;
    const char* s = getenv("SHIPYARD_COLL_SESSION");
    const char* p = getenv("MASTER_PORT");
    base = std::string(s && *s ? s : "nccl") + "-" + (p ? p : "0");
  } else {
    if (!g_native.count((void*)comm)) return nullptr;               // not ours and no NCCL to ask: never guess
    Native* n = reinterpret_cast<Native*>(comm);
    world = n->world; rank = n->rank; dev = n->device; base = n->session;
  }
  if (world < 2 || world > 8) return nullptr;                      // single rank / multi-box communicators stay on NCCL
  const int ord = g_ordinal[world]++;
  std::string session = base + "-w" + std::to_string(world) + "-c" + std::to_string(ord);
  sy_comm* c = nullptr;
  if (sy_comm_init(&c, rank, world, session.c_str(), dev, 0, SY_TRANSPORT_AUTO) != SY_OK) {
    fprintf(stderr, "[shipyard-preload] cannot attach to communicator (%s); falling back to NCCL\n", sy_last_error());
    return nullptr;
  }
  e.sy = c; e.world = world;
  return c;
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
      g_hit.fetch_add(1);
      return ncclSuccess;
    }
  }
  g_fwd.fetch_add(1);
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
    if (sy_reduce_scatter(c, sendbuff, recvbuff, recvcount, sdt, sdt, scale, sop, stream) == SY_OK) { g_hit.fetch_add(1); return ncclSuccess; }
  }
  g_fwd.fetch_add(1);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)>("ncclReduceScatter");
  return f ? f(sendbuff, recvbuff, recvcount, dt, op, comm, stream) : ncclInvalidUsage;
}

ncclResult_t ncclAllGather(const void* sendbuff, void* recvbuff, size_t sendcount, ncclDataType_t dt, ncclComm_t comm, cudaStream_t stream) {
  sy_comm* c = bind(comm);
  if (c && sy_allgather(c, sendbuff, recvbuff, sendcount * dtype_bytes(dt), SY_U8, stream) == SY_OK) { g_hit.fetch_add(1); return ncclSuccess; }
  g_fwd.fetch_add(1);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)>("ncclAllGather");
  return f ? f(sendbuff, recvbuff, sendcount, dt, comm, stream) : ncclInvalidUsage;
}

ncclResult_t ncclBroadcast(const void* sendbuff, void* recvbuff, size_t count, ncclDataType_t dt, int root, ncclComm_t comm, cudaStream_t stream) {
  sy_comm* c = bind(comm);
  if (c) {
    const void* in = sy_comm_rank(c) == root ? sendbuff : recvbuff;
    if (sy_broadcast(c, in, recvbuff, count * dtype_bytes(dt), SY_U8, root, stream) == SY_OK) { g_hit.fetch_add(1); return ncclSuccess; }
  }
  g_fwd.fetch_add(1);
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
    if (sy_reduce(c, sendbuff, recvbuff, count, sdt, sop, root, stream) == SY_OK) { g_hit.fetch_add(1); return ncclSuccess; }
  }
  g_fwd.fetch_add(1);
  auto f = real<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t)>("ncclReduce");
  return f ? f(sendbuff, recvbuff, count, dt, op, root, comm, stream) : ncclInvalidUsage;
}

// ---- communicator management: forwarded when a real NCCL exists, native otherwise ----------------------
ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  if (auto f = real<ncclResult_t (*)(ncclUniqueId*)>("ncclGetUniqueId")) return f(id);
  memset(id, 0, sizeof *id);
  snprintf(id->internal, sizeof id->internal, "sy-%d-%ld", getpid(), (long)random());
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) {
  if (auto f = real<ncclResult_t (*)(ncclComm_t*, int, ncclUniqueId, int)>("ncclCommInitRank")) return f(comm, nranks, id, rank);
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
ncclResult_t ncclGroupStart() { if (auto f = real<ncclResult_t (*)()>("ncclGroupStart")) return f(); return ncclSuccess; }
ncclResult_t ncclGroupEnd() { if (auto f = real<ncclResult_t (*)()>("ncclGroupEnd")) return f(); return ncclSuccess; }
const char* ncclGetErrorString(ncclResult_t r) {
  if (auto f = real<const char* (*)(ncclResult_t)>("ncclGetErrorString")) return f(r);
  return r == ncclSuccess ? "no error" : "shipyard-preload: error";
}
ncclResult_t ncclGetVersion(int* v) {
  if (auto f = real<ncclResult_t (*)(int*)>("ncclGetVersion")) return f(v);
  *v = 22809; return ncclSuccess;
}

unsigned long long shipyard_preload_hits() { return g_hit.load(); }
unsigned long long shipyard_preload_forwards() { return g_fwd.load(); }

}  // extern "C"
