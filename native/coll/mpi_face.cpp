// libshipyard_mpi — the MPI face of the shipyard collectives.
//
// MPI_* symbols the recipe workloads call resolve here: device buffers go
// to the sm_100a kernels of libshipyard_coll, host buffers to its shared-memory stub transport.
// Ranks are discovered from the task runner's environment (no mpirun / ssh: the reference's
// multi-instance containers run sshd on port 23 for that, convoy/settings.py:4391-4443).
// Point-to-point uses per-(src,dst) mailboxes in the host symmetric heap with a non-blocking
// progress engine, so Irecv/Isend/Waitall patterns of any size cannot deadlock.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>
#include <atomic>
#include <string>
#include <vector>
#include "mpi.h"
#include "sy_coll.h"

namespace {

struct PairBox {                       // one per ordered (src -> dst), lives in dst's heap
  std::atomic<uint64_t> seq;           // chunks published by the sender
  std::atomic<uint64_t> ack;           // chunks consumed by the receiver
  int tag; int pad; uint64_t total; uint64_t len;   // header of the current chunk
  char payload[1];
};
const size_t BOX_PAYLOAD = 1u << 20;
const size_t BOX_BYTES = 4096 + BOX_PAYLOAD;

struct Req {
  bool active = false, is_send = false, done = false;
  char* buf = nullptr; size_t bytes = 0, off = 0; int peer = 0, tag = 0;
  bool device = false; std::vector<char> bounce;
  MPI_Status st{};
};

struct State {
  bool inited = false, finalized = false;
  int rank = 0, world = 1, device = -1;
  sy_comm* host = nullptr;             // stub transport (always)
  sy_comm* dev = nullptr;              // GPU transport (lazy)
  char* boxes = nullptr;               // my inbound mailboxes: world x BOX_BYTES (host heap)
  size_t boxes_off = 0;
  std::vector<uint64_t> sent, rcvd;    // per-peer chunk counters
  std::vector<Req> reqs;
  std::string session;
  cudaStream_t stream = nullptr;
} g;

int env_int(const char* const* names, int dflt) {
  for (; *names; ++names) { const char* v = getenv(*names); if (v && *v) return atoi(v); }
  return dflt;
}
[[noreturn]] void die(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  fprintf(stderr, "[shipyard-mpi rank %d] ", g.rank); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
  va_end(ap);
  _exit(70);
}
#define SYCHECK(call) do { int _r = (call); if (_r != SY_OK) die("%s failed (%d): %s", #call, _r, sy_last_error()); } while (0)

size_t dt_size(MPI_Datatype dt) {
  switch (dt) {
    case MPI_CHAR: case MPI_SIGNED_CHAR: case MPI_UNSIGNED_CHAR: case MPI_BYTE: case MPI_UINT8_T: return 1;
    case MPI_SHORT: case MPI_UNSIGNED_SHORT: case MPIX_BFLOAT16: case MPIX_FLOAT16: return 2;
    case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: case MPI_INT32_T: case MPI_UINT32_T: return 4;
    case MPI_LONG: case MPI_UNSIGNED_LONG: case MPI_LONG_LONG: case MPI_LONG_LONG_INT: case MPI_UNSIGNED_LONG_LONG:
    case MPI_DOUBLE: case MPI_INT64_T: case MPI_UINT64_T: return 8;
  }
  return 0;
}
int sy_dt(MPI_Datatype dt) {   // reduction-capable mapping
  switch (dt) {
    case MPI_FLOAT: return SY_F32; case MPI_DOUBLE: return SY_F64;
    case MPI_INT: case MPI_INT32_T: return SY_I32;
    case MPI_LONG: case MPI_LONG_LONG: case MPI_LONG_LONG_INT: case MPI_INT64_T: return SY_I64;
    case MPIX_BFLOAT16: return SY_BF16; case MPIX_FLOAT16: return SY_F16;
  }
  return -1;
}
int sy_opof(MPI_Op op) { return op == MPI_SUM ? SY_SUM : op == MPI_MAX ? SY_MAX : op == MPI_MIN ? SY_MIN : op == MPI_PROD ? SY_PROD : -1; }

bool is_device_ptr(const void* p) {
  if (!p || p == MPI_IN_PLACE) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

sy_comm* dev_comm() {
  if (g.dev) return g.dev;
  if (g.device < 0) {
    static const char* n[] = {"SHIPYARD_GPU", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", nullptr};
    g.device = env_int(n, 0);
  }
  std::string s = g.session + "-dev";
  SYCHECK(sy_comm_init(&g.dev, g.rank, g.world, s.c_str(), g.device, 0, SY_TRANSPORT_AUTO));
  cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking);
  return g.dev;
}
bool g_async = false;      // MPIX_Set_device_async(1): device-buffer collectives return after the enqueue (the caller times / syncs)
void dev_sync() {
  if (g_async) return;
  cudaError_t e = cudaStreamSynchronize(g.stream);
  if (e != cudaSuccess) die("stream sync: %s", cudaGetErrorString(e));
  if (sy_comm_status(g.dev) != 0) die("collective watchdog fired: a peer rank did not arrive");
}

PairBox* box_on(int dst, int src) {     // mailbox for src->dst (in dst's heap)
  return (PairBox*)((char*)sy_heap_base(g.host, dst) + g.boxes_off + (size_t)src * BOX_BYTES);
}

// ---- non-blocking progress on one request; returns true when it completed ----------------------
bool progress(Req& r) {
  if (r.done) return true;
  if (r.is_send) {
    PairBox* b = box_on(r.peer, g.rank);
    while (r.off < r.bytes || (r.bytes == 0 && r.off == 0)) {
      if (b->ack.load(std::memory_order_acquire) != g.sent[r.peer]) return false;   // previous chunk not consumed yet
      size_t n = r.bytes - r.off < BOX_PAYLOAD ? r.bytes - r.off : BOX_PAYLOAD;
      memcpy(b->payload, r.buf + r.off, n);
      b->tag = r.tag; b->total = r.bytes; b->len = n;
      g.sent[r.peer] += 1;
      b->seq.store(g.sent[r.peer], std::memory_order_release);
      r.off += n;
      if (r.bytes == 0) { r.off = 1; break; }
    }
    r.done = true;
    return true;
  }
  PairBox* b = box_on(g.rank, r.peer);
  for (;;) {
    if (b->seq.load(std::memory_order_acquire) == g.rcvd[r.peer]) return false;     // nothing new
    if (r.tag != MPI_ANY_TAG && b->tag != r.tag)
      die("MPI_Recv tag mismatch from rank %d: expected %d, got %d (messages are matched in order per peer)", r.peer, r.tag, b->tag);
    if (b->total > r.bytes) die("MPI_Recv from rank %d: message of %zu bytes truncated (buffer %zu)", r.peer, (size_t)b->total, r.bytes);
    size_t n = (size_t)b->len, total = (size_t)b->total;
    memcpy(r.buf + r.off, b->payload, n);
    r.st.MPI_SOURCE = r.peer; r.st.MPI_TAG = b->tag; r.st._count = (int)total;
    g.rcvd[r.peer] += 1;
    b->ack.store(g.rcvd[r.peer], std::memory_order_release);
    r.off += n;
    if (r.off >= total) { r.done = true; return true; }
  }
}

int new_req(bool is_send, void* buf, size_t bytes, int peer, int tag) {
  int id = -1;
  for (size_t i = 1; i < g.reqs.size(); ++i) if (!g.reqs[i].active) { id = (int)i; break; }
  if (id < 0) { g.reqs.emplace_back(); id = (int)g.reqs.size() - 1; }
  Req& r = g.reqs[id];
  r = Req(); r.active = true; r.is_send = is_send; r.bytes = bytes; r.peer = peer; r.tag = tag;
  r.device = is_device_ptr(buf);
  if (r.device) {            // device buffers are bounced through host memory for point-to-point
    r.bounce.resize(bytes ? bytes : 1);
    if (is_send) cudaMemcpy(r.bounce.data(), buf, bytes, cudaMemcpyDeviceToHost);
    r.buf = r.bounce.data();
  } else r.buf = (char*)buf;
  return id;
}
struct DevDst { void* p; };
std::vector<DevDst> g_devdst;

void wait_all(int n, MPI_Request* ids, MPI_Status* sts) {
  double t0 = MPI_Wtime(); unsigned spins = 0;
  for (;;) {
    bool all = true;
    for (int i = 0; i < n; ++i) {
      if (ids[i] == MPI_REQUEST_NULL) continue;
      if (!progress(g.reqs[ids[i]])) all = false;
    }
    if (all) break;
    if ((++spins & 0x3fff) == 0) {
      if (MPI_Wtime() - t0 > 120.0) die("point-to-point wait timed out after 120 s (peer dead or unmatched send/recv)");
      usleep(50);
    }
  }
  for (int i = 0; i < n; ++i) {
    if (ids[i] == MPI_REQUEST_NULL) continue;
    Req& r = g.reqs[ids[i]];
    if (r.device && !r.is_send && (size_t)ids[i] < g_devdst.size() && g_devdst[ids[i]].p)
      cudaMemcpy(g_devdst[ids[i]].p, r.bounce.data(), r.off, cudaMemcpyHostToDevice);
    if (sts) sts[i] = r.st;
    r.active = false; r.bounce.clear(); r.bounce.shrink_to_fit();
    ids[i] = MPI_REQUEST_NULL;
  }
}

// run a collective on host or device buffers
template <typename HostFn, typename DevFn> int coll(const void* a, const void* b, HostFn hf, DevFn df) {
  const bool dev = is_device_ptr(a) || is_device_ptr(b);
  if (dev) { sy_comm* c = dev_comm(); SYCHECK(df(c, (sy_stream_t)g.stream)); dev_sync(); }
  else SYCHECK(hf(g.host));
  return MPI_SUCCESS;
}

}  // namespace

extern "C" {

int MPI_Init(int*, char***) {
  if (g.inited) return MPI_SUCCESS;
  static const char* rn[] = {"SHIPYARD_RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "RANK", nullptr};
  static const char* wn[] = {"SHIPYARD_WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "WORLD_SIZE", nullptr};
  g.rank = env_int(rn, 0); g.world = env_int(wn, 1);
  const char* s = getenv("SHIPYARD_COLL_SESSION");
  g.session = std::string("mpi-") + (s && *s ? s : (getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "solo")) +
              (g.world == 1 ? "-" + std::to_string(getpid()) : "");
  size_t heap = (64ul << 20) + (size_t)g.world * BOX_BYTES + (160ul << 20);
  SYCHECK(sy_comm_init(&g.host, g.rank, g.world, g.session.c_str(), -1, heap, SY_TRANSPORT_STUB));
  g.boxes = (char*)sy_sym_alloc(g.host, (size_t)g.world * BOX_BYTES);
  if (!g.boxes) die("cannot allocate point-to-point mailboxes: %s", sy_last_error());
  g.boxes_off = g.boxes - (char*)sy_heap_base(g.host, g.rank);
  memset(g.boxes, 0, (size_t)g.world * BOX_BYTES);
  g.sent.assign(g.world, 0); g.rcvd.assign(g.world, 0);
  g.reqs.resize(1);
  SYCHECK(sy_barrier(g.host, nullptr));
  g.inited = true;
  return MPI_SUCCESS;
}
int MPI_Init_thread(int* a, char*** b, int required, int* provided) { if (provided) *provided = required < MPI_THREAD_SERIALIZED ? required : MPI_THREAD_SERIALIZED; return MPI_Init(a, b); }
int MPI_Initialized(int* f) { *f = g.inited ? 1 : 0; return MPI_SUCCESS; }
int MPI_Finalized(int* f) { *f = g.finalized ? 1 : 0; return MPI_SUCCESS; }
int MPI_Finalize(void) {
  if (!g.inited || g.finalized) return MPI_SUCCESS;
  sy_barrier(g.host, nullptr);
  if (g.dev) { cudaStreamSynchronize(g.stream); sy_comm_destroy(g.dev); g.dev = nullptr; }
  sy_comm_destroy(g.host); g.host = nullptr;
  g.finalized = true;
  return MPI_SUCCESS;
}
int MPI_Abort(MPI_Comm, int code) { fprintf(stderr, "[shipyard-mpi rank %d] MPI_Abort(%d)\n", g.rank, code); _exit(code ? code : 1); }
int MPI_Comm_rank(MPI_Comm c, int* r) { *r = c == MPI_COMM_SELF ? 0 : g.rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int* s) { *s = c == MPI_COMM_SELF ? 1 : g.world; return MPI_SUCCESS; }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n) { *n = c; return MPI_SUCCESS; }
int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return MPI_SUCCESS; }
int MPI_Get_processor_name(char* name, int* len) { snprintf(name, MPI_MAX_PROCESSOR_NAME, "b200-box-rank%d", g.rank); *len = (int)strlen(name); return MPI_SUCCESS; }
int MPI_Error_string(int code, char* s, int* len) { snprintf(s, MPI_MAX_ERROR_STRING, "shipyard-mpi error %d", code); *len = (int)strlen(s); return MPI_SUCCESS; }
int MPI_Type_size(MPI_Datatype dt, int* size) { *size = (int)dt_size(dt); return *size ? MPI_SUCCESS : MPI_ERR_OTHER; }
int MPI_Get_count(const MPI_Status* st, MPI_Datatype dt, int* count) { size_t s = dt_size(dt); *count = s ? (int)(st->_count / s) : 0; return MPI_SUCCESS; }
double MPI_Wtime(void) { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + tv.tv_usec * 1e-6; }
double MPI_Wtick(void) { return 1e-6; }

int MPI_Barrier(MPI_Comm c) {
  if (c == MPI_COMM_SELF || g.world == 1) return MPI_SUCCESS;
  SYCHECK(sy_barrier(g.host, nullptr));
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void* sb, void* rb, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm) {
  const void* in = sb == MPI_IN_PLACE ? rb : sb;
  int sdt = sy_dt(dt), sop = sy_opof(op);
  if (sdt < 0 || sop < 0) die("MPI_Allreduce: unsupported datatype/op (%d, %d)", dt, op);
  return coll(in, rb,
              [&](sy_comm* c) { return sy_allreduce(c, in, rb, (size_t)count, sdt, sdt, 1.0f, sop, SY_ALGO_AUTO, nullptr); },
              [&](sy_comm* c, sy_stream_t s) { return sy_allreduce(c, in, rb, (size_t)count, sdt, sdt, 1.0f, sop, SY_ALGO_AUTO, s); });
}
int MPI_Reduce(const void* sb, void* rb, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm) {
  const void* in = sb == MPI_IN_PLACE ? rb : sb;
  int sdt = sy_dt(dt), sop = sy_opof(op);
  if (sdt < 0 || sop < 0) die("MPI_Reduce: unsupported datatype/op");
  return coll(in, rb, [&](sy_comm* c) { return sy_reduce(c, in, rb, (size_t)count, sdt, sop, root, nullptr); },
              [&](sy_comm* c, sy_stream_t s) { return sy_reduce(c, in, rb, (size_t)count, sdt, sop, root, s); });
}
int MPI_Reduce_scatter_block(const void* sb, void* rb, int rc, MPI_Datatype dt, MPI_Op op, MPI_Comm) {
  int sdt = sy_dt(dt), sop = sy_opof(op);
  if (sdt < 0 || sop < 0 || sb == MPI_IN_PLACE) die("MPI_Reduce_scatter_block: unsupported arguments");
  return coll(sb, rb, [&](sy_comm* c) { return sy_reduce_scatter(c, sb, rb, (size_t)rc, sdt, sdt, 1.0f, sop, nullptr); },
              [&](sy_comm* c, sy_stream_t s) { return sy_reduce_scatter(c, sb, rb, (size_t)rc, sdt, sdt, 1.0f, sop, s); });
}
int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm) {
  size_t bytes = (size_t)count * dt_size(dt);
  return coll(buf, buf, [&](sy_comm* c) { return sy_broadcast(c, buf, buf, bytes, SY_U8, root, nullptr); },
              [&](sy_comm* c, sy_stream_t s) { return sy_broadcast(c, buf, buf, bytes, SY_U8, root, s); });
}
int MPI_Allgather(const void* sb, int sc, MPI_Datatype sdt, void* rb, int, MPI_Datatype, MPI_Comm) {
  size_t bytes = (size_t)sc * dt_size(sdt);
  std::vector<char> tmp;
  const void* in = sb;
  if (sb == MPI_IN_PLACE) {
    if (is_device_ptr(rb)) die("MPI_Allgather: MPI_IN_PLACE on device buffers is not supported");
    tmp.assign((char*)rb + (size_t)g.rank * bytes, (char*)rb + (size_t)(g.rank + 1) * bytes); in = tmp.data();
  }
  return coll(in, rb, [&](sy_comm* c) { return sy_allgather(c, in, rb, bytes, SY_U8, nullptr); },
              [&](sy_comm* c, sy_stream_t s) { return sy_allgather(c, in, rb, bytes, SY_U8, s); });
}
int MPI_Alltoall(const void* sb, int sc, MPI_Datatype sdt, void* rb, int, MPI_Datatype, MPI_Comm) {
  size_t bytes = (size_t)sc * dt_size(sdt);
  if (sb == MPI_IN_PLACE) die("MPI_Alltoall: MPI_IN_PLACE is not supported");
  return coll(sb, rb, [&](sy_comm* c) { return sy_alltoall(c, sb, rb, bytes, SY_U8, nullptr); },
              [&](sy_comm* c, sy_stream_t s) { return sy_alltoall(c, sb, rb, bytes, SY_U8, s); });
}
int MPI_Gather(const void* sb, int sc, MPI_Datatype sdt, void* rb, int, MPI_Datatype, int root, MPI_Comm) {
  size_t bytes = (size_t)sc * dt_size(sdt);
  return coll(sb, rb, [&](sy_comm* c) { return sy_gather(c, sb, rb, bytes, SY_U8, root, nullptr); },
              [&](sy_comm* c, sy_stream_t s) { return sy_gather(c, sb, rb, bytes, SY_U8, root, s); });
}
int MPI_Scatter(const void* sb, int, MPI_Datatype, void* rb, int rc, MPI_Datatype rdt, int root, MPI_Comm) {
  size_t bytes = (size_t)rc * dt_size(rdt);
  return coll(sb, rb, [&](sy_comm* c) { return sy_scatter(c, sb, rb, bytes, SY_U8, root, nullptr); },
              [&](sy_comm* c, sy_stream_t s) { return sy_scatter(c, sb, rb, bytes, SY_U8, root, s); });
}

// v-variants: exchange through the uniform primitives with max-count padding (host buffers)
static void need_host(const void* a, const void* b, const char* fn) {
  if (is_device_ptr(a) || is_device_ptr(b)) die("%s: device buffers are not supported by the v-variants; use the uniform collective", fn);
}
int MPI_Allgatherv(const void* sb, int sc, MPI_Datatype sdt, void* rb, const int* rcs, const int* displs, MPI_Datatype rdt, MPI_Comm) {
  need_host(sb, rb, "MPI_Allgatherv");
  size_t es = dt_size(rdt); int mx = 0;
  for (int r = 0; r < g.world; ++r) mx = rcs[r] > mx ? rcs[r] : mx;
  std::vector<char> in((size_t)mx * es + 1), out((size_t)mx * es * g.world + 1);
  const void* src = sb == MPI_IN_PLACE ? (char*)rb + (size_t)displs[g.rank] * es : sb;
  memcpy(in.data(), src, (size_t)(sb == MPI_IN_PLACE ? rcs[g.rank] : sc) * dt_size(sb == MPI_IN_PLACE ? rdt : sdt));
  SYCHECK(sy_allgather(g.host, in.data(), out.data(), (size_t)mx * es, SY_U8, nullptr));
  for (int r = 0; r < g.world; ++r) memcpy((char*)rb + (size_t)displs[r] * es, out.data() + (size_t)r * mx * es, (size_t)rcs[r] * es);
  return MPI_SUCCESS;
}
int MPI_Gatherv(const void* sb, int sc, MPI_Datatype sdt, void* rb, const int* rcs, const int* displs, MPI_Datatype rdt, int root, MPI_Comm) {
  need_host(sb, rb, "MPI_Gatherv");
  // every rank learns the maximum contribution, then a padded gather
  int mine = sc, mx = 0;
  SYCHECK(sy_allreduce(g.host, &mine, &mx, 1, SY_I32, SY_I32, 1.0f, SY_MAX, SY_ALGO_AUTO, nullptr));
  size_t es = dt_size(sdt);
  std::vector<char> in((size_t)mx * es + 1), out((size_t)mx * es * g.world + 1);
  memcpy(in.data(), sb, (size_t)sc * es);
  SYCHECK(sy_gather(g.host, in.data(), out.data(), (size_t)mx * es, SY_U8, root, nullptr));
  if (g.rank == root)
    for (int r = 0; r < g.world; ++r) memcpy((char*)rb + (size_t)displs[r] * dt_size(rdt), out.data() + (size_t)r * mx * es, (size_t)rcs[r] * dt_size(rdt));
  return MPI_SUCCESS;
}
int MPI_Alltoallv(const void* sb, const int* scs, const int* sdis, MPI_Datatype sdt, void* rb, const int* rcs, const int* rdis, MPI_Datatype rdt, MPI_Comm) {
  need_host(sb, rb, "MPI_Alltoallv");
  int mine = 0, mx = 0;
  for (int r = 0; r < g.world; ++r) mine = scs[r] > mine ? scs[r] : mine;
  SYCHECK(sy_allreduce(g.host, &mine, &mx, 1, SY_I32, SY_I32, 1.0f, SY_MAX, SY_ALGO_AUTO, nullptr));
  size_t es = dt_size(sdt), blk = (size_t)mx * es;
  std::vector<char> in(blk * g.world + 1), out(blk * g.world + 1);
  for (int r = 0; r < g.world; ++r) memcpy(in.data() + (size_t)r * blk, (const char*)sb + (size_t)sdis[r] * es, (size_t)scs[r] * es);
  SYCHECK(sy_alltoall(g.host, in.data(), out.data(), blk, SY_U8, nullptr));
  for (int r = 0; r < g.world; ++r) memcpy((char*)rb + (size_t)rdis[r] * dt_size(rdt), out.data() + (size_t)r * blk, (size_t)rcs[r] * dt_size(rdt));
  return MPI_SUCCESS;
}

// ---- point to point --------------------------------------------------------------------------------
int MPI_Isend(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm, MPI_Request* req) {
  if (dest == MPI_PROC_NULL) { *req = MPI_REQUEST_NULL; return MPI_SUCCESS; }
  if (dest < 0 || dest >= g.world) die("MPI_Isend: bad destination %d", dest);
  *req = new_req(true, (void*)buf, (size_t)count * dt_size(dt), dest, tag);
  progress(g.reqs[*req]);
  return MPI_SUCCESS;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm, MPI_Request* req) {
  if (src == MPI_PROC_NULL) { *req = MPI_REQUEST_NULL; return MPI_SUCCESS; }
  if (src == MPI_ANY_SOURCE) die("MPI_ANY_SOURCE is not supported (per-peer in-order matching)");
  if (src < 0 || src >= g.world) die("MPI_Irecv: bad source %d", src);
  int id = new_req(false, buf, (size_t)count * dt_size(dt), src, tag);
  if (g.reqs[id].device) { if (g_devdst.size() <= (size_t)id) g_devdst.resize(id + 1); g_devdst[id].p = buf; }
  else if (g_devdst.size() > (size_t)id) g_devdst[id].p = nullptr;
  *req = id;
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request* req, MPI_Status* st) { if (*req != MPI_REQUEST_NULL) wait_all(1, req, st); return MPI_SUCCESS; }
int MPI_Waitall(int n, MPI_Request reqs[], MPI_Status sts[]) { wait_all(n, reqs, sts); return MPI_SUCCESS; }
int MPI_Send(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm c) {
  MPI_Request r; MPI_Isend(buf, count, dt, dest, tag, c, &r); return MPI_Wait(&r, MPI_STATUS_IGNORE);
}
int MPI_Recv(void* buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Status* st) {
  MPI_Request r; MPI_Irecv(buf, count, dt, src, tag, c, &r); return MPI_Wait(&r, st);
}
int MPI_Sendrecv(const void* sb, int sc, MPI_Datatype sdt, int dest, int stag, void* rb, int rc, MPI_Datatype rdt, int src, int rtag,
                 MPI_Comm c, MPI_Status* st) {
  MPI_Request r[2]; MPI_Status s[2];
  MPI_Irecv(rb, rc, rdt, src, rtag, c, &r[0]);
  MPI_Isend(sb, sc, sdt, dest, stag, c, &r[1]);
  wait_all(2, r, s);
  if (st) *st = s[0];
  return MPI_SUCCESS;
}

int MPIX_Query_shipyard_transport(char* name, int len) {
  const char* t = "stub";
  if (g.dev) t = sy_comm_transport(g.dev) == SY_TRANSPORT_NVLS ? "nvls" : "p2p";
  snprintf(name, len, "%s", t);
  return MPI_SUCCESS;
}
void* MPIX_Sym_alloc(size_t bytes) { return sy_sym_alloc(dev_comm(), bytes); }
// Extensions for device-timed benchmarking (bench/mpibench.cpp --compare): the stream the device-buffer collectives run on, an
// asynchronous mode in which they return right after the kernel is enqueued, and the explicit synchronisation (with the watchdog check).
void* MPIX_Device_stream(void) { dev_comm(); return (void*)g.stream; }
int MPIX_Set_device_async(int on) { g_async = on != 0; return MPI_SUCCESS; }
int MPIX_Device_sync(void) {
  if (!g.dev) return MPI_SUCCESS;
  cudaError_t e = cudaStreamSynchronize(g.stream);
  if (e != cudaSuccess) die("stream sync: %s", cudaGetErrorString(e));
  if (sy_comm_status(g.dev) != 0) die("collective watchdog fired: a peer rank did not arrive");
  return MPI_SUCCESS;
}
void* MPIX_Shipyard_comm(void) { return (void*)dev_comm(); }

}  // extern "C"
