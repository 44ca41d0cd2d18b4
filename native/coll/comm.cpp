// Public API: communicator lifecycle, symmetric-heap allocator, algorithm selection
// and staging for non-symmetric user buffers.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "internal.h"

static size_t env_sz(const char* name, size_t dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  char* end = nullptr; double x = strtod(v, &end);
  if (end && (*end == 'k' || *end == 'K')) x *= 1024.0;
  else if (end && (*end == 'm' || *end == 'M')) x *= 1024.0 * 1024.0;
  else if (end && (*end == 'g' || *end == 'G')) x *= 1024.0 * 1024.0 * 1024.0;
  return (size_t)x;
}

static bool is_stub(const sy_comm* c) { return c->transport == SY_TRANSPORT_STUB; }

static bool sym_off(sy_comm* c, const void* p, size_t* off) {
  const char* b = c->dev.heap[c->rank];
  if (p && (const char*)p >= b && (const char*)p < b + c->heap_bytes) { *off = (const char*)p - b; return true; }
  return false;
}

extern "C" int sy_comm_init(sy_comm** out, int rank, int world, const char* session, int device,
                            size_t heap_bytes, int transport) {
  if (!out || rank < 0 || world < 1 || world > SY_MAXR || rank >= world || !session) {
    sy_set_error("sy_comm_init: bad arguments (rank=%d world=%d; max world is %d)", rank, world, SY_MAXR);
    return SY_ERR_ARG;
  }
  const char* tenv = getenv("SHIPYARD_COLL_TRANSPORT");
  if (tenv && transport == SY_TRANSPORT_AUTO) {
    if (!strcmp(tenv, "stub")) transport = SY_TRANSPORT_STUB;
    else if (!strcmp(tenv, "p2p")) transport = SY_TRANSPORT_P2P;
    else if (!strcmp(tenv, "nvls")) transport = SY_TRANSPORT_NVLS;
  }
  if (device < 0) transport = SY_TRANSPORT_STUB;
  sy_comm* c = new sy_comm();
  c->rank = rank; c->world = world; c->device = device; c->session = session;
  c->dev.rank = rank; c->dev.world = world;
  c->timeout_ms = (long)env_sz("SHIPYARD_COLL_TIMEOUT_MS", 20000);
  c->max_blocks = (long)env_sz("SHIPYARD_COLL_MAX_BLOCKS", 128);
  c->threads = (long)env_sz("SHIPYARD_COLL_THREADS", 512);
  c->ll_max_bytes = (long)env_sz("SHIPYARD_COLL_LL_MAX", 16384);
  c->oneshot_max_bytes = (long)env_sz("SHIPYARD_COLL_ONESHOT_MAX", 256 << 10);
  c->mailbox_max_bytes = (long)env_sz("SHIPYARD_COLL_MAILBOX_MAX", 1 << 20);
  c->lm_max_bytes = (long)env_sz("SHIPYARD_COLL_LM_MAX", SY_LM_MAX_PAYLOAD);
  if (c->lm_max_bytes > (long)SY_LM_MAX_PAYLOAD) c->lm_max_bytes = (long)SY_LM_MAX_PAYLOAD;
  c->nvls_copy = (long)env_sz("SHIPYARD_COLL_NVLS_COPY", 1);
  c->nvls_min_world = (long)env_sz("SHIPYARD_COLL_NVLS_MIN_WORLD", 4);
  if (heap_bytes == 0) heap_bytes = env_sz("SHIPYARD_COLL_HEAP", transport == SY_TRANSPORT_STUB ? (256ul << 20) : (1ul << 30));
  size_t min_heap = SY_USER_OFF + (16ul << 20);
  if (heap_bytes < min_heap) heap_bytes = min_heap;
  c->heap_bytes = (heap_bytes + (2ul << 20) - 1) / (2ul << 20) * (2ul << 20);
  c->hub = hub_create(rank, world, c->session, (int)env_sz("SHIPYARD_COLL_BOOT_TIMEOUT_MS", 120000));
  if (!c->hub) { delete c; return SY_ERR_SYS; }
  int rc;
  if (transport == SY_TRANSPORT_STUB) {
    rc = stub_init(c);
    c->status_host = (uint32_t*)calloc(1, sizeof(uint32_t));   // the stub has no mapped status word: a private one
  } else {
    rc = gpu_init(c, transport);
  }
  if (rc != SY_OK) {
    if (is_stub(c) || transport == SY_TRANSPORT_STUB) stub_destroy(c); else gpu_destroy(c);
    hub_destroy(c->hub); delete c; return rc;
  }
  // staging region for non-symmetric user buffers: two halves (in / out)
  size_t stage = env_sz("SHIPYARD_COLL_STAGE", is_stub(c) ? (64ul << 20) : (128ul << 20));
  size_t avail = c->heap_bytes - SY_USER_OFF;
  if (stage > avail / 2) stage = avail / 2;
  stage = stage / (4ul << 20) * (4ul << 20);
  c->stage_bytes = stage;
  c->stage_off = c->heap_bytes - stage;   // carved from the top; the bump allocator grows from below
  *out = c;
  return SY_OK;
}

extern "C" int sy_comm_destroy(sy_comm* c) {
  if (!c) return SY_OK;
  if (is_stub(c)) { if (c->world > 1) hub_barrier(c->hub); stub_destroy(c); free(c->status_host); }
  else gpu_destroy(c);
  hub_destroy(c->hub);
  delete c;
  return SY_OK;
}
extern "C" int sy_comm_rank(const sy_comm* c) { return c->rank; }
extern "C" int sy_comm_world(const sy_comm* c) { return c->world; }
extern "C" int sy_comm_transport(const sy_comm* c) { return c->transport; }
extern "C" int sy_comm_has_multicast(const sy_comm* c) { return c->has_mc ? 1 : 0; }
extern "C" int sy_comm_status(sy_comm* c) { return c->status_host ? (int)*c->status_host : 0; }
extern "C" uint64_t sy_comm_launch_count(const sy_comm* c) { return c->launches; }
extern "C" size_t sy_heap_bytes(const sy_comm* c) { return c->heap_bytes; }
// device-side view (heap pointers, multicast VA, epoch counters) for kernels living in other libraries
extern "C" size_t sy_comm_device_view(sy_comm* c, void* out, size_t cap) {
  CommDev d = c->dev;
  d.timeout_ns = (unsigned long long)c->timeout_ms * 1000000ull;
  if (out && cap >= sizeof d) memcpy(out, &d, sizeof d);
  return sizeof d;
}
extern "C" void sy_comm_count_launch(sy_comm* c) { c->launches++; }
extern "C" void* sy_heap_base(sy_comm* c, int peer) { return peer >= 0 && peer < c->world ? c->dev.heap[peer] : nullptr; }
extern "C" void* sy_mc_base(sy_comm* c) { return c->dev.mc; }
extern "C" int sy_is_symmetric(sy_comm* c, const void* p) { size_t o; return sym_off(c, p, &o) ? 1 : 0; }

extern "C" void* sy_sym_alloc(sy_comm* c, size_t bytes) {
  size_t al = 256;
  size_t off = (c->bump + al - 1) / al * al;
  size_t need = (bytes + al - 1) / al * al;
  if (off + need > c->stage_off) {
    sy_set_error("sym_alloc: heap exhausted (want %zu, free %zu); raise SHIPYARD_COLL_HEAP", need,
                 c->stage_off > off ? c->stage_off - off : 0);
    return nullptr;
  }
  c->bump = off + need;
  return c->dev.heap[c->rank] + off;
}
extern "C" int sy_sym_reset(sy_comm* c) { c->bump = SY_USER_OFF; return SY_OK; }

extern "C" int sy_set_tuning(sy_comm* c, const char* k, long v) {
  if (!strcmp(k, "max_blocks")) c->max_blocks = v < 1 ? 1 : (v > SY_MAX_BLOCKS ? SY_MAX_BLOCKS : v);
  else if (!strcmp(k, "threads")) c->threads = v < 32 ? 32 : (v > 512 ? 512 : v / 32 * 32);
  else if (!strcmp(k, "ll_max_bytes")) c->ll_max_bytes = v > (long)SY_LL_MAX_PAYLOAD ? (long)SY_LL_MAX_PAYLOAD : v;
  else if (!strcmp(k, "oneshot_max_bytes")) c->oneshot_max_bytes = v > (long)SY_OS_SLOT ? (long)SY_OS_SLOT : v;
  else if (!strcmp(k, "mailbox_max_bytes")) c->mailbox_max_bytes = v > (long)SY_OS_SLOT ? (long)SY_OS_SLOT : v;
  else if (!strcmp(k, "nvls_min_bytes")) c->nvls_min_bytes = v;
  else if (!strcmp(k, "timeout_ms")) c->timeout_ms = v;
  else if (!strcmp(k, "nvls_copy")) c->nvls_copy = v;
  else if (!strcmp(k, "bcast_sag_min_bytes")) c->bcast_sag_min_bytes = v;
  else if (!strcmp(k, "ag_p2p_min_bytes")) c->ag_p2p_min_bytes = v;
  else if (!strcmp(k, "lm_max_bytes")) c->lm_max_bytes = v > (long)SY_LM_MAX_PAYLOAD ? (long)SY_LM_MAX_PAYLOAD : v;
  else if (!strcmp(k, "nvls_min_world")) c->nvls_min_world = v;
  else return SY_ERR_ARG;
  return SY_OK;
}
extern "C" long sy_get_tuning(sy_comm* c, const char* k) {
  if (!strcmp(k, "max_blocks")) return c->max_blocks;
  if (!strcmp(k, "threads")) return c->threads;
  if (!strcmp(k, "ll_max_bytes")) return c->ll_max_bytes;
  if (!strcmp(k, "oneshot_max_bytes")) return c->oneshot_max_bytes;
  if (!strcmp(k, "mailbox_max_bytes")) return c->mailbox_max_bytes;
  if (!strcmp(k, "nvls_min_bytes")) return c->nvls_min_bytes;
  if (!strcmp(k, "timeout_ms")) return c->timeout_ms;
  if (!strcmp(k, "nvls_copy")) return c->nvls_copy;
  if (!strcmp(k, "bcast_sag_min_bytes")) return c->bcast_sag_min_bytes;
  if (!strcmp(k, "ag_p2p_min_bytes")) return c->ag_p2p_min_bytes;
  if (!strcmp(k, "lm_max_bytes")) return c->lm_max_bytes;
  if (!strcmp(k, "nvls_min_world")) return c->nvls_min_world;
  return -1;
}

extern "C" size_t sy_shard_begin(const sy_comm* c, size_t count, int rank) {
  size_t units = count / 8, base = units / c->world, rem = units % c->world;
  return ((size_t)rank * base + ((size_t)rank < rem ? (size_t)rank : rem)) * 8;
}
extern "C" size_t sy_shard_count(const sy_comm* c, size_t count, int rank) {
  size_t units = count / 8, base = units / c->world, rem = units % c->world;
  return (base + ((size_t)rank < rem ? 1 : 0)) * 8;
}

#define CUDA_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { sy_set_error(#x ": %s", cudaGetErrorString(_e)); return SY_ERR_CUDA; } } while (0)

static char* stage_half(sy_comm* c, int half) { return c->dev.heap[c->rank] + c->stage_off + (size_t)half * (c->stage_bytes / 2); }
static size_t stage_half_off(sy_comm* c, int half) { return c->stage_off + (size_t)half * (c->stage_bytes / 2); }

static bool mailbox_ok(const sy_comm* c, const void* in, const void* out, size_t bytes);
static bool lm_ok(const sy_comm* c, const void* in, const void* out, size_t bytes);
static bool nvls_dtype_ok(int a, int b) {
  if (a == b) return a == SY_F32 || a == SY_BF16 || a == SY_F16;
  return (a == SY_BF16 && b == SY_F32) || (a == SY_F32 && b == SY_BF16);
}
static bool pair_ok(int a, int b) {
  if (a == b) return a == SY_F32 || a == SY_BF16 || a == SY_F16 || a == SY_F64 || a == SY_I32 || a == SY_I64;
  return (a == SY_BF16 && b == SY_F32) || (a == SY_F32 && b == SY_BF16);
}

extern "C" int sy_allreduce(sy_comm* c, const void* in, void* out, size_t count, int dt_in, int dt_out,
                            float scale, int op, int algo, sy_stream_t stream) {
  if (!pair_ok(dt_in, dt_out)) { sy_set_error("allreduce: unsupported dtype pair %d->%d", dt_in, dt_out); return SY_ERR_UNSUPPORTED; }
  if (count == 0) return SY_OK;
  if (is_stub(c)) return stub_allreduce(c, in, out, count, dt_in, dt_out, scale, op);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t si = sy_dtype_size(dt_in), so = sy_dtype_size(dt_out);
  if (c->world == 1) return k_local_cast(c, in, out, count, dt_in, dt_out, scale, stream);
  size_t in_off = 0, out_off = 0;
  const bool in_sym = sym_off(c, in, &in_off), out_sym = sym_off(c, out, &out_off);
  const size_t bytes = count * si;
  const bool auto_algo = algo == SY_ALGO_AUTO;
  if (algo == SY_ALGO_AUTO) {
    if (bytes <= (size_t)c->ll_max_bytes) algo = SY_ALGO_LL;
    else if (bytes <= (size_t)c->oneshot_max_bytes) algo = SY_ALGO_ONESHOT;
    else algo = (c->has_mc && c->world >= c->nvls_min_world && op == SY_SUM && nvls_dtype_ok(dt_in, dt_out)) ? SY_ALGO_TWOSHOT_NVLS : SY_ALGO_TWOSHOT_P2P;
  }
  // 16 KB .. 256 KB sums: the multi-block LL kernel instead of the one-shot mailboxes (no fence + flag round trip)
  if (algo == SY_ALGO_ONESHOT && auto_algo && op == SY_SUM && dt_in == dt_out && (dt_in == SY_F32 || dt_in == SY_BF16) && lm_ok(c, in, out, bytes))
    return k_lm(c, in, out, bytes, 4, dt_in, stream, scale);
  if (algo == SY_ALGO_TWOSHOT_NVLS && !(c->has_mc && op == SY_SUM && nvls_dtype_ok(dt_in, dt_out))) {
    sy_set_error("allreduce: NVLS path unavailable for this call"); return SY_ERR_UNSUPPORTED;
  }
  if (algo == SY_ALGO_LL) {
    if (bytes > SY_LL_MAX_PAYLOAD) { sy_set_error("allreduce: LL limited to %lu bytes", SY_LL_MAX_PAYLOAD); return SY_ERR_ARG; }
    return k_allreduce(c, in, out, 0, 0, false, false, count, dt_in, dt_out, scale, op, SY_ALGO_LL, stream);
  }
  if (algo == SY_ALGO_ONESHOT) {
    // chunk through the 1 MB mailbox slot (a unit never straddles a chunk: 64K-element multiples)
    const size_t maxe = SY_OS_SLOT / (si > so ? si : so);
    for (size_t b = 0; b < count; b += maxe) {
      size_t n = count - b < maxe ? count - b : maxe;
      int rc = k_allreduce(c, (const char*)in + b * si, (char*)out + b * so, 0, 0, false, false, n, dt_in, dt_out,
                           scale, op, SY_ALGO_ONESHOT, stream);
      if (rc) return rc;
    }
    return SY_OK;
  }
  // two-shot: symmetric operands are used in place; others are staged
  if (in_sym && out_sym)
    return k_allreduce(c, in, out, in_off, out_off, true, true, count, dt_in, dt_out, scale, op, algo, stream);
  const size_t half = c->stage_bytes / 2;
  const size_t maxe = half / (si > so ? si : so) / 64 * 64;
  if (maxe == 0) { sy_set_error("allreduce: no staging space"); return SY_ERR_NOMEM; }
  for (size_t b = 0; b < count; b += maxe) {
    size_t n = count - b < maxe ? count - b : maxe;
    size_t io = in_sym ? in_off + b * si : stage_half_off(c, 0);
    size_t oo = out_sym ? out_off + b * so : stage_half_off(c, 1);
    if (!in_sym) CUDA_TRY(cudaMemcpyAsync(stage_half(c, 0), (const char*)in + b * si, n * si, cudaMemcpyDeviceToDevice, s));
    int rc = k_allreduce(c, nullptr, nullptr, io, oo, true, true, n, dt_in, dt_out, scale, op, algo, stream);
    if (rc) return rc;
    if (!out_sym) CUDA_TRY(cudaMemcpyAsync((char*)out + b * so, stage_half(c, 1), n * so, cudaMemcpyDeviceToDevice, s));
  }
  return SY_OK;
}

extern "C" int sy_reduce_scatter(sy_comm* c, const void* in, void* out, size_t count, int dt_in, int dt_out,
                                 float scale, int op, sy_stream_t stream) {
  if (!pair_ok(dt_in, dt_out)) return SY_ERR_UNSUPPORTED;
  if (count == 0) return SY_OK;
  if (is_stub(c)) return stub_reduce_scatter(c, in, out, count, dt_in, dt_out, scale, op);
  if (c->world == 1) return k_local_cast(c, in, out, count, dt_in, dt_out, scale, stream);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t si = sy_dtype_size(dt_in);
  // small / medium shards: one-shot through the mailboxes (push my block p into rank p's slot, one flag per block, the receiver sums
  // its W slots in fp32 in rank order): no barriers, any device pointers.  NCCL was 1.4-1.6x faster than the barrier-based kernel
  // below 1 MB (profiles/round2_multi_gpu.md, N = 4)
  if (op == SY_SUM && dt_in == dt_out && (dt_in == SY_F32 || dt_in == SY_BF16) && lm_ok(c, in, out, count * si))
    return k_lm(c, in, out, count * si, 3, dt_in, stream, scale);
  if (op == SY_SUM && dt_in == dt_out && (dt_in == SY_F32 || dt_in == SY_BF16) && mailbox_ok(c, in, out, count * si))
    return k_mailbox(c, in, out, count * si, 3, dt_in, stream, scale);
  size_t in_off;
  const size_t total = count * si * c->world;
  if (!sym_off(c, in, &in_off)) {
    if (total > c->stage_bytes / 2) {
      // chunk along the per-rank count: stage [world x nc] columns of the input per round
      const size_t so_ = sy_dtype_size(dt_out);
      const size_t nc_max = (c->stage_bytes / 2 / c->world / si) / 64 * 64;
      if (nc_max == 0) { sy_set_error("reduce_scatter: no staging space"); return SY_ERR_NOMEM; }
      for (size_t b = 0; b < count; b += nc_max) {
        const size_t nc = count - b < nc_max ? count - b : nc_max;
        CUDA_TRY(cudaMemcpy2DAsync(stage_half(c, 0), nc * si, (const char*)in + b * si, count * si, nc * si, c->world, cudaMemcpyDeviceToDevice, s));
        int rc2 = sy_reduce_scatter(c, stage_half(c, 0), (char*)out + b * so_, nc, dt_in, dt_out, scale, op, stream);
        if (rc2) return rc2;
      }
      return SY_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(stage_half(c, 0), in, total, cudaMemcpyDeviceToDevice, s));
    in_off = stage_half_off(c, 0);
  }
  bool nvls = c->has_mc && c->world >= c->nvls_min_world && op == SY_SUM && nvls_dtype_ok(dt_in, dt_out) && total >= (size_t)c->nvls_min_bytes;
  int rc = k_reduce_scatter(c, in_off, out, count, dt_in, dt_out, scale, op, nvls, stream);
  if (rc == SY_ERR_UNSUPPORTED) {
    // unaligned shard: all-reduce the whole thing into staging, then copy my shard out
    if (total > c->stage_bytes / 2) { sy_set_error("reduce_scatter: unaligned count too large for staging"); return SY_ERR_NOMEM; }
    rc = k_allreduce(c, nullptr, nullptr, in_off, stage_half_off(c, 1), true, true, count * c->world, dt_in, dt_out,
                     scale, op, SY_ALGO_TWOSHOT_P2P, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, stage_half(c, 1) + (size_t)c->rank * count * sy_dtype_size(dt_out),
                             count * sy_dtype_size(dt_out), cudaMemcpyDeviceToDevice, s));
  }
  return rc;
}

// helper for collectives whose OUTPUT must be symmetric (peers write into it)
struct OutStage { bool staged; size_t off; };
static int out_target(sy_comm* c, void* out, size_t bytes, OutStage* t) {
  if (sym_off(c, out, &t->off)) { t->staged = false; return SY_OK; }
  if (bytes > c->stage_bytes / 2) { sy_set_error("output larger than staging (%zu > %zu); use a symmetric buffer", bytes, c->stage_bytes / 2); return SY_ERR_NOMEM; }
  t->staged = true; t->off = stage_half_off(c, 1);
  return SY_OK;
}

// small / medium byte movers go through the one-shot mailboxes: any device pointers, no staging copy, no barriers
static bool mailbox_ok(const sy_comm* c, const void* in, const void* out, size_t bytes) {
  return bytes <= (size_t)c->mailbox_max_bytes && bytes <= SY_OS_SLOT && bytes % 16 == 0 &&
         ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0 && !getenv("SHIPYARD_COLL_NO_MAILBOX");
}

// latency-bound range: the multi-block LL kernel (payload and flag in one 8-byte atom; see k_lm_k).  `span` = bytes of in / out touched
static bool lm_ok(const sy_comm* c, const void* in, const void* out, size_t bytes) {
  return bytes <= (size_t)c->lm_max_bytes && bytes <= SY_LM_MAX_PAYLOAD && bytes % 4 == 0 && c->heap_bytes >= SY_USER_OFF &&
         ((((uintptr_t)in) | ((uintptr_t)out)) & 3) == 0;
}

extern "C" int sy_allgather(sy_comm* c, const void* in, void* out, size_t count, int dt, sy_stream_t stream) {
  if (count == 0) return SY_OK;
  if (is_stub(c)) return stub_allgather(c, in, out, count, dt);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = count * sy_dtype_size(dt);
  if (c->world == 1) { if (in != out) CUDA_TRY(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, s)); return SY_OK; }
  if (lm_ok(c, in, out, bytes)) return k_lm(c, in, out, bytes, 0, 0, stream);
  if (mailbox_ok(c, in, out, bytes)) return k_mailbox(c, in, out, bytes, 0, 0, stream);
  { size_t o_; if (!sym_off(c, out, &o_) && bytes * c->world > c->stage_bytes / 2) {
      // plain (non-symmetric) output larger than the staging half: gather it in column chunks through the staging buffer
      const size_t piece = (c->stage_bytes / 2 / c->world) & ~(size_t)255;
      for (size_t off = 0; off < bytes; off += piece) {
        const size_t n = bytes - off < piece ? bytes - off : piece;
        int rc2 = k_allgather(c, (const char*)in + off, stage_half_off(c, 1), n, SY_U8, c->has_mc && c->nvls_copy && c->world >= c->nvls_min_world, stream);
        if (rc2) return rc2;
        CUDA_TRY(cudaMemcpy2DAsync((char*)out + off, bytes, stage_half(c, 1), n, n, c->world, cudaMemcpyDeviceToDevice, s));
      }
      return SY_OK;
  } }
  OutStage t; int rc = out_target(c, out, bytes * c->world, &t); if (rc) return rc;
  // multimem.st replicates in the switch (egress bytes/rank instead of bytes x (world - 1)), which wins while the message is latency-
  // bound; from ~16 MB of output on the same sweep shows plain peer stores (the all-to-all data path: 680 GB/s bus bandwidth at
  // 1 GB) ahead of the multicast path (528-616 GB/s) — profiles/round2_multi_gpu.md, all-gather vs all-to-all rows at N = 4 / 8
  const bool ag_nvls = c->has_mc && c->nvls_copy && c->world >= c->nvls_min_world && bytes >= (size_t)(c->nvls_min_bytes / c->world) &&
                       bytes * c->world < (size_t)c->ag_p2p_min_bytes;
  rc = k_allgather(c, in, t.off, count, dt, ag_nvls, stream);
  if (rc) return rc;
  if (t.staged) CUDA_TRY(cudaMemcpyAsync(out, stage_half(c, 1), bytes * c->world, cudaMemcpyDeviceToDevice, s));
  return SY_OK;
}

extern "C" int sy_broadcast(sy_comm* c, const void* in, void* out, size_t count, int dt, int root, sy_stream_t stream) {
  if (count == 0) return SY_OK;
  if (root < 0 || root >= c->world) return SY_ERR_ARG;
  if (is_stub(c)) return stub_broadcast(c, in, out, count, dt, root);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = count * sy_dtype_size(dt);
  if (c->world == 1) { if (in != out) CUDA_TRY(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, s)); return SY_OK; }
  if (lm_ok(c, c->rank == root ? in : out, out, bytes)) return k_lm(c, c->rank == root ? in : out, out, bytes, 2, root, stream);
  if (mailbox_ok(c, c->rank == root ? in : out, out, bytes)) return k_mailbox(c, c->rank == root ? in : out, out, bytes, 2, root, stream);
  { size_t o_; if (!sym_off(c, out, &o_) && bytes > c->stage_bytes / 2) {
      const size_t piece = (c->stage_bytes / 2) & ~(size_t)255;
      for (size_t off = 0; off < bytes; off += piece) {
        const size_t n = bytes - off < piece ? bytes - off : piece;
        int rc2 = k_broadcast(c, c->rank == root ? (const char*)in + off : nullptr, stage_half_off(c, 1), n, root,
                              c->has_mc && c->nvls_copy && c->world >= c->nvls_min_world, stream);
        if (rc2) return rc2;
        CUDA_TRY(cudaMemcpyAsync((char*)out + off, stage_half(c, 1), n, cudaMemcpyDeviceToDevice, s));
      }
      return SY_OK;
  } }
  OutStage t; int rc = out_target(c, out, bytes, &t); if (rc) return rc;
  rc = k_broadcast(c, c->rank == root ? in : nullptr, t.off, bytes, root, c->has_mc && c->nvls_copy && c->world >= c->nvls_min_world && bytes >= (size_t)c->nvls_min_bytes, stream);
  if (rc) return rc;
  if (t.staged) CUDA_TRY(cudaMemcpyAsync(out, stage_half(c, 1), bytes, cudaMemcpyDeviceToDevice, s));
  return SY_OK;
}

extern "C" int sy_alltoall(sy_comm* c, const void* in, void* out, size_t count, int dt, sy_stream_t stream) {
  if (count == 0) return SY_OK;
  if (is_stub(c)) return stub_alltoall(c, in, out, count, dt);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = count * sy_dtype_size(dt);
  if (c->world == 1) { if (in != out) CUDA_TRY(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, s)); return SY_OK; }
  if (lm_ok(c, in, out, bytes)) return k_lm(c, in, out, bytes, 1, 0, stream);
  if (mailbox_ok(c, in, out, bytes)) return k_mailbox(c, in, out, bytes, 1, 0, stream);
  { size_t o_; if (!sym_off(c, out, &o_) && bytes * c->world > c->stage_bytes / 2) {
      // column chunks: stage [world x n] of the input, exchange, scatter the [world x n] result back with a strided copy
      const size_t piece = (c->stage_bytes / 2 / c->world) & ~(size_t)255;
      for (size_t off = 0; off < bytes; off += piece) {
        const size_t n = bytes - off < piece ? bytes - off : piece;
        CUDA_TRY(cudaMemcpy2DAsync(stage_half(c, 0), n, (const char*)in + off, bytes, n, c->world, cudaMemcpyDeviceToDevice, s));
        int rc2 = k_alltoall(c, stage_half(c, 0), stage_half_off(c, 1), n, stream);
        if (rc2) return rc2;
        CUDA_TRY(cudaMemcpy2DAsync((char*)out + off, bytes, stage_half(c, 1), n, n, c->world, cudaMemcpyDeviceToDevice, s));
      }
      return SY_OK;
  } }
  OutStage t; int rc = out_target(c, out, bytes * c->world, &t); if (rc) return rc;
  rc = k_alltoall(c, in, t.off, bytes, stream);
  if (rc) return rc;
  if (t.staged) CUDA_TRY(cudaMemcpyAsync(out, stage_half(c, 1), bytes * c->world, cudaMemcpyDeviceToDevice, s));
  return SY_OK;
}

extern "C" int sy_reduce(sy_comm* c, const void* in, void* out, size_t count, int dt, int op, int root, sy_stream_t stream) {
  if (count == 0) return SY_OK;
  if (root < 0 || root >= c->world) return SY_ERR_ARG;
  if (is_stub(c)) return stub_reduce(c, in, out, count, dt, op, root);
  if (c->world == 1) return k_local_cast(c, in, out, count, dt, dt, 1.0f, stream);
  cudaStream_t s = (cudaStream_t)stream;
  size_t in_off; const size_t bytes = count * sy_dtype_size(dt);
  if (!sym_off(c, in, &in_off)) {
    if (bytes > c->stage_bytes / 2) {
      // plain input larger than the staging half: reduce it in element chunks (each chunk is an independent rooted reduction)
      const size_t es = sy_dtype_size(dt), maxe = (c->stage_bytes / 2 / es) / 64 * 64;
      if (maxe == 0) { sy_set_error("reduce: no staging space"); return SY_ERR_NOMEM; }
      for (size_t b = 0; b < count; b += maxe) {
        const size_t n = count - b < maxe ? count - b : maxe;
        int rc2 = sy_reduce(c, (const char*)in + b * es, out ? (char*)out + b * es : nullptr, n, dt, op, root, stream);
        if (rc2) return rc2;
      }
      return SY_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(stage_half(c, 0), in, bytes, cudaMemcpyDeviceToDevice, s));
    in_off = stage_half_off(c, 0);
  }
  return k_reduce_rooted(c, in_off, out, count, dt, op, root, stream);
}

extern "C" int sy_gather(sy_comm* c, const void* in, void* out, size_t count, int dt, int root, sy_stream_t stream) {
  if (count == 0) return SY_OK;
  if (root < 0 || root >= c->world) return SY_ERR_ARG;
  if (is_stub(c)) return stub_gather(c, in, out, count, dt, root);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = count * sy_dtype_size(dt);
  if (c->world == 1) { if (in != out) CUDA_TRY(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, s)); return SY_OK; }
  // every rank must agree on the destination offset: non-roots cannot know the root's `out`,
  // so the rooted gather always lands in the staging half and the root copies out
  if (bytes * c->world > c->stage_bytes / 2) {
    // column chunks: gather [world x n] pieces into the staging half, the root scatters them into place with a strided copy
    const size_t piece = (c->stage_bytes / 2 / c->world) & ~(size_t)255;
    if (piece == 0) { sy_set_error("gather: no staging space"); return SY_ERR_NOMEM; }
    for (size_t off = 0; off < bytes; off += piece) {
      const size_t n = bytes - off < piece ? bytes - off : piece;
      int rc2 = k_gather(c, (const char*)in + off, stage_half_off(c, 1), n, root, stream);
      if (rc2) return rc2;
      if (c->rank == root) CUDA_TRY(cudaMemcpy2DAsync((char*)out + off, bytes, stage_half(c, 1), n, n, c->world, cudaMemcpyDeviceToDevice, s));
    }
    return SY_OK;
  }
  int rc = k_gather(c, in, stage_half_off(c, 1), bytes, root, stream);
  if (rc) return rc;
  if (c->rank == root) CUDA_TRY(cudaMemcpyAsync(out, stage_half(c, 1), bytes * c->world, cudaMemcpyDeviceToDevice, s));
  return SY_OK;
}

extern "C" int sy_scatter(sy_comm* c, const void* in, void* out, size_t count, int dt, int root, sy_stream_t stream) {
  if (count == 0) return SY_OK;
  if (root < 0 || root >= c->world) return SY_ERR_ARG;
  if (is_stub(c)) return stub_scatter(c, in, out, count, dt, root);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = count * sy_dtype_size(dt);
  if (c->world == 1) { if (in != out) CUDA_TRY(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, s)); return SY_OK; }
  if (bytes * c->world > c->stage_bytes / 2) {
    const size_t piece = (c->stage_bytes / 2 / c->world) & ~(size_t)255;
    if (piece == 0) { sy_set_error("scatter: no staging space"); return SY_ERR_NOMEM; }
    for (size_t off = 0; off < bytes; off += piece) {
      const size_t n = bytes - off < piece ? bytes - off : piece;
      if (c->rank == root) CUDA_TRY(cudaMemcpy2DAsync(stage_half(c, 0), n, (const char*)in + off, bytes, n, c->world, cudaMemcpyDeviceToDevice, s));
      int rc2 = k_scatter(c, stage_half_off(c, 0), (char*)out + off, n, root, stream);
      if (rc2) return rc2;
    }
    return SY_OK;
  }
  if (c->rank == root) CUDA_TRY(cudaMemcpyAsync(stage_half(c, 0), in, bytes * c->world, cudaMemcpyDeviceToDevice, s));
  return k_scatter(c, stage_half_off(c, 0), out, bytes, root, stream);
}

extern "C" int sy_barrier(sy_comm* c, sy_stream_t stream) {
  if (is_stub(c)) return stub_barrier(c);
  if (c->world == 1) return SY_OK;
  return k_barrier(c, stream);
}

// Fault injection for the flag protocol (tests only): SHIPYARD_FAULT_INJECT=drop_signal:<rank>:<n> makes rank <rank>'s n-th
// (1-based) put_signal deliver the payload but never raise the flag, so the peer's bounded wait must end in SY_ERR_TIMEOUT
// (the watchdog) instead of a hang.  (The task runner interprets the kill_rank:... form of the same variable.)
static bool fault_drop_signal(const sy_comm* c) {
  static int target = -2, nth = 0, calls = 0;
  if (target == -2) {
    target = -1;
    const char* fi = getenv("SHIPYARD_FAULT_INJECT");
    int r = 0, n = 0;
    if (fi && sscanf(fi, "drop_signal:%d:%d", &r, &n) == 2) { target = r; nth = n; }
  }
  return target == c->rank && ++calls == nth;
}

extern "C" int sy_put_signal(sy_comm* c, const void* src, size_t dst_off, size_t bytes, int peer, int sig, sy_stream_t stream) {
  if (peer < 0 || peer >= c->world || sig < 0 || sig >= SY_NSIG || dst_off + bytes > c->heap_bytes) return SY_ERR_ARG;
  if (fault_drop_signal(c)) {
    if (is_stub(c)) { memcpy(c->dev.heap[peer] + dst_off, src, bytes); return SY_OK; }
    CUDA_TRY(cudaMemcpyAsync(c->dev.heap[peer] + dst_off, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SY_OK;
  }
  if (is_stub(c)) return stub_put_signal(c, src, dst_off, bytes, peer, sig);
  return k_put_signal(c, src, dst_off, bytes, peer, sig, stream);
}
extern "C" int sy_wait_signal(sy_comm* c, int sig, uint32_t expected, sy_stream_t stream) {
  if (sig < 0 || sig >= SY_NSIG) return SY_ERR_ARG;
  if (is_stub(c)) return stub_wait_signal(c, sig, expected);
  return k_wait_signal(c, sig, expected, stream);
}

extern "C" int sy_halo_exchange(sy_comm* c, const void* src, int dt, const sy_halo_desc* descs, int ndesc,
                                const int* wait_sig, int nwait, sy_stream_t stream) {
  if (is_stub(c)) {
    // CPU reference: strided gather + memcpy, then wait on cumulative counters
    static thread_local uint32_t expect[SY_NSIG];
    size_t es = sy_dtype_size(dt);
    for (int i = 0; i < ndesc; ++i) {
      const sy_halo_desc& d = descs[i];
      size_t n = (size_t)d.nx * d.ny * d.nz;
      std::vector<char> tmp(n * es);
      for (size_t k = 0; k < n; ++k) {
        long x = k % d.nx, y = (k / d.nx) % d.ny, z = k / ((size_t)d.nx * d.ny);
        memcpy(tmp.data() + k * es, (const char*)src + (d.src_elem_off + x * d.sx + y * d.sy + z * d.sz) * es, es);
      }
      int rc = stub_put_signal(c, tmp.data(), (size_t)d.dst_off, n * es, d.peer, d.sig_idx);
      if (rc) return rc;
    }
    for (int i = 0; i < nwait; ++i) {
      int rc = stub_wait_signal(c, wait_sig[i], ++expect[wait_sig[i]]);
      if (rc) return rc;
    }
    return SY_OK;
  }
  return k_halo(c, src, dt, descs, ndesc, wait_sig, nwait, stream);
}

extern "C" int sy_fused_allreduce_sgd(sy_comm* c, void* grads, int dt_grad, void* params, int dt_param, float* master,
                                      float* mom, const float* hyper, size_t count, int zero_grads, sy_stream_t stream) {
  if (is_stub(c)) return stub_fused_sgd(c, grads, dt_grad, params, dt_param, master, mom, hyper, count, zero_grads);
  size_t g_off, p_off;
  if (!sym_off(c, grads, &g_off) || !sym_off(c, params, &p_off)) {
    sy_set_error("fused_allreduce_sgd: grads and params must come from sy_sym_alloc"); return SY_ERR_ARG;
  }
  return k_fused_sgd(c, g_off, dt_grad, p_off, dt_param, master, mom, hyper, count, zero_grads, stream);
}

extern "C" int sy_fused_allreduce_adam(sy_comm* c, float* grad, float* param, float* m, float* v, float* hyper, size_t count,
                                       float scale, int zero_grad, sy_stream_t stream) {
  if (count == 0) return SY_OK;
  if (is_stub(c)) return stub_fused_adam(c, grad, param, m, v, hyper, count, scale, zero_grad);
  return k_oneshot_adam(c, grad, param, m, v, hyper, count, scale, zero_grad, stream);
}

extern "C" int sy_allreduce_fp8_blockscaled(sy_comm* c, const void* in, int dt_in, void* out_q, void* out_scales,
                                            size_t count, float scale, sy_stream_t stream) {
  if (is_stub(c)) return stub_allreduce_fp8(c, in, dt_in, out_q, out_scales, count, scale);
  size_t in_off, q_off, s_off;
  if (!sym_off(c, in, &in_off) || !sym_off(c, out_q, &q_off) || !sym_off(c, out_scales, &s_off)) {
    sy_set_error("allreduce_fp8: in/out_q/out_scales must be symmetric allocations"); return SY_ERR_ARG;
  }
  return k_allreduce_fp8(c, in_off, dt_in, out_q, out_scales, count, scale, stream);
}
