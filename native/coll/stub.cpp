// Host shared-memory transport: same API as the GPU transports, CPU reductions.
// Lets `shipyard jobs add` multi-instance tasks, the MPI face and every
// collective run on a box with no GPU (BASELINE.json config #1: world_size=2
// dry-run with the stub collectives shim).  Stands in for the MPI / NCCL libraries the reference's recipe
// images bring along (e.g. /root/reference/recipes/mpiBench-OpenMPI/docker/Dockerfile:21-35).
#include <atomic>
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <sched.h>
#include <string.h>
#include <vector>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>
#include "internal.h"
#include "numeric.h"

struct StubHdr {
  std::atomic<uint64_t> arrive;
  char pad[4096 - sizeof(std::atomic<uint64_t>)];
};

static std::string shm_name(const sy_comm* c) { return "/shipyard-coll-" + c->session; }

int stub_init(sy_comm* c) {
  size_t total = sizeof(StubHdr) + (size_t)c->world * c->heap_bytes;
  std::string name = shm_name(c);
  int fd = -1;
  if (c->rank == 0) {
    shm_unlink(name.c_str());
    fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)total) < 0) {
      sy_set_error("stub: shm_open/ftruncate %s: %s", name.c_str(), strerror(errno));
      if (fd >= 0) close(fd);
      return SY_ERR_SYS;
    }
  }
  if (hub_barrier(c->hub) < 0) return SY_ERR_SYS;
  if (c->rank != 0) {
    fd = shm_open(name.c_str(), O_RDWR, 0600);
    if (fd < 0) { sy_set_error("stub: shm_open %s: %s", name.c_str(), strerror(errno)); return SY_ERR_SYS; }
  }
  void* p = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { sy_set_error("stub: mmap: %s", strerror(errno)); return SY_ERR_SYS; }
  c->shm_base = p; c->shm_bytes = total;
  if (c->rank == 0) new (p) StubHdr();
  if (hub_barrier(c->hub) < 0) return SY_ERR_SYS;
  if (c->rank == 0) shm_unlink(name.c_str());  // everyone has it mapped; no residue
  for (int r = 0; r < c->world; ++r)
    c->dev.heap[r] = (char*)p + sizeof(StubHdr) + (size_t)r * c->heap_bytes;
  c->dev.mc = nullptr;
  c->transport = SY_TRANSPORT_STUB;
  return SY_OK;
}

void stub_destroy(sy_comm* c) {
  if (c->shm_base) munmap(c->shm_base, c->shm_bytes);
  c->shm_base = nullptr;
}

int stub_barrier(sy_comm* c) {
  if (c->world == 1) return SY_OK;
  StubHdr* h = (StubHdr*)c->shm_base;
  c->stub_gen += 1;
  uint64_t target = c->stub_gen * (uint64_t)c->world;
  h->arrive.fetch_add(1, std::memory_order_acq_rel);
  struct timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
  unsigned spins = 0;
  while (h->arrive.load(std::memory_order_acquire) < target) {
    if (++spins > 200) { sched_yield(); }
    if ((spins & 0xfff) == 0) {
      struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
      double ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
      if (ms > (double)c->timeout_ms) {
        sy_set_error("stub: barrier timeout on rank %d (peer died?)", c->rank);
        if (c->status_host) *c->status_host = SY_ERR_TIMEOUT;
        return SY_ERR_TIMEOUT;
      }
    }
  }
  return SY_OK;
}

// ---- staging: collectives read peers' data from the symmetric heap ----------
static char* stage_ptr(sy_comm* c, int peer, int half) {
  return c->dev.heap[peer] + c->stage_off + (size_t)half * (c->stage_bytes / 2);
}
static bool in_heap(sy_comm* c, const void* p, size_t* off) {
  const char* b = c->dev.heap[c->rank];
  if ((const char*)p >= b && (const char*)p < b + c->heap_bytes) { *off = (const char*)p - b; return true; }
  return false;
}

template <typename Acc> struct OpApply {
  static inline Acc run(int op, Acc a, Acc b) {
    switch (op) {
      case SY_SUM: return a + b;
      case SY_MAX: return a > b ? a : b;
      case SY_MIN: return a < b ? a : b;
      case SY_PROD: return a * b;
    }
    return a;
  }
};

static inline bool is_float(int dt) { return dt == SY_F32 || dt == SY_BF16 || dt == SY_F16; }

// reduce element range [lo,hi) of `srcs[r]` (world pointers) into dst (dt_out), rank order
static void reduce_range(int world, const void* const* srcs, void* dst, size_t lo, size_t hi,
                         size_t dst_lo, int dt_in, int dt_out, float scale, int op) {
  for (size_t i = lo; i < hi; ++i) {
    size_t o = dst_lo + (i - lo);
    if (dt_in == SY_F64) {
      double a = ((const double*)srcs[0])[i];
      for (int r = 1; r < world; ++r) a = OpApply<double>::run(op, a, ((const double*)srcs[r])[i]);
      if (scale != 1.0f) a *= (double)scale;
      syn::store_f(dst, o, dt_out, a);
    } else if (is_float(dt_in)) {
      float a = syn::load_f32(srcs[0], i, dt_in);
      for (int r = 1; r < world; ++r) a = OpApply<float>::run(op, a, syn::load_f32(srcs[r], i, dt_in));
      if (scale != 1.0f) a *= scale;
      syn::store_f(dst, o, dt_out, (double)a);
    } else {
      int64_t a = syn::load_i64(srcs[0], i, dt_in);
      for (int r = 1; r < world; ++r) a = OpApply<int64_t>::run(op, a, syn::load_i64(srcs[r], i, dt_in));
      syn::store_i(dst, o, dt_out, a);
    }
  }
}

// publish `in` (bytes) so that peers can read it; returns per-rank pointers
static int publish(sy_comm* c, const void* in, size_t bytes, const void** srcs) {
  size_t off;
  if (in_heap(c, in, &off)) {
    for (int r = 0; r < c->world; ++r) srcs[r] = c->dev.heap[r] + off;
  } else {
    if (bytes > c->stage_bytes / 2) { sy_set_error("stub: message larger than staging"); return SY_ERR_NOMEM; }
    memcpy(stage_ptr(c, c->rank, 0), in, bytes);
    for (int r = 0; r < c->world; ++r) srcs[r] = stage_ptr(c, r, 0);
  }
  return SY_OK;
}

int stub_allreduce(sy_comm* c, const void* in, void* out, size_t count, int dt_in, int dt_out,
                   float scale, int op) {
  const void* srcs[SY_MAXR];
  size_t esz = sy_dtype_size(dt_in), chunk_elems = (c->stage_bytes / 2) / esz;
  size_t off; bool sym = in_heap(c, in, &off);
  if (sym) chunk_elems = count ? count : 1;
  // in-place on symmetric memory needs a private result buffer until all ranks have read
  for (size_t base = 0; base < count || (count == 0 && base == 0); base += chunk_elems) {
    size_t n = count - base < chunk_elems ? count - base : chunk_elems;
    int e = publish(c, (const char*)in + base * esz, n * esz, srcs);
    if (e) return e;
    if ((e = stub_barrier(c))) return e;
    std::vector<char> tmp(n * sy_dtype_size(dt_out));
    reduce_range(c->world, srcs, tmp.data(), 0, n, 0, dt_in, dt_out, scale, op);
    if ((e = stub_barrier(c))) return e;
    memcpy((char*)out + base * sy_dtype_size(dt_out), tmp.data(), tmp.size());
    if (count == 0) break;
  }
  // symmetric outputs written after the second barrier: make them visible before return
  return stub_barrier(c);
}

int stub_reduce_scatter(sy_comm* c, const void* in, void* out, size_t count, int dt_in, int dt_out,
                        float scale, int op) {
  const void* srcs[SY_MAXR];
  size_t esz = sy_dtype_size(dt_in);
  int e = publish(c, in, (size_t)c->world * count * esz, srcs);
  if (e) return e;
  if ((e = stub_barrier(c))) return e;
  std::vector<char> tmp(count * sy_dtype_size(dt_out));
  reduce_range(c->world, srcs, tmp.data(), (size_t)c->rank * count, (size_t)(c->rank + 1) * count, 0,
               dt_in, dt_out, scale, op);
  if ((e = stub_barrier(c))) return e;
  memcpy(out, tmp.data(), tmp.size());
  return stub_barrier(c);
}

int stub_allgather(sy_comm* c, const void* in, void* out, size_t count, int dt) {
  const void* srcs[SY_MAXR];
  size_t bytes = count * sy_dtype_size(dt);
  int e = publish(c, in, bytes, srcs);
  if (e) return e;
  if ((e = stub_barrier(c))) return e;
  std::vector<char> tmp((size_t)c->world * bytes);
  for (int r = 0; r < c->world; ++r) memcpy(tmp.data() + (size_t)r * bytes, srcs[r], bytes);
  if ((e = stub_barrier(c))) return e;
  memcpy(out, tmp.data(), tmp.size());
  return stub_barrier(c);
}

int stub_broadcast(sy_comm* c, const void* in, void* out, size_t count, int dt, int root) {
  // rooted: always relay through the root's staging half (chunked), so no agreement on
  // buffer symmetry is needed between root and non-roots
  size_t bytes = count * sy_dtype_size(dt), chunk = c->stage_bytes / 2;
  int e;
  for (size_t base = 0; base < bytes || base == 0; base += chunk) {
    size_t n = bytes - base < chunk ? bytes - base : chunk;
    if (c->rank == root) memcpy(stage_ptr(c, root, 1), (const char*)in + base, n);
    if ((e = stub_barrier(c))) return e;
    if (c->rank != root || out != in) memcpy((char*)out + base, stage_ptr(c, root, 1), n);
    if ((e = stub_barrier(c))) return e;
    if (bytes == 0) break;
  }
  return SY_OK;
}

int stub_alltoall(sy_comm* c, const void* in, void* out, size_t count, int dt) {
  const void* srcs[SY_MAXR];
  size_t bytes = count * sy_dtype_size(dt);
  int e = publish(c, in, (size_t)c->world * bytes, srcs);
  if (e) return e;
  if ((e = stub_barrier(c))) return e;
  std::vector<char> tmp((size_t)c->world * bytes);
  for (int r = 0; r < c->world; ++r)
    memcpy(tmp.data() + (size_t)r * bytes, (const char*)srcs[r] + (size_t)c->rank * bytes, bytes);
  if ((e = stub_barrier(c))) return e;
  memcpy(out, tmp.data(), tmp.size());
  return stub_barrier(c);
}

int stub_reduce(sy_comm* c, const void* in, void* out, size_t count, int dt, int op, int root) {
  const void* srcs[SY_MAXR];
  int e = publish(c, in, count * sy_dtype_size(dt), srcs);
  if (e) return e;
  if ((e = stub_barrier(c))) return e;
  std::vector<char> tmp;
  if (c->rank == root) {
    tmp.resize(count * sy_dtype_size(dt));
    reduce_range(c->world, srcs, tmp.data(), 0, count, 0, dt, dt, 1.0f, op);
  }
  if ((e = stub_barrier(c))) return e;
  if (c->rank == root) memcpy(out, tmp.data(), tmp.size());
  return stub_barrier(c);
}

int stub_gather(sy_comm* c, const void* in, void* out, size_t count, int dt, int root) {
  const void* srcs[SY_MAXR];
  size_t bytes = count * sy_dtype_size(dt);
  int e = publish(c, in, bytes, srcs);
  if (e) return e;
  if ((e = stub_barrier(c))) return e;
  std::vector<char> tmp;
  if (c->rank == root) {
    tmp.resize((size_t)c->world * bytes);
    for (int r = 0; r < c->world; ++r) memcpy(tmp.data() + (size_t)r * bytes, srcs[r], bytes);
  }
  if ((e = stub_barrier(c))) return e;
  if (c->rank == root) memcpy(out, tmp.data(), tmp.size());
  return stub_barrier(c);
}

int stub_scatter(sy_comm* c, const void* in, void* out, size_t count, int dt, int root) {
  size_t bytes = count * sy_dtype_size(dt);
  if ((size_t)c->world * bytes > c->stage_bytes / 2) { sy_set_error("stub: scatter too large"); return SY_ERR_NOMEM; }
  int e;
  if (c->rank == root) memcpy(stage_ptr(c, root, 1), in, (size_t)c->world * bytes);
  if ((e = stub_barrier(c))) return e;
  memcpy(out, stage_ptr(c, root, 1) + (size_t)c->rank * bytes, bytes);
  return stub_barrier(c);
}

int stub_put_signal(sy_comm* c, const void* src, size_t dst_off, size_t bytes, int peer, int sig) {
  if (dst_off + bytes > c->heap_bytes || sig < 0 || sig >= SY_NSIG) return SY_ERR_ARG;
  memcpy(c->dev.heap[peer] + dst_off, src, bytes);
  std::atomic<uint32_t>* s = (std::atomic<uint32_t>*)(c->dev.heap[peer] + SY_SIG_OFF) + sig;
  s->fetch_add(1, std::memory_order_release);
  return SY_OK;
}

int stub_wait_signal(sy_comm* c, int sig, uint32_t expected) {
  if (sig < 0 || sig >= SY_NSIG) return SY_ERR_ARG;
  std::atomic<uint32_t>* s = (std::atomic<uint32_t>*)(c->dev.heap[c->rank] + SY_SIG_OFF) + sig;
  struct timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
  unsigned spins = 0;
  while ((int32_t)(s->load(std::memory_order_acquire) - expected) < 0) {
    if (++spins > 200) sched_yield();
    if ((spins & 0xfff) == 0) {
      struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
      double ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
      if (ms > (double)c->timeout_ms) { sy_set_error("stub: wait_signal timeout"); return SY_ERR_TIMEOUT; }
    }
  }
  return SY_OK;
}

int stub_fused_sgd(sy_comm* c, void* grads, int dt_grad, void* params, int dt_param, float* master,
                   float* mom, const float* hyper, size_t count, int zero_grads) {
  size_t goff, poff;
  if (!in_heap(c, grads, &goff) || !in_heap(c, params, &poff)) {
    sy_set_error("fused_sgd: grads/params must be symmetric allocations"); return SY_ERR_ARG;
  }
  float lr = hyper[0], mu = hyper[1], wd = hyper[2], scale = hyper[3];
  size_t lo = sy_shard_begin(c, count, c->rank), n = sy_shard_count(c, count, c->rank);
  int e;
  if ((e = stub_barrier(c))) return e;
  for (size_t k = 0; k < n; ++k) {
    size_t i = lo + k;
    float g = 0.f;
    for (int r = 0; r < c->world; ++r) g += syn::load_f32(c->dev.heap[r] + goff, i, dt_grad);
    g *= scale;
    g += wd * master[k];
    float m = mu * mom[k] + g;
    mom[k] = m;
    float w = master[k] - lr * m;
    master[k] = w;
    for (int r = 0; r < c->world; ++r) syn::store_f(c->dev.heap[r] + poff, i, dt_param, (double)w);
  }
  if ((e = stub_barrier(c))) return e;
  if (zero_grads) memset(grads, 0, count * sy_dtype_size(dt_grad));
  return stub_barrier(c);
}

int stub_fused_adam(sy_comm* c, float* grad, float* param, float* m, float* v, float* hyper, size_t count, float scale, int zero_grad) {
  const void* srcs[SY_MAXR];
  int e = publish(c, grad, count * sizeof(float), srcs);
  if (e) return e;
  if ((e = stub_barrier(c))) return e;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], t = hyper[4] + 1.0f;
  const float c1 = 1.0f - powf(b1, t), c2 = 1.0f - powf(b2, t);
  std::vector<float> g(count);
  for (size_t i = 0; i < count; ++i) {
    float a = 0.f;
    for (int r = 0; r < c->world; ++r) a += ((const float*)srcs[r])[i];
    g[i] = a * scale;
  }
  if ((e = stub_barrier(c))) return e;          // everyone has read every gradient before anyone zeroes it
  for (size_t i = 0; i < count; ++i) {
    m[i] = b1 * m[i] + (1.0f - b1) * g[i];
    v[i] = b2 * v[i] + (1.0f - b2) * g[i] * g[i];
    param[i] -= (lr / c1) * m[i] / (sqrtf(v[i]) / sqrtf(c2) + eps);
  }
  hyper[4] = t;
  if (zero_grad) memset(grad, 0, count * sizeof(float));
  return stub_barrier(c);
}

int stub_allreduce_fp8(sy_comm* c, const void* in, int dt_in, void* out_q, void* out_scales,
                       size_t count, float scale) {
  const void* srcs[SY_MAXR];
  int e = publish(c, in, count * sy_dtype_size(dt_in), srcs);
  if (e) return e;
  if ((e = stub_barrier(c))) return e;
  uint8_t* q = (uint8_t*)out_q; uint8_t* sc = (uint8_t*)out_scales;
  for (size_t b = 0; b < (count + 31) / 32; ++b) {
    float v[32]; float amax = 0.f;
    size_t n = count - b * 32 < 32 ? count - b * 32 : 32;
    for (size_t k = 0; k < n; ++k) {
      float a = 0.f;
      for (int r = 0; r < c->world; ++r) a += syn::load_f32(srcs[r], b * 32 + k, dt_in);
      v[k] = a * scale; amax = fmaxf(amax, fabsf(v[k]));
    }
    uint8_t e8 = syn::e8m0_from_amax(amax);
    float inv = syn::e8m0_inv_scale(e8);
    sc[b] = e8;
    for (size_t k = 0; k < n; ++k) q[b * 32 + k] = syn::f32_to_e4m3(v[k] * inv);
  }
  return stub_barrier(c);
}
