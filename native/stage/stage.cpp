// libshipyard_stage — pinned staging arena + async host->HBM data mover.
//
// The B200-native replacement for the reference's image pre-load (cascade: `docker pull`
// with bounded concurrency, /root/reference/cascade/cascade.py:500-646) and its blobxfer /
// scp data ingress (/root/reference/convoy/data.py:492-978, scripts/shipyard_blobxfer.sh):
// artefacts and input batches are read from disk (or taken from host memory) in chunks into a
// cudaHostAlloc'd arena and pushed with cudaMemcpyAsync on per-worker copy streams into
// per-GPU HBM.  `concurrency` workers bound the number of simultaneous transfers (the
// reference's concurrent_source_downloads); each worker double-buffers its arena slice so
// disk reads overlap PCIe copies.  A ticket's completion is a CUDA event, so the first task
// step can be event-chained behind the copy (sy_stage_stream_wait) instead of blocking the host.
//
// device < 0 selects host-only mode (no CUDA): data lands in a malloc'd buffer — used on CPU
// boxes and by the control-plane tests.
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

static double now_s() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }

enum TicketState { T_QUEUED = 0, T_RUNNING = 1, T_DONE = 2, T_FAILED = 3 };

struct Ticket {
  long id = 0;
  std::string path;            // file source (empty => host memory source)
  std::string dst_path;        // file -> file copy (the task-side data mover): no device involved
  const void* host_src = nullptr;
  bool pinned_src = false;     // host_src is page-locked: one cudaMemcpyAsync straight from it, no bounce through the arena
  cudaEvent_t wait_ev = nullptr;   // the copy stream waits for this event first (e.g. "the step that read dptr last is done")
  size_t offset = 0, bytes = 0;
  void* dptr = nullptr;        // destination (device, or host in host-only mode)
  bool own_dptr = false;
  int state = T_QUEUED;
  size_t done_bytes = 0;
  double t_submit = 0, t_start = 0, t_end = 0;
  cudaEvent_t done_ev = nullptr;
  bool ev_recorded = false;
  std::string error;
};

struct sy_stage {
  int device = -1;
  size_t arena_bytes = 0, chunk_bytes = 0;
  int concurrency = 1, chunks_per_worker = 2;
  char* arena = nullptr;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv, cv_done;
  std::deque<long> queue;
  std::map<long, Ticket*> tickets;
  long next_id = 1;
  bool stop = false;
  std::atomic<unsigned long long> total_bytes{0};
  std::atomic<unsigned long long> memcpy_calls{0};
  double busy_seconds = 0;
};

static thread_local std::string g_err;
extern "C" const char* sy_stage_last_error() { return g_err.c_str(); }

static void worker_main(sy_stage* s, int wi) {
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> chunk_ev(s->chunks_per_worker, nullptr);
  const bool gpu = s->device >= 0;
  if (gpu) {
    cudaSetDevice(s->device);
    cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    for (auto& e : chunk_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  }
  char* slice = s->arena + (size_t)wi * s->chunks_per_worker * s->chunk_bytes;
  for (;;) {
    Ticket* t = nullptr;
    {
      std::unique_lock<std::mutex> lk(s->mu);
      s->cv.wait(lk, [&] { return s->stop || !s->queue.empty(); });
      if (s->stop && s->queue.empty()) break;
      long id = s->queue.front(); s->queue.pop_front();
      t = s->tickets[id];
      t->state = T_RUNNING; t->t_start = now_s();
    }
    int fd = -1; bool ok = true; std::string err;
    if (!t->path.empty()) {
      fd = open(t->path.c_str(), O_RDONLY | O_CLOEXEC);
      if (fd < 0) { ok = false; err = "open " + t->path + ": " + strerror(errno); }
      else {
#ifdef POSIX_FADV_SEQUENTIAL
        posix_fadvise(fd, (off_t)t->offset, (off_t)t->bytes, POSIX_FADV_SEQUENTIAL);
#endif
      }
    }
    size_t done = 0; int ci = 0;
    std::vector<bool> chunk_busy(s->chunks_per_worker, false);
    if (ok && !t->dst_path.empty()) {
      // file -> file: copy_file_range keeps the bytes in the kernel (reflink / server-side copy where the file system can); when it is
      // refused (EXDEV, EINVAL, ENOSYS, old kernels) the worker's arena slice is the bounce buffer.  Mode and mtime are carried over.
      struct stat st{};
      if (fstat(fd, &st) != 0) { ok = false; err = "fstat " + t->path + ": " + strerror(errno); }
      int out = ok ? open(t->dst_path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, st.st_mode & 07777) : -1;
      if (ok && out < 0) { ok = false; err = "open " + t->dst_path + ": " + strerror(errno); }
      if (ok) {
        const size_t total = t->bytes ? t->bytes : (size_t)st.st_size;
        t->bytes = total;
        bool use_cfr = true;
        off_t in_off = (off_t)t->offset, out_off = 0;
        while (ok && done < total) {
          if (use_cfr) {
            ssize_t r = copy_file_range(fd, &in_off, out, &out_off, total - done, 0);
            if (r > 0) { done += (size_t)r; { std::lock_guard<std::mutex> lk(s->mu); t->done_bytes = done; } continue; }
            if (r == 0) { ok = false; err = "short read on " + t->path; break; }
            if (errno == EINTR) continue;
            use_cfr = false;                                  // fall through to the bounce path from the current offsets
          }
          const size_t n = total - done < s->chunk_bytes ? total - done : s->chunk_bytes;
          size_t got = 0;
          while (got < n) {
            ssize_t r = pread(fd, slice + got, n - got, (off_t)(t->offset + done + got));
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) { ok = false; err = "short read on " + t->path; break; }
            got += (size_t)r;
          }
          size_t put = 0;
          while (ok && put < n) {
            ssize_t w = pwrite(out, slice + put, n - put, (off_t)(done + put));
            if (w < 0 && errno == EINTR) continue;
            if (w <= 0) { ok = false; err = "write " + t->dst_path + ": " + strerror(errno); break; }
            put += (size_t)w;
          }
          if (ok) { done += n; std::lock_guard<std::mutex> lk(s->mu); t->done_bytes = done; }
        }
        if (ok) {
          fchmod(out, st.st_mode & 07777);
          struct timespec ts[2] = {st.st_atim, st.st_mtim};
          futimens(out, ts);
        }
      }
      if (out >= 0 && close(out) != 0 && ok) { ok = false; err = "close " + t->dst_path + ": " + strerror(errno); }
      done = t->bytes;                                        // skip the device loop below
    }
    if (ok && gpu && t->wait_ev) {
      cudaError_t e = cudaStreamWaitEvent(stream, t->wait_ev, 0);
      if (e != cudaSuccess) { ok = false; err = std::string("cudaStreamWaitEvent: ") + cudaGetErrorString(e); }
    }
    if (ok && t->pinned_src && t->bytes) {
      // page-locked source (a pinned input batch): DMA straight from it, the arena is not involved
      if (gpu) {
        cudaError_t e = cudaMemcpyAsync(t->dptr, t->host_src, t->bytes, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) { ok = false; err = std::string("cudaMemcpyAsync(pinned): ") + cudaGetErrorString(e); }
        s->memcpy_calls.fetch_add(1);
      } else {
        memcpy(t->dptr, t->host_src, t->bytes);
      }
      done = t->bytes;
      { std::lock_guard<std::mutex> lk(s->mu); t->done_bytes = done; }
    }
    while (ok && done < t->bytes) {
      const size_t n = t->bytes - done < s->chunk_bytes ? t->bytes - done : s->chunk_bytes;
      char* buf = slice + (size_t)ci * s->chunk_bytes;
      if (gpu && chunk_busy[ci]) { cudaEventSynchronize(chunk_ev[ci]); chunk_busy[ci] = false; }   // chunk free again?
      if (fd >= 0) {
        size_t got = 0;
        while (got < n) {
          ssize_t r = pread(fd, buf + got, n - got, (off_t)(t->offset + done + got));
          if (r < 0 && errno == EINTR) continue;
          if (r <= 0) { ok = false; err = "short read on " + t->path; break; }
          got += (size_t)r;
        }
        if (!ok) break;
      } else {
        memcpy(buf, (const char*)t->host_src + done, n);   // pageable -> pinned bounce
      }
      if (gpu) {
        cudaError_t e = cudaMemcpyAsync((char*)t->dptr + done, buf, n, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) { ok = false; err = std::string("cudaMemcpyAsync: ") + cudaGetErrorString(e); break; }
        cudaEventRecord(chunk_ev[ci], stream); chunk_busy[ci] = true;
        s->memcpy_calls.fetch_add(1);
      } else {
        memcpy((char*)t->dptr + done, buf, n);
      }
      done += n; ci = (ci + 1) % s->chunks_per_worker;
      { std::lock_guard<std::mutex> lk(s->mu); t->done_bytes = done; }
    }
    if (fd >= 0) close(fd);
    if (gpu && ok) {
      cudaEventRecord(t->done_ev, stream);
      { std::lock_guard<std::mutex> lk(s->mu); t->ev_recorded = true; }
      s->cv_done.notify_all();
      cudaError_t e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess) { ok = false; err = std::string("stream sync: ") + cudaGetErrorString(e); }
    }
    {
      std::lock_guard<std::mutex> lk(s->mu);
      t->t_end = now_s();
      t->state = ok ? T_DONE : T_FAILED;
      t->error = err;
      t->ev_recorded = true;
      s->busy_seconds += t->t_end - t->t_start;
      if (ok) s->total_bytes.fetch_add(t->bytes);
    }
    s->cv_done.notify_all();
  }
  if (gpu) { for (auto& e : chunk_ev) cudaEventDestroy(e); cudaStreamDestroy(stream); }
}

extern "C" int sy_stage_create(sy_stage** out, int device, size_t arena_bytes, int concurrency, int chunks_per_worker) {
  if (!out || concurrency < 1 || concurrency > 64) { g_err = "bad arguments"; return 1; }
  sy_stage* s = new sy_stage();
  s->device = device; s->concurrency = concurrency;
  s->chunks_per_worker = chunks_per_worker < 2 ? 2 : chunks_per_worker;
  if (arena_bytes < (size_t)concurrency * s->chunks_per_worker * (1u << 20)) arena_bytes = (size_t)concurrency * s->chunks_per_worker * (1u << 20);
  s->chunk_bytes = arena_bytes / ((size_t)concurrency * s->chunks_per_worker) / 4096 * 4096;
  s->arena_bytes = s->chunk_bytes * concurrency * s->chunks_per_worker;
  if (device >= 0) {
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s->arena, s->arena_bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { g_err = std::string("pinned arena: ") + cudaGetErrorString(e); delete s; return 2; }
  } else {
    s->arena = (char*)malloc(s->arena_bytes);
    if (!s->arena) { g_err = "arena malloc failed"; delete s; return 3; }
  }
  for (int i = 0; i < concurrency; ++i) s->workers.emplace_back(worker_main, s, i);
  *out = s;
  return 0;
}

extern "C" int sy_stage_destroy(sy_stage* s) {
  if (!s) return 0;
  { std::lock_guard<std::mutex> lk(s->mu); s->stop = true; }
  s->cv.notify_all();
  for (auto& w : s->workers) w.join();
  for (auto& kv : s->tickets) {
    Ticket* t = kv.second;
    if (t->own_dptr && t->dptr) { if (s->device >= 0) cudaFree(t->dptr); else free(t->dptr); }
    if (t->done_ev) cudaEventDestroy(t->done_ev);
    delete t;
  }
  if (s->device >= 0) cudaFreeHost(s->arena); else free(s->arena);
  delete s;
  return 0;
}

static long submit(sy_stage* s, Ticket* t) {
  if (!t->dptr && t->dst_path.empty()) {
    if (s->device >= 0) {
      cudaSetDevice(s->device);
      cudaError_t e = cudaMalloc(&t->dptr, t->bytes ? t->bytes : 1);
      if (e != cudaSuccess) { g_err = std::string("cudaMalloc: ") + cudaGetErrorString(e); delete t; return -1; }
    } else {
      t->dptr = malloc(t->bytes ? t->bytes : 1);
    }
    t->own_dptr = true;
  }
  if (s->device >= 0) cudaEventCreateWithFlags(&t->done_ev, cudaEventDisableTiming);
  t->t_submit = now_s();
  std::lock_guard<std::mutex> lk(s->mu);
  t->id = s->next_id++;
  s->tickets[t->id] = t;
  s->queue.push_back(t->id);
  s->cv.notify_one();
  return t->id;
}

extern "C" long sy_stage_submit_file(sy_stage* s, const char* path, void* dptr, size_t offset, size_t bytes) {
  struct stat st;
  if (stat(path, &st) != 0) { g_err = std::string("stat ") + path + ": " + strerror(errno); return -1; }
  if ((size_t)st.st_size < offset) { g_err = "offset beyond end of file"; return -1; }
  if (bytes == 0 || offset + bytes > (size_t)st.st_size) bytes = (size_t)st.st_size - offset;
  Ticket* t = new Ticket();
  t->path = path; t->offset = offset; t->bytes = bytes; t->dptr = dptr;
  return submit(s, t);
}

// file -> file copy on a worker thread (bytes = 0: the whole file from `offset`); wait / query / release as for any ticket
extern "C" long sy_stage_submit_copy(sy_stage* s, const char* src, const char* dst, size_t offset, size_t bytes) {
  if (!src || !dst || !src[0] || !dst[0]) { g_err = "submit_copy: empty path"; return -1; }
  Ticket* t = new Ticket();
  t->path = src; t->dst_path = dst; t->offset = offset; t->bytes = bytes;
  return submit(s, t);
}

extern "C" long sy_stage_submit_host(sy_stage* s, const void* host, size_t bytes, void* dptr) {
  Ticket* t = new Ticket();
  t->host_src = host; t->bytes = bytes; t->dptr = dptr;
  return submit(s, t);
}

// Page-locked host source (cudaHostAlloc / cudaHostRegister / torch pin_memory): copied with ONE cudaMemcpyAsync on a worker's copy
// stream, after `wait_event` (a cudaEvent_t, may be null) has completed on the device — the double-buffered input path of a
// training loop: "copy batch i+1 into the slot as soon as step i-1, which read that slot, is done", without blocking the host.
extern "C" long sy_stage_submit_pinned(sy_stage* s, const void* host_pinned, size_t bytes, void* dptr, void* wait_event) {
  if (!dptr && s->device >= 0 && bytes == 0) { g_err = "empty transfer"; return -1; }
  Ticket* t = new Ticket();
  t->host_src = host_pinned; t->pinned_src = true; t->bytes = bytes; t->dptr = dptr; t->wait_ev = (cudaEvent_t)wait_event;
  return submit(s, t);
}

// page-locked host memory from the stager (so callers without a CUDA framework can fill batches in place)
extern "C" void* sy_stage_pinned_alloc(sy_stage* s, size_t bytes) {
  void* p = nullptr;
  if (s->device >= 0) { cudaSetDevice(s->device); if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; } }
  else p = malloc(bytes);
  return p;
}
extern "C" void sy_stage_pinned_free(sy_stage* s, void* p) { if (!p) return; if (s->device >= 0) cudaFreeHost(p); else free(p); }

// 0 = done, 1 = timeout, 2 = failed, 3 = unknown ticket
extern "C" int sy_stage_wait(sy_stage* s, long id, double timeout_s) {
  std::unique_lock<std::mutex> lk(s->mu);
  auto it = s->tickets.find(id);
  if (it == s->tickets.end()) return 3;
  Ticket* t = it->second;
  auto pred = [&] { return t->state == T_DONE || t->state == T_FAILED; };
  if (timeout_s < 0) s->cv_done.wait(lk, pred);
  else if (!s->cv_done.wait_for(lk, std::chrono::duration<double>(timeout_s), pred)) return 1;
  if (t->state == T_FAILED) { g_err = t->error; return 2; }
  return 0;
}

// make `stream` wait for the ticket's copy without blocking the host beyond event recording
extern "C" int sy_stage_stream_wait(sy_stage* s, long id, void* stream) {
  if (s->device < 0) return sy_stage_wait(s, id, -1);
  std::unique_lock<std::mutex> lk(s->mu);
  auto it = s->tickets.find(id);
  if (it == s->tickets.end()) return 3;
  Ticket* t = it->second;
  s->cv_done.wait(lk, [&] { return t->ev_recorded; });
  if (t->state == T_FAILED) { g_err = t->error; return 2; }
  cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, t->done_ev, 0);
  return e == cudaSuccess ? 0 : 2;
}

extern "C" void* sy_stage_ptr(sy_stage* s, long id) {
  std::lock_guard<std::mutex> lk(s->mu);
  auto it = s->tickets.find(id);
  return it == s->tickets.end() ? nullptr : it->second->dptr;
}

// out[0]=state out[1]=bytes out[2]=done_bytes ; secs[0]=queue wait secs[1]=transfer seconds
extern "C" int sy_stage_query(sy_stage* s, long id, unsigned long long* out, double* secs) {
  std::lock_guard<std::mutex> lk(s->mu);
  auto it = s->tickets.find(id);
  if (it == s->tickets.end()) return 3;
  Ticket* t = it->second;
  out[0] = (unsigned long long)t->state; out[1] = t->bytes; out[2] = t->done_bytes;
  secs[0] = (t->t_start > 0 ? t->t_start : now_s()) - t->t_submit;
  secs[1] = t->t_start > 0 ? ((t->t_end > 0 ? t->t_end : now_s()) - t->t_start) : 0.0;
  return 0;
}

extern "C" int sy_stage_release(sy_stage* s, long id) {
  std::lock_guard<std::mutex> lk(s->mu);
  auto it = s->tickets.find(id);
  if (it == s->tickets.end()) return 3;
  Ticket* t = it->second;
  if (t->state != T_DONE && t->state != T_FAILED) return 1;
  if (t->own_dptr && t->dptr) { if (s->device >= 0) cudaFree(t->dptr); else free(t->dptr); }
  if (t->done_ev) cudaEventDestroy(t->done_ev);
  delete t;
  s->tickets.erase(it);
  return 0;
}

extern "C" void sy_stage_stats(sy_stage* s, unsigned long long* out, double* busy_seconds) {
  std::lock_guard<std::mutex> lk(s->mu);
  out[0] = s->total_bytes.load(); out[1] = s->memcpy_calls.load(); out[2] = s->arena_bytes; out[3] = s->chunk_bytes;
  *busy_seconds = s->busy_seconds;
}
