// ThreadSanitizer driver for the staging library (host mode): several client threads submit, wait, query and release
// tickets concurrently while the library's worker threads move the bytes.  Built with -fsanitize=thread by
// `python native/build.py --sanitizers`; a data race makes the process exit non-zero (TSAN_OPTIONS exitcode).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>

struct sy_stage;
extern "C" {
int sy_stage_create(sy_stage** out, int device, size_t arena_bytes, int concurrency, int chunks_per_worker);
int sy_stage_destroy(sy_stage* s);
long sy_stage_submit_file(sy_stage* s, const char* path, void* dptr, size_t offset, size_t bytes);
long sy_stage_submit_host(sy_stage* s, const void* host, size_t bytes, void* dptr);
int sy_stage_wait(sy_stage* s, long id, double timeout_s);
void* sy_stage_ptr(sy_stage* s, long id);
int sy_stage_query(sy_stage* s, long id, unsigned long long* out, double* secs);
int sy_stage_release(sy_stage* s, long id);
void sy_stage_stats(sy_stage* s, unsigned long long* out, double* busy_seconds);
}

int main(int argc, char** argv) {
  const char* file = argc > 1 ? argv[1] : nullptr;
  sy_stage* st = nullptr;
  if (sy_stage_create(&st, -1, 4 << 20, 4, 2) != 0) { fprintf(stderr, "create failed\n"); return 2; }
  std::atomic<int> failures{0};
  std::vector<std::thread> clients;
  for (int c = 0; c < 4; ++c) {
    clients.emplace_back([&, c] {
      std::vector<unsigned char> src(300000 + 1000 * c), dst(src.size());
      for (int it = 0; it < 25; ++it) {
        for (size_t i = 0; i < src.size(); ++i) src[i] = (unsigned char)(i * 7 + c + it);
        memset(dst.data(), 0, dst.size());
        long id = sy_stage_submit_host(st, src.data(), src.size(), dst.data());
        if (id < 0) { ++failures; continue; }
        long fid = file ? sy_stage_submit_file(st, file, nullptr, 0, 0) : -1;
        if (sy_stage_wait(st, id, 30.0) != 0) ++failures;
        unsigned long long q[4]; double secs[2];
        if (sy_stage_query(st, id, q, secs) != 0) ++failures;
        if (memcmp(src.data(), dst.data(), src.size()) != 0) ++failures;
        sy_stage_release(st, id);
        if (fid >= 0) { if (sy_stage_wait(st, fid, 30.0) != 0) ++failures; sy_stage_release(st, fid); }
      }
    });
  }
  for (auto& t : clients) t.join();
  unsigned long long stats[4]; double busy;
  sy_stage_stats(st, stats, &busy);
  sy_stage_destroy(st);
  printf("tsan_stage: %llu bytes moved, %d failures\n", stats[0], failures.load());
  return failures.load() ? 1 : 0;
}
