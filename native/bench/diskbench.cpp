// shipyard-diskbench: storage micro-benchmark for the box's scratch / shared file systems.
//
// The reference's DiskSpd recipe (/root/reference/recipes/DiskSpd-Windows/config/jobs.yaml: `-c8192k -d1 testfile.dat`) runs
// Microsoft's DiskSpd from a Windows container.  On the B200 box the question that tool answers is the one the input pipeline
// cares about: how fast can the NVMe scratch (or the NFS / Gluster mount of `fs cluster add`) feed the pinned staging buffers of
// libshipyard_stage.  This is a small native equivalent with the same flag spelling for the options the recipe uses:
//
//   shipyard-diskbench [-c<size>] [-d<seconds>] [-b<block>] [-t<threads>] [-w<percent>] [-r] [-S] [-W<seconds>] [-j] [-k] <file>
//
//   -c<size>   create (or grow) the file to <size> bytes first; suffix k / m / g (powers of 1024).  Default: use the file as is.
//   -d<sec>    measured duration (default 10), -W<sec> warm-up before it (default 0)
//   -b<size>   block size per I/O (default 64k), -t<n> threads (default 1), each with its own fd and its own region of the file
//   -w<pct>    percentage of writes (default 0 = read only), -r random offsets (default sequential, wrapping inside the region)
//   -S         O_DIRECT (bypass the page cache); buffers are 4 KB aligned either way
//   -j         one JSON object on stdout instead of the table, -k keep a file created by -c (default: delete it at exit)
//
// Per-I/O latency goes into a log2 histogram (1 us .. ~1 s buckets, 4 sub-buckets each) so percentiles cost no allocation.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

using Clock = std::chrono::steady_clock;

struct Options {
  uint64_t create = 0;
  double duration = 10.0, warmup = 0.0;
  uint64_t block = 64 * 1024;
  int threads = 1;
  int write_pct = 0;
  bool random = false, direct = false, json = false, keep = false;
  std::string path;
};

bool parse_size(const char* s, uint64_t* out) {
  char* end = nullptr;
  errno = 0;
  double v = strtod(s, &end);
  if (end == s || errno || v < 0) return false;
  uint64_t mul = 1;
  if (*end == 'k' || *end == 'K') mul = 1ull << 10, ++end;
  else if (*end == 'm' || *end == 'M') mul = 1ull << 20, ++end;
  else if (*end == 'g' || *end == 'G') mul = 1ull << 30, ++end;
  if (*end == 'b' || *end == 'B') ++end;
  if (*end) return false;
  *out = (uint64_t)(v * (double)mul);
  return true;
}

constexpr int kSub = 4;                       // sub-buckets per power of two
constexpr int kBuckets = 32 * kSub;

inline int bucket_of(uint64_t ns) {
  uint64_t us4 = ns * kSub / 1000;            // quarter-microseconds keep the first buckets distinct
  if (us4 < kSub) return (int)us4;
  int lg = 63 - __builtin_clzll(us4);         // us4 in [2^lg, 2^(lg+1))
  int sub = (int)((us4 >> (lg - 2)) & 3);
  int b = (lg - 1) * kSub + sub;              // lg = 2 -> buckets 4..7
  return b < kBuckets ? b : kBuckets - 1;
}

inline double bucket_upper_us(int b) {
  if (b < kSub) return (double)(b + 1) / kSub;
  int lg = b / kSub + 1, sub = b % kSub;
  return (double)((1ull << lg) + (uint64_t)(sub + 1) * (1ull << (lg - 2))) / kSub;
}

struct Worker {
  uint64_t ios = 0, bytes = 0, read_ios = 0, write_ios = 0, errors = 0;
  uint64_t hist[kBuckets] = {};
  uint64_t max_ns = 0;
  int err = 0;
};

struct Xorshift {
  uint64_t s;
  explicit Xorshift(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 1) {}
  uint64_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
};

std::atomic<int> g_phase{0};                  // 0 warm-up, 1 measured, 2 stop

void run_worker(const Options& o, int tid, uint64_t region_off, uint64_t region_blocks, Worker* w) {
  int flags = (o.write_pct > 0 ? O_RDWR : O_RDONLY) | (o.direct ? O_DIRECT : 0);
  int fd = open(o.path.c_str(), flags);
  if (fd < 0) { w->err = errno; return; }
  void* buf = nullptr;
  if (posix_memalign(&buf, 4096, o.block)) { w->err = ENOMEM; close(fd); return; }
  memset(buf, 0x5a + tid, o.block);
  Xorshift rng(1234567 + tid);
  uint64_t seq = 0;
  while (true) {
    int phase = g_phase.load(std::memory_order_relaxed);
    if (phase == 2) break;
    uint64_t blk = o.random ? rng.next() % region_blocks : (seq++ % region_blocks);
    off_t off = (off_t)(region_off + blk * o.block);
    bool wr = o.write_pct >= 100 || (o.write_pct > 0 && (int)(rng.next() % 100) < o.write_pct);
    auto t0 = Clock::now();
    ssize_t n = wr ? pwrite(fd, buf, o.block, off) : pread(fd, buf, o.block, off);
    auto t1 = Clock::now();
    if (phase != 1) continue;
    if (n != (ssize_t)o.block) { ++w->errors; if (!w->err) w->err = n < 0 ? errno : EIO; if (w->errors > 16) break; continue; }
    uint64_t ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
    ++w->ios; w->bytes += o.block; (wr ? w->write_ios : w->read_ios)++;
    ++w->hist[bucket_of(ns)];
    w->max_ns = std::max(w->max_ns, ns);
  }
  free(buf);
  close(fd);
}

double percentile_us(const uint64_t* hist, uint64_t total, double p) {
  if (!total) return 0.0;
  uint64_t want = (uint64_t)(p * (double)total + 0.5), acc = 0;
  if (want < 1) want = 1;
  for (int b = 0; b < kBuckets; ++b) { acc += hist[b]; if (acc >= want) return bucket_upper_us(b); }
  return bucket_upper_us(kBuckets - 1);
}

int usage() {
  fprintf(stderr, "usage: shipyard-diskbench [-c<size>] [-d<sec>] [-W<sec>] [-b<block>] [-t<threads>] [-w<pct>] [-r] [-S] [-j] [-k] <file>\n");
  return 2;
}

}  // namespace

int main(int argc, char** argv) {
  Options o;
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    if (a[0] != '-' || !a[1]) {
      if (!o.path.empty()) return usage();
      o.path = a;
      continue;
    }
    const char* v = a + 2;
    uint64_t sz = 0;
    switch (a[1]) {
      case 'c': if (!parse_size(v, &sz) || !sz) return usage(); o.create = sz; break;
      case 'b': if (!parse_size(v, &sz) || !sz) return usage(); o.block = sz; break;
      case 'd': o.duration = atof(v); if (o.duration <= 0) return usage(); break;
      case 'W': o.warmup = atof(v); if (o.warmup < 0) return usage(); break;
      case 't': o.threads = atoi(v); if (o.threads < 1 || o.threads > 1024) return usage(); break;
      case 'w': o.write_pct = *v ? atoi(v) : 100; if (o.write_pct < 0 || o.write_pct > 100) return usage(); break;
      case 'r': o.random = true; break;
      case 'S': o.direct = true; break;
      case 'j': o.json = true; break;
      case 'k': o.keep = true; break;
      case 'h': usage(); return 0;
      default: return usage();
    }
  }
  if (o.path.empty()) return usage();
  if (o.direct && o.block % 4096) { fprintf(stderr, "diskbench: -S needs a block size that is a multiple of 4096\n"); return 2; }

  bool created = false;
  if (o.create) {
    struct stat st;
    bool existed = stat(o.path.c_str(), &st) == 0;
    int fd = open(o.path.c_str(), O_RDWR | O_CREAT, 0644);
    if (fd < 0) { fprintf(stderr, "diskbench: cannot create %s: %s\n", o.path.c_str(), strerror(errno)); return 1; }
    created = !existed;
    // write real data (not a sparse file): reads of holes never touch the device
    std::vector<char> chunk(1 << 20, 0x33);
    uint64_t have = existed ? (uint64_t)st.st_size : 0;
    if (have < o.create) {
      if (lseek(fd, (off_t)have, SEEK_SET) < 0) { perror("lseek"); return 1; }
      for (uint64_t left = o.create - have; left;) {
        size_t n = (size_t)std::min<uint64_t>(left, chunk.size());
        ssize_t wr = write(fd, chunk.data(), n);
        if (wr <= 0) { fprintf(stderr, "diskbench: write failed while creating %s: %s\n", o.path.c_str(), strerror(errno)); close(fd); return 1; }
        left -= (uint64_t)wr;
      }
      fsync(fd);
    }
    close(fd);
  }
  struct stat st;
  if (stat(o.path.c_str(), &st) != 0) { fprintf(stderr, "diskbench: %s: %s (use -c<size> to create it)\n", o.path.c_str(), strerror(errno)); return 1; }
  uint64_t fsize = (uint64_t)st.st_size;
  uint64_t blocks = fsize / o.block;
  if (blocks < (uint64_t)o.threads) {
    fprintf(stderr, "diskbench: file of %llu bytes holds %llu blocks of %llu bytes; need at least one per thread (%d)\n",
            (unsigned long long)fsize, (unsigned long long)blocks, (unsigned long long)o.block, o.threads);
    if (created && !o.keep) unlink(o.path.c_str());
    return 1;
  }

  std::vector<Worker> ws(o.threads);
  std::vector<std::thread> th;
  uint64_t per = blocks / o.threads;
  g_phase.store(o.warmup > 0 ? 0 : 1);
  auto t_start = Clock::now();
  for (int t = 0; t < o.threads; ++t) th.emplace_back(run_worker, std::cref(o), t, (uint64_t)t * per * o.block, per, &ws[t]);
  if (o.warmup > 0) {
    std::this_thread::sleep_for(std::chrono::duration<double>(o.warmup));
    t_start = Clock::now();
    g_phase.store(1);
  }
  std::this_thread::sleep_for(std::chrono::duration<double>(o.duration));
  g_phase.store(2);
  auto t_end = Clock::now();
  for (auto& t : th) t.join();
  double secs = std::chrono::duration<double>(t_end - t_start).count();

  Worker tot;
  for (auto& w : ws) {
    tot.ios += w.ios; tot.bytes += w.bytes; tot.read_ios += w.read_ios; tot.write_ios += w.write_ios; tot.errors += w.errors;
    tot.max_ns = std::max(tot.max_ns, w.max_ns);
    if (w.err && !tot.err) tot.err = w.err;
    for (int b = 0; b < kBuckets; ++b) tot.hist[b] += w.hist[b];
  }
  if (created && !o.keep) unlink(o.path.c_str());
  if (tot.err && !tot.ios) { fprintf(stderr, "diskbench: I/O failed: %s%s\n", strerror(tot.err), o.direct && tot.err == EINVAL ? " (file system without O_DIRECT support? drop -S)" : ""); return 1; }

  double mbps = (double)tot.bytes / secs / 1e6, iops = (double)tot.ios / secs;
  double p50 = percentile_us(tot.hist, tot.ios, 0.50), p95 = percentile_us(tot.hist, tot.ios, 0.95), p99 = percentile_us(tot.hist, tot.ios, 0.99);
  if (o.json) {
    printf("{\"file\": \"%s\", \"file_bytes\": %llu, \"block_bytes\": %llu, \"threads\": %d, \"write_pct\": %d, \"pattern\": \"%s\", "
           "\"direct\": %s, \"seconds\": %.3f, \"ios\": %llu, \"read_ios\": %llu, \"write_ios\": %llu, \"errors\": %llu, \"bytes\": %llu, "
           "\"mb_per_s\": %.2f, \"iops\": %.1f, \"lat_us\": {\"p50\": %.2f, \"p95\": %.2f, \"p99\": %.2f, \"max\": %.2f}}\n",
           o.path.c_str(), (unsigned long long)fsize, (unsigned long long)o.block, o.threads, o.write_pct, o.random ? "random" : "sequential",
           o.direct ? "true" : "false", secs, (unsigned long long)tot.ios, (unsigned long long)tot.read_ios, (unsigned long long)tot.write_ios,
           (unsigned long long)tot.errors, (unsigned long long)tot.bytes, mbps, iops, p50, p95, p99, (double)tot.max_ns / 1e3);
  } else {
    printf("file %s (%llu bytes)  block %llu  threads %d  %s  write %d%%  %s\n", o.path.c_str(), (unsigned long long)fsize,
           (unsigned long long)o.block, o.threads, o.random ? "random" : "sequential", o.write_pct, o.direct ? "O_DIRECT" : "buffered");
    printf("%-8s %14s %12s %12s %10s %10s %10s %10s\n", "thread", "bytes", "I/Os", "MB/s", "IOPS", "", "", "");
    for (int t = 0; t < o.threads; ++t)
      printf("%-8d %14llu %12llu %12.2f %10.1f\n", t, (unsigned long long)ws[t].bytes, (unsigned long long)ws[t].ios,
             (double)ws[t].bytes / secs / 1e6, (double)ws[t].ios / secs);
    printf("%-8s %14llu %12llu %12.2f %10.1f\n", "total", (unsigned long long)tot.bytes, (unsigned long long)tot.ios, mbps, iops);
    printf("latency us: p50 %.2f  p95 %.2f  p99 %.2f  max %.2f   errors %llu\n", p50, p95, p99, (double)tot.max_ns / 1e3,
           (unsigned long long)tot.errors);
  }
  return tot.errors ? 1 : 0;
}
