// shipyard-mpibench — collective micro-benchmark in the style of LLNL mpiBench / OSU.
//
// Re-authored from the public description of mpiBench (the reference only downloads and runs
// it: /root/reference/recipes/mpiBench-OpenMPI/docker/Dockerfile:21-26, config/docker/jobs.yaml:6):
// for each operation, the message size doubles from -b to -e; each size is timed over -i
// iterations after warm-up and the average / minimum / maximum time over ranks is printed.
// Extension: -d runs on device buffers (cudaMalloc), which routes the MPI_* calls to the
// sm_100a collective kernels through libshipyard_mpi.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <mpi.h>
#include <nccl.h>
#include <algorithm>
#include "sy_coll.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <string>
#include <vector>

static size_t parse_size(const char* s) {
  char* end; double v = strtod(s, &end);
  if (*end == 'K' || *end == 'k') v *= 1024; else if (*end == 'M' || *end == 'm') v *= 1024 * 1024; else if (*end == 'G' || *end == 'g') v *= 1024.0 * 1024 * 1024;
  return (size_t)v;
}


// ---------------------------------------------------------------------------------------------------------------------------------
// --compare: every operation and size on the shipyard kernels AND on the real NCCL library, in this one process, with one harness:
// device buffers, >= 20 warm-up calls, >= 20 individually timed calls (CUDA events on the collective's own stream, an L2 flush and a
// cross-rank barrier before each, so ranks enter together and no call finds its input in L2), the MEDIAN per rank and the MAX over
// ranks.  NCCL is dlopen'ed (the library PyTorch bundles when present, else the system one) and called at the C level; its
// communicator is bootstrapped with the unique id broadcast through MPI.  One JSON line per (op, size).
struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string path;
  bool load() {
    std::vector<std::string> cands;
    if (const char* e = getenv("SHIPYARD_NCCL_LIB")) cands.push_back(e);
    if (const char* py = getenv("SHIPYARD_PYTHON")) {           // <venv>/bin/python -> <venv>/lib/python*/site-packages/nvidia/nccl/lib
      std::string p = py; size_t s = p.rfind("/bin/");
      if (s != std::string::npos) {
        const std::string cmd = "ls " + p.substr(0, s) + "/lib/python*/site-packages/nvidia/nccl/lib/libnccl.so.2 2>/dev/null | head -1";
        if (FILE* f = popen(cmd.c_str(), "r")) { char b[1024]; if (fgets(b, sizeof b, f)) { std::string l = b; while (!l.empty() && (l.back() == '\n' || l.back() == ' ')) l.pop_back(); if (!l.empty()) cands.push_back(l); } pclose(f); }
      }
    }
    cands.push_back("libnccl.so.2"); cands.push_back("libnccl.so");
    for (auto& c : cands) { h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL); if (h) { path = c; break; } }
    if (!h) return false;
#define SYM(n) n = (decltype(n))dlsym(h, "nccl" #n); if (!n) return false
    SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(AllReduce); SYM(AllGather); SYM(ReduceScatter); SYM(Broadcast);
    SYM(Send); SYM(Recv); SYM(GroupStart); SYM(GroupEnd); SYM(GetVersion);
#undef SYM
    return true;
  }
};

static int run_compare(size_t beg, size_t end, int factor, int iters, int warm, const std::vector<std::string>& ops_in, int rank, int world) {
  std::vector<std::string> ops = ops_in;
  if (ops.empty()) ops = {"allreduce", "allgather", "reduce_scatter", "alltoall", "broadcast", "allreduce_fp8"};
  sy_comm* sc = (sy_comm*)MPIX_Shipyard_comm();
  cudaStream_t st = (cudaStream_t)MPIX_Device_stream();
  MPIX_Set_device_async(1);
  Nccl nc; ncclComm_t ncomm = nullptr; int nver = 0;
  const bool have_nccl = nc.load();
  if (have_nccl) {
    ncclUniqueId id; memset(&id, 0, sizeof id);
    if (rank == 0) nc.GetUniqueId(&id);
    MPI_Bcast(&id, (int)sizeof id, MPI_BYTE, 0, MPI_COMM_WORLD);          // host buffer: the shared-memory face
    if (nc.CommInitRank(&ncomm, world, id, rank) != ncclSuccess) { fprintf(stderr, "mpibench: ncclCommInitRank failed\n"); ncomm = nullptr; }
    nc.GetVersion(&nver);
  }
  const size_t maxb = end;                                               // total bytes of the per-rank buffer (NCCL-tests convention)
  // symmetric buffers (zero-copy path of the shipyard kernels) and plain cudaMalloc buffers (what an unmodified program passes)
  char* sym_in = (char*)MPIX_Sym_alloc(maxb + 256); char* sym_out = (char*)MPIX_Sym_alloc(maxb + 256);
  char *pl_in = nullptr, *pl_out = nullptr, *flush = nullptr; const size_t flush_bytes = 256ul << 20;
  cudaMalloc(&pl_in, maxb + 256); cudaMalloc(&pl_out, maxb + 256); cudaMalloc(&flush, flush_bytes);
  // block-scaled fp8 output lives in the symmetric heap too (the kernel multicasts the quantised result): carved out of sym_out
  void* q8 = sym_out; void* q8s = sym_out + ((maxb / 2 + 4095) & ~(size_t)4095);
  if (!sym_in || !sym_out || !pl_in || !pl_out || !flush) { fprintf(stderr, "mpibench: buffer allocation failed (raise SHIPYARD_COLL_HEAP for --compare up to %zu bytes)\n", maxb); return 2; }
  cudaMemsetAsync(sym_in, 0, maxb, st); cudaMemsetAsync(pl_in, 0, maxb, st); cudaStreamSynchronize(st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  char tr[16]; MPIX_Query_shipyard_transport(tr, sizeof tr);
  if (rank == 0) printf("{\"bench\": \"shipyard-mpibench --compare\", \"world\": %d, \"transport\": \"%s\", \"nccl\": \"%s\", \"nccl_version\": %d, "
                        "\"warmup\": %d, \"iters\": %d, \"timing\": \"cuda events per call, L2 flush + barrier before each, median per rank, max over ranks\"}\n",
                        world, tr, have_nccl ? nc.path.c_str() : "unavailable", nver, warm, iters);
  auto measure = [&](const std::function<void()>& fn, size_t bytes) -> double {
    for (int i = 0; i < warm; ++i) fn();
    cudaStreamSynchronize(st);
    std::vector<float> ts;
    for (int i = 0; i < iters; ++i) {
      if (bytes >= (1ul << 20)) cudaMemsetAsync(flush, i & 0xff, flush_bytes, st);     // evict the operands from L2
      sy_barrier(sc, st);                                                             // ranks enter together (device-side flag barrier)
      cudaEventRecord(e0, st); fn(); cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ts.push_back(ms * 1e3f);
    }
    std::sort(ts.begin(), ts.end());
    double med = ts[ts.size() / 2], mx = 0;
    MPI_Allreduce(&med, &mx, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
    return mx;
  };
  bool fp8_warned = false;
  for (auto& op : ops) {
    for (size_t total = beg; total <= end; total *= (size_t)factor) {
      const size_t n = std::max<size_t>((size_t)world, total / 4 / world * world);     // fp32 elements of the per-rank buffer
      const size_t bytes = n * 4, per = n / world;
      double busf = op == "allreduce" || op == "allreduce_fp8" ? 2.0 * (world - 1) / world : op == "broadcast" ? 1.0 : (double)(world - 1) / world;
      std::function<void()> f_sym, f_plain, f_nccl;
      if (op == "allreduce") {
        f_sym = [&] { sy_allreduce(sc, sym_in, sym_out, n, SY_F32, SY_F32, 1.0f, SY_SUM, SY_ALGO_AUTO, st); };
        f_plain = [&] { sy_allreduce(sc, pl_in, pl_out, n, SY_F32, SY_F32, 1.0f, SY_SUM, SY_ALGO_AUTO, st); };
        if (ncomm) f_nccl = [&] { nc.AllReduce(pl_in, pl_out, n, ncclFloat32, ncclSum, ncomm, st); };
      } else if (op == "allreduce_fp8") {
        // block-scaled fp8 output (e4m3 + one e8m0 scale per 32 elements) from bf16 input; NCCL comparator: the bf16 all-reduce it replaces
        f_sym = [&] { if (sy_allreduce_fp8_blockscaled(sc, sym_in, SY_BF16, q8, q8s, n * 2, 1.0f, st) != SY_OK && rank == 0 && !fp8_warned) { fp8_warned = true; fprintf(stderr, "mpibench: allreduce_fp8: %s\n", sy_last_error()); } };
        if (ncomm) f_nccl = [&] { nc.AllReduce(pl_in, pl_out, n * 2, ncclBfloat16, ncclSum, ncomm, st); };
      } else if (op == "allgather") {
        f_sym = [&] { sy_allgather(sc, sym_in, sym_out, per * 4, SY_U8, st); };
        f_plain = [&] { sy_allgather(sc, pl_in, pl_out, per * 4, SY_U8, st); };
        if (ncomm) f_nccl = [&] { nc.AllGather(pl_in, pl_out, per, ncclFloat32, ncomm, st); };
      } else if (op == "reduce_scatter") {
        f_sym = [&] { sy_reduce_scatter(sc, sym_in, sym_out, per, SY_F32, SY_F32, 1.0f, SY_SUM, st); };
        f_plain = [&] { sy_reduce_scatter(sc, pl_in, pl_out, per, SY_F32, SY_F32, 1.0f, SY_SUM, st); };
        if (ncomm) f_nccl = [&] { nc.ReduceScatter(pl_in, pl_out, per, ncclFloat32, ncclSum, ncomm, st); };
      } else if (op == "alltoall") {
        f_sym = [&] { sy_alltoall(sc, sym_in, sym_out, per * 4, SY_U8, st); };
        f_plain = [&] { sy_alltoall(sc, pl_in, pl_out, per * 4, SY_U8, st); };
        if (ncomm) f_nccl = [&] {
          nc.GroupStart();
          for (int r = 0; r < world; ++r) { nc.Send(pl_in + (size_t)r * per * 4, per, ncclFloat32, r, ncomm, st); nc.Recv(pl_out + (size_t)r * per * 4, per, ncclFloat32, r, ncomm, st); }
          nc.GroupEnd();
        };
      } else if (op == "broadcast") {
        f_sym = [&] { sy_broadcast(sc, sym_in, sym_in, bytes, SY_U8, 0, st); };
        f_plain = [&] { sy_broadcast(sc, pl_in, pl_in, bytes, SY_U8, 0, st); };
        if (ncomm) f_nccl = [&] { nc.Broadcast(pl_in, pl_in, n, ncclFloat32, 0, ncomm, st); };
      } else { if (rank == 0) fprintf(stderr, "mpibench: unknown --compare op %s\n", op.c_str()); break; }
      const double t_sym = f_sym ? measure(f_sym, bytes) : -1, t_plain = f_plain ? measure(f_plain, bytes) : -1, t_nccl = f_nccl ? measure(f_nccl, bytes) : -1;
      MPIX_Device_sync();
      if (rank == 0) {
        const double best = t_plain > 0 && t_plain < t_sym ? t_plain : t_sym;
        printf("{\"op\": \"%s\", \"bytes\": %zu, \"world\": %d, \"sy_sym_us\": %.2f, \"sy_plain_us\": %.2f, \"nccl_us\": %.2f, \"speedup_sym_vs_nccl\": %.3f, "
               "\"speedup_plain_vs_nccl\": %.3f, \"sy_busbw_gbs\": %.1f, \"nccl_busbw_gbs\": %.1f, \"frac_of_770gbs\": %.3f}\n",
               op.c_str(), bytes, world, t_sym, t_plain, t_nccl, t_nccl > 0 ? t_nccl / t_sym : 0.0, t_nccl > 0 && t_plain > 0 ? t_nccl / t_plain : 0.0,
               bytes * busf / best / 1e3, t_nccl > 0 ? bytes * busf / t_nccl / 1e3 : 0.0, bytes * busf / best / 1e3 / 770.0);
        fflush(stdout);
      }
    }
  }
  MPIX_Device_sync();
  if (ncomm) nc.CommDestroy(ncomm);
  MPIX_Set_device_async(0);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// --osu <benchmark> [OSU flags]: the OSU micro-benchmark front end for the collective latency tests the reference recipe launches
// (`collective/osu_allreduce -f`, /root/reference/recipes/OSUMicroBenchmarks-Infiniband-MVAPICH/config/jobs.yaml:5-16): same flags
// (-f full statistics, -m [min:]max message sizes in bytes, -i iterations, -x warm-up iterations, -d cuda = device buffers), same
// columns (size in bytes, average / minimum / maximum latency over the ranks in microseconds, iterations), host-timed with MPI_Wtime
// around each call like the original.
// ---------------------------------------------------------------------------------------------------------------------------------
static int run_osu(int argc, char** argv, const char* bench) {
  std::string name = bench;
  const size_t slash = name.rfind('/');
  if (slash != std::string::npos) name = name.substr(slash + 1);
  struct { const char* osu; const char* title; } known[] = {
    {"osu_allreduce", "Allreduce"}, {"osu_reduce", "Reduce"}, {"osu_bcast", "Broadcast"}, {"osu_allgather", "Allgather"},
    {"osu_alltoall", "All-to-All Personalized Exchange"}, {"osu_gather", "Gather"}, {"osu_scatter", "Scatter"}, {"osu_barrier", "Barrier"}};
  const char* title = nullptr;
  for (auto& k : known) if (name == k.osu) title = k.title;
  bool full = false, device = false; size_t mn = 4, mx = 1 << 20; int iters = 1000, warm = 200; bool iters_set = false;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-f")) full = true;
    else if (!strcmp(argv[i], "-d") && i + 1 < argc) device = !strcmp(argv[++i], "cuda") || !strcmp(argv[i], "managed");
    else if (!strcmp(argv[i], "-i") && i + 1 < argc) { iters = atoi(argv[++i]); iters_set = true; }
    else if (!strcmp(argv[i], "-x") && i + 1 < argc) warm = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-m") && i + 1 < argc) {
      std::string m = argv[++i]; const size_t c = m.find(':');
      if (c == std::string::npos) mx = parse_size(m.c_str()); else { if (c > 0) mn = parse_size(m.substr(0, c).c_str()); mx = parse_size(m.substr(c + 1).c_str()); }
    }
  }
  MPI_Init(&argc, &argv);
  int rank, world;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank); MPI_Comm_size(MPI_COMM_WORLD, &world);
  if (name == "osu_latency" || name == "osu_bw") {
    // point-to-point tests between ranks 0 and 1 (host buffers: the MPI face's Send / Recv path): ping-pong latency (half round trip) and
    // windowed bandwidth (64 non-blocking sends per iteration, then a 4-byte acknowledgement), as in the OSU originals
    if (world < 2) { if (rank == 0) fprintf(stderr, "mpibench --osu %s needs 2 ranks\n", name.c_str()); MPI_Finalize(); return 3; }
    if (device) { if (rank == 0) fprintf(stderr, "mpibench --osu %s: host buffers only\n", name.c_str()); MPI_Finalize(); return 3; }
    const bool bw = name == "osu_bw";
    if (mx == (size_t)(1 << 20) && bw) mx = 4 << 20;
    if (mn <= 4) mn = 1;                                  // OSU's point-to-point tests start at one byte
    const int window = 64;
    char* sb = (char*)calloc(1, mx + 64); char* rb = (char*)calloc(1, mx + 64);
    if (rank == 0) printf("# OSU MPI %s Test (shipyard-mpibench over libshipyard_mpi, ranks 0 <-> 1, host buffers)\n# %-8s %18s\n",
                          bw ? "Bandwidth" : "Latency", "Size", bw ? "Bandwidth (MB/s)" : "Latency (us)");
    for (size_t bytes = mn; bytes <= mx; bytes *= 2) {
      int it_n = iters_set ? iters : (bytes > 8192 ? (bw ? 20 : 100) : (bw ? 100 : 1000));
      int wm_n = bytes > 8192 ? 2 : (warm < 10 ? warm : 10);
      MPI_Barrier(MPI_COMM_WORLD);
      double t0 = 0;
      if (rank < 2) {
        std::vector<MPI_Request> reqs(window);
        for (int it = -wm_n; it < it_n; ++it) {
          if (it == 0) t0 = MPI_Wtime();
          if (!bw) {
            if (rank == 0) { MPI_Send(sb, (int)bytes, MPI_CHAR, 1, 1, MPI_COMM_WORLD); MPI_Recv(rb, (int)bytes, MPI_CHAR, 1, 1, MPI_COMM_WORLD, MPI_STATUS_IGNORE); }
            else { MPI_Recv(rb, (int)bytes, MPI_CHAR, 0, 1, MPI_COMM_WORLD, MPI_STATUS_IGNORE); MPI_Send(sb, (int)bytes, MPI_CHAR, 0, 1, MPI_COMM_WORLD); }
          } else if (rank == 0) {
            for (int w = 0; w < window; ++w) MPI_Isend(sb, (int)bytes, MPI_CHAR, 1, 100, MPI_COMM_WORLD, &reqs[w]);
            MPI_Waitall(window, reqs.data(), MPI_STATUSES_IGNORE);
            MPI_Recv(rb, 4, MPI_CHAR, 1, 101, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
          } else {
            for (int w = 0; w < window; ++w) MPI_Irecv(rb, (int)bytes, MPI_CHAR, 0, 100, MPI_COMM_WORLD, &reqs[w]);
            MPI_Waitall(window, reqs.data(), MPI_STATUSES_IGNORE);
            MPI_Send(sb, 4, MPI_CHAR, 0, 101, MPI_COMM_WORLD);
          }
        }
        const double el = MPI_Wtime() - t0;
        if (rank == 0) {
          if (bw) printf("%-10zu %18.2f\n", bytes, (double)bytes * window * it_n / el / 1e6);
          else printf("%-10zu %18.2f\n", bytes, el * 1e6 / (2.0 * it_n));
          fflush(stdout);
        }
      }
    }
    MPI_Barrier(MPI_COMM_WORLD);
    free(sb); free(rb);
    MPI_Finalize();
    return 0;
  }
  if (!title) { if (rank == 0) fprintf(stderr, "mpibench --osu: unknown benchmark %s (collective latency tests, osu_latency, osu_bw)\n", name.c_str()); MPI_Finalize(); return 2; }
  if (device) {
    const char* g = getenv("SHIPYARD_GPU"); int ndev = 0; cudaGetDeviceCount(&ndev);
    if (ndev == 0) { if (rank == 0) fprintf(stderr, "mpibench --osu: -d cuda requested but no GPU visible\n"); MPI_Finalize(); return 3; }
    cudaSetDevice(g ? atoi(g) % ndev : rank % ndev);
  }
  if (mn < 4) mn = 4;
  const size_t maxb = mx * (size_t)world;
  void *sbuf = nullptr, *rbuf = nullptr;
  if (device) { cudaMalloc(&sbuf, maxb + 64); cudaMalloc(&rbuf, maxb + 64); cudaMemset(sbuf, 0, maxb + 64); cudaMemset(rbuf, 0, maxb + 64); }
  else { sbuf = calloc(1, maxb + 64); rbuf = calloc(1, maxb + 64); }
  char tr[16]; MPIX_Query_shipyard_transport(tr, sizeof tr);
  if (rank == 0) {
    printf("# OSU MPI%s %s Latency Test (shipyard-mpibench over libshipyard_mpi, %d ranks, transport %s)\n", device ? "-CUDA" : "", title, world, device ? tr : "host");
    if (full) printf("# %-8s %18s %18s %18s %12s\n", "Size", "Avg Latency(us)", "Min Latency(us)", "Max Latency(us)", "Iterations");
    else printf("# %-8s %18s\n", "Size", "Avg Latency(us)");
  }
  const bool barrier = name == "osu_barrier";
  for (size_t bytes = mn; bytes <= mx; bytes *= 2) {
    const int count = (int)(bytes / 4);
    int it_n = iters, wm_n = warm;
    if (!iters_set && bytes > 8192) { it_n = 100; wm_n = warm < 10 ? warm : 10; }           // OSU's large-message defaults
    auto once = [&]() {
      if (barrier) MPI_Barrier(MPI_COMM_WORLD);
      else if (name == "osu_allreduce") MPI_Allreduce(sbuf, rbuf, count, MPI_FLOAT, MPI_SUM, MPI_COMM_WORLD);
      else if (name == "osu_reduce") MPI_Reduce(sbuf, rbuf, count, MPI_FLOAT, MPI_SUM, 0, MPI_COMM_WORLD);
      else if (name == "osu_bcast") MPI_Bcast(sbuf, count, MPI_FLOAT, 0, MPI_COMM_WORLD);
      else if (name == "osu_allgather") MPI_Allgather(sbuf, count, MPI_FLOAT, rbuf, count, MPI_FLOAT, MPI_COMM_WORLD);
      else if (name == "osu_alltoall") MPI_Alltoall(sbuf, count, MPI_FLOAT, rbuf, count, MPI_FLOAT, MPI_COMM_WORLD);
      else if (name == "osu_gather") MPI_Gather(sbuf, count, MPI_FLOAT, rbuf, count, MPI_FLOAT, 0, MPI_COMM_WORLD);
      else if (name == "osu_scatter") MPI_Scatter(sbuf, count, MPI_FLOAT, rbuf, count, MPI_FLOAT, 0, MPI_COMM_WORLD);
    };
    for (int w = 0; w < wm_n; ++w) once();
    MPI_Barrier(MPI_COMM_WORLD);
    double total = 0;
    for (int it = 0; it < it_n; ++it) { const double t0 = MPI_Wtime(); once(); total += MPI_Wtime() - t0; }
    double us = total * 1e6 / it_n, lo, hi, sum;
    MPI_Allreduce(&us, &lo, 1, MPI_DOUBLE, MPI_MIN, MPI_COMM_WORLD);
    MPI_Allreduce(&us, &hi, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
    MPI_Allreduce(&us, &sum, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    if (rank == 0) {
      if (barrier) { if (full) printf("%-10s %18.2f %18.2f %18.2f %12d\n", "", sum / world, lo, hi, it_n); else printf("%-10s %18.2f\n", "", sum / world); }
      else if (full) printf("%-10zu %18.2f %18.2f %18.2f %12d\n", bytes, sum / world, lo, hi, it_n);
      else printf("%-10zu %18.2f\n", bytes, sum / world);
      fflush(stdout);
    }
    if (barrier) break;
  }
  if (device) { cudaFree(sbuf); cudaFree(rbuf); } else { free(sbuf); free(rbuf); }
  MPI_Finalize();
  return 0;
}

int main(int argc, char** argv) {
  for (int i = 1; i + 1 < argc; ++i)
    if (!strcmp(argv[i], "--osu")) return run_osu(argc, argv, argv[i + 1]);
  size_t beg = 8, end = 1024; int iters = 100; bool device = false; bool check = false;
  bool compare = false; int factor = 2, warm = 20;
  std::vector<std::string> ops;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-b") && i + 1 < argc) beg = parse_size(argv[++i]);
    else if (!strcmp(argv[i], "-e") && i + 1 < argc) end = parse_size(argv[++i]);
    else if (!strcmp(argv[i], "-i") && i + 1 < argc) iters = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-d")) device = true;
    else if (!strcmp(argv[i], "-c")) check = true;
    else if (!strcmp(argv[i], "--compare")) { compare = true; device = true; if (iters == 100) iters = 25; }
    else if (!strcmp(argv[i], "--factor") && i + 1 < argc) factor = atoi(argv[++i]) < 2 ? 2 : atoi(argv[i]);
    else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warm = atoi(argv[++i]);
    else if (argv[i][0] != '-') ops.push_back(argv[i]);
  }
  // the LLNL mpiBench operation set; the vector variants run with uniform counts (host buffers only in this MPI face)
  if (ops.empty() && !compare) {
    ops = {"Barrier", "Bcast", "Alltoall", "Allgather", "Gather", "Scatter", "Allreduce", "Reduce"};
    if (!device) { ops.push_back("Alltoallv"); ops.push_back("Allgatherv"); ops.push_back("Gatherv"); }
  }
  MPI_Init(&argc, &argv);
  int rank, world;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank); MPI_Comm_size(MPI_COMM_WORLD, &world);
  if (device) {
    const char* g = getenv("SHIPYARD_GPU");
    int ndev = 0; cudaGetDeviceCount(&ndev);
    if (ndev == 0) { if (rank == 0) fprintf(stderr, "mpibench: -d requested but no GPU visible\n"); MPI_Finalize(); return 3; }
    cudaSetDevice(g ? atoi(g) % ndev : rank % ndev);
  }
  if (compare) {
    if (world < 2) { if (rank == 0) fprintf(stderr, "mpibench: --compare needs at least 2 ranks\n"); MPI_Finalize(); return 3; }
    const int rc = run_compare(beg < 1024 ? 1024 : beg, end, factor, iters < 20 ? 20 : iters, warm < 20 ? 20 : warm, ops, rank, world);
    MPI_Finalize();
    return rc;
  }
  const size_t maxb = end * (size_t)world;
  void *sbuf = nullptr, *rbuf = nullptr;
  if (device) { cudaMalloc(&sbuf, maxb + 64); cudaMalloc(&rbuf, maxb + 64); cudaMemset(sbuf, 0, maxb + 64); }
  else { sbuf = calloc(1, maxb + 64); rbuf = calloc(1, maxb + 64); }
  std::vector<float> host(end / 4 + 4);
  if (rank == 0) printf("# shipyard-mpibench ranks=%d buffers=%s iters=%d\n# %-12s %12s %8s %12s %12s %12s\n", world,
                        device ? "device" : "host", iters, "op", "bytes", "iters", "avg_us", "min_us", "max_us");
  int failures = 0;
  for (auto& op : ops) {
    for (size_t bytes = beg; bytes <= end; bytes *= 2) {
      if (op == "Barrier" && bytes != beg) break;
      const int count = (int)(bytes / 4 ? bytes / 4 : 1);
      if (check && op == "Allreduce") {
        for (int i = 0; i < count; ++i) host[i] = (float)(rank + 1) + (float)(i % 7);
        if (device) cudaMemcpy(sbuf, host.data(), (size_t)count * 4, cudaMemcpyHostToDevice); else memcpy(sbuf, host.data(), (size_t)count * 4);
      }
      std::vector<int> counts(world, count), displs(world);
      for (int r = 0; r < world; ++r) displs[r] = r * count;
      if (device && (op == "Alltoallv" || op == "Allgatherv" || op == "Gatherv")) {
        if (rank == 0 && bytes == beg) printf("  %-12s skipped (vector collectives take host buffers)\n", op.c_str());
        break;
      }
      auto once = [&]() {
        if (op == "Barrier") MPI_Barrier(MPI_COMM_WORLD);
        else if (op == "Alltoallv") MPI_Alltoallv(sbuf, counts.data(), displs.data(), MPI_FLOAT, rbuf, counts.data(), displs.data(), MPI_FLOAT, MPI_COMM_WORLD);
        else if (op == "Allgatherv") MPI_Allgatherv(sbuf, count, MPI_FLOAT, rbuf, counts.data(), displs.data(), MPI_FLOAT, MPI_COMM_WORLD);
        else if (op == "Gatherv") MPI_Gatherv(sbuf, count, MPI_FLOAT, rbuf, counts.data(), displs.data(), MPI_FLOAT, 0, MPI_COMM_WORLD);
        else if (op == "Bcast") MPI_Bcast(sbuf, count, MPI_FLOAT, 0, MPI_COMM_WORLD);
        else if (op == "Alltoall") MPI_Alltoall(sbuf, count, MPI_FLOAT, rbuf, count, MPI_FLOAT, MPI_COMM_WORLD);
        else if (op == "Allgather") MPI_Allgather(sbuf, count, MPI_FLOAT, rbuf, count, MPI_FLOAT, MPI_COMM_WORLD);
        else if (op == "Gather") MPI_Gather(sbuf, count, MPI_FLOAT, rbuf, count, MPI_FLOAT, 0, MPI_COMM_WORLD);
        else if (op == "Scatter") MPI_Scatter(sbuf, count, MPI_FLOAT, rbuf, count, MPI_FLOAT, 0, MPI_COMM_WORLD);
        else if (op == "Allreduce") MPI_Allreduce(sbuf, rbuf, count, MPI_FLOAT, MPI_SUM, MPI_COMM_WORLD);
        else if (op == "Reduce") MPI_Reduce(sbuf, rbuf, count, MPI_FLOAT, MPI_SUM, 0, MPI_COMM_WORLD);
      };
      for (int w = 0; w < 3; ++w) once();
      if (check && op == "Allreduce") {
        if (device) cudaMemcpy(host.data(), rbuf, (size_t)count * 4, cudaMemcpyDeviceToHost); else memcpy(host.data(), rbuf, (size_t)count * 4);
        for (int i = 0; i < count; ++i) {
          float want = (float)world * (float)(i % 7) + (float)world * (world + 1) / 2.0f;
          if (host[i] != want) { ++failures; if (failures < 4) fprintf(stderr, "rank %d: Allreduce[%d] = %g, want %g\n", rank, i, host[i], want); break; }
        }
      }
      MPI_Barrier(MPI_COMM_WORLD);
      double t0 = MPI_Wtime();
      for (int it = 0; it < iters; ++it) once();
      double us = (MPI_Wtime() - t0) * 1e6 / iters, mn, mx, sum;
      MPI_Allreduce(&us, &mn, 1, MPI_DOUBLE, MPI_MIN, MPI_COMM_WORLD);
      MPI_Allreduce(&us, &mx, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
      MPI_Allreduce(&us, &sum, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
      if (rank == 0) printf("  %-12s %12zu %8d %12.2f %12.2f %12.2f\n", op.c_str(), op == "Barrier" ? (size_t)0 : bytes, iters, sum / world, mn, mx);
    }
  }
  int total_fail = 0;
  MPI_Allreduce(&failures, &total_fail, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  char tr[16]; MPIX_Query_shipyard_transport(tr, sizeof tr);
  if (rank == 0) printf("# device-buffer transport: %s ; check failures: %d\n", tr, total_fail);
  if (device) { cudaFree(sbuf); cudaFree(rbuf); } else { free(sbuf); free(rbuf); }
  MPI_Finalize();
  return total_fail ? 1 : 0;
}
