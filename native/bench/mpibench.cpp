// shipyard-mpibench — collective micro-benchmark in the style of LLNL mpiBench / OSU.
//
// Re-authored from the public description of mpiBench (the reference only downloads and runs
// it: /root/reference/recipes/mpiBench-OpenMPI/docker/Dockerfile:21-26, config/docker/jobs.yaml:6):
// for each operation, the message size doubles from -b to -e; each size is timed over -i
// iterations after warm-up and the average / minimum / maximum time over ranks is printed.
// Extension: -d runs on device buffers (cudaMalloc), which routes the MPI_* calls to the
// sm_100a collective kernels through libshipyard_mpi.
#include <cuda_runtime.h>
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

static size_t parse_size(const char* s) {
  char* end; double v = strtod(s, &end);
  if (*end == 'K' || *end == 'k') v *= 1024; else if (*end == 'M' || *end == 'm') v *= 1024 * 1024; else if (*end == 'G' || *end == 'g') v *= 1024.0 * 1024 * 1024;
  return (size_t)v;
}

int main(int argc, char** argv) {
  size_t beg = 8, end = 1024; int iters = 100; bool device = false; bool check = false;
  std::vector<std::string> ops;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-b") && i + 1 < argc) beg = parse_size(argv[++i]);
    else if (!strcmp(argv[i], "-e") && i + 1 < argc) end = parse_size(argv[++i]);
    else if (!strcmp(argv[i], "-i") && i + 1 < argc) iters = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-d")) device = true;
    else if (!strcmp(argv[i], "-c")) check = true;
    else if (argv[i][0] != '-') ops.push_back(argv[i]);
  }
  // the LLNL mpiBench operation set; the vector variants run with uniform counts (host buffers only in this MPI face)
  if (ops.empty()) {
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
