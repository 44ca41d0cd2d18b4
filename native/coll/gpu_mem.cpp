// GPU transports: symmetric heap creation + cross-process mapping.
//
// Preferred path: CUDA VMM (cuMemCreate, POSIX-fd export, SCM_RIGHTS exchange,
// cuMemMap of every peer's allocation) plus one NVLS multicast object bound
// over all heaps (cuMulticast*), giving a multicast VA for multimem.* PTX.
// Fallback: cudaMalloc + cudaIpc handles (P2P only).
// (SURVEY.md §5.8 "Bootstrap"; hard parts §7.4 items 2-3.)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "internal.h"

#define DRV_FUNCS(X)                                                                          \
  X(cuDeviceGet) X(cuDeviceGetAttribute) X(cuMemGetAllocationGranularity) X(cuMemCreate)      \
  X(cuMemExportToShareableHandle) X(cuMemImportFromShareableHandle) X(cuMemAddressReserve)    \
  X(cuMemMap) X(cuMemSetAccess) X(cuMemUnmap) X(cuMemRelease) X(cuMemAddressFree)             \
  X(cuMulticastCreate) X(cuMulticastAddDevice) X(cuMulticastBindMem)                          \
  X(cuMulticastGetGranularity) X(cuMulticastUnbind) X(cuGetErrorString)

struct Drv {
#define X(n) decltype(&n) n##_ = nullptr;
  DRV_FUNCS(X)
#undef X
  bool ok = false, mc_ok = false;
};

static bool load_drv(Drv& d) {
  bool all = true, mc = true;
#define X(n)                                                                                   \
  {                                                                                            \
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;                                     \
    if (cudaGetDriverEntryPoint(#n, &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||       \
        q != cudaDriverEntryPointSuccess) {                                                    \
      if (strncmp(#n, "cuMulticast", 11) == 0) mc = false; else all = false;                   \
    } else d.n##_ = (decltype(&n))fn;                                                          \
  }
  DRV_FUNCS(X)
#undef X
  cudaGetLastError();
  d.ok = all; d.mc_ok = all && mc;
  return all;
}

struct GpuImpl {
  Drv drv;
  bool vmm = false;
  CUmemGenericAllocationHandle h[SY_MAXR] = {};
  CUdeviceptr va[SY_MAXR] = {};
  bool mapped[SY_MAXR] = {};
  CUmemGenericAllocationHandle mc_h = 0;
  CUdeviceptr mc_va = 0;
  bool mc_bound = false, mc_mapped = false;
  size_t map_bytes = 0;
  void* ipc_ptr[SY_MAXR] = {};
  int cu_dev = 0;
};

#define CU_TRY(call, what)                                                                     \
  do {                                                                                         \
    CUresult _r = (call);                                                                      \
    if (_r != CUDA_SUCCESS) {                                                                  \
      const char* s = nullptr; if (g->drv.cuGetErrorString_) g->drv.cuGetErrorString_(_r, &s); \
      sy_set_error("%s failed: %s (%d)", what, s ? s : "?", (int)_r);                          \
      return SY_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)
#define RT_TRY(call, what)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) { sy_set_error("%s failed: %s", what, cudaGetErrorString(_e)); return SY_ERR_CUDA; } \
  } while (0)

static int agree_all(sy_comm* c, int mine) {  // logical AND across ranks
  int all[SY_MAXR];
  if (hub_allgather(c->hub, &mine, sizeof mine, all) < 0) return 0;
  int ok = 1;
  for (int r = 0; r < c->world; ++r) ok &= (all[r] != 0);
  return ok;
}

static int init_vmm(sy_comm* c, GpuImpl* g, bool want_mc) {
  CUdevice dev; CU_TRY(g->drv.cuDeviceGet_(&dev, c->device), "cuDeviceGet");
  g->cu_dev = dev;
  int fd_ok = 0, mc_supp = 0;
  g->drv.cuDeviceGetAttribute_(&fd_ok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
  if (g->drv.mc_ok) g->drv.cuDeviceGetAttribute_(&mc_supp, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
  if (!agree_all(c, fd_ok)) { sy_set_error("vmm: posix-fd handles unsupported"); return SY_ERR_UNSUPPORTED; }
  bool use_mc = want_mc && c->world > 1 && agree_all(c, mc_supp);

  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  CU_TRY(g->drv.cuMemGetAllocationGranularity_(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "granularity");
  CUmulticastObjectProp mprop = {};
  if (use_mc) {
    mprop.numDevices = (unsigned)c->world;
    mprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mprop.size = c->heap_bytes;
    size_t mg = 0;
    if (g->drv.cuMulticastGetGranularity_(&mg, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > gran) gran = mg;
  }
  if (gran < (2ul << 20)) gran = 2ul << 20;
  size_t bytes = (c->heap_bytes + gran - 1) / gran * gran;
  c->heap_bytes = bytes; g->map_bytes = bytes;

  CU_TRY(g->drv.cuMemCreate_(&g->h[c->rank], bytes, &prop, 0), "cuMemCreate(heap)");
  int myfd = -1;
  CU_TRY(g->drv.cuMemExportToShareableHandle_(&myfd, g->h[c->rank], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export fd");
  int fds[SY_MAXR];
  for (int r = 0; r < SY_MAXR; ++r) fds[r] = -1;
  if (hub_allgather_fd(c->hub, myfd, fds) < 0) return SY_ERR_SYS;
  close(myfd);
  CUmemAccessDesc acc = {};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (int r = 0; r < c->world; ++r) {
    if (r != c->rank)
      CU_TRY(g->drv.cuMemImportFromShareableHandle_(&g->h[r], (void*)(uintptr_t)fds[r], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "import peer heap");
    close(fds[r]);
    CU_TRY(g->drv.cuMemAddressReserve_(&g->va[r], bytes, gran, 0, 0), "reserve VA");
    CU_TRY(g->drv.cuMemMap_(g->va[r], bytes, 0, g->h[r], 0), "cuMemMap(peer)");
    g->mapped[r] = true;
    CU_TRY(g->drv.cuMemSetAccess_(g->va[r], bytes, &acc, 1), "cuMemSetAccess(peer)");
    c->dev.heap[r] = (char*)g->va[r];
  }
  g->vmm = true;

  if (use_mc) {
    // any failure below degrades to P2P instead of failing the communicator
    int ok = 1, mfd = -1;
    mprop.size = bytes;
    if (c->rank == 0) {
      if (g->drv.cuMulticastCreate_(&g->mc_h, &mprop) != CUDA_SUCCESS ||
          g->drv.cuMemExportToShareableHandle_(&mfd, g->mc_h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS)
        ok = 0;
    }
    if (!agree_all(c, ok)) { use_mc = false; }
    if (use_mc) {
      if (hub_bcast_fd(c->hub, &mfd) < 0) return SY_ERR_SYS;
      if (c->rank != 0 &&
          g->drv.cuMemImportFromShareableHandle_(&g->mc_h, (void*)(uintptr_t)mfd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS)
        ok = 0;
      if (mfd >= 0) close(mfd);
      if (ok && g->drv.cuMulticastAddDevice_(g->mc_h, dev) != CUDA_SUCCESS) ok = 0;
      if (!agree_all(c, ok)) use_mc = false;
    }
    if (use_mc) {
      if (g->drv.cuMulticastBindMem_(g->mc_h, 0, g->h[c->rank], 0, bytes, 0) != CUDA_SUCCESS) ok = 0;
      else g->mc_bound = true;
      if (ok && (g->drv.cuMemAddressReserve_(&g->mc_va, bytes, gran, 0, 0) != CUDA_SUCCESS ||
                 g->drv.cuMemMap_(g->mc_va, bytes, 0, g->mc_h, 0) != CUDA_SUCCESS)) ok = 0;
      else if (ok) g->mc_mapped = true;
      if (ok && g->drv.cuMemSetAccess_(g->mc_va, bytes, &acc, 1) != CUDA_SUCCESS) ok = 0;
      if (!agree_all(c, ok)) use_mc = false;
    }
    if (use_mc) { c->dev.mc = (char*)g->mc_va; c->has_mc = true; }
    else if (getenv("SHIPYARD_COLL_DEBUG")) fprintf(stderr, "[shipyard-coll] NVLS multicast unavailable, using P2P\n");
  }
  return SY_OK;
}

static int init_ipc(sy_comm* c, GpuImpl* g) {
  // legacy IPC: peers must enable access to each other's devices
  int devs[SY_MAXR];
  if (hub_allgather(c->hub, &c->device, sizeof(int), devs) < 0) return SY_ERR_SYS;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank || devs[r] == c->device) continue;
    int can = 0; cudaDeviceCanAccessPeer(&can, c->device, devs[r]);
    if (!can) { sy_set_error("ipc: device %d cannot access peer device %d", c->device, devs[r]); return SY_ERR_UNSUPPORTED; }
    cudaError_t e = cudaDeviceEnablePeerAccess(devs[r], 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { sy_set_error("ipc: enable peer access: %s", cudaGetErrorString(e)); return SY_ERR_CUDA; }
    cudaGetLastError();
  }
  size_t bytes = (c->heap_bytes + (2ul << 20) - 1) / (2ul << 20) * (2ul << 20);
  c->heap_bytes = bytes;
  RT_TRY(cudaMalloc(&g->ipc_ptr[c->rank], bytes), "cudaMalloc(heap)");
  cudaIpcMemHandle_t mine, all[SY_MAXR];
  RT_TRY(cudaIpcGetMemHandle(&mine, g->ipc_ptr[c->rank]), "cudaIpcGetMemHandle");
  if (hub_allgather(c->hub, &mine, sizeof mine, all) < 0) return SY_ERR_SYS;
  for (int r = 0; r < c->world; ++r) {
    if (r != c->rank)
      RT_TRY(cudaIpcOpenMemHandle(&g->ipc_ptr[r], all[r], cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    c->dev.heap[r] = (char*)g->ipc_ptr[r];
  }
  return SY_OK;
}

int gpu_init(sy_comm* c, int requested) {
  GpuImpl* g = new GpuImpl();
  c->impl = g;
  RT_TRY(cudaSetDevice(c->device), "cudaSetDevice");
  RT_TRY(cudaFree(0), "context init");
  const char* mem = getenv("SHIPYARD_COLL_MEM");
  bool want_vmm = !(mem && strcmp(mem, "ipc") == 0);
  bool want_mc = requested != SY_TRANSPORT_P2P;
  int rc = SY_ERR_UNSUPPORTED;
  int have = want_vmm && load_drv(g->drv);
  if (agree_all(c, have)) {
    rc = init_vmm(c, g, want_mc);
    if (!agree_all(c, rc == SY_OK)) rc = rc == SY_OK ? SY_ERR_UNSUPPORTED : rc;
  }
  if (rc != SY_OK) {
    if (want_vmm && getenv("SHIPYARD_COLL_DEBUG")) fprintf(stderr, "[shipyard-coll] VMM path failed (%s); trying cudaIpc\n", sy_last_error());
    if (g->vmm) { sy_set_error("vmm partially initialised; cannot fall back"); return rc; }
    rc = init_ipc(c, g);
    if (!agree_all(c, rc == SY_OK)) return rc == SY_OK ? SY_ERR_UNSUPPORTED : rc;
  }
  if (requested == SY_TRANSPORT_NVLS && !c->has_mc) {
    sy_set_error("NVLS transport requested but multicast is unavailable on this box");
    return SY_ERR_UNSUPPORTED;
  }
  c->transport = c->has_mc ? SY_TRANSPORT_NVLS : SY_TRANSPORT_P2P;

  // control region + counters
  RT_TRY(cudaMemset(c->dev.heap[c->rank], 0, SY_USER_OFF), "zero control region");
  RT_TRY(cudaMalloc(&c->dev.epoch, SY_MAX_BLOCKS * sizeof(uint32_t)), "cudaMalloc(epoch)");
  RT_TRY(cudaMemset(c->dev.epoch, 0, SY_MAX_BLOCKS * sizeof(uint32_t)), "zero epoch");
  RT_TRY(cudaMalloc(&c->dev.seq, SY_SEQ_WORDS * sizeof(uint32_t)), "cudaMalloc(seq)");
  RT_TRY(cudaMemset(c->dev.seq, 0, SY_SEQ_WORDS * sizeof(uint32_t)), "zero seq");
  RT_TRY(cudaHostAlloc(&c->status_host, sizeof(uint32_t), cudaHostAllocMapped), "status word");
  *c->status_host = 0;
  RT_TRY(cudaHostGetDevicePointer((void**)&c->dev.status, c->status_host, 0), "status devptr");
  RT_TRY(cudaDeviceSynchronize(), "sync after init");
  if (hub_barrier(c->hub) < 0) return SY_ERR_SYS;
  return SY_OK;
}

void gpu_destroy(sy_comm* c) {
  GpuImpl* g = (GpuImpl*)c->impl;
  if (!g) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->hub) hub_barrier(c->hub);  // nobody unmaps while a peer may still be touching it
  if (g->vmm) {
    if (g->mc_mapped) { g->drv.cuMemUnmap_(g->mc_va, g->map_bytes); g->drv.cuMemAddressFree_(g->mc_va, g->map_bytes); }
    if (g->mc_bound) g->drv.cuMulticastUnbind_(g->mc_h, g->cu_dev, 0, g->map_bytes);
    if (g->mc_h) g->drv.cuMemRelease_(g->mc_h);
    for (int r = 0; r < c->world; ++r) {
      if (g->mapped[r]) { g->drv.cuMemUnmap_(g->va[r], g->map_bytes); g->drv.cuMemAddressFree_(g->va[r], g->map_bytes); }
      if (g->h[r]) g->drv.cuMemRelease_(g->h[r]);
    }
  } else {
    for (int r = 0; r < c->world; ++r) {
      if (!g->ipc_ptr[r]) continue;
      if (r == c->rank) continue;
      cudaIpcCloseMemHandle(g->ipc_ptr[r]);
    }
    if (c->hub) hub_barrier(c->hub);
    if (g->ipc_ptr[c->rank]) cudaFree(g->ipc_ptr[c->rank]);
  }
  if (c->dev.epoch) cudaFree(c->dev.epoch);
  if (c->dev.seq) cudaFree(c->dev.seq);
  if (c->status_host) cudaFreeHost(c->status_host);
  c->status_host = nullptr;
  delete g; c->impl = nullptr;
}
