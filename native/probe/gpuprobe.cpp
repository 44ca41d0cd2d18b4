// shipyard-gpuprobe — enumerate the box's GPUs and their interconnect as JSON.
//
// The "node" half of pool provisioning (/root/reference/scripts/shipyard_nodeprep.sh:380-422
// detects VM size / RDMA class from IMDS and /dev/infiniband; :626-873 installs and checks the
// NVIDIA driver).  Here the facts come straight from the CUDA runtime + driver API: device
// list, memory, SM count, compute capability, P2P access/atomics matrix, NVLS multicast and
// POSIX-fd (VMM) export support, driver/runtime versions.  NVML (dlopen, optional) adds clocks,
// power limit, persistence mode and the NVLink link count.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

typedef int (*nvml_fn0)();
typedef int (*nvml_get_handle)(unsigned, void**);
typedef int (*nvml_get_uint)(void*, unsigned*);
typedef int (*nvml_get_clock)(void*, int, unsigned*);
typedef int (*nvml_get_mode)(void*, int*);
typedef int (*nvml_nvlink_state)(void*, unsigned, int*);

int main(int argc, char** argv) {
  bool pretty = argc > 1 && !strcmp(argv[1], "--pretty");
  const char* nl = pretty ? "\n" : "";
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  int drv = 0, rt = 0;
  cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
  if (e != cudaSuccess) n = 0;
  void* nvml = dlopen("libnvidia-ml.so.1", RTLD_LAZY);
  nvml_fn0 nvml_init = nvml ? (nvml_fn0)dlsym(nvml, "nvmlInit_v2") : nullptr;
  bool have_nvml = nvml_init && nvml_init() == 0;
  auto sym = [&](const char* s) { return have_nvml ? dlsym(nvml, s) : nullptr; };
  auto h_by_index = (nvml_get_handle)sym("nvmlDeviceGetHandleByIndex_v2");
  auto max_clock = (nvml_get_clock)sym("nvmlDeviceGetMaxClockInfo");
  auto power_limit = (nvml_get_uint)sym("nvmlDeviceGetPowerManagementLimit");
  auto persistence = (nvml_get_mode)sym("nvmlDeviceGetPersistenceMode");
  auto nvlink_state = (nvml_nvlink_state)sym("nvmlDeviceGetNvLinkState");

  printf("{\"driver_version\": %d, \"runtime_version\": %d, \"error\": %s, \"nvml\": %s, \"gpus\": [%s", drv, rt,
         e == cudaSuccess ? "null" : (std::string("\"") + cudaGetErrorString(e) + "\"").c_str(), have_nvml ? "true" : "false", nl);
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, d);
    size_t free_b = 0, total_b = 0;
    cudaSetDevice(d); cudaMemGetInfo(&free_b, &total_b);
    int mc = 0, fdh = 0;
    {
      void* fn = nullptr; cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuDeviceGetAttribute", &fn, cudaEnableDefault, &q) == cudaSuccess && fn) {
        auto get = (CUresult(*)(int*, CUdevice_attribute, CUdevice))fn;
        get(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d);
        get(&fdh, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, d);
      }
    }
    unsigned sm_max = 0, mem_max = 0, plimit = 0; int pm = -1, links = 0;
    void* h = nullptr;
    if (have_nvml && h_by_index && h_by_index((unsigned)d, &h) == 0) {
      if (max_clock) { max_clock(h, 1 /*SM*/, &sm_max); max_clock(h, 2 /*MEM*/, &mem_max); }
      if (power_limit) power_limit(h, &plimit);
      if (persistence) persistence(h, &pm);
      if (nvlink_state) for (unsigned l = 0; l < 18; ++l) { int on = 0; if (nvlink_state(h, l, &on) == 0 && on) ++links; }
    }
    printf("%s{\"index\": %d, \"name\": \"%s\", \"uuid_prefix\": \"%02x%02x%02x%02x\", \"cc\": \"%d.%d\", \"sms\": %d, "
           "\"memory_total\": %zu, \"memory_free\": %zu, \"l2_bytes\": %d, \"pci_bus_id\": %d, \"multicast\": %s, "
           "\"posix_fd_handles\": %s, \"sm_max_mhz\": %u, \"mem_max_mhz\": %u, \"power_limit_mw\": %u, "
           "\"persistence_mode\": %d, \"nvlink_links_active\": %d, \"p2p\": [",
           d ? ", " : "", d, p.name, (unsigned char)p.uuid.bytes[0], (unsigned char)p.uuid.bytes[1],
           (unsigned char)p.uuid.bytes[2], (unsigned char)p.uuid.bytes[3], p.major, p.minor, p.multiProcessorCount,
           total_b, free_b, p.l2CacheSize, p.pciBusID, mc ? "true" : "false", fdh ? "true" : "false", sm_max, mem_max,
           plimit, pm, links);
    for (int q = 0; q < n; ++q) {
      int can = d == q ? 1 : 0, atom = d == q ? 1 : 0;
      if (d != q) { cudaDeviceCanAccessPeer(&can, d, q); cudaDeviceGetP2PAttribute(&atom, cudaDevP2PAttrNativeAtomicSupported, d, q); }
      printf("%s{\"peer\": %d, \"access\": %s, \"atomics\": %s}", q ? ", " : "", q, can ? "true" : "false", atom ? "true" : "false");
    }
    printf("]}%s", nl);
  }
  printf("]}\n");
  return 0;
}
