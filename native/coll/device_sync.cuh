// Device-side cross-GPU synchronisation shared by the collective kernels and the fused
// GEMM + collective kernels: monotonic epoch flags in every rank's symmetric heap, release/acquire
// at system scope, and a watchdog that reports a stuck peer through a mapped host word instead of hanging.
#pragma once
#include <stdint.h>
#include "internal.h"

#ifndef DEVI
#define DEVI __device__ __forceinline__
#endif

DEVI void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
DEVI uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
DEVI void red_add_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
DEVI unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// cross-GPU block barrier (flag words in every rank's heap, monotonic epochs)
DEVI bool spin_until_ge(const uint32_t* p, uint32_t target, const CommDev& c) {
  unsigned it = 0; unsigned long long t0 = 0;
  while ((int32_t)(ld_acquire_sys(p) - target) < 0) {
    if (((++it) & 0x3ff) == 0) {
      unsigned long long t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > c.timeout_ns) {  // watchdog: report instead of hanging the box
        *reinterpret_cast<volatile uint32_t*>(c.status) = SY_ERR_TIMEOUT;
        __threadfence_system();
        return false;
      }
    }
  }
  return true;
}

// All threads of the block call this.  Orders every prior memory operation of
// the block before the signal and every later one after the wait.
DEVI void block_barrier(const CommDev& c, uint32_t& ep) {
  __syncthreads();
  ep += 1;
  if ((int)threadIdx.x < c.world) {
    const int p = threadIdx.x;
    uint32_t* remote = reinterpret_cast<uint32_t*>(c.heap[p] + SY_FLAGS_OFF) + (blockIdx.x * SY_MAXR + c.rank);
    __threadfence_system();
    st_release_sys(remote, ep);
    const uint32_t* local = reinterpret_cast<const uint32_t*>(c.heap[c.rank] + SY_FLAGS_OFF) + (blockIdx.x * SY_MAXR + p);
    spin_until_ge(local, ep, c);
  }
  __syncthreads();
}
DEVI uint32_t epoch_load(const CommDev& c) { return c.epoch[blockIdx.x]; }
DEVI void epoch_store(const CommDev& c, uint32_t ep) {
  if (threadIdx.x == 0) c.epoch[blockIdx.x] = ep;
}

