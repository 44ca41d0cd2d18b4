// Internal definitions shared by the host runtime and the kernels.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include "sy_coll.h"

#define SY_MAXR 8            // ranks on one NVSwitch box
#define SY_MAX_BLOCKS 256    // upper bound on grid size of any collective kernel
#define SY_NSIG 256          // point-to-point signal words per rank

// ---- symmetric heap layout (identical offsets on every rank) ---------------
// [0,            16K)  barrier flags   uint32 flags[SY_MAX_BLOCKS][SY_MAXR] (8K used, padded)
// [16K,          32K)  one-shot flags  uint32 os_flags[2][SY_MAX_BLOCKS][SY_MAXR]
// [32K,          36K)  p2p signal words uint32 sig[SY_NSIG]
// [64K,   64K+LLTOT )  LL mailboxes    2 parity x SY_MAXR x SY_LL_SLOT bytes
// [...,   ...+OSTOT )  one-shot mailboxes 2 parity x SY_MAXR x SY_OS_SLOT bytes
// [SY_USER_OFF, end )  user allocations (bump allocator) + lazily carved staging
#define SY_FLAGS_OFF 0ul
#define SY_OSFLAGS_OFF (16ul << 10)
#define SY_SIG_OFF (32ul << 10)
#define SY_LL_OFF (64ul << 10)
#define SY_LL_MAX_PAYLOAD (16ul << 10)               // bytes of payload per rank
#define SY_LL_SLOT (2 * SY_LL_MAX_PAYLOAD)           // 8B line carries 4B payload
#define SY_LL_TOT (2ul * SY_MAXR * SY_LL_SLOT)       // 512 KB
#define SY_OS_OFF (SY_LL_OFF + SY_LL_TOT)
#define SY_OS_SLOT (1ul << 20)                       // 1 MB payload per rank
#define SY_OS_TOT (2ul * SY_MAXR * SY_OS_SLOT)       // 16 MB
// [17M,  17M+LMTOT )  multi-block LL mailboxes 2 parity x SY_MAXR x SY_LM_SLOT bytes (only ever written in LL line format)
#define SY_LM_OFF (17ul << 20)
#define SY_LM_MAX_PAYLOAD (256ul << 10)              // bytes of payload per writer
#define SY_LM_SLOT (2 * SY_LM_MAX_PAYLOAD)           // 16 B line = two 8-byte atoms {word, flag}
#define SY_LM_TOT (2ul * SY_MAXR * SY_LM_SLOT)       // 8 MB; followed by 2 x SY_MAXR 16-byte token lines (broadcast)
#define SY_USER_OFF (32ul << 20)                     // 32 MB, 2 MB aligned

struct CommDev {
  int rank;
  int world;
  char* heap[SY_MAXR];   // local VA of every rank's heap (heap[rank] is ours)
  char* mc;              // multicast VA aliasing all heaps, or nullptr
  uint32_t* epoch;       // local: per-block barrier epoch counters [SY_MAX_BLOCKS]
  uint32_t* seq;         // local: [0]=LL sequence, [1]=one-shot sequence
  uint32_t* status;      // mapped pinned host word: watchdog writes an error code here
  unsigned long long timeout_ns;
};

struct Hub;  // socket rendezvous (bootstrap.cpp)

struct sy_comm {
  int rank = 0, world = 1, device = -1;
  int transport = SY_TRANSPORT_STUB;
  bool has_mc = false;
  std::string session;
  Hub* hub = nullptr;
  size_t heap_bytes = 0;
  size_t bump = SY_USER_OFF;          // next free offset
  size_t stage_off = 0, stage_bytes = 0;  // staging region (2 halves: in / out)
  CommDev dev{};
  // host-side bookkeeping
  uint64_t launches = 0;
  uint32_t* status_host = nullptr;
  // tuning
  long max_blocks = 128, threads = 512;
  long ll_max_bytes = 16384, oneshot_max_bytes = 256 << 10, nvls_min_bytes = 256 << 10;   // LL up to its 16 KB payload: 10 us vs 18 us (one-shot) at 16 KB, N = 4 (round-2 sweep)
  long lm_max_bytes = 256 << 10;        // per-writer payload up to which gathers / exchanges / sum-reductions use the multi-block LL kernel (k_lm_k)
  long mailbox_max_bytes = 1 << 20;     // per-writer payload up to which all-gather / all-to-all / broadcast use the mailbox kernel
  long timeout_ms = 20000;
  long nvls_min_world = 4;  // below this world size the P2P paths win (measured at N=2)
  long nvls_copy = 1;       // all-gather / broadcast through multimem.st when multicast exists
  long ag_p2p_min_bytes = 16 << 20;     // all-gathers whose OUTPUT is at least this big use direct peer stores instead of multimem.st (see sy_allgather)
  long bcast_sag_min_bytes = 8 << 20;   // broadcasts from this size on run as pipelined scatter + all-gather (world >= 4, multicast)
  // VMM handles (opaque to other TUs)
  void* impl = nullptr;
  // stub transport state
  void* shm_base = nullptr; size_t shm_bytes = 0; uint64_t stub_gen = 0;
};

void sy_set_error(const char* fmt, ...);

// ---- bootstrap.cpp ----------------------------------------------------------
Hub* hub_create(int rank, int world, const std::string& session, int timeout_ms);
void hub_destroy(Hub*);
// every rank contributes `len` bytes; all[] receives world*len bytes in rank order
int hub_allgather(Hub*, const void* mine, size_t len, void* all);
// every rank contributes one fd; fds_out receives `world` fds (own slot = dup of own fd)
int hub_allgather_fd(Hub*, int myfd, int* fds_out);
// rank 0 contributes fd; everyone receives it
int hub_bcast_fd(Hub*, int* fd_inout);
int hub_barrier(Hub*);

// ---- stub.cpp ---------------------------------------------------------------
int stub_init(sy_comm* c);
void stub_destroy(sy_comm* c);
int stub_barrier(sy_comm* c);
int stub_allreduce(sy_comm* c, const void* in, void* out, size_t count, int dt_in, int dt_out,
                   float scale, int op);
int stub_reduce_scatter(sy_comm* c, const void* in, void* out, size_t count, int dt_in, int dt_out,
                        float scale, int op);
int stub_allgather(sy_comm* c, const void* in, void* out, size_t count, int dt);
int stub_broadcast(sy_comm* c, const void* in, void* out, size_t count, int dt, int root);
int stub_alltoall(sy_comm* c, const void* in, void* out, size_t count, int dt);
int stub_reduce(sy_comm* c, const void* in, void* out, size_t count, int dt, int op, int root);
int stub_gather(sy_comm* c, const void* in, void* out, size_t count, int dt, int root);
int stub_scatter(sy_comm* c, const void* in, void* out, size_t count, int dt, int root);
int stub_put_signal(sy_comm* c, const void* src, size_t dst_off, size_t bytes, int peer, int sig);
int stub_wait_signal(sy_comm* c, int sig, uint32_t expected);
int stub_fused_sgd(sy_comm* c, void* grads, int dt_grad, void* params, int dt_param, float* master,
                   float* mom, const float* hyper, size_t count, int zero_grads);
int stub_allreduce_fp8(sy_comm* c, const void* in, int dt_in, void* out_q, void* out_scales,
                       size_t count, float scale);

// ---- gpu_mem.cpp ------------------------------------------------------------
int gpu_init(sy_comm* c, int requested_transport);
void gpu_destroy(sy_comm* c);

// ---- kernels.cu launchers ---------------------------------------------------
struct LaunchCfg { int blocks; int threads; void* stream; };
int k_allreduce(sy_comm* c, const void* in, void* out, size_t in_off, size_t out_off, bool in_sym,
                bool out_sym, size_t count, int dt_in, int dt_out, float scale, int op, int algo,
                void* stream);
int k_reduce_scatter(sy_comm* c, size_t in_off, void* out, size_t count, int dt_in, int dt_out,
                     float scale, int op, bool nvls, void* stream);
int k_mailbox(sy_comm* c, const void* in, void* out, size_t bytes, int mode, int root, void* stream, float scale = 1.0f);
int k_lm(sy_comm* c, const void* in, void* out, size_t bytes, int mode, int dt, void* stream, float scale = 1.0f);
int k_allgather(sy_comm* c, const void* in, size_t out_off, size_t count, int dt, bool nvls,
                void* stream);
int k_broadcast(sy_comm* c, const void* in, size_t out_off, size_t bytes, int root, bool nvls,
                void* stream);
int k_alltoall(sy_comm* c, const void* in, size_t out_off, size_t bytes_per_peer, void* stream);
int k_gather(sy_comm* c, const void* in, size_t out_off, size_t bytes, int root, void* stream);
int k_scatter(sy_comm* c, size_t in_off, void* out, size_t bytes, int root, void* stream);
int k_barrier(sy_comm* c, void* stream);
int k_put_signal(sy_comm* c, const void* src, size_t dst_off, size_t bytes, int peer, int sig,
                 void* stream);
int k_wait_signal(sy_comm* c, int sig, uint32_t expected, void* stream);
int k_halo(sy_comm* c, const void* src, int dt, const sy_halo_desc* descs, int ndesc,
           const int* wait_sig, int nwait, void* stream);
int k_oneshot_adam(sy_comm* c, float* grad, float* param, float* m, float* v, float* hyper, size_t count, float scale, int zero_grad, void* stream);
int stub_fused_adam(sy_comm* c, float* grad, float* param, float* m, float* v, float* hyper, size_t count, float scale, int zero_grad);
int k_fused_sgd(sy_comm* c, size_t grads_off, int dt_grad, size_t params_off, int dt_param,
                float* master, float* mom, const float* hyper, size_t count, int zero_grads,
                void* stream);
int k_allreduce_fp8(sy_comm* c, size_t in_off, int dt_in, void* out_q, void* out_scales,
                    size_t count, float scale, void* stream);
int k_reduce_rooted(sy_comm* c, size_t in_off, void* out, size_t count, int dt, int op, int root,
                    void* stream);
int k_local_cast(sy_comm* c, const void* in, void* out, size_t count, int dt_in, int dt_out,
                 float scale, void* stream);
#define SY_SEQ_WORDS (16 + SY_NSIG + 32)   // [0]=one-shot seq [1]=LL seq [8..10]=done counters [16..]=signal expects [16+NSIG..]=halo block counters

static inline size_t sy_dtype_size(int dt) {
  switch (dt) {
    case SY_F32: case SY_I32: return 4;
    case SY_BF16: case SY_F16: return 2;
    case SY_F64: case SY_I64: return 8;
    case SY_U8: return 1;
  }
  return 0;
}
