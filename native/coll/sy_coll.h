// libshipyard_coll — single-box NVSwitch collectives for the shipyard B200 launcher.
//
// One process per GPU.  Every rank owns a *symmetric heap* (a cuMemCreate
// allocation exported as a POSIX fd, or a cudaIpc handle as fallback) that is
// mapped into every peer, plus — when the fabric supports it — one NVLS
// multicast object bound over all heaps.  Collectives are single CUDA kernels
// that do P2P / multimem loads and stores from inside the kernel and fuse the
// scale + dtype cast (and, for training, the SGD update) into the same pass.
//
// A host shared-memory "stub" transport with the same API makes every
// collective runnable on a CPU-only box (world_size >= 1), which is how the
// control-plane tests exercise multi-instance tasks without GPUs.
//
// Replaces the data plane the reference delegates to container images
// (mpirun + NCCL/MPI inside user containers: convoy/batch.py:4362-4486,
// SURVEY.md §2D/§2E rows K1-K12).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sy_comm sy_comm;
typedef void* sy_stream_t;  // cudaStream_t (ignored by the stub transport)

enum sy_dtype {
  SY_F32 = 0, SY_BF16 = 1, SY_F16 = 2, SY_F64 = 3, SY_I32 = 4, SY_I64 = 5, SY_U8 = 6,
};
enum sy_op { SY_SUM = 0, SY_MAX = 1, SY_MIN = 2, SY_PROD = 3 };
enum sy_algo {
  SY_ALGO_AUTO = 0,
  SY_ALGO_LL = 1,            // one-shot, 8B data+flag lines, no barrier (tiny messages)
  SY_ALGO_ONESHOT = 2,       // one-shot push into peer mailboxes + per-block flags
  SY_ALGO_TWOSHOT_P2P = 3,   // pull reduce-scatter + push all-gather over P2P
  SY_ALGO_TWOSHOT_NVLS = 4,  // multimem.ld_reduce + multimem.st through the switch
};
enum sy_transport {
  SY_TRANSPORT_AUTO = 0, SY_TRANSPORT_STUB = 1, SY_TRANSPORT_P2P = 2, SY_TRANSPORT_NVLS = 3,
};
enum sy_err {
  SY_OK = 0, SY_ERR_ARG = 1, SY_ERR_CUDA = 2, SY_ERR_SYS = 3, SY_ERR_UNSUPPORTED = 4,
  SY_ERR_NOMEM = 5, SY_ERR_TIMEOUT = 6,
};

// ---- lifecycle ------------------------------------------------------------
// `session` must be identical on all ranks of the communicator and unique per
// communicator (the task runner derives it from job/task id).
// device < 0 selects the stub (host shared memory) transport.
int sy_comm_init(sy_comm** out, int rank, int world, const char* session, int device,
                 size_t heap_bytes, int transport);
int sy_comm_destroy(sy_comm* c);
int sy_comm_rank(const sy_comm* c);
int sy_comm_world(const sy_comm* c);
int sy_comm_transport(const sy_comm* c);     // resolved transport
int sy_comm_has_multicast(const sy_comm* c);
const char* sy_last_error(void);
// device-side watchdog/status word (0 = ok); non-zero after a flag-wait timeout
int sy_comm_status(sy_comm* c);
// number of kernels this library launched on behalf of `c` (bench accounting)
uint64_t sy_comm_launch_count(const sy_comm* c);

// ---- symmetric heap -------------------------------------------------------
// Collective bump allocator: every rank must call with the same sizes in the
// same order, so an allocation has the same offset in every heap.
void* sy_sym_alloc(sy_comm* c, size_t bytes);
int sy_sym_reset(sy_comm* c);                       // free everything (collective)
void* sy_heap_base(sy_comm* c, int peer);           // local VA of peer's heap
void* sy_mc_base(sy_comm* c);                       // multicast VA or NULL
size_t sy_heap_bytes(const sy_comm* c);
int sy_is_symmetric(sy_comm* c, const void* p);
// copy of the device-side communicator view (struct CommDev) for fused compute+collective kernels
size_t sy_comm_device_view(sy_comm* c, void* out, size_t cap);
void sy_comm_count_launch(sy_comm* c);

// ---- tuning ---------------------------------------------------------------
// knob names: "max_blocks", "threads", "ll_max_bytes", "oneshot_max_bytes",
// "nvls_min_bytes", "timeout_ms", "nvls_copy"
int sy_set_tuning(sy_comm* c, const char* knob, long value);
long sy_get_tuning(sy_comm* c, const char* knob);

// ---- collectives ----------------------------------------------------------
// out[i] = cast_out( scale * reduce_r in_r[i] ).  `in`/`out` may be any device
// pointers; symmetric-heap pointers take the zero-copy path.
int sy_allreduce(sy_comm* c, const void* in, void* out, size_t count, int dt_in, int dt_out,
                 float scale, int op, int algo, sy_stream_t stream);
// in: world*count elements; out: count elements (this rank's shard)
int sy_reduce_scatter(sy_comm* c, const void* in, void* out, size_t count, int dt_in, int dt_out,
                      float scale, int op, sy_stream_t stream);
// in: count elements; out: world*count elements
int sy_allgather(sy_comm* c, const void* in, void* out, size_t count, int dt, sy_stream_t stream);
int sy_broadcast(sy_comm* c, const void* in, void* out, size_t count, int dt, int root,
                 sy_stream_t stream);
// in/out: world blocks of `count` elements
int sy_alltoall(sy_comm* c, const void* in, void* out, size_t count, int dt, sy_stream_t stream);
int sy_reduce(sy_comm* c, const void* in, void* out, size_t count, int dt, int op, int root,
              sy_stream_t stream);
int sy_gather(sy_comm* c, const void* in, void* out, size_t count, int dt, int root,
              sy_stream_t stream);
int sy_scatter(sy_comm* c, const void* in, void* out, size_t count, int dt, int root,
               sy_stream_t stream);
int sy_barrier(sy_comm* c, sy_stream_t stream);

// Point-to-point "push": copy `bytes` from local `src` into peer's heap at
// `dst_off`, then bump peer's signal word `sig_idx`.  wait spins locally.
// (HPCG halo exchange, K9.)
int sy_put_signal(sy_comm* c, const void* src, size_t dst_off, size_t bytes, int peer, int sig_idx,
                  sy_stream_t stream);
int sy_wait_signal(sy_comm* c, int sig_idx, uint32_t expected_count, sy_stream_t stream);

// Fused pack + push of up to 26 strided halo regions (HPCG 27-point stencil).
typedef struct sy_halo_desc {
  int peer;            // destination rank
  int sig_idx;         // signal word on the destination
  long dst_off;        // byte offset in the destination heap
  int nx, ny, nz;      // extent of the boundary region (elements)
  long sx, sy, sz;     // element strides in the source array
  long src_elem_off;   // first element of the region in the source array
} sy_halo_desc;
int sy_halo_exchange(sy_comm* c, const void* src, int dt, const sy_halo_desc* descs, int ndesc,
                     const int* wait_sig, int nwait, sy_stream_t stream);

// Fused gradient all-reduce + SGD(momentum) step + parameter all-gather
// (ZeRO-1 style, one kernel):
//   g   = scale * sum_r grads_r[i]                  (reduce-scatter, shard of this rank)
//   g  += wd * master[i];  mom = mu*mom + g;  master -= lr*mom
//   params_r[i] = cast(master[i]) on every rank      (all-gather through the switch)
//   grads (local) zeroed for the next accumulation
// grads/params must be symmetric allocations of `count` elements; master/mom
// are local fp32 arrays of sy_shard_count(count) elements.  hyper points at
// 4 floats in device memory {lr, momentum, weight_decay, scale} so the launch
// is CUDA-graph capturable with a changing learning rate.
size_t sy_shard_begin(const sy_comm* c, size_t count, int rank);
size_t sy_shard_count(const sy_comm* c, size_t count, int rank);
int sy_fused_allreduce_sgd(sy_comm* c, void* grads, int dt_grad, void* params, int dt_param,
                           float* master, float* mom, const float* hyper, size_t count,
                           int zero_grads, sy_stream_t stream);

// One-shot gradient all-reduce (averaged by `scale`) fused with the Adam update of every rank's own parameter replica:
// grad/param/m/v are local fp32 arrays of `count` elements (count % 4 == 0, <= 256 Ki elements), hyper = device floats
// {lr, beta1, beta2, eps, step}; step is advanced by the kernel.  One launch per training step.
int sy_fused_allreduce_adam(sy_comm* c, float* grad, float* param, float* m, float* v, float* hyper, size_t count,
                            float scale, int zero_grad, sy_stream_t stream);

// All-reduce whose output is block-scaled fp8: out_q[i] (e4m3) and one e8m0
// scale byte per 32 elements (MX format), reduced in fp32 (K11).
int sy_allreduce_fp8_blockscaled(sy_comm* c, const void* in, int dt_in, void* out_q,
                                 void* out_scales, size_t count, float scale, sy_stream_t stream);

#ifdef __cplusplus
}
#endif
