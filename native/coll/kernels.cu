// sm_100a collective kernels over NVSwitch peer memory.
//
// Every collective is ONE kernel: the cross-GPU synchronisation (flag barrier on
// P2P-mapped words), the data movement (P2P ld/st or multimem.ld_reduce /
// multimem.st through the switch) and the arithmetic that follows the
// collective (1/N scale, dtype cast, SGD update, fp8 block quantisation) happen
// in the same pass, so the reduced gradient never makes an extra HBM round trip.
//
// These kernels are NVLink-bandwidth / latency bound; tensor cores do not apply
// here (see native/gemm for the tcgen05 GEMM + reduce-scatter fusion, K10).
// Call-site inventory: SURVEY.md §2E K1-K9, K11.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "internal.h"

#define DEVI __device__ __forceinline__
#include "device_sync.cuh"

// ---------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------
struct alignas(16) V16 { uint32_t x, y, z, w; };

DEVI V16 ld16(const void* p) {  // streaming 16B load (peer or local), no L1 allocation
  V16 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
DEVI void st16(void* p, const V16& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w) : "memory");
}
DEVI V16 ld16_volatile(const void* p) {
  V16 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
DEVI void st16_volatile(void* p, const V16& v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w) : "memory");
}
// NVLS: reduce the same address across every GPU bound to the multicast object, in the switch
DEVI V16 mc_ld_reduce_f32(const void* mc) {
  V16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}
DEVI V16 mc_ld_reduce_bf16(const void* mc) {
  V16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}
DEVI V16 mc_ld_reduce_f16(const void* mc) {
  V16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}
// one store, replicated by the switch into every bound GPU's memory
DEVI void mc_st16(void* mc, const V16& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
DEVI void mc_st4(void* mc, uint32_t v) {
  asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------
// dtype traits: a "unit" is EPU elements; IN_V / OUT_V 16-byte vectors
// ---------------------------------------------------------------------------
template <int DT> struct DtSize;
template <> struct DtSize<SY_F32> { static constexpr int v = 4; };
template <> struct DtSize<SY_BF16> { static constexpr int v = 2; };
template <> struct DtSize<SY_F16> { static constexpr int v = 2; };
template <> struct DtSize<SY_F64> { static constexpr int v = 8; };
template <> struct DtSize<SY_I32> { static constexpr int v = 4; };
template <> struct DtSize<SY_I64> { static constexpr int v = 8; };

template <int DT> struct Acc { using t = float; };
template <> struct Acc<SY_F64> { using t = double; };
template <> struct Acc<SY_I32> { using t = int; };
template <> struct Acc<SY_I64> { using t = long long; };

// unpack N elements of type DT packed in `words` into acc array
template <int DT, int N> struct Codec;
template <int N> struct Codec<SY_F32, N> {
  DEVI static void unpack(const uint32_t* w, float* a) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = __uint_as_float(w[i]);
  }
  DEVI static void pack(const float* a, uint32_t* w) {
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = __float_as_uint(a[i]);
  }
};
template <int N> struct Codec<SY_BF16, N> {
  DEVI static void unpack(const uint32_t* w, float* a) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      a[2 * i] = __uint_as_float(w[i] << 16);
      a[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  DEVI static void pack(const float* a, uint32_t* w) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
};
template <int N> struct Codec<SY_F16, N> {
  DEVI static void unpack(const uint32_t* w, float* a) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 f = __half22float2(h);
      a[2 * i] = f.x; a[2 * i + 1] = f.y;
    }
  }
  DEVI static void pack(const float* a, uint32_t* w) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      __half2 h = __floats2half2_rn(a[2 * i], a[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
};
template <int N> struct Codec<SY_F64, N> {
  DEVI static void unpack(const uint32_t* w, double* a) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]);
  }
  DEVI static void pack(const double* a, uint32_t* w) {
#pragma unroll
    for (int i = 0; i < N; ++i) { w[2 * i] = (uint32_t)__double2loint(a[i]); w[2 * i + 1] = (uint32_t)__double2hiint(a[i]); }
  }
};
template <int N> struct Codec<SY_I32, N> {
  DEVI static void unpack(const uint32_t* w, int* a) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = (int)w[i];
  }
  DEVI static void pack(const int* a, uint32_t* w) {
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = (uint32_t)a[i];
  }
};
template <int N> struct Codec<SY_I64, N> {
  DEVI static void unpack(const uint32_t* w, long long* a) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = (long long)(((unsigned long long)w[2 * i + 1] << 32) | w[2 * i]);
  }
  DEVI static void pack(const long long* a, uint32_t* w) {
#pragma unroll
    for (int i = 0; i < N; ++i) { w[2 * i] = (uint32_t)(unsigned long long)a[i]; w[2 * i + 1] = (uint32_t)((unsigned long long)a[i] >> 32); }
  }
};

template <int OP, typename T> DEVI T op_apply(T a, T b) {
  if (OP == SY_SUM) return a + b;
  if (OP == SY_MAX) return a > b ? a : b;
  if (OP == SY_MIN) return a < b ? a : b;
  return a * b;
}
template <typename TA, typename TO> DEVI TO acc_cast(TA a) { return (TO)a; }

// scalar element load/store (tails, unaligned ranges)
template <int DT> DEVI typename Acc<DT>::t ld_elem(const void* base, size_t i) {
  if (DT == SY_F32) return (typename Acc<DT>::t)((const float*)base)[i];
  if (DT == SY_BF16) return (typename Acc<DT>::t)__bfloat162float(((const __nv_bfloat16*)base)[i]);
  if (DT == SY_F16) return (typename Acc<DT>::t)__half2float(((const __half*)base)[i]);
  if (DT == SY_F64) return (typename Acc<DT>::t)((const double*)base)[i];
  if (DT == SY_I32) return (typename Acc<DT>::t)((const int*)base)[i];
  return (typename Acc<DT>::t)((const long long*)base)[i];
}
template <int DT, typename TA> DEVI void st_elem(void* base, size_t i, TA v) {
  if (DT == SY_F32) ((float*)base)[i] = (float)v;
  else if (DT == SY_BF16) ((__nv_bfloat16*)base)[i] = __float2bfloat16_rn((float)v);
  else if (DT == SY_F16) ((__half*)base)[i] = __float2half_rn((float)v);
  else if (DT == SY_F64) ((double*)base)[i] = (double)v;
  else if (DT == SY_I32) ((int*)base)[i] = (int)v;
  else ((long long*)base)[i] = (long long)v;
}

template <int DT_IN, int DT_OUT> struct Unit {
  static constexpr int SI = DtSize<DT_IN>::v, SO = DtSize<DT_OUT>::v;
  static constexpr int EPU = 16 / (SI < SO ? SI : SO);   // elements per unit
  static constexpr int IN_V = EPU * SI / 16;             // 16B vectors per unit on the input side
  static constexpr int OUT_V = EPU * SO / 16;
  using acc_t = typename Acc<DT_IN>::t;
  using oacc_t = typename Acc<DT_OUT>::t;
};

// balanced split of `units` over world ranks
DEVI void shard_of(size_t units, int world, int r, size_t& begin, size_t& n) {
  size_t base = units / world, rem = units % world;
  begin = (size_t)r * base + ((size_t)r < rem ? (size_t)r : rem);
  n = base + ((size_t)r < rem ? 1 : 0);
}

// ---------------------------------------------------------------------------
// reduce one unit (EPU elements at unit index u) across all ranks over P2P
// ---------------------------------------------------------------------------
// NW: compile-time upper bound on the world size (2, 4 or 8) so the register file only
// holds NW in-flight vectors per unit and small worlds can unroll over more units.
template <int DT_IN, int DT_OUT, int OP, int NW>
DEVI void p2p_load_unit(const CommDev& c, size_t in_off, size_t u, V16 (&v)[NW][Unit<DT_IN, DT_OUT>::IN_V]) {
  using U = Unit<DT_IN, DT_OUT>;
#pragma unroll
  for (int r = 0; r < NW; ++r) {
    if (r < c.world) {
      const char* src = c.heap[r] + in_off + u * (size_t)(U::IN_V * 16);
#pragma unroll
      for (int k = 0; k < U::IN_V; ++k) v[r][k] = ld16(src + k * 16);
    }
  }
}
template <int DT_IN, int DT_OUT, int OP, int NW>
DEVI void p2p_finish_unit(const CommDev& c, V16 (&v)[NW][Unit<DT_IN, DT_OUT>::IN_V], float scale,
                          uint32_t (&outw)[Unit<DT_IN, DT_OUT>::OUT_V * 4]) {
  using U = Unit<DT_IN, DT_OUT>;
  using acc_t = typename U::acc_t;
  acc_t acc[U::EPU];
  Codec<DT_IN, U::EPU>::unpack(reinterpret_cast<const uint32_t*>(v[0]), acc);
#pragma unroll
  for (int r = 1; r < NW; ++r) {
    if (r < c.world) {
      acc_t a[U::EPU];
      Codec<DT_IN, U::EPU>::unpack(reinterpret_cast<const uint32_t*>(v[r]), a);
#pragma unroll
      for (int i = 0; i < U::EPU; ++i) acc[i] = op_apply<OP>(acc[i], a[i]);
    }
  }
  typename U::oacc_t o[U::EPU];
#pragma unroll
  for (int i = 0; i < U::EPU; ++i) {
    if (DT_IN == SY_F32 || DT_IN == SY_BF16 || DT_IN == SY_F16) o[i] = (typename U::oacc_t)((float)acc[i] * scale);
    else if (DT_IN == SY_F64) o[i] = (typename U::oacc_t)((double)acc[i] * (double)scale);
    else o[i] = (typename U::oacc_t)acc[i];
  }
  Codec<DT_OUT, U::EPU>::pack(o, outw);
}

template <int DT_IN, int DT_OUT, int OP>
DEVI void p2p_reduce_unit(const CommDev& c, size_t in_off, size_t u, float scale,
                          uint32_t (&outw)[Unit<DT_IN, DT_OUT>::OUT_V * 4]) {
  using U = Unit<DT_IN, DT_OUT>;
  using acc_t = typename U::acc_t;
  V16 v[SY_MAXR][U::IN_V];
  // issue every peer load before consuming any (8 x 16B in flight per unit at N=8, one per
  // peer, so all 8 NVLink destinations are exercised at once); accumulate in rank order so
  // the result does not depend on which rank reduces
#pragma unroll
  for (int r = 0; r < SY_MAXR; ++r) {
    if (r < c.world) {
      const char* src = c.heap[r] + in_off + u * (size_t)(U::IN_V * 16);
#pragma unroll
      for (int k = 0; k < U::IN_V; ++k) v[r][k] = ld16(src + k * 16);
    }
  }
  acc_t acc[U::EPU];
  Codec<DT_IN, U::EPU>::unpack(reinterpret_cast<const uint32_t*>(v[0]), acc);
#pragma unroll
  for (int r = 1; r < SY_MAXR; ++r) {
    if (r < c.world) {
      acc_t a[U::EPU];
      Codec<DT_IN, U::EPU>::unpack(reinterpret_cast<const uint32_t*>(v[r]), a);
#pragma unroll
      for (int i = 0; i < U::EPU; ++i) acc[i] = op_apply<OP>(acc[i], a[i]);
    }
  }
  typename U::oacc_t o[U::EPU];
#pragma unroll
  for (int i = 0; i < U::EPU; ++i) {
    if (DT_IN == SY_F32 || DT_IN == SY_BF16 || DT_IN == SY_F16) o[i] = (typename U::oacc_t)((float)acc[i] * scale);
    else if (DT_IN == SY_F64) o[i] = (typename U::oacc_t)((double)acc[i] * (double)scale);
    else o[i] = (typename U::oacc_t)acc[i];
  }
  Codec<DT_OUT, U::EPU>::pack(o, outw);
}

template <int DT_IN, int DT_OUT, int OP>
DEVI typename Acc<DT_OUT>::t p2p_reduce_elem(const CommDev& c, size_t in_off, size_t i, float scale) {
  using acc_t = typename Acc<DT_IN>::t;
  acc_t acc = ld_elem<DT_IN>(c.heap[0] + in_off, i);
  for (int r = 1; r < c.world; ++r) acc = op_apply<OP>(acc, ld_elem<DT_IN>(c.heap[r] + in_off, i));
  if (DT_IN == SY_F32 || DT_IN == SY_BF16 || DT_IN == SY_F16) return (typename Acc<DT_OUT>::t)((float)acc * scale);
  if (DT_IN == SY_F64) return (typename Acc<DT_OUT>::t)((double)acc * (double)scale);
  return (typename Acc<DT_OUT>::t)acc;
}

// ---------------------------------------------------------------------------
// K1/K3: two-shot all-reduce over P2P: pull reduce-scatter + push all-gather,
// fused scale + cast.  in/out are symmetric-heap offsets (in-place allowed).
// mode 0: all-reduce; mode 1: reduce-scatter into local `rs_out` (count = per-rank count)
// ---------------------------------------------------------------------------
template <int DT_IN, int DT_OUT, int OP, int NW, int UNR>
__global__ void __launch_bounds__(512)
k_twoshot_p2p(const __grid_constant__ CommDev c, size_t in_off, size_t out_off, size_t count,
              float scale, int mode, void* rs_out) {
  using U = Unit<DT_IN, DT_OUT>;
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);  // every rank's input is ready, and nobody still reads `out`
  size_t ebeg, eend;     // element range this rank reduces
  if (mode == 0) {
    size_t units = count / U::EPU, ub, un;
    shard_of(units, c.world, c.rank, ub, un);
    ebeg = ub * U::EPU; eend = (ub + un) * U::EPU;
    if (c.rank == c.world - 1) eend = count;  // tail elements go to the last rank
  } else {
    ebeg = (size_t)c.rank * count; eend = ebeg + count;
  }
  const size_t ufirst = (ebeg + U::EPU - 1) / U::EPU, ulast = eend / U::EPU;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (ulast > ufirst) {
    // UNR units per thread per iteration: UNR x world 16B loads in flight before the first use
    for (size_t ub = ufirst + tid; ub < ulast; ub += (size_t)UNR * nth) {
      V16 v[UNR][NW][U::IN_V];
#pragma unroll
      for (int j = 0; j < UNR; ++j)
        if (ub + (size_t)j * nth < ulast) p2p_load_unit<DT_IN, DT_OUT, OP, NW>(c, in_off, ub + (size_t)j * nth, v[j]);
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t u = ub + (size_t)j * nth;
        if (u >= ulast) break;
        uint32_t w[U::OUT_V * 4];
        p2p_finish_unit<DT_IN, DT_OUT, OP, NW>(c, v[j], scale, w);
        if (mode == 0) {
#pragma unroll
          for (int p = 0; p < NW; ++p)
            if (p < c.world) {
              int q = c.rank + p; if (q >= c.world) q -= c.world;
              char* dst = c.heap[q] + out_off + u * (size_t)(U::OUT_V * 16);
#pragma unroll
              for (int k = 0; k < U::OUT_V; ++k) st16(dst + k * 16, *reinterpret_cast<V16*>(&w[4 * k]));
            }
        } else {
          char* dst = (char*)rs_out + (u * U::EPU - ebeg) * U::SO;
#pragma unroll
          for (int k = 0; k < U::OUT_V; ++k) st16(dst + k * 16, *reinterpret_cast<V16*>(&w[4 * k]));
        }
      }
    }
  }
  // scalar head/tail (elements outside whole units)
  {
    size_t head_end = ufirst * U::EPU < eend ? ufirst * U::EPU : eend;
    size_t tail_beg = ulast * U::EPU > head_end ? ulast * U::EPU : head_end;
    size_t nscal = (head_end - ebeg) + (eend - tail_beg);
    for (size_t s = tid; s < nscal; s += nth) {
      size_t i = s < (head_end - ebeg) ? ebeg + s : tail_beg + (s - (head_end - ebeg));
      auto v = p2p_reduce_elem<DT_IN, DT_OUT, OP>(c, in_off, i, scale);
      if (mode == 0) { for (int p = 0; p < c.world; ++p) st_elem<DT_OUT>(c.heap[p] + out_off, i, v); }
      else st_elem<DT_OUT>(rs_out, i - ebeg, v);
    }
  }
  block_barrier(c, ep);  // pushes visible everywhere; peers are done reading my input
  epoch_store(c, ep);
}

// ---------------------------------------------------------------------------
// K1: two-shot all-reduce through the switch (NVLS): multimem.ld_reduce own shard,
// scale/cast in registers, multimem.st the result to every GPU.
// DT in {F32,BF16,F16}; OUT may differ (bf16->f32, f32->bf16).
// ---------------------------------------------------------------------------
template <int DT> DEVI V16 mc_ld_reduce(const void* p);
template <> DEVI V16 mc_ld_reduce<SY_F32>(const void* p) { return mc_ld_reduce_f32(p); }
template <> DEVI V16 mc_ld_reduce<SY_BF16>(const void* p) { return mc_ld_reduce_bf16(p); }
template <> DEVI V16 mc_ld_reduce<SY_F16>(const void* p) { return mc_ld_reduce_f16(p); }

template <int DT_IN, int DT_OUT>
DEVI void nvls_reduce_unit(const CommDev& c, size_t in_off, size_t u, float scale,
                           uint32_t (&outw)[Unit<DT_IN, DT_OUT>::OUT_V * 4]) {
  using U = Unit<DT_IN, DT_OUT>;
  V16 s[U::IN_V];
  const char* src = c.mc + in_off + u * (size_t)(U::IN_V * 16);
#pragma unroll
  for (int k = 0; k < U::IN_V; ++k) s[k] = mc_ld_reduce<DT_IN>(src + k * 16);
  if (DT_IN == DT_OUT && scale == 1.0f) {
#pragma unroll
    for (int k = 0; k < U::IN_V * 4; ++k) outw[k] = reinterpret_cast<uint32_t*>(s)[k];
    return;
  }
  float a[U::EPU];
  Codec<DT_IN, U::EPU>::unpack(reinterpret_cast<const uint32_t*>(s), a);
#pragma unroll
  for (int i = 0; i < U::EPU; ++i) a[i] *= scale;
  Codec<DT_OUT, U::EPU>::pack(a, outw);
}

template <int DT_IN, int DT_OUT, int UNR>
__global__ void __launch_bounds__(512)
k_twoshot_nvls(const __grid_constant__ CommDev c, size_t in_off, size_t out_off, size_t count,
               float scale, int mode, void* rs_out) {
  using U = Unit<DT_IN, DT_OUT>;
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);
  size_t ebeg, eend;
  if (mode == 0) {
    size_t units = count / U::EPU, ub, un;
    shard_of(units, c.world, c.rank, ub, un);
    ebeg = ub * U::EPU; eend = (ub + un) * U::EPU;
    if (c.rank == c.world - 1) eend = count;
  } else { ebeg = (size_t)c.rank * count; eend = ebeg + count; }
  const size_t ufirst = (ebeg + U::EPU - 1) / U::EPU, ulast = eend / U::EPU;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (ulast > ufirst) {
    // UNR multimem.ld_reduce in flight per thread hide the switch round trip
    for (size_t ub = ufirst + tid; ub < ulast; ub += (size_t)UNR * nth) {
      V16 raw[UNR][U::IN_V];
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t u = ub + (size_t)j * nth;
        if (u < ulast) {
#pragma unroll
          for (int k = 0; k < U::IN_V; ++k) raw[j][k] = mc_ld_reduce<DT_IN>(c.mc + in_off + u * (size_t)(U::IN_V * 16) + k * 16);
        }
      }
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t u = ub + (size_t)j * nth;
        if (u >= ulast) break;
        uint32_t w0[U::OUT_V * 4];
        if (DT_IN == DT_OUT && scale == 1.0f) {
#pragma unroll
          for (int k = 0; k < U::IN_V * 4; ++k) w0[k] = reinterpret_cast<uint32_t*>(raw[j])[k];
        } else {
          float a[U::EPU];
          Codec<DT_IN, U::EPU>::unpack(reinterpret_cast<const uint32_t*>(raw[j]), a);
#pragma unroll
          for (int i = 0; i < U::EPU; ++i) a[i] *= scale;
          Codec<DT_OUT, U::EPU>::pack(a, w0);
        }
#pragma unroll
        for (int k = 0; k < U::OUT_V; ++k) {
          if (mode == 0) mc_st16(c.mc + out_off + u * (size_t)(U::OUT_V * 16) + k * 16, *reinterpret_cast<V16*>(&w0[4 * k]));
          else st16((char*)rs_out + (u * U::EPU - ebeg) * U::SO + k * 16, *reinterpret_cast<V16*>(&w0[4 * k]));
        }
      }
    }
  }
  {  // scalar head/tail over plain P2P
    size_t head_end = ufirst * U::EPU < eend ? ufirst * U::EPU : eend;
    size_t tail_beg = ulast * U::EPU > head_end ? ulast * U::EPU : head_end;
    size_t nscal = (head_end - ebeg) + (eend - tail_beg);
    for (size_t s = tid; s < nscal; s += nth) {
      size_t i = s < (head_end - ebeg) ? ebeg + s : tail_beg + (s - (head_end - ebeg));
      auto v = p2p_reduce_elem<DT_IN, DT_OUT, SY_SUM>(c, in_off, i, scale);
      if (mode == 0) { for (int p = 0; p < c.world; ++p) st_elem<DT_OUT>(c.heap[p] + out_off, i, v); }
      else st_elem<DT_OUT>(rs_out, i - ebeg, v);
    }
  }
  block_barrier(c, ep);
  epoch_store(c, ep);
}

// ---------------------------------------------------------------------------
// K2/K3: one-shot all-reduce (push): every rank writes its data into its slot of
// every peer's mailbox, one flag per block, then reduces locally.  Any device
// pointers for in/out; no start barrier (mailboxes are double-buffered by parity).
// ---------------------------------------------------------------------------
DEVI char* os_slot(const CommDev& c, int on_rank, uint32_t parity, int writer) {
  return c.heap[on_rank] + SY_OS_OFF + ((size_t)parity * SY_MAXR + writer) * SY_OS_SLOT;
}
DEVI void seq_finish(const CommDev& c, int which) {
  // last block to finish bumps the global sequence (all blocks have read it by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    uint32_t done = atomicAdd(&c.seq[8 + which], 1u);
    if (done == gridDim.x - 1) { c.seq[8 + which] = 0; c.seq[which] += 1; __threadfence(); }
  }
}

template <int DT_IN, int DT_OUT, int OP>
__global__ void __launch_bounds__(512)
k_oneshot(const __grid_constant__ CommDev c, const void* in, void* out, size_t count, float scale) {
  using U = Unit<DT_IN, DT_OUT>;
  const uint32_t seq = c.seq[0] + 1, parity = seq & 1;
  const size_t units = (count + U::EPU - 1) / U::EPU;   // last unit may be partial
  const size_t per_block = (units + gridDim.x - 1) / gridDim.x;
  const size_t u0 = (size_t)blockIdx.x * per_block, u1 = (u0 + per_block < units) ? u0 + per_block : units;
  const bool in_al = (((uintptr_t)in) & 15) == 0, out_al = (((uintptr_t)out) & 15) == 0;
  // phase 1: push my data to everyone (self included)
  for (size_t u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
    V16 v[U::IN_V];
    const size_t e0 = u * U::EPU;
    if (in_al && e0 + U::EPU <= count) {
#pragma unroll
      for (int k = 0; k < U::IN_V; ++k) v[k] = ld16((const char*)in + u * (size_t)(U::IN_V * 16) + k * 16);
    } else {
      typename U::acc_t a[U::EPU];
#pragma unroll
      for (int i = 0; i < U::EPU; ++i) a[i] = e0 + i < count ? ld_elem<DT_IN>(in, e0 + i) : (typename U::acc_t)0;
      Codec<DT_IN, U::EPU>::pack(a, reinterpret_cast<uint32_t*>(v));
    }
#pragma unroll
    for (int j = 0; j < SY_MAXR; ++j)
      if (j < c.world) {
        int p = c.rank + j; if (p >= c.world) p -= c.world;
        char* dst = os_slot(c, p, parity, c.rank) + u * (size_t)(U::IN_V * 16);
#pragma unroll
        for (int k = 0; k < U::IN_V; ++k) st16(dst + k * 16, v[k]);
      }
  }
  __syncthreads();
  if ((int)threadIdx.x < c.world) {
    const int p = threadIdx.x;
    uint32_t* remote = reinterpret_cast<uint32_t*>(c.heap[p] + SY_OSFLAGS_OFF) + ((parity * SY_MAX_BLOCKS + blockIdx.x) * SY_MAXR + c.rank);
    __threadfence_system();
    st_release_sys(remote, seq);
    const uint32_t* local = reinterpret_cast<const uint32_t*>(c.heap[c.rank] + SY_OSFLAGS_OFF) + ((parity * SY_MAX_BLOCKS + blockIdx.x) * SY_MAXR + p);
    spin_until_ge(local, seq, c);
  }
  __syncthreads();
  // phase 2: reduce my mailboxes in rank order (bitwise identical on every rank)
  for (size_t u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
    V16 v[SY_MAXR][U::IN_V];
#pragma unroll
    for (int r = 0; r < SY_MAXR; ++r)
      if (r < c.world) {
        const char* src = os_slot(c, c.rank, parity, r) + u * (size_t)(U::IN_V * 16);
#pragma unroll
        for (int k = 0; k < U::IN_V; ++k) v[r][k] = ld16(src + k * 16);
      }
    typename U::acc_t acc[U::EPU];
    Codec<DT_IN, U::EPU>::unpack(reinterpret_cast<const uint32_t*>(v[0]), acc);
#pragma unroll
    for (int r = 1; r < SY_MAXR; ++r)
      if (r < c.world) {
        typename U::acc_t a[U::EPU];
        Codec<DT_IN, U::EPU>::unpack(reinterpret_cast<const uint32_t*>(v[r]), a);
#pragma unroll
        for (int i = 0; i < U::EPU; ++i) acc[i] = op_apply<OP>(acc[i], a[i]);
      }
    typename U::oacc_t o[U::EPU];
#pragma unroll
    for (int i = 0; i < U::EPU; ++i) {
      if (DT_IN == SY_F32 || DT_IN == SY_BF16 || DT_IN == SY_F16) o[i] = (typename U::oacc_t)((float)acc[i] * scale);
      else if (DT_IN == SY_F64) o[i] = (typename U::oacc_t)((double)acc[i] * (double)scale);
      else o[i] = (typename U::oacc_t)acc[i];
    }
    const size_t e0 = u * U::EPU;
    if (out_al && e0 + U::EPU <= count) {
      uint32_t w[U::OUT_V * 4];
      Codec<DT_OUT, U::EPU>::pack(o, w);
#pragma unroll
      for (int k = 0; k < U::OUT_V; ++k) st16((char*)out + u * (size_t)(U::OUT_V * 16) + k * 16, *reinterpret_cast<V16*>(&w[4 * k]));
    } else {
#pragma unroll
      for (int i = 0; i < U::EPU; ++i) if (e0 + i < count) st_elem<DT_OUT>(out, e0 + i, o[i]);
    }
  }
  seq_finish(c, 0);
}

// ---------------------------------------------------------------------------
// One-shot gradient all-reduce fused with Adam (the TensorFlow-Distributed recipe: ~80 k parameters = the latency-bound
// regime).  Every rank pushes its fp32 gradient into every peer's mailbox, waits for the per-block flags, reduces in rank
// order (bitwise identical everywhere), averages, and applies the Adam update to its own replica of the parameters in the
// same pass — one launch per training step instead of all-reduce + ~6 optimizer kernels.
// hyper (device, fp32): [lr, beta1, beta2, eps, step]; `step` is advanced by the last block to finish.
// count % 4 == 0, all pointers 16-byte aligned.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
k_oneshot_adam_k(const __grid_constant__ CommDev c, float* __restrict__ grad, float* __restrict__ param, float* __restrict__ m,
                 float* __restrict__ v, float* __restrict__ hyper, size_t count, float scale, int zero_grad) {
  const uint32_t seq = c.seq[0] + 1, parity = seq & 1;
  const size_t units = count / 4;
  const size_t per_block = (units + gridDim.x - 1) / gridDim.x;
  const size_t u0 = (size_t)blockIdx.x * per_block, u1 = (u0 + per_block < units) ? u0 + per_block : units;
  for (size_t u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
    const V16 g = ld16((const char*)grad + u * 16);
#pragma unroll
    for (int j = 0; j < SY_MAXR; ++j)
      if (j < c.world) {
        int p = c.rank + j; if (p >= c.world) p -= c.world;
        st16(os_slot(c, p, parity, c.rank) + u * 16, g);
      }
  }
  __syncthreads();
  if ((int)threadIdx.x < c.world) {
    const int p = threadIdx.x;
    uint32_t* remote = reinterpret_cast<uint32_t*>(c.heap[p] + SY_OSFLAGS_OFF) + ((parity * SY_MAX_BLOCKS + blockIdx.x) * SY_MAXR + c.rank);
    __threadfence_system();
    st_release_sys(remote, seq);
    const uint32_t* local = reinterpret_cast<const uint32_t*>(c.heap[c.rank] + SY_OSFLAGS_OFF) + ((parity * SY_MAX_BLOCKS + blockIdx.x) * SY_MAXR + p);
    spin_until_ge(local, seq, c);
  }
  __syncthreads();
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], t = hyper[4] + 1.0f;
  const float c1 = 1.0f - __powf(b1, t), c2 = 1.0f - __powf(b2, t);
  const float step_size = lr / c1, inv_sqrt_c2 = rsqrtf(c2);
  for (size_t u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < SY_MAXR; ++r)
      if (r < c.world) {
        const V16 q = ld16(os_slot(c, c.rank, parity, r) + u * 16);
        acc[0] += __uint_as_float(q.x); acc[1] += __uint_as_float(q.y); acc[2] += __uint_as_float(q.z); acc[3] += __uint_as_float(q.w);
      }
    float4 pm = reinterpret_cast<float4*>(m)[u], pv = reinterpret_cast<float4*>(v)[u], pp = reinterpret_cast<float4*>(param)[u];
    float* mm = &pm.x; float* vv = &pv.x; float* ppp = &pp.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float g = acc[i] * scale;
      mm[i] = b1 * mm[i] + (1.0f - b1) * g;
      vv[i] = b2 * vv[i] + (1.0f - b2) * g * g;
      ppp[i] -= step_size * mm[i] / (sqrtf(vv[i]) * inv_sqrt_c2 + eps);
    }
    reinterpret_cast<float4*>(m)[u] = pm; reinterpret_cast<float4*>(v)[u] = pv; reinterpret_cast<float4*>(param)[u] = pp;
    if (zero_grad) reinterpret_cast<float4*>(grad)[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // last block: bump the mailbox sequence AND the Adam step (every block has read both by now)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    uint32_t done = atomicAdd(&c.seq[8], 1u);
    if (done == gridDim.x - 1) { c.seq[8] = 0; c.seq[0] += 1; hyper[4] = t; __threadfence(); }
  }
}

// ---------------------------------------------------------------------------
// K2: LL all-reduce for tiny messages: 16B lines {w0, flag, w1, flag}; data and
// flag travel in the same 8-byte atom, so there is no fence and no barrier —
// one NVLink store latency end to end.  Single block.
// ---------------------------------------------------------------------------
template <int DT_IN, int DT_OUT, int OP>
__global__ void __launch_bounds__(1024)
k_ll(const __grid_constant__ CommDev c, const void* in, void* out, size_t count, float scale) {
  constexpr int SI = DtSize<DT_IN>::v;
  constexpr int EPL = 8 / SI;  // elements per 8-byte payload line
  using acc_t = typename Acc<DT_IN>::t;
  const uint32_t seq = c.seq[1] + 1, parity = seq & 1, flag = seq;
  const size_t lines = (count + EPL - 1) / EPL;
  for (size_t l = threadIdx.x; l < lines; l += blockDim.x) {
    acc_t a[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) a[i] = l * EPL + i < count ? ld_elem<DT_IN>(in, l * EPL + i) : (acc_t)0;
    uint32_t w[2];
    Codec<DT_IN, EPL>::pack(a, w);
    V16 line = {w[0], flag, w[1], flag};
#pragma unroll
    for (int j = 0; j < SY_MAXR; ++j)
      if (j < c.world) {
        int p = c.rank + j; if (p >= c.world) p -= c.world;
        st16_volatile(c.heap[p] + SY_LL_OFF + ((size_t)parity * SY_MAXR + c.rank) * SY_LL_SLOT + l * 16, line);
      }
  }
  for (size_t l = threadIdx.x; l < lines; l += blockDim.x) {
    acc_t acc[EPL];
    for (int r = 0; r < c.world; ++r) {
      const char* src = c.heap[c.rank] + SY_LL_OFF + ((size_t)parity * SY_MAXR + r) * SY_LL_SLOT + l * 16;
      V16 v; unsigned it = 0; unsigned long long t0 = 0;
      for (;;) {
        v = ld16_volatile(src);
        if (v.y == flag && v.w == flag) break;
        if (((++it) & 0x3ff) == 0) {
          unsigned long long t = globaltimer_ns();
          if (t0 == 0) t0 = t;
          else if (t - t0 > c.timeout_ns) { *reinterpret_cast<volatile uint32_t*>(c.status) = SY_ERR_TIMEOUT; __threadfence_system(); break; }
        }
      }
      uint32_t w[2] = {v.x, v.z};
      acc_t a[EPL];
      Codec<DT_IN, EPL>::unpack(w, a);
#pragma unroll
      for (int i = 0; i < EPL; ++i) acc[i] = r == 0 ? a[i] : op_apply<OP>(acc[i], a[i]);
    }
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      if (l * EPL + i < count) {
        if (DT_IN == SY_F32 || DT_IN == SY_BF16 || DT_IN == SY_F16) st_elem<DT_OUT>(out, l * EPL + i, (float)acc[i] * scale);
        else if (DT_IN == SY_F64) st_elem<DT_OUT>(out, l * EPL + i, (double)acc[i] * (double)scale);
        else st_elem<DT_OUT>(out, l * EPL + i, acc[i]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) c.seq[1] = seq;
}

// world == 1 (or local post-processing): out = cast(scale * in)
template <int DT_IN, int DT_OUT>
__global__ void k_local_scale_cast(const void* in, void* out, size_t count, float scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    auto v = ld_elem<DT_IN>(in, i);
    if (DT_IN == SY_F32 || DT_IN == SY_BF16 || DT_IN == SY_F16) st_elem<DT_OUT>(out, i, (float)v * scale);
    else if (DT_IN == SY_F64) st_elem<DT_OUT>(out, i, (double)v * (double)scale);
    else st_elem<DT_OUT>(out, i, v);
  }
}

// ---------------------------------------------------------------------------
// byte movers: all-gather / broadcast / all-to-all / gather / scatter
// ---------------------------------------------------------------------------
DEVI void copy_bytes_block(char* dst, const char* src, size_t bytes, size_t tid, size_t nth, bool mc) {
  const bool al = ((((uintptr_t)dst) | ((uintptr_t)src) | bytes) & 15) == 0;
  if (al) {
    size_t n = bytes / 16, i = tid;
    for (; i + 3 * nth < n; i += 4 * nth) {   // four 16B loads in flight per thread
      V16 a = ld16(src + i * 16), b = ld16(src + (i + nth) * 16), cc = ld16(src + (i + 2 * nth) * 16), d = ld16(src + (i + 3 * nth) * 16);
      if (mc) { mc_st16(dst + i * 16, a); mc_st16(dst + (i + nth) * 16, b); mc_st16(dst + (i + 2 * nth) * 16, cc); mc_st16(dst + (i + 3 * nth) * 16, d); }
      else { st16(dst + i * 16, a); st16(dst + (i + nth) * 16, b); st16(dst + (i + 2 * nth) * 16, cc); st16(dst + (i + 3 * nth) * 16, d); }
    }
    for (; i < n; i += nth) { V16 a = ld16(src + i * 16); if (mc) mc_st16(dst + i * 16, a); else st16(dst + i * 16, a); }
  } else if (!mc && ((((uintptr_t)dst) | ((uintptr_t)src) | bytes) & 3) == 0) {
    for (size_t i = tid; i < bytes / 4; i += nth) ((uint32_t*)dst)[i] = ((const uint32_t*)src)[i];
  } else {
    for (size_t i = tid; i < bytes; i += nth) dst[i] = src[i];
  }
}


// Deal the blocks of the grid to destination peers: with nb >= world every block
// owns one (peer, lane) pair so all NVLink destinations are written concurrently;
// smaller grids loop over peers.  Calls f(peer, lane, lanes).
template <typename F> DEVI void for_my_peers(const CommDev& c, F f) {
  const int nb = gridDim.x, w = c.world, b = blockIdx.x;
  if (nb >= w) {
    const int lanes = nb / w;
    if (b < lanes * w) { int p = c.rank + (b % w); if (p >= w) p -= w; f(p, b / w, lanes); }
  } else {
    for (int j = b; j < w; j += nb) { int p = c.rank + j; if (p >= w) p -= w; f(p, 0, 1); }
  }
}

// all-gather: my `bytes` -> offset rank*bytes of `out_off` in every heap
__global__ void __launch_bounds__(512)
k_allgather_k(const __grid_constant__ CommDev c, const void* in, size_t out_off, size_t bytes, int nvls) {
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);
  const bool al = ((((uintptr_t)in) | bytes | out_off) & 15) == 0;
  if (nvls && al && c.mc) {
    copy_bytes_block(c.mc + out_off + (size_t)c.rank * bytes, (const char*)in, bytes,
                     (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x, true);
  } else {
    for_my_peers(c, [&](int p, int lane, int lanes) {
      copy_bytes_block(c.heap[p] + out_off + (size_t)c.rank * bytes, (const char*)in, bytes,
                       (size_t)lane * blockDim.x + threadIdx.x, (size_t)lanes * blockDim.x, false);
    });
  }
  block_barrier(c, ep);
  epoch_store(c, ep);
}

__global__ void __launch_bounds__(512)
k_broadcast_k(const __grid_constant__ CommDev c, const void* in, size_t out_off, size_t bytes, int root, int nvls) {
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);
  if (c.rank == root) {
    const bool al = ((((uintptr_t)in) | bytes | out_off) & 15) == 0;
    if (nvls && al && c.mc) {
      copy_bytes_block(c.mc + out_off, (const char*)in, bytes, (size_t)blockIdx.x * blockDim.x + threadIdx.x,
                       (size_t)gridDim.x * blockDim.x, true);
    } else {
      for_my_peers(c, [&](int p, int lane, int lanes) {
        if (c.heap[p] + out_off == (const char*)in) return;  // in-place on root
        copy_bytes_block(c.heap[p] + out_off, (const char*)in, bytes, (size_t)lane * blockDim.x + threadIdx.x,
                         (size_t)lanes * blockDim.x, false);
      });
    }
  }
  block_barrier(c, ep);
  epoch_store(c, ep);
}

// Large broadcast (K7) as a pipelined scatter + all-gather in ONE kernel.  A single source cannot drive multimem.st above ~290 GB/s
// (profiles/coll_sweeps.md: 256 MB at 8 GPUs took 925 us, NCCL 473 us), but it can drive its NVLink egress with plain P2P stores:
// the root deals shard j (bytes / world) to rank j, and every rank re-broadcasts its own shard through the switch (multimem.st),
// so all eight egress links carry the all-gather while the root's link carries the scatter.  The data is cut into chunks; block b
// owns chunks b, b + grid, ...; per chunk: root scatters -> block barrier (same block index on every rank) -> every rank multicasts its
// piece.  While the other ranks multicast chunk k the root is already scattering chunk k + 1, so its link never idles:
// ~ bytes / link bandwidth in total instead of 2x.  Requires bytes % (world * 16) == 0 and 16-byte aligned buffers.
__global__ void __launch_bounds__(512)
k_broadcast_sag_k(const __grid_constant__ CommDev c, const void* in, size_t out_off, size_t bytes, int root, size_t chunk) {
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);                                   // nobody still reads `out`; the root's input is ready
  const size_t shard = bytes / (size_t)c.world;
  const size_t nchunks = (shard + chunk - 1) / chunk;
  for (size_t k = blockIdx.x; k < nchunks; k += gridDim.x) {
    const size_t o = k * chunk, n = shard - o < chunk ? shard - o : chunk;
    if (c.rank == root) {
      for (int j = 0; j < c.world; ++j) {                 // piece (j, k) -> rank j's own shard region (its own heap for j == root)
        int p = root + 1 + j; if (p >= c.world) p -= c.world;          // start with the next rank: all egress queues fill evenly
        char* dst = c.heap[p] + out_off + (size_t)p * shard + o;
        const char* src = (const char*)in + (size_t)p * shard + o;
        if (dst != src) copy_bytes_block(dst, src, n, threadIdx.x, blockDim.x, false);
      }
    }
    block_barrier(c, ep);                                 // piece (rank, k) has landed in my heap (release/acquire at system scope)
    copy_bytes_block(c.mc + out_off + (size_t)c.rank * shard + o, c.heap[c.rank] + out_off + (size_t)c.rank * shard + o, n,
                     threadIdx.x, blockDim.x, true);
  }
  // blocks with fewer chunks than the others still have to take part in the per-chunk barriers of their block index only, so no
  // cross-block coupling exists; one last barrier makes every multicast store visible before anybody returns
  block_barrier(c, ep);
  epoch_store(c, ep);
}

// all-to-all: block p of `in` -> offset rank*bytes of out in peer p (K4: uniform NVSwitch,
// so no ring schedule; all peers are written concurrently)
__global__ void __launch_bounds__(512)
k_alltoall_k(const __grid_constant__ CommDev c, const void* in, size_t out_off, size_t bytes) {
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);
  for_my_peers(c, [&](int p, int lane, int lanes) {
    copy_bytes_block(c.heap[p] + out_off + (size_t)c.rank * bytes, (const char*)in + (size_t)p * bytes, bytes,
                     (size_t)lane * blockDim.x + threadIdx.x, (size_t)lanes * blockDim.x, false);
  });
  block_barrier(c, ep);
  epoch_store(c, ep);
}

__global__ void __launch_bounds__(512)
k_gather_k(const __grid_constant__ CommDev c, const void* in, size_t out_off, size_t bytes, int root) {
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);
  copy_bytes_block(c.heap[root] + out_off + (size_t)c.rank * bytes, (const char*)in, bytes,
                   (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x, false);
  block_barrier(c, ep);
  epoch_store(c, ep);
}

// scatter (pull): out <- root's in[rank]
__global__ void __launch_bounds__(512)
k_scatter_k(const __grid_constant__ CommDev c, size_t in_off, void* out, size_t bytes, int root) {
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);
  copy_bytes_block((char*)out, c.heap[root] + in_off + (size_t)c.rank * bytes, bytes,
                   (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x, false);
  block_barrier(c, ep);
  epoch_store(c, ep);
}

// Mailbox byte movers for small / medium messages (<= 1 MB per writer): the sender pushes straight into its slot of the
// receiver's parity-double-buffered one-shot mailbox and raises one flag per block; the receiver waits for the flags of
// the writers it needs and copies the payload out of its own HBM.  No start barrier (nobody's `out` is touched remotely)
// and no end barrier (the next mailbox operation uses the other parity), so the cost is one NVLink store latency + a
// local copy instead of two cross-GPU barriers.
//   mode 0 all-gather : my `bytes` -> slot[me] on every rank; out[r*bytes ..] <- slot[r]
//   mode 1 all-to-all : in[p*bytes ..] -> slot[me] on rank p;  out[r*bytes ..] <- slot[r]
//   mode 2 broadcast  : root's `bytes` -> slot[root] on every rank; out <- slot[root]
//   mode 3 reduce-scatter (sum, fp32 accumulate): in[p*bytes ..] -> slot[me] on rank p; out <- scale * sum_r slot[r], summed in rank
//          order (bit-identical on every run); `root` carries the element type (SY_F32 / SY_BF16), `scale` the fused 1/N
__global__ void __launch_bounds__(512)
k_mailbox_k(const __grid_constant__ CommDev c, const char* __restrict__ in, char* __restrict__ out, size_t bytes, int mode, int root,
            float scale = 1.0f) {
  const uint32_t seq = c.seq[0] + 1, parity = seq & 1;
  const size_t units = bytes / 16;
  const size_t per_block = (units + gridDim.x - 1) / gridDim.x;
  const size_t u0 = (size_t)blockIdx.x * per_block, u1 = (u0 + per_block < units) ? u0 + per_block : units;
  const bool writer = mode != 2 || c.rank == root;
  if (writer) {
    for (size_t u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
      if (mode == 1 || mode == 3) {
        // all loads first, then all stores: ld16/st16 are volatile asm and keep their program order, so a load -> store -> load chain
        // would pay one cold HBM latency per peer (world x ~0.7 us per unit)
        V16 v[SY_MAXR];
#pragma unroll
        for (int j = 0; j < SY_MAXR; ++j)
          if (j < c.world) {
            int p = c.rank + j; if (p >= c.world) p -= c.world;
            v[j] = ld16(in + (size_t)p * bytes + u * 16);
          }
#pragma unroll
        for (int j = 0; j < SY_MAXR; ++j)
          if (j < c.world) {
            int p = c.rank + j; if (p >= c.world) p -= c.world;
            st16(os_slot(c, p, parity, c.rank) + u * 16, v[j]);
          }
      } else {
        const V16 v = ld16(in + u * 16);
#pragma unroll
        for (int j = 0; j < SY_MAXR; ++j)
          if (j < c.world) {
            int p = c.rank + j; if (p >= c.world) p -= c.world;
            st16(os_slot(c, p, parity, c.rank) + u * 16, v);
          }
      }
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < c.world) {
    const int p = threadIdx.x;
    // flags: writers tell every receiver "my block-b payload has landed".  In a broadcast EVERY rank flags every rank (the non-roots
    // without payload) and waits for all of them: the parity double-buffering of the mailboxes is only safe if finishing operation
    // k + 1 implies having heard from every peer in k + 1, i.e. that every peer has finished copying out of operation k.  (With
    // root-only flags a non-root could finish a broadcast, enter the next all-gather and overwrite a slot that a slower non-root
    // was still copying from.)
    {
      uint32_t* remote = reinterpret_cast<uint32_t*>(c.heap[p] + SY_OSFLAGS_OFF) + ((parity * SY_MAX_BLOCKS + blockIdx.x) * SY_MAXR + c.rank);
      __threadfence_system();
      st_release_sys(remote, seq);
      const uint32_t* local = reinterpret_cast<const uint32_t*>(c.heap[c.rank] + SY_OSFLAGS_OFF) + ((parity * SY_MAX_BLOCKS + blockIdx.x) * SY_MAXR + p);
      spin_until_ge(local, seq, c);
    }
  }
  __syncthreads();
  if (mode == 2) {
    for (size_t u = u0 + threadIdx.x; u < u1; u += blockDim.x) st16(out + u * 16, ld16(os_slot(c, c.rank, parity, root) + u * 16));
  } else if (mode == 3) {
    for (size_t u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
      V16 v[SY_MAXR];
#pragma unroll
      for (int r = 0; r < SY_MAXR; ++r) if (r < c.world) v[r] = ld16(os_slot(c, c.rank, parity, r) + u * 16);
      V16 o;
      if (root == SY_F32) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < SY_MAXR; ++r) if (r < c.world) {
          a[0] += __uint_as_float(v[r].x); a[1] += __uint_as_float(v[r].y); a[2] += __uint_as_float(v[r].z); a[3] += __uint_as_float(v[r].w);
        }
        o.x = __float_as_uint(a[0] * scale); o.y = __float_as_uint(a[1] * scale); o.z = __float_as_uint(a[2] * scale); o.w = __float_as_uint(a[3] * scale);
      } else {
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < SY_MAXR; ++r) if (r < c.world) {
          float t[8]; const uint32_t w4[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
          Codec<SY_BF16, 8>::unpack(w4, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] += t[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] *= scale;
        uint32_t w4[4]; Codec<SY_BF16, 8>::pack(a, w4);
        o.x = w4[0]; o.y = w4[1]; o.z = w4[2]; o.w = w4[3];
      }
      st16(out + u * 16, o);
    }
  } else {
    for (size_t u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
      V16 v[SY_MAXR];
#pragma unroll
      for (int r = 0; r < SY_MAXR; ++r) if (r < c.world) v[r] = ld16(os_slot(c, c.rank, parity, r) + u * 16);
#pragma unroll
      for (int r = 0; r < SY_MAXR; ++r) if (r < c.world) st16(out + (size_t)r * bytes + u * 16, v[r]);
    }
  }
  seq_finish(c, 0);
}

// ---------------------------------------------------------------------------
// Multi-block LL ("LM") kernel for the latency-bound range (<= 256 KB per writer): k_ll's line format — 16-byte lines of two 8-byte
// atoms {payload word, flag} — on a grid of blocks and for every all-hear-all collective.  Data and flag share one 8-byte store, so
// there is no fence, no separate flag store and no barrier: the receiver polls the lines themselves.  One NVLink store latency end
// to end where the mailbox kernel (k_mailbox_k) pays payload flight + fence + flag flight + a second pass over its own HBM.
// Lines live in a region that is only ever written in this format (SY_LM_OFF), double-buffered by the parity of the LL sequence
// (c.seq[1], shared with k_ll): a rank can enter operation k + 2 only after it has heard from every peer in operation k + 1, i.e.
// after every peer has finished reading operation k.
//   mode 0 all-gather      : in[l] -> slot[me] on every rank;        out[r * bytes + l] <- slot[r]
//   mode 1 all-to-all      : in[p * bytes + l] -> slot[me] on rank p; out[r * bytes + l] <- slot[r]
//   mode 3 reduce-scatter  : in[p * bytes + l] -> slot[me] on rank p; out[l] <- scale * sum_r slot[r]   (fp32 accumulate, rank order)
//   mode 4 all-reduce      : in[l] -> slot[me] on every rank;        out[l] <- scale * sum_r slot[r]
//   mode 2 broadcast       : root's in[l] -> slot[root] on every rank; out[l] <- slot[root]  (`dt` carries the root).  To stay
//          all-hear-all every rank also sends one token line to every rank when it enters the kernel and waits for all tokens
//          before it leaves.
// `bytes` (per writer) is a multiple of 4; in / out are 4-byte aligned; in == out is allowed (a thread reads every input word of its
// line index before it writes any output word of that index, and no other thread touches that index).
// ---------------------------------------------------------------------------
DEVI char* lm_slot(const CommDev& c, int on_rank, uint32_t parity, int writer) {
  return c.heap[on_rank] + SY_LM_OFF + ((size_t)parity * SY_MAXR + writer) * SY_LM_SLOT;
}
DEVI char* lm_token(const CommDev& c, int on_rank, uint32_t parity, int writer) {
  return c.heap[on_rank] + SY_LM_OFF + SY_LM_TOT + ((size_t)parity * SY_MAXR + writer) * 16;
}
DEVI V16 lm_poll(const CommDev& c, const char* src, uint32_t flag) {
  V16 v; unsigned it = 0; unsigned long long t0 = 0;
  for (;;) {
    v = ld16_volatile(src);
    if (v.y == flag && v.w == flag) break;
    if (((++it) & 0x3ff) == 0) {
      unsigned long long t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > c.timeout_ns) { *reinterpret_cast<volatile uint32_t*>(c.status) = SY_ERR_TIMEOUT; __threadfence_system(); break; }
    }
  }
  return v;
}
__global__ void __launch_bounds__(256)
k_lm_k(const __grid_constant__ CommDev c, const uint32_t* in, uint32_t* out, size_t bytes, int mode, int dt,
       float scale) {
  const uint32_t seq = c.seq[1] + 1, parity = seq & 1, flag = seq;
  const size_t words = bytes / 4, lines = (words + 1) / 2, wpr = words;      // wpr: words per rank block of in / out
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  const bool exchange = mode == 1 || mode == 3;
  if (mode == 2) {
    const int root = dt;
    if (blockIdx.x == 0 && (int)threadIdx.x < c.world) {
      const V16 token = {0u, flag, 0u, flag};
      st16_volatile(lm_token(c, (int)threadIdx.x, parity, c.rank), token);
    }
    for (size_t l = tid; l < lines; l += nth) {
      const bool two = 2 * l + 1 < words;
      if (c.rank == root) {
        const V16 line = {in[2 * l], flag, two ? in[2 * l + 1] : 0u, flag};
#pragma unroll
        for (int j = 0; j < SY_MAXR; ++j)
          if (j < c.world) {
            int p = c.rank + j; if (p >= c.world) p -= c.world;
            st16_volatile(lm_slot(c, p, parity, root) + l * 16, line);
          }
      }
      const V16 v = lm_poll(c, lm_slot(c, c.rank, parity, root) + l * 16, flag);
      out[2 * l] = v.x; if (two) out[2 * l + 1] = v.z;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < c.world) lm_poll(c, lm_token(c, c.rank, parity, (int)threadIdx.x), flag);
    seq_finish(c, 1);
    return;
  }
  for (size_t l = tid; l < lines; l += nth) {
    const bool two = 2 * l + 1 < words;
    // ---- send: all loads first (volatile asm keeps program order), then one 16-byte store per peer
    uint32_t w0[SY_MAXR], w1[SY_MAXR];
    if (exchange) {
#pragma unroll
      for (int j = 0; j < SY_MAXR; ++j)
        if (j < c.world) {
          int p = c.rank + j; if (p >= c.world) p -= c.world;
          w0[j] = in[(size_t)p * wpr + 2 * l]; w1[j] = two ? in[(size_t)p * wpr + 2 * l + 1] : 0u;
        }
    } else {
      w0[0] = in[2 * l]; w1[0] = two ? in[2 * l + 1] : 0u;
    }
#pragma unroll
    for (int j = 0; j < SY_MAXR; ++j)
      if (j < c.world) {
        int p = c.rank + j; if (p >= c.world) p -= c.world;
        const V16 line = {exchange ? w0[j] : w0[0], flag, exchange ? w1[j] : w1[0], flag};
        st16_volatile(lm_slot(c, p, parity, c.rank) + l * 16, line);
      }
    // ---- receive: poll my own slots (local HBM / L2), every writer's line in flight at once
    V16 v[SY_MAXR];
    uint32_t pending = (1u << c.world) - 1u;
    unsigned it = 0; unsigned long long t0 = 0;
    while (pending) {
#pragma unroll
      for (int r = 0; r < SY_MAXR; ++r)
        if (r < c.world && (pending >> r & 1u)) v[r] = ld16_volatile(lm_slot(c, c.rank, parity, r) + l * 16);
#pragma unroll
      for (int r = 0; r < SY_MAXR; ++r)
        if (r < c.world && (pending >> r & 1u) && v[r].y == flag && v[r].w == flag) pending &= ~(1u << r);
      if (pending && ((++it) & 0x3ff) == 0) {
        unsigned long long t = globaltimer_ns();
        if (t0 == 0) t0 = t;
        else if (t - t0 > c.timeout_ns) { *reinterpret_cast<volatile uint32_t*>(c.status) = SY_ERR_TIMEOUT; __threadfence_system(); break; }
      }
    }
    if (mode == 0 || mode == 1) {
#pragma unroll
      for (int r = 0; r < SY_MAXR; ++r)
        if (r < c.world) { out[(size_t)r * wpr + 2 * l] = v[r].x; if (two) out[(size_t)r * wpr + 2 * l + 1] = v[r].z; }
    } else if (dt == SY_F32) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int r = 0; r < SY_MAXR; ++r) if (r < c.world) { a0 += __uint_as_float(v[r].x); a1 += __uint_as_float(v[r].z); }
      out[2 * l] = __float_as_uint(a0 * scale); if (two) out[2 * l + 1] = __float_as_uint(a1 * scale);
    } else {                                    // bf16 pairs: low half = even element
      float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < SY_MAXR; ++r) if (r < c.world) {
        a[0] += __uint_as_float(v[r].x << 16); a[1] += __uint_as_float(v[r].x & 0xffff0000u);
        a[2] += __uint_as_float(v[r].z << 16); a[3] += __uint_as_float(v[r].z & 0xffff0000u);
      }
      uint32_t o[2]; float sc[4] = {a[0] * scale, a[1] * scale, a[2] * scale, a[3] * scale};
      Codec<SY_BF16, 4>::pack(sc, o);
      out[2 * l] = o[0]; if (two) out[2 * l + 1] = o[1];
    }
  }
  seq_finish(c, 1);
}

__global__ void k_barrier_k(const __grid_constant__ CommDev c) {
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);
  epoch_store(c, ep);
}

// rooted reduce: root pulls and reduces everything (P2P)
template <int DT, int OP>
__global__ void __launch_bounds__(512)
k_reduce_k(const __grid_constant__ CommDev c, size_t in_off, void* out, size_t count, int root) {
  using U = Unit<DT, DT>;
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);
  if (c.rank == root) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    const size_t units = ((((uintptr_t)out) & 15) == 0) ? count / U::EPU : 0;
    for (size_t u = tid; u < units; u += nth) {
      uint32_t w[U::OUT_V * 4];
      p2p_reduce_unit<DT, DT, OP>(c, in_off, u, 1.0f, w);
#pragma unroll
      for (int k = 0; k < U::OUT_V; ++k) st16((char*)out + u * (size_t)(U::OUT_V * 16) + k * 16, *reinterpret_cast<V16*>(&w[4 * k]));
    }
    for (size_t i = units * U::EPU + tid; i < count; i += nth)
      st_elem<DT>(out, i, p2p_reduce_elem<DT, DT, OP>(c, in_off, i, 1.0f));
  }
  block_barrier(c, ep);
  epoch_store(c, ep);
}

// ---------------------------------------------------------------------------
// K9: point-to-point put + signal, and fused halo pack + push + wait
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
k_put_signal_k(const __grid_constant__ CommDev c, const void* src, size_t dst_off, size_t bytes, int peer, int sig) {
  copy_bytes_block(c.heap[peer] + dst_off, (const char*)src, bytes, (size_t)blockIdx.x * blockDim.x + threadIdx.x,
                   (size_t)gridDim.x * blockDim.x, false);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    // the last block to finish raises the signal once
    uint32_t done = atomicAdd(&c.seq[10], 1u);
    if (done == gridDim.x - 1) {
      c.seq[10] = 0;
      red_add_release_sys(reinterpret_cast<uint32_t*>(c.heap[peer] + SY_SIG_OFF) + sig, 1u);
    }
  }
}
__global__ void k_wait_signal_k(const __grid_constant__ CommDev c, int sig, uint32_t expected) {
  if (threadIdx.x == 0) spin_until_ge(reinterpret_cast<const uint32_t*>(c.heap[c.rank] + SY_SIG_OFF) + sig, expected, c);
}

struct HaloArgs {
  int ndesc, nwait;
  sy_halo_desc d[26];
  int wait_sig[26];
};
// One block per boundary region: gather the strided region straight out of the
// field array and store it into the neighbour's ghost buffer over NVLink (no
// pack buffer, no separate copy), then signal; finally wait for my own
// neighbours' signals.  expected counts live in device memory (graph-safe).
// kHaloBPD blocks share one descriptor (a 512 KB face pushed by a single block is latency-bound at ~50 us); the last block
// of a descriptor to finish (device counter) raises the neighbour's signal.  Rows that are contiguous and 16-byte aligned
// move as 16-byte vectors.
constexpr int kHaloBPD = 8;
template <typename T>
__global__ void __launch_bounds__(256)
k_halo_k(const __grid_constant__ CommDev c, const T* __restrict__ src, const __grid_constant__ HaloArgs a,
         uint32_t* expect, uint32_t* done) {
  const int di = blockIdx.x / kHaloBPD, part = blockIdx.x % kHaloBPD;
  if (di < a.ndesc) {
    const sy_halo_desc& d = a.d[di];
    T* dst = reinterpret_cast<T*>(c.heap[d.peer] + d.dst_off);
    const long n = (long)d.nx * d.ny * d.nz;
    const long per = (n + kHaloBPD - 1) / kHaloBPD;
    const long i0 = (long)part * per, i1 = i0 + per < n ? i0 + per : n;
    constexpr int VE = 16 / sizeof(T);
    const bool contiguous = d.sx == 1 && (d.ny == 1 || d.sy == d.nx) && (d.nz == 1 || d.sz == (long)d.nx * d.ny);
    const bool vec_ok = contiguous && (((uintptr_t)(src + d.src_elem_off) | (uintptr_t)dst) & 15) == 0 && (per % VE) == 0;
    if (vec_ok) {
      const V16* s16 = reinterpret_cast<const V16*>(src + d.src_elem_off);
      V16* d16 = reinterpret_cast<V16*>(dst);
      const long v0 = i0 / VE, v1 = i1 / VE;
      long v = v0 + threadIdx.x;
      for (; v + 3 * (long)blockDim.x < v1; v += 4 * (long)blockDim.x) {        // four 16 B loads in flight per thread
        const V16 q0 = s16[v], q1 = s16[v + blockDim.x], q2 = s16[v + 2 * blockDim.x], q3 = s16[v + 3 * blockDim.x];
        d16[v] = q0; d16[v + blockDim.x] = q1; d16[v + 2 * blockDim.x] = q2; d16[v + 3 * blockDim.x] = q3;
      }
      for (; v < v1; v += blockDim.x) d16[v] = s16[v];
      for (long i = v1 * VE + threadIdx.x; i < i1; i += blockDim.x) dst[i] = src[d.src_elem_off + i];   // tail of the last part
    } else {
      for (long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        long x = i % d.nx, y = (i / d.nx) % d.ny, z = i / ((long)d.nx * d.ny);
        dst[i] = src[d.src_elem_off + x * d.sx + y * d.sy + z * d.sz];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(&done[di], 1u) == kHaloBPD - 1) {      // every part of this face has been pushed and fenced
        done[di] = 0;
        __threadfence_system();
        red_add_release_sys(reinterpret_cast<uint32_t*>(c.heap[d.peer] + SY_SIG_OFF) + d.sig_idx, 1u);
      }
    }
  }
  // waits: block 0 only; one thread per expected signal
  if (blockIdx.x == 0 && (int)threadIdx.x < a.nwait) {
    const int s = a.wait_sig[threadIdx.x];
    const uint32_t e = expect[s] + 1;
    spin_until_ge(reinterpret_cast<const uint32_t*>(c.heap[c.rank] + SY_SIG_OFF) + s, e, c);
    expect[s] = e;
  }
}

// ---------------------------------------------------------------------------
// fused gradient all-reduce + SGD(momentum) + parameter all-gather (ZeRO-1, one kernel)
// ---------------------------------------------------------------------------
template <int DT_G, int DT_P, bool NVLS>
__global__ void __launch_bounds__(512)
k_fused_sgd_k(const __grid_constant__ CommDev c, size_t g_off, size_t p_off, float* __restrict__ master,
              float* __restrict__ mom, const float* __restrict__ hyper, size_t count, int zero_grads) {
  constexpr int EPU = 8;                              // 8 elements per thread-iteration
  constexpr int GV = EPU * DtSize<DT_G>::v / 16;      // 1 (bf16) or 2 (f32) vectors
  constexpr int PV = EPU * DtSize<DT_P>::v / 16;
  const float lr = hyper[0], mu = hyper[1], wd = hyper[2], scale = hyper[3];
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);
  size_t ub, un;
  shard_of(count / EPU, c.world, c.rank, ub, un);
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (size_t k = tid; k < un; k += nth) {
    const size_t u = ub + k;
    float g[EPU];
    if (NVLS) {
      V16 s[GV];
#pragma unroll
      for (int v = 0; v < GV; ++v) s[v] = mc_ld_reduce<DT_G>(c.mc + g_off + u * (size_t)(GV * 16) + v * 16);
      Codec<DT_G, EPU>::unpack(reinterpret_cast<const uint32_t*>(s), g);
    } else {
      V16 s[SY_MAXR][GV];
#pragma unroll
      for (int j = 0; j < SY_MAXR; ++j)
        if (j < c.world) {
          int p = c.rank + j; if (p >= c.world) p -= c.world;
#pragma unroll
          for (int v = 0; v < GV; ++v) s[j][v] = ld16(c.heap[p] + g_off + u * (size_t)(GV * 16) + v * 16);
        }
#pragma unroll
      for (int i = 0; i < EPU; ++i) g[i] = 0.f;
#pragma unroll
      for (int j = 0; j < SY_MAXR; ++j)
        if (j < c.world) {
          float a[EPU];
          Codec<DT_G, EPU>::unpack(reinterpret_cast<const uint32_t*>(s[j]), a);
#pragma unroll
          for (int i = 0; i < EPU; ++i) g[i] += a[i];
        }
    }
    // optimizer state for this shard is local fp32: 2 x 16B each
    float4 m0 = reinterpret_cast<const float4*>(master)[2 * k], m1 = reinterpret_cast<const float4*>(master)[2 * k + 1];
    float4 v0 = reinterpret_cast<const float4*>(mom)[2 * k], v1 = reinterpret_cast<const float4*>(mom)[2 * k + 1];
    float w[EPU] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float b[EPU] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int i = 0; i < EPU; ++i) {
      float gi = fmaf(wd, w[i], g[i] * scale);
      b[i] = fmaf(mu, b[i], gi);
      w[i] = fmaf(-lr, b[i], w[i]);
    }
    reinterpret_cast<float4*>(master)[2 * k] = make_float4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<float4*>(master)[2 * k + 1] = make_float4(w[4], w[5], w[6], w[7]);
    reinterpret_cast<float4*>(mom)[2 * k] = make_float4(b[0], b[1], b[2], b[3]);
    reinterpret_cast<float4*>(mom)[2 * k + 1] = make_float4(b[4], b[5], b[6], b[7]);
    uint32_t pw[PV * 4];
    Codec<DT_P, EPU>::pack(w, pw);
    if (NVLS) {
#pragma unroll
      for (int v = 0; v < PV; ++v) mc_st16(c.mc + p_off + u * (size_t)(PV * 16) + v * 16, *reinterpret_cast<V16*>(&pw[4 * v]));
    } else {
#pragma unroll
      for (int j = 0; j < SY_MAXR; ++j)
        if (j < c.world) {
          int p = c.rank + j; if (p >= c.world) p -= c.world;
#pragma unroll
          for (int v = 0; v < PV; ++v) st16(c.heap[p] + p_off + u * (size_t)(PV * 16) + v * 16, *reinterpret_cast<V16*>(&pw[4 * v]));
        }
    }
  }
  block_barrier(c, ep);   // params visible everywhere; all peers done reading my grads
  epoch_store(c, ep);
  if (zero_grads) {       // clear my gradient buffer for the next accumulation
    const size_t nv = count * DtSize<DT_G>::v / 16;
    V16 z = {0, 0, 0, 0};
    for (size_t i = tid; i < nv; i += nth) st16(c.heap[c.rank] + g_off + i * 16, z);
  }
}

// ---------------------------------------------------------------------------
// K11: all-reduce with block-scaled fp8 output (MX: e4m3 + e8m0 per 32 elements).
// Reduce-scatter in fp32 registers, quantise own shard, broadcast q + scales.
// One thread owns one 32-element block; 4 adjacent threads merge their scale
// bytes into one 32-bit store.  count % 128 == 0.
// ---------------------------------------------------------------------------
template <int DT_IN, bool NVLS>
__global__ void __launch_bounds__(256)
k_ar_fp8_k(const __grid_constant__ CommDev c, size_t in_off, size_t q_off, size_t s_off, size_t count, float scale) {
  constexpr int SI = DtSize<DT_IN>::v;
  constexpr int NV = 32 * SI / 16;   // 16B vectors per 32-element block: 4 (bf16) / 8 (f32)
  uint32_t ep = epoch_load(c);
  block_barrier(c, ep);
  size_t gb, gn;                      // groups of 128 elements (4 blocks)
  shard_of(count / 128, c.world, c.rank, gb, gn);
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  const size_t nblk = gn * 4;
  // trip count is uniform over the whole grid (full-mask shuffles below); nth % 4 == 0 keeps
  // the 4 lanes of a 128-element group together
  const size_t iters = (nblk + nth - 1) / nth;
  for (size_t it = 0; it < iters; ++it) {
    const size_t k = it * nth + tid;
    const bool live = k < nblk;
    const size_t blk = gb * 4 + (live ? k : 0);
    float x[32];
    float amax = 0.f;
    if (live) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        constexpr int EPV = 16 / SI;
        float a[EPV];
        if (NVLS) {
          V16 s = mc_ld_reduce<DT_IN>(c.mc + in_off + (blk * NV + v) * 16);
          Codec<DT_IN, EPV>::unpack(reinterpret_cast<const uint32_t*>(&s), a);
        } else {
#pragma unroll
          for (int i = 0; i < EPV; ++i) a[i] = 0.f;
          for (int r = 0; r < c.world; ++r) {
            V16 s = ld16(c.heap[r] + in_off + (blk * NV + v) * 16);
            float t[EPV];
            Codec<DT_IN, EPV>::unpack(reinterpret_cast<const uint32_t*>(&s), t);
#pragma unroll
            for (int i = 0; i < EPV; ++i) a[i] += t[i];
          }
        }
#pragma unroll
        for (int i = 0; i < EPV; ++i) { x[v * EPV + i] = a[i] * scale; amax = fmaxf(amax, fabsf(x[v * EPV + i])); }
      }
    }
    // shared exponent: floor(log2(amax)) - 8 (e4m3 emax = 8), +1 when the mantissa exceeds 1.75
    int se = 0;
    if (amax > 0.f) {
      const uint32_t ab = __float_as_uint(amax);
      se = (int)((ab >> 23) & 0xff) - 8 + ((ab & 0x7fffffu) > 0x600000u ? 1 : 0);  // amax/scale <= 448
      if (se < 0) se = 0; if (se > 254) se = 254;
    }
    const float inv = __uint_as_float((uint32_t)(254 - se) << 23);   // 2^(127-se)
    uint32_t q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_fp8x4_e4m3 f(make_float4(x[4 * i] * inv, x[4 * i + 1] * inv, x[4 * i + 2] * inv, x[4 * i + 3] * inv));
      q[i] = *reinterpret_cast<uint32_t*>(&f);
    }
    // merge 4 scale bytes across the 4 lanes of the group
    uint32_t sb = (uint32_t)se << (8 * (threadIdx.x & 3));
    sb |= __shfl_xor_sync(0xffffffffu, sb, 1);
    sb |= __shfl_xor_sync(0xffffffffu, sb, 2);
    if (live) {
      V16 q0 = {q[0], q[1], q[2], q[3]}, q1 = {q[4], q[5], q[6], q[7]};
      if (NVLS) {
        mc_st16(c.mc + q_off + blk * 32, q0); mc_st16(c.mc + q_off + blk * 32 + 16, q1);
        if ((threadIdx.x & 3) == 0) mc_st4(c.mc + s_off + blk, sb);
      } else {
        for (int p = 0; p < c.world; ++p) {
          st16(c.heap[p] + q_off + blk * 32, q0); st16(c.heap[p] + q_off + blk * 32 + 16, q1);
          if ((threadIdx.x & 3) == 0) *reinterpret_cast<uint32_t*>(c.heap[p] + s_off + blk) = sb;
        }
      }
    }
  }
  block_barrier(c, ep);
  epoch_store(c, ep);
}

// ===========================================================================
// host launchers
// ===========================================================================
static inline int grid_for(sy_comm* c, size_t work_items, int threads, int per_thread = 1) {
  size_t b = (work_items + (size_t)threads * per_thread - 1) / ((size_t)threads * per_thread);
  if (b < 1) b = 1;
  if (b > (size_t)c->max_blocks) b = c->max_blocks;
  if (b > SY_MAX_BLOCKS) b = SY_MAX_BLOCKS;
  return (int)b;
}
#define LAUNCH_CHECK(c)                                                                     \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) { sy_set_error("kernel launch: %s", cudaGetErrorString(_e)); return SY_ERR_CUDA; } \
    (c)->launches++;                                                                        \
  } while (0)

static inline CommDev devof(sy_comm* c) {
  CommDev d = c->dev;
  d.timeout_ns = (unsigned long long)c->timeout_ms * 1000000ull;
  return d;
}

#define PAIR(a, b) ((a) * 16 + (b))

template <int DT_IN, int DT_OUT>
static int launch_ar_op(sy_comm* c, const void* in, void* out, size_t in_off, size_t out_off, size_t count,
                        float scale, int op, int algo, cudaStream_t s) {
  using U = Unit<DT_IN, DT_OUT>;
  CommDev d = devof(c);
  const int th = (int)c->threads;
#define OPSWITCH(KERNEL_CALL)                                                 \
  switch (op) {                                                               \
    case SY_SUM: { constexpr int OP = SY_SUM; KERNEL_CALL; break; }           \
    case SY_MAX: { constexpr int OP = SY_MAX; KERNEL_CALL; break; }           \
    case SY_MIN: { constexpr int OP = SY_MIN; KERNEL_CALL; break; }           \
    case SY_PROD: { constexpr int OP = SY_PROD; KERNEL_CALL; break; }         \
    default: return SY_ERR_ARG;                                               \
  }
  if (algo == SY_ALGO_LL) {
    int lth = 1024; size_t lines = (count * DtSize<DT_IN>::v + 7) / 8;
    if (lines < 1024) lth = (int)((lines + 31) / 32 * 32); if (lth < 32) lth = 32;
    OPSWITCH((k_ll<DT_IN, DT_OUT, OP><<<1, lth, 0, s>>>(d, in, out, count, scale)));
  } else if (algo == SY_ALGO_ONESHOT) {
    size_t units = (count + U::EPU - 1) / U::EPU;
    int g = grid_for(c, units, th);
    OPSWITCH((k_oneshot<DT_IN, DT_OUT, OP><<<g, th, 0, s>>>(d, in, out, count, scale)));
  } else if (algo == SY_ALGO_TWOSHOT_P2P) {
    size_t units = count / U::EPU / (size_t)c->world + 1;
    if (op == SY_SUM && c->world <= 2) {
      int g = grid_for(c, units, th, 4);
      k_twoshot_p2p<DT_IN, DT_OUT, SY_SUM, 2, 4><<<g, th, 0, s>>>(d, in_off, out_off, count, scale, 0, nullptr);
    } else if (op == SY_SUM && c->world <= 4) {
      int g = grid_for(c, units, th, 2);
      k_twoshot_p2p<DT_IN, DT_OUT, SY_SUM, 4, 2><<<g, th, 0, s>>>(d, in_off, out_off, count, scale, 0, nullptr);
    } else {
      int g = grid_for(c, units, th);
      OPSWITCH((k_twoshot_p2p<DT_IN, DT_OUT, OP, SY_MAXR, 1><<<g, th, 0, s>>>(d, in_off, out_off, count, scale, 0, nullptr)));
    }
  } else return SY_ERR_ARG;
#undef OPSWITCH
  LAUNCH_CHECK(c);
  return SY_OK;
}

template <int DT_IN, int DT_OUT>
static int launch_ar_nvls(sy_comm* c, size_t in_off, size_t out_off, size_t count, float scale, cudaStream_t s) {
  using U = Unit<DT_IN, DT_OUT>;
  size_t units = count / U::EPU / (size_t)c->world + 1;
  int g = grid_for(c, units, (int)c->threads, 4);
  k_twoshot_nvls<DT_IN, DT_OUT, 4><<<g, (int)c->threads, 0, s>>>(devof(c), in_off, out_off, count, scale, 0, nullptr);
  LAUNCH_CHECK(c);
  return SY_OK;
}

int k_allreduce(sy_comm* c, const void* in, void* out, size_t in_off, size_t out_off, bool, bool,
                size_t count, int dt_in, int dt_out, float scale, int op, int algo, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (c->world == 1 && algo != SY_ALGO_LL && algo != SY_ALGO_ONESHOT) algo = SY_ALGO_ONESHOT;
  if (algo == SY_ALGO_TWOSHOT_NVLS) {
    if (op != SY_SUM) return SY_ERR_UNSUPPORTED;
    switch (PAIR(dt_in, dt_out)) {
      case PAIR(SY_F32, SY_F32): return launch_ar_nvls<SY_F32, SY_F32>(c, in_off, out_off, count, scale, s);
      case PAIR(SY_BF16, SY_BF16): return launch_ar_nvls<SY_BF16, SY_BF16>(c, in_off, out_off, count, scale, s);
      case PAIR(SY_F16, SY_F16): return launch_ar_nvls<SY_F16, SY_F16>(c, in_off, out_off, count, scale, s);
      case PAIR(SY_BF16, SY_F32): return launch_ar_nvls<SY_BF16, SY_F32>(c, in_off, out_off, count, scale, s);
      case PAIR(SY_F32, SY_BF16): return launch_ar_nvls<SY_F32, SY_BF16>(c, in_off, out_off, count, scale, s);
      default: return SY_ERR_UNSUPPORTED;
    }
  }
  switch (PAIR(dt_in, dt_out)) {
    case PAIR(SY_F32, SY_F32): return launch_ar_op<SY_F32, SY_F32>(c, in, out, in_off, out_off, count, scale, op, algo, s);
    case PAIR(SY_BF16, SY_BF16): return launch_ar_op<SY_BF16, SY_BF16>(c, in, out, in_off, out_off, count, scale, op, algo, s);
    case PAIR(SY_F16, SY_F16): return launch_ar_op<SY_F16, SY_F16>(c, in, out, in_off, out_off, count, scale, op, algo, s);
    case PAIR(SY_F64, SY_F64): return launch_ar_op<SY_F64, SY_F64>(c, in, out, in_off, out_off, count, scale, op, algo, s);
    case PAIR(SY_I32, SY_I32): return launch_ar_op<SY_I32, SY_I32>(c, in, out, in_off, out_off, count, scale, op, algo, s);
    case PAIR(SY_I64, SY_I64): return launch_ar_op<SY_I64, SY_I64>(c, in, out, in_off, out_off, count, scale, op, algo, s);
    case PAIR(SY_BF16, SY_F32): return launch_ar_op<SY_BF16, SY_F32>(c, in, out, in_off, out_off, count, scale, op, algo, s);
    case PAIR(SY_F32, SY_BF16): return launch_ar_op<SY_F32, SY_BF16>(c, in, out, in_off, out_off, count, scale, op, algo, s);
    default: return SY_ERR_UNSUPPORTED;
  }
}

template <int DT_IN, int DT_OUT>
static int launch_rs(sy_comm* c, size_t in_off, void* out, size_t count, float scale, int op, bool nvls, cudaStream_t s) {
  using U = Unit<DT_IN, DT_OUT>;
  CommDev d = devof(c);
  int g = grid_for(c, count / U::EPU + 1, (int)c->threads, nvls ? 2 : 1);
  // vector stores into `out` need (rank*count) to start on a unit boundary and out 16B aligned
  bool vec_ok = (count % U::EPU == 0) && ((((uintptr_t)out) & 15) == 0);
  if (!vec_ok) nvls = false;
  constexpr bool kNvlsType = (DT_IN == SY_F32 || DT_IN == SY_BF16 || DT_IN == SY_F16);
  if (!kNvlsType) nvls = false;
  if (nvls) {
    if constexpr (kNvlsType)
      k_twoshot_nvls<DT_IN, DT_OUT, 4><<<g, (int)c->threads, 0, s>>>(d, in_off, 0, count, scale, 1, out);
  } else if (vec_ok && op == SY_SUM && c->world <= 2) {
    k_twoshot_p2p<DT_IN, DT_OUT, SY_SUM, 2, 4><<<g, (int)c->threads, 0, s>>>(d, in_off, 0, count, scale, 1, out);
  } else if (vec_ok && op == SY_SUM && c->world <= 4) {
    k_twoshot_p2p<DT_IN, DT_OUT, SY_SUM, 4, 2><<<g, (int)c->threads, 0, s>>>(d, in_off, 0, count, scale, 1, out);
  } else if (vec_ok) {
    switch (op) {
      case SY_SUM: k_twoshot_p2p<DT_IN, DT_OUT, SY_SUM, SY_MAXR, 1><<<g, (int)c->threads, 0, s>>>(d, in_off, 0, count, scale, 1, out); break;
      case SY_MAX: k_twoshot_p2p<DT_IN, DT_OUT, SY_MAX, SY_MAXR, 1><<<g, (int)c->threads, 0, s>>>(d, in_off, 0, count, scale, 1, out); break;
      case SY_MIN: k_twoshot_p2p<DT_IN, DT_OUT, SY_MIN, SY_MAXR, 1><<<g, (int)c->threads, 0, s>>>(d, in_off, 0, count, scale, 1, out); break;
      default: k_twoshot_p2p<DT_IN, DT_OUT, SY_PROD, SY_MAXR, 1><<<g, (int)c->threads, 0, s>>>(d, in_off, 0, count, scale, 1, out); break;
    }
  } else {
    return SY_ERR_UNSUPPORTED;  // caller falls back to the staged element path
  }
  LAUNCH_CHECK(c);
  return SY_OK;
}

int k_reduce_scatter(sy_comm* c, size_t in_off, void* out, size_t count, int dt_in, int dt_out, float scale,
                     int op, bool nvls, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (op != SY_SUM) nvls = false;
  switch (PAIR(dt_in, dt_out)) {
    case PAIR(SY_F32, SY_F32): return launch_rs<SY_F32, SY_F32>(c, in_off, out, count, scale, op, nvls, s);
    case PAIR(SY_BF16, SY_BF16): return launch_rs<SY_BF16, SY_BF16>(c, in_off, out, count, scale, op, nvls, s);
    case PAIR(SY_F16, SY_F16): return launch_rs<SY_F16, SY_F16>(c, in_off, out, count, scale, op, nvls, s);
    case PAIR(SY_BF16, SY_F32): return launch_rs<SY_BF16, SY_F32>(c, in_off, out, count, scale, op, nvls, s);
    case PAIR(SY_F32, SY_BF16): return launch_rs<SY_F32, SY_BF16>(c, in_off, out, count, scale, op, nvls, s);
    case PAIR(SY_F64, SY_F64): return launch_rs<SY_F64, SY_F64>(c, in_off, out, count, scale, op, false, s);
    case PAIR(SY_I32, SY_I32): return launch_rs<SY_I32, SY_I32>(c, in_off, out, count, scale, op, false, s);
    case PAIR(SY_I64, SY_I64): return launch_rs<SY_I64, SY_I64>(c, in_off, out, count, scale, op, false, s);
    default: return SY_ERR_UNSUPPORTED;
  }
}

// in/out: any device pointers (out is this rank's own buffer); bytes per writer, multiple of 16, <= SY_OS_SLOT
int k_mailbox(sy_comm* c, const void* in, void* out, size_t bytes, int mode, int root, void* stream, float scale) {
  int g = grid_for(c, bytes / 16 + 1, 256);
  k_mailbox_k<<<g, 256, 0, (cudaStream_t)stream>>>(devof(c), (const char*)in, (char*)out, bytes, mode, root, scale);
  LAUNCH_CHECK(c);
  return SY_OK;
}
// in/out: any device pointers, 4-byte aligned; bytes per writer: multiple of 4, <= SY_LM_MAX_PAYLOAD
int k_lm(sy_comm* c, const void* in, void* out, size_t bytes, int mode, int dt, void* stream, float scale) {
  int g = grid_for(c, (bytes / 4 + 1) / 2 + 1, 256);
  k_lm_k<<<g, 256, 0, (cudaStream_t)stream>>>(devof(c), (const uint32_t*)in, (uint32_t*)out, bytes, mode, dt, scale);
  LAUNCH_CHECK(c);
  return SY_OK;
}
int k_allgather(sy_comm* c, const void* in, size_t out_off, size_t count, int dt, bool nvls, void* stream) {
  size_t bytes = count * sy_dtype_size(dt);
  int g = grid_for(c, bytes / 16 + 1, (int)c->threads, 4);
  if (!nvls) { g = g / c->world * c->world; if (g < c->world) g = bytes >= (64u << 10) ? c->world : 1; }
  k_allgather_k<<<g, (int)c->threads, 0, (cudaStream_t)stream>>>(devof(c), in, out_off, bytes, nvls ? 1 : 0);
  LAUNCH_CHECK(c);
  return SY_OK;
}
int k_broadcast(sy_comm* c, const void* in, size_t out_off, size_t bytes, int root, bool nvls, void* stream) {
  // (the choice must be identical on every rank: it may only depend on symmetric quantities, never on the root's input pointer)
  if (nvls && c->has_mc && c->world >= 4 && bytes >= (size_t)c->bcast_sag_min_bytes && bytes % ((size_t)c->world * 16) == 0 &&
      (out_off & 15) == 0) {
    // every block must see the same number of per-chunk barriers on every rank: the chunk schedule depends only on (bytes, world, grid)
    const size_t shard = bytes / (size_t)c->world;
    size_t chunk = 128ul << 10;
    int blocks = (int)c->max_blocks < 64 ? (int)c->max_blocks : 64;
    const size_t nchunks = (shard + chunk - 1) / chunk;
    if ((size_t)blocks > nchunks) blocks = (int)nchunks;
    k_broadcast_sag_k<<<blocks, 512, 0, (cudaStream_t)stream>>>(devof(c), in, out_off, bytes, root, chunk);
    LAUNCH_CHECK(c);
    return SY_OK;
  }
  int g = grid_for(c, bytes / 16 + 1, (int)c->threads, 2);
  if (!nvls && g < c->world && bytes >= (64u << 10)) g = c->world;
  k_broadcast_k<<<g, (int)c->threads, 0, (cudaStream_t)stream>>>(devof(c), in, out_off, bytes, root, nvls ? 1 : 0);
  LAUNCH_CHECK(c);
  return SY_OK;
}
int k_alltoall(sy_comm* c, const void* in, size_t out_off, size_t bytes, void* stream) {
  int g = grid_for(c, bytes * c->world / 16 + 1, (int)c->threads, 2);
  if (g < c->world) g = c->world;
  k_alltoall_k<<<g, (int)c->threads, 0, (cudaStream_t)stream>>>(devof(c), in, out_off, bytes);
  LAUNCH_CHECK(c);
  return SY_OK;
}
int k_gather(sy_comm* c, const void* in, size_t out_off, size_t bytes, int root, void* stream) {
  int g = grid_for(c, bytes / 16 + 1, (int)c->threads, 2);
  k_gather_k<<<g, (int)c->threads, 0, (cudaStream_t)stream>>>(devof(c), in, out_off, bytes, root);
  LAUNCH_CHECK(c);
  return SY_OK;
}
int k_scatter(sy_comm* c, size_t in_off, void* out, size_t bytes, int root, void* stream) {
  int g = grid_for(c, bytes / 16 + 1, (int)c->threads, 2);
  k_scatter_k<<<g, (int)c->threads, 0, (cudaStream_t)stream>>>(devof(c), in_off, out, bytes, root);
  LAUNCH_CHECK(c);
  return SY_OK;
}
int k_barrier(sy_comm* c, void* stream) {
  k_barrier_k<<<1, 32, 0, (cudaStream_t)stream>>>(devof(c));
  LAUNCH_CHECK(c);
  return SY_OK;
}

template <int DT> static int launch_reduce(sy_comm* c, size_t in_off, void* out, size_t count, int op, int root, cudaStream_t s) {
  using U = Unit<DT, DT>;
  int g = grid_for(c, count / U::EPU + 1, (int)c->threads);
  CommDev d = devof(c);
  switch (op) {
    case SY_SUM: k_reduce_k<DT, SY_SUM><<<g, (int)c->threads, 0, s>>>(d, in_off, out, count, root); break;
    case SY_MAX: k_reduce_k<DT, SY_MAX><<<g, (int)c->threads, 0, s>>>(d, in_off, out, count, root); break;
    case SY_MIN: k_reduce_k<DT, SY_MIN><<<g, (int)c->threads, 0, s>>>(d, in_off, out, count, root); break;
    default: k_reduce_k<DT, SY_PROD><<<g, (int)c->threads, 0, s>>>(d, in_off, out, count, root); break;
  }
  LAUNCH_CHECK(c);
  return SY_OK;
}
int k_reduce_rooted(sy_comm* c, size_t in_off, void* out, size_t count, int dt, int op, int root, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  switch (dt) {
    case SY_F32: return launch_reduce<SY_F32>(c, in_off, out, count, op, root, s);
    case SY_BF16: return launch_reduce<SY_BF16>(c, in_off, out, count, op, root, s);
    case SY_F16: return launch_reduce<SY_F16>(c, in_off, out, count, op, root, s);
    case SY_F64: return launch_reduce<SY_F64>(c, in_off, out, count, op, root, s);
    case SY_I32: return launch_reduce<SY_I32>(c, in_off, out, count, op, root, s);
    case SY_I64: return launch_reduce<SY_I64>(c, in_off, out, count, op, root, s);
  }
  return SY_ERR_UNSUPPORTED;
}

int k_put_signal(sy_comm* c, const void* src, size_t dst_off, size_t bytes, int peer, int sig, void* stream) {
  int g = grid_for(c, bytes / 16 + 1, (int)c->threads, 2);
  k_put_signal_k<<<g, (int)c->threads, 0, (cudaStream_t)stream>>>(devof(c), src, dst_off, bytes, peer, sig);
  LAUNCH_CHECK(c);
  return SY_OK;
}
int k_wait_signal(sy_comm* c, int sig, uint32_t expected, void* stream) {
  k_wait_signal_k<<<1, 32, 0, (cudaStream_t)stream>>>(devof(c), sig, expected);
  LAUNCH_CHECK(c);
  return SY_OK;
}

int k_halo(sy_comm* c, const void* src, int dt, const sy_halo_desc* descs, int ndesc, const int* wait_sig,
           int nwait, void* stream) {
  if (ndesc > 26 || nwait > 26 || ndesc < 0 || nwait < 0) return SY_ERR_ARG;
  HaloArgs a; a.ndesc = ndesc; a.nwait = nwait;
  for (int i = 0; i < ndesc; ++i) a.d[i] = descs[i];
  for (int i = 0; i < nwait; ++i) a.wait_sig[i] = wait_sig[i];
  uint32_t* expect = c->dev.seq + 16;  // SY_NSIG expected counters follow the sequence words
  uint32_t* done = c->dev.seq + 16 + SY_NSIG;   // per-descriptor block counters (self-resetting)
  int g = ndesc > 0 ? ndesc * kHaloBPD : 1;
  CommDev d = devof(c);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dt) {
    case SY_F64: k_halo_k<double><<<g, 256, 0, s>>>(d, (const double*)src, a, expect, done); break;
    case SY_F32: k_halo_k<float><<<g, 256, 0, s>>>(d, (const float*)src, a, expect, done); break;
    case SY_BF16: case SY_F16: k_halo_k<uint16_t><<<g, 256, 0, s>>>(d, (const uint16_t*)src, a, expect, done); break;
    default: return SY_ERR_UNSUPPORTED;
  }
  LAUNCH_CHECK(c);
  return SY_OK;
}

int k_oneshot_adam(sy_comm* c, float* grad, float* param, float* m, float* v, float* hyper, size_t count, float scale, int zero_grad, void* stream) {
  if (count % 4 || (((uintptr_t)grad | (uintptr_t)param | (uintptr_t)m | (uintptr_t)v) & 15)) { sy_set_error("allreduce_adam: count %% 4, 16-byte aligned tensors"); return SY_ERR_ARG; }
  if (count * 4 > SY_OS_SLOT) { sy_set_error("allreduce_adam: gradient larger than the one-shot mailbox (1 MB)"); return SY_ERR_UNSUPPORTED; }
  CommDev d = devof(c);
  const int th = 256;
  int g = grid_for(c, count / 4, th);
  k_oneshot_adam_k<<<g, th, 0, (cudaStream_t)stream>>>(d, grad, param, m, v, hyper, count, scale, zero_grad);
  LAUNCH_CHECK(c);
  return SY_OK;
}

int k_fused_sgd(sy_comm* c, size_t g_off, int dt_grad, size_t p_off, int dt_param, float* master, float* mom,
                const float* hyper, size_t count, int zero_grads, void* stream) {
  if (count % 8) { sy_set_error("fused_sgd: count must be a multiple of 8"); return SY_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  CommDev d = devof(c);
  const bool nvls = c->has_mc && c->world >= c->nvls_min_world && !getenv("SHIPYARD_COLL_NO_NVLS");
  int g = grid_for(c, count / 8 / (size_t)c->world + 1, (int)c->threads);
  // zeroing the whole local gradient buffer wants a wide grid too
  if (zero_grads) { int gz = grid_for(c, count * sy_dtype_size(dt_grad) / 16 + 1, (int)c->threads, 4); if (gz > g) g = gz; }
  const int th = (int)c->threads;
#define FS(DG, DP)                                                                                          \
  do {                                                                                                      \
    if (nvls) k_fused_sgd_k<DG, DP, true><<<g, th, 0, s>>>(d, g_off, p_off, master, mom, hyper, count, zero_grads); \
    else k_fused_sgd_k<DG, DP, false><<<g, th, 0, s>>>(d, g_off, p_off, master, mom, hyper, count, zero_grads);     \
  } while (0)
  switch (PAIR(dt_grad, dt_param)) {
    case PAIR(SY_BF16, SY_BF16): FS(SY_BF16, SY_BF16); break;
    case PAIR(SY_F32, SY_F32): FS(SY_F32, SY_F32); break;
    case PAIR(SY_BF16, SY_F32): FS(SY_BF16, SY_F32); break;
    case PAIR(SY_F32, SY_BF16): FS(SY_F32, SY_BF16); break;
    default: return SY_ERR_UNSUPPORTED;
  }
#undef FS
  LAUNCH_CHECK(c);
  return SY_OK;
}

int k_allreduce_fp8(sy_comm* c, size_t in_off, int dt_in, void* out_q, void* out_scales, size_t count,
                    float scale, void* stream) {
  if (count % 128) { sy_set_error("allreduce_fp8: count must be a multiple of 128"); return SY_ERR_ARG; }
  char* base = c->dev.heap[c->rank];
  size_t q_off = (char*)out_q - base, s_off = (char*)out_scales - base;
  const bool nvls = c->has_mc && c->world >= c->nvls_min_world && !getenv("SHIPYARD_COLL_NO_NVLS");
  int g = grid_for(c, count / 32 / (size_t)c->world + 1, 256);
  CommDev d = devof(c);
  cudaStream_t s = (cudaStream_t)stream;
  if (dt_in == SY_BF16) { if (nvls) k_ar_fp8_k<SY_BF16, true><<<g, 256, 0, s>>>(d, in_off, q_off, s_off, count, scale); else k_ar_fp8_k<SY_BF16, false><<<g, 256, 0, s>>>(d, in_off, q_off, s_off, count, scale); }
  else if (dt_in == SY_F32) { if (nvls) k_ar_fp8_k<SY_F32, true><<<g, 256, 0, s>>>(d, in_off, q_off, s_off, count, scale); else k_ar_fp8_k<SY_F32, false><<<g, 256, 0, s>>>(d, in_off, q_off, s_off, count, scale); }
  else return SY_ERR_UNSUPPORTED;
  LAUNCH_CHECK(c);
  return SY_OK;
}

int k_local_cast(sy_comm* c, const void* in, void* out, size_t count, int dt_in, int dt_out, float scale, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int g = grid_for(c, count + 1, 256, 4);
#define LC(A, B) k_local_scale_cast<A, B><<<g, 256, 0, s>>>(in, out, count, scale)
  switch (PAIR(dt_in, dt_out)) {
    case PAIR(SY_F32, SY_F32): LC(SY_F32, SY_F32); break;
    case PAIR(SY_BF16, SY_BF16): LC(SY_BF16, SY_BF16); break;
    case PAIR(SY_F16, SY_F16): LC(SY_F16, SY_F16); break;
    case PAIR(SY_F64, SY_F64): LC(SY_F64, SY_F64); break;
    case PAIR(SY_I32, SY_I32): LC(SY_I32, SY_I32); break;
    case PAIR(SY_I64, SY_I64): LC(SY_I64, SY_I64); break;
    case PAIR(SY_BF16, SY_F32): LC(SY_BF16, SY_F32); break;
    case PAIR(SY_F32, SY_BF16): LC(SY_F32, SY_BF16); break;
    default: return SY_ERR_UNSUPPORTED;
  }
#undef LC
  LAUNCH_CHECK(c);
  return SY_OK;
}
