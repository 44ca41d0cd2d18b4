// Host-side scalar conversions for the stub transport (bf16/f16/e4m3/e8m0).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "sy_coll.h"

namespace syn {

static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f;
}
static inline uint16_t f32_to_bf16(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // quiet NaN
  u += 0x7fffu + ((u >> 16) & 1u);  // round to nearest even
  return (uint16_t)(u >> 16);
}
static inline float f16_to_f32(uint16_t h) {
  uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ff, u;
  if (e == 0) {
    if (m == 0) u = s;
    else { int sh = 0; while (!(m & 0x400)) { m <<= 1; ++sh; } m &= 0x3ff; u = s | ((127 - 15 - sh + 1) << 23) | (m << 13); }
  } else if (e == 31) u = s | 0x7f800000u | (m << 13);
  else u = s | ((e - 15 + 127) << 23) | (m << 13);
  float f; memcpy(&f, &u, 4); return f;
}
static inline uint16_t f32_to_f16(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  uint32_t s = (u >> 16) & 0x8000u; int32_t e = (int32_t)((u >> 23) & 0xff) - 127 + 15; uint32_t m = u & 0x7fffffu;
  if (((u >> 23) & 0xff) == 0xff) return (uint16_t)(s | 0x7c00u | (m ? 0x200u : 0));
  if (e >= 31) return (uint16_t)(s | 0x7c00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t)s;
    m |= 0x800000u; int sh = 14 - e; uint32_t r = m >> sh; uint32_t rem = m & ((1u << sh) - 1), half = 1u << (sh - 1);
    if (rem > half || (rem == half && (r & 1))) ++r;
    return (uint16_t)(s | r);
  }
  uint32_t r = (uint32_t)(e << 10) | (m >> 13); uint32_t rem = m & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) ++r;
  return (uint16_t)(s | r);
}

// e4m3fn: 1-4-3, bias 7, max 448, no inf, NaN = 0x7f
static inline uint8_t f32_to_e4m3(float f) {
  if (f != f) return 0x7f;
  uint8_t s = f < 0 ? 0x80 : 0; float a = fabsf(f);
  if (a >= 448.f) return s | 0x7e;
  if (a < 0.0009765625f) return s;  // < 2^-10 (half of min subnormal 2^-9) rounds to 0
  int e; float m = frexpf(a, &e);   // a = m*2^e, m in [0.5,1)
  int E = e - 1 + 7;                // biased exponent with 1.xxx mantissa
  if (E <= 0) {                     // subnormal: value = k * 2^-9, k in 0..7
    float k = a * 512.f; float r = nearbyintf(k);
    if (r >= 8.f) return s | 0x08;
    return s | (uint8_t)r;
  }
  float frac = m * 2.f - 1.f;       // [0,1)
  float r = nearbyintf(frac * 8.f);
  if (r >= 8.f) { r = 0.f; ++E; if (E > 15) return s | 0x7e; }
  uint8_t v = (uint8_t)((E << 3) | (int)r);
  if (v > 0x7e) v = 0x7e;
  return s | v;
}
static inline float e4m3_to_f32(uint8_t v) {
  float s = (v & 0x80) ? -1.f : 1.f; int E = (v >> 3) & 0xf, m = v & 7;
  if (E == 15 && m == 7) return NAN;
  if (E == 0) return s * (float)m * 0.001953125f;  // 2^-9
  return s * ldexpf(1.f + m / 8.f, E - 7);
}
// e8m0 block scale: value 2^(e-127); choose so that amax/scale <= 448 (MX convention:
// shared exponent = floor(log2(amax)) - emax_elem, emax_elem(e4m3) = 8)
static inline uint8_t e8m0_from_amax(float amax) {
  if (!(amax > 0.f)) return 127 - 8 < 0 ? 0 : 0;  // all-zero block: smallest scale
  int e; float m = frexpf(amax, &e);  // amax = m*2^e, m in [0.5,1) => floor(log2) = e-1
  int se = (e - 1) - 8 + 127;
  if (m > 0.875f) se += 1;            // keep amax/scale <= 448 (1.75 * 2^8): no saturation loss
  if (se < 0) se = 0;
  if (se > 254) se = 254;
  return (uint8_t)se;
}
static inline float e8m0_inv_scale(uint8_t e) { return ldexpf(1.f, 127 - (int)e); }
static inline float e8m0_scale(uint8_t e) { return ldexpf(1.f, (int)e - 127); }

static inline float load_f32(const void* p, size_t i, int dt) {
  switch (dt) {
    case SY_F32: return ((const float*)p)[i];
    case SY_BF16: return bf16_to_f32(((const uint16_t*)p)[i]);
    case SY_F16: return f16_to_f32(((const uint16_t*)p)[i]);
    case SY_F64: return (float)((const double*)p)[i];
    case SY_I32: return (float)((const int32_t*)p)[i];
    case SY_I64: return (float)((const int64_t*)p)[i];
    case SY_U8: return (float)((const uint8_t*)p)[i];
  }
  return 0.f;
}
static inline int64_t load_i64(const void* p, size_t i, int dt) {
  switch (dt) {
    case SY_I32: return ((const int32_t*)p)[i];
    case SY_I64: return ((const int64_t*)p)[i];
    case SY_U8: return ((const uint8_t*)p)[i];
    default: return (int64_t)load_f32(p, i, dt);
  }
}
static inline void store_f(void* p, size_t i, int dt, double v) {
  switch (dt) {
    case SY_F32: ((float*)p)[i] = (float)v; break;
    case SY_BF16: ((uint16_t*)p)[i] = f32_to_bf16((float)v); break;
    case SY_F16: ((uint16_t*)p)[i] = f32_to_f16((float)v); break;
    case SY_F64: ((double*)p)[i] = v; break;
    case SY_I32: ((int32_t*)p)[i] = (int32_t)v; break;
    case SY_I64: ((int64_t*)p)[i] = (int64_t)v; break;
    case SY_U8: ((uint8_t*)p)[i] = (uint8_t)v; break;
  }
}
static inline void store_i(void* p, size_t i, int dt, int64_t v) {
  switch (dt) {
    case SY_I32: ((int32_t*)p)[i] = (int32_t)v; break;
    case SY_I64: ((int64_t*)p)[i] = v; break;
    case SY_U8: ((uint8_t*)p)[i] = (uint8_t)v; break;
    default: store_f(p, i, dt, (double)v);
  }
}

}  // namespace syn
