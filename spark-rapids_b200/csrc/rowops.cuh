// rowops.cuh — row hashing / equality over a set of key columns, shared by hash aggregate (a4),
// hash join (a6) and distinct count.  Semantics follow the reference's key rules:
//  * group-by: NULL is its own group, NaN == NaN, -0.0 == 0.0 (GpuAggregateExec.scala:565-568,
//    NormalizeFloatingNumbers.scala:29-38);
//  * join: NULL never matches unless compareNullsEqual (GpuHashJoin.scala:602-640).
// The hash here is internal (never observable); Spark-visible murmur3 lives in hash.cu.
#pragma once
#include "common.cuh"

namespace b2 {

constexpr int MAX_KEYS = 8;
struct KeyCol {
  const void* data;
  const uint32_t* valid;
  const int32_t* offsets;
  int32_t dtype;
  int32_t width;
  int32_t pack;   // bytes this column takes in a packed 8-byte group key (0 = not packable); strings: 1 length byte + chars
  int32_t pad;
};
struct KeyCols {
  int32_t n;
  KeyCol c[MAX_KEYS];
};

inline KeyCols key_cols_of(const Table* t, const int* idx, int n) {
  if (n > MAX_KEYS) throw Error(B2_ERR_UNSUPPORTED, "more than 8 key columns");
  KeyCols k; memset(&k, 0, sizeof(k));
  k.n = n;
  for (int i = 0; i < n; i++) {
    if (idx[i] < 0 || idx[i] >= (int)t->cols.size()) throw Error(B2_ERR_INVALID, "key column index out of range");
    const Column* c = t->cols[idx[i]];
    k.c[i].data = c->data.p; k.c[i].valid = c->validity(); k.c[i].offsets = c->offsets.as<int32_t>();
    k.c[i].dtype = c->dtype; k.c[i].width = dtype_width(c->dtype);
  }
  return k;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

__device__ __forceinline__ uint64_t key_bits(const KeyCol& k, int64_t r) {
  switch (k.dtype) {
    case B2_FLOAT32: {
      float f = reinterpret_cast<const float*>(k.data)[r];
      if (f != f) return 0x7fc00000u;
      if (f == 0.0f) return 0;
      return __float_as_uint(f);
    }
    case B2_FLOAT64: {
      double d = reinterpret_cast<const double*>(k.data)[r];
      if (d != d) return 0x7ff8000000000000ull;
      if (d == 0.0) return 0;
      return (uint64_t)__double_as_longlong(d);
    }
    default:
      switch (k.width) {
        case 1: return reinterpret_cast<const uint8_t*>(k.data)[r];
        case 2: return reinterpret_cast<const uint16_t*>(k.data)[r];
        case 4: return reinterpret_cast<const uint32_t*>(k.data)[r];
        default: return reinterpret_cast<const uint64_t*>(k.data)[r];
      }
  }
}

__device__ __forceinline__ uint32_t row_hash(const KeyCols& ks, int64_t r) {
  uint64_t h = 0x9e3779b97f4a7c15ull;
  for (int i = 0; i < ks.n; i++) {
    const KeyCol& k = ks.c[i];
    if (!row_valid(k.valid, r)) { h = mix64(h ^ 0x5bd1e995u); continue; }
    if (k.dtype == B2_STRING) {
      const int32_t b = k.offsets[r], e = k.offsets[r + 1];
      const uint8_t* p = reinterpret_cast<const uint8_t*>(k.data);
      uint64_t s = 0xcbf29ce484222325ull;
      for (int32_t q = b; q < e; q++) { s ^= p[q]; s *= 0x100000001b3ull; }
      h = mix64(h ^ s ^ (uint64_t)(e - b));
    } else if (k.width == 16) {
      const uint64_t* p = reinterpret_cast<const uint64_t*>(k.data) + 2 * r;
      h = mix64(h ^ p[0]); h = mix64(h ^ p[1]);
    } else {
      h = mix64(h ^ key_bits(k, r));
    }
  }
  return (uint32_t)(h ^ (h >> 32));
}

// rows ra of a and rb of b (same schema)
__device__ __forceinline__ bool rows_equal(const KeyCols& a, int64_t ra, const KeyCols& b, int64_t rb, bool nulls_equal) {
  for (int i = 0; i < a.n; i++) {
    const KeyCol& x = a.c[i];
    const KeyCol& y = b.c[i];
    const bool vx = row_valid(x.valid, ra), vy = row_valid(y.valid, rb);
    if (!vx || !vy) {
      if (vx != vy || !nulls_equal) return false;
      continue;
    }
    if (x.dtype == B2_STRING) {
      const int32_t bx = x.offsets[ra], lx = x.offsets[ra + 1] - bx;
      const int32_t by = y.offsets[rb], ly = y.offsets[rb + 1] - by;
      if (lx != ly) return false;
      const uint8_t* px = reinterpret_cast<const uint8_t*>(x.data) + bx;
      const uint8_t* py = reinterpret_cast<const uint8_t*>(y.data) + by;
      for (int32_t q = 0; q < lx; q++) if (px[q] != py[q]) return false;
    } else if (x.width == 16) {
      const uint64_t* px = reinterpret_cast<const uint64_t*>(x.data) + 2 * ra;
      const uint64_t* py = reinterpret_cast<const uint64_t*>(y.data) + 2 * rb;
      if (px[0] != py[0] || px[1] != py[1]) return false;
    } else {
      if (key_bits(x, ra) != key_bits(y, rb)) return false;
    }
  }
  return true;
}
#endif

}  // namespace b2
