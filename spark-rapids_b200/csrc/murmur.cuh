// murmur.cuh — Spark's Murmur3_x86_32 + HashExpression per-type rules as device functions, shared by hash.cu
// (Hash.murmurHash32 / GpuHashPartitioning) and the fused partition->peer-store exchange kernel (exchange.cu).
// Reference: GpuMurmur3Hash.compute (HashFunctions.scala:196-209); the algorithm is Spark's
// org.apache.spark.unsafe.hash.Murmur3_x86_32 (external spec, restated in oracle/spark_hash.py, pinned by known answers).
#pragma once
#include "rowops.cuh"

namespace b2 {
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__device__ __forceinline__ uint32_t mix_k1(uint32_t k1) { k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u; return k1; }
__device__ __forceinline__ uint32_t mix_h1(uint32_t h1, uint32_t k1) { h1 ^= k1; h1 = rotl32(h1, 13); return h1 * 5u + 0xe6546b64u; }
__device__ __forceinline__ uint32_t fmix(uint32_t h1, uint32_t len) {
  h1 ^= len; h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
  return h1;
}
__device__ __forceinline__ uint32_t hash_int(uint32_t v, uint32_t seed) { return fmix(mix_h1(seed, mix_k1(v)), 4); }
__device__ __forceinline__ uint32_t hash_long(uint64_t v, uint32_t seed) {
  uint32_t h1 = mix_h1(seed, mix_k1((uint32_t)v));
  h1 = mix_h1(h1, mix_k1((uint32_t)(v >> 32)));
  return fmix(h1, 8);
}
// Murmur3_x86_32.hashUnsafeBytes: whole little-endian words, then each trailing byte (sign extended) as its own block
__device__ __forceinline__ uint32_t hash_bytes(const uint8_t* p, int len, uint32_t seed) {
  uint32_t h1 = seed;
  const int aligned = len & ~3;
  for (int i = 0; i < aligned; i += 4) {
    uint32_t w = (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24);
    h1 = mix_h1(h1, mix_k1(w));
  }
  for (int i = aligned; i < len; i++) h1 = mix_h1(h1, mix_k1((uint32_t)(int32_t)(int8_t)p[i]));
  return fmix(h1, (uint32_t)len);
}

__device__ __forceinline__ uint32_t murmur_col(const KeyCol& k, int64_t r, uint32_t seed) {
  if (!row_valid(k.valid, r)) return seed;
  switch (k.dtype) {
    case B2_BOOL8: return hash_int(reinterpret_cast<const int8_t*>(k.data)[r] != 0 ? 1u : 0u, seed);
    case B2_INT8: return hash_int((uint32_t)(int32_t)reinterpret_cast<const int8_t*>(k.data)[r], seed);
    case B2_INT16: return hash_int((uint32_t)(int32_t)reinterpret_cast<const int16_t*>(k.data)[r], seed);
    case B2_INT32: case B2_DATE32: return hash_int(reinterpret_cast<const uint32_t*>(k.data)[r], seed);
    case B2_INT64: case B2_TIMESTAMP_US: return hash_long(reinterpret_cast<const uint64_t*>(k.data)[r], seed);
    case B2_FLOAT32: {
      float f = reinterpret_cast<const float*>(k.data)[r];
      uint32_t b = (f != f) ? 0x7fc00000u : (f == 0.0f ? 0u : __float_as_uint(f));  // floatToIntBits, -0.0 -> 0.0
      return hash_int(b, seed);
    }
    case B2_FLOAT64: {
      double d = reinterpret_cast<const double*>(k.data)[r];
      uint64_t b = (d != d) ? 0x7ff8000000000000ull : (d == 0.0 ? 0ull : (uint64_t)__double_as_longlong(d));
      return hash_long(b, seed);
    }
    case B2_DECIMAL32: return hash_long((uint64_t)(int64_t)reinterpret_cast<const int32_t*>(k.data)[r], seed);
    case B2_DECIMAL64: return hash_long(reinterpret_cast<const uint64_t*>(k.data)[r], seed);
    case B2_DECIMAL128: {
      // precision > 18: hashUnsafeBytes(BigInteger.toByteArray()) = minimal big-endian two's complement
      const uint64_t* p = reinterpret_cast<const uint64_t*>(k.data) + 2 * r;
      uint8_t be[16];
#pragma unroll
      for (int i = 0; i < 8; i++) { be[i] = (uint8_t)(p[1] >> (8 * (7 - i))); be[8 + i] = (uint8_t)(p[0] >> (8 * (7 - i))); }
      const uint8_t sign = (be[0] & 0x80) ? 0xff : 0x00;
      int start = 0;
      while (start < 15 && be[start] == sign && ((be[start + 1] & 0x80) == (sign & 0x80))) start++;
      return hash_bytes(be + start, 16 - start, seed);
    }
    case B2_STRING: {
      const int32_t b = k.offsets[r], e = k.offsets[r + 1];
      return hash_bytes(reinterpret_cast<const uint8_t*>(k.data) + b, e - b, seed);
    }
  }
  return seed;
}

#endif
}  // namespace b2
