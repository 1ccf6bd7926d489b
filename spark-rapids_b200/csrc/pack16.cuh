// pack16.cuh — V shared-memory elements starting at an arbitrary (element-aligned) index, packed into one 16-byte vector in
// registers (the scatter kernels of hash.cu / exchange.cu store 16-byte vectors to aligned DESTINATIONS, so the source side
// of a vector is in general not 16-byte aligned)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
namespace b2 {
template <typename T> __device__ __forceinline__ uint4 pack16(const T* p);
template <> __device__ __forceinline__ uint4 pack16<uint64_t>(const uint64_t* p) {
  const uint64_t a = p[0], b = p[1];
  return make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
}
template <> __device__ __forceinline__ uint4 pack16<uint32_t>(const uint32_t* p) { return make_uint4(p[0], p[1], p[2], p[3]); }
template <> __device__ __forceinline__ uint4 pack16<uint16_t>(const uint16_t* p) {
  return make_uint4(p[0] | (uint32_t)p[1] << 16, p[2] | (uint32_t)p[3] << 16, p[4] | (uint32_t)p[5] << 16, p[6] | (uint32_t)p[7] << 16);
}
template <> __device__ __forceinline__ uint4 pack16<uint8_t>(const uint8_t* p) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; i++) w[i] = p[4 * i] | (uint32_t)p[4 * i + 1] << 8 | (uint32_t)p[4 * i + 2] << 16 | (uint32_t)p[4 * i + 3] << 24;
  return make_uint4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ uint4 pack16<uint4>(const uint4* p) { return p[0]; }
}  // namespace b2
