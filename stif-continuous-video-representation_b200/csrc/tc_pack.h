// tc_pack.h -- host-side packing of bf16 weight matrices into the shared-memory image the
// tcgen05 kernels consume (K-major, 128-byte swizzle; see tc_primitives.cuh for the layout).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "tc_primitives.cuh"

namespace stif {

inline float bf16_round_host(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}
inline uint16_t bf16_bits_host(float x) {
  float r = bf16_round_host(x);
  uint32_t u;
  std::memcpy(&u, &r, 4);
  return (uint16_t)(u >> 16);
}

// w: [rows, K] fp32 row-major, K a multiple of 64, rows a multiple of 8.
// Image: K/64 consecutive K-blocks, each `rows` x 128 bytes in SW128 order.  Appends to `img`.
inline void append_sw128_image(std::vector<uint8_t>& img, const float* w, int rows, int K) {
  const size_t base = img.size();
  img.resize(base + (size_t)rows * K * 2, 0);
  for (int kb = 0; kb < K / 64; ++kb)
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < 64; ++k) {
        uint16_t b = bf16_bits_host(w[(size_t)r * K + kb * 64 + k]);
        std::memcpy(&img[base + (size_t)kb * rows * 128 + tc::sw128_offset(r, k)], &b, 2);
      }
}

}  // namespace stif
