// nfb_save.cuh — transposed FP16 stores into the per-tile training record (nfb_layout.h: img_offset).
// A row thread owns one sample row r of a 128-row tile and holds 32 consecutive features of it as 16 packed f16x2
// (feature k0 + 2j in the low half of word j).  Element (k, r) of an image with `rows` features lives at
//   (r >> 6) * rows * 128 + k * 128 + ((((r & 63) >> 3) ^ (k & 7)) << 4) + (r & 7) * 2,
// so the 32 lanes of a warp (consecutive r) fill 64 bytes of one 128-byte line per feature.
#pragma once
#include <stdint.h>

namespace nfb {

// Byte offset of this thread's sample row inside an image, excluding the feature-dependent part.
__device__ __forceinline__ uint32_t img_row_base(int rows, int r) { return (uint32_t)((r >> 6) * rows * 128 + (r & 7) * 2); }

// k0 must be a multiple of 8 (so (k0 + j) & 7 == j & 7).
__device__ __forceinline__ void store_t32(uint8_t* __restrict__ img_row /* image + img_row_base */, int r, int k0,
                                          const uint32_t (&h)[16]) {
  const uint32_t cr = (uint32_t)((r & 63) >> 3);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int ka = 2 * j, kb = 2 * j + 1;
    uint8_t* pa = img_row + (size_t)(k0 + ka) * 128 + ((cr ^ (uint32_t)(ka & 7)) << 4);
    uint8_t* pb = img_row + (size_t)(k0 + kb) * 128 + ((cr ^ (uint32_t)(kb & 7)) << 4);
    *reinterpret_cast<uint16_t*>(pa) = (uint16_t)(h[j] & 0xFFFFu);
    *reinterpret_cast<uint16_t*>(pb) = (uint16_t)(h[j] >> 16);
  }
}

}  // namespace nfb
