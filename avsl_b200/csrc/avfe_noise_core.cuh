// The pairwise-summation tree of numpy's float32 add.reduce, addressed like a heap (see
// avfe_noise.cu).  __host__ __device__ so that tests/hostcheck can walk it on the CPU.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace avfe {

constexpr int kLeafMax = 128;       // numpy's PW_BLOCKSIZE
constexpr int kMaxDepth = 20;       // 2^21 heap slots per waveform; covers > 10^8 samples

// state of heap node k for an n-element array: 0 = does not exist, 1 = leaf piece, 2 = inner node
__host__ __device__ inline int locate_node(uint32_t k, uint32_t n, uint32_t& off, uint32_t& len) {
  int depth = 0;
  for (uint32_t t = k; t > 1; t >>= 1) ++depth;
  off = 0;
  len = n;
  for (int d = depth - 1; d >= 0; --d) {
    if (len <= (uint32_t)kLeafMax) return 0;
    const uint32_t half = (len >> 1) & ~7u;
    if ((k >> d) & 1u) {
      off += half;
      len -= half;
    } else {
      len = half;
    }
  }
  return len <= (uint32_t)kLeafMax ? 1 : 2;
}

// The leaf piece that holds element `pos` (< n): heap index, offset and length.
__host__ __device__ inline uint32_t locate_piece(uint32_t pos, uint32_t n, uint32_t& off, uint32_t& len) {
  uint32_t k = 1;
  off = 0;
  len = n;
  while (len > (uint32_t)kLeafMax) {
    const uint32_t half = (len >> 1) & ~7u;
    if (pos < off + half) {
      len = half;
      k = 2 * k;
    } else {
      off += half;
      len -= half;
      k = 2 * k + 1;
    }
  }
  return k;
}

// A depth at which every node of every clip up to max_len samples is a leaf: the larger child of
// an m-element node has at most m/2 + 8 elements.
inline int tree_depth(int64_t max_len) {
  int d = 0;
  for (int64_t m = max_len; m > kLeafMax; m = (m + 1) / 2 + 8) ++d;
  return d;
}

}  // namespace avfe
