// Host-side plumbing shared by the kernel launchers: error capture and TMA tensor-map creation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace b200w {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define B200W_CUDA(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      throw ::b200w::Error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +     \
                           " (" __FILE__ ":" + std::to_string(__LINE__) + ")");            \
  } while (0)

#define B200W_CHECK(cond, msg)                                                             \
  do {                                                                                     \
    if (!(cond))                                                                           \
      throw ::b200w::Error(std::string("check failed: " #cond " — ") + (msg) +             \
                           " (" __FILE__ ":" + std::to_string(__LINE__) + ")");            \
  } while (0)

// 2-D bf16 row-major tensor [rows, cols] with leading dimension ld (elements).
// Box = [box_rows, box_cols]; box_cols * 2 bytes must be <= 128 (SWIZZLE_128B span).
// Out-of-bounds elements are zero-filled by the hardware (used for K / M / N tails).
CUtensorMap make_tmap_bf16_2d(const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                              uint32_t box_rows, uint32_t box_cols);

int sm_count();

}  // namespace b200w
