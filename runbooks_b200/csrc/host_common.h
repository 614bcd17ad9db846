// Host-side plumbing shared by the kernel launchers: error capture and TMA tensor-map creation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>

#include <nvtx3/nvToolsExt.h>

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

// SM count of the CURRENT device (cached per device: a process may hold contexts on several GPUs)
int sm_count();

// First-use work that is per DEVICE, not per process: cudaFuncSetAttribute(MaxDynamicSharedMemorySize)
// applies to the current device only, so a process-wide `static bool` breaks the second GPU of a
// one-process-many-contexts host (INTEGRATION.md's cgo layout).
//   static PerDeviceOnce once; once.run([&] { cudaFuncSetAttribute(...); });
class PerDeviceOnce {
 public:
  template <typename F>
  void run(F&& f) {
    int dev = 0;
    B200W_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    std::lock_guard<std::mutex> lock(mu_);
    if (done_ & bit) return;
    f();
    done_ |= bit;
  }

 private:
  std::mutex mu_;
  uint64_t done_ = 0;
};

// Kernel launch with the programmatic-stream-serialization attribute (see ptx.cuh pdl_wait): only for
// kernels that call pdl_wait() before they touch anything their predecessor wrote.
// cluster_y > 1 additionally groups that many consecutive blockIdx.y CTAs into a thread-block cluster.
template <typename... KArgs, typename... Args>
void launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_y,
                        Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  int n = 1;
  if (cluster_y > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 1;
    attr[1].val.clusterDim.y = static_cast<unsigned>(cluster_y);
    attr[1].val.clusterDim.z = 1;
    n = 2;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  B200W_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...));
}
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  launch_pdl_cluster(kern, grid, block, smem, s, 1, std::forward<Args>(args)...);
}

// NVTX range for the phases of a step (header-only NVTX v3: a no-op costing one predictable branch unless a tool
// such as nsys / ncu --nvtx is attached). SURVEY.md 5, tracing row.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// gemm_bf16 leaves this many SMs free (grid = SMs - reserve): set around the backward that runs
// concurrently with the NCCL gradient all-reduce, so that NCCL's CTAs find SMs without waiting for a
// persistent GEMM CTA to end and the GEMM's CTAs never queue behind NCCL's. Per host thread.
void gemm_set_sm_reserve(int n);
int gemm_sm_reserve();

}  // namespace b200w
