// Kernel launch helper: plain <<<>>> semantics, plus the programmatic-dependent-launch attribute when enabled for the
// calling thread (mmd_decoder_step turns it on around its launch sequence).
#pragma once
#include <cuda_runtime.h>

namespace mmd {

extern thread_local bool g_use_pdl;

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: "already done" flags are kept per device of the
// calling thread (one mmd_ctx per device may live in one process).
struct PerDeviceFlag {
  bool set[16] = {};
  bool& cur() {
    int d = 0;
    cudaGetDevice(&d);
    return set[d & 15];
  }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace mmd
