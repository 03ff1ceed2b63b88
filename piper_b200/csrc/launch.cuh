// Kernel launch helper with programmatic dependent launch (PDL).
//
// The pipeline is ~150 short kernels on one stream; without PDL every boundary costs the full launch latency plus the next
// kernel's prologue (TMEM allocation, mbarrier initialisation, tensor-map prefetch) after the previous grid has drained.
// With the programmatic-stream-serialisation attribute the next grid may be scheduled as soon as every CTA of the current
// one has executed `griddepcontrol.launch_dependents` (all kernels do so first thing); its CTAs run their prologue and then
// block in `griddepcontrol.wait`, which returns only when the prerequisite grid has COMPLETED and its memory is visible -
// so every kernel keeps plain stream-order semantics for everything it reads or writes after that point.
// Measured on the B200 (profiles/r02_ab.md): no gain on the headline step (414 vs 423 M samples/s with graphs, 420 vs 418
// without) - the step is bound by the kernels, not by the gaps between them - so it is OFF by default; PIPER_B200_PDL=1
// launches with the attribute (without it the two instructions are no-ops).  All GPU tests pass either way.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace pb200 {

inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = std::getenv("PIPER_B200_PDL");
    on = e ? (std::atoi(e) != 0) : 0;
  }
  return on != 0;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  if (pdl_enabled()) {
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#ifdef __CUDACC__
// Let the next grid in the stream be scheduled early (its prologue overlaps our tail) ...
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
// ... and do not touch global memory that another grid produces or still reads before this returns.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
#endif

}  // namespace pb200
