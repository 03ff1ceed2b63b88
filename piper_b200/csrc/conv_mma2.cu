// EXPERIMENTAL - OFF BY DEFAULT (PIPER_B200_V2=1).  CUDA instantiation of the second-generation persistent tensor-core
// convolution: the body lives in conv2_body.inl (what changes against conv_mma.cu is described there), the primitives
// in tc_policy_dev.cuh, plan / pack / argument fill in conv2_host.h.  The same body runs on a CPU model of the primitives
// in tests/test_conv2_sim.py (every epilogue, both split precisions, ragged batches); it has NOT yet run on a GPU.
#include "kernels.cuh"

#include <cuda_bf16.h>

#include <cstdlib>
#include <stdexcept>

#define MRF_FN __device__ __forceinline__
#define MRF_NOINLINE __device__ __noinline__
#include "tc_policy_dev.cuh"
#include "conv2_host.h"
#include "conv2_body.inl"

namespace pb200 {
void count_launch();

namespace {

template <int PREC, int MT>
__global__ void __launch_bounds__(conv2::C2_THREADS, 1) conv2_kernel(const __grid_constant__ MmaConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) conv2::Barriers<uint64_t> bar;
  __shared__ uint32_t tmem_base_s;
  DevPrim::Ctx cx;
  conv2::conv2_body<DevPrim, PREC, MT>(a, cx, smem, bar, &tmem_base_s);
}

}  // namespace

bool conv2_plan(int ci, int rows, int k, int dil, int prec, int chains, Conv2Layer& l) {
  static int g_opts = -1;                               // PIPER_B200_V2_OPTS: plan options (conv2_host.h)
  if (g_opts < 0) {
    const char* e = std::getenv("PIPER_B200_V2_OPTS");
    g_opts = e ? std::atoi(e) : 0;
  }
  conv2::Plan p;
  if (!conv2::plan(ci, rows, k, dil, prec, chains, p, g_opts)) return false;
  l.tf32 = p.tf32; l.prec = p.prec; l.n_tile = p.n_tile; l.n_tiles = p.n_tiles; l.mt = p.mt; l.kc = p.kc; l.stage_rows = p.stage_rows;
  l.raw_stride = p.raw_stride; l.t_slots = p.t_slots; l.tmem_cols = p.tmem_cols; l.chains = p.chains; l.mh_stride = p.mh_stride;
  l.smem = p.smem; l.w_bytes = p.w_bytes;
  return true;
}

static conv2::Plan to_plan(const Conv2Layer& l) {
  conv2::Plan p;
  p.ok = true; p.tf32 = l.tf32; p.prec = l.prec; p.n_tile = l.n_tile; p.n_tiles = l.n_tiles; p.mt = l.mt; p.kc = l.kc; p.stage_rows = l.stage_rows;
  p.raw_stride = l.raw_stride; p.t_slots = l.t_slots; p.tmem_cols = l.tmem_cols; p.chains = l.chains; p.mh_stride = l.mh_stride;
  p.smem = l.smem; p.w_bytes = l.w_bytes;
  return p;
}

void conv2_pack(const float* wsrc, int ci, int k, int rows_p, const Conv2Layer& l, uint8_t* out) {
  conv2::pack(wsrc, ci, k, rows_p, to_plan(l), out);
}

// a: x / y / y2 / r / w (conv2_pack layout) / bias / len / shape / epilogue fields set by the caller
bool launch_conv2(MmaConvArgs a, const Conv2Layer& l, int B, int max_len, cudaStream_t st) {
  if (B <= 0 || max_len <= 0) return true;
  const conv2::Plan p = to_plan(l);
  const int grid = conv2::fill_args(a, p, B, max_len);
  static int g_small_too = -1;                          // PIPER_B200_V2=2: also take launches with fewer tiles than SMs
  if (g_small_too < 0) {
    const char* e = std::getenv("PIPER_B200_V2");
    g_small_too = (e && std::atoi(e) >= 2) ? 1 : 0;
  }
  if (a.total_tiles < 148 && !g_small_too) return false;   // small launches stay on the one-tile-per-CTA kernel
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(conv2_kernel<0, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv2_kernel<0, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv2_kernel<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv2_kernel<2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv2_kernel<2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set[dev & 63] = true;
  }
  if (p.prec == 1) conv2_kernel<1, 128><<<grid, conv2::C2_THREADS, p.smem, st>>>(a);
  else if (p.prec == 2 && p.mt == 256) conv2_kernel<2, 256><<<grid, conv2::C2_THREADS, p.smem, st>>>(a);
  else if (p.prec == 2) conv2_kernel<2, 128><<<grid, conv2::C2_THREADS, p.smem, st>>>(a);
  else if (p.mt == 256) conv2_kernel<0, 256><<<grid, conv2::C2_THREADS, p.smem, st>>>(a);
  else conv2_kernel<0, 128><<<grid, conv2::C2_THREADS, p.smem, st>>>(a);
  count_launch();
  return true;
}

}  // namespace pb200
