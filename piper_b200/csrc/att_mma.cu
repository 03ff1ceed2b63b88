// DEFAULT since round 2 (PIPER_B200_ATT3=0 selects the CUDA-core kernel).  CUDA instantiation of the tensor-core relative-position attention;
// the body is att_body.inl (design notes there), the primitives tc_policy_dev.cuh.  The same body runs on the CPU model of
// the primitives in tests/test_att_sim.py; on the B200 it is covered by the whole -m gpu suite.
#include "kernels.cuh"
#include "launch.cuh"

#include <cmath>
#include <cstdlib>
#include <stdexcept>

#define MRF_FN __device__ __forceinline__
#include "tc_policy_dev.cuh"
#include "att_body.inl"

namespace pb200 {
void count_launch();

namespace {
__global__ void __launch_bounds__(att::A_THREADS, 1) att_kernel(const __grid_constant__ att::Args a,
                                                                 const __grid_constant__ CUtensorMap tmq,
                                                                 const __grid_constant__ CUtensorMap tmk) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) att::Barriers<uint64_t> bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u);   // tensor copies: 128-byte aligned
  DevPrim::Ctx cx;
  att::att_body<DevPrim>(a, cx, smem, bar, &tmem_base_s, &tmq, &tmk);
}
}  // namespace

// false: shape outside what the kernel handles (caller keeps the CUDA-core kernel)
bool launch_rel_attention_tc(View qkv, View out, const float* rel_k, const float* rel_v, int H, int n_heads, int window,
                             const int* len, int B, int Tmax, cudaStream_t st, int* tail_thr) {
  if (tail_thr) *tail_thr = 0;
  if (B <= 0 || Tmax <= 0) return true;
  const int dk = H / n_heads;
  if (window != 4 || dk % 16 != 0 || dk > att::A_MAXDK || dk * n_heads != H) return false;
  att::Args a;
  a.qkv = qkv; a.out = out; a.rel_k = rel_k; a.rel_v = rel_v; a.len = len;
  a.H = H; a.dk = dk; a.n_heads = n_heads; a.q_tiles = (Tmax + att::A_QT - 1) / att::A_QT;
  // More tiles than SMs (one CTA per SM): short last tiles go to the CUDA-core kernel, which the caller launches behind
  // this one (PIPER_B200_ATT_TAIL=0 keeps every tile here)
  static int g_tail = -1;
  if (g_tail < 0) {
    const char* e = std::getenv("PIPER_B200_ATT_TAIL");
    g_tail = e ? std::atoi(e) : 16;                      // rows; at most 128
  }
  if (tail_thr && g_tail > 0 && a.q_tiles >= 2 && (long long)a.q_tiles * n_heads * B > 148) a.tail_thr = *tail_thr = g_tail;
  const int smem = att::smem_bytes(dk) + 128;
  static int g_tm = -1;                                   // PIPER_B200_ATT_TM: tensor-map loads of the q / k / v windows (default on)
  if (g_tm < 0) {
    const char* e = std::getenv("PIPER_B200_ATT_TM");
    g_tm = e ? (std::atoi(e) != 0) : 1;
  }
  CUtensorMap tmq{}, tmk{};
  if (g_tm) {
    a.flat = qkv.bs < (long long)qkv.cs ? 1 : 0;
    a.tm = 1;
    TmapDesc dq, dk_;
    att::att_tmaps(a, B, dq, dk_);
    if (!encode_tmap(dq, &tmq) || !encode_tmap(dk_, &tmk)) a.tm = a.flat = 0;   // refused: per-row bulk copies
  }
  // The opt-in limit (227 KB per block) covers static + dynamic shared memory, so asking for 227 KB of dynamic memory is
  // refused: ask for what this head width needs (found on the first GPU run: the launch failed with "invalid argument").
  static int attr_bytes[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_bytes[dev & 63] < smem) {
    const cudaError_t e = cudaFuncSetAttribute(att_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return false;                              // head width too large for one CTA: keep the CUDA-core kernel
    }
    attr_bytes[dev & 63] = smem;
  }
  launch_k(att_kernel, dim3(a.q_tiles * n_heads * B), dim3(att::A_THREADS), smem, st, a, tmq, tmk);
  count_launch();
  return true;
}

}  // namespace pb200
