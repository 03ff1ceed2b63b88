// DEFAULT since round 2 (engine mask bit 16; PIPER_B200_MMA=15 runs the stage layer-wise).  Written at the end of round 1 from the
// per-launch measurements in profiles/r01_layer_report.txt; first run on a B200 in round 2 (profiles/r02_summary.md).
// (The layer-wise path in conv_mma2.cu remains for the other stages.)  tools/fused_mrf_proto.py is the CPU statement of the same
// tile algorithm and tests/test_fused_mrf_proto.py pins it against the oracle.
//
// One multi-receptive-field stage of the HiFi-GAN generator in a single persistent kernel, for the C = 32 stage (the last
// stage of every piper preset; 2.6 of the 4.6 ms the resblocks take in the medium voice):
//
//     y = ( ResBlock_0(x) + ResBlock_1(x) + ResBlock_2(x) ) / 3                      models.py:356-363
//     ResBlock2:  v = v + conv_{k,d}(lrelu(v))            for d in dilations         modules.py:355-364
//     ResBlock1:  v = v + conv_{k,1}(lrelu(conv_{k,d}(lrelu(v))))                    modules.py:301-314
//
// Layer-wise this is 14 passes over a [B][32][L] fp32 tensor (6 convs x (read + write) + 2 MRF read-modify-writes);
// here x is read once and y written once.  A tile is 256 consecutive positions ("rows") starting hv positions before
// the first output it will store; every convolution of every chain is computed on all 256 rows with the SAME row <->
// position mapping, so a thread of the epilogue owns one position for the whole tile.  A conv of half-width w makes the
// outermost w rows of its output meaningless; after the whole chain the middle TO = 256 - 2 hv rows are exact and are the
// ones stored (hv = the summed half-widths of all convs but the first, whose halo comes from real neighbouring data).
// The reference zero-pads the input of EVERY conv at the utterance edges: operands are written as 0 for positions
// outside [0, L) whatever the recompute produced there.
//
// Split precision: FP16 hi/lo operands (fp16x3, 22 mantissa bits at K = 16 per instruction - DESIGN.md section 3), with the two weight halves stacked along N so a k-step is two
// instructions instead of three (the SS-form instruction is paced by reading A from shared memory):
//     D[128][ 0..31] += A_hi * W_hi^T,  D[128][32..63] += A_hi * W_lo^T      one tcgen05.mma, N = 64, B = [W_hi ; W_lo]
//     D[128][32..63] += A_lo * W_hi^T                                         one tcgen05.mma, N = 32
// The epilogue adds the main and the correction half in fp32 (round to nearest).
//
// Warp roles (13 warps): 0 stage-input TMA, 1 weight-tap TMA ring, 2 MMA issue, 3-4 converter (fp32 -> lrelu -> hi/lo
// operand), 5-12 epilogue.  GEMMs are issued step-major over the chains, G(c0,s0) G(c1,s0) G(c2,s0) G(c0,s1) ..., so
// the epilogue of one chain's conv overlaps the MMAs of the other two; two TMEM accumulator sets form the ring between
// the two roles.  The fp32 value a later residual add needs ("carrier") is parked in spare TMEM columns.
//
// TMEM columns: [0,256) two accumulator sets x two 128-row tiles x (32 main + 32 correction);
//               [256,448) carriers, chain c / row tile m at 256 + 64 c + 32 m.
#include "kernels.cuh"
#include "launch.cuh"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>

#define MRF_FN __device__ __forceinline__
#include "tc_policy_dev.cuh"
#include "mrf_fused_body.inl"
#include "conv2_host.h"   // f32_to_f16_rn / f16_to_f32

namespace pb200 {
void count_launch();

namespace {
using namespace mrf;

__global__ void __launch_bounds__(F_THREADS, 1) mrf_fused_kernel(const __grid_constant__ MrfFusedArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) FBarriers<uint64_t> bar;
  __shared__ uint32_t tmem_base_s;
  DevPrim::Ctx cx;
  mrf_fused_body<DevPrim>(a, cx, smem, bar, &tmem_base_s);
}

// conv of chain c, step s (ResBlock1 alternates convs1 / convs2; ResBlock2 walks convs)
const ConvW& step_conv(const ResBlockW& rb, int pair, int s) { return pair == 1 ? rb.c1.at(s) : ((s & 1) ? rb.c2.at(s / 2) : rb.c1.at(s / 2)); }

}  // namespace

bool plan_mrf_fused(const std::vector<ResBlockW>& stage, int resblock_kind, int channels, int post_k, MrfFusedPlan& p) {
  p = MrfFusedPlan{};
  if (channels != F_C || stage.empty() || int(stage.size()) > MRF_MAX_CHAINS) return false;
  p.n_chains = int(stage.size());
  p.pair = resblock_kind == 1 ? 2 : 1;
  p.n_steps = p.pair * int(stage[0].c1.size());
  if (p.n_steps < 2 || p.n_steps > MRF_MAX_STEPS) return false;
  p.hv = 0;
  for (int c = 0; c < p.n_chains; ++c) {
    const ResBlockW& rb = stage[c];
    if (int(rb.c1.size()) * p.pair != p.n_steps || (p.pair == 2 && rb.c2.size() != rb.c1.size())) return false;
    p.k[c] = rb.k;
    int later = 0;
    for (int s = 0; s < p.n_steps; ++s) {
      const ConvW& cw = step_conv(rb, p.pair, s);
      if (cw.ci != F_C || cw.rows != F_C || cw.k != rb.k || (cw.k & 1) == 0 || cw.up != 1) return false;
      const int hw = (cw.k - 1) / 2 * cw.dil;
      if (cw.pad != hw) return false;                                   // 'same' padding only
      if (s == 0 ? hw > F_G0 : hw > F_GA) return false;
      p.dil[c][s] = cw.dil;
      if (s > 0) later += hw;
    }
    p.hv = std::max(p.hv, later);
  }
  if (post_k > 0) {                                                    // fused conv_post: one more halo of (k - 1) / 2
    if ((post_k & 1) == 0 || post_k > F_POST_MAXK) return false;
    p.post_k = post_k;
    p.hv += (post_k - 1) / 2;
  }
  p.to = F_M - 2 * p.hv;
  if (p.to < 64) return false;
  int taps = 0;
  for (int c = 0; c < p.n_chains; ++c) taps += p.k[c] * p.n_steps;
  p.w_bytes = size_t(taps) * F_TAP_BYTES;
  p.n_bias = p.n_steps * p.n_chains * F_C;
  p.ok = true;
  return true;
}

// Tap tiles in consumption order (step-major over the chains), each [ci / 8][64 rows = W_hi co 0..31 | W_lo co 0..31][8 ci]
// fp16; biases [step][chain][32].  `blob` is the engine's fp32 weight blob: a conv is [ci][k][rows_p], row fastest.
void pack_mrf_fused(const float* blob, const std::vector<ResBlockW>& stage, const MrfFusedPlan& p, uint8_t* w, float* bias) {
  uint16_t* out = reinterpret_cast<uint16_t*>(w);
  for (int s = 0; s < p.n_steps; ++s)
    for (int c = 0; c < p.n_chains; ++c) {
      const ConvW& cw = step_conv(stage[c], p.pair, s);
      for (int i = 0; i < F_C; ++i) bias[(s * p.n_chains + c) * F_C + i] = cw.b >= 0 ? blob[cw.b + i] : 0.f;
      for (int j = 0; j < cw.k; ++j, out += F_TAP_BYTES / 2)
        for (int ci = 0; ci < F_C; ++ci)
          for (int co = 0; co < F_C; ++co) {
            const float v = blob[cw.w + (int64_t(ci) * cw.k + j) * cw.rows_p + co];
            const uint16_t hi = conv2::f32_to_f16_rn(v), lo = conv2::f32_to_f16_rn(v - conv2::f16_to_f32(hi));
            const int g = ci / 8, e = ci % 8;
            out[(g * 64 + co) * 8 + e] = hi;
            out[(g * 64 + 32 + co) * 8 + e] = lo;
          }
    }
}

void launch_mrf_fused(MrfFusedArgs a, const MrfFusedPlan& p, int B, int max_len, cudaStream_t st) {
  if (B <= 0 || max_len <= 0) return;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(mrf_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM);
    attr_set[dev & 63] = true;
  }
  mrf_fill_args(a, p, B, max_len);
  const long long total = a.total_tiles;
  const int grid = (int)std::min<long long>(total, 148);
  launch_k(mrf_fused_kernel, dim3(grid), dim3(F_THREADS), F_SMEM, st, a);
  count_launch();
}

}  // namespace pb200
