// EXPERIMENTAL - OFF BY DEFAULT (engine mask bit 16 / PIPER_B200_MMA=31).  Written at the end of round 1 from the
// per-launch measurements in profiles/r01_layer_report.txt; compiled and checked with ptxas here, NOT yet run on a GPU.
// The shipped path is the layer-wise one in conv_mma.cu.  tools/fused_mrf_proto.py is the CPU statement of the same
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
// Split precision as in conv_mma.cu (bf16x3), but with the two weight halves stacked along N so a k-step is two
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

#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>

#define MRF_FN __device__ __forceinline__
#include "mrf_fused_body.inl"

namespace pb200 {
void count_launch();

namespace {
using namespace mrf;

// ---- tcgen05 / TMA / mbarrier primitives (same forms as conv_mma.cu, where they have been exercised on hardware) ----
struct DevPrim {
  using Mbar = uint64_t;
  struct Ctx {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int block() const { return blockIdx.x; }
    __device__ __forceinline__ int grid() const { return gridDim.x; }
  };
  static __device__ __forceinline__ uint32_t saddr(Ctx&, const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
  static __device__ __forceinline__ int bcast0(Ctx&, int v) { return __shfl_sync(0xffffffffu, v, 0); }
  static __device__ __forceinline__ void syncwarp() { __syncwarp(); }
  static __device__ __forceinline__ void syncthreads(Ctx&) { __syncthreads(); }
  static __device__ __forceinline__ bool elect_one(Ctx&) {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t}\n"
        : "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred != 0;
  }
  static __device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  static __device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
  static __device__ __forceinline__ void fence_tc_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
  static __device__ __forceinline__ void fence_tc_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
  static __device__ __forceinline__ void mbar_init(Ctx& c, Mbar* m, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(saddr(c, m)), "r"(count) : "memory");
  }
  static __device__ __forceinline__ void mbar_arrive(Ctx& c, Mbar* m) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(saddr(c, m)) : "memory");
  }
  static __device__ __forceinline__ void mbar_expect_tx(Ctx& c, Mbar* m, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(saddr(c, m)), "r"(bytes) : "memory");
  }
  // bounded: this kernel is experimental - a protocol bug must end as a trap (launch failure), not as a hung GPU
  static __device__ __forceinline__ void mbar_wait(Ctx& c, Mbar* m, uint32_t parity) {
    const uint32_t a = saddr(c, m);
    const long long t0 = clock64();
    for (;;) {
      uint32_t done;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(a), "r"(parity)
          : "memory");
      if (done) return;
      if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
  // TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
  static __device__ __forceinline__ void bulk_g2s(Ctx& c, uint32_t dst_saddr, const void* gsrc, uint32_t bytes, Mbar* m) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_saddr),
                 "l"(gsrc), "r"(bytes), "r"(saddr(c, m))
                 : "memory");
  }
  static __device__ __forceinline__ void tmem_alloc(Ctx& c, uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(saddr(c, slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  static __device__ __forceinline__ void tmem_dealloc(Ctx&, uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(cols) : "memory");
  }
  static __device__ __forceinline__ void mma_bf16(Ctx&, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    constexpr uint32_t HI = (128u >> 4) | (1u << 14);     // SBO = 128 bytes, descriptor version 1
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(HI)
        : "memory");
  }
  static __device__ __forceinline__ void mma_commit(Ctx& c, Mbar* m) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(saddr(c, m)) : "memory");
  }
  static __device__ __forceinline__ void tmem_ld16(Ctx&, uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
  }
  static __device__ __forceinline__ void tmem_st16(Ctx&, uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
  }
  static __device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
  static __device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
  static __device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);   // .x = a (low half), .y = b
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

__global__ void __launch_bounds__(F_THREADS, 1) mrf_fused_kernel(const __grid_constant__ MrfFusedArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) FBarriers<uint64_t> bar;
  __shared__ uint32_t tmem_base_s;
  DevPrim::Ctx cx;
  mrf_fused_body<DevPrim>(a, cx, smem, bar, &tmem_base_s);
}

inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x40);
  const uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return uint16_t((u + r) >> 16);
}
inline float bf16_to_f32(uint16_t h) {
  uint32_t u = uint32_t(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// conv of chain c, step s (ResBlock1 alternates convs1 / convs2; ResBlock2 walks convs)
const ConvW& step_conv(const ResBlockW& rb, int pair, int s) { return pair == 1 ? rb.c1.at(s) : ((s & 1) ? rb.c2.at(s / 2) : rb.c1.at(s / 2)); }

}  // namespace

bool plan_mrf_fused(const std::vector<ResBlockW>& stage, int resblock_kind, int channels, MrfFusedPlan& p) {
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
// bf16; biases [step][chain][32].  `blob` is the engine's fp32 weight blob: a conv is [ci][k][rows_p], row fastest.
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
            const uint16_t hi = f32_to_bf16_rn(v), lo = f32_to_bf16_rn(v - bf16_to_f32(hi));
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
  mrf_fused_kernel<<<grid, F_THREADS, F_SMEM, st>>>(a);
  count_launch();
}

}  // namespace pb200
