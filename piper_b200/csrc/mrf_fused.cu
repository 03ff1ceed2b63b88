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

namespace pb200 {
void count_launch();

namespace {

constexpr int F_C = 32;                   // channels of the stage
constexpr int F_M = 256;                  // rows per GEMM (two 128-row MMA tiles)
constexpr int F_G0 = 12;                  // guard rows of the stage-input operand: >= half-width of every first conv
constexpr int F_GA = 36;                  // guard rows of the chain operands:      >= half-width of every later conv
constexpr int F_R0 = F_M + 2 * F_G0;      // 280 rows
constexpr int F_RA = F_M + 2 * F_GA;      // 328 rows
constexpr int F_XS = F_R0 + 8;            // row stride (floats) of the staged fp32 input
constexpr int F_W_SLOTS = 6;
constexpr int F_TAP_BYTES = 4 * 64 * 16;  // one tap: [ci / 8][W_hi rows 0..31 | W_lo rows 0..31][8 x bf16]
constexpr int F_A0_PART = 4 * F_R0 * 16, F_AC_PART = 4 * F_RA * 16;
constexpr int F_OFF_A0 = F_C * F_XS * 4;
constexpr int F_OFF_AC = F_OFF_A0 + 2 * F_A0_PART;
constexpr int F_OFF_W = F_OFF_AC + MRF_MAX_CHAINS * 2 * F_AC_PART;
constexpr int F_OFF_BIAS = F_OFF_W + F_W_SLOTS * F_TAP_BYTES;
constexpr int F_SMEM = F_OFF_BIAS + MRF_MAX_CHAINS * MRF_MAX_STEPS * F_C * 4;
static_assert(F_SMEM <= 227 * 1024, "fused MRF stage does not fit shared memory");
constexpr int F_CONV_WARP0 = 3, F_EPI_WARP0 = 5, F_CONV_THREADS = 64, F_EPI_THREADS = 256;
constexpr int F_THREADS = 13 * 32;   // 416 threads: up to 152 registers each
constexpr uint32_t F_TMEM_CARRIER = 256;

// ---- tcgen05 / TMA / mbarrier primitives (same forms as conv_mma.cu, where they have been exercised on hardware) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  constexpr uint32_t HI = (128u >> 4) | (1u << 14);     // SBO = 128 bytes, descriptor version 1
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(HI)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(mbar))
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
// bounded: this kernel is experimental - a protocol bug must end as a trap (launch failure), not as a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  const uint32_t a = smem_u32(mbar);
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
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}\n"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);   // .x = a (low half), .y = b
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}

// 8 consecutive channels of one row -> one 16-byte operand row of the hi part and one of the lo part
__device__ __forceinline__ void store_split8(uint8_t* hi_row, uint8_t* lo_row, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 8; e += 2) {
    const float ph = __bfloat162float(__float2bfloat16_rn(v[e])), qh = __bfloat162float(__float2bfloat16_rn(v[e + 1]));
    hi[e >> 1] = pack_bf16(ph, qh);
    lo[e >> 1] = pack_bf16(v[e] - ph, v[e + 1] - qh);
  }
  *reinterpret_cast<uint4*>(hi_row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(lo_row) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

struct FBarriers {
  uint64_t raw_full, raw_free, a0_full, a0_free, w_full[F_W_SLOTS], w_empty[F_W_SLOTS], acc_full[2], acc_empty[2],
      a_full[MRF_MAX_CHAINS];
};

__global__ void __launch_bounds__(F_THREADS, 1) mrf_fused_kernel(const __grid_constant__ MrfFusedArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) FBarriers bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  float* xs = reinterpret_cast<float*>(smem);
  uint8_t* A0 = smem + F_OFF_A0;
  uint8_t* AC = smem + F_OFF_AC;
  uint8_t* Wr = smem + F_OFF_W;
  float* bias_s = reinterpret_cast<float*>(smem + F_OFF_BIAS);
  const int n_chains = a.n_chains, n_steps = a.n_steps, pair = a.pair;
  const int tpi = a.tiles_per_item, total = a.total_tiles;

  // ---- prologue: zero the chain operands once (their guard rows are never written again), biases to shared memory
  for (int i = tid; i < MRF_MAX_CHAINS * 2 * F_AC_PART / 16; i += F_THREADS)
    reinterpret_cast<uint4*>(AC)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < n_chains * n_steps * F_C; i += F_THREADS) bias_s[i] = a.bias[i];
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&bar.raw_full, 1); mbar_init(&bar.raw_free, F_EPI_THREADS);
    mbar_init(&bar.a0_full, F_CONV_THREADS); mbar_init(&bar.a0_free, 1);
    for (int i = 0; i < F_W_SLOTS; ++i) { mbar_init(&bar.w_full[i], 1); mbar_init(&bar.w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar.acc_full[i], 1); mbar_init(&bar.acc_empty[i], F_EPI_THREADS); }
    for (int i = 0; i < MRF_MAX_CHAINS; ++i) mbar_init(&bar.a_full[i], F_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");       // the zeroed operands are read by the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  // tile id -> (item, first stored position); every role walks the same list and skips the same tiles
  auto decode = [&](int tile, int& b, int& t0, int& L) {
    const int tb = tile % tpi;
    b = tile / tpi;
    t0 = tb * a.to;
    L = a.len[b] * a.len_scale;
    return t0 < L;
  };

  if (warp == 0) {
    // ---------------------------------------------------------------------- stage input: 32 fp32 rows of the window.
    // Whole warp converged, copies predicated on an elected lane: operands stay in uniform registers (conv_mma.cu, UNI)
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int b, t0, L;
      bool ok = decode(tile, b, t0, L);
      L = __shfl_sync(0xffffffffu, L, 0);
      ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
      if (!ok) continue;
      const int t_lo = t0 - a.hv - F_G0;                               // position of operand row 0
      const int t_base = t_lo & ~3;                                    // shared-memory column 0 <-> position t_base
      const int g0 = max(t_lo, 0) & ~3;
      const int g1 = min((min(t_lo + F_R0, L) + 3) & ~3, a.x.cs);
      const uint32_t row_bytes = (uint32_t)(g1 - g0) * 4;
      if (ti >= 1) mbar_wait(&bar.raw_free, (ti - 1) & 1);
      if (elect_one()) mbar_expect_tx(&bar.raw_full, row_bytes * (uint32_t)F_C);
      const float* src = a.x.p + (long long)b * a.x.bs + g0;
      uint32_t d = smem_u32(xs + (g0 - t_base));
      const uint32_t mb = smem_u32(&bar.raw_full);
      const long long s_step = a.x.cs;
#pragma unroll 4
      for (int c = 0; c < F_C; ++c, d += F_XS * 4, src += s_step) {
        if (elect_one())
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
                       "l"(src), "r"(row_bytes), "r"(mb)
                       : "memory");
      }
      __syncwarp();
      ++ti;
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------- weight taps, in the order the MMAs use them
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int b, t0, L;
      bool ok = decode(tile, b, t0, L);
      ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
      if (!ok) continue;
      const uint8_t* src = a.w;
      for (int s = 0; s < n_steps; ++s)
        for (int c = 0; c < n_chains; ++c)
          for (int j = 0; j < a.k[c]; ++j, ++it, src += F_TAP_BYTES) {
            const int slot = it % F_W_SLOTS;
            if (it >= F_W_SLOTS) mbar_wait(&bar.w_empty[slot], ((it / F_W_SLOTS) - 1) & 1);
            if (elect_one()) {
              mbar_expect_tx(&bar.w_full[slot], F_TAP_BYTES);
              bulk_g2s(Wr + slot * F_TAP_BYTES, src, F_TAP_BYTES, &bar.w_full[slot]);
            }
            __syncwarp();
          }
    }
  } else if (warp == 2) {
    // -------------------------------------------------------------------- MMA issue (whole warp converged; only the
    // tcgen05 instructions are predicated on the elected lane so that every operand stays warp-uniform)
    const uint32_t tmem_du = __shfl_sync(0xffffffffu, tmem_d, 0);
    constexpr uint32_t idesc64 = make_idesc_bf16(128, 64), idesc32 = make_idesc_bf16(128, 32);
    uint32_t g_it = 0, w_it = 0, ti = 0, a_par = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int b, t0, L;
      bool ok = decode(tile, b, t0, L);
      ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;                    // depends on a global load: make it uniform
      if (!ok) continue;
      for (int s = 0; s < n_steps; ++s)
        for (int c = 0; c < n_chains; ++c, ++g_it) {
          const uint32_t slot = g_it & 1;
          if (g_it >= 2) mbar_wait(&bar.acc_empty[slot], ((g_it >> 1) - 1) & 1);
          uint32_t abase, part, lbo, guard;
          if (s == 0) {
            if (c == 0) mbar_wait(&bar.a0_full, ti & 1);
            abase = smem_u32(A0); part = F_A0_PART; lbo = F_R0 * 16; guard = F_G0;
          } else {
            mbar_wait(&bar.a_full[c], (a_par >> c) & 1);
            a_par ^= 1u << c;
            abase = smem_u32(AC + c * 2 * F_AC_PART); part = F_AC_PART; lbo = F_RA * 16; guard = F_GA;
          }
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const int k = a.k[c], dil = a.dil[c][s];
          const uint32_t row0 = guard - (uint32_t)((k - 1) / 2 * dil);
          const uint32_t d0 = tmem_du + slot * 128u;
          const uint32_t a_step = 2u * (lbo >> 4);
          for (int j = 0; j < k; ++j, ++w_it) {
            const int ws = w_it % F_W_SLOTS;
            mbar_wait(&bar.w_full[ws], (w_it / F_W_SLOTS) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t row = row0 + (uint32_t)(j * dil);
            uint32_t ah = desc_lo(abase + row * 16, lbo), al = desc_lo(abase + part + row * 16, lbo);
            uint32_t wb = desc_lo(smem_u32(Wr + ws * F_TAP_BYTES), 64 * 16);
#pragma unroll 1
            for (int kk = 0; kk < F_C / 16; ++kk) {
              const uint32_t accf = (j > 0 || kk > 0) ? 1u : 0u;
              if (elect_one()) {
                mma_bf16(d0, ah, wb, idesc64, accf);                     // rows   0..127: main | correction (hi*lo)
                mma_bf16(d0 + 32u, al, wb, idesc32, 1u);                 //                correction += lo*hi
                mma_bf16(d0 + 64u, ah + 128u, wb, idesc64, accf);        // rows 128..255
                mma_bf16(d0 + 96u, al + 128u, wb, idesc32, 1u);
              }
              __syncwarp();
              ah += a_step; al += a_step; wb += 2u * 64u;
            }
            if (elect_one()) mma_commit(&bar.w_empty[ws]);
            __syncwarp();
          }
          if (s == 0 && c == n_chains - 1) {
            if (elect_one()) mma_commit(&bar.a0_free);                   // the stage-input operand may be overwritten
            __syncwarp();
          }
          if (elect_one()) mma_commit(&bar.acc_full[slot]);
          __syncwarp();
        }
      ++ti;
    }
  } else if (warp < F_EPI_WARP0) {
    // ---------------------------------------------------------------------- converter: x -> lrelu -> hi/lo operand
    const int ctid = tid - F_CONV_WARP0 * 32;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int b, t0, L;
      if (!decode(tile, b, t0, L)) continue;
      const int t_lo = t0 - a.hv - F_G0;
      const int off = t_lo - (t_lo & ~3);
      mbar_wait(&bar.raw_full, ti & 1);
      if (ti >= 1) mbar_wait(&bar.a0_free, (ti - 1) & 1);
      for (int idx = ctid; idx < (F_C / 8) * F_R0; idx += F_CONV_THREADS) {
        const int g = idx / F_R0, rho = idx - g * F_R0;
        const int pos = t_lo + rho;
        const bool live = pos >= 0 && pos < L;                           // outside the utterance: zeros, whatever the
        float v[8];                                                      // (unwritten / stale) shared memory holds
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float x = live ? xs[(g * 8 + e) * F_XS + off + rho] : 0.f;
          v[e] = x > 0.f ? x : x * a.slope;
        }
        store_split8(A0 + (g * F_R0 + rho) * 16, A0 + F_A0_PART + (g * F_R0 + rho) * 16, v);
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      mbar_arrive(&bar.a0_full);
      ++ti;
    }
  } else {
    // ---------------------------------------------------------------------- epilogue: one position per thread
    const int ew = warp - F_EPI_WARP0;                 // 0..7 (any four consecutive warps cover the four lane quadrants)
    const int q = warp & 3, m = ew >> 2;               // TMEM lane quadrant is fixed by warp id % 4; m = 128-row tile
    const int r = m * 128 + q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float n_f = (float)n_chains;
    uint32_t g_it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int b, t0, L;
      if (!decode(tile, b, t0, L)) continue;
      const int p0 = t0 - a.hv;
      const int pos = p0 + r;
      const bool inside = pos >= 0 && pos < L;
      const int t_lo = p0 - F_G0;
      const float* xrow = xs + (t_lo - (t_lo & ~3)) + F_G0 + r;           // this position in the staged input
      float sum[F_C];
#pragma unroll
      for (int i = 0; i < F_C; ++i) sum[i] = 0.f;
      for (int s = 0; s < n_steps; ++s) {
        const bool closes = (s % pair) == pair - 1;                      // this conv ends a residual unit
        const bool last = s == n_steps - 1;
        for (int c = 0; c < n_chains; ++c, ++g_it) {
          const uint32_t slot = g_it & 1;
          mbar_wait(&bar.acc_full[slot], (g_it >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t tb = tmem_d + lane_addr + slot * 128u + (uint32_t)m * 64u;
          float v[F_C];
          {
            float p[16], qv[16];
            tmem_ld16(tb, p);
            tmem_ld16(tb + 32u, qv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = p[i] + qv[i];
            tmem_ld16(tb + 16u, p);
            tmem_ld16(tb + 48u, qv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[16 + i] = p[i] + qv[i];
          }
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          mbar_arrive(&bar.acc_empty[slot]);
          const float* bs = bias_s + (s * n_chains + c) * F_C;
#pragma unroll
          for (int i = 0; i < F_C; ++i) v[i] += bs[i];
          const uint32_t carrier = tmem_d + lane_addr + F_TMEM_CARRIER + (uint32_t)c * 64u + (uint32_t)m * 32u;
          if (closes) {
            if (s == pair - 1) {                                         // residual = the stage input
#pragma unroll
              for (int i = 0; i < F_C; ++i) v[i] += xrow[i * F_XS];
            } else {                                                     // residual = the value parked by the last unit
              float p[16];
              tmem_ld16(carrier, p);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += p[i];
              tmem_ld16(carrier + 16u, p);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[16 + i] += p[i];
            }
          }
          if (s == pair - 1 && c == n_chains - 1) mbar_arrive(&bar.raw_free);   // last read of the staged input
          if (last) {
#pragma unroll
            for (int i = 0; i < F_C; ++i) sum[i] += v[i];
            if (c == n_chains - 1 && r >= a.hv && r < a.hv + a.to && pos < L) {
              float* yb = a.y.p + (long long)b * a.y.bs + pos;
#pragma unroll
              for (int i = 0; i < F_C; ++i) yb[(long long)i * a.y.cs] = sum[i] / n_f;
            }
          } else {
            if (closes) {
              tmem_st16(carrier, v);
              tmem_st16(carrier + 16u, v + 16);
              asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            }
            // operand of this chain's next conv: lrelu, zero outside the utterance (every conv pads its own input)
            uint8_t* hi = AC + c * 2 * F_AC_PART + (F_GA + r) * 16;
#pragma unroll
            for (int g = 0; g < F_C / 8; ++g) {
              float w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float x = v[g * 8 + e];
                w[e] = inside ? (x > 0.f ? x : x * a.slope) : 0.f;
              }
              store_split8(hi + g * F_RA * 16, hi + F_AC_PART + g * F_RA * 16, w);
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            mbar_arrive(&bar.a_full[c]);
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(512u) : "memory");
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
  a.n_chains = p.n_chains; a.n_steps = p.n_steps; a.pair = p.pair; a.hv = p.hv; a.to = p.to;
  for (int c = 0; c < MRF_MAX_CHAINS; ++c) {
    a.k[c] = p.k[c];
    for (int s = 0; s < MRF_MAX_STEPS; ++s) a.dil[c][s] = p.dil[c][s];
  }
  a.tiles_per_item = (max_len + p.to - 1) / p.to;
  const long long total = (long long)a.tiles_per_item * B;
  a.total_tiles = (int)total;
  const int grid = (int)std::min<long long>(total, 148);
  mrf_fused_kernel<<<grid, F_THREADS, F_SMEM, st>>>(a);
  count_launch();
}

}  // namespace pb200
