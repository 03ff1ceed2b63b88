// Conv1d on the 5th-generation tensor cores: tcgen05.mma (kind::f16, BF16 operands, FP32 accumulators in
// TMEM) with error-compensated split precision ("bf16x3").
//
//   x = x_hi + x_lo,  w = w_hi + w_lo   (hi = bf16(v), lo = bf16(v - hi): 16 mantissa bits kept)
//   y = x_hi*w_hi + x_hi*w_lo + x_lo*w_hi          (three MMAs, products exact in fp32, fp32 accumulate)
//
// Single-pass BF16/TF32 misses the 1e-3 waveform bar (SURVEY.md finding 7); the 3-term split measured on the
// CPU emulation gives 6.5e-5 on the generator of the real voice.
//
// GEMM view of a stride-1 Conv1d layer (C_in -> C_out, K taps, dilation d), one CTA per 128 output positions:
//   D[128 t][C_out] += A_j[128 t][KC ci] * W_j[C_out][KC ci]^T        for every tap j and channel chunk
// A_j is NOT materialised per tap: the activation tile (with its halo) is staged once per channel chunk in
// the canonical no-swizzle K-major layout  [ci/8][row][8 ci]  where consecutive rows (time steps) are 16 bytes
// apart, so tap j is the same buffer with the descriptor start address advanced by j*d*16 bytes.
// The fp32 -> (leaky-relu) -> bf16 hi/lo conversion happens while staging; weights are pre-split and
// pre-laid-out at load time and stream through a two-deep cp.async ring; the epilogue reads the
// accumulators with tcgen05.ld (lane = output position -> coalesced stores along time) and applies
// bias / residual / MRF-average exactly like the CUDA-core kernel in conv1d.cu.
//
// Reference ops: ResBlock1/ResBlock2 convolutions of the HiFi-GAN generator (modules.py:301-314,355-364).
#include "kernels.cuh"

#include <cuda_bf16.h>

#include <cstring>
#include <vector>

#include <stdexcept>
#include <string>

namespace pb200 {
void count_launch();

namespace {

constexpr int MMA_M = 128;          // output positions per CTA == TMEM lanes
constexpr int MMA_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor, SWIZZLE_NONE, K-major "interleave" canonical layout:
//   element (row, k) at  start + (row % 8) * 16 + (row / 8) * SBO + (k / 8) * LBO + (k % 8) * 2   bytes
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64)).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Instruction descriptor for kind::f16: D=F32 (bits 4-5 = 1), A=B=BF16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28  (cute::UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
      : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(mbar))
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  const uint32_t a = smem_u32(mbar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(a), "r"(parity)
      : "memory");
}

__device__ __forceinline__ void cp_async16_mma(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
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

// dynamic shared memory layout (all offsets multiples of 128 bytes):
//   A_hi [KC/8][R][8] bf16 | A_lo same | W ring: 2 x { hi [KC/8][N][8], lo [KC/8][N][8] }
__global__ void __launch_bounds__(MMA_THREADS) conv_mma_kernel(const MmaConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar_w[2];   // weight-ring slot consumed by the tensor core
  __shared__ __align__(8) uint64_t mbar_a;      // activation chunk consumed
  __shared__ uint32_t tmem_base_s;

  const int b = blockIdx.y;
  const int L = a.len[b] * a.len_scale;
  const int t0 = blockIdx.x * MMA_M;
  if (t0 >= L) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.co, KC = a.kc;
  const int R = a.rows;                          // staged rows: 128 + (k-1)*dil rounded up to 8
  const int a_part_bytes = KC * R * 2;
  const int w_part_bytes = KC * N * 2;
  uint8_t* A_hi = smem;
  uint8_t* A_lo = smem + a_part_bytes;
  uint8_t* W_ring = smem + 2 * a_part_bytes;     // slot s at + s * 2 * w_part_bytes
  const int n_kc = a.ci / KC;
  const int n_units = n_kc * a.k;                // weight units: (kc, tap), kc outer
  const size_t unit_bytes = size_t(2) * w_part_bytes;

  // ---- one-time setup: TMEM allocation (warp 0), barriers (one thread)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(&mbar_w[0], 1);
    mbar_init(&mbar_w[1], 1);
    mbar_init(&mbar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(MMA_M, N);

  auto load_unit = [&](int u) {                  // cp.async one (kc, tap) weight unit into ring slot u & 1
    const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w) + size_t(u) * unit_bytes;
    uint8_t* dst = W_ring + size_t(u & 1) * unit_bytes;
    for (int i = tid * 16; i < (int)unit_bytes; i += MMA_THREADS * 16) cp_async16_mma(dst + i, src + i);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  const float* xb = a.x.p + (long long)b * a.x.bs;
  uint32_t ph_w[2] = {0, 0}, ph_a = 0;
  bool first_mma = true;
  load_unit(0);
  int u = 0;
  for (int kc = 0; kc < n_kc; ++kc) {
    // ---- stage the activation chunk: fp32 -> leaky-relu -> bf16 hi/lo, layout [g][row][8]
    if (kc > 0) {                                // previous chunk's MMAs must have finished reading A
      mbar_wait(&mbar_a, ph_a);
      ph_a ^= 1;
    }
    const int c0 = kc * KC;
    const int items = (KC / 8) * R;
    for (int idx = tid; idx < items; idx += MMA_THREADS) {
      const int g = idx / R, r = idx - g * R;
      const int t = t0 - a.pad + r;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (t >= 0 && t < L) {
        const float* xr = xb + (long long)(c0 + g * 8) * a.x.cs + t;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(xr + (long long)e * a.x.cs);
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        float p = v[e], q = v[e + 1];
        if (a.pre == PRE_LRELU) {
          p = p > 0.f ? p : p * a.slope;
          q = q > 0.f ? q : q * a.slope;
        }
        const float ph = __bfloat162float(__float2bfloat16_rn(p)), qh = __bfloat162float(__float2bfloat16_rn(q));
        hi[e >> 1] = pack_bf16(ph, qh);
        lo[e >> 1] = pack_bf16(p - ph, q - qh);
      }
      const int off = (g * R + r) * 16;
      *reinterpret_cast<uint4*>(A_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(A_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }

    for (int j = 0; j < a.k; ++j, ++u) {
      // prefetch the next weight unit into the other slot once the tensor core has released it
      if (u + 1 < n_units) {
        if (u + 1 >= 2) {
          mbar_wait(&mbar_w[(u + 1) & 1], ph_w[(u + 1) & 1]);
          ph_w[(u + 1) & 1] ^= 1;
        }
        load_unit(u + 1);
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      }
      // generic-proxy writes (st.shared / cp.async) -> visible to the tensor core's async proxy
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo);
        const uint32_t w_hi = smem_u32(W_ring + size_t(u & 1) * unit_bytes), w_lo = w_hi + w_part_bytes;
        const uint32_t a_lbo = R * 16, w_lbo = N * 16;
        const uint32_t shift = (uint32_t)(j * a.dil) * 16;
        for (int kb = 0; kb < KC / 16; ++kb) {
          const uint64_t ah = make_desc(a_hi + 2 * kb * a_lbo + shift, a_lbo, 128);
          const uint64_t al = make_desc(a_lo + 2 * kb * a_lbo + shift, a_lbo, 128);
          const uint64_t wh = make_desc(w_hi + 2 * kb * w_lbo, w_lbo, 128);
          const uint64_t wl = make_desc(w_lo + 2 * kb * w_lbo, w_lbo, 128);
          mma_bf16(tmem_d, ah, wh, idesc, !first_mma);
          first_mma = false;
          mma_bf16(tmem_d, ah, wl, idesc, true);
          mma_bf16(tmem_d, al, wh, idesc, true);
        }
        mma_commit(&mbar_w[u & 1]);                    // ring slot free when these MMAs retire
        if (j == a.k - 1) mma_commit(&mbar_a);         // ... and so is the activation chunk
      }
    }
  }
  // ---- all MMAs of this tile issued: the last commit on mbar_a covers them (commits are cumulative)
  mbar_wait(&mbar_a, ph_a);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4)..+31 (its sub-partition), column half w/4
  {
    const int q = warp & 3, half = warp >> 2;
    const int t = t0 + q * 32 + lane;
    const int ncol = N / 2;
    float* yb = a.y.p ? a.y.p + (long long)b * a.y.bs : nullptr;
    float* y2b = a.y2.p ? a.y2.p + (long long)b * a.y2.bs : nullptr;
    const float* rb = a.r.p ? a.r.p + (long long)b * a.r.bs : nullptr;
    for (int cb = 0; cb < ncol; cb += 16) {
      const int col0 = half * ncol + cb;
      float v[16];
      tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
      if (t < L) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int co = col0 + i;
          float val = v[i] + (a.bias ? __ldg(a.bias + co) : 0.f);
          switch (a.epi) {
            case EPI_BIAS: yb[(long long)co * a.y.cs + t] = val; break;
            case EPI_RES: yb[(long long)co * a.y.cs + t] = val + rb[(long long)co * a.r.cs + t]; break;
            case EPI_MRF: {
              const float v2 = val + rb[(long long)co * a.r.cs + t];
              float* o = y2b + (long long)co * a.y2.cs + t;
              if (a.mrf == 0) *o = v2;
              else if (a.mrf == 1) *o = *o + v2;
              else *o = (*o + v2) / (float)a.mrf_n;
              break;
            }
            default: break;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
}

}  // namespace

bool mma_conv_supported(int ci, int co, int k, int dil) {
  if (ci % 16 != 0 || co % 32 != 0 || co < 32 || co > 256) return false;
  const int rows = (MMA_M + (k - 1) * dil + 7) & ~7;
  const int kc = mma_conv_chunk(ci, co, k, dil);
  return kc > 0 && rows * 16 < (1 << 18);
}

// channel chunk KC (multiple of 16, divides ci) such that A (hi+lo) + 2 weight slots fit ~100 KB (2 CTAs / SM)
int mma_conv_chunk(int ci, int co, int k, int dil) {
  const int rows = (MMA_M + (k - 1) * dil + 7) & ~7;
  for (int kc = ci; kc >= 16; kc -= 16) {
    if (ci % kc) continue;
    const size_t bytes = size_t(2) * kc * rows * 2 + size_t(2) * 2 * kc * co * 2;
    if (bytes <= 110 * 1024) return kc;
  }
  return 0;
}

void launch_conv_mma(MmaConvArgs a, int B, int max_len, cudaStream_t st) {
  if (B <= 0 || max_len <= 0) return;
  a.rows = (MMA_M + (a.k - 1) * a.dil + 7) & ~7;
  a.kc = mma_conv_chunk(a.ci, a.co, a.k, a.dil);
  if (a.kc <= 0) throw std::runtime_error("conv_mma: layer shape not supported by the tensor-core path");
  a.tmem_cols = a.co <= 32 ? 32 : a.co <= 64 ? 64 : a.co <= 128 ? 128 : 256;
  const size_t smem = size_t(2) * a.kc * a.rows * 2 + size_t(2) * 2 * a.kc * a.co * 2;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(conv_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set[dev & 63] = true;
  }
  dim3 grid((max_len + MMA_M - 1) / MMA_M, B);
  conv_mma_kernel<<<grid, MMA_THREADS, smem, st>>>(a);
  count_launch();
}

// Host-side packing of one Conv1d weight [Co][Ci][K] (fp32) into the tensor-core layout:
//   units (kc, tap) in issue order, each unit = { hi [KC/8][Co][8], lo [KC/8][Co][8] } bf16
static inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x40);   // NaN
  const uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return uint16_t((u + r) >> 16);
}
static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = uint32_t(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

void pack_conv_mma(const float* w, int co, int ci, int k, int kc, std::vector<uint16_t>& out) {
  const int n_kc = ci / kc;
  const size_t part = size_t(kc) * co;
  out.assign(size_t(n_kc) * k * 2 * part, 0);
  for (int c = 0; c < n_kc; ++c)
    for (int j = 0; j < k; ++j) {
      uint16_t* unit = out.data() + (size_t(c) * k + j) * 2 * part;
      for (int g = 0; g < kc / 8; ++g)
        for (int n = 0; n < co; ++n)
          for (int e = 0; e < 8; ++e) {
            const float v = w[(size_t(n) * ci + (c * kc + g * 8 + e)) * k + j];
            const uint16_t hi = f32_to_bf16_rn(v);
            const uint16_t lo = f32_to_bf16_rn(v - bf16_to_f32(hi));
            unit[(size_t(g) * co + n) * 8 + e] = hi;
            unit[part + (size_t(g) * co + n) * 8 + e] = lo;
          }
    }
}

}  // namespace pb200
