// Conv1d / ConvTranspose1d on the 5th-generation tensor cores: tcgen05.mma with FP32 accumulators in TMEM and
// error-compensated split precision.
//
//   x = x_hi + x_lo,  w = w_hi + w_lo,   y = x_hi*w_hi + x_hi*w_lo + x_lo*w_hi   (three MMAs per k-step)
//
//   * kind::f16  / BF16 operands ("bf16x3", 16 mantissa bits kept): HiFi-GAN generator layers
//   * kind::tf32 / TF32 operands ("tf32x3", 21 mantissa bits kept): flow (WaveNet) and text-encoder layers,
//     whose outputs feed exp()/ceil() and need fp32-grade accuracy (DESIGN.md §Precision has the measurements;
//     single-pass BF16/TF32 misses the 1e-3 waveform bar, SURVEY.md finding 7).
//
// GEMM view of one layer (C_in -> rows, K taps, dilation d); one CTA owns MT (128 or 256) output positions x
// one tile of up to 256 output rows:
//     D[MT t][N] += A_j[MT t][KC ci] * W_j[N][KC ci]^T      for every tap j and every channel chunk
// A_j is never materialised per tap: the activation chunk (with its halo) is staged once in the canonical
// no-swizzle K-major layout [ci/E][row][E] (E = 16 bytes of elements) where consecutive rows (time steps) are
// 16 bytes apart, so tap j is the same buffer with the descriptor start address advanced by j*d*16 bytes.
//
// Warp roles (320 threads):
//   warps 0-7  producers: global fp32 -> (leaky-relu) -> hi/lo split -> shared, double-buffered per channel chunk;
//              then the epilogue: tcgen05.ld (lane = output position -> coalesced stores along time), bias and
//              the same fused epilogues as conv1d.cu (residual, MRF average, WaveNet gate, WN residual/skip,
//              coupling subtract, relu, ConvTranspose pixel-shuffle)
//   warp  8    owns the TMEM allocation; its lane 0 issues every tcgen05.mma and the tcgen05.commit that
//              releases ring slots (mbarrier arrive when the MMAs that read the slot have retired)
//   warp  9    lane 0 feeds the weight ring with cp.async.bulk (TMA 1-D) + mbarrier complete_tx
//
// Reference ops: modules.py:184-209 (WN), 301-314 / 355-364 (ResBlock1/2), models.py:348-368 (Generator),
// attentions.py:215-223,386-407 (1x1 projections, FFN), all lowered by voice.cc.
#include "kernels.cuh"
#include "launch.cuh"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace pb200 {
void count_launch();

namespace {

constexpr int PROD_THREADS = 256;                 // 8 producer / epilogue warps
constexpr int MMA_THREADS = PROD_THREADS + 64;    // + warp 8 (MMA issuer, TMEM owner) + warp 9 (weight TMA)
constexpr int MAX_A_SLOTS = 2;
constexpr int MAX_W_SLOTS = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor, SWIZZLE_NONE, K-major canonical ("interleave") layout:
//   element (row, k) at  start + (row % 8) * 16 + (row / 8) * SBO + (k / E) * LBO + (k % E) * elem_bytes
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64)=0)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// The MMA issuer is a single thread: descriptor arithmetic per instruction must stay at a couple of integer ops or the
// issue loop, not the tensor pipe, paces small-N layers.  hi word (SBO, version) is constant; the lo word is
// (LBO >> 4) << 16 | (start >> 4), and advancing the start address is one 32-bit add.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__device__ __forceinline__ uint64_t desc_join(uint32_t lo) {
  constexpr uint32_t hi = (128u >> 4) | (1u << 14);     // SBO = 128 bytes, descriptor version 1
  return ((uint64_t)hi << 32) | lo;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (bits 4-5 = 1), A/B format at bits 7-9 / 10-12
// (1 = BF16, 2 = TF32), both operands K-major, N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc(bool tf32, int M, int N) {
  return (1u << 4) | ((tf32 ? 2u : 1u) << 7) | ((tf32 ? 2u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

template <bool TF32>
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
  if (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
  }
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
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  const uint32_t a = smem_u32(mbar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
// Watchdog form for EXPERIMENTAL kernel variants: a protocol bug becomes a trap (launch failure) after ~2 s instead of
// a hung GPU.  BOUNDED = false is exactly mbar_wait.
template <bool BOUNDED>
__device__ __forceinline__ void mbar_wait_t(uint64_t* mbar, uint32_t parity) {
  if (!BOUNDED) {
    mbar_wait(mbar, parity);
    return;
  }
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
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);   // .x = a (low half), .y = b
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
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

__device__ __forceinline__ float sigmoid_acc(float v) { return 1.f / (1.f + expf(-v)); }
// WaveNet gate tanh(a) * sigmoid(b) (commons.py:99-106) as ONE out-of-line copy: tanhf / expf expand to ~100 instructions
// and the epilogues call it 8 times per 16-row chunk; inlined, that alone was several KB of the kernel's hot path.
__device__ __noinline__ float wn_gate(float a, float b) { return tanhf(a) * sigmoid_acc(b); }

template <bool TF32, bool ACC>
__device__ __forceinline__ void mma_ss_imm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  // accumulate flag as a compile-time predicate: no setp in the issue loop
  if (TF32) {
    if (ACC) asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
  } else {
    if (ACC) asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
  }
}

// One lane of a converged warp (elect.sync).  The MMA / TMA issue loops are executed by the WHOLE warp with only the
// instruction itself predicated on the elected lane: inside an `if (lane == 0)` region the compiler has to build the
// 64-bit descriptors in vector registers and move them to the uniform register file (R2UR) for every UTCHMMA, which
// measured at ~200 cycles per instruction; in warp-uniform code they live in uniform registers (tools/mma_bench.py).
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

// ---- lean issue path --------------------------------------------------------------------------------------------
// Measured (tools/mma_bench.py): a tcgen05.mma retires in ~16-100 cycles (N = 32..256) but every instruction the
// single issuing thread executes between two MMAs costs ~5-6 cycles of dependent latency, so ~30 instructions of
// descriptor / index arithmetic per MMA throttled the first versions of this kernel to one MMA per ~200 cycles.
// The loop below spends ~3 instructions per MMA: descriptors are (constant hi word, running lo word) pairs
// assembled inside the asm statement, the accumulate flag is a predicate set once per k-step.
template <bool TF32>
__device__ __forceinline__ void mma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  constexpr uint32_t HI = (128u >> 4) | (1u << 14);     // SBO = 128 bytes, descriptor version 1
  if (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(HI)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(HI)
        : "memory");
  }
}

// All k-steps of one tap for up to two row halves.  acc_main / acc_corr: 0 => the first k-step overwrites that
// accumulator.  Descriptor lo words advance by a_step / w_step (16-byte units) per k-step.
template <bool TF32, int MH>
__device__ __forceinline__ void issue_tap(uint32_t ah, uint32_t al, uint32_t wh, uint32_t wl, int ksteps, uint32_t a_step,
                                          uint32_t w_step, bool two_halves, uint32_t dm, uint32_t dc, uint32_t mh_cols,
                                          uint32_t idesc, uint32_t acc_main, uint32_t acc_corr) {
  // Executed by the WHOLE (converged) MMA warp: all operands are warp-uniform, so the compiler keeps them in
  // uniform registers, which is what UTCHMMA reads; only the instruction itself is predicated on the elected lane.
#pragma unroll 1
  for (int kb = 0; kb < ksteps; ++kb) {
    if (elect_one()) {
      mma_lo<TF32>(dm, ah, wh, idesc, acc_main);
      mma_lo<TF32>(dc, ah, wl, idesc, acc_corr);
      mma_lo<TF32>(dc, al, wh, idesc, 1u);
    }
    if (MH == 2 && two_halves) {
      if (elect_one()) {
        mma_lo<TF32>(dm + mh_cols, ah + 128u, wh, idesc, acc_main);
        mma_lo<TF32>(dc + mh_cols, ah + 128u, wl, idesc, acc_corr);
        mma_lo<TF32>(dc + mh_cols, al + 128u, wh, idesc, 1u);
      }
    }
    __syncwarp();
    acc_main = 1u;
    acc_corr = 1u;
    ah += a_step; al += a_step; wh += w_step; wl += w_step;
  }
}

struct Barriers {
  uint64_t a_full[MAX_A_SLOTS], a_empty[MAX_A_SLOTS], w_full[MAX_W_SLOTS], w_empty[MAX_W_SLOTS], d_full;
};

// dynamic shared memory: A ring  A_SLOTS x { hi [KC/E][R][16 B], lo same }  |  W ring  W_SLOTS x { hi [KC/E][N][16 B], lo }
template <bool TF32, int MT>
__global__ void __launch_bounds__(MMA_THREADS) conv_mma_kernel(const MmaConvArgs a) {
  pdl_launch_dependents();   // programmatic dependent launch (launch.cuh)
  constexpr int ES = TF32 ? 4 : 2;         // operand element bytes
  constexpr int E = 16 / ES;               // elements per 16-byte K chunk
  constexpr int KSTEP = 2 * E;             // K per MMA (32 bytes): 16 (bf16) / 8 (tf32)
  constexpr int MH = MT / 128;             // accumulators (row halves) per CTA
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) Barriers bar;
  __shared__ uint32_t tmem_base_s;

  const int b = blockIdx.y;
  const int L = a.len[b] * a.len_scale;
  const int Lq = L + a.q_extra;
  const int t0 = blockIdx.x * MT;
  if (t0 >= Lq) return;
  const int n0 = blockIdx.z * a.n_tile;                  // first output row of this CTA
  const int N = min(a.n_tile, a.rows - n0);              // rows in this tile (multiple of 16)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int KC = a.kc, R = a.stage_rows;
  const int A_SLOTS = a.a_slots, W_SLOTS = a.w_slots;
  const int a_part = KC * R * ES, w_part = KC * a.n_tile * ES;      // ring slots are sized for a full N tile
  uint8_t* A_ring = smem;
  uint8_t* W_ring = smem + size_t(A_SLOTS) * 2 * a_part;
  const int n_kc = a.ci / KC;
  const int n_units = n_kc * a.k;
  const uint32_t unit_bytes = 2u * (uint32_t)w_part;
  const int mh_live = (Lq - t0 > 128 && MH > 1) ? 2 : 1;            // skip the second row half at the ragged end

  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < A_SLOTS; ++i) { mbar_init(&bar.a_full[i], PROD_THREADS); mbar_init(&bar.a_empty[i], 1); }
    for (int i = 0; i < W_SLOTS; ++i) { mbar_init(&bar.w_full[i], 1); mbar_init(&bar.w_empty[i], 1); }
    mbar_init(&bar.d_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  pdl_wait();                // the prologue above overlapped the previous grid; nothing below runs before it has completed

  if (warp >= 8) {
    // =========================== control warps: weight TMA (warp 9) and MMA issue (warp 8), one lane each ===
    if (warp == 9 && lane == 0) {
      // global layout [n tile][tap][hi|lo][ci/E][n_tile][E]: a (chunk, tap) unit is two contiguous runs
      const size_t g_row = size_t(a.n_tile) * 16, part_all = size_t(a.ci / E) * g_row;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w) + size_t(blockIdx.z) * a.k * 2 * part_all;
      for (int u = 0; u < n_units; ++u) {
        const int s = u % W_SLOTS;
        if (u >= W_SLOTS) mbar_wait(&bar.w_empty[s], ((u / W_SLOTS) - 1) & 1);
        const int kc = u / a.k, j = u - kc * a.k;
        const uint8_t* hi = wsrc + size_t(j) * 2 * part_all + size_t(kc) * (KC / E) * g_row;
        mbar_expect_tx(&bar.w_full[s], unit_bytes);
        bulk_g2s(W_ring + size_t(s) * unit_bytes, hi, (uint32_t)w_part, &bar.w_full[s]);
        bulk_g2s(W_ring + size_t(s) * unit_bytes + w_part, hi + part_all, (uint32_t)w_part, &bar.w_full[s]);
      }
    } else if (warp == 8 && lane == 0) {
      const uint32_t idesc = make_idesc(TF32, 128, N);
      const uint32_t a_lbo = (uint32_t)R * 16, w_lbo = (uint32_t)a.n_tile * 16;
      // The tensor core truncates (round-toward-zero) when it adds into the fp32 accumulator, a systematic bias
      // that grows with the number of accumulation steps.  Where fp32-grade accuracy is needed (tf32x3 layers)
      // the K range is cut into `chains` independent accumulators and the two small correction terms
      // (hi*lo, lo*hi) go to an accumulator of their own; the epilogue adds them in fp32 with round-to-nearest.
      uint32_t started = 0;                               // bit (mh * 8 + accumulator) set once it holds data
      int u = 0;
      for (int kc = 0; kc < n_kc; ++kc) {
        const int as = kc % A_SLOTS;
        mbar_wait(&bar.a_full[as], (kc / A_SLOTS) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t a_hi = smem_u32(A_ring + size_t(as) * 2 * a_part), a_lo = a_hi + a_part;
        for (int j = 0; j < a.k; ++j, ++u) {
          const int ws = u % W_SLOTS;
          mbar_wait(&bar.w_full[ws], (u / W_SLOTS) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t w_hi = smem_u32(W_ring + size_t(ws) * unit_bytes), w_lo = w_hi + w_part;
          const uint32_t shift = (uint32_t)(j * a.dil) * 16;
          const int chain = (u * a.chains) / n_units;
          const int corr = a.sep_corr ? a.chains : chain;
          for (int kb = 0; kb < KC / KSTEP; ++kb) {
            const uint64_t wh = make_desc(w_hi + 2 * kb * w_lbo, w_lbo, 128);
            const uint64_t wl = make_desc(w_lo + 2 * kb * w_lbo, w_lbo, 128);
            for (int mh = 0; mh < mh_live; ++mh) {
              const uint32_t row_off = shift + (uint32_t)mh * 128 * 16;
              const uint64_t ah = make_desc(a_hi + 2 * kb * a_lbo + row_off, a_lbo, 128);
              const uint64_t al = make_desc(a_lo + 2 * kb * a_lbo + row_off, a_lbo, 128);
              const uint32_t dm = tmem_d + (uint32_t)(mh * a.mh_stride + chain * a.acc_cols);
              const uint32_t dc = tmem_d + (uint32_t)(mh * a.mh_stride + corr * a.acc_cols);
              const uint32_t bm = 1u << (mh * 8 + chain), bc = 1u << (mh * 8 + corr);
              mma_ss<TF32>(dm, ah, wh, idesc, (started & bm) != 0);
              started |= bm;
              mma_ss<TF32>(dc, ah, wl, idesc, (started & bc) != 0);
              started |= bc;
              mma_ss<TF32>(dc, al, wh, idesc, true);
            }
          }
          mma_commit(&bar.w_empty[ws]);                 // ring slot free once these MMAs have read it
        }
        mma_commit(&bar.a_empty[as]);
      }
      mma_commit(&bar.d_full);                          // commits are cumulative: every MMA above has retired
    }
  } else {
    // =========================== producers: stage activation chunks ========================================
    const float* xb = a.x.p + (long long)b * a.x.bs;
    for (int kc = 0; kc < n_kc; ++kc) {
      const int as = kc % A_SLOTS;
      if (kc >= A_SLOTS) mbar_wait(&bar.a_empty[as], ((kc / A_SLOTS) - 1) & 1);
      uint8_t* A_hi = A_ring + size_t(as) * 2 * a_part;
      uint8_t* A_lo = A_hi + a_part;
      const int c0 = kc * KC;
      const int items = (KC / E) * R;
      // two work items per thread per trip: all 2*E global loads are issued before any conversion
      for (int base = tid; base < items; base += 2 * PROD_THREADS) {
        float v[2][E];
        int offs[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int idx = base + s * PROD_THREADS;
          offs[s] = -1;
#pragma unroll
          for (int e = 0; e < E; ++e) v[s][e] = 0.f;
          if (idx < items) {
            const int g = idx / R, r = idx - g * R;
            const int t = t0 - a.pad + r;
            offs[s] = (g * R + r) * 16;
            if (t >= 0 && t < L) {
              const float* xr = xb + (long long)(c0 + g * E) * a.x.cs + t;
#pragma unroll
              for (int e = 0; e < E; ++e) v[s][e] = __ldg(xr + (long long)e * a.x.cs);
            }
          }
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (offs[s] < 0) continue;
          if (a.pre == PRE_LRELU) {
#pragma unroll
            for (int e = 0; e < E; ++e) v[s][e] = v[s][e] > 0.f ? v[s][e] : v[s][e] * a.slope;
          }
          if (TF32) {
            float h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = to_tf32(v[s][e]);
            *reinterpret_cast<float4*>(A_hi + offs[s]) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(A_lo + offs[s]) =
                make_float4(v[s][0] - h[0], v[s][1] - h[1], v[s][2] - h[2], v[s][3] - h[3]);
          } else {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              const float ph = __bfloat162float(__float2bfloat16_rn(v[s][e % E])),
                          qh = __bfloat162float(__float2bfloat16_rn(v[s][(e + 1) % E]));
              hi[e >> 1] = pack_bf16(ph, qh);
              lo[e >> 1] = pack_bf16(v[s][e % E] - ph, v[s][(e + 1) % E] - qh);
            }
            *reinterpret_cast<uint4*>(A_hi + offs[s]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(A_lo + offs[s]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> tensor-core proxy
      mbar_arrive(&bar.a_full[as]);
    }

    // =========================== epilogue ====================================================================
    mbar_wait(&bar.d_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int q = warp & 3, half = warp >> 2;
    float* yb = a.y.p ? a.y.p + (long long)b * a.y.bs : nullptr;
    float* y2b = a.y2.p ? a.y2.p + (long long)b * a.y2.bs : nullptr;
    const float* rb = a.r.p ? a.r.p + (long long)b * a.r.bs : nullptr;
    const int n_chunks = N / 16;
    for (int mh = 0; mh < mh_live; ++mh) {
      const int t = t0 + mh * 128 + q * 32 + lane;
      for (int c = half; c < n_chunks; c += 2) {
        float v[16];
        const uint32_t tbase = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * a.mh_stride + c * 16);
        tmem_ld16(tbase, v);
        const int n_acc = a.chains + (a.sep_corr ? 1 : 0);
        for (int ai = 1; ai < n_acc; ++ai) {             // fp32 round-to-nearest combine of the partial sums
          float p[16];
          tmem_ld16(tbase + (uint32_t)(ai * a.acc_cols), p);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += p[i];
        }
        if (t >= Lq) continue;
        const int row0 = n0 + c * 16;
        if (a.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(a.bias + row0 + i);
        }
        if (a.bias_item) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(a.bias_item + (long long)b * a.bias_item_stride + row0 + i);
        }
        if (a.epi == EPI_GATE) {
#pragma unroll
          for (int i = 0; i < 16; i += 2)
            yb[(long long)((row0 + i) >> 1) * a.y.cs + t] = wn_gate(v[i], v[i + 1]);
          continue;
        }
        // one compact loop per epilogue kind (the switch stays outside the unrolled loops: smaller instruction footprint)
        switch (a.epi) {
          case EPI_RELU:
#pragma unroll
            for (int i = 0; i < 16; ++i) yb[(long long)(row0 + i) * a.y.cs + t] = fmaxf(v[i], 0.f);
            break;
          case EPI_RES:
#pragma unroll
            for (int i = 0; i < 16; ++i) yb[(long long)(row0 + i) * a.y.cs + t] = v[i] + rb[(long long)(row0 + i) * a.r.cs + t];
            break;
          case EPI_SUBFROM:
#pragma unroll
            for (int i = 0; i < 16; ++i) yb[(long long)(row0 + i) * a.y.cs + t] = rb[(long long)(row0 + i) * a.r.cs + t] - v[i];
            break;
          case EPI_WN:
            for (int i = 0; i < 16; ++i) {
              const int row = row0 + i;
              if (row < a.split) yb[(long long)row * a.y.cs + t] = rb[(long long)row * a.r.cs + t] + v[i];
              else {
                float* o = y2b + (long long)(row - a.split) * a.y2.cs + t;
                *o = a.first ? v[i] : *o + v[i];
              }
            }
            break;
          case EPI_UPSAMPLE:
            for (int i = 0; i < 16; ++i) {
              const int row = row0 + i;
              const int co = row / a.up, phi = row - co * a.up;
              const int to = t * a.up + phi - a.up_pad;
              if (to >= 0 && to < L * a.up) yb[(long long)co * a.y.cs + to] = v[i];
            }
            break;
          case EPI_MRF:
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float v2 = v[i] + rb[(long long)(row0 + i) * a.r.cs + t];
              float* o = y2b + (long long)(row0 + i) * a.y2.cs + t;
              if (a.mrf == 0) *o = v2;
              else if (a.mrf == 1) *o = *o + v2;
              else *o = (*o + v2) / (float)a.mrf_n;
            }
            break;
          default:
#pragma unroll
            for (int i = 0; i < 16; ++i) yb[(long long)(row0 + i) * a.y.cs + t] = v[i];
            break;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
}


// =====================================================================================================================
// Persistent variant: one CTA per SM walks a static list of tiles; every phase of tile i+1 overlaps tile i.
//
//   warp 0  lane 0 : TMA of the raw fp32 activation rows (cp.async.bulk, one copy per channel row) -> raw ring
//   warp 1  lane 0 : TMA of the weight units                                                      -> W ring
//   warp 2         : TMEM owner; lane 0 issues tcgen05.mma / tcgen05.commit
//   warps 3-6      : converters  raw fp32 (smem) -> leaky-relu -> hi/lo split -> operand layout   -> A ring
//   warps 7-14     : epilogue    TMEM (double-buffered) -> registers -> fused epilogue -> global
//
// All hand-offs are mbarriers; global memory is touched only by the TMA engine (loads) and the epilogue (residual
// loads + stores), so the number of bytes in flight no longer depends on how many threads happen to be staging.
constexpr int P_CONV_WARP0 = 3, P_CONV_THREADS = 128;
constexpr int P_EPI_WARP0 = 7, P_EPI_THREADS = 256;
constexpr int P_THREADS = 32 * 15;
constexpr int P_RAW_SLOTS = 2, P_A_SLOTS = 2, P_W_SLOTS = 4, P_T_SLOTS = 2;

// ConvTranspose pixel-shuffle store of one 16-row chunk: rows (co*UP + phase) are contiguous, so each thread owns UP
// consecutive output samples per output channel -> vector stores, a warp writes 32*UP*4 contiguous bytes per channel.
template <int UP>
__device__ __forceinline__ void store_upsampled(const float (&v)[16], float* yb, int cs, int row0, int t, int up_pad, int Lout) {
#pragma unroll
  for (int i = 0; i < 16; i += UP) {
    const int co = (row0 + i) / UP;
    const int to = t * UP - up_pad;
    float* dst = yb + (long long)co * cs + to;
    if (to >= 0 && to + UP <= Lout) {
      if (UP % 4 == 0 && (to & 3) == 0) {
#pragma unroll
        for (int j = 0; j < UP; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[i + j], v[i + j + 1], v[i + j + 2], v[i + j + 3]);
      } else if (UP % 2 == 0 && (to & 1) == 0) {
#pragma unroll
        for (int j = 0; j < UP; j += 2) *reinterpret_cast<float2*>(dst + j) = make_float2(v[i + j], v[i + j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < UP; ++j) dst[j] = v[i + j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < UP; ++j)
        if (to + j >= 0 && to + j < Lout) dst[j] = v[i + j];
    }
  }
}

struct PBarriers {
  uint64_t raw_full[P_RAW_SLOTS], raw_empty[P_RAW_SLOTS], a_full[P_A_SLOTS], a_empty[P_A_SLOTS], w_full[P_W_SLOTS],
      w_empty[P_W_SLOTS], t_full[P_T_SLOTS], t_empty[P_T_SLOTS];
};

// UNI (experimental, PIPER_B200_UNI=1, off by default): the two TMA warps run their loops converged with only the copy
// predicated on an elected lane, like the MMA warp, so the copy operands stay in uniform registers.  In the shipped
// form (`if (lane == 0)`) every cp.async.bulk costs ~18 dependent instructions (vector address arithmetic, 3-4 R2UR, an
// ELECT / BRA.U.ANY loop): ~150 cycles per activation ROW, which fits the launch times of every many-channel layer in
// profiles/r01_layer_report.txt (tiles per CTA x C_in x ~76 ns) - see DESIGN.md section 8.
template <bool TF32, int MT, bool UNI = false>
__global__ void __launch_bounds__(P_THREADS, 1) conv_mma_persist_kernel(const MmaConvArgs a) {
  pdl_launch_dependents();   // programmatic dependent launch (launch.cuh)
  constexpr int ES = TF32 ? 4 : 2;
  constexpr int E = 16 / ES;
  constexpr int KSTEP = 2 * E;
  constexpr int MH = MT / 128;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) PBarriers bar;
  __shared__ uint32_t tmem_base_s;

  // warp index through a shuffle ("canonical warp index"): the role branches below are then provably warp-uniform and
  // the MMA warp's operands can live in uniform registers
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int KC = a.kc, R = a.stage_rows, RS = a.raw_stride;
  const int raw_bytes = KC * RS * 4, a_part = KC * R * ES, w_part = KC * a.n_tile * ES;
  uint8_t* RAW_ring = smem;
  uint8_t* A_ring = RAW_ring + size_t(P_RAW_SLOTS) * raw_bytes;
  uint8_t* W_ring = A_ring + size_t(P_A_SLOTS) * 2 * a_part;
  const int n_kc = a.ci / KC;
  const int n_units = n_kc * a.k;
  const uint32_t unit_bytes = 2u * (uint32_t)w_part;
  const int tpi = a.tiles_per_item, total = a.total_tiles;
  const int t_slots = a.t_slots;                      // 1 or 2 TMEM accumulator sets
  const int set_cols = MH * a.mh_stride;

  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < P_RAW_SLOTS; ++i) { mbar_init(&bar.raw_full[i], 1); mbar_init(&bar.raw_empty[i], P_CONV_THREADS); }
    for (int i = 0; i < P_A_SLOTS; ++i) { mbar_init(&bar.a_full[i], P_CONV_THREADS); mbar_init(&bar.a_empty[i], 1); }
    for (int i = 0; i < P_W_SLOTS; ++i) { mbar_init(&bar.w_full[i], 1); mbar_init(&bar.w_empty[i], 1); }
    for (int i = 0; i < P_T_SLOTS; ++i) { mbar_init(&bar.t_full[i], 1); mbar_init(&bar.t_empty[i], P_EPI_THREADS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  pdl_wait();                // the prologue above overlapped the previous grid; nothing below runs before it has completed

  // tile id -> (output-row tile, item, time block); every role walks the same list and skips the same tiles
  auto decode = [&](int tile, int& nt, int& b, int& t0, int& L, int& Lq) {
    const int tb = tile % tpi;
    const int rest = tile / tpi;
    b = rest % a.batch;
    nt = rest / a.batch;
    t0 = tb * MT;
    L = a.len[b] * a.len_scale;
    Lq = L + a.q_extra;
    return t0 < Lq;
  };

  if (UNI && warp == 0) {
    // ---------------------------------------------------------------------- raw activation rows via TMA, uniform issue
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int nt, b, t0, L, Lq;
      bool ok = decode(tile, nt, b, t0, L, Lq);
      L = __shfl_sync(0xffffffffu, L, 0);                              // loaded from global: make uniformity explicit
      ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
      if (!ok) continue;
      const int t_lo = t0 - a.pad;
      const int t_base = t_lo & ~3;
      const int g0 = max(t_lo, 0) & ~3;
      const int g1 = min((min(t_lo + R, L) + 3) & ~3, a.x.cs);
      const uint32_t row_bytes = (uint32_t)(g1 - g0) * 4;
      const float* xb = a.x.p + (long long)b * a.x.bs + g0;
      for (int kc = 0; kc < n_kc; ++kc, ++it) {
        const int s = it % P_RAW_SLOTS;
        if (it >= P_RAW_SLOTS) mbar_wait_t<UNI>(&bar.raw_empty[s], ((it / P_RAW_SLOTS) - 1) & 1);
        if (elect_one()) mbar_expect_tx(&bar.raw_full[s], row_bytes * (uint32_t)KC);
        const uint32_t dst = smem_u32(RAW_ring + size_t(s) * raw_bytes) + (uint32_t)(g0 - t_base) * 4;
        const float* src = xb + (long long)(kc * KC) * a.x.cs;
        const uint32_t mb = smem_u32(&bar.raw_full[s]);
        const uint32_t d_step = (uint32_t)RS * 4u;
        const long long s_step = a.x.cs;
        uint32_t d = dst;
#pragma unroll 4
        for (int c = 0; c < KC; ++c, d += d_step, src += s_step) {     // running addresses: ~4 uniform ops per copy
          if (elect_one())
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
                         "l"(src), "r"(row_bytes), "r"(mb)
                         : "memory");
        }
        __syncwarp();
      }
    }
  } else if (UNI && warp == 1) {
    // ---------------------------------------------------------------------- weight units via TMA, uniform issue
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int nt, b, t0, L, Lq;
      bool ok = decode(tile, nt, b, t0, L, Lq);
      ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
      if (!ok) continue;
      const size_t g_row = size_t(a.n_tile) * 16, part_all = size_t(a.ci / E) * g_row;
      const uint8_t* wsrc = a.w + size_t(nt) * a.k * 2 * part_all;
      for (int u = 0; u < n_units; ++u, ++it) {
        const int s = it % P_W_SLOTS;
        if (it >= P_W_SLOTS) mbar_wait_t<UNI>(&bar.w_empty[s], ((it / P_W_SLOTS) - 1) & 1);
        const int kc = u / a.k, j = u - kc * a.k;
        const uint8_t* hi = wsrc + size_t(j) * 2 * part_all + size_t(kc) * (KC / E) * g_row;
        if (elect_one()) {
          mbar_expect_tx(&bar.w_full[s], unit_bytes);
          bulk_g2s(W_ring + size_t(s) * unit_bytes, hi, (uint32_t)w_part, &bar.w_full[s]);
          bulk_g2s(W_ring + size_t(s) * unit_bytes + w_part, hi + part_all, (uint32_t)w_part, &bar.w_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ raw activation rows via TMA
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int nt, b, t0, L, Lq;
        if (!decode(tile, nt, b, t0, L, Lq)) continue;
        const int t_lo = t0 - a.pad;
        const int t_base = t_lo & ~3;                                   // smem column 0 <-> time t_base
        const int g0 = max(t_lo, 0) & ~3;                               // first / one-past-last float fetched
        const int g1 = min((min(t_lo + R, L) + 3) & ~3, a.x.cs);
        const uint32_t row_bytes = (uint32_t)(g1 - g0) * 4;
        const float* xb = a.x.p + (long long)b * a.x.bs + g0;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int s = it % P_RAW_SLOTS;
          if (it >= P_RAW_SLOTS) mbar_wait_t<UNI>(&bar.raw_empty[s], ((it / P_RAW_SLOTS) - 1) & 1);
          mbar_expect_tx(&bar.raw_full[s], row_bytes * (uint32_t)KC);
          uint8_t* dst = RAW_ring + size_t(s) * raw_bytes + (size_t)(g0 - t_base) * 4;
          for (int c = 0; c < KC; ++c)
            bulk_g2s(dst + (size_t)c * RS * 4, xb + (long long)(kc * KC + c) * a.x.cs, row_bytes, &bar.raw_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ weight units via TMA
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int nt, b, t0, L, Lq;
        if (!decode(tile, nt, b, t0, L, Lq)) continue;
        const size_t g_row = size_t(a.n_tile) * 16, part_all = size_t(a.ci / E) * g_row;
        const uint8_t* wsrc = a.w + size_t(nt) * a.k * 2 * part_all;
        for (int u = 0; u < n_units; ++u, ++it) {
          const int s = it % P_W_SLOTS;
          if (it >= P_W_SLOTS) mbar_wait_t<UNI>(&bar.w_empty[s], ((it / P_W_SLOTS) - 1) & 1);
          const int kc = u / a.k, j = u - kc * a.k;
          const uint8_t* hi = wsrc + size_t(j) * 2 * part_all + size_t(kc) * (KC / E) * g_row;
          mbar_expect_tx(&bar.w_full[s], unit_bytes);
          bulk_g2s(W_ring + size_t(s) * unit_bytes, hi, (uint32_t)w_part, &bar.w_full[s]);
          bulk_g2s(W_ring + size_t(s) * unit_bytes + w_part, hi + part_all, (uint32_t)w_part, &bar.w_full[s]);
        }
      }
    }
  } else if (warp == 2) {
    {
      // ------------------------------------------------------------------ MMA issue: the WHOLE warp runs the loop,
      // only the tcgen05 instructions are predicated on one elected lane (see issue_tap)
      const uint32_t tmem_du = __shfl_sync(0xffffffffu, tmem_d, 0);
      const uint32_t a_lbo = (uint32_t)R * 16, w_lbo = (uint32_t)a.n_tile * 16;
      const uint32_t idesc = make_idesc(TF32, 128, a.n_tile);
      const uint32_t a_step = 2u * (uint32_t)R, w_step = 2u * (uint32_t)a.n_tile;   // 16-byte units per k-step
      uint32_t a_it = 0, w_it = 0, t_it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int nt, b, t0, L, Lq;
        bool ok = decode(tile, nt, b, t0, L, Lq);
        Lq = __shfl_sync(0xffffffffu, Lq, 0);                      // loaded from global: make its uniformity explicit
        ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
        if (!ok) continue;
        const int mh_live = (Lq - t0 > 128 && MH > 1) ? 2 : 1;
        const int ts = t_it % t_slots;
        if (t_it >= (uint32_t)t_slots) mbar_wait_t<UNI>(&bar.t_empty[ts], ((t_it / t_slots) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t d_set = tmem_du + (uint32_t)(ts * set_cols);
        uint32_t started = 0;
        int u = 0;
        for (int kc = 0; kc < n_kc; ++kc, ++a_it) {
          const int as = a_it % P_A_SLOTS;
          mbar_wait_t<UNI>(&bar.a_full[as], (a_it / P_A_SLOTS) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t a_hi = smem_u32(A_ring + size_t(as) * 2 * a_part);
          const uint32_t ah_base = desc_lo(a_hi, a_lbo), al_base = desc_lo(a_hi + a_part, a_lbo);
          for (int j = 0; j < a.k; ++j, ++u, ++w_it) {
            const int ws = w_it % P_W_SLOTS;
            mbar_wait_t<UNI>(&bar.w_full[ws], (w_it / P_W_SLOTS) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int chain = (u * a.chains) / n_units;
            const int corr = a.sep_corr ? a.chains : chain;
            const uint32_t bm = 1u << chain, bc = 1u << corr;
            const uint32_t acc_m = (started & bm) ? 1u : 0u;
            started |= bm;
            const uint32_t acc_c = (started & bc) ? 1u : 0u;
            started |= bc;
            const uint32_t wh = desc_lo(smem_u32(W_ring + size_t(ws) * unit_bytes), w_lbo);
            const uint32_t row = (uint32_t)(j * a.dil);
            issue_tap<TF32, MH>(ah_base + row, al_base + row, wh, wh + ((uint32_t)w_part >> 4), KC / KSTEP, a_step, w_step,
                                mh_live == 2, d_set + (uint32_t)(chain * a.acc_cols), d_set + (uint32_t)(corr * a.acc_cols),
                                (uint32_t)a.mh_stride, idesc, acc_m, acc_c);
            if (elect_one()) mma_commit(&bar.w_empty[ws]);
            __syncwarp();
          }
          if (elect_one()) mma_commit(&bar.a_empty[as]);
          __syncwarp();
        }
        if (elect_one()) mma_commit(&bar.t_full[ts]);
        __syncwarp();
        ++t_it;
      }
    }
  } else if (warp < P_EPI_WARP0) {
    // -------------------------------------------------------------------- converters
    const int ctid = tid - P_CONV_WARP0 * 32;
    uint32_t raw_it = 0, a_it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int nt, b, t0, L, Lq;
      if (!decode(tile, nt, b, t0, L, Lq)) continue;
      const int t_lo = t0 - a.pad;
      const int off = t_lo - (t_lo & ~3);                               // smem column of stage row 0
      for (int kc = 0; kc < n_kc; ++kc, ++raw_it, ++a_it) {
        const int rs = raw_it % P_RAW_SLOTS, as = a_it % P_A_SLOTS;
        mbar_wait_t<UNI>(&bar.raw_full[rs], (raw_it / P_RAW_SLOTS) & 1);
        if (a_it >= P_A_SLOTS) mbar_wait_t<UNI>(&bar.a_empty[as], ((a_it / P_A_SLOTS) - 1) & 1);
        const float* raw = reinterpret_cast<const float*>(RAW_ring + size_t(rs) * raw_bytes) + off;
        uint8_t* A_hi = A_ring + size_t(as) * 2 * a_part;
        uint8_t* A_lo = A_hi + a_part;
        for (int g = 0; g < KC / E; ++g) {
          const float* rg = raw + (size_t)(g * E) * RS;
          for (int r = ctid; r < R; r += P_CONV_THREADS) {
            const int t = t_lo + r;
            const bool live = t >= 0 && t < L;                          // outside the utterance: zeros, whatever the
            float v[E];                                                 // (unwritten / stale) smem holds
#pragma unroll
            for (int e = 0; e < E; ++e) {
              float x = live ? rg[(size_t)e * RS + r] : 0.f;
              if (a.pre == PRE_LRELU) x = x > 0.f ? x : x * a.slope;
              v[e] = x;
            }
            const int o = (g * R + r) * 16;
            if (TF32) {
              float h[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = to_tf32(v[e % E]);
              *reinterpret_cast<float4*>(A_hi + o) = make_float4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<float4*>(A_lo + o) =
                  make_float4(v[0] - h[0], v[1 % E] - h[1], v[2 % E] - h[2], v[3 % E] - h[3]);
            } else {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                const float ph = __bfloat162float(__float2bfloat16_rn(v[e % E])),
                            qh = __bfloat162float(__float2bfloat16_rn(v[(e + 1) % E]));
                hi[e >> 1] = pack_bf16(ph, qh);
                lo[e >> 1] = pack_bf16(v[e % E] - ph, v[(e + 1) % E] - qh);
              }
              *reinterpret_cast<uint4*>(A_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(A_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        mbar_arrive(&bar.a_full[as]);
        mbar_arrive(&bar.raw_empty[rs]);
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue
    const int ew = warp - P_EPI_WARP0;                 // 0..7
    const int q = warp & 3, half = ew >> 2;            // TMEM lane quadrant is fixed by warp id % 4
    uint32_t t_it = 0;
    const int n_chunks = a.n_tile / 16;
    const int n_acc = a.chains + (a.sep_corr ? 1 : 0);
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int nt, b, t0, L, Lq;
      if (!decode(tile, nt, b, t0, L, Lq)) continue;
      const int mh_live = (Lq - t0 > 128 && MH > 1) ? 2 : 1;
      const int ts = t_it % t_slots;
      mbar_wait_t<UNI>(&bar.t_full[ts], (t_it / t_slots) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t d_set = tmem_d + (uint32_t)(ts * set_cols);
      float* yb = a.y.p ? a.y.p + (long long)b * a.y.bs : nullptr;
      float* y2b = a.y2.p ? a.y2.p + (long long)b * a.y2.bs : nullptr;
      const float* rb = a.r.p ? a.r.p + (long long)b * a.r.bs : nullptr;
      const int n0 = nt * a.n_tile;
      // this thread's share of the tile: chunks c = half, half+2, ... of each live row half
      const int my_chunks = (n_chunks - half + 1) / 2;
      const int work = mh_live * my_chunks;
      for (int wi = 0; wi < work; ++wi) {
        const int mh = wi / my_chunks, c = half + 2 * (wi - mh * my_chunks);
        const int t = t0 + mh * 128 + q * 32 + lane;
        float v[16];
        const uint32_t tbase = d_set + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * a.mh_stride + c * 16);
        tmem_ld16(tbase, v);
        for (int ai = 1; ai < n_acc; ++ai) {
          float p[16];
          tmem_ld16(tbase + (uint32_t)(ai * a.acc_cols), p);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += p[i];
        }
        if (wi == work - 1) {                          // all of this thread's TMEM reads are done: release the set
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          mbar_arrive(&bar.t_empty[ts]);
        }
        if (t >= Lq) continue;
        const int row0 = n0 + c * 16;
        if (a.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(a.bias + row0 + i);
        }
        if (a.bias_item) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(a.bias_item + (long long)b * a.bias_item_stride + row0 + i);
        }
        if (a.epi == EPI_GATE) {
#pragma unroll
          for (int i = 0; i < 16; i += 2)
            yb[(long long)((row0 + i) >> 1) * a.y.cs + t] = wn_gate(v[i], v[i + 1]);
          continue;
        }
        if (a.epi == EPI_RES || a.epi == EPI_MRF || a.epi == EPI_SUBFROM) {
          float rv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) rv[i] = rb[(long long)(row0 + i) * a.r.cs + t];   // 16 loads in flight
          if (a.epi == EPI_SUBFROM) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = rv[i] - v[i];
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += rv[i];
          }
        }
        if (a.epi == EPI_UPSAMPLE && (a.up == 8 || a.up == 4 || a.up == 2)) {
          if (a.up == 8) store_upsampled<8>(v, yb, a.y.cs, row0, t, a.up_pad, L * 8);
          else if (a.up == 4) store_upsampled<4>(v, yb, a.y.cs, row0, t, a.up_pad, L * 4);
          else store_upsampled<2>(v, yb, a.y.cs, row0, t, a.up_pad, L * 2);
          continue;
        }
        if (a.epi == EPI_MRF) {
          if (a.mrf == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y2b[(long long)(row0 + i) * a.y2.cs + t] = v[i];
          } else {
            float ov[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = y2b[(long long)(row0 + i) * a.y2.cs + t];
            const float inv_n = (float)a.mrf_n;
            if (a.mrf == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) y2b[(long long)(row0 + i) * a.y2.cs + t] = ov[i] + v[i];
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) y2b[(long long)(row0 + i) * a.y2.cs + t] = (ov[i] + v[i]) / inv_n;
            }
          }
          continue;
        }
        // the switch is OUTSIDE the unrolled loops: one compact 16-wide loop per epilogue kind keeps the epilogue's
        // instruction footprint small (15 warps in 5 roles share the instruction caches)
        switch (a.epi) {
          case EPI_RELU:
#pragma unroll
            for (int i = 0; i < 16; ++i) yb[(long long)(row0 + i) * a.y.cs + t] = fmaxf(v[i], 0.f);
            break;
          case EPI_WN:
            if (row0 + 16 <= a.split) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                yb[(long long)(row0 + i) * a.y.cs + t] = rb[(long long)(row0 + i) * a.r.cs + t] + v[i];
            } else if (row0 >= a.split) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float* o = y2b + (long long)(row0 + i - a.split) * a.y2.cs + t;
                *o = a.first ? v[i] : *o + v[i];
              }
            } else {
              for (int i = 0; i < 16; ++i) {
                const int row = row0 + i;
                if (row < a.split) yb[(long long)row * a.y.cs + t] = rb[(long long)row * a.r.cs + t] + v[i];
                else {
                  float* o = y2b + (long long)(row - a.split) * a.y2.cs + t;
                  *o = a.first ? v[i] : *o + v[i];
                }
              }
            }
            break;
          case EPI_UPSAMPLE:                              // generic stride (the common ones took the vector path above)
            for (int i = 0; i < 16; ++i) {
              const int row = row0 + i;
              const int co = row / a.up, phi = row - co * a.up;
              const int to = t * a.up + phi - a.up_pad;
              if (to >= 0 && to < L * a.up) yb[(long long)co * a.y.cs + to] = v[i];
            }
            break;
          default:                                        // EPI_BIAS, and EPI_RES / EPI_SUBFROM after the residual fold
#pragma unroll
            for (int i = 0; i < 16; ++i) yb[(long long)(row0 + i) * a.y.cs + t] = v[i];
            break;
        }
      }
      if (work == 0) {                                   // (cannot happen: n_chunks >= 1) keep the protocol total
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        mbar_arrive(&bar.t_empty[ts]);
      }
      ++t_it;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
}

// (developer microbenchmark below; its findings are summarised in DESIGN.md)
// ---- microbenchmark: cycles per tcgen05.mma (SS operands, no-swizzle K-major) for a given N, dependent vs rotating
// accumulators.  Developer tool (tools/mma_bench.py); results are quoted in DESIGN.md.
__global__ void __launch_bounds__(128) mma_bench_kernel(int N, int tf32, int n_acc, int iters, int a_rows_shift,
                                                        unsigned long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) { mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  pdl_wait();                // the prologue above overlapped the previous grid; nothing below runs before it has completed
  if (warp == 0) {
    // whole warp runs the loop; only the tcgen05 instructions are predicated on the elected lane
    const int mode = a_rows_shift;
    const int M = (mode & 256) ? 64 : 128;
    const uint32_t layout = (uint32_t)((mode >> 4) & 7);
    const uint32_t idesc = make_idesc(tf32 != 0, M, N);
    const uint32_t a_lbo = layout ? 16 : 256 * 16, w_lbo = layout ? 16 : (uint32_t)N * 16;
    const uint32_t a0 = desc_lo(smem_u32(smem), a_lbo), w0 = desc_lo(smem_u32(smem + 16 * 1024), w_lbo);
    const uint64_t hi_extra = ((uint64_t)layout << 61) | (layout ? ((uint64_t)((1024u >> 4) - (128u >> 4)) << 32) : 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem_d + (uint32_t)((i % n_acc) * N);
      const uint64_t ad = desc_join(a0 + (uint32_t)((i * (mode & 3)) & 63)) + hi_extra, wd = desc_join(w0) + hi_extra;
      const int reps = (mode & 512) ? 0 : (mode & 1024) ? 4 : 1;
      if (elect_one()) {
        for (int r = 0; r < reps; ++r) {
          if (tf32) mma_ss_imm<true, true>(d, ad, wd, idesc);
          else mma_ss_imm<false, true>(d, ad, wd, idesc);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) mma_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (tid == 0) {
      out[0] = (unsigned long long)(t1 - t0);
      out[1] = (unsigned long long)(t2 - t0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(512u) : "memory");
}

int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

}  // namespace

// Tiling of one layer for the tensor-core path.  Returns false when the shape is outside what the kernel handles.
bool mma_plan(int ci, int rows, int k, int dil, bool tf32, MmaPlan& p) {
  const int es = tf32 ? 4 : 2, kstep = tf32 ? 8 : 16;
  if (ci % kstep != 0 || rows % 16 != 0 || rows < 16) return false;
  // accumulators per row half: tf32x3 layers use 2 K-chains + 1 correction accumulator (see the MMA issuer)
  p.chains = tf32 ? 2 : 1;
  p.sep_corr = tf32;
  const int n_acc = p.chains + (p.sep_corr ? 1 : 0);
  // output-row tile: multiple of 16 dividing the rows evenly; <= 256 columns of TMEM per CTA for all its
  // accumulators when several are needed (two CTAs per SM), <= 256 rows otherwise
  // multi-accumulator (tf32x3) layers: 128-row tiles when the reduction is long (the epilogue is a small share of the
  // tile; 3 x 128 columns fill TMEM so it cannot be double-buffered), 64-row tiles with double-buffered TMEM otherwise
  const int max_tile = n_acc > 1 ? ((ci * k >= 900 && rows >= 256) ? 128 : 64) : 256;
  int nt = 1;
  while (rows / nt > max_tile || rows % nt != 0 || (rows / nt) % 16 != 0) {
    if (++nt > 64) return false;
  }
  p.n_tile = rows / nt;
  p.n_tiles = nt;
  p.acc_cols = n_acc > 1 ? ((p.n_tile + 31) & ~31) : pow2_cols(p.n_tile);
  // Occupancy first: most layers are bound by memory latency/bandwidth or by weight streaming, which several
  // small co-resident CTAs hide better than one deeply pipelined CTA.  Try footprints of <= 56 KB (4 CTAs / SM,
  // single activation slot), then 100 KB and 200 KB (double-buffered activations).
  struct Try { size_t budget; int a_slots, w_slots; };
  const Try tries[] = {{size_t(56) << 10, 1, 3}, {size_t(100) << 10, 2, 3}, {size_t(200) << 10, 2, 3}};
  for (const Try& tr : tries)
    for (int mt : {256, 128}) {
      if (n_acc > 1 && mt != 128) continue;
      if (mt / 128 * n_acc * p.acc_cols > 512) continue;
      const int stage_rows = (mt + (k - 1) * dil + 7) & ~7;
      if (stage_rows * 16 >= (1 << 18)) continue;
      for (int kc = ci; kc >= kstep; kc -= kstep) {     // largest channel chunk that divides ci and fits
        if (ci % kc) continue;
        const size_t bytes = size_t(tr.a_slots) * 2 * kc * stage_rows * es + size_t(tr.w_slots) * 2 * kc * p.n_tile * es;
        if (bytes <= tr.budget) {
          p.mt = mt; p.kc = kc; p.stage_rows = stage_rows; p.smem = bytes; p.tf32 = tf32;
          p.a_slots = tr.a_slots; p.w_slots = tr.w_slots;
          p.mh_stride = n_acc * p.acc_cols;
          p.tmem_cols = pow2_cols(mt / 128 * p.mh_stride);
          return true;
        }
      }
    }
  return false;
}

// Persistent configuration for a layer: tile height, channel chunk and TMEM double-buffering within 200 KB.
static bool persist_cfg(const MmaConvArgs& a, const MmaPlan& p, int& mt, int& kc, int& stage_rows, int& raw_stride,
                        int& t_slots, size_t& smem) {
  const int es = p.tf32 ? 4 : 2, kstep = p.tf32 ? 8 : 16;
  const int n_acc = p.chains + (p.sep_corr ? 1 : 0);
  for (int want_slots : {2, 1})
    for (int m : {256, 128}) {
      if (n_acc > 1 && m != 128) continue;
      const int set_cols = m / 128 * n_acc * p.acc_cols;
      if (set_cols * want_slots > 512) continue;
      const int rows = (m + (a.k - 1) * a.dil + 7) & ~7;
      if (rows * 16 >= (1 << 18)) continue;
      const int rs = rows + 8;
      for (int c = a.ci; c >= kstep; c -= kstep) {
        if (a.ci % c) continue;
        const size_t bytes = size_t(P_RAW_SLOTS) * c * rs * 4 + size_t(P_A_SLOTS) * 2 * c * rows * es +
                             size_t(P_W_SLOTS) * 2 * c * p.n_tile * es;
        if (bytes <= (size_t(196) << 10)) {
          mt = m; kc = c; stage_rows = rows; raw_stride = rs; t_slots = want_slots; smem = bytes;
          return true;
        }
      }
    }
  return false;
}

static int g_mma_persist = -1;   // env PIPER_B200_PERSIST (default on)

void launch_conv_mma(MmaConvArgs a, const MmaPlan& p, int B, int max_len, cudaStream_t st) {
  if (B <= 0 || max_len <= 0) return;
  a.n_tile = p.n_tile; a.acc_cols = p.acc_cols;
  a.sep_corr = p.sep_corr ? 1 : 0; a.mh_stride = p.mh_stride;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(conv_mma_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_mma_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_mma_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_mma_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_mma_persist_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_mma_persist_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_mma_persist_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set[dev & 63] = true;
  }
  if (g_mma_persist < 0) {
    const char* e = std::getenv("PIPER_B200_PERSIST");
    g_mma_persist = e ? std::atoi(e) : 1;
  }

  // ---- persistent, fully pipelined kernel when there is enough work to keep every SM busy for several tiles
  int mt = 0, kc = 0, rows = 0, rs = 0, t_slots = 0;
  size_t smem = 0;
  if (g_mma_persist && persist_cfg(a, p, mt, kc, rows, rs, t_slots, smem)) {
    const int tpi = (max_len + mt - 1) / mt;
    const long long total = (long long)tpi * B * p.n_tiles;
    if (total >= 148) {
      a.kc = kc; a.stage_rows = rows; a.raw_stride = rs; a.t_slots = t_slots;
      a.chains = std::min(p.chains, (a.ci / kc) * a.k);
      const int n_acc = p.chains + (p.sep_corr ? 1 : 0);
      a.tmem_cols = pow2_cols(t_slots * (mt / 128) * n_acc * p.acc_cols);
      a.tiles_per_item = tpi; a.total_tiles = (int)total; a.batch = B;
      const int grid = (int)std::min<long long>(total, 148);
      static int g_uni = -1;                           // experimental uniform-issue TMA warps (see the kernel)
      if (g_uni < 0) {
        const char* e = std::getenv("PIPER_B200_UNI");
        g_uni = e ? std::atoi(e) : 0;
        if (g_uni) {
          cudaFuncSetAttribute(conv_mma_persist_kernel<false, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          cudaFuncSetAttribute(conv_mma_persist_kernel<false, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          cudaFuncSetAttribute(conv_mma_persist_kernel<true, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        }
      }
      if (g_uni) {
        if (p.tf32) launch_k(conv_mma_persist_kernel<true, 128, true>, dim3(grid), dim3(P_THREADS), smem, st, a);
        else if (mt == 256) launch_k(conv_mma_persist_kernel<false, 256, true>, dim3(grid), dim3(P_THREADS), smem, st, a);
        else launch_k(conv_mma_persist_kernel<false, 128, true>, dim3(grid), dim3(P_THREADS), smem, st, a);
      } else if (p.tf32) launch_k(conv_mma_persist_kernel<true, 128>, dim3(grid), dim3(P_THREADS), smem, st, a);
      else if (mt == 256) launch_k(conv_mma_persist_kernel<false, 256>, dim3(grid), dim3(P_THREADS), smem, st, a);
      else launch_k(conv_mma_persist_kernel<false, 128>, dim3(grid), dim3(P_THREADS), smem, st, a);
      count_launch();
      return;
    }
  }

  // ---- one tile per CTA (small problems / latency regime)
  a.kc = p.kc; a.stage_rows = p.stage_rows; a.tmem_cols = p.tmem_cols;
  a.a_slots = p.a_slots; a.w_slots = p.w_slots;
  size_t smem_bytes = p.smem;
  // Experimental (PIPER_B200_SMALL=1, off by default): the occupancy-first plan gives small layers a single
  // activation slot and a 16-channel chunk, i.e. a chain of ci/16 dependent load -> convert -> MMA -> commit round
  // trips (~2 us each: the ~30 us floor of the small encoder / duration-predictor launches in
  // profiles/r01_layer_report.txt).  With fewer than ~3 CTAs per SM in the grid there is nothing co-resident to hide
  // that chain, so take the largest double-buffered chunk that fits instead.  The weight layout does not depend on kc.
  static int g_small = -1;
  if (g_small < 0) {
    const char* e = std::getenv("PIPER_B200_SMALL");
    g_small = e ? std::atoi(e) : 0;
  }
  if (g_small && p.a_slots == 1) {
    const long long ctas = (long long)((max_len + p.mt - 1) / p.mt) * B * p.n_tiles;
    if (ctas <= 3 * 148) {
      const int es = p.tf32 ? 4 : 2, kstep = p.tf32 ? 8 : 16;
      for (int kc = a.ci; kc > p.kc; kc -= kstep) {
        if (a.ci % kc) continue;
        const size_t bytes = size_t(2) * 2 * kc * p.stage_rows * es + size_t(p.w_slots) * 2 * kc * p.n_tile * es;
        if (bytes <= (size_t(200) << 10)) {
          a.kc = kc; a.a_slots = 2; smem_bytes = bytes;
          break;
        }
      }
    }
  }
  // never more chains than weight units, or an accumulator would be read without ever being written
  a.chains = std::min(p.chains, (a.ci / a.kc) * a.k);
  mt = p.mt;
  if (mt == 256 && (long long)((max_len + 255) / 256) * B * p.n_tiles < 148) {
    mt = 128;                                        // 128-row tiles double the CTA count
    a.stage_rows = (128 + (a.k - 1) * a.dil + 7) & ~7;
    a.tmem_cols = pow2_cols(p.mh_stride);
  }
  dim3 grid((max_len + mt - 1) / mt, B, p.n_tiles);
  if (p.tf32) {
    if (mt == 256) launch_k(conv_mma_kernel<true, 256>, dim3(grid), dim3(MMA_THREADS), smem_bytes, st, a);
    else launch_k(conv_mma_kernel<true, 128>, dim3(grid), dim3(MMA_THREADS), smem_bytes, st, a);
  } else {
    if (mt == 256) launch_k(conv_mma_kernel<false, 256>, dim3(grid), dim3(MMA_THREADS), smem_bytes, st, a);
    else launch_k(conv_mma_kernel<false, 128>, dim3(grid), dim3(MMA_THREADS), smem_bytes, st, a);
  }
  count_launch();
}

void run_mma_bench(int N, int tf32, int n_acc, int iters, int shift, unsigned long long out[2]) {
  unsigned long long* d = nullptr;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  launch_k(mma_bench_kernel, dim3(1), dim3(128), 48 * 1024, nullptr, N, tf32, n_acc, iters, shift, d);
  cudaDeviceSynchronize();
  cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
}

// ---- host-side packing ---------------------------------------------------------------------------------------
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
static inline float f32_to_tf32_rna(float f) {   // cvt.rna.tf32.f32: round to nearest, ties away from zero
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return f;
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&f, &u, 4);
  return f;
}

// Source weights in the engine's fp32 layout wsrc[ci][k][rows_p] (row fastest).  Output layout
//   [n tile][tap][hi | lo][ci / E][n_tile][E]      E = 8 bf16 or 4 tf32 (fp32 container)
// so any channel chunk of any tap is one contiguous run per part, whatever chunk size a launch picks.
void pack_conv_mma(const float* wsrc, int ci, int k, int rows, int rows_p, const MmaPlan& p, std::vector<uint8_t>& out) {
  const int es = p.tf32 ? 4 : 2, E = 16 / es;
  const size_t part_all = size_t(ci) * p.n_tile * es;
  out.assign(size_t(p.n_tiles) * k * 2 * part_all, 0);
  (void)rows;
  for (int nt = 0; nt < p.n_tiles; ++nt)
    for (int j = 0; j < k; ++j) {
      uint8_t* hi_base = out.data() + (size_t(nt) * k + j) * 2 * part_all;
      uint8_t* lo_base = hi_base + part_all;
      for (int cin = 0; cin < ci; ++cin)
        for (int n = 0; n < p.n_tile; ++n) {
          const float v = wsrc[(size_t(cin) * k + j) * rows_p + nt * p.n_tile + n];
          const size_t pos = (size_t(cin / E) * p.n_tile + n) * E + (cin % E);
          if (p.tf32) {
            const float hi = f32_to_tf32_rna(v), lo = v - hi;
            memcpy(hi_base + pos * 4, &hi, 4);
            memcpy(lo_base + pos * 4, &lo, 4);
          } else {
            const uint16_t hi = f32_to_bf16_rn(v);
            const uint16_t lo = f32_to_bf16_rn(v - bf16_to_f32(hi));
            memcpy(hi_base + pos * 2, &hi, 2);
            memcpy(lo_base + pos * 2, &lo, 2);
          }
        }
    }
}

}  // namespace pb200
