// Body of the experimental fused MRF stage kernel (see mrf_fused.cu for what it computes and how), written against a small
// primitive policy P so that the SAME source is compiled twice:
//   * mrf_fused.cu        P = DevPrim: inline PTX (tcgen05 / TMA / mbarrier), the CUDA kernel
//   * sim/mrf_sim.cpp     P = SimPrim: a functional CPU model (416 std::threads per CTA, blocking mbarriers, TMEM as an
//                         array, tcgen05.mma decoded from its shared-memory / instruction descriptors) used by
//                         tests/test_fused_mrf_sim.py to check index arithmetic, operand layouts and the barrier protocol
//                         against the oracle without a GPU.
// Nothing here may use CUDA built-ins directly; everything hardware-specific goes through P.
#pragma once

namespace pb200 {
namespace mrf {

constexpr int F_C = 32;                   // channels of the stage
constexpr int F_M = 256;                  // rows per GEMM (two 128-row MMA tiles)
constexpr int F_G0 = 12;                  // guard rows of the stage-input operand: >= half-width of every first conv
constexpr int F_GA = 36;                  // guard rows of the chain operands:      >= half-width of every later conv
constexpr int F_R0 = F_M + 2 * F_G0;      // 280 rows
constexpr int F_RA = F_M + 2 * F_GA;      // 328 rows
constexpr int F_XS = F_R0 + 8;            // row stride (floats) of the staged fp32 input
constexpr int F_W_SLOTS = 6;
constexpr int F_TAP_BYTES = 4 * 64 * 16;  // one tap: [ci / 8][W_hi rows 0..31 | W_lo rows 0..31][8 x fp16]
constexpr int F_A0_PART = 4 * F_R0 * 16, F_AC_PART = 4 * F_RA * 16;
constexpr int F_OFF_A0 = F_C * F_XS * 4;
constexpr int F_OFF_AC = F_OFF_A0 + 2 * F_A0_PART;
constexpr int F_OFF_W = F_OFF_AC + MRF_MAX_CHAINS * 2 * F_AC_PART;
constexpr int F_OFF_BIAS = F_OFF_W + F_W_SLOTS * F_TAP_BYTES;
constexpr int F_OFF_POST = F_OFF_BIAS + MRF_MAX_CHAINS * MRF_MAX_STEPS * F_C * 4;   // conv_post weights [32][8]
constexpr int F_POST_MAXK = 7;
constexpr int F_SMEM = F_OFF_POST + F_C * 8 * 4;
static_assert(F_SMEM <= 227 * 1024, "fused MRF stage does not fit shared memory");
// 16 epilogue warps: two per (TMEM lane quadrant, 128-row tile), each owning 16 of the 32 channels of its row.  With 8 warps
// (one thread = one row x 32 channels) the epilogue executed ~3000 instructions per row per tile and was the kernel's
// bottleneck at 38 % issue utilisation (tensor pipe 24 % active): twice the warps = twice the latency hiding.
constexpr int F_CONV_WARP0 = 3, F_EPI_WARP0 = 5, F_CONV_THREADS = 64, F_EPI_THREADS = 512;
constexpr int F_HC = F_C / 2;             // channels per epilogue thread
constexpr int F_THREADS = 21 * 32;        // 672 threads
constexpr uint32_t F_TMEM_CARRIER = 256;

template <class Mbar>
struct FBarriers {
  Mbar raw_full, raw_free, a0_full, a0_free, w_full[F_W_SLOTS], w_empty[F_W_SLOTS], acc_full[2], acc_empty[2],
      a_full[MRF_MAX_CHAINS];
};

MRF_FN uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
// Operands are FP16 hi/lo pairs (fp16x3: 22 mantissa bits at kind::f16's K = 16).
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = F16 (format 0 at bits 7-9 / 10-12),
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24
MRF_FN constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
MRF_FN int imax(int a, int b) { return a > b ? a : b; }
MRF_FN int imin(int a, int b) { return a < b ? a : b; }

// 8 consecutive channels of one row -> one 16-byte operand row of the hi part and one of the lo part
template <class P>
MRF_FN void store_split8(uint8_t* hi_row, uint8_t* lo_row, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 8; e += 2) P::split2_f16(v[e], v[e + 1], hi[e >> 1], lo[e >> 1]);   // 3 conversions per pair instead of 5
  *reinterpret_cast<uint4*>(hi_row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(lo_row) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

template <class P>
MRF_FN void mrf_fused_body(const MrfFusedArgs& a, typename P::Ctx& cx, uint8_t* smem, FBarriers<typename P::Mbar>& bar,
                           uint32_t* tmem_base_s) {
  P::pdl_launch();
  const int tid = cx.tid(), lane = tid & 31, warp = P::bcast0(cx, tid >> 5);
  float* xs = reinterpret_cast<float*>(smem);
  uint8_t* A0 = smem + F_OFF_A0;
  uint8_t* AC = smem + F_OFF_AC;
  uint8_t* Wr = smem + F_OFF_W;
  float* bias_s = reinterpret_cast<float*>(smem + F_OFF_BIAS);
  float* post_s = reinterpret_cast<float*>(smem + F_OFF_POST);
  const int n_chains = a.n_chains, n_steps = a.n_steps, pair = a.pair;
  const int tpi = a.tiles_per_item, total = a.total_tiles;
  const int block = cx.block(), grid = cx.grid();

  // ---- prologue: zero the chain operands once (their guard rows are never written again), biases to shared memory
  for (int i = tid; i < MRF_MAX_CHAINS * 2 * F_AC_PART / 16; i += F_THREADS)
    reinterpret_cast<uint4*>(AC)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < n_chains * n_steps * F_C; i += F_THREADS) bias_s[i] = a.bias[i];
  if (a.post_w)
    for (int i = tid; i < F_C * a.post_k; i += F_THREADS) post_s[(i / a.post_k) * 8 + (i % a.post_k)] = a.post_w[i];
  if (warp == 2) P::tmem_alloc(cx, tmem_base_s, 512u);
  if (tid == 0) {
    P::mbar_init(cx, &bar.raw_full, 1); P::mbar_init(cx, &bar.raw_free, F_EPI_THREADS);
    P::mbar_init(cx, &bar.a0_full, F_CONV_THREADS); P::mbar_init(cx, &bar.a0_free, 1);
    for (int i = 0; i < F_W_SLOTS; ++i) { P::mbar_init(cx, &bar.w_full[i], 1); P::mbar_init(cx, &bar.w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { P::mbar_init(cx, &bar.acc_full[i], 1); P::mbar_init(cx, &bar.acc_empty[i], F_EPI_THREADS); }
    for (int i = 0; i < MRF_MAX_CHAINS; ++i) P::mbar_init(cx, &bar.a_full[i], F_EPI_THREADS);
    P::fence_mbar_init();
  }
  P::fence_async_proxy();                   // the zeroed operands are read by the tensor core
  P::fence_tc_before();
  P::syncthreads(cx);
  P::fence_tc_after();
  const uint32_t tmem_d = *tmem_base_s;
  P::pdl_sync();                                       // programmatic dependent launch: the prologue above overlapped the previous grid

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
    for (int tile = block; tile < total; tile += grid) {
      int b, t0, L;
      bool ok = decode(tile, b, t0, L);
      L = P::bcast0(cx, L);
      ok = P::bcast0(cx, (int)ok) != 0;
      if (!ok) continue;
      const int t_lo = t0 - a.hv - F_G0;                               // position of operand row 0
      const int t_base = t_lo & ~3;                                    // shared-memory column 0 <-> position t_base
      const int g0 = imax(t_lo, 0) & ~3;
      const int g1 = imin((imin(t_lo + F_R0, L) + 3) & ~3, a.x.cs);
      const uint32_t row_bytes = (uint32_t)(g1 - g0) * 4;
      if (ti >= 1) P::mbar_wait(cx, &bar.raw_free, (ti - 1) & 1);
      if (P::elect_one(cx)) P::mbar_expect_tx(cx, &bar.raw_full, row_bytes * (uint32_t)F_C);
      const float* src = a.x.p + (long long)b * a.x.bs + g0;
      uint32_t d = P::saddr(cx, xs + (g0 - t_base));
      const long long s_step = a.x.cs;
#pragma unroll 4
      for (int c = 0; c < F_C; ++c, d += F_XS * 4, src += s_step) {
        if (P::elect_one(cx)) P::bulk_g2s(cx, d, src, row_bytes, &bar.raw_full);
      }
      P::syncwarp();
      ++ti;
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------- weight taps, in the order the MMAs use them
    uint32_t it = 0;
    for (int tile = block; tile < total; tile += grid) {
      int b, t0, L;
      bool ok = decode(tile, b, t0, L);
      ok = P::bcast0(cx, (int)ok) != 0;
      if (!ok) continue;
      const uint8_t* src = a.w;
      for (int s = 0; s < n_steps; ++s)
        for (int c = 0; c < n_chains; ++c)
          for (int j = 0; j < a.k[c]; ++j, ++it, src += F_TAP_BYTES) {
            const int slot = it % F_W_SLOTS;
            if (it >= F_W_SLOTS) P::mbar_wait(cx, &bar.w_empty[slot], ((it / F_W_SLOTS) - 1) & 1);
            if (P::elect_one(cx)) {
              P::mbar_expect_tx(cx, &bar.w_full[slot], F_TAP_BYTES);
              P::bulk_g2s(cx, P::saddr(cx, Wr + slot * F_TAP_BYTES), src, F_TAP_BYTES, &bar.w_full[slot]);
            }
            P::syncwarp();
          }
    }
  } else if (warp == 2) {
    // -------------------------------------------------------------------- MMA issue (whole warp converged; only the
    // tcgen05 instructions are predicated on the elected lane so that every operand stays warp-uniform)
    const uint32_t tmem_du = (uint32_t)P::bcast0(cx, (int)tmem_d);
    constexpr uint32_t idesc64 = make_idesc_f16(128, 64), idesc32 = make_idesc_f16(128, 32);
    uint32_t g_it = 0, w_it = 0, ti = 0, a_par = 0;
    for (int tile = block; tile < total; tile += grid) {
      int b, t0, L;
      bool ok = decode(tile, b, t0, L);
      ok = P::bcast0(cx, (int)ok) != 0;                                  // depends on a global load: make it uniform
      if (!ok) continue;
      for (int s = 0; s < n_steps; ++s)
        for (int c = 0; c < n_chains; ++c, ++g_it) {
          const uint32_t slot = g_it & 1;
          if (g_it >= 2) P::mbar_wait(cx, &bar.acc_empty[slot], ((g_it >> 1) - 1) & 1);
          uint32_t abase, part, lbo, guard;
          if (s == 0) {
            if (c == 0) P::mbar_wait(cx, &bar.a0_full, ti & 1);
            abase = P::saddr(cx, A0); part = F_A0_PART; lbo = F_R0 * 16; guard = F_G0;
          } else {
            P::mbar_wait(cx, &bar.a_full[c], (a_par >> c) & 1);
            a_par ^= 1u << c;
            abase = P::saddr(cx, AC + c * 2 * F_AC_PART); part = F_AC_PART; lbo = F_RA * 16; guard = F_GA;
          }
          P::fence_tc_after();
          const int k = a.k[c], dil = a.dil[c][s];
          const uint32_t row0 = guard - (uint32_t)((k - 1) / 2 * dil);
          const uint32_t d0 = tmem_du + slot * 128u;
          const uint32_t a_step = 2u * (lbo >> 4);
          for (int j = 0; j < k; ++j, ++w_it) {
            const int ws = w_it % F_W_SLOTS;
            P::mbar_wait(cx, &bar.w_full[ws], (w_it / F_W_SLOTS) & 1);
            P::fence_tc_after();
            const uint32_t row = row0 + (uint32_t)(j * dil);
            uint32_t ah = desc_lo(abase + row * 16, lbo), al = desc_lo(abase + part + row * 16, lbo);
            uint32_t wb = desc_lo(P::saddr(cx, Wr + ws * F_TAP_BYTES), 64 * 16);
#pragma unroll 1
            for (int kk = 0; kk < F_C / 16; ++kk) {
              const uint32_t accf = (j > 0 || kk > 0) ? 1u : 0u;
              if (P::elect_one(cx)) {
                P::mma_f16(cx, d0, ah, wb, idesc64, accf);              // rows   0..127: main | correction (hi*lo)
                P::mma_f16(cx, d0 + 32u, al, wb, idesc32, 1u);          //                correction += lo*hi
                P::mma_f16(cx, d0 + 64u, ah + 128u, wb, idesc64, accf); // rows 128..255
                P::mma_f16(cx, d0 + 96u, al + 128u, wb, idesc32, 1u);
              }
              P::syncwarp();
              ah += a_step; al += a_step; wb += 2u * 64u;
            }
            if (P::elect_one(cx)) P::mma_commit(cx, &bar.w_empty[ws]);
            P::syncwarp();
          }
          if (s == 0 && c == n_chains - 1) {
            if (P::elect_one(cx)) P::mma_commit(cx, &bar.a0_free);      // the stage-input operand may be overwritten
            P::syncwarp();
          }
          if (P::elect_one(cx)) P::mma_commit(cx, &bar.acc_full[slot]);
          P::syncwarp();
        }
      ++ti;
    }
  } else if (warp < F_EPI_WARP0) {
    // ---------------------------------------------------------------------- converter: x -> lrelu -> hi/lo operand
    const int ctid = tid - F_CONV_WARP0 * 32;
    uint32_t ti = 0;
    for (int tile = block; tile < total; tile += grid) {
      int b, t0, L;
      if (!decode(tile, b, t0, L)) continue;
      const int t_lo = t0 - a.hv - F_G0;
      const int off = t_lo - (t_lo & ~3);
      P::mbar_wait(cx, &bar.raw_full, ti & 1);
      if (ti >= 1) P::mbar_wait(cx, &bar.a0_free, (ti - 1) & 1);
      for (int idx = ctid; idx < (F_C / 8) * F_R0; idx += F_CONV_THREADS) {
        const int g = idx / F_R0, rho = idx - g * F_R0;
        const int pos = t_lo + rho;
        const bool live = pos >= 0 && pos < L;                           // outside the utterance: zeros, whatever the
        float v[8];                                                      // (unwritten / stale) shared memory holds
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float x = live ? xs[(g * 8 + e) * F_XS + off + rho] : 0.f;
          v[e] = fmaxf(x, x * a.slope);
        }
        store_split8<P>(A0 + (g * F_R0 + rho) * 16, A0 + F_A0_PART + (g * F_R0 + rho) * 16, v);
      }
      P::fence_async_proxy();
      P::mbar_arrive(cx, &bar.a0_full);
      ++ti;
    }
  } else {
    // ---------------------------------------------------------------------- epilogue: one (position, channel half) per thread
    const int ew = warp - F_EPI_WARP0;                 // 0..15 (any four consecutive warps cover the four lane quadrants)
    const int q = warp & 3, m = (ew >> 2) & 1, hc = ew >> 3;   // TMEM lane quadrant (fixed by warp id % 4), 128-row tile, channel half
    const int r = m * 128 + q * 32 + lane;
    const int c0 = hc * F_HC;                          // first channel of this thread
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float inv_n = 1.f / (float)n_chains;          // xs / num_kernels as a multiplication (<= 1 ulp; the divisions were 6 % of the kernel)
    uint32_t g_it = 0;
    for (int tile = block; tile < total; tile += grid) {
      int b, t0, L;
      if (!decode(tile, b, t0, L)) continue;
      const int p0 = t0 - a.hv;
      const int pos = p0 + r;
      const bool inside = pos >= 0 && pos < L;
      const int t_lo = p0 - F_G0;
      const float* xrow = xs + (t_lo - (t_lo & ~3)) + F_G0 + r + c0 * F_XS;   // this position, first own channel, in the staged input
      float sum[F_HC];
#pragma unroll
      for (int i = 0; i < F_HC; ++i) sum[i] = 0.f;
      for (int s = 0; s < n_steps; ++s) {
        const bool closes = (s % pair) == pair - 1;                      // this conv ends a residual unit
        const bool last = s == n_steps - 1;
        for (int c = 0; c < n_chains; ++c, ++g_it) {
          const uint32_t slot = g_it & 1;
          P::mbar_wait(cx, &bar.acc_full[slot], (g_it >> 1) & 1);
          P::fence_tc_after();
          const uint32_t tb = tmem_d + lane_addr + slot * 128u + (uint32_t)m * 64u + (uint32_t)c0;
          float v[F_HC];
          {
            float qv[16];
            P::tmem_ld16(cx, tb, v);                                     // main
            P::tmem_ld16(cx, tb + 32u, qv);                              // correction (hi*lo + lo*hi)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += qv[i];
          }
          P::fence_tc_before();
          P::mbar_arrive(cx, &bar.acc_empty[slot]);
          const float* bs = bias_s + (s * n_chains + c) * F_C + c0;
#pragma unroll
          for (int i = 0; i < F_HC; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bs + i);
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
          const uint32_t carrier = tmem_d + lane_addr + F_TMEM_CARRIER + (uint32_t)c * 64u + (uint32_t)m * 32u + (uint32_t)c0;
          if (closes) {
            if (s == pair - 1) {                                         // residual = the stage input
#pragma unroll
              for (int i = 0; i < F_HC; ++i) v[i] += xrow[i * F_XS];
            } else {                                                     // residual = the value parked by the last unit
              float p[16];
              P::tmem_ld16(cx, carrier, p);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += p[i];
            }
          }
          if (s == pair - 1 && c == n_chains - 1) P::mbar_arrive(cx, &bar.raw_free);   // last read of the staged input
          if (last) {
#pragma unroll
            for (int i = 0; i < F_HC; ++i) sum[i] += v[i];
            if (c == n_chains - 1) {
              const bool stored = r >= a.hv && r < a.hv + a.to && pos < L;
              if (a.post_w == nullptr) {
                if (stored) {
                  float* yb = a.y.p + (long long)b * a.y.bs + (long long)c0 * a.y.cs + pos;
#pragma unroll
                  for (int i = 0; i < F_HC; ++i) yb[(long long)i * a.y.cs] = sum[i] * inv_n;
                }
              } else {
                // Fused generator tail (models.py:364-366): the stage output never goes to HBM.  y -> leaky-relu, zero outside
                // the utterance (conv_post pads its own input) -> the payload rows of chain 0's operand buffer, free until the
                // next tile's first epilogue and never its zero guard rows: 8 blocks [256 rows][4 channels] of 16-byte rows.
                // Each thread parks its 16 channels (4 blocks); the channel-half-0 thread of a row then sums all 32.
#pragma unroll
                for (int bj = 0; bj < F_HC / 4; ++bj) {
                  const int bi = hc * (F_HC / 4) + bj;
                  float w4[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float yv = sum[bj * 4 + e] * inv_n;
                    w4[e] = inside ? fmaxf(yv, yv * a.post_slope) : 0.f;
                  }
                  uint8_t* blk = AC + (bi < 4 ? bi * F_RA * 16 : F_AC_PART + (bi - 4) * F_RA * 16) + F_GA * 16;
                  *reinterpret_cast<float4*>(blk + r * 16) = make_float4(w4[0], w4[1], w4[2], w4[3]);
                }
                P::bar_sync(cx, 1, F_EPI_THREADS);
                if (stored && hc == 0) {
                  const int half_k = (a.post_k - 1) / 2;
                  float acc = 0.f;
                  for (int j = 0; j < a.post_k; ++j) {
                    const int rr = r + j - half_k;                     // within [hv - half_k, 256 - hv + half_k): exact y rows
#pragma unroll
                    for (int bi = 0; bi < F_C / 4; ++bi) {
                      const uint8_t* blk = AC + (bi < 4 ? bi * F_RA * 16 : F_AC_PART + (bi - 4) * F_RA * 16) + F_GA * 16;
                      const float4 y4 = *reinterpret_cast<const float4*>(blk + rr * 16);
                      const float* pw = post_s + bi * 4 * 8 + j;
                      acc += y4.x * pw[0] + y4.y * pw[8] + y4.z * pw[16] + y4.w * pw[24];
                    }
                  }
                  a.audio[a.out_off[b] + pos] = tanhf(acc);
                }
                P::bar_sync(cx, 1, F_EPI_THREADS);                      // the buffer is rewritten by the next tile's first epilogue
              }
            }
          } else {
            if (closes) {
              P::tmem_st16(cx, carrier, v);
              P::tmem_wait_st();
            }
            // operand of this chain's next conv: lrelu, zero outside the utterance (every conv pads its own input);
            // this thread's 16 channels are operand groups 2 hc and 2 hc + 1
            uint8_t* hi = AC + c * 2 * F_AC_PART + (F_GA + r) * 16 + (2 * hc) * F_RA * 16;
            if (inside) {
              const float slope = a.slope;                   // 0 < slope < 1: leaky_relu(x) = max(x, slope x)
#pragma unroll
              for (int g = 0; g < F_HC / 8; ++g) {
                float w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = fmaxf(v[g * 8 + e], v[g * 8 + e] * slope);
                store_split8<P>(hi + g * F_RA * 16, hi + F_AC_PART + g * F_RA * 16, w);
              }
            } else {
#pragma unroll
              for (int g = 0; g < F_HC / 8; ++g) {
                *reinterpret_cast<uint4*>(hi + g * F_RA * 16) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(hi + F_AC_PART + g * F_RA * 16) = make_uint4(0u, 0u, 0u, 0u);
              }
            }
            P::fence_async_proxy();
            P::mbar_arrive(cx, &bar.a_full[c]);
          }
        }
      }
    }
  }

  P::fence_tc_before();
  P::syncthreads(cx);
  if (warp == 2) P::tmem_dealloc(cx, tmem_d, 512u);
}

}  // namespace mrf
}  // namespace pb200
