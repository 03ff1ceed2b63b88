// Relative-position multi-head attention of the text encoder (attentions.py:225-272) on the tensor cores (PIPER_B200_ATT3),
// written against the primitive policy P (tc_policy_dev.cuh on the GPU, tests/sim/sim_prim.h on the CPU) and checked by
// tests/test_att_sim.py without a GPU.
//
// One CTA per (utterance, head, tile of 128 queries).  Keys / values are walked in blocks of 64, twice:
//   pass 1   S~ = Q_hi K_hi^T (one product, enough to place the softmax offset)        -> row maximum m
//   pass 2   S  = Q K^T with all three fp16x3 products, p = exp(S - m), l += sum p,     O += P V (fp16x3)
// so the running output never has to be rescaled.  Every thread of the four softmax warps owns one query row (one TMEM
// lane).  The banded relative-position terms stay off the tensor cores except for the 9 logits per row:
//   S[i][j] += (q_i / sqrt(dk)) . emb_rel_k[j - i + w]   |j - i| <= w   from one tiny GEMM  Rl = Q Ek^T  (N = 16)
//   O[i]    += sum_r p[i][i + r - w] emb_rel_v[r]        9 probabilities kept in registers per row, applied at the end
// Operands are fp16 hi/lo pairs (22 mantissa bits) with the B operands stacked [hi ; lo] along N:
//   D[:, 0..N) += A_hi B_hi^T | D[:, N..2N) += A_hi B_lo^T    (one instruction, N' = 2N)     D[:, N..2N) += A_lo B_hi^T
//
// Roles (10 warps): 0 TMA (q tile, then k / v blocks), 1 MMA issue, 2-5 converters (fp32 -> fp16 hi/lo operand layouts),
// 6-9 softmax / epilogue.  The blocks are processed one after the other (single operand buffers): the gain over the CUDA-core
// kernel is the arithmetic, not the overlap - 190 tensor-core instructions per tile instead of ~70k FMA warp-instructions.
//
// TMEM columns: [0,128) S (main | correction), [128, 128 + 2 dk) O (main | correction), [384, 416) Rl (main | correction).
#pragma once
#ifndef MRF_HD
#ifdef __CUDACC__
#define MRF_HD __host__ __device__ __forceinline__
#else
#define MRF_HD inline
#endif
#endif

namespace pb200 {
namespace att {

constexpr int A_QT = 128;                 // queries per CTA
constexpr int A_KB = 64;                  // keys per block
constexpr int A_MAXDK = 128;
constexpr int A_NREL = 9;                 // 2 * window + 1, window = 4
constexpr int A_CONV_WARP0 = 2, A_CONV_THREADS = 128, A_SM_WARP0 = 6, A_SM_THREADS = 128;
constexpr int A_THREADS = 10 * 32;
constexpr uint32_t A_TM_S = 0, A_TM_O = 128, A_TM_RL = 384;

struct Args {
  View qkv, out;                          // [B][3H][Lp] (q | k | v rows), [B][H][Lp]
  const float* rel_k = nullptr;           // [9][dk]
  const float* rel_v = nullptr;           // [9][dk]
  const int* len = nullptr;
  int H = 0, dk = 0, n_heads = 0, q_tiles = 0;
  int tm = 0;                             // 1: q / k / v windows by tensor-map TMA (tmq: box [dk][136], tmk: box [dk][72])
  int flat = 0;                           // with tm: qkv is laid out [channel][item][slot] -> coordinates (t, item, channel)
  int tail_thr = 0;                       // > 0: the LAST query tile of an utterance is left to the CUDA-core kernel when it holds <= tail_thr rows (encoder.cu)
};

// tensor-map descriptors of the qkv view for the two boxes (host side; shared with the CPU model)
inline void att_tmaps(const Args& a, int B, TmapDesc& tq, TmapDesc& tk) {
  for (TmapDesc* t : {&tq, &tk}) {
    t->base = a.qkv.p;
    if (a.flat) {                         // [3H][B][slot]: dims (t, item, channel)
      t->dims[0] = int(a.qkv.bs); t->dims[1] = B; t->dims[2] = 3 * a.H;
      t->stride1 = a.qkv.bs * 4; t->stride2 = (long long)a.qkv.cs * 4;
      t->box[1] = 1; t->box[2] = a.dk;
    } else {                              // [B][3H][pitch]: dims (t, channel, item)
      t->dims[0] = a.qkv.cs; t->dims[1] = 3 * a.H; t->dims[2] = B;
      t->stride1 = (long long)a.qkv.cs * 4; t->stride2 = a.qkv.bs * 4;
      t->box[1] = a.dk; t->box[2] = 1;
    }
  }
  tq.box[0] = A_QT + 8;
  tk.box[0] = A_KB + 8;
}

// shared-memory map (bytes), G = dk / 8
MRF_HD int off_raw(int) { return 0; }                                           // fp32 rows: q [dk][136] or k | v [2 dk][72]
MRF_HD int raw_bytes(int dk) { return dk * (A_QT + 8) * 4 > 2 * dk * (A_KB + 8) * 4 ? dk * (A_QT + 8) * 4 : 2 * dk * (A_KB + 8) * 4; }
MRF_HD int off_q(int dk) { return raw_bytes(dk); }                              // hi | lo, each [G][128][16 B]
MRF_HD int off_k(int dk) { return off_q(dk) + 2 * (dk / 8) * A_QT * 16; }       // [G][64 hi + 64 lo][16 B]
MRF_HD int off_v(int dk) { return off_k(dk) + (dk / 8) * 2 * A_KB * 16; }       // [8][dk hi + dk lo][16 B]
MRF_HD int off_p(int dk) { return off_v(dk) + (A_KB / 8) * 2 * dk * 16; }       // hi | lo, each [8][128][16 B]
MRF_HD int off_ek(int dk) { return off_p(dk) + 2 * (A_KB / 8) * A_QT * 16; }    // [G][16 hi + 16 lo][16 B]
MRF_HD int off_ev(int dk) { return off_ek(dk) + (dk / 8) * 32 * 16; }           // fp32 [9][dk]
MRF_HD int smem_bytes(int dk) { return off_ev(dk) + A_NREL * dk * 4; }

template <class Mbar>
struct Barriers {
  Mbar raw_full, raw_free, q_full, op_full, s_full, s_free, p_full, kv_free, rl_full;
};

MRF_FN uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
// kind::f16, FP16 operands (format 0), D = F32, K-major A and B
MRF_FN constexpr uint32_t idesc_f16(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
MRF_FN int imin(int a, int b) { return a < b ? a : b; }

template <class P>
MRF_FN void split8(uint8_t* hi_row, uint8_t* lo_row, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 8; e += 2) {
    const float ph = P::f16_round(v[e]), qh = P::f16_round(v[e + 1]);
    hi[e >> 1] = P::pack_f16(ph, qh);
    lo[e >> 1] = P::pack_f16(v[e] - ph, v[e + 1] - qh);
  }
  *reinterpret_cast<uint4*>(hi_row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(lo_row) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

template <class P>
MRF_FN void att_body(const Args& a, typename P::Ctx& cx, uint8_t* smem, Barriers<typename P::Mbar>& bar, uint32_t* tmem_base_s,
                     const typename P::TensorMap* tmq = nullptr, const typename P::TensorMap* tmk = nullptr) {
  P::pdl_launch();
  const int tid = cx.tid(), lane = tid & 31, warp = P::bcast0(cx, tid >> 5);
  const int dk = a.dk, G = dk / 8;
  const int unit = cx.block();
  const int qt = unit % a.q_tiles, h = (unit / a.q_tiles) % a.n_heads, b = unit / (a.q_tiles * a.n_heads);
  const int T = a.len[b];
  const int q0 = qt * A_QT;
  float* raw = reinterpret_cast<float*>(smem + off_raw(dk));
  uint8_t* Qop = smem + off_q(dk);
  uint8_t* Kop = smem + off_k(dk);
  uint8_t* Vop = smem + off_v(dk);
  uint8_t* Pop = smem + off_p(dk);
  uint8_t* Eko = smem + off_ek(dk);
  float* Ev = reinterpret_cast<float*>(smem + off_ev(dk));
  const int q_part = G * A_QT * 16, p_part = (A_KB / 8) * A_QT * 16;
  const int n_blk = (T + A_KB - 1) / A_KB;
  // uniform per CTA: a tile past the utterance has nothing to do; neither has a short last tile that the launcher gave to
  // the CUDA-core kernel (a 259-id utterance is two full tiles + 3 rows, and a 3-row tile costs a CTA as much as a full one:
  // 32 utterances x 2 heads x 3 tiles = 192 CTAs on 148 SMs were two waves)
  const bool active = q0 < T && !(a.tail_thr > 0 && q0 > 0 && T - q0 <= a.tail_thr);

  if (warp == 1) P::tmem_alloc(cx, tmem_base_s, 512u);
  if (tid == 0) {
    P::mbar_init(cx, &bar.raw_full, 1); P::mbar_init(cx, &bar.raw_free, A_CONV_THREADS);
    P::mbar_init(cx, &bar.q_full, A_CONV_THREADS); P::mbar_init(cx, &bar.op_full, A_CONV_THREADS);
    P::mbar_init(cx, &bar.s_full, 1); P::mbar_init(cx, &bar.s_free, A_SM_THREADS);
    P::mbar_init(cx, &bar.p_full, A_SM_THREADS); P::mbar_init(cx, &bar.kv_free, 1); P::mbar_init(cx, &bar.rl_full, 1);
    P::fence_mbar_init();
  }
  P::fence_tc_before();
  P::syncthreads(cx);
  P::fence_tc_after();
  const uint32_t tmem_d = *tmem_base_s;
  P::pdl_sync();                                       // programmatic dependent launch: the prologue above overlapped the previous grid
  const float* base = a.qkv.p + (long long)b * a.qkv.bs;
  const float* qg = base + (long long)(h * dk) * a.qkv.cs;
  const float* kg = base + (long long)(a.H + h * dk) * a.qkv.cs;
  const float* vg = base + (long long)(2 * a.H + h * dk) * a.qkv.cs;
  constexpr int RSQ = A_QT + 8, RSK = A_KB + 8;       // raw row strides (floats)

  if (active && warp == 0 && a.tm) {
    // ---------------------------------------------------------------------- TMA by tensor map: one box per window
    // (the per-row form below issues dk .. 2 dk bulk copies of 256 bytes per block from one lane: ~3 us of a ~7 us block step)
    const uint32_t rawa = P::saddr(cx, raw);
    const int chq = h * dk, chk = a.H + h * dk, chv = 2 * a.H + h * dk;
    auto box = [&](uint32_t dst, const typename P::TensorMap* tm, int x, int ch) {
      if (a.flat) P::tma_load_3d(cx, dst, tm, x, b, ch, &bar.raw_full);
      else P::tma_load_3d(cx, dst, tm, x, ch, b, &bar.raw_full);
    };
    uint32_t it = 0;
    if (P::elect_one(cx)) {
      P::mbar_expect_tx(cx, &bar.raw_full, (uint32_t)(dk * RSQ * 4));
      box(rawa, tmq, q0, chq);
    }
    P::syncwarp();
    ++it;
    for (int pass = 0; pass < 2; ++pass)
      for (int blk = 0; blk < n_blk; ++blk, ++it) {
        const int j0 = blk * A_KB;
        P::mbar_wait(cx, &bar.raw_free, (it - 1) & 1);
        if (P::elect_one(cx)) {
          P::mbar_expect_tx(cx, &bar.raw_full, (uint32_t)((pass ? 2 : 1) * dk * RSK * 4));
          box(rawa, tmk, j0, chk);
          if (pass) box(rawa + (uint32_t)(dk * RSK * 4), tmk, j0, chv);
        }
        P::syncwarp();
      }
  } else if (active && warp == 0) {
    // ---------------------------------------------------------------------- TMA: the q tile, then k (and v) blocks
    uint32_t it = 0;                                   // raw buffer uses so far
    {
      const int g1 = imin((imin(q0 + A_QT, T) + 3) & ~3, a.qkv.cs);
      const uint32_t row_bytes = (uint32_t)(g1 - q0) * 4;
      if (P::elect_one(cx)) P::mbar_expect_tx(cx, &bar.raw_full, row_bytes * (uint32_t)dk);
      uint32_t d = P::saddr(cx, raw);
      const float* src = qg + q0;
      for (int c = 0; c < dk; ++c, d += RSQ * 4, src += a.qkv.cs)
        if (P::elect_one(cx)) P::bulk_g2s(cx, d, src, row_bytes, &bar.raw_full);
      P::syncwarp();
      ++it;
    }
    for (int pass = 0; pass < 2; ++pass)
      for (int blk = 0; blk < n_blk; ++blk, ++it) {
        const int j0 = blk * A_KB;
        const int g1 = imin((imin(j0 + A_KB, T) + 3) & ~3, a.qkv.cs);
        const uint32_t row_bytes = (uint32_t)(g1 - j0) * 4;
        const int rows = pass ? 2 * dk : dk;           // pass 1 needs the keys only
        P::mbar_wait(cx, &bar.raw_free, (it - 1) & 1);
        if (P::elect_one(cx)) P::mbar_expect_tx(cx, &bar.raw_full, row_bytes * (uint32_t)rows);
        uint32_t d = P::saddr(cx, raw);
        const float* src = kg + j0;
        for (int c = 0; c < dk; ++c, d += RSK * 4, src += a.qkv.cs)
          if (P::elect_one(cx)) P::bulk_g2s(cx, d, src, row_bytes, &bar.raw_full);
        if (pass) {
          src = vg + j0;
          for (int c = 0; c < dk; ++c, d += RSK * 4, src += a.qkv.cs)
            if (P::elect_one(cx)) P::bulk_g2s(cx, d, src, row_bytes, &bar.raw_full);
        }
        P::syncwarp();
      }
  } else if (active && warp == 1) {
    // ---------------------------------------------------------------------- MMA issue
    const uint32_t td = (uint32_t)P::bcast0(cx, (int)tmem_d);
    const uint32_t q_lbo = A_QT * 16, k_lbo = 2 * A_KB * 16, v_lbo = 2u * (uint32_t)dk * 16, p_lbo = A_QT * 16, e_lbo = 32 * 16;
    const uint32_t qh0 = desc_lo(P::saddr(cx, Qop), q_lbo), ql0 = desc_lo(P::saddr(cx, Qop + q_part), q_lbo);
    const int ks_q = dk / 16;                          // k-steps over the head dimension
    // ---- relative-key logits Rl = Q Ek^T (N = 16: 9 used)
    P::mbar_wait(cx, &bar.q_full, 0);
    P::fence_tc_after();
    {
      uint32_t ah = qh0, al = ql0, eb = desc_lo(P::saddr(cx, Eko), e_lbo);
      for (int kb = 0; kb < ks_q; ++kb) {
        if (P::elect_one(cx)) {
          P::mma_f16(cx, td + A_TM_RL, ah, eb, idesc_f16(32), kb ? 1u : 0u);
          P::mma_f16(cx, td + A_TM_RL + 16u, al, eb, idesc_f16(16), 1u);
        }
        P::syncwarp();
        ah += 2u * (q_lbo >> 4); al += 2u * (q_lbo >> 4); eb += 2u * (e_lbo >> 4);
      }
      if (P::elect_one(cx)) P::mma_commit(cx, &bar.rl_full);
      P::syncwarp();
    }
    uint32_t it = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int blk = 0; blk < n_blk; ++blk, ++it) {
        P::mbar_wait(cx, &bar.op_full, it & 1);
        if (it >= 1) P::mbar_wait(cx, &bar.s_free, (it - 1) & 1);
        P::fence_tc_after();
        uint32_t ah = qh0, al = ql0, kb_d = desc_lo(P::saddr(cx, Kop), k_lbo);
        for (int kb = 0; kb < ks_q; ++kb) {
          if (P::elect_one(cx)) {
            if (pass == 0) {
              P::mma_f16(cx, td + A_TM_S, ah, kb_d, idesc_f16(A_KB), kb ? 1u : 0u);          // hi x hi only
            } else {
              P::mma_f16(cx, td + A_TM_S, ah, kb_d, idesc_f16(2 * A_KB), kb ? 1u : 0u);      // main | hi x lo
              P::mma_f16(cx, td + A_TM_S + A_KB, al, kb_d, idesc_f16(A_KB), 1u);             //        lo x hi
            }
          }
          P::syncwarp();
          ah += 2u * (q_lbo >> 4); al += 2u * (q_lbo >> 4); kb_d += 2u * (k_lbo >> 4);
        }
        if (P::elect_one(cx)) P::mma_commit(cx, &bar.s_full);
        P::syncwarp();
        if (pass == 1) {
          P::mbar_wait(cx, &bar.p_full, blk & 1);
          P::fence_tc_after();
          uint32_t ph = desc_lo(P::saddr(cx, Pop), p_lbo), pl = desc_lo(P::saddr(cx, Pop + p_part), p_lbo);
          uint32_t vb = desc_lo(P::saddr(cx, Vop), v_lbo);
          for (int kb = 0; kb < A_KB / 16; ++kb) {
            const uint32_t acc = (blk > 0 || kb > 0) ? 1u : 0u;
            if (P::elect_one(cx)) {
              P::mma_f16(cx, td + A_TM_O, ph, vb, idesc_f16(2 * dk), acc);                  // main | hi x lo
              P::mma_f16(cx, td + A_TM_O + (uint32_t)dk, pl, vb, idesc_f16(dk), 1u);        //        lo x hi
            }
            P::syncwarp();
            ph += 2u * (p_lbo >> 4); pl += 2u * (p_lbo >> 4); vb += 2u * (v_lbo >> 4);
          }
        }
        if (P::elect_one(cx)) P::mma_commit(cx, &bar.kv_free);      // K (and V, P) of this block have been consumed
        P::syncwarp();
      }
  } else if (active && warp >= A_CONV_WARP0 && warp < A_SM_WARP0) {
    // ---------------------------------------------------------------------- converters
    const int ctid = tid - A_CONV_WARP0 * 32;
    // relative-position tables (no dependency on the TMA): Ek as a stacked B operand (rows 9..15 zero), Ev in fp32
    for (int idx = ctid; idx < G * 16; idx += A_CONV_THREADS) {
      const int g = idx / 16, r = idx - g * 16;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = r < A_NREL ? P::ldg(a.rel_k + r * dk + g * 8 + e) : 0.f;
      split8<P>(Eko + (g * 32 + r) * 16, Eko + (g * 32 + 16 + r) * 16, v);
    }
    for (int idx = ctid; idx < A_NREL * dk; idx += A_CONV_THREADS) Ev[idx] = P::ldg(a.rel_v + idx);
    // the query tile, scaled by 1 / sqrt(dk) before the product as the reference does (attentions.py:232)
    uint32_t it = 0;
    P::mbar_wait(cx, &bar.raw_full, 0);
    for (int idx = ctid; idx < G * A_QT; idx += A_CONV_THREADS) {
      const int g = idx / A_QT, r = idx - g * A_QT;
      const bool live = q0 + r < T;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = live ? raw[(g * 8 + e) * RSQ + r] / sqrtf((float)dk) : 0.f;
      split8<P>(Qop + (g * A_QT + r) * 16, Qop + q_part + (g * A_QT + r) * 16, v);
    }
    P::fence_async_proxy();
    P::mbar_arrive(cx, &bar.q_full);
    P::mbar_arrive(cx, &bar.raw_free);
    ++it;
    for (int pass = 0; pass < 2; ++pass)
      for (int blk = 0; blk < n_blk; ++blk, ++it) {
        const int j0 = blk * A_KB;
        P::mbar_wait(cx, &bar.raw_full, it & 1);
        if (it >= 2) P::mbar_wait(cx, &bar.kv_free, it & 1);          // block it-2 (the previous one) fully consumed
        // K as the B operand of S = Q K^T: rows = keys (hi 0..63 | lo 64..127), 16-byte row = 8 consecutive d
        for (int idx = ctid; idx < G * A_KB; idx += A_CONV_THREADS) {
          const int g = idx / A_KB, kq = idx - g * A_KB;
          const bool live = j0 + kq < T;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = live ? raw[(g * 8 + e) * RSK + kq] : 0.f;
          split8<P>(Kop + (g * 2 * A_KB + kq) * 16, Kop + (g * 2 * A_KB + A_KB + kq) * 16, v);
        }
        if (pass) {
          // V as the B operand of O = P V: rows = d (hi 0..dk-1 | lo dk..2dk-1), 16-byte row = 8 consecutive keys
          for (int idx = ctid; idx < (A_KB / 8) * dk; idx += A_CONV_THREADS) {
            const int kgp = idx / dk, d = idx - kgp * dk;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (j0 + kgp * 8 + e < T) ? raw[(dk + d) * RSK + kgp * 8 + e] : 0.f;
            split8<P>(Vop + (kgp * 2 * dk + d) * 16, Vop + (kgp * 2 * dk + dk + d) * 16, v);
          }
        }
        P::fence_async_proxy();
        P::mbar_arrive(cx, &bar.op_full);
        P::mbar_arrive(cx, &bar.raw_free);
      }
  } else if (active && warp >= A_SM_WARP0) {
    // ---------------------------------------------------------------------- softmax / epilogue: one query row per thread
    const int q = warp & 3;                            // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;
    const int i = q0 + r;                              // query position
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float rl[A_NREL];
    {
      P::mbar_wait(cx, &bar.rl_full, 0);
      P::fence_tc_after();
      float m16[16], c16[16];
      P::tmem_ld16(cx, tmem_d + lane_addr + A_TM_RL, m16);
      P::tmem_ld16(cx, tmem_d + lane_addr + A_TM_RL + 16u, c16);
#pragma unroll
      for (int k = 0; k < A_NREL; ++k) rl[k] = m16[k] + c16[k];
    }
    float m_run = -INFINITY, l_run = 0.f;
    float u[A_NREL];
#pragma unroll
    for (int k = 0; k < A_NREL; ++k) u[k] = 0.f;
    uint32_t it = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int blk = 0; blk < n_blk; ++blk, ++it) {
        const int j0 = blk * A_KB;
        P::mbar_wait(cx, &bar.s_full, it & 1);
        P::fence_tc_after();
        if (pass == 1 && blk >= 1) P::mbar_wait(cx, &bar.kv_free, (it - 1) & 1);   // GEMM2 of the previous block has read P
#pragma unroll 1
        for (int c = 0; c < A_KB / 16; ++c) {
          float s[16];
          P::tmem_ld16(cx, tmem_d + lane_addr + A_TM_S + (uint32_t)(c * 16), s);
          if (pass) {
            float cr[16];
            P::tmem_ld16(cx, tmem_d + lane_addr + A_TM_S + A_KB + (uint32_t)(c * 16), cr);
#pragma unroll
            for (int e = 0; e < 16; ++e) s[e] += cr[e];
          }
          float p[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int j = j0 + c * 16 + e;
            const int rel = j - i + 4;
            float sc = s[e];
            if (rel >= 0 && rel < A_NREL) {
              float add = 0.f;
#pragma unroll
              for (int k = 0; k < A_NREL; ++k) add = rel == k ? rl[k] : add;
              sc += add;
            }
            if (pass == 0) {
              if (j < T) m_run = fmaxf(m_run, sc);
              p[e] = 0.f;
            } else {
              const float pv = j < T ? expf(sc - m_run) : 0.f;      // keys past the utterance: masked_fill(-1e4) -> exactly 0
              l_run += pv;
              p[e] = pv;
              if (rel >= 0 && rel < A_NREL) {
#pragma unroll
                for (int k = 0; k < A_NREL; ++k) u[k] = rel == k ? pv : u[k];
              }
            }
          }
          if (pass) {
            // P as the A operand of O = P V: 16-byte row = 8 consecutive keys of this query
            split8<P>(Pop + ((2 * c) * A_QT + r) * 16, Pop + p_part + ((2 * c) * A_QT + r) * 16, p);
            split8<P>(Pop + ((2 * c + 1) * A_QT + r) * 16, Pop + p_part + ((2 * c + 1) * A_QT + r) * 16, p + 8);
          }
        }
        P::fence_tc_before();
        P::mbar_arrive(cx, &bar.s_free);
        if (pass) {
          P::fence_async_proxy();
          P::mbar_arrive(cx, &bar.p_full);
        }
      }
    // ---- O = (main + correction) + banded relative values, normalised, stored channel-major
    P::mbar_wait(cx, &bar.kv_free, (it - 1) & 1);
    P::fence_tc_after();
    {
      // tcgen05.ld is .sync.aligned: the WHOLE warp must execute it, also the lanes whose query lies past the utterance
      // (a ragged tile end) - only the stores are predicated.  (The first hardware run hung here: the loads sat inside
      // `if (i < T)`; the CPU model executes lanes independently and cannot see a partial-warp collective.)
      const float inv_l = 1.f / l_run;
      float* ob = a.out.p + (long long)b * a.out.bs + (long long)(h * dk) * a.out.cs + i;
#pragma unroll 1
      for (int c = 0; c < dk / 16; ++c) {
        float o[16], cr[16];
        P::tmem_ld16(cx, tmem_d + lane_addr + A_TM_O + (uint32_t)(c * 16), o);
        P::tmem_ld16(cx, tmem_d + lane_addr + A_TM_O + (uint32_t)(dk + c * 16), cr);
        if (i < T) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float acc = o[e] + cr[e];
#pragma unroll
            for (int k = 0; k < A_NREL; ++k) acc = fmaf(u[k], Ev[k * dk + c * 16 + e], acc);
            ob[(long long)(c * 16 + e) * a.out.cs] = acc * inv_l;
          }
        }
      }
    }
  }
  P::fence_tc_before();
  P::syncthreads(cx);
  if (warp == 1) P::tmem_dealloc(cx, tmem_d, 512u);
}

}  // namespace att
}  // namespace pb200
