// Second-generation persistent tensor-core convolution (the default conv kernel since round 2), written against
// the primitive policy P (tc_policy_dev.cuh on the GPU, tests/sim/sim_prim.h on the CPU) so that its logic is checked by
// tests/test_conv2_sim.py without a GPU.  Same GEMM view, data layout, ragged-batch rules and
// fused epilogues as conv_mma_persist_kernel (conv_mma.cu); what changes, from the round-1 measurements:
//
//   * both TMA warps run converged and issue from an elected lane, so copy operands live in uniform registers (the
//     shipped kernel spends ~150 cycles per activation-row copy building addresses in vector registers)
//   * the two weight halves are stacked along N, [W_hi ; W_lo], so a k-step is TWO instructions instead of three:
//         D[:, 0..N)   += A_hi W_hi^T   and   D[:, N..2N) += A_hi W_lo^T      one tcgen05.mma, N' = 2N
//         D[:, N..2N)  += A_lo W_hi^T                                          one tcgen05.mma, N' = N
//     (the SS-form instruction is paced by reading its operands from shared memory; A is read twice instead of three
//     times).  Each K-chain therefore owns a (main | correction) accumulator pair; the epilogue adds all of them in
//     fp32 round-to-nearest, which keeps the chained / separated accumulation that tf32x3 layers need.
//   * a weight unit is ONE bulk copy (the stacked layout is contiguous per channel chunk and tap).
#pragma once

namespace pb200 {
namespace conv2 {

constexpr int C2_CONV_WARP0 = 3, C2_CONV_THREADS = 128;
constexpr int C2_EPI_WARP0 = 7, C2_EPI_THREADS = 256;
constexpr int C2_THREADS = 32 * 15;
constexpr int C2_RAW_SLOTS = 2, C2_A_SLOTS_MAX = 12, C2_W_SLOTS = 4, C2_T_SLOTS = 2;   // (operand slots per launch: MmaConvArgs::a_slots, 2 .. 12)

template <class Mbar>
struct Barriers {
  Mbar raw_full[C2_RAW_SLOTS], raw_empty[C2_RAW_SLOTS], a_full[C2_A_SLOTS_MAX], a_empty[C2_A_SLOTS_MAX], w_full[C2_W_SLOTS],
      w_empty[C2_W_SLOTS], t_full[C2_T_SLOTS], t_empty[C2_T_SLOTS];
};

MRF_FN uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
// Operand precision of a layer.  All three are error-compensated three-product schemes (x = hi + lo, w = hi + lo):
//   PREC_BF16  kind::f16, BF16 operands: 16 mantissa bits kept                 (the shipped generator path)
//   PREC_TF32  kind::tf32: 21 bits kept, but only K = 8 per instruction        (the shipped encoder / flow path)
//   PREC_F16   kind::f16, FP16 operands: 22 bits kept at K = 16 per instruction - tf32x3-class accuracy at bf16x3 cost,
//              half the operand bytes in shared memory and L2 (tools/precision_study.py part 3; operands must stay below
//              65504 in magnitude, which every activation and weight of the graph does by orders of magnitude)
enum : int { PREC_BF16 = 0, PREC_TF32 = 1, PREC_F16 = 2 };
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A / B format at bits 7-9 / 10-12 (0 = F16, 1 = BF16, 2 = TF32)
MRF_FN constexpr uint32_t make_idesc(int prec, int M, int N) {
  const uint32_t fmt = prec == PREC_TF32 ? 2u : prec == PREC_BF16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
MRF_FN int imax(int a, int b) { return a > b ? a : b; }
MRF_FN int imin(int a, int b) { return a < b ? a : b; }

// WaveNet gate tanh(a) * sigmoid(b) (commons.py:99-106), one out-of-line copy on the GPU (instruction footprint)
MRF_NOINLINE float wn_gate(float a, float b) { return tanhf(a) * (1.f / (1.f + expf(-b))); }

template <int UP>
MRF_FN void store_upsampled(const float (&v)[16], float* yb, int cs, int row0, int t, int up_pad, int Lout) {
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

// TM: the activation rows of a channel chunk arrive as one tensor-map TMA copy (cp.async.bulk.tensor.3d, box
// [KC channels][R / boxes columns], any start column, out-of-bounds zero filled) instead of KC per-row bulk copies.
// ASLOTS: operand-ring slots, 2 = the streaming ring (compile-time), 0 = one slot per channel chunk of a position
// (MmaConvArgs::a_slots, A-stationary launches)
template <class P, int PREC, int MT, bool TM = false, int ASLOTS = 2>
MRF_FN void conv2_body(const MmaConvArgs& a, typename P::Ctx& cx, uint8_t* smem, Barriers<typename P::Mbar>& bar,
                       uint32_t* tmem_base_s, const typename P::TensorMap* tmx = nullptr) {
  constexpr bool TF32 = PREC == PREC_TF32;
  constexpr int ES = TF32 ? 4 : 2;
  constexpr int E = 16 / ES;
  constexpr int KSTEP = 2 * E;
  constexpr int MH = MT / 128;

  P::pdl_launch();
  const int tid = cx.tid(), lane = tid & 31, warp = P::bcast0(cx, tid >> 5);
  const int block = cx.block(), grid = cx.grid();
  const int KC = a.kc, R = a.stage_rows, RS = a.raw_stride, NT = a.n_tile;
  const int raw_bytes = KC * RS * 4, a_part = KC * R * ES;
  const uint32_t unit_bytes = (uint32_t)(KC / E) * 2u * (uint32_t)NT * 16u;       // one (chunk, tap): stacked hi | lo rows
  uint8_t* RAW_ring = smem;
  uint8_t* A_ring = RAW_ring + size_t(C2_RAW_SLOTS) * raw_bytes;
  const int a_slots = ASLOTS > 0 ? ASLOTS : (a.a_slots < 2 ? 2 : (a.a_slots > C2_A_SLOTS_MAX ? C2_A_SLOTS_MAX : a.a_slots));   // operand ring: 2, or every chunk of a position (A-stationary)
  uint8_t* W_ring = A_ring + size_t(a_slots) * 2 * a_part;
  const int n_kc = a.ci / KC;
  const int n_units = n_kc * a.k;
  const int tpi = a.tiles_per_item, total = a.total_tiles;
  const int t_slots = a.t_slots;                       // 1 or 2 TMEM accumulator sets
  const int pair_cols = 2 * NT;                        // (main | correction) of one chain
  const int set_cols = MH * a.mh_stride;               // mh_stride = chains * pair_cols

  if (warp == 2) P::tmem_alloc(cx, tmem_base_s, (uint32_t)a.tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < C2_RAW_SLOTS; ++i) { P::mbar_init(cx, &bar.raw_full[i], 1); P::mbar_init(cx, &bar.raw_empty[i], C2_CONV_THREADS); }
    for (int i = 0; i < a_slots; ++i) { P::mbar_init(cx, &bar.a_full[i], C2_CONV_THREADS); P::mbar_init(cx, &bar.a_empty[i], 1); }
    for (int i = 0; i < C2_W_SLOTS; ++i) { P::mbar_init(cx, &bar.w_full[i], 1); P::mbar_init(cx, &bar.w_empty[i], 1); }
    for (int i = 0; i < C2_T_SLOTS; ++i) { P::mbar_init(cx, &bar.t_full[i], 1); P::mbar_init(cx, &bar.t_empty[i], C2_EPI_THREADS); }
    P::fence_mbar_init();
  }
  P::fence_tc_before();
  P::syncthreads(cx);
  P::fence_tc_after();
  const uint32_t tmem_d = *tmem_base_s;
  P::pdl_sync();                                       // programmatic dependent launch: the prologue above overlapped the previous grid

  // tile id -> (output-row tile, item, time block); every role walks the same list and skips the same tiles
  const int flat_tg = a.flat_tg;                       // > 0: tiles on the concatenated time axis (see MmaConvArgs)
  // Tile walk.  Default: tile = (nt, item, time block) with the time block fastest, CTAs take tiles round-robin.
  // A-stationary (a.astat): tile = (item, time block, nt) with nt fastest and every CTA takes a CONTIGUOUS range, so the
  // output-row tiles of one position follow each other on one CTA: the activation roles (TMA, converters) work once per
  // group of such tiles, the MMA warp waits for the operand once per group and releases it after the group's last tile.
  const int astat = a.astat, n_nt = a.n_tiles;
  const int per_cta = (total + grid - 1) / grid;
  const int tile_lo = astat ? imin(block * per_cta, total) : block;
  const int tile_hi = astat ? imin(tile_lo + per_cta, total) : total;
  const int tile_step = astat ? 1 : grid;
  auto group_first = [&](int tile) { return !astat || tile == tile_lo || tile % n_nt == 0; };
  auto group_last = [&](int tile) { return !astat || tile + 1 == tile_hi || tile % n_nt == n_nt - 1; };
  auto decode = [&](int tile, int& nt, int& b, int& t0, int& L, int& Lq) {
    int tb;
    if (astat) {
      nt = tile % n_nt;
      const int pos = tile / n_nt;
      tb = pos % tpi;
      b = pos / tpi;
    } else {
      tb = tile % tpi;
      const int rest = tile / tpi;
      b = rest % a.batch;
      nt = rest / a.batch;
    }
    t0 = tb * MT;
    L = flat_tg ? a.flat_n * flat_tg : a.len[b] * a.len_scale;
    Lq = L + a.q_extra;
    return t0 < Lq;
  };
  // flat mode: is concatenated position g inside an utterance, and whose?
  // (g / flat_tg by multiplication with ceil(2^32 / flat_tg): exact while g * flat_tg < 2^32, which the host guarantees for
  // a flat launch; a 32-bit integer division is ~25 dependent instructions, and this runs per row in three roles)
  const uint32_t tg_magic = flat_tg > 1 ? 0xFFFFFFFFu / (uint32_t)flat_tg + 1u : 0u;
  auto flat_live = [&](int g, int& item) -> bool {
    if (g < 0) return false;
    item = (int)(((unsigned long long)(uint32_t)g * tg_magic) >> 32);
    return item < a.flat_n && g - item * flat_tg < a.len[item] * a.len_scale;
  };

  if (TM && warp == 0) {
    // ---------------------------------------------------------------------- raw activation window via tensor-map TMA
    if (P::elect_one(cx)) P::tma_prefetch_desc(tmx);
    P::syncwarp();
    const int n_box = a.tm_boxes, Wb = RS / n_box;
    const uint32_t box_bytes = (uint32_t)KC * (uint32_t)Wb * 4u;
    uint32_t it = 0;
    for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
      int nt, b, t0, L, Lq;
      bool ok = decode(tile, nt, b, t0, L, Lq);
      ok = P::bcast0(cx, (int)ok) != 0;
      if (!ok || !group_first(tile)) continue;                         // one load per position
      const int t_base = (t0 - a.pad) & ~3;                            // 16-byte aligned start (may be negative: zero fill)
      for (int kc = 0; kc < n_kc; ++kc, ++it) {
        const int s = it % C2_RAW_SLOTS;
        if (it >= C2_RAW_SLOTS) P::mbar_wait(cx, &bar.raw_empty[s], ((it / C2_RAW_SLOTS) - 1) & 1);
        if (P::elect_one(cx)) {
          P::mbar_expect_tx(cx, &bar.raw_full[s], box_bytes * (uint32_t)n_box);
          const uint32_t d = P::saddr(cx, RAW_ring + size_t(s) * raw_bytes);
          P::tma_load_3d(cx, d, tmx, t_base, kc * KC, b, &bar.raw_full[s]);
          if (n_box == 2) P::tma_load_3d(cx, d + box_bytes, tmx, t_base + Wb, kc * KC, b, &bar.raw_full[s]);
        }
        P::syncwarp();
      }
    }
  } else if (warp == 0) {
    // ---------------------------------------------------------------------- raw activation rows via TMA, uniform issue
    uint32_t it = 0;
    for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
      int nt, b, t0, L, Lq;
      bool ok = decode(tile, nt, b, t0, L, Lq);
      L = P::bcast0(cx, L);                                            // loaded from global: make uniformity explicit
      ok = P::bcast0(cx, (int)ok) != 0;
      if (!ok || !group_first(tile)) continue;
      const int t_lo = t0 - a.pad;
      const int t_base = t_lo & ~3;                                    // smem column 0 <-> time t_base
      const int g0 = imax(t_lo, 0) & ~3;                               // first / one-past-last float fetched
      const int g1 = imin((imin(t_lo + R, L) + 3) & ~3, a.x.cs);
      const bool any = g1 > g0;                                        // a tile that lies wholly in the ConvTranspose tail
      const uint32_t row_bytes = any ? (uint32_t)(g1 - g0) * 4 : 0u;   // reads nothing: every row is masked to zero
      const float* xb = a.x.p + (long long)b * a.x.bs + g0;
      for (int kc = 0; kc < n_kc; ++kc, ++it) {
        const int s = it % C2_RAW_SLOTS;
        if (it >= C2_RAW_SLOTS) P::mbar_wait(cx, &bar.raw_empty[s], ((it / C2_RAW_SLOTS) - 1) & 1);
        if (P::elect_one(cx)) P::mbar_expect_tx(cx, &bar.raw_full[s], row_bytes * (uint32_t)KC);
        uint32_t d = P::saddr(cx, RAW_ring + size_t(s) * raw_bytes) + (uint32_t)(g0 - t_base) * 4;
        const float* src = xb + (long long)(kc * KC) * a.x.cs;
        const uint32_t d_step = (uint32_t)RS * 4u;
        const long long s_step = a.x.cs;
        if (any) {
#pragma unroll 4
          for (int c = 0; c < KC; ++c, d += d_step, src += s_step) {
            if (P::elect_one(cx)) P::bulk_g2s(cx, d, src, row_bytes, &bar.raw_full[s]);
          }
        }
        P::syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------- weight units via TMA: one copy per unit
    uint32_t it = 0;
    const size_t nt_bytes = size_t(a.k) * (a.ci / E) * 2 * NT * 16;    // all taps and chunks of one output-row tile
    for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
      int nt, b, t0, L, Lq;
      bool ok = decode(tile, nt, b, t0, L, Lq);
      ok = P::bcast0(cx, (int)ok) != 0;
      if (!ok) continue;
      const uint8_t* wsrc = a.w + size_t(nt) * nt_bytes;
      int kc = 0, j = 0;                                               // unit u = (chunk kc, tap j), tap fastest
      for (int u = 0; u < n_units; ++u, ++it, ++j) {
        const int s = it % C2_W_SLOTS;
        if (it >= C2_W_SLOTS) P::mbar_wait(cx, &bar.w_empty[s], ((it / C2_W_SLOTS) - 1) & 1);
        if (j == a.k) { j = 0; ++kc; }
        const uint8_t* src = wsrc + (size_t(j) * (a.ci / E) + size_t(kc) * (KC / E)) * 2 * NT * 16;
        if (P::elect_one(cx)) {
          P::mbar_expect_tx(cx, &bar.w_full[s], unit_bytes);
          P::bulk_g2s(cx, P::saddr(cx, W_ring + size_t(s) * unit_bytes), src, unit_bytes, &bar.w_full[s]);
        }
        P::syncwarp();
      }
    }
  } else if (warp == 2) {
    // ---------------------------------------------------------------------- MMA issue: whole warp converged, the
    // tcgen05 instructions predicated on one elected lane so that every operand stays warp-uniform
    const uint32_t tmem_du = (uint32_t)P::bcast0(cx, (int)tmem_d);
    const uint32_t a_lbo = (uint32_t)R * 16, w_lbo = 2u * (uint32_t)NT * 16;
    const uint32_t idesc2 = make_idesc(PREC, 128, 2 * NT), idesc1 = make_idesc(PREC, 128, NT);
    const uint32_t a_step = 2u * (uint32_t)R, w_step = 4u * (uint32_t)NT;          // 16-byte units per k-step
    uint32_t a_slot = 0, a_round = 0, w_it = 0, t_it = 0;
    // optional per-role wait counters (a.prof, developer diagnostic): cycles the MMA warp waits for [0] the TMEM set,
    // [1] an operand chunk, [2] a weight unit, and [3] its whole loop; [4] epilogue waiting for an accumulator, [5] its loop;
    // [6] converters waiting for raw data, [7] for a free operand slot, [8] their loop
    const bool prof = a.prof != nullptr;
    long long pw_t = 0, pw_a = 0, pw_w = 0;
    const long long p_start = prof ? P::clock() : 0;
    for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
      int nt, b, t0, L, Lq;
      bool ok = decode(tile, nt, b, t0, L, Lq);
      Lq = P::bcast0(cx, Lq);
      ok = P::bcast0(cx, (int)ok) != 0;
      if (!ok) continue;
      const int mh_live = (Lq - t0 > 128 && MH > 1) ? 2 : 1;
      const int ts = t_it % t_slots;
      long long pc = prof ? P::clock() : 0;
      if (t_it >= (uint32_t)t_slots) P::mbar_wait(cx, &bar.t_empty[ts], ((t_it / t_slots) - 1) & 1);
      if (prof) pw_t += P::clock() - pc;
      P::fence_tc_after();
      const uint32_t d_set = tmem_du + (uint32_t)(ts * set_cols);
      uint32_t started = 0;
      int u = 0;
      const bool g_first = group_first(tile), g_last = group_last(tile);
      uint32_t as = a_slot, a_par = a_round;                           // operand-ring slot / round of this group's first chunk
      for (int kc = 0; kc < n_kc; ++kc) {
        pc = prof ? P::clock() : 0;
        if (g_first) P::mbar_wait(cx, &bar.a_full[as], a_par & 1);     // (a later tile of the group: still resident)
        if (prof) pw_a += P::clock() - pc;
        P::fence_tc_after();
        const uint32_t a_hi = P::saddr(cx, A_ring + size_t(as) * 2 * a_part);
        const uint32_t ah_base = desc_lo(a_hi, a_lbo), al_base = desc_lo(a_hi + a_part, a_lbo);
        for (int j = 0; j < a.k; ++j, ++u, ++w_it) {
          const int ws = w_it % C2_W_SLOTS;
          pc = prof ? P::clock() : 0;
          P::mbar_wait(cx, &bar.w_full[ws], (w_it / C2_W_SLOTS) & 1);
          if (prof) pw_w += P::clock() - pc;
          P::fence_tc_after();
          const int chain = u * a.chains >= n_units ? 1 : 0;            // = (u * chains) / n_units for chains <= 2, without the division
          uint32_t acc = (started >> chain) & 1u;
          started |= 1u << chain;
          const uint32_t d_pair = d_set + (uint32_t)(chain * pair_cols);
          const uint32_t row = (uint32_t)(j * a.dil);
          uint32_t ah = ah_base + row, al = al_base + row;
          uint32_t wb = desc_lo(P::saddr(cx, W_ring + size_t(ws) * unit_bytes), w_lbo);
#pragma unroll 1
          for (int kb = 0; kb < KC / KSTEP; ++kb) {
            if (a.mma3) {
              // three instructions, every accumulator region addressed with one fixed (base, N): main = hi*hi,
              // correction = hi*lo + lo*hi.  (The stacked form below writes [0, 2N) and then accumulates into [N, 2N):
              // two PARTIALLY overlapping destinations in flight, which raced on hardware - see DESIGN.md.)
              for (int mh = 0; mh < (MH == 2 && mh_live == 2 ? 2 : 1); ++mh) {
                const uint32_t dm = d_pair + (uint32_t)(mh * a.mh_stride), ro = (uint32_t)mh * 128u;
                if (P::elect_one(cx)) {
                  if (TF32) {
                    P::mma_tf32(cx, dm, ah + ro, wb, idesc1, acc);
                    P::mma_tf32(cx, dm + (uint32_t)NT, ah + ro, wb + (uint32_t)NT, idesc1, acc);
                    P::mma_tf32(cx, dm + (uint32_t)NT, al + ro, wb, idesc1, 1u);
                  } else {
                    P::mma_f16(cx, dm, ah + ro, wb, idesc1, acc);
                    P::mma_f16(cx, dm + (uint32_t)NT, ah + ro, wb + (uint32_t)NT, idesc1, acc);
                    P::mma_f16(cx, dm + (uint32_t)NT, al + ro, wb, idesc1, 1u);
                  }
                }
                P::syncwarp();
              }
            } else {
            if (P::elect_one(cx)) {
              if (TF32) {
                P::mma_tf32(cx, d_pair, ah, wb, idesc2, acc);                            // main | hi*lo
                P::mma_tf32(cx, d_pair + (uint32_t)NT, al, wb, idesc1, 1u);              //        lo*hi
              } else {
                P::mma_f16(cx, d_pair, ah, wb, idesc2, acc);
                P::mma_f16(cx, d_pair + (uint32_t)NT, al, wb, idesc1, 1u);
              }
            }
            if (MH == 2 && mh_live == 2) {
              if (P::elect_one(cx)) {
                if (TF32) {
                  P::mma_tf32(cx, d_pair + (uint32_t)a.mh_stride, ah + 128u, wb, idesc2, acc);
                  P::mma_tf32(cx, d_pair + (uint32_t)(a.mh_stride + NT), al + 128u, wb, idesc1, 1u);
                } else {
                  P::mma_f16(cx, d_pair + (uint32_t)a.mh_stride, ah + 128u, wb, idesc2, acc);
                  P::mma_f16(cx, d_pair + (uint32_t)(a.mh_stride + NT), al + 128u, wb, idesc1, 1u);
                }
              }
            }
            }
            P::syncwarp();
            acc = 1u;
            ah += a_step; al += a_step; wb += w_step;
          }
          if (P::elect_one(cx)) P::mma_commit(cx, &bar.w_empty[ws]);
          P::syncwarp();
        }
        if (g_last) {                                                  // the group's last tile releases the operand chunk
          if (P::elect_one(cx)) P::mma_commit(cx, &bar.a_empty[as]);
          P::syncwarp();
        }
        if (++as == (uint32_t)a_slots) { as = 0; ++a_par; }
      }
      if (g_last) { a_slot = as; a_round = a_par; }
      if (P::elect_one(cx)) P::mma_commit(cx, &bar.t_full[ts]);
      P::syncwarp();
      ++t_it;
    }
    if (prof && lane == 0) {
      P::prof_add(a.prof + 0, pw_t); P::prof_add(a.prof + 1, pw_a); P::prof_add(a.prof + 2, pw_w);
      P::prof_add(a.prof + 3, P::clock() - p_start);
    }
  } else if (warp < C2_EPI_WARP0) {
    // ---------------------------------------------------------------------- converters
    const int ctid = tid - C2_CONV_WARP0 * 32;
    uint32_t raw_it = 0, as = 0, a_round = 0;                           // next operand slot and how often the ring has wrapped
    const bool prof = a.prof != nullptr && ctid == 0;
    long long pw_r = 0, pw_e = 0;
    const long long p_start = prof ? P::clock() : 0;
    for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
      int nt, b, t0, L, Lq;
      if (!decode(tile, nt, b, t0, L, Lq) || !group_first(tile)) continue;   // one conversion per position
      const int t_lo = t0 - a.pad;
      const int off = t_lo - (t_lo & ~3);                               // smem column of stage row 0
      const int Wb = TM ? RS / a.tm_boxes : RS;                         // TM: dense boxes [box][KC][Wb]
      for (int kc = 0; kc < n_kc; ++kc, ++raw_it) {
        const int rs = raw_it % C2_RAW_SLOTS;
        long long pc = prof ? P::clock() : 0;
        P::mbar_wait(cx, &bar.raw_full[rs], (raw_it / C2_RAW_SLOTS) & 1);
        if (prof) { const long long n = P::clock(); pw_r += n - pc; pc = n; }
        if (a_round > 0) P::mbar_wait(cx, &bar.a_empty[as], (a_round - 1) & 1);
        if (prof) pw_e += P::clock() - pc;
        const float* raw = reinterpret_cast<const float*>(RAW_ring + size_t(rs) * raw_bytes) + off;
        uint8_t* A_hi = A_ring + size_t(as) * 2 * a_part;
        uint8_t* A_lo = A_hi + a_part;
        // One staged row (time step) per thread, channel groups in the inner loop: the addresses advance by constants, the
        // in / out-of-utterance test is hoisted, and four groups are in flight per thread (the first version walked the
        // groups in the outer loop - ~19 instructions per element, one load latency at a time; on the B200 the four
        // converter warps, not the TMA issue or the tensor pipe, paced every small launch: profiles/r02_*.txt).
        const int G = KC / E;
        // work item = (row, share of the channel groups): with R = 136 rows and 128 threads a row-only split would leave
        // 120 threads idle in the second pass
        const int S = (G % 4 == 0 && G >= 16) ? 4 : (G % 2 == 0 && G >= 8) ? 2 : 1, GS = G / S;
        int sg = 0, r = ctid;                                           // item = sg * R + r, advanced without a division
        for (int item = ctid; item < R * S; item += C2_CONV_THREADS, r += C2_CONV_THREADS) {
          while (r >= R) { r -= R; ++sg; }
          const int t = t_lo + r;
          int item_unused = 0;
          const bool live = flat_tg ? flat_live(t, item_unused) : (t >= 0 && t < L);   // outside the utterance: zeros, whatever the
          const int bx = (TM && off + r >= Wb) ? 1 : 0;                  // (unwritten / stale) smem holds
          const size_t g_src = (size_t)E * Wb, g_dst = (size_t)R * 16;
          const float* src = raw + (size_t)bx * KC * Wb + (r - bx * Wb) + (size_t)(sg * GS) * g_src;   // (raw points at column `off`)
          uint8_t* dh = A_hi + r * 16 + (size_t)(sg * GS) * g_dst;
          uint8_t* dl = A_lo + r * 16 + (size_t)(sg * GS) * g_dst;
          if (!live) {
            for (int g = 0; g < GS; ++g) {
              *reinterpret_cast<uint4*>(dh + g * g_dst) = make_uint4(0u, 0u, 0u, 0u);
              *reinterpret_cast<uint4*>(dl + g * g_dst) = make_uint4(0u, 0u, 0u, 0u);
            }
            continue;
          }
          const bool lrelu = a.pre == PRE_LRELU;
          const float slope = a.slope;
#pragma unroll 4
          for (int g = 0; g < GS; ++g) {
            float v[E];
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = src[g * g_src + (size_t)e * Wb];
            if (lrelu) {                                               // 0 < slope < 1: leaky_relu(x) = max(x, slope x)
#pragma unroll
              for (int e = 0; e < E; ++e) v[e] = fmaxf(v[e], v[e] * slope);
            }
            if (TF32) {
              float h[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = P::to_tf32(v[e % E]);
              *reinterpret_cast<float4*>(dh + g * g_dst) = make_float4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<float4*>(dl + g * g_dst) =
                  make_float4(v[0] - h[0], v[1 % E] - h[1], v[2 % E] - h[2], v[3 % E] - h[3]);
            } else {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                if (PREC == PREC_F16) P::split2_f16(v[e % E], v[(e + 1) % E], hi[e >> 1], lo[e >> 1]);
                else P::split2_bf16(v[e % E], v[(e + 1) % E], hi[e >> 1], lo[e >> 1]);
              }
              *reinterpret_cast<uint4*>(dh + g * g_dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(dl + g * g_dst) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
        P::fence_async_proxy();
        P::mbar_arrive(cx, &bar.a_full[as]);
        P::mbar_arrive(cx, &bar.raw_empty[rs]);
        if (++as == (uint32_t)a_slots) { as = 0; ++a_round; }
      }
    }
    if (prof) { P::prof_add(a.prof + 6, pw_r); P::prof_add(a.prof + 7, pw_e); P::prof_add(a.prof + 8, P::clock() - p_start); }
  } else {
    // ---------------------------------------------------------------------- epilogue
    // The tensors an epilogue READS from global memory (the residual, the MRF / WaveNet-skip running sums) do not depend
    // on the accumulator, so their loads are issued ahead of their use - the first item's of a tile before the accumulator
    // is even waited for (while the previous tile's last item is still being stored).  Measured on the B200
    // (profiles/r02_ncu_*): with the loads issued after tcgen05.ld, the 8 epilogue warps kept 16 KB in flight per SM and
    // sat in `long scoreboard` for 45 % of all samples of the generator's 32-channel convs, holding the TMEM set (and
    // through it the MMA warp and the converters) - 2.4 TB/s where HBM gives 6.5.
    // Order inside an item (second revision, from the per-instruction samples of the 64-channel ResBlock launches where
    // this role paces the kernel): tcgen05.ld (both accumulators of a pair behind one wait) -> every use of the prefetched values -> the NEXT item's
    // loads into the SAME registers -> stores.  The first revision prefetched into a second register set and copied it
    // over at the end of the item: 32-48 moves per item, 32 more live registers, and the first item of every tile had
    // nothing in flight while it waited.  Addresses are one 64-bit base per tensor and item plus i * (channel stride).
    const int ew = warp - C2_EPI_WARP0;                 // 0..7
    const int q = warp & 3, half = ew >> 2;             // TMEM lane quadrant is fixed by warp id % 4
    uint32_t t_it = 0;
    const int n_chunks = NT / 16;
    const int n_acc = 2 * a.chains;                     // (main | correction) per chain, NT columns apart
    const int epi = a.epi;
    const bool bias_vec = a.bias != nullptr && (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0;
    const bool prof = a.prof != nullptr && ew == 0 && lane == 0;
    long long pw_f = 0;
    const long long p_start = prof ? P::clock() : 0;
    const bool r_all = epi == EPI_RES || epi == EPI_MRF || epi == EPI_SUBFROM;     // residual for every row
    const bool o_all = epi == EPI_MRF && a.mrf != 0;                               // running MRF sum for every row
    const bool any_aux = r_all || o_all || epi == EPI_WN;
    const long long r_cs = a.r.cs, y_cs = a.y.cs, y2_cs = a.y2.cs;
    const int my_chunks = (n_chunks - half + 1) / 2;    // this thread's chunks c = half, half + 2, ... of each live row half
    // kind of auxiliary read of a 16-row chunk: bit 0 = residual rows (r), bit 1 = running-sum rows (y2)
    auto aux_kind = [&](int row0) -> int {
      if (epi == EPI_WN) return row0 + 16 <= a.split ? 1 : (row0 >= a.split && !a.first ? 2 : 0);
      return (r_all ? 1 : 0) | (o_all ? 2 : 0);
    };
    // issue the auxiliary reads of work item wi of tile (nt, b, t0, Lq)
    auto aux_issue = [&](int nt, int b, int t0, int Lq, int wi, float (&rv)[16], float (&ov)[16]) {
      const int mh = wi >= my_chunks ? 1 : 0, c = half + 2 * (wi - mh * my_chunks);   // (work <= 2 * my_chunks)
      const int t = t0 + mh * 128 + q * 32 + lane;
      int item = b;
      if (flat_tg ? !flat_live(t, item) : t >= Lq) return;
      const int row0 = nt * NT + c * 16, kind = aux_kind(row0);
      if (kind & 1) {
        const float* pr = a.r.p + (long long)b * a.r.bs + (long long)row0 * r_cs + t;
#pragma unroll
        for (int i = 0; i < 16; ++i) rv[i] = pr[i * r_cs];
      }
      if (kind & 2) {
        const int o0 = epi == EPI_WN ? row0 - a.split : row0;
        const float* po = a.y2.p + (long long)b * a.y2.bs + (long long)o0 * y2_cs + t;
#pragma unroll
        for (int i = 0; i < 16; ++i) ov[i] = po[i * y2_cs];
      }
    };
    // first live tile at or after `tile` (every role skips the same tiles); false: none left
    auto next_live = [&](int& tile, int& nt, int& b, int& t0, int& L, int& Lq) -> bool {
      for (; tile < tile_hi; tile += tile_step)
        if (decode(tile, nt, b, t0, L, Lq)) return true;
      return false;
    };
    float rv[16], ov[16];
    int tile = tile_lo, nt, b, t0, L, Lq;
    bool have = next_live(tile, nt, b, t0, L, Lq);
    if (have && any_aux && ((Lq - t0 > 128 && MH > 1) ? 2 : 1) * my_chunks > 0) aux_issue(nt, b, t0, Lq, 0, rv, ov);
    while (have) {
      const int mh_live = (Lq - t0 > 128 && MH > 1) ? 2 : 1;
      const int ts = t_it % t_slots;
      float* yb = a.y.p ? a.y.p + (long long)b * a.y.bs : nullptr;
      float* y2b = a.y2.p ? a.y2.p + (long long)b * a.y2.bs : nullptr;
      const float* rb = a.r.p ? a.r.p + (long long)b * a.r.bs : nullptr;
      const int n0 = nt * NT;
      const int work = mh_live * my_chunks;
      // the tile after this one (its first auxiliary reads are issued from inside this tile's last item)
      int tile_n = tile + tile_step, nt_n = 0, b_n = 0, t0_n = 0, L_n = 0, Lq_n = 0;
      const bool have_n = next_live(tile_n, nt_n, b_n, t0_n, L_n, Lq_n);
      const bool pre_n = have_n && any_aux && ((Lq_n - t0_n > 128 && MH > 1) ? 2 : 1) * my_chunks > 0;
      const long long pc = prof ? P::clock() : 0;
      P::mbar_wait(cx, &bar.t_full[ts], (t_it / t_slots) & 1);
      if (prof) pw_f += P::clock() - pc;
      P::fence_tc_after();
      const uint32_t d_set = tmem_d + (uint32_t)(ts * set_cols);
      for (int wi = 0; wi < work; ++wi) {
        const int mh = wi >= my_chunks ? 1 : 0, c = half + 2 * (wi - mh * my_chunks);   // (work <= 2 * my_chunks)
        const int t = t0 + mh * 128 + q * 32 + lane;
        const int row0 = n0 + c * 16;
        int item = b;
        const bool live = flat_tg ? flat_live(t, item) : t < Lq;
        float v[16];
        const uint32_t tbase = d_set + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * a.mh_stride + c * 16);
        {                                                // fp32 round-to-nearest combine of the partial sums: (main | correction)
          float p[16];                                   // pairs, both loads of a pair behind one wait
          P::tmem_ld16x2(cx, tbase, tbase + (uint32_t)NT, v, p);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += p[i];
          if (n_acc == 4) {
            float p2[16];
            P::tmem_ld16x2(cx, tbase + (uint32_t)(2 * NT), tbase + (uint32_t)(3 * NT), p, p2);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += p[i];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += p2[i];
          }
        }
        if (wi == work - 1) {                            // all of this thread's TMEM reads are done: release the set
          P::fence_tc_before();
          P::mbar_arrive(cx, &bar.t_empty[ts]);
        }
        // ---- everything that uses the prefetched values (v becomes what is stored); `dst` picks the destination
        int dst = 0;                                     // 0: y rows row0.., 1: y2 rows row0.., 2: y2 rows row0 - split.., 3: special store
        if (live) {
          if (a.bias) {
            if (bias_vec) {                                  // 16-byte aligned bias vector: four 128-bit loads
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(a.bias + row0 + i);
                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += P::ldg(a.bias + row0 + i);
            }
          }
          if (a.bias_item) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += P::ldg(a.bias_item + (long long)item * a.bias_item_stride + row0 + i);
          }
          if (epi == EPI_MRF) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += rv[i];
            if (a.mrf == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += ov[i];
            } else if (a.mrf != 0) {
              // xs / num_kernels (models.py:363) as a multiplication by the rounded reciprocal: <= 1 ulp from the division,
              // and 16 fp32 divisions per item were 14 % of this kernel's issue slots on the 64-channel stage
              const float inv_n = 1.f / (float)a.mrf_n;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = (ov[i] + v[i]) * inv_n;
            }
            dst = 1;
          } else if (epi == EPI_WN) {
            const int kind = aux_kind(row0);
            if (kind == 1) {                               // residual stream, in place
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = rv[i] + v[i];
            } else if (row0 >= a.split) {                  // skip sum
              if (kind == 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = ov[i] + v[i];
              }
              dst = 2;
            } else {
              dst = 3;                                     // a chunk that straddles the split (no real layer has one)
            }
          } else if (epi == EPI_RES) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += rv[i];
          } else if (epi == EPI_SUBFROM) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = rv[i] - v[i];
          } else if (epi == EPI_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : 0.f;
          }
        }
        // ---- the next item's reads, into the registers that were just consumed: they fly during the stores below and
        // the next tcgen05.ld (the next item may be the first one of the next tile)
        {
          const bool more = wi + 1 < work;               // (one call site: the instruction footprint of this role matters)
          if (more ? any_aux : pre_n) aux_issue(more ? nt : nt_n, more ? b : b_n, more ? t0 : t0_n, more ? Lq : Lq_n, more ? wi + 1 : 0, rv, ov);
        }
        // ---- stores
        if (live) {
          if (epi == EPI_GATE) {
            float* py = yb + (long long)(row0 >> 1) * y_cs + t;
#pragma unroll
            for (int i = 0; i < 16; i += 2) py[(i >> 1) * y_cs] = wn_gate(v[i], v[i + 1]);
          } else if (epi == EPI_UPSAMPLE) {
            if (a.up == 8) store_upsampled<8>(v, yb, a.y.cs, row0, t, a.up_pad, L * 8);
            else if (a.up == 4) store_upsampled<4>(v, yb, a.y.cs, row0, t, a.up_pad, L * 4);
            else if (a.up == 2) store_upsampled<2>(v, yb, a.y.cs, row0, t, a.up_pad, L * 2);
            else {
              for (int i = 0; i < 16; ++i) {               // generic stride
                const int row = row0 + i;
                const int co = row / a.up, phi = row - co * a.up;
                const int to = t * a.up + phi - a.up_pad;
                if (to >= 0 && to < L * a.up) yb[(long long)co * a.y.cs + to] = v[i];
              }
            }
          } else if (dst == 3) {
            for (int i = 0; i < 16; ++i) {
              const int row = row0 + i;
              if (row < a.split) yb[(long long)row * a.y.cs + t] = rb[(long long)row * a.r.cs + t] + v[i];
              else {
                float* o = y2b + (long long)(row - a.split) * a.y2.cs + t;
                *o = a.first ? v[i] : *o + v[i];
              }
            }
          } else if (dst == 0) {
            float* py = yb + (long long)row0 * y_cs + t;
#pragma unroll
            for (int i = 0; i < 16; ++i) py[i * y_cs] = v[i];
          } else {
            float* py = y2b + (long long)(dst == 2 ? row0 - a.split : row0) * y2_cs + t;
#pragma unroll
            for (int i = 0; i < 16; ++i) py[i * y2_cs] = v[i];
          }
        }
      }
      if (work == 0) {                                   // a 16-row tile leaves the odd half of the warps without a chunk:
        P::fence_tc_before();                            // they still owe the accumulator set their arrival
        P::mbar_arrive(cx, &bar.t_empty[ts]);
      }
      ++t_it;
      tile = tile_n; nt = nt_n; b = b_n; t0 = t0_n; L = L_n; Lq = Lq_n; have = have_n;
    }
    if (prof) { P::prof_add(a.prof + 4, pw_f); P::prof_add(a.prof + 5, P::clock() - p_start); }
  }
  P::fence_tc_before();
  P::syncthreads(cx);
  if (warp == 2) P::tmem_dealloc(cx, tmem_d, (uint32_t)a.tmem_cols);
}

}  // namespace conv2
}  // namespace pb200
