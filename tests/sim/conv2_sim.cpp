// CPU functional model run of the experimental second-generation tensor-core convolution: the SAME kernel body as the
// GPU build (piper_b200/csrc/conv2_body.inl) against the primitive model in sim_prim.h, with the host plan / pack /
// argument fill of conv2_host.h.  Test infrastructure only (tests/test_conv2_sim.py).
#include "../../piper_b200/csrc/conv2_host.h"

#include <cmath>

#include "sim_prim.h"

#define MRF_FN inline
#define MRF_NOINLINE inline
#include "../../piper_b200/csrc/conv2_body.inl"

using namespace pb200;
using namespace simtc;

// info = {ok, n_tile, n_tiles, mt, kc, stage_rows, raw_stride, t_slots, tmem_cols, chains, mh_stride, smem, w_bytes}
extern "C" void conv2_sim_plan(int ci, int rows, int k, int dil, int prec, int chains, long long* info) {
  conv2::Plan p;
  conv2::plan(ci, rows, k, dil, prec, chains, p);
  const long long v[13] = {p.ok, p.n_tile, p.n_tiles, p.mt, p.kc, p.stage_rows, p.raw_stride, p.t_slots, p.tmem_cols, p.chains,
                           p.mh_stride, (long long)p.smem, (long long)p.w_bytes};
  for (int i = 0; i < 13; ++i) info[i] = v[i];
}

// One convolution layer on a ragged batch.  x [B][ci][cs_x], w [rows][ci][k] fp32 (rows = output rows: C_out, or
// C_out * up for a lowered ConvTranspose), y / y2 / r as the epilogue needs.  desc = {ci, rows, k, dil, pad, q_extra, pre,
// epi, split, first, up, up_pad, mrf, mrf_n, prec (0 bf16x3 / 1 tf32x3 / 2 fp16x3), len_scale, cs_x, cs_y, cs_y2, cs_r, C_y, C_y2, C_r, grid, chains (0 = default), plan opts}.
// info (out) = {n_tile, n_tiles, mt, kc, t_slots, chains, total_tiles, smem bytes}.
extern "C" int conv2_sim_run(const float* x, const float* w, const float* bias, const float* bias_item, int bias_item_stride,
                             float* y, float* y2, const float* r, const int* len, int B, const int* desc, float slope, int max_len,
                             int* info, char* err, int errcap) {
  try {
    const int ci = desc[0], rows = desc[1], k = desc[2], dil = desc[3];
    const int prec = desc[14], chains = desc[24] > 0 ? desc[24] : (prec == 1 ? 2 : 1);
    const bool tf32 = prec == 1;
    conv2::Plan p;
    if (!conv2::plan(ci, rows, k, dil, prec, chains, p, desc[25])) throw std::runtime_error("shape outside the kernel's plan");
    // engine layout of the weights: [ci][k][rows_p], row fastest
    const int rows_p = rows;
    std::vector<float> wsrc(size_t(ci) * k * rows_p);
    for (int co = 0; co < rows; ++co)
      for (int c = 0; c < ci; ++c)
        for (int j = 0; j < k; ++j) wsrc[(size_t(c) * k + j) * rows_p + co] = w[(size_t(co) * ci + c) * k + j];
    std::vector<uint8_t> packed_raw(p.w_bytes + 128);
    uint8_t* packed = packed_raw.data() + ((128 - reinterpret_cast<uintptr_t>(packed_raw.data()) % 128) % 128);
    conv2::pack(wsrc.data(), ci, k, rows_p, p, packed);

    MmaConvArgs a;
    a.x = View{const_cast<float*>(x), (long long)ci * desc[16], desc[16]};
    a.y = View{y, (long long)desc[20] * desc[17], desc[17]};
    a.y2 = View{y2, (long long)desc[21] * desc[18], desc[18]};
    a.r = View{const_cast<float*>(r), (long long)desc[22] * desc[19], desc[19]};
    a.w = packed; a.bias = bias; a.bias_item = bias_item; a.bias_item_stride = bias_item_stride;
    a.len = len; a.len_scale = desc[15];
    a.ci = ci; a.rows = rows; a.k = k; a.dil = dil; a.pad = desc[4]; a.q_extra = desc[5];
    a.pre = desc[6]; a.slope = slope; a.epi = desc[7]; a.split = desc[8]; a.first = desc[9];
    a.up = desc[10]; a.up_pad = desc[11]; a.mrf = desc[12]; a.mrf_n = desc[13];
    const bool flat = (desc[25] & 16) != 0;              // opts bit 4: flat mode - the caller passes [C][B][Tg] tensors (cs = B * Tg)
    if (flat) {
      a.flat_tg = desc[16] / B; a.flat_n = B;
      a.x.bs = a.flat_tg; a.y.bs = desc[17] / B; a.y2.bs = desc[18] / B; a.r.bs = desc[19] / B;
    }
    const int launch_B = flat ? 1 : B;
    if (flat) max_len = B * a.flat_tg;
    const bool tm = (desc[25] & 4) != 0;                 // opts bit 2: tensor-map TMA for the activation window
    a.mma3 = (desc[25] & 8) ? 1 : 0;                     // opts bit 3: three instructions per k-step on matching accumulator regions
    TmapDesc td;
    size_t smem_bytes = p.smem;
    int grid = conv2::fill_args(a, p, launch_B, max_len, tm, &td, (desc[25] & 32) != 0, &smem_bytes);   // opts bit 5: A-stationary order
    if (!tm && a.astat) grid = conv2::fill_args(a, p, launch_B, max_len, false, nullptr, false, &smem_bytes);   // as the launcher: A-stationary needs the tensor-map kernel
    if (desc[23] > 0) grid = std::min(grid, desc[23]);
    if (info) {
      info[0] = p.n_tile; info[1] = p.n_tiles; info[2] = p.mt; info[3] = p.kc; info[4] = p.t_slots; info[5] = a.chains;
      info[6] = a.total_tiles; info[7] = (int)smem_bytes;
    }
    for (int block = 0; block < grid; ++block) {
      std::unique_ptr<SimCta> cta(new SimCta);
      SmemBuf smem(smem_bytes);
      cta->smem = smem.p; cta->smem_bytes = (int)smem_bytes; cta->control_warps = conv2::C2_CONV_WARP0;
      for (auto& row : cta->tmem) for (float& v : row) v = std::numeric_limits<float>::quiet_NaN();
      std::unique_ptr<conv2::Barriers<SimMbar>> bar(new conv2::Barriers<SimMbar>);
      const std::string e = run_cta(*cta, conv2::C2_THREADS, block, grid, [&](SimPrim::Ctx& cx) {
        if (tm && a.astat) {                             // the A-stationary instantiation: operand ring sized at run time
          if (tf32) conv2::conv2_body<SimPrim, 1, 128, true, 0>(a, cx, cta->smem, *bar, &cta->tmem_base, &td);
          else if (prec == 2 && p.mt == 256) conv2::conv2_body<SimPrim, 2, 256, true, 0>(a, cx, cta->smem, *bar, &cta->tmem_base, &td);
          else if (prec == 2) conv2::conv2_body<SimPrim, 2, 128, true, 0>(a, cx, cta->smem, *bar, &cta->tmem_base, &td);
          else if (p.mt == 256) conv2::conv2_body<SimPrim, 0, 256, true, 0>(a, cx, cta->smem, *bar, &cta->tmem_base, &td);
          else conv2::conv2_body<SimPrim, 0, 128, true, 0>(a, cx, cta->smem, *bar, &cta->tmem_base, &td);
        } else if (tm) {
          if (tf32) conv2::conv2_body<SimPrim, 1, 128, true>(a, cx, cta->smem, *bar, &cta->tmem_base, &td);
          else if (prec == 2 && p.mt == 256) conv2::conv2_body<SimPrim, 2, 256, true>(a, cx, cta->smem, *bar, &cta->tmem_base, &td);
          else if (prec == 2) conv2::conv2_body<SimPrim, 2, 128, true>(a, cx, cta->smem, *bar, &cta->tmem_base, &td);
          else if (p.mt == 256) conv2::conv2_body<SimPrim, 0, 256, true>(a, cx, cta->smem, *bar, &cta->tmem_base, &td);
          else conv2::conv2_body<SimPrim, 0, 128, true>(a, cx, cta->smem, *bar, &cta->tmem_base, &td);
        } else if (tf32) conv2::conv2_body<SimPrim, 1, 128>(a, cx, cta->smem, *bar, &cta->tmem_base);
        else if (prec == 2 && p.mt == 256) conv2::conv2_body<SimPrim, 2, 256>(a, cx, cta->smem, *bar, &cta->tmem_base);
        else if (prec == 2) conv2::conv2_body<SimPrim, 2, 128>(a, cx, cta->smem, *bar, &cta->tmem_base);
        else if (p.mt == 256) conv2::conv2_body<SimPrim, 0, 256>(a, cx, cta->smem, *bar, &cta->tmem_base);
        else conv2::conv2_body<SimPrim, 0, 128>(a, cx, cta->smem, *bar, &cta->tmem_base);
      });
      if (!e.empty()) throw std::runtime_error("CTA " + std::to_string(block) + ": " + e);
    }
    return 0;
  } catch (const std::exception& e) {
    if (err && errcap > 0) std::snprintf(err, size_t(errcap), "%s", e.what());
    return 1;
  }
}
