// CPU functional model run of the experimental tensor-core attention (piper_b200/csrc/att_body.inl) - test infrastructure.
#include "../../piper_b200/csrc/kernels.cuh"

#include <cmath>

#include "sim_prim.h"

#define MRF_FN inline
#include "../../piper_b200/csrc/att_body.inl"

using namespace pb200;
using namespace simtc;

// qkv [B][3H][cs], out [B][H][cs], rel_k / rel_v [9][dk], len [B]
// tail_thr > 0: the launcher's tail mode - a last query tile of at most tail_thr rows is left to the CUDA-core kernel
extern "C" int att_sim_run2(const float* qkv, float* out, const float* rel_k, const float* rel_v, const int* len, int B, int H,
                            int n_heads, int cs, int Tmax, char* err, int errcap, int tm, int flat, int tail_thr) {
  try {
    att::Args a;
    a.tail_thr = tail_thr;
    // flat: the caller passes [channel][item][cs] tensors
    a.qkv = flat ? View{const_cast<float*>(qkv), (long long)cs, B * cs} : View{const_cast<float*>(qkv), (long long)3 * H * cs, cs};
    a.out = flat ? View{out, (long long)cs, B * cs} : View{out, (long long)H * cs, cs};
    a.tm = tm; a.flat = flat;
    a.rel_k = rel_k; a.rel_v = rel_v; a.len = len;
    a.H = H; a.n_heads = n_heads; a.dk = H / n_heads; a.q_tiles = (Tmax + att::A_QT - 1) / att::A_QT;
    if (a.dk % 16 || a.dk > att::A_MAXDK) throw std::runtime_error("head width not handled");
    const int grid = a.q_tiles * n_heads * B, smem = att::smem_bytes(a.dk);
    TmapDesc tq, tk;
    if (tm) att::att_tmaps(a, B, tq, tk);
    for (int block = 0; block < grid; ++block) {
      std::unique_ptr<SimCta> cta(new SimCta);
      SmemBuf sm(smem);
      cta->smem = sm.p; cta->smem_bytes = smem; cta->control_warps = att::A_CONV_WARP0;
      for (auto& row : cta->tmem) for (float& v : row) v = std::numeric_limits<float>::quiet_NaN();
      std::unique_ptr<att::Barriers<SimMbar>> bar(new att::Barriers<SimMbar>);
      const std::string e = run_cta(*cta, att::A_THREADS, block, grid, [&](SimPrim::Ctx& cx) {
        att::att_body<SimPrim>(a, cx, cta->smem, *bar, &cta->tmem_base, &tq, &tk);
      });
      if (!e.empty()) throw std::runtime_error("CTA " + std::to_string(block) + ": " + e);
    }
    return 0;
  } catch (const std::exception& e) {
    if (err && errcap > 0) std::snprintf(err, size_t(errcap), "%s", e.what());
    return 1;
  }
}

extern "C" int att_sim_run(const float* qkv, float* out, const float* rel_k, const float* rel_v, const int* len, int B, int H,
                           int n_heads, int cs, int Tmax, char* err, int errcap, int tm, int flat) {
  return att_sim_run2(qkv, out, rel_k, rel_v, len, B, H, n_heads, cs, Tmax, err, errcap, tm, flat, 0);
}
