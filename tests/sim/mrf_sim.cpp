// CPU functional model of the experimental fused MRF stage kernel: compiles the SAME kernel body as the GPU build
// (piper_b200/csrc/mrf_fused_body.inl) against SimPrim below and runs every thread of a CTA as a std::thread.
//
//   * shared memory  = a byte array; "shared addresses" are offsets into it, exactly what the descriptors encode
//   * TMEM           = float[128 lanes][512 columns]; tcgen05.ld / st check the warp's lane-quadrant rule
//   * tcgen05.mma    = decoded from the shared-memory descriptor lo words (start, LBO; SBO = 128, no swizzle, K-major)
//                      and the instruction descriptor (N); queued and executed LATER, in order, by a "tensor core" thread
//                      with random pauses; tcgen05.commit arrives when everything queued before it has run
//   * cp.async.bulk  = queued to a "TMA" thread: memcpy + complete_tx some time later (alignment rules asserted)
//   * mbarrier       = blocking phase barrier with arrival and transaction counts; a wait that lasts 60 s aborts the run
//
// What it can show: index arithmetic, operand layouts, TMEM column maps and the barrier protocol - liveness (phases,
// counts, deadlocks) and sufficiency (a result read before its barrier is stale, an operand overwritten before the engine
// consumed it is seen overwritten) - reproduce the oracle.  What it cannot: memory-model fences, PTX encodings, speed.  Test infrastructure only (tests/test_fused_mrf_sim.py).
#include "../../piper_b200/csrc/kernels.cuh"

#include <cmath>

#include "sim_prim.h"

#define MRF_FN inline
#include "../../piper_b200/csrc/mrf_fused_body.inl"

using namespace pb200;
using namespace pb200::mrf;
using namespace simtc;

// plan = the 28 ints pb200_debug_mrf_pack returns.  x / y: [B][32][cs] fp32 (bs = batch stride in floats).
// post_w != NULL (with a plan made for it): conv_post + tanh fused, audio[out_off[b] + t] written instead of y.
extern "C" int mrf_sim_run(const float* x, float* y, const int* len, int B, long long bs, int cs, int len_scale,
                           const uint8_t* w, const float* bias, const int* plan, int max_len, int grid, const float* post_w,
                           float* audio, const long long* out_off, char* err, int errcap) {
  try {
    MrfFusedPlan p;
    p.ok = plan[0] != 0; p.n_chains = plan[1]; p.n_steps = plan[2]; p.pair = plan[3]; p.hv = plan[4]; p.to = plan[5];
    for (int c = 0; c < MRF_MAX_CHAINS; ++c) {
      p.k[c] = plan[6 + c];
      for (int s = 0; s < MRF_MAX_STEPS; ++s) p.dil[c][s] = plan[9 + c * MRF_MAX_STEPS + s];
    }
    p.post_k = plan[27];
    if (!p.ok) throw std::runtime_error("plan not ok");
    if ((p.post_k > 0) != (post_w != nullptr)) throw std::runtime_error("plan and post weights disagree");
    MrfFusedArgs a;
    a.x = View{const_cast<float*>(x), bs, cs};
    a.y = View{y, bs, cs};
    a.len = len; a.len_scale = len_scale; a.w = w; a.bias = bias; a.slope = 0.1f;
    a.post_w = post_w; a.post_slope = 0.01f; a.audio = audio; a.out_off = out_off;
    mrf_fill_args(a, p, B, max_len);
    if (grid <= 0 || grid > a.total_tiles) grid = a.total_tiles;
    std::string first_err;
    for (int block = 0; block < grid; ++block) {
      std::unique_ptr<SimCta> cta(new SimCta);
      SmemBuf smem(F_SMEM);                               // stale garbage (NaN-ish patterns), as on hardware
      cta->smem = smem.p; cta->smem_bytes = F_SMEM; cta->control_warps = F_CONV_WARP0;
      for (auto& row : cta->tmem) for (float& v : row) v = std::numeric_limits<float>::quiet_NaN();
      std::unique_ptr<FBarriers<SimMbar>> bar(new FBarriers<SimMbar>);
      const std::string e = run_cta(*cta, F_THREADS, block, grid, [&](SimPrim::Ctx& cx) {
        mrf_fused_body<SimPrim>(a, cx, cta->smem, *bar, &cta->tmem_base);
      });
      if (!e.empty()) { first_err = "CTA " + std::to_string(block) + ": " + e; break; }
    }
    if (!first_err.empty()) throw std::runtime_error(first_err);
    return 0;
  } catch (const std::exception& e) {
    if (err && errcap > 0) { std::snprintf(err, size_t(errcap), "%s", e.what()); }
    return 1;
  }
}
