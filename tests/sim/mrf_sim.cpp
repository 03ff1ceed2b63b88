// CPU functional model of the experimental fused MRF stage kernel: compiles the SAME kernel body as the GPU build
// (piper_b200/csrc/mrf_fused_body.inl) against SimPrim below and runs every thread of a CTA as a std::thread.
//
//   * shared memory  = a byte array; "shared addresses" are offsets into it, exactly what the descriptors encode
//   * TMEM           = float[128 lanes][512 columns]; tcgen05.ld / st check the warp's lane-quadrant rule
//   * tcgen05.mma    = decoded from the shared-memory descriptor lo words (start, LBO; SBO = 128, no swizzle, K-major)
//                      and the instruction descriptor (N), executed synchronously; tcgen05.commit = immediate arrive
//   * cp.async.bulk  = memcpy + complete_tx, with the TMA alignment rules asserted (16-byte addresses and sizes)
//   * mbarrier       = blocking phase barrier with arrival and transaction counts; a wait that lasts 20 s aborts the run
//
// What it can show: index arithmetic, operand layouts, TMEM column maps and the barrier protocol (phases, counts,
// deadlocks) are consistent and reproduce the oracle.  What it cannot: asynchrony of the real MMA / TMA pipelines,
// memory-model fences, PTX encodings.  Test infrastructure only (tests/test_fused_mrf_sim.py).
#include "../../piper_b200/csrc/kernels.cuh"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <limits>
#include <memory>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#define MRF_FN inline
#include "../../piper_b200/csrc/mrf_fused_body.inl"

namespace {
using namespace pb200;
using namespace pb200::mrf;

struct SimAbort : std::runtime_error { using std::runtime_error::runtime_error; };

struct SimCta;
struct SimMbar {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, pending = 0;
  long long tx = 0;
  uint32_t phase = 0;
  const char* name = "?";
};

struct CtaBarrier {   // __syncthreads for n threads, reusable
  std::mutex m; std::condition_variable cv; int n = 0, waiting = 0; unsigned gen = 0;
  void arrive_and_wait(std::atomic<bool>& abort) {
    std::unique_lock<std::mutex> l(m);
    const unsigned g = gen;
    if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); return; }
    while (gen == g) {
      if (cv.wait_for(l, std::chrono::seconds(20)) == std::cv_status::timeout || abort.load())
        if (gen == g) { abort = true; throw SimAbort("__syncthreads never completed"); }
    }
  }
};

struct SimCta {
  alignas(128) uint8_t smem[F_SMEM];
  float tmem[128][512];
  FBarriers<SimMbar> bar;
  uint32_t tmem_base = 0xdeadbeef;
  CtaBarrier sync;
  std::atomic<bool> abort{false};
  std::mutex err_m;
  std::string err;
  void fail(const std::string& e) {
    std::lock_guard<std::mutex> l(err_m);
    if (err.empty()) err = e;
    abort = true;
  }
};

inline uint16_t f2bf(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x40);
  return uint16_t((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
inline float bf2f(uint16_t h) { uint32_t u = uint32_t(h) << 16; float f; memcpy(&f, &u, 4); return f; }

struct SimPrim {
  using Mbar = SimMbar;
  struct Ctx {
    int tid_, block_, grid_;
    SimCta* cta;
    int tid() const { return tid_; }
    int block() const { return block_; }
    int grid() const { return grid_; }
  };
  static void check(Ctx& c, bool ok, const char* what) {
    if (!ok) { c.cta->fail(std::string("check failed: ") + what + " (tid " + std::to_string(c.tid_) + ")"); throw SimAbort(what); }
  }
  static uint32_t saddr(Ctx& c, const void* p) {
    const long long off = (const uint8_t*)p - c.cta->smem;
    check(c, off >= 0 && off < F_SMEM, "pointer outside shared memory");
    return (uint32_t)off;
  }
  static int bcast0(Ctx&, int v) { return v; }           // values broadcast in the kernel are warp-uniform by construction
  static void syncwarp() {}
  static void syncthreads(Ctx& c) { c.cta->sync.arrive_and_wait(c.cta->abort); }
  static bool elect_one(Ctx& c) { return (c.tid_ & 31) == 0; }
  static void fence_mbar_init() {}
  static void fence_async_proxy() {}
  static void fence_tc_before() {}
  static void fence_tc_after() {}

  static void mbar_init(Ctx&, Mbar* m, uint32_t count) { m->count = m->pending = (int)count; m->tx = 0; m->phase = 0; }
  static void complete_locked(Mbar* m) {
    if (m->pending == 0 && m->tx == 0) { m->phase ^= 1u; m->pending = m->count; m->cv.notify_all(); }
  }
  static void mbar_arrive(Ctx& c, Mbar* m) {
    std::lock_guard<std::mutex> l(m->m);
    check(c, m->pending > 0, "more arrivals than the barrier expects in one phase");
    --m->pending;
    complete_locked(m);
  }
  static void mbar_expect_tx(Ctx& c, Mbar* m, uint32_t bytes) {
    std::lock_guard<std::mutex> l(m->m);
    check(c, m->pending > 0, "more arrivals than the barrier expects in one phase");
    m->tx += bytes;
    --m->pending;
    complete_locked(m);
  }
  static void complete_tx(Mbar* m, uint32_t bytes) {
    std::lock_guard<std::mutex> l(m->m);
    m->tx -= bytes;
    complete_locked(m);
  }
  static void mbar_wait(Ctx& c, Mbar* m, uint32_t parity) {
    // Control warps (TMA x 2, MMA) run converged on hardware: all 32 lanes observe a barrier phase together and only the
    // elected lane acts.  std::threads are not in lockstep - a lane that reaches a parity wait late could find the phase
    // flipped twice and wait forever - so in the model only the elected lane of a control warp blocks.
    if ((c.tid_ >> 5) < F_CONV_WARP0 && (c.tid_ & 31) != 0) return;
    std::unique_lock<std::mutex> l(m->m);
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(20);
    while (m->phase == parity) {                          // the phase with this parity has not completed yet
      if (c.cta->abort.load()) throw SimAbort("aborted");
      if (m->cv.wait_until(l, std::chrono::steady_clock::now() + std::chrono::milliseconds(200)) == std::cv_status::timeout &&
          std::chrono::steady_clock::now() > deadline) {
        l.unlock();
        c.cta->fail("deadlock: thread " + std::to_string(c.tid_) + " (warp " + std::to_string(c.tid_ >> 5) +
                    ") waited 20 s on an mbarrier, parity " + std::to_string(parity));
        throw SimAbort("deadlock");
      }
    }
  }
  static void bulk_g2s(Ctx& c, uint32_t dst, const void* src, uint32_t bytes, Mbar* m) {
    check(c, (dst & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (bytes & 15) == 0 && bytes > 0,
          "cp.async.bulk needs 16-byte aligned addresses and size");
    check(c, dst + bytes <= (uint32_t)F_SMEM, "cp.async.bulk writes past shared memory");
    memcpy(c.cta->smem + dst, src, bytes);
    complete_tx(m, bytes);
  }
  static void tmem_alloc(Ctx& c, uint32_t* slot, uint32_t cols) { check(c, cols == 512, "tmem columns"); *slot = 0; }
  static void tmem_dealloc(Ctx&, uint32_t, uint32_t) {}
  static void mma_bf16(Ctx& c, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    const uint32_t a0 = (a_lo & 0x3FFFu) << 4, albo = ((a_lo >> 16) & 0x3FFFu) << 4;
    const uint32_t b0 = (b_lo & 0x3FFFu) << 4, blbo = ((b_lo >> 16) & 0x3FFFu) << 4;
    const int N = int((idesc >> 17) & 0x3Fu) << 3, M = int((idesc >> 24) & 0x1Fu) << 4;
    check(c, M == 128 && (N == 32 || N == 64), "instruction descriptor shape");
    check(c, ((idesc >> 4) & 3u) == 1 && ((idesc >> 7) & 7u) == 1 && ((idesc >> 10) & 7u) == 1, "instruction descriptor formats");
    const uint32_t col0 = tmem_d & 0xFFFFu;
    check(c, (tmem_d >> 16) == 0 && col0 + N <= 512, "accumulator address");
    auto elem = [&](uint32_t start, uint32_t lbo, int row, int k) -> float {
      const uint32_t addr = start + uint32_t(k / 8) * lbo + uint32_t(row / 8) * 128u + uint32_t(row % 8) * 16u + uint32_t(k % 8) * 2u;
      if (addr + 2 > (uint32_t)F_SMEM) { c.cta->fail("tcgen05.mma operand outside shared memory"); throw SimAbort("operand"); }
      uint16_t h; memcpy(&h, c.cta->smem + addr, 2);
      return bf2f(h);
    };
    float B[64][16];
    for (int n = 0; n < N; ++n) for (int k = 0; k < 16; ++k) B[n][k] = elem(b0, blbo, n, k);
    for (int m = 0; m < 128; ++m) {
      float A[16];
      for (int k = 0; k < 16; ++k) A[k] = elem(a0, albo, m, k);
      for (int n = 0; n < N; ++n) {
        float s = 0.f;
        for (int k = 0; k < 16; ++k) s += A[k] * B[n][k];
        float& d = c.cta->tmem[m][col0 + n];
        d = acc ? d + s : s;
      }
    }
  }
  static void mma_commit(Ctx& c, Mbar* m) { mbar_arrive(c, m); }   // the model executes MMAs synchronously
  static void tmem_ld16(Ctx& c, uint32_t taddr, float (&v)[16]) {
    const uint32_t lane0 = taddr >> 16, col = taddr & 0xFFFFu;
    check(c, lane0 == 32u * ((c.tid_ >> 5) & 3), "tcgen05.ld outside the warp's TMEM lane quadrant (warp id % 4)");
    check(c, col + 16 <= 512, "tcgen05.ld column range");
    for (int i = 0; i < 16; ++i) v[i] = c.cta->tmem[lane0 + (c.tid_ & 31)][col + i];
  }
  static void tmem_st16(Ctx& c, uint32_t taddr, const float* v) {
    const uint32_t lane0 = taddr >> 16, col = taddr & 0xFFFFu;
    check(c, lane0 == 32u * ((c.tid_ >> 5) & 3), "tcgen05.st outside the warp's TMEM lane quadrant (warp id % 4)");
    check(c, col + 16 <= 512, "tcgen05.st column range");
    for (int i = 0; i < 16; ++i) c.cta->tmem[lane0 + (c.tid_ & 31)][col + i] = v[i];
  }
  static void tmem_wait_st() {}
  static float bf16_round(float v) { return bf2f(f2bf(v)); }
  static uint32_t pack_bf16(float a, float b) { return uint32_t(f2bf(a)) | (uint32_t(f2bf(b)) << 16); }
};

}  // namespace

// plan = the 27 ints pb200_debug_mrf_pack returns.  x / y: [B][32][cs] fp32 (bs = batch stride in floats).
extern "C" int mrf_sim_run(const float* x, float* y, const int* len, int B, long long bs, int cs, int len_scale,
                           const uint8_t* w, const float* bias, const int* plan, int max_len, int grid, char* err, int errcap) {
  try {
    MrfFusedPlan p;
    p.ok = plan[0] != 0; p.n_chains = plan[1]; p.n_steps = plan[2]; p.pair = plan[3]; p.hv = plan[4]; p.to = plan[5];
    for (int c = 0; c < MRF_MAX_CHAINS; ++c) {
      p.k[c] = plan[6 + c];
      for (int s = 0; s < MRF_MAX_STEPS; ++s) p.dil[c][s] = plan[9 + c * MRF_MAX_STEPS + s];
    }
    if (!p.ok) throw std::runtime_error("plan not ok");
    MrfFusedArgs a;
    a.x = View{const_cast<float*>(x), bs, cs};
    a.y = View{y, bs, cs};
    a.len = len; a.len_scale = len_scale; a.w = w; a.bias = bias; a.slope = 0.1f;
    mrf_fill_args(a, p, B, max_len);
    if (grid <= 0 || grid > a.total_tiles) grid = a.total_tiles;
    std::string first_err;
    for (int block = 0; block < grid; ++block) {
      std::unique_ptr<SimCta> cta(new SimCta);
      memset(cta->smem, 0xCD, sizeof cta->smem);          // stale garbage (NaN-ish patterns), as on hardware
      for (auto& row : cta->tmem) for (float& v : row) v = std::numeric_limits<float>::quiet_NaN();
      cta->sync.n = F_THREADS;
      std::vector<std::thread> th;
      for (int t = 0; t < F_THREADS; ++t)
        th.emplace_back([&, t] {
          SimPrim::Ctx cx{t, block, grid, cta.get()};
          try {
            mrf_fused_body<SimPrim>(a, cx, cta->smem, cta->bar, &cta->tmem_base);
          } catch (const SimAbort&) {
          } catch (const std::exception& e) {
            cta->fail(e.what());
          }
        });
      for (auto& t : th) t.join();
      if (!cta->err.empty()) { first_err = "CTA " + std::to_string(block) + ": " + cta->err; break; }
    }
    if (!first_err.empty()) throw std::runtime_error(first_err);
    return 0;
  } catch (const std::exception& e) {
    if (err && errcap > 0) { std::snprintf(err, size_t(errcap), "%s", e.what()); }
    return 1;
  }
}
